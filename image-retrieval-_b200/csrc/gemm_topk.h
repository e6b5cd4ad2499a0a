// tcgen05 / TMEM / TMA path for L2 and cosine-family top-k (bf16 stores directly, fp32 stores through the
// three-term bf16 split): host-side interface.
#pragma once
#include "common.cuh"

namespace b200ir {

bool gemm_path_supported(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags);
// prepared per-store state (row scales, max norm, bf16 hi / lo planes of an fp32 store); 0 = shape not served by this path
size_t gemm_index_bytes(int dtype, int64_t N, int D);
int gemm_index_build(int dtype, const void* X, int64_t N, int D, unsigned char* index, cudaStream_t st);
size_t gemm_workspace_bytes(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags, bool have_index);
// byte offset inside the workspace of the int32 count of queries the last search re-did with the exact scan
size_t gemm_fallback_counter_offset(int dtype, int64_t nq, int64_t N, int D, int k, int flags, bool have_index);
// index == nullptr: the per-store state is rebuilt inside the workspace on every call
int run_gemm_topk(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k, int64_t index_offset,
                  int flags, const MetricParams& mp, float* out_score, int64_t* out_idx, unsigned char* ws,
                  const unsigned char* index, cudaStream_t st);

}  // namespace b200ir
