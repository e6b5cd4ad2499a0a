// tcgen05 / TMEM / TMA path for bf16 L2 and cosine-family top-k: host-side interface.
#pragma once
#include "common.cuh"

namespace b200ir {

bool gemm_path_supported(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags);
size_t gemm_workspace_bytes(int metric, int64_t nq, int64_t N, int D, int k, int flags);
int run_gemm_topk(int metric, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k, int64_t index_offset,
                  int flags, const MetricParams& mp, float* out_score, int64_t* out_idx, unsigned char* ws,
                  cudaStream_t st);

}  // namespace b200ir
