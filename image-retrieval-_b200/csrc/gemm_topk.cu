// tcgen05 / TMEM / TMA path (bring-up stub: path disabled until the kernel lands).
#include "gemm_topk.h"

namespace b200ir {

bool gemm_path_supported(int, int, int64_t, int64_t, int, int, int) { return false; }
size_t gemm_workspace_bytes(int, int64_t, int64_t, int, int, int) { return 0; }
int run_gemm_topk(int, const void*, int64_t, const void*, int64_t, int, int, int64_t, int, const MetricParams&, float*,
                  int64_t*, unsigned char*, cudaStream_t) { return B200IR_E_SHAPE; }

}  // namespace b200ir
