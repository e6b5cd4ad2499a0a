// One (kind, dtype) instantiation family of the scan kernel per object file (parallel builds).
#include "scan_topk.cuh"

#ifndef SCAN_KIND
#error "compile with -DSCAN_KIND=K_L1|K_L2|K_LINF|K_DOT|K_MULTI -DSCAN_BF16=0|1"
#endif
#define CAT3(a, b, c) a##b##c
#define FN(kind, suffix) CAT3(launch_scan_, kind, suffix)

namespace b200ir {
#if defined(SCAN_EVAL)
cudaError_t launch_scan_eval_f32(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, size_t smem, cudaStream_t st) { return launch_scan_eval_inst(tmX, tmQ, a, smem, st); }
#elif SCAN_BF16
cudaError_t FN(SCAN_KIND, _bf16)(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, int TQ, size_t smem, cudaStream_t st) {
  return launch_scan_tq<SCAN_KIND, __nv_bfloat16>(tmX, tmQ, a, TQ, smem, st);
}
#else
cudaError_t FN(SCAN_KIND, _f32)(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, int TQ, size_t smem, cudaStream_t st) {
  return launch_scan_tq<SCAN_KIND, float>(tmX, tmQ, a, TQ, smem, st);
}
#endif
}  // namespace b200ir
