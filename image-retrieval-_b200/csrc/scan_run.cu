// Scan driver: query preparation + dispatch to the per-(kind, dtype) kernel families.
#include "scan_topk.cuh"
#include "profile.h"

namespace b200ir {

cudaError_t launch_prep_queries(int dtype, const void* Q, int nq, int D, int nq_pad, int D_pad, float* Qf, float* qn,
                                cudaStream_t st) {
  const int warps_per_block = 8;
  const int blocks = int(ceil_div64(nq_pad, warps_per_block));
  ProfileScope ps(PT_PREP, st);
  if (dtype == B200IR_F32)
    prep_queries_kernel<float><<<blocks, warps_per_block * 32, 0, st>>>(static_cast<const float*>(Q), nq, D, nq_pad, D_pad, Qf, qn);
  else
    prep_queries_kernel<__nv_bfloat16><<<blocks, warps_per_block * 32, 0, st>>>(static_cast<const __nv_bfloat16*>(Q), nq, D, nq_pad, D_pad, Qf, qn);
  return cudaGetLastError();
}

cudaError_t run_scan(const ScanPlan& pl, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k,
                     const MetricParams& mp, unsigned char* ws, float* out_all, cudaStream_t st) {
  float* Qf = reinterpret_cast<float*>(ws + pl.off_qf);
  float* qn = reinterpret_cast<float*>(ws + pl.off_qn);
  cudaError_t e = launch_prep_queries(dtype, Q, int(nq), D, pl.nq_pad, pl.D_pad, Qf, qn, st);
  if (e != cudaSuccess) return e;

  ScanArgs a{};
  a.X = X; a.N = N; a.D = D; a.Qf = Qf; a.qnorm = qn; a.nq = int(nq); a.D_pad = pl.D_pad;
  a.G = pl.G; a.P = pl.P; a.rows_per_part = pl.rows_per_part; a.k = k; a.sortn = pl.sortn;
  const int esz = dtype == B200IR_F32 ? 4 : 2;
  a.aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (int64_t(D) * esz) % 16 == 0) ? 1 : 0;
  a.partial = reinterpret_cast<uint64_t*>(ws + pl.off_partial);
  a.out_all = out_all;
  a.mp = mp;
  const bool f32 = dtype == B200IR_F32;
  ProfileScope ps(PT_SCAN, st);
  switch (scan_kind_of(mp.metric)) {
    case K_L1:    return f32 ? launch_scan_K_L1_f32(a, pl.TQ, pl.smem, st)    : launch_scan_K_L1_bf16(a, pl.TQ, pl.smem, st);
    case K_L2:    return f32 ? launch_scan_K_L2_f32(a, pl.TQ, pl.smem, st)    : launch_scan_K_L2_bf16(a, pl.TQ, pl.smem, st);
    case K_LINF:  return f32 ? launch_scan_K_LINF_f32(a, pl.TQ, pl.smem, st)  : launch_scan_K_LINF_bf16(a, pl.TQ, pl.smem, st);
    case K_DOT:   return f32 ? launch_scan_K_DOT_f32(a, pl.TQ, pl.smem, st)   : launch_scan_K_DOT_bf16(a, pl.TQ, pl.smem, st);
    case K_MULTI: return f32 ? launch_scan_K_MULTI_f32(a, pl.TQ, pl.smem, st) : launch_scan_K_MULTI_bf16(a, pl.TQ, pl.smem, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace b200ir
