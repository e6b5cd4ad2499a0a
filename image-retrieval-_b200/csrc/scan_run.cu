// Scan driver: query preparation + dispatch to the per-(kind, dtype) kernel families.
#include "scan_topk.cuh"
#include "select.cuh"
#include "profile.h"

#include <cstring>

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

// Tensor maps of the database shard (128-row x 128-byte boxes, 128-byte swizzle) and of the prepared fp32 queries
// (TQ-row boxes).  Falls back to the cp.async loader (use_tma = 0) when rows are not 16-byte aligned or the driver
// entry point is unavailable.
static void setup_tma(ScanArgs& a, int dtype, int TQ, int nq_pad, size_t smem, CUtensorMap* tmX, CUtensorMap* tmQ, int TR = 1) {
  memset(tmX, 0, sizeof(*tmX));
  memset(tmQ, 0, sizeof(*tmQ));
  a.bar_off = int(smem) - 64;
  a.use_tma = 0;
  if (!a.aligned || a.N <= 0) return;
  const int esz = dtype == B200IR_F32 ? 4 : 2;
  const int DKE = kRowChunkBytes / esz;
  if (!tma::encode_2d(tmX, esz, dtype == B200IR_BF16, a.X, a.N, a.D, int64_t(a.D) * esz, DKE, kScanThreads * TR, true)) return;
  if (!tma::encode_2d(tmQ, 4, false, a.Qf, nq_pad, a.D_pad, int64_t(a.D_pad) * 4, DKE, TQ, false)) return;
  a.use_tma = 1;
}

static cudaError_t dispatch_scan(int kind, bool f32, const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, int TQ, size_t smem,
                                 cudaStream_t st) {
  ProfileScope ps(PT_SCAN, st);
  switch (kind) {
    case K_L1:    return f32 ? launch_scan_K_L1_f32(tmX, tmQ, a, TQ, smem, st)    : launch_scan_K_L1_bf16(tmX, tmQ, a, TQ, smem, st);
    case K_L2:    return f32 ? launch_scan_K_L2_f32(tmX, tmQ, a, TQ, smem, st)    : launch_scan_K_L2_bf16(tmX, tmQ, a, TQ, smem, st);
    case K_LINF:  return f32 ? launch_scan_K_LINF_f32(tmX, tmQ, a, TQ, smem, st)  : launch_scan_K_LINF_bf16(tmX, tmQ, a, TQ, smem, st);
    case K_DOT:   return f32 ? launch_scan_K_DOT_f32(tmX, tmQ, a, TQ, smem, st)   : launch_scan_K_DOT_bf16(tmX, tmQ, a, TQ, smem, st);
    case K_MULTI: return f32 ? launch_scan_K_MULTI_f32(tmX, tmQ, a, TQ, smem, st) : launch_scan_K_MULTI_bf16(tmX, tmQ, a, TQ, smem, st);
    case K_MULTI6: return f32 ? launch_scan_K_MULTI6_f32(tmX, tmQ, a, TQ, smem, st) : launch_scan_K_MULTI6_bf16(tmX, tmQ, a, TQ, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t run_scan(const ScanPlan& pl, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k,
                     const MetricParams& mp, unsigned char* ws, float* out_all, cudaStream_t st, const uint64_t* after, int kind_mask) {
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
  a.after = after;
  a.nl = 1;
  int kind = scan_kind_of(mp.metric);
  if (kind_mask != 0) {
    kind = K_MULTI6;
    a.nl = 0;
    for (int m = 0; m < RK_COUNT; ++m) a.lslot[m] = (kind_mask >> m) & 1 ? (signed char)(a.nl++) : (signed char)-1;
  }
  CUtensorMap tmX, tmQ;
  setup_tma(a, dtype, pl.TQ, pl.nq_pad, pl.smem, &tmX, &tmQ, pl.TR);
  return dispatch_scan(kind, dtype == B200IR_F32, tmX, tmQ, a, pl.TR == 2 ? 16 : pl.TQ, pl.smem, st);
}

FallbackPlan make_fallback_plan(int dtype, int64_t nq, int64_t N, int D, int k) {
  FallbackPlan fp{};
  const int esz = dtype == B200IR_F32 ? 4 : 2;
  const int DKE = kRowChunkBytes / esz;
  const int TQ = 8;
  fp.D_pad = int(round_up64(D, DKE));
  fp.nq_pad = int(round_up64(nq, TQ));
  fp.sortn = scan_sortn(k);
  fp.smem = size_t(kScanStages) * scan_stage_bytes(TQ, DKE) + size_t(TQ) * fp.sortn * 8 + TQ * 16 + 64;
  int ctas_per_sm = int((227 * 1024) / (fp.smem + 1024));
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  const int64_t target = int64_t(kNumSMs) * ctas_per_sm;
  const int64_t ntiles = ceil_div64(N, kScanThreads);
  size_t off = 0;
  fp.off_qf = off; off += round_up64(size_t(fp.nq_pad) * fp.D_pad * 4, 256);
  fp.off_qn = off; off += round_up64(size_t(fp.nq_pad) * 4, 256);
  const int sizes[kFallbackTiers] = {8, 64, 512, 1 << 30};
  int base = 0;
  for (int t = 0; t < kFallbackTiers && base < fp.nq_pad; ++t) {
    const int cnt = int(fp.nq_pad - base < sizes[t] ? fp.nq_pad - base : sizes[t]);
    fp.base[t] = base; fp.count[t] = cnt; fp.G[t] = cnt / TQ;
    int64_t P = target / fp.G[t];
    P = P < 1 ? 1 : (P > ntiles ? ntiles : P);
    P = P < 1 ? 1 : P;
    fp.rows_per_part[t] = round_up64(ceil_div64(N, P), kScanThreads);
    fp.P[t] = int(ceil_div64(N, fp.rows_per_part[t]));
    if (fp.P[t] < 1) fp.P[t] = 1;
    fp.off_partial[t] = off; off += round_up64(size_t(cnt) * fp.P[t] * k * 8, 256);
    base += cnt;
    fp.ntiers = t + 1;
  }
  fp.total_bytes = off;
  return fp;
}

cudaError_t run_fallback(const FallbackPlan& fp, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k,
                         const MetricParams& mp, const int* fb_count, const int* fb_list, unsigned char* ws, int64_t index_offset,
                         float* out_score, int64_t* out_idx, cudaStream_t st) {
  (void)nq;
  float* Qf = reinterpret_cast<float*>(ws + fp.off_qf);
  float* qn = reinterpret_cast<float*>(ws + fp.off_qn);
  {
    ProfileScope ps(PT_PREP, st);
    const int blocks = int(ceil_div64(fp.nq_pad, 8));
    if (dtype == B200IR_F32)
      prep_queries_gather_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(Q), fb_list, fb_count, D, fp.nq_pad, fp.D_pad, Qf, qn);
    else
      prep_queries_gather_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(Q), fb_list, fb_count, D, fp.nq_pad, fp.D_pad, Qf, qn);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const int esz = dtype == B200IR_F32 ? 4 : 2;
  for (int t = 0; t < fp.ntiers; ++t) {
    ScanArgs a{};
    a.X = X; a.N = N; a.D = D; a.D_pad = fp.D_pad;
    a.Qf = Qf + size_t(fp.base[t]) * fp.D_pad; a.qnorm = qn + fp.base[t]; a.nq = fp.count[t];
    a.G = fp.G[t]; a.P = fp.P[t]; a.rows_per_part = fp.rows_per_part[t]; a.k = k; a.sortn = fp.sortn;
    a.aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (int64_t(D) * esz) % 16 == 0) ? 1 : 0;
    a.partial = reinterpret_cast<uint64_t*>(ws + fp.off_partial[t]);
    a.out_all = nullptr;
    a.mp = mp;
    a.gate = fb_count; a.gate_base = fp.base[t];
    CUtensorMap tmX, tmQ;
    setup_tma(a, dtype, 8, fp.count[t], fp.smem, &tmX, &tmQ);
    cudaError_t e = dispatch_scan(scan_kind_of(mp.metric), dtype == B200IR_F32, tmX, tmQ, a, 8, fp.smem, st);
    if (e != cudaSuccess) return e;
    FinalizeOpts fo;
    fo.qmap = fb_list + fp.base[t]; fo.gate = fb_count; fo.gate_base = fp.base[t];
    e = launch_finalize(a.partial, fp.count[t], int64_t(fp.P[t]) * k, k, mp, index_offset, out_score, out_idx, st, fo);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

size_t allpairs_eval_workspace_bytes(int64_t N, int D, int nthr) {
  const int D_pad = int(round_up64(D, 32));
  const int64_t n_pad = round_up64(N, kEvalTQ);
  return size_t(round_up64(size_t(n_pad) * D_pad * 4, 256)) + size_t(round_up64(size_t(n_pad) * 4, 256)) +
         size_t(round_up64(size_t(nthr) * 8, 256));
}

cudaError_t run_allpairs_eval(const float* X, const int32_t* cat, const int32_t* col, int64_t N, int D, int nbins,
                              const float* lo, const float* hi, const double* thresholds_host, int nthr,
                              unsigned long long* hist, unsigned long long* thr_counts, unsigned char* ws, cudaStream_t st,
                              int part, int nparts) {
  const int D_pad = int(round_up64(D, 32));
  const int64_t n_pad = round_up64(N, kEvalTQ);
  size_t off = 0;
  float* Qf = reinterpret_cast<float*>(ws + off); off += round_up64(size_t(n_pad) * D_pad * 4, 256);
  float* qn = reinterpret_cast<float*>(ws + off); off += round_up64(size_t(n_pad) * 4, 256);
  double* thr_d = reinterpret_cast<double*>(ws + off);
  cudaError_t e = cudaMemcpyAsync(thr_d, thresholds_host, size_t(nthr) * 8, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(hist, 0, size_t(kEvalMetrics) * 4 * nbins * 8, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(thr_counts, 0, size_t(kEvalMetrics) * 2 * (nthr + 1) * 8, st);
  if (e != cudaSuccess) return e;
  e = launch_prep_queries(B200IR_F32, X, int(N), D, int(n_pad), D_pad, Qf, qn, st);
  if (e != cudaSuccess) return e;
  ScanArgs a{};
  a.X = X; a.N = N; a.D = D; a.Qf = Qf; a.qnorm = qn; a.nq = int(N); a.D_pad = D_pad;
  // query groups are dealt to the parts cyclically (the pair grid is triangular: low groups meet the most rows)
  const int groups = int(n_pad / kEvalTQ);
  a.g_first = part; a.g_stride = nparts;
  a.G = part < groups ? (groups - part + nparts - 1) / nparts : 0;
  a.P = 1; a.rows_per_part = round_up64(N, kScanThreads); a.k = 1; a.sortn = 256;
  if (a.G == 0) return cudaSuccess;
  a.aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (int64_t(D) * 4) % 16 == 0) ? 1 : 0;
  a.partial = nullptr;
  a.out_all = nullptr;
  a.mp.metric = B200IR_OPTIMIZED; a.mp.flags = 0; a.mp.D = D;
  a.cat = cat; a.col = col; a.hist = hist; a.thr_counts = thr_counts; a.thresholds = thr_d; a.nbins = nbins; a.nthr = nthr;
  for (int m = 0; m < kEvalMetrics; ++m) { a.lo[m] = lo[m]; a.inv_w[m] = float(nbins) / (hi[m] - lo[m]); }
  CUtensorMap tmX, tmQ;
  const size_t smem = eval_smem_bytes(nbins, nthr);
  setup_tma(a, B200IR_F32, kEvalTQ, int(n_pad), smem, &tmX, &tmQ);
  ProfileScope ps(PT_SCAN, st);
  return launch_scan_eval_f32(tmX, tmQ, a, smem, st);
}

}  // namespace b200ir
