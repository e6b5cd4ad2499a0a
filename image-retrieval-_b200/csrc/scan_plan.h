// Host-side plan of the CUDA-core scan: tiling, partitioning and workspace layout.
#pragma once
#include "common.cuh"
#include "tma_util.h"

namespace b200ir {

enum ScanKind { K_L1 = 0, K_L2 = 1, K_LINF = 2, K_DOT = 3, K_MULTI = 4, K_EVAL = 5, K_MULTI6 = 6 };

// K_MULTI6: ONE pass keeps up to six candidate lists per query, one per ranking (search_with_multiple_metrics,
// app_pipeline.py:296-328 / image_search.py:199-219, runs one scan + sort per metric instead)
enum RankKind { RK_COS = 0, RK_L1 = 1, RK_L2 = 2, RK_LINF = 3, RK_MAG = 4, RK_OPT = 5, RK_COUNT = 6 };
__host__ __device__ inline int rank_kind_of(int metric) {
  switch (metric) {
    case B200IR_L1: return RK_L1;
    case B200IR_L2: return RK_L2;
    case B200IR_LINF: return RK_LINF;
    case B200IR_MAG_DIFF: return RK_MAG;
    case B200IR_OPTIMIZED: return RK_OPT;
    default: return RK_COS;   // cosine similarity / distance / angle share the descending-cosine ranking
  }
}

constexpr int kScanThreads = 128;   // == rows per tile (one row per thread)
constexpr int kScanStages = 2;
constexpr int kEvalStages = 1;     // all-pairs evaluation: the bins take 40 KB, a one-stage ring lets three CTAs share an SM
constexpr int kRowChunkBytes = 128; // bytes of one row staged per pipeline step

__host__ __device__ inline int scan_kind_of(int metric) {
  switch (metric) {
    case B200IR_L1: return K_L1;
    case B200IR_L2: return K_L2;
    case B200IR_LINF: return K_LINF;
    case B200IR_OPTIMIZED: return K_MULTI;
    default: return K_DOT;   // cosine family + magnitude difference
  }
}

struct ScanArgs {
  const void* X;          // [N, D] database shard, element type T
  int64_t N;
  int D;
  const float* Qf;        // prepared queries fp32 [nq_pad, D_pad], zero padded
  const float* qnorm;     // [nq_pad] |q|
  int nq;
  int D_pad;              // multiple of the row chunk (32 fp32 / 64 bf16 elements)
  int G;                  // query groups of TQ
  int P;                  // row partitions
  int64_t rows_per_part;  // multiple of 128
  int k;
  int sortn;              // 256 or 512: capacity of the per-query candidate buffer
  int aligned;            // rows are 16-byte aligned -> cp.async path
  int use_tma;            // tensor maps valid: one thread issues cp.async.bulk.tensor per stage instead of 9 cp.async per thread
  int bar_off;            // byte offset of the stage mbarriers in dynamic shared memory
  uint64_t* partial;      // [nq, P, k] sorted keys            (top-k mode)
  float* out_all;         // [nq, N] metric values, or nullptr  (pairwise mode)
  MetricParams mp;
  // fallback mode (queries that failed the tensor path's exactness certificate): the launch serves positions
  // [gate_base, gate_base + nq) of a device-side list whose length *gate is only known on the device
  const int* gate;        // nullptr: plain launch
  int gate_base;
  // evaluation mode across GPUs: this launch serves query groups g_first, g_first + g_stride, ... (g_stride 0 == 1)
  int g_first, g_stride;
  // paging (result pages beyond B200IR_MAX_K): only keys strictly after the cursor of the previous page are candidates
  const uint64_t* after;  // [nq] or nullptr
  // K_MULTI6: list slot of each ranking kind (-1: not requested) and the number of lists kept per query
  signed char lslot[RK_COUNT];
  int nl;
  // all-pairs evaluation mode (K_EVAL): queries == database rows, pairs i < j only
  const int32_t* cat;     // [N] object category of each row
  const int32_t* col;     // [N] colour of each row
  unsigned long long* hist;        // [5][4][nbins] counts per (metric, relationship type, bin)
  unsigned long long* thr_counts;  // [5][2][nthr + 1] first-threshold-index counts for the two same-object labels
  const double* thresholds;        // [nthr] ascending (device)
  int nbins, nthr;
  float lo[5], inv_w[5];           // bin = clamp((v - lo) * inv_w, 0, nbins - 1)
};

constexpr int kEvalMetrics = 5;    // cosine_distance, l1, l2, linf, magnitude_difference (mi_analysis.py:183-189)
constexpr int kEvalTQ = 8;
// one pipeline stage: 16 KB database tile + the fp32 query chunk, padded so that every tile stays 1024-byte aligned
// (required by the 128-byte TMA swizzle)
inline size_t scan_stage_bytes(int TQ, int DKE, int TR = 1) {
  return size_t(kScanThreads) * TR * kRowChunkBytes + size_t(round_up64(size_t(TQ) * DKE * 4, 1024));
}
// density bins in shared memory: 16-bit counters, two per 32-bit word
__host__ __device__ inline int eval_hist_words(int nbins) { return (kEvalMetrics * 4 * nbins + 1) / 2; }
inline size_t eval_smem_bytes(int nbins, int nthr) {
  return size_t(kEvalStages) * scan_stage_bytes(kEvalTQ, 32) +
         size_t(eval_hist_words(nbins)) * 4 + size_t(kEvalMetrics) * 2 * (nthr + 1) * 4 + size_t(nthr) * 8 + 16 + 64;
}
cudaError_t launch_scan_eval_f32(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, size_t smem, cudaStream_t st);

struct ScanPlan {
  int TQ, TR, G, P, sortn, D_pad, nq_pad;      // TR: database rows per thread (2 only with TQ = 8, single-list kinds)
  int64_t rows_per_part;
  size_t smem;
  size_t off_qf, off_qn, off_partial, total_bytes;
};

inline int scan_sortn(int k) { return k <= 128 ? 256 : 512; }

// nl > 1: multi-list scan (K_MULTI6) keeping nl candidate lists per query
inline ScanPlan make_scan_plan(int metric, int dtype, int64_t nq, int64_t N, int D, int k, bool pairwise, int nl = 1) {
  ScanPlan pl{};
  const int kind = scan_kind_of(metric);
  const int esz = dtype == B200IR_F32 ? 4 : 2;
  const int DKE = kRowChunkBytes / esz;
  int tq_max = 8;
  if (nl > 1 && nl * scan_sortn(k) * 8 * 8 > 100 * 1024) tq_max = 4;     // keep the lists of a CTA under ~100 KB
  (void)kind;
  pl.TQ = nq <= 1 ? 1 : (nq <= 4 ? 4 : tq_max);
  // two rows per thread for the 8-query pass of the single-list kinds (a 256-row tile can add 256 candidates: 512-key lists)
  pl.TR = (pl.TQ == 8 && nl == 1 && kind != K_MULTI && N >= 4 * kScanThreads) ? 2 : 1;
  pl.G = int(ceil_div64(nq, pl.TQ));
  pl.nq_pad = pl.G * pl.TQ;
  pl.D_pad = int(round_up64(D, DKE));
  pl.sortn = pairwise ? 256 : (pl.TR == 2 ? 512 : scan_sortn(k));
  pl.smem = size_t(kScanStages) * scan_stage_bytes(pl.TQ, DKE, pl.TR) + size_t(pl.TQ) * nl * pl.sortn * 8 + pl.TQ * nl * 16 + 64;
  int ctas_per_sm = int((227 * 1024) / (pl.smem + 1024));      // 1 KB per resident CTA is reserved by the driver
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  const int64_t target = int64_t(kNumSMs) * ctas_per_sm;
  int64_t P = target / pl.G;
  if (P < 1) P = 1;
  {
    // G x P CTAs run in waves of `target`: when the default P leaves more than 5 % of the last wave idle (G = 128,
    // target = 296: P = 2 fills 86 %), take the smallest P whose last wave is at least 95 % full (P = 9: 97 %)
    const int64_t ctas0 = pl.G * P;
    const double eff0 = double(ctas0) / double(ceil_div64(ctas0, target) * target);
    if (eff0 < 0.95 && pl.G > 1) {
      const int64_t max_p = N / (int64_t(16) * kScanThreads * pl.TR) > 1 ? N / (int64_t(16) * kScanThreads * pl.TR) : 1;
      double best_eff = eff0;
      for (int64_t cand = 1; cand <= 64 && cand <= max_p; ++cand) {
        const int64_t ctas = pl.G * cand;
        const double eff = double(ctas) / double(ceil_div64(ctas, target) * target);
        if (eff > best_eff + 1e-9) { best_eff = eff; P = cand; }
        if (eff >= 0.95) break;
      }
    }
  }
  const int tile_rows = kScanThreads * pl.TR;
  const int64_t ntiles = ceil_div64(N, tile_rows);
  if (P > ntiles) P = ntiles;
  if (P < 1) P = 1;
  pl.rows_per_part = round_up64(ceil_div64(N, P), tile_rows);
  if (pl.rows_per_part < tile_rows) pl.rows_per_part = tile_rows;
  pl.P = int(ceil_div64(N, pl.rows_per_part));
  if (pl.P < 1) pl.P = 1;
  size_t off = 0;
  pl.off_qf = off; off += round_up64(size_t(pl.nq_pad) * pl.D_pad * 4, 256);
  pl.off_qn = off; off += round_up64(size_t(pl.nq_pad) * 4, 256);
  pl.off_partial = off;
  if (!pairwise) off += round_up64(size_t(nl) * nq * pl.P * k * 8, 256);     // [nl][nq][P][k]
  pl.total_bytes = off;
  return pl;
}

// one translation unit per (kind, dtype): scan_inst.cu compiled with -DSCAN_KIND / -DSCAN_BF16
#define B200IR_DECL_SCAN(kind) \
  cudaError_t launch_scan_##kind##_f32(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, int TQ, size_t smem, cudaStream_t st); \
  cudaError_t launch_scan_##kind##_bf16(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, int TQ, size_t smem, cudaStream_t st);
B200IR_DECL_SCAN(K_L1) B200IR_DECL_SCAN(K_L2) B200IR_DECL_SCAN(K_LINF) B200IR_DECL_SCAN(K_DOT) B200IR_DECL_SCAN(K_MULTI)
B200IR_DECL_SCAN(K_MULTI6)
#undef B200IR_DECL_SCAN

// Exact re-do of the queries a tensor-core search could not certify (gemm_topk.cu).  Their number is only known on the
// device, so the work is cut into tiers of 8 / 64 / 512 / rest list positions, each a scan launch shaped for its size
// (few queries -> many row partitions) whose CTAs exit at once when the list is shorter than their tier.
constexpr int kFallbackTiers = 4;
struct FallbackPlan {
  int ntiers, D_pad, nq_pad, sortn;
  int base[kFallbackTiers], count[kFallbackTiers], G[kFallbackTiers], P[kFallbackTiers];
  int64_t rows_per_part[kFallbackTiers];
  size_t smem, off_qf, off_qn, off_partial[kFallbackTiers], total_bytes;
};
FallbackPlan make_fallback_plan(int dtype, int64_t nq, int64_t N, int D, int k);
cudaError_t run_fallback(const FallbackPlan& fp, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k,
                         const MetricParams& mp, const int* fb_count, const int* fb_list, unsigned char* ws, int64_t index_offset,
                         float* out_score, int64_t* out_idx, cudaStream_t st);

cudaError_t launch_prep_queries(int dtype, const void* Q, int nq, int D, int nq_pad, int D_pad, float* Qf, float* qn,
                                cudaStream_t st);

// Runs prep + scan.  In top-k mode leaves [nq, P, k] sorted keys at ws + plan.off_partial.  `after`: paging cursor
// per query (or nullptr).  kind_mask != 0: multi-list scan, [nl][nq][P][k] keys, list l = l-th set bit of kind_mask.
cudaError_t run_scan(const ScanPlan& pl, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k,
                     const MetricParams& mp, unsigned char* ws, float* out_all, cudaStream_t st,
                     const uint64_t* after = nullptr, int kind_mask = 0);

size_t allpairs_eval_workspace_bytes(int64_t N, int D, int nthr);
cudaError_t run_allpairs_eval(const float* X, const int32_t* cat, const int32_t* col, int64_t N, int D, int nbins,
                              const float* lo, const float* hi, const double* thresholds_host, int nthr,
                              unsigned long long* hist, unsigned long long* thr_counts, unsigned char* ws, cudaStream_t st,
                              int part = 0, int nparts = 1);

}  // namespace b200ir
