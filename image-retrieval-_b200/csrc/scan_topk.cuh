// CUDA-core distance scan with fused top-k (K5/K6 of SURVEY.md section 2.2, plus the exact
// fp32 forms of L2 / cosine / magnitude / weighted "optimized" similarity).
//
// Replaces the per-pair Python loops of app_pipeline.py:156-168 / :296-328 and
// geometric_metrics.py:12-57: one CTA streams a contiguous range of database rows through
// shared memory (cp.async, 128-byte swizzled row chunks), every thread owns ONE database row of
// the current 128-row tile and accumulates its distance to TQ queries at once (queries are
// broadcast reads from shared memory), and winners go through a per-query threshold filter into
// a shared-memory candidate buffer that is compacted by a warp-wide bitonic sort.  The distance
// matrix never reaches HBM; each CTA emits k sorted 64-bit keys per query.
#pragma once
#include "common.cuh"
#include "scan_plan.h"

namespace b200ir {

// |q| and fp32 copy of the queries, padded with zeros to [nq_pad, D_pad].  One warp per row.
template <typename T>
__global__ void prep_queries_kernel(const T* __restrict__ Q, int nq, int D, int nq_pad, int D_pad,
                                    float* __restrict__ Qf, float* __restrict__ qnorm) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nq_pad) return;
  float ss = 0.f;
  for (int d0 = 0; d0 < D_pad; d0 += 32) {
    const int d = d0 + lane;
    float v = 0.f;
    if (warp < nq && d < D) v = to_f32<T>(Q[int64_t(warp) * D + d]);
    if (d < D_pad) Qf[int64_t(warp) * D_pad + d] = v;
    float p = v * v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    ss += p;
  }
  if (lane == 0) qnorm[warp] = sqrtf(ss);
}

// Same for the first *count rows of a device-side query list (fallback of the tensor path): row w of the output is
// query list[w]; rows past the list up to the next multiple of 8 are zeroed, the rest is never read.
template <typename T>
__global__ void prep_queries_gather_kernel(const T* __restrict__ Q, const int* __restrict__ list, const int* __restrict__ count,
                                           int D, int nq_pad, int D_pad, float* __restrict__ Qf, float* __restrict__ qnorm) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int n = *count;
  if (warp >= nq_pad || warp >= ((n + 7) & ~7)) return;
  const int src = warp < n ? list[warp] : -1;
  float ss = 0.f;
  for (int d0 = 0; d0 < D_pad; d0 += 32) {
    const int d = d0 + lane;
    float v = 0.f;
    if (src >= 0 && d < D) v = to_f32<T>(Q[int64_t(src) * D + d]);
    if (d < D_pad) Qf[int64_t(warp) * D_pad + d] = v;
    float p = v * v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    ss += p;
  }
  if (lane == 0) qnorm[warp] = sqrtf(ss);
}

constexpr int kSelectMaxK = 32;     // up to this k the candidate buffers are compacted by selection, above by sorting

// Warp-cooperative compaction of one query's candidate buffer: sort, keep the best k.
template <int E>
__device__ __noinline__ void compact_candidates(uint64_t* keys, int* cnt, uint64_t* thr, int k, int lane) {
  const int n = *cnt;
  uint64_t r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    r[e] = i < n ? keys[i] : kKeyInf;
  }
  warp_sort<E>(r, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    if (i < k) keys[i] = r[e];
  }
  const int kept = n < k ? n : k;
  // threshold = current k-th best key (a candidate must be strictly smaller to enter)
  const int src_lane = (k - 1) / E, src_e = (k - 1) % E;
  uint64_t kth = kKeyInf;
#pragma unroll
  for (int e = 0; e < E; ++e) if (e == src_e) kth = r[e];
  kth = shfl_u64(kth, src_lane);
  if (lane == 0) { *cnt = kept; *thr = (n >= k) ? kth : kKeyInf; }
  __syncwarp();
}

// Small k: k rounds of warp arg-min instead of sorting the whole buffer.  A CTA scans only a few thousand rows, so the
// compactions of its cold start (every row is a candidate until k of them are known) are a fixed cost per CTA and per
// query; for k <= 32 the selection is ~5x cheaper than the 32*E-key bitonic sort.
template <int E>
__device__ __noinline__ void compact_select(uint64_t* keys, int* cnt, uint64_t* thr, int k, int lane) {
  const int n = *cnt;
  uint64_t r[E];
  uint64_t m = kKeyInf;                                   // smallest key this lane still holds
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    r[e] = i < n ? keys[i] : kKeyInf;
    m = r[e] < m ? r[e] : m;
  }
  __syncwarp();                                           // every key is in registers before the front of the list is rewritten
  uint64_t kth = kKeyInf;
  for (int j = 0; j < k; ++j) {
    uint64_t g = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const uint64_t other = shfl_xor_u64(g, o); g = other < g ? other : g; }
    if (g == kKeyInf) break;                              // fewer than k candidates
    if (m == g) {                                         // keys are unique (they carry the row index): exactly one owner
      keys[j] = g;
      m = kKeyInf;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        r[e] = r[e] == g ? kKeyInf : r[e];
        m = r[e] < m ? r[e] : m;
      }
    }
    kth = g;
  }
  if (lane == 0) { *cnt = n < k ? n : k; *thr = (n >= k) ? kth : kKeyInf; }
  __syncwarp();
}

// Two differences at once with ONE issue slot (sm_100 packed fp32: sub.rn.f32x2 -> SASS FADD2): the 8-query pass is
// issue-bound (ncu: issue slots 73 % busy against 54 % for the FMA pipe), and the packed form keeps the arithmetic of
// every element exactly as the scalar code has it.
__device__ __forceinline__ void sub2(float x0, float x1, float q0, float q1, float& d0, float& d1) {
  unsigned long long xa, qa, da;
  asm("mov.b64 %0, {%1, %2};" : "=l"(xa) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(qa) : "f"(q0), "f"(q1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(da) : "l"(xa), "l"(qa));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(da));
}
__device__ __forceinline__ float fmax3_abs(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c)));     // 3-input max (sm_100)
  return d;
}

// two consecutive elements of a row against two consecutive elements of a query
template <int KIND>
__device__ __forceinline__ void accum2(float (&a)[(KIND == K_MULTI || KIND == K_EVAL || KIND == K_MULTI6) ? 4 : 1], float x0, float x1,
                                       float q0, float q1) {
  if constexpr (KIND == K_DOT) { a[0] = fmaf(x0, q0, a[0]); a[0] = fmaf(x1, q1, a[0]); }
  else {
    float d0, d1;
    sub2(x0, x1, q0, q1, d0, d1);
    if constexpr (KIND == K_L1) { a[0] += fabsf(d0); a[0] += fabsf(d1); }
    else if constexpr (KIND == K_L2) { a[0] = fmaf(d0, d0, a[0]); a[0] = fmaf(d1, d1, a[0]); }
    else if constexpr (KIND == K_LINF) a[0] = fmax3_abs(a[0], d0, d1);
    else {
      a[0] = fmaf(x0, q0, a[0]); a[0] = fmaf(x1, q1, a[0]);
      a[1] += fabsf(d0); a[1] += fabsf(d1);
      a[2] = fmaf(d0, d0, a[2]); a[2] = fmaf(d1, d1, a[2]);
      a[3] = fmax3_abs(a[3], d0, d1);
    }
  }
}

// rank value r (smaller = better) of one (query,row) pair from its accumulators
template <int KIND>
__device__ __forceinline__ float finish_rank(const float* acc, float xsq, float qn, const MetricParams& mp) {
  if constexpr (KIND == K_L1 || KIND == K_L2 || KIND == K_LINF) {
    return acc[0];
  } else {
    const float xn = sqrtf(xsq);
    float cs = 0.f;
    if (qn != 0.f && xn != 0.f) cs = acc[0] / (qn * xn);        // geometric_metrics.py:14-18
    if constexpr (KIND == K_DOT) {
      if (mp.metric == B200IR_MAG_DIFF) return fabsf(qn - xn);   // geometric_metrics.py:57
      if (mp.flags & B200IR_FLAG_ABS_SCORE) cs = fabsf(cs);      // app_pipeline.py:167
      return -cs;
    } else {
      const float fD = float(mp.D);
      float sim = mp.w[0] * cs - mp.w[1] * (acc[1] / fD) - mp.w[2] * (sqrtf(acc[2]) / sqrtf(fD))
                  - mp.w[3] * acc[3] - mp.w[4] * fabsf(qn - xn);   // geometric_metrics.py:85-92
      if (mp.flags & B200IR_FLAG_ABS_SCORE) sim = fabsf(sim);
      return -sim;
    }
  }
}

// TR = database rows per thread (1, or 2 for the 8-query instantiations of the single-list kinds): with two rows per
// thread every broadcast query load feeds two rows (half the LDS per FADD) and a thread carries 16 independent
// accumulator chains, which is what the issue-bound 8-query pass and the batch regime (one pass per 8 queries) need.
template <int KIND, typename T, int TQ, int TR = 1>
__global__ void __launch_bounds__(kScanThreads) scan_topk_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                 const __grid_constant__ CUtensorMap tmQ, const ScanArgs a) {
  constexpr int DKE = kRowChunkBytes / int(sizeof(T));     // elements of a row per pipeline step
  constexpr int NA = (KIND == K_MULTI || KIND == K_EVAL || KIND == K_MULTI6) ? 4 : 1;
  constexpr bool NEED_XSQ = (KIND == K_DOT || KIND == K_MULTI || KIND == K_EVAL || KIND == K_MULTI6);
  const int NL = KIND == K_MULTI6 ? a.nl : 1;              // candidate lists per query
  constexpr int TILE_ROWS = kScanThreads * TR;             // database rows per tile
  constexpr int XT_BYTES = TILE_ROWS * kRowChunkBytes;     // 16 KB (32 KB) database tile per stage
  constexpr int QC_BYTES = TQ * DKE * 4;                   // fp32 query chunk per stage
  constexpr int STAGE_BYTES = XT_BYTES + (QC_BYTES + 1023) / 1024 * 1024;   // tiles stay 1024-byte aligned (TMA swizzle)
  constexpr int NST = KIND == K_EVAL ? kEvalStages : kScanStages;           // ring depth

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stage_base = smem;
  uint64_t* keys_s = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES);
  uint64_t* thr_s = keys_s + size_t(TQ) * NL * a.sortn;
  int* cnt_s = reinterpret_cast<int*>(thr_s + TQ * NL);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = a.g_first + (blockIdx.x % a.G) * (a.g_stride > 0 ? a.g_stride : 1);
  const int p = blockIdx.x / a.G;
  int nq_eff = a.nq;
  if (a.gate != nullptr) {                                 // fallback tier: serve only the list positions that exist
    const int avail = *a.gate - a.gate_base;
    if (g * TQ >= avail) return;
    nq_eff = min(nq_eff, avail);
  }
  int64_t row_begin = int64_t(p) * a.rows_per_part;
  const int64_t row_end = min(a.N, row_begin + a.rows_per_part);
  // evaluation mode: bins live where the key buffers would be; tiles whose rows are all <= the first query are skipped
  // Density bins are 16-bit, two per word (half the shared memory -> two CTAs per SM); they are flushed to the 64-bit
  // global bins every kEvalFlushTiles tiles, before any of them can reach 2^16 (a tile adds at most 128 * TQ to a bin).
  uint32_t* ev_hist = reinterpret_cast<uint32_t*>(smem + NST * STAGE_BYTES);
  const int ev_hist_words = KIND == K_EVAL ? eval_hist_words(a.nbins) : 0;
  uint32_t* ev_thr = ev_hist + ev_hist_words;
  double* ev_thresholds = reinterpret_cast<double*>(
      smem + NST * STAGE_BYTES + round_up64((ev_hist_words + kEvalMetrics * 2 * ((KIND == K_EVAL ? a.nthr : 0) + 1)) * 4, 8));
  constexpr int kEvalFlushTiles = 65535 / (kScanThreads * TQ);
  auto flush_bins = [&]() {
    __syncthreads();
    const int nh = kEvalMetrics * 4 * a.nbins;
    for (int i = tid; i < ev_hist_words; i += kScanThreads) {
      const uint32_t w = ev_hist[i];
      if (w) {
        if (w & 0xffffu) atomicAdd(&a.hist[2 * i], (unsigned long long)(w & 0xffffu));
        if ((w >> 16) && 2 * i + 1 < nh) atomicAdd(&a.hist[2 * i + 1], (unsigned long long)(w >> 16));
        ev_hist[i] = 0;
      }
    }
    __syncthreads();
  };
  if constexpr (KIND == K_EVAL) {
    const int64_t sdiff = int64_t(g) * TQ - row_begin;
    if (sdiff >= kScanThreads - 1) row_begin += ((sdiff - (kScanThreads - 1)) / kScanThreads + 1) * kScanThreads;
    if (row_begin > row_end) row_begin = row_end;
    const int nh = ev_hist_words + kEvalMetrics * 2 * (a.nthr + 1);
    for (int i = tid; i < nh; i += kScanThreads) ev_hist[i] = 0;
    for (int i = tid; i < a.nthr; i += kScanThreads) ev_thresholds[i] = a.thresholds[i];
  }
  const int ntiles = int(ceil_div64(row_end - row_begin, TILE_ROWS));
  const int nchunks = a.D_pad / DKE;
  const int total = ntiles * nchunks;
  const bool topk_mode = a.out_all == nullptr;
  const unsigned char* Xb = static_cast<const unsigned char*>(a.X);
  const int64_t row_bytes = int64_t(a.D) * int64_t(sizeof(T));

  if (KIND != K_EVAL && tid < TQ * NL) { thr_s[tid] = kKeyInf; cnt_s[tid] = 0; }

  // Loader state advances incrementally (no divisions in the steady state).  Thread t copies the 16-byte
  // piece c = t & 7 of rows r0 + 16 i (r0 = t >> 3, i = 0..7): (r & 7) == (r0 & 7) for all of them, so the
  // swizzled destination is one per-thread constant plus i * 2 KB.
  const int ld_c = tid & 7, ld_r0 = tid >> 3;
  const uint32_t ld_dst = uint32_t(ld_r0 * kRowChunkBytes + ((ld_c ^ (ld_r0 & 7)) << 4));
  const unsigned char* ld_src = Xb + (row_begin + ld_r0) * row_bytes + ld_c * 16;
  const int64_t ld_istride = 16 * row_bytes;
  const float* ld_q = a.Qf + int64_t(g * TQ + tid / (DKE / 4)) * a.D_pad + (tid % (DKE / 4)) * 4;
  const uint32_t ld_qdst = uint32_t(XT_BYTES + ((tid / (DKE / 4)) * DKE + (tid % (DKE / 4)) * 4) * 4);
  int iss_tile = 0, iss_chunk = 0, iss_stage = 0;
  // TMA path: stage `s` completes on mbarrier full[s] (transaction bytes of the tile box + the query box)
  const uint32_t full_bar0 = smem_u32(smem + a.bar_off);
  if (a.use_tma) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    if (tid == 0) {
      for (int s = 0; s < NST; ++s) tma::mbar_init(full_bar0 + 8u * s, 1);
      tma::mbar_fence_init();
      tma::prefetch_desc(&tmX);
      tma::prefetch_desc(&tmQ);
    }
    __syncthreads();
  }

  auto issue = [&]() {
    if (a.use_tma) {
      // one elected thread: two bulk tensor copies per stage; the hardware applies the same 16-byte XOR swizzle and
      // zero-fills rows past N / columns past D
      if (iss_tile < ntiles) {
        if (tid == 0) {
          const uint32_t sbu = smem_u32(stage_base + iss_stage * STAGE_BYTES);
          const uint32_t bar = full_bar0 + 8u * iss_stage;
          tma::mbar_expect_tx(bar, XT_BYTES + QC_BYTES);
          tma::load_2d(sbu, &tmX, bar, iss_chunk * DKE, int(row_begin + int64_t(iss_tile) * TILE_ROWS));
          tma::load_2d(sbu + XT_BYTES, &tmQ, bar, iss_chunk * DKE, g * TQ);
        }
        if (++iss_chunk == nchunks) { iss_chunk = 0; ++iss_tile; }
        if (++iss_stage == NST) iss_stage = 0;
      }
      return;
    }
    if (iss_tile < ntiles) {
      unsigned char* sb = stage_base + iss_stage * STAGE_BYTES;
      const uint32_t sbu = smem_u32(sb);
      const int64_t row0 = row_begin + int64_t(iss_tile) * TILE_ROWS;
      const int64_t col_byte0 = int64_t(iss_chunk) * kRowChunkBytes;
      const unsigned char* src0 = ld_src + int64_t(iss_tile) * (TILE_ROWS * row_bytes) + col_byte0;
      const bool full = a.aligned && (row0 + TILE_ROWS <= row_end) && (col_byte0 + kRowChunkBytes <= row_bytes);
      if (full) {
#pragma unroll
        for (int i = 0; i < 8 * TR; ++i) cp_async_16(sbu + ld_dst + i * 2048, src0 + i * ld_istride, 16);
      } else {
#pragma unroll
        for (int i = 0; i < 8 * TR; ++i) {
          const int64_t grow = row0 + ld_r0 + 16 * i;
          const int64_t cb = col_byte0 + ld_c * 16;
          int nbytes = 0;
          if (grow < row_end) nbytes = int(max(int64_t(0), min(int64_t(16), row_bytes - cb)));
          const unsigned char* src = nbytes > 0 ? src0 + i * ld_istride : Xb;
          if (a.aligned) {
            cp_async_16(sbu + ld_dst + i * 2048, src, nbytes);
          } else {
            // rows not 16-byte aligned: element loads, staged through registers
            T tmp[16 / sizeof(T)];
#pragma unroll
            for (int e = 0; e < int(16 / sizeof(T)); ++e)
              tmp[e] = (int(e * sizeof(T)) < nbytes) ? reinterpret_cast<const T*>(src)[e] : T(0.f);
            *reinterpret_cast<uint4*>(sb + ld_dst + i * 2048) = *reinterpret_cast<uint4*>(tmp);
          }
        }
      }
      // fp32 query chunk: TQ rows x DKE floats (the workspace copy is zero padded: no guards needed)
      static_assert(TQ * DKE / 4 <= kScanThreads, "query chunk must fit one cp.async per thread");
      if (tid < TQ * DKE / 4) cp_async_16(sbu + ld_qdst, ld_q + iss_chunk * DKE, 16);
      if (++iss_chunk == nchunks) { iss_chunk = 0; ++iss_tile; }
      if (++iss_stage == NST) iss_stage = 0;
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < NST - 1; ++s) issue();

  float acc[TR][TQ][NA];
  float xsq[TR];
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    xsq[r] = 0.f;
#pragma unroll
    for (int t = 0; t < TQ; ++t)
#pragma unroll
      for (int j = 0; j < NA; ++j) acc[r][t][j] = 0.f;
  }

  float qn[TQ];
#pragma unroll
  for (int t = 0; t < TQ; ++t) qn[t] = a.qnorm[g * TQ + t];
  const bool paged = a.after != nullptr;                   // paging: candidates must sort strictly after the query's cursor

  int tile = 0, chunk = 0, stage = 0;
  for (int it = 0; it < total; ++it) {
    if (a.use_tma) {
      __syncthreads();                       // everyone is done with the stage that is refilled next
      issue();
      tma::mbar_wait(full_bar0 + 8u * stage, (it / NST) & 1);
    } else if constexpr (NST >= 2) {
      cp_async_wait<(NST >= 2 ? NST - 2 : 0)>();
      __syncthreads();
      issue();
    } else {
      __syncthreads();                       // one-stage ring: refill, then wait for this very chunk
      issue();
      cp_async_wait<0>();
      __syncthreads();
    }

    const unsigned char* sb = stage_base + stage * STAGE_BYTES;
    if (++stage == NST) stage = 0;
    const unsigned char* xrow = sb + tid * kRowChunkBytes;                 // row r of this thread: tile row tid + 128 r
    const float* qs = reinterpret_cast<const float*>(sb + XT_BYTES);

    float part[TR][TQ][NA];
    float xpart[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      xpart[r] = 0.f;
#pragma unroll
      for (int t = 0; t < TQ; ++t)
#pragma unroll
        for (int j = 0; j < NA; ++j) part[r][t][j] = (KIND == K_LINF || j == 3) ? acc[r][t][j] : 0.f;
    }

#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 raw[TR];
#pragma unroll
      for (int r = 0; r < TR; ++r)
        raw[r] = *reinterpret_cast<const uint4*>(xrow + r * (kScanThreads * kRowChunkBytes) + ((j ^ (tid & 7)) << 4));
      if constexpr (sizeof(T) == 4) {
        float x[TR][4];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          x[r][0] = __uint_as_float(raw[r].x); x[r][1] = __uint_as_float(raw[r].y);
          x[r][2] = __uint_as_float(raw[r].z); x[r][3] = __uint_as_float(raw[r].w);
          if constexpr (NEED_XSQ) {
#pragma unroll
            for (int e = 0; e < 4; ++e) xpart[r] = fmaf(x[r][e], x[r][e], xpart[r]);
          }
        }
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
          const float4 q4 = *reinterpret_cast<const float4*>(qs + t * DKE + j * 4);
#pragma unroll
          for (int r = 0; r < TR; ++r) {
            accum2<KIND>(part[r][t], x[r][0], x[r][1], q4.x, q4.y);
            accum2<KIND>(part[r][t], x[r][2], x[r][3], q4.z, q4.w);
          }
        }
      } else {
        float x[TR][8];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          x[r][0] = bf16_lo(raw[r].x); x[r][1] = bf16_hi(raw[r].x); x[r][2] = bf16_lo(raw[r].y); x[r][3] = bf16_hi(raw[r].y);
          x[r][4] = bf16_lo(raw[r].z); x[r][5] = bf16_hi(raw[r].z); x[r][6] = bf16_lo(raw[r].w); x[r][7] = bf16_hi(raw[r].w);
          if constexpr (NEED_XSQ) {
#pragma unroll
            for (int e = 0; e < 8; ++e) xpart[r] = fmaf(x[r][e], x[r][e], xpart[r]);
          }
        }
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
          const float4 qa = *reinterpret_cast<const float4*>(qs + t * DKE + j * 8);
          const float4 qb = *reinterpret_cast<const float4*>(qs + t * DKE + j * 8 + 4);
#pragma unroll
          for (int r = 0; r < TR; ++r) {
            accum2<KIND>(part[r][t], x[r][0], x[r][1], qa.x, qa.y);
            accum2<KIND>(part[r][t], x[r][2], x[r][3], qa.z, qa.w);
            accum2<KIND>(part[r][t], x[r][4], x[r][5], qb.x, qb.y);
            accum2<KIND>(part[r][t], x[r][6], x[r][7], qb.z, qb.w);
          }
        }
      }
    }
    // two-level summation: chunk partial -> row total (keeps fp32 error ~ sqrt-free (DKE + D/DKE) eps)
#pragma unroll
    for (int r = 0; r < TR; ++r) {
#pragma unroll
      for (int t = 0; t < TQ; ++t)
#pragma unroll
        for (int j = 0; j < NA; ++j) acc[r][t][j] = (KIND == K_LINF || j == 3) ? part[r][t][j] : acc[r][t][j] + part[r][t][j];
      xsq[r] += xpart[r];
    }

    if (++chunk == nchunks) {
      chunk = 0;
#pragma unroll
      for (int r = 0; r < TR; ++r) {
      const int64_t grow = row_begin + int64_t(tile) * TILE_ROWS + r * kScanThreads + tid;
      const bool valid = grow < row_end;
#pragma unroll
      for (int t = 0; t < TQ; ++t) {
        const int q = g * TQ + t;
        if constexpr (KIND == K_EVAL) {
          if (valid && q < nq_eff && grow > q) {
            // the five evaluation metrics of mi_analysis.py:183-189 from one pass (geometric_metrics.py:114-129)
            const float xn = sqrtf(xsq[r]);
            float cs = 0.f;
            if (qn[t] != 0.f && xn != 0.f) cs = acc[r][t][0] / (qn[t] * xn);
            const float fD = float(a.mp.D);
            const float vals[kEvalMetrics] = {1.0f - cs, acc[r][t][1] / fD, sqrtf(acc[r][t][2]) / sqrtf(fD), acc[r][t][3], fabsf(qn[t] - xn)};
            const int rel = (a.cat[q] == a.cat[grow] ? 0 : 2) + (a.col[q] == a.col[grow] ? 0 : 1);   // mi_analysis.py:176-181
#pragma unroll
            for (int m = 0; m < kEvalMetrics; ++m) {
              int bin = int(floorf((vals[m] - a.lo[m]) * a.inv_w[m]));
              bin = bin < 0 ? 0 : (bin >= a.nbins ? a.nbins - 1 : bin);
              const int slot = (m * 4 + rel) * a.nbins + bin;
              atomicAdd(&ev_hist[slot >> 1], 1u << ((slot & 1) * 16));
              if (rel <= 1) {
                // first threshold index with d <= threshold (thresholds ascending); nthr = none (mi_analysis.py:783-784)
                const double dv = double(vals[m]);
                int lo_i = 0, hi_i = a.nthr;
                while (lo_i < hi_i) {
                  const int mid = (lo_i + hi_i) >> 1;
                  if (dv <= ev_thresholds[mid]) hi_i = mid; else lo_i = mid + 1;
                }
                atomicAdd(&ev_thr[(m * 2 + rel) * (a.nthr + 1) + lo_i], 1u);
              }
            }
          }
        } else
        if (valid && q < nq_eff) {
          if constexpr (KIND == K_MULTI6) {
            // all six rankings of this (query, row) pair from the four accumulators (geometric_metrics.py:12-57, :85-92)
            const float xn = sqrtf(xsq[r]);
            float cs = 0.f;
            if (qn[t] != 0.f && xn != 0.f) cs = acc[r][t][0] / (qn[t] * xn);
            const float fD = float(a.mp.D);
            const float mag = fabsf(qn[t] - xn);
            float sim = a.mp.w[0] * cs - a.mp.w[1] * (acc[r][t][1] / fD) - a.mp.w[2] * (sqrtf(acc[r][t][2]) / sqrtf(fD)) - a.mp.w[3] * acc[r][t][3] -
                        a.mp.w[4] * mag;
            if (a.mp.flags & B200IR_FLAG_ABS_SCORE) { cs = fabsf(cs); sim = fabsf(sim); }
            const float rk[RK_COUNT] = {-cs, acc[r][t][1], acc[r][t][2], acc[r][t][3], mag, -sim};
#pragma unroll
            for (int m = 0; m < RK_COUNT; ++m) {
              const int ls = a.lslot[m];
              if (ls >= 0) {
                const int L = t * NL + ls;
                const uint64_t key = make_key(rk[m], uint32_t(grow));
                if (key < thr_s[L]) {
                  const int slot = atomicAdd(&cnt_s[L], 1);
                  keys_s[size_t(L) * a.sortn + slot] = key;
                }
              }
            }
          } else {
          const float rv = finish_rank<KIND>(acc[r][t], xsq[r], qn[t], a.mp);
          if (topk_mode) {
            const uint64_t key = make_key(rv, uint32_t(grow));
            if (key < thr_s[t] && (!paged || key > a.after[g * TQ + t])) {     // cursor read only on the (rare) hit path
              const int slot = atomicAdd(&cnt_s[t], 1);
              keys_s[size_t(t) * a.sortn + slot] = key;
            }
          } else {
            a.out_all[int64_t(q) * a.N + grow] = rank_to_score(rv, a.mp.metric, a.mp.flags, a.mp.D);
          }
          }
        }
#pragma unroll
        for (int j = 0; j < NA; ++j) acc[r][t][j] = 0.f;
      }
      xsq[r] = 0.f;
      }
      ++tile;
      if constexpr (KIND == K_EVAL) {
        if (tile % kEvalFlushTiles == 0) flush_bins();
      }
      if (KIND != K_EVAL && topk_mode) {
        __syncthreads();
        for (int t = warp; t < TQ * NL; t += kScanThreads / 32) {
          if (cnt_s[t] > a.sortn - TILE_ROWS) {
            if (a.k <= kSelectMaxK) {
              if (a.sortn == 256) compact_select<8>(keys_s + size_t(t) * a.sortn, &cnt_s[t], &thr_s[t], a.k, lane);
              else compact_select<16>(keys_s + size_t(t) * a.sortn, &cnt_s[t], &thr_s[t], a.k, lane);
            } else if (a.sortn == 256) compact_candidates<8>(keys_s + size_t(t) * a.sortn, &cnt_s[t], &thr_s[t], a.k, lane);
            else compact_candidates<16>(keys_s + size_t(t) * a.sortn, &cnt_s[t], &thr_s[t], a.k, lane);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  if constexpr (KIND == K_EVAL) {
    __syncthreads();
    const int nt = kEvalMetrics * 2 * (a.nthr + 1);
    flush_bins();
    for (int i = tid; i < nt; i += kScanThreads) if (ev_thr[i]) atomicAdd(&a.thr_counts[i], (unsigned long long)ev_thr[i]);
    return;
  }
  if (topk_mode) {
    __syncthreads();
    for (int L = warp; L < TQ * NL; L += kScanThreads / 32) {
      const int t = L / NL, ls = L % NL;
      const int q = g * TQ + t;
      if (q >= nq_eff) continue;
      uint64_t* kb = keys_s + size_t(L) * a.sortn;
      if (a.k <= kSelectMaxK) {
        if (a.sortn == 256) compact_select<8>(kb, &cnt_s[L], &thr_s[L], a.k, lane);
        else compact_select<16>(kb, &cnt_s[L], &thr_s[L], a.k, lane);
      } else if (a.sortn == 256) compact_candidates<8>(kb, &cnt_s[L], &thr_s[L], a.k, lane);
      else compact_candidates<16>(kb, &cnt_s[L], &thr_s[L], a.k, lane);
      const int kept = cnt_s[L];
      uint64_t* dst = a.partial + ((int64_t(ls) * a.nq + q) * a.P + p) * a.k;      // [nl][nq][P][k]
      for (int i = lane; i < a.k; i += 32) dst[i] = i < kept ? kb[i] : kKeyInf;
    }
  }
}


template <int KIND, typename T, int TQ, int TR = 1>
inline cudaError_t launch_scan_inst(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, size_t smem, cudaStream_t st) {
  auto kern = scan_topk_kernel<KIND, T, TQ, TR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  kern<<<a.G * a.P, kScanThreads, smem, st>>>(tmX, tmQ, a);
  return cudaGetLastError();
}

inline cudaError_t launch_scan_eval_inst(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, size_t smem, cudaStream_t st) {
  return launch_scan_inst<K_EVAL, float, kEvalTQ>(tmX, tmQ, a, smem, st);
}

template <int KIND, typename T>
inline cudaError_t launch_scan_tq(const CUtensorMap& tmX, const CUtensorMap& tmQ, const ScanArgs& a, int TQ, size_t smem, cudaStream_t st) {
  switch (TQ) {
    case 1: return launch_scan_inst<KIND, T, 1>(tmX, tmQ, a, smem, st);
    case 4: return launch_scan_inst<KIND, T, 4>(tmX, tmQ, a, smem, st);
    case 8: return launch_scan_inst<KIND, T, 8>(tmX, tmQ, a, smem, st);
    case 16:                                              // TQ code 16 = 8 queries x 2 rows per thread (single-list kinds)
      if constexpr (KIND == K_L1 || KIND == K_L2 || KIND == K_LINF || KIND == K_DOT) return launch_scan_inst<KIND, T, 8, 2>(tmX, tmQ, a, smem, st);
      else return cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace b200ir
