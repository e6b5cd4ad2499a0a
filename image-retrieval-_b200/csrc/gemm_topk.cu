// tcgen05 / TMEM / TMA path: bf16 cosine-family and L2 top-k as a dense contraction
// (K1-K4 of SURVEY.md section 2.2) with the top-k selection fused into the accumulator epilogue.
//
// Replaces the scan loops of app_pipeline.py:156-172 / :296-328 for bf16 stores: the query batch
// and the database are the A and B operands (both K-major) of C = Q . X^T.
//
//   grid      : one persistent CTA per SM, 384 threads, warp-specialised:
//                 warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor, 128B-swizzled tiles)
//                 warp 1 lane 0 : MMA issuer    (tcgen05.mma.cta_group::1.kind::f16, M128 N256 K16)
//                 warp 2        : TMEM allocator (512 columns = two 128x256 fp32 accumulators)
//                 warps 4..11   : epilogue, two warps per TMEM lane quarter sharing one candidate list per query
//   work unit : (128-query tile, contiguous range of 256-row database tiles).  The query tile stays
//               resident in shared memory (A, up to 128 KB for D = 512) for the whole unit; database
//               tiles stream through a 3-stage 32 KB ring (B).
//   epilogue  : tcgen05.ld 32 columns at a time; v = dot * rnorm_x (cosine) or 2 dot - |x|^2 (L2);
//               a per-thread running threshold rejects almost everything; survivors are appended
//               to the query's candidate list (L2-resident global scratch) which a warp-wide
//               bitonic sort compacts to the best k' when it fills.  The distance matrix never
//               exists in memory.  Each unit emits k' sorted keys per query; a final kernel merges
//               the partitions, re-ranks the k' candidates with exact fp32 arithmetic on the CUDA
//               cores (cancellation-free) and writes the k winners.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "gemm_topk.h"
#include "scan_plan.h"
#include "profile.h"
#include "select.cuh"

namespace b200ir {

namespace gemm {

constexpr int BM = 128;            // queries per tile (UMMA M, TMEM lanes)
#ifndef GEMM_BN
#define GEMM_BN 256
#endif
constexpr int BN = GEMM_BN;        // database rows per tile (UMMA N, TMEM columns per accumulator)
constexpr int BK = 64;             // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_KB = 8;          // D <= 512: the query tile stays resident in shared memory
constexpr int MAX_D_STREAM = 2048; // beyond 512 the query k-blocks are streamed (A-streamed kernel variant)
constexpr int TMEM_COLS = 512;
constexpr int NBUF = TMEM_COLS / BN;         // accumulators in flight (2 x 256 or 4 x 128 columns)
constexpr int B_STAGES = BN == 256 ? 3 : 5;
constexpr int A_KB_BYTES = BM * BK * 2;      // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int SMEM_A = MAX_KB * A_KB_BYTES;  // 128 KB
constexpr int SMEM_B = B_STAGES * B_STAGE_BYTES;   // 96 KB
constexpr int SMEM_SCALE = NBUF * BN * 4;    // per-column scale (rnorm / sqnorm), one buffer per accumulator
constexpr int SMEM_BARS = BN == 256 ? 192 : 320;
constexpr int SMEM_CNT = 3 * BM * 2;        // per-query 16-bit counters: front (warp 0 of the quarter), back (warp 1), sorted prefix
constexpr int SMEM_TOTAL = SMEM_A + SMEM_B + SMEM_SCALE + SMEM_BARS + SMEM_CNT;   // 232,384 B <= 232,448
// A-streamed variant (D > 512: the query tile no longer fits next to the ring): no resident A, a 13-slot ring of 16 KB
constexpr int STREAM_SLOTS = 13;
constexpr int SMEM_BARS_STREAM = 320;
constexpr int SMEM_TOTAL_STREAM = STREAM_SLOTS * (B_STAGE_BYTES / 2) + SMEM_SCALE + SMEM_BARS_STREAM + SMEM_CNT;
constexpr int THREADS = 384;           // 4 control warps + 8 epilogue warps (two per TMEM lane quarter)
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

enum Mode { MODE_COS = 0, MODE_ABSCOS = 1, MODE_L2 = 2 };

struct Args {
  int nq;
  int64_t N;
  int D;
  int num_kb;
  int num_qtiles;
  int num_qgroups;       // query tiles grouped per cluster (1 per CTA): ceil(num_qtiles / cluster size)
  int P;                 // database partitions
  int tiles_per_part;
  int total_tiles;
  int kp;                // candidates kept per (query, partition): k' >= k
  int cap;               // candidate list capacity (256 or 512)
  const float* colscale; // [total_tiles*256]: 1/|x| (cosine) or |x|^2 (L2), NaN past row N
  uint64_t* cand;        // [gridDim.x][128][cap]
  uint64_t* partial;     // [nq][P][kp]
  uint32_t* thr_g;       // [num_qtiles*128] best published k'-th rank value per query (ordered bits), 0xffffffff = none
  uint32_t* lvl;         // [num_qtiles*128][kLevels][P] per-partition order statistics (see publish_levels), or nullptr
  unsigned long long* dbg;   // optional [gridDim.x][16] cycle counters (B200IR_GEMM_DEBUG=1), else nullptr
  int opt;               // experiment switches (B200IR_GEMM_OPT): bit 0 = L2 prefetch of the tile two ahead
};

#define DBG_ON (a.dbg && !(a.opt & 64))            /* B200IR_GEMM_OPT bit 6: keep only the per-round timers */
#define DBG_T0() (DBG_ON ? clock64() : 0ll)
#define DBG_ADD(slot, t0) do { if (DBG_ON) a.dbg[blockIdx.x * 16 + (slot)] += (unsigned long long)(clock64() - (t0)); } while (0)

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
// 2-CTA (cta_group::2) variants: the transaction bytes are signalled on the barrier of the LEADER CTA (rank 0 of the
// pair: peer bit of the shared::cluster address cleared), MMAs are issued by the leader for both SMs, and commits
// arrive on the same barrier offset in both CTAs.
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows
// x 128 B) >> 4 in [32,46), descriptor version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6)=1, A bf16 [7,10)=1, B bf16 [10,13)=1,
// A/B K-major (bits 15,16 = 0), N>>3 in [17,23), M>>4 in [24,29).
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);
constexpr uint32_t kInstrDesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t((2 * BM) >> 4) << 24);   // M = 256 across the CTA pair

// ------------------------------------------------------------------------------------ epilogue helpers
// Warp-cooperative compaction of the candidate lists of the lanes in `mask`: sort, keep the best kp.
// If `out` is set the sorted list goes to the unit's output slot instead of back to the scratch list.
// Cross-unit threshold sharing: after a compaction the k'-th best rank value of a query is published with
// atomicMin; any unit working on the same query may discard candidates that are strictly worse (at least k'
// better rows exist somewhere).  Imported thresholds are non-strict (equal scores may still win on index).
__device__ __forceinline__ void import_threshold(const uint32_t* thr_g_q, float& thr) {
  const uint32_t g = *reinterpret_cast<const volatile uint32_t*>(thr_g_q);
  if (g != 0xffffffffu) {
    const float gv = -ordered_to_f32(g);
    thr = fmaxf(thr, nextafterf(gv, -INFINITY));
  }
}

// Warp-cooperative compaction of the candidate lists of the query rows in `mask` (bit L = row L of this TMEM lane
// quarter): sort, keep the best kp, publish the k'-th rank value.  Counters live in shared memory (`cnt_q`, one
// per row of the quarter) because two epilogue warps append to the same list.  If `out` is set the sorted list
// goes to the unit's output slot instead of back to the scratch list.
// Threshold sharing between the units (database partitions) of one query.  A unit publishes, for levels j = 0..3, the
// rank value of its ceil(kp / 2^j)-th best candidate so far (a statement "this partition holds that many rows at least
// this good", which only gets stronger with time).  If 2^j different partitions each hold ceil(kp / 2^j) rows at least
// as good as t, the database holds kp of them, so the 2^j-th best published level-j value is a valid threshold for
// every unit of the query.  Level 0 alone (the best k'-th value of any single partition) is what a per-unit list can
// give; the deeper levels make the threshold track the union of the partitions seen so far instead of one of them.
constexpr int kLevels = 4;
__device__ __forceinline__ void publish_level_values(const uint32_t (&mine)[kLevels], int lane, uint32_t* thr_g_q, uint32_t* lvl_q,
                                                     int p, int P);
template <int E>
__device__ __forceinline__ void publish_levels(const uint64_t (&r)[E], int n, int kp, int lane, uint32_t* thr_g_q, uint32_t* lvl_q,
                                               int p, int P) {
  uint32_t mine[kLevels];
#pragma unroll
  for (int j = 0; j < kLevels; ++j) {
    const int m = (kp + (1 << j) - 1) >> j;
    const int src_lane = (m - 1) / E, src_e = (m - 1) % E;
    uint32_t hi = 0xffffffffu;
#pragma unroll
    for (int e = 0; e < E; ++e) if (e == src_e) hi = uint32_t(r[e] >> 32);
    hi = __shfl_sync(0xffffffffu, hi, src_lane);
    mine[j] = n >= m ? hi : 0xffffffffu;
  }
  publish_level_values(mine, lane, thr_g_q, lvl_q, p, P);
}

// `mine[j]`: a rank value (ordered bits) such that this partition holds at least ceil(kp / 2^j) candidates at least that
// good (0xffffffff: no such statement yet)
__device__ __forceinline__ void publish_level_values(const uint32_t (&mine)[kLevels], int lane, uint32_t* thr_g_q, uint32_t* lvl_q,
                                                     int p, int P) {
  if (lvl_q == nullptr) {                                     // more partitions than lanes: level 0 only
    if (lane == 0 && mine[0] != 0xffffffffu) atomicMin(thr_g_q, mine[0]);
    return;
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kLevels; ++j)
      if (mine[j] != 0xffffffffu) *reinterpret_cast<volatile uint32_t*>(lvl_q + j * P + p) = mine[j];
  }
  __syncwarp();
  uint32_t best = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < kLevels; ++j) {
    uint32_t x[1] = {lane < P ? *reinterpret_cast<const volatile uint32_t*>(lvl_q + j * P + lane) : 0xffffffffu};
    warp_sort<1>(x, lane);                                    // ascending: lane i holds the (i+1)-th best published value
    best = min(best, __shfl_sync(0xffffffffu, x[0], (1 << j) - 1));
  }
  if (lane == 0 && best != 0xffffffffu) atomicMin(thr_g_q, best);
}

template <int E>
__device__ __noinline__ void compact_one(uint64_t* list, int cap, int kp, int n0, int n1, unsigned short* cnt_front,
                                         unsigned short* cnt_back, unsigned short* srt, int lane, uint64_t* dst,
                                         uint32_t* thr_g_q, uint32_t* lvl_q, int p, int P) {
  // the list is filled from both ends: warp 0 of the quarter appends at [0, n0), warp 1 at (cap - n1, cap]
  const int n = n0 + n1;
  uint64_t r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    r[e] = i < n0 ? list[i] : (i < n ? list[cap - 1 - (i - n0)] : kKeyInf);
  }
  warp_sort<E>(r, lane);
  uint64_t* out = dst ? dst : list;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    if (i < kp) out[i] = r[e];
  }
  if (dst == nullptr && lane == 0) {
    *cnt_front = (unsigned short)(n < kp ? n : kp);   // the compacted list lives at the front, sorted
    *cnt_back = 0;
    *srt = *cnt_front;
  }
  publish_levels<E>(r, n, kp, lane, thr_g_q, lvl_q, p, P);
  __syncwarp();
}

// Incremental compaction: the first `s` entries of the list are the sorted survivors of the previous compaction, so
// only the h <= 32*E entries appended since then are sorted; the best 32*E of both come out of one bitonic merge
// (C[i] = min(old[i], new[32E-1-i]) is bitonic and holds the 32E smallest).  Needs s <= 32*E and kp <= 32*E.
template <int E>
__device__ __noinline__ void compact_merge(uint64_t* list, int cap, int kp, int s, int n0, int n1, unsigned short* cnt_front,
                                           unsigned short* cnt_back, unsigned short* srt, int lane, uint64_t* dst,
                                           uint32_t* thr_g_q, uint32_t* lvl_q, int p, int P) {
  const int hf = n0 - s, h = hf + n1, n = n0 + n1;
  uint64_t rn[E], r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    rn[e] = i < hf ? list[s + i] : (i < h ? list[cap - 1 - (i - hf)] : kKeyInf);
  }
  warp_sort<E>(rn, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    const uint64_t old = i < s ? list[i] : kKeyInf;
    const uint64_t rev = shfl_u64(rn[E - 1 - e], 31 - lane);       // new[32E - 1 - i]
    r[e] = old < rev ? old : rev;
  }
  warp_bitonic_merge<E>(r, lane);
  uint64_t* out = dst ? dst : list;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    if (i < kp) out[i] = r[e];
  }
  if (dst == nullptr && lane == 0) {
    *cnt_front = (unsigned short)(n < kp ? n : kp);
    *cnt_back = 0;
    *srt = *cnt_front;
  }
  publish_levels<E>(r, n, kp, lane, thr_g_q, lvl_q, p, P);
  __syncwarp();
}

// Compaction by SELECTION (the in-loop compactions; the unit end still sorts): the list only has to shrink to its best
// ~kp keys and yield a threshold, it does not have to be ordered.  A bisection on the 32-bit rank value (one
// __reduce_add_sync count per step, ~10-14 steps for a 200-key list) finds a pivot with at least kp keys at or below
// it; the survivors are ballot-compacted to the front with the WORST survivor parked at position kp - 1, where the
// owner threads read their new threshold (at least kp rows are at least that good, so anything not strictly better
// cannot enter).  The counts seen along the bisection give valid statements for the shared threshold levels for free.
// ~400 instructions instead of the 1300-2500 of a sort / merge; falls back to the sorter when ties keep the pivot
// from separating (more than kp + 48 survivors).
template <int E>
__device__ __noinline__ bool compact_select(uint64_t* list, int cap, int kp, int n0, int n1, unsigned short* cnt_front,
                                            unsigned short* cnt_back, unsigned short* srt, int lane, uint32_t* thr_g_q,
                                            uint32_t* lvl_q, int p, int P) {
  const int n = n0 + n1;
  uint64_t r[E];
  uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    r[e] = i < n0 ? list[i] : (i < n ? list[cap - 1 - (i - n0)] : kKeyInf);
    if (i < n) { const uint32_t h = uint32_t(r[e] >> 32); lo = min(lo, h); hi = max(hi, h); }
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  int need[kLevels];
  uint32_t best[kLevels];
#pragma unroll
  for (int j = 0; j < kLevels; ++j) { need[j] = (kp + (1 << j) - 1) >> j; best[j] = hi; }     // all n >= kp keys are <= hi
  uint32_t L = lo, H = hi;                                   // invariant: count(rank <= H) >= kp
  int cH = n;
  for (int it = 0; it < 32 && L < H; ++it) {
    const uint32_t mid = L + ((H - L) >> 1);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (r[e] != kKeyInf && uint32_t(r[e] >> 32) <= mid) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
#pragma unroll
    for (int j = 1; j < kLevels; ++j) if (c >= need[j] && mid < best[j]) best[j] = mid;
    if (c >= kp) { H = mid; cH = c; if (c <= kp + 8) break; } else L = mid + 1;
  }
  if (cH > kp + 48) return false;                            // ties: let the sorter cut by (rank, row)
  best[0] = H;
  // worst survivor (largest key at or below the pivot): goes to position kp - 1
  uint64_t wkey = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) if (r[e] != kKeyInf && uint32_t(r[e] >> 32) <= H && r[e] > wkey) wkey = r[e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const uint64_t other = shfl_xor_u64(wkey, o); wkey = other > wkey ? other : wkey; }
  __syncwarp();                                              // every key is in registers before the front is rewritten
  int base = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool keep = r[e] != kKeyInf && uint32_t(r[e] >> 32) <= H && r[e] != wkey;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      int pos = base + __popc(m & ((1u << lane) - 1));
      pos += pos >= kp - 1 ? 1 : 0;                          // skip the slot reserved for the worst survivor
      list[pos] = r[e];
    }
    base += __popc(m);
  }
  if (lane == 0) {
    list[kp - 1] = wkey;
    *cnt_front = (unsigned short)(base + 1);
    *cnt_back = 0;
    *srt = 0;                                                // nothing is sorted
  }
  publish_level_values(best, lane, thr_g_q, lvl_q, p, P);
  __syncwarp();
  return true;
}

// Warp-cooperative compaction of the candidate lists of the query rows in `mask` (bit L = row L of this TMEM lane
// quarter): sort, keep the best kp, publish the k'-th rank value.  The sorter width follows the list length.
// If `out` is set the sorted list goes to the unit's output slot instead of back to the scratch list.
__device__ __forceinline__ void warp_compact(uint64_t* warp_lists, int cap, int kp, unsigned short* cnt_q, uint32_t mask, int lane,
                                             uint64_t* out, int64_t out_stride, int valid_lanes, uint32_t* thr_g_warp,
                                             uint32_t* lvl_warp, int p, int P, bool use_select = false) {
  while (mask) {
    const int L = __ffs(mask) - 1;
    mask &= mask - 1;
    const int n0 = cnt_q[L], n1 = cnt_q[32 + L];
    uint64_t* list = warp_lists + size_t(L) * cap;
    uint64_t* dst = nullptr;
    if (out != nullptr) {
      if (L >= valid_lanes) continue;
      dst = out + int64_t(L) * out_stride;
    }
    uint32_t* lvl_q = lvl_warp ? lvl_warp + size_t(L) * kLevels * P : nullptr;
    const int s = cnt_q[64 + L];                               // sorted prefix left by the previous compaction
    const int h = n0 + n1 - s;
    unsigned short *cf = cnt_q + L, *cb = cnt_q + 32 + L, *srt = cnt_q + 64 + L;
    if (use_select && dst == nullptr && n0 + n1 > kp) {      // in-loop compaction: selection first, the sorter on ties
      const bool done = n0 + n1 <= 256 ? compact_select<8>(list, cap, kp, n0, n1, cf, cb, srt, lane, thr_g_warp + L, lvl_q, p, P)
                                       : compact_select<16>(list, cap, kp, n0, n1, cf, cb, srt, lane, thr_g_warp + L, lvl_q, p, P);
      if (done) continue;
    }
    if (s > 0 && h <= 128 && kp <= 128) compact_merge<4>(list, cap, kp, s, n0, n1, cf, cb, srt, lane, dst, thr_g_warp + L, lvl_q, p, P);
    else if (s > 0 && h <= 256) compact_merge<8>(list, cap, kp, s, n0, n1, cf, cb, srt, lane, dst, thr_g_warp + L, lvl_q, p, P);
    else if (n0 + n1 <= 256) compact_one<8>(list, cap, kp, n0, n1, cf, cb, srt, lane, dst, thr_g_warp + L, lvl_q, p, P);
    else compact_one<16>(list, cap, kp, n0, n1, cf, cb, srt, lane, dst, thr_g_warp + L, lvl_q, p, P);
  }
}

// every other set bit of `m`, starting with set bit number `which` (0 or 1): splits compaction work between the
// two epilogue warps of a quarter
__device__ __forceinline__ uint32_t alternate_bits(uint32_t m, int which) {
  uint32_t out = 0;
  int k = 0;
  while (m) {
    const uint32_t low = m & (0u - m);
    if ((k & 1) == which) out |= low;
    m ^= low;
    ++k;
  }
  return out;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));     // sm_100 3-input max; NaN inputs are skipped
  return d;
}

__device__ __forceinline__ void mul2(uint32_t a0, uint32_t a1, float b0, float b1, float& d0, float& d1) {
  unsigned long long a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

template <int MODE>
__device__ __forceinline__ float score_of(float dot, float s) {
  if constexpr (MODE == MODE_COS) return dot * s;
  else if constexpr (MODE == MODE_ABSCOS) return fabsf(dot) * s;
  else return fmaf(2.0f, dot, -s);          // -(|x|^2 - 2 q.x): larger = closer
}

// ------------------------------------------------------------------------------------ main kernel
// NT = 1: bf16 store, one MMA pass per k-block (A = the query tile, resident; B = database rows, streamed).
// NT = 3: fp32 store as an error-compensated bf16 split x = hi + lo (hi = bf16(x), lo = bf16(x - hi), 16 mantissa
//         bits together): q.x ~= q_hi.x_hi + q_hi.x_lo + q_lo.x_hi, all three accumulated into the same TMEM
//         accumulator.  q_hi is the resident A operand; x_hi, x_lo and this CTA's q_lo k-block take one ring slot
//         each (the same 16 KB), so a k-block costs three slots and twelve MMAs and the bytes staged per MMA are the
//         same as in the bf16 kernel.  tmA2 / tmB2 are the tensor maps of the lo planes (NT = 1: unused copies).
// ARES = true:  D <= 512, the query tile (q / q_hi) is resident in shared memory for the whole unit (above).
// ARES = false: any D (host limit 2048): nothing is resident, the query k-blocks ride the ring next to the database
//         k-blocks (bf16: x, q = 2 slots per k-block; fp32: x_hi, q_hi, x_lo, q_lo = 4 slots).  Twice the L2 -> SM bytes
//         per FLOP, so this variant is L2-bandwidth bound well below the resident kernel - and still tens of times
//         faster than the CUDA-core scan those shapes used to fall back to.
template <int MODE, int NCTA, int NT, bool ARES>
__global__ void __launch_bounds__(THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2, const Args a) {
  static_assert(NT == 1 || (NT == 3 && NCTA == 2), "the split mode is built for the CTA-pair kernel only");
  static_assert(ARES || NCTA == 2, "the A-streamed variant is built for the CTA-pair kernel only");
  // NCTA == 2: the CTA pair of a cluster works on two query tiles and shares every database tile: each CTA stages
  // HALF of the tile's rows (16 KB per k-block instead of 32 KB, so the ring is 6 deep), the leader issues
  // tcgen05.mma.cta_group::2 (M = 256 across the pair) and each SM's tensor core reads both halves: shared-memory
  // traffic per FLOP drops by a third, which is what bounds the 1-CTA kernel.
  constexpr int NST = !ARES ? STREAM_SLOTS : (NCTA == 2 ? 2 * B_STAGES : B_STAGES);          // ring depth
  constexpr int STAGE_BYTES = B_STAGE_BYTES / NCTA;                  // bytes of one k-block staged per CTA
  constexpr int BN_CTA = BN / NCTA;                                  // database rows staged per CTA
  constexpr int A_BYTES = ARES ? SMEM_A : 0;
  constexpr int RING_BYTES = NST * STAGE_BYTES;
  constexpr int BARS_BYTES = ARES ? SMEM_BARS : SMEM_BARS_STREAM;
  constexpr int SLOTS = ARES ? NT : (NT == 1 ? 2 : 4);               // ring slots per k-block
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  float* sScale = reinterpret_cast<float*>(smem + A_BYTES + RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + RING_BYTES + SMEM_SCALE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2 + 4 * NBUF);
  unsigned short* cnt_s = reinterpret_cast<unsigned short*>(smem + A_BYTES + RING_BYTES + SMEM_SCALE + BARS_BYTES);

  const uint32_t bar0 = smem_u32(bars);
  auto B_FULL = [&](int s) { return bar0 + 8u * s; };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
  const uint32_t A_FULL = bar0 + 8u * (2 * NST), A_EMPTY = bar0 + 8u * (2 * NST + 1);
  auto T_FULL = [&](int b) { return bar0 + 8u * (2 * NST + 2 + b); };
  auto T_EMPTY = [&](int b) { return bar0 + 8u * (2 * NST + 2 + NBUF + b); };     // local: this CTA's epilogue released buffer b
  auto S_FULL = [&](int b) { return bar0 + 8u * (2 * NST + 2 + 2 * NBUF + b); };
  auto TE_MMA = [&](int b) { return bar0 + 8u * (2 * NST + 2 + 3 * NBUF + b); };  // leader: every epilogue warp of the cluster released b
  static_assert((2 * NST + 2 + 4 * NBUF) * 8 + 8 <= BARS_BYTES, "barrier area");
  const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / NCTA, nclusters = gridDim.x / NCTA;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if ((smem_u32(smem) & 1023u) != 0) __trap();     // SWIZZLE_128B tiles need 1024-byte alignment

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (NT == 3) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
    mbar_init(A_FULL, 1);
    mbar_init(A_EMPTY, 1);
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(T_FULL(b), 1); mbar_init(T_EMPTY(b), 8); mbar_init(S_FULL(b), 1); mbar_init(TE_MMA(b), 8 * NCTA);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();       // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_units = a.num_qgroups * a.P;
  // Unit order.  Partition-major: all clusters stream the same database range together, so a tile is read from DRAM once
  // per round and served from L2 to everybody else.  Query-major (-DGEMM_QUERY_MAJOR=1: the P partitions of a query
  // group run at the same time and pool their candidates through the shared thresholds sooner) was measured 39 % SLOWER
  // on the headline shape (12.28 vs 8.82 ms, also at k = 10): P concurrent streams at a fixed 91 MB stride lose the L2
  // sharing and more than eat the shorter cold phase.  Kept as a compile-time switch for the record.
#ifndef GEMM_QUERY_MAJOR
#define GEMM_QUERY_MAJOR 0
#endif
  constexpr bool part_major = GEMM_QUERY_MAJOR == 0;
  auto unit_qgroup = [&](int unit) { return part_major ? unit % a.num_qgroups : unit / a.P; };
  auto unit_part = [&](int unit) { return part_major ? unit / a.num_qgroups : unit % a.P; };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t kiter = 0, uiter = 0, titer = 0;
      for (int unit = cluster_id; unit < num_units; unit += nclusters, ++uiter) {
        const int qt = unit_qgroup(unit) * NCTA + int(rank), p = unit_part(unit);
        long long t0 = DBG_T0();
        if constexpr (ARES) {
          mbar_wait(A_EMPTY, (uiter & 1) ^ 1);                   // previous unit's MMAs are done with A
          DBG_ADD(1, t0);
          if (rank == 0) mbar_expect_tx(A_FULL, uint32_t(NCTA) * uint32_t(a.num_kb) * A_KB_BYTES);
          for (int kb = 0; kb < a.num_kb; ++kb) {
            if constexpr (NCTA == 2) tma_load_2d_2cta(smem_u32(sA + kb * A_KB_BYTES), &tmA, A_FULL, kb * BK, qt * BM);
            else tma_load_2d(smem_u32(sA + kb * A_KB_BYTES), &tmA, A_FULL, kb * BK, qt * BM);
          }
        }
        const int tile0 = p * a.tiles_per_part;
        const int tile1 = min(a.total_tiles, tile0 + a.tiles_per_part);
        for (int t = tile0; t < tile1; ++t, ++titer) {
          for (int kb = 0; kb < a.num_kb; ++kb) {
#pragma unroll
            for (int part = 0; part < SLOTS; ++part, ++kiter) {       // resident A: x [, x_lo, q_lo]; streamed A: x, q | x_hi, q_hi, x_lo, q_lo
              const int s = kiter % NST;
              const uint32_t ph = (kiter / NST) & 1;
              t0 = DBG_T0();
              mbar_wait(B_EMPTY(s), ph ^ 1);
              DBG_ADD(0, t0);
              if (rank == 0) mbar_expect_tx(B_FULL(s), B_STAGE_BYTES);                  // both halves land on the leader's barrier
              const uint32_t dst = smem_u32(sB + s * STAGE_BYTES);
              if constexpr (!ARES) {
                const bool is_q = (part & 1) != 0;                    // odd slots carry this CTA's query k-block
                const CUtensorMap* map = is_q ? (part == 1 ? &tmA : &tmA2) : (part == 0 ? &tmB : &tmB2);
                if (is_q) tma_load_2d_2cta(dst, map, B_FULL(s), kb * BK, qt * BM);
                else tma_load_2d_2cta(dst, map, B_FULL(s), kb * BK, t * BN + int(rank) * BN_CTA);
              } else if constexpr (NT == 3) {
                if (part == 2) tma_load_2d_2cta(dst, &tmA2, B_FULL(s), kb * BK, qt * BM);
                else tma_load_2d_2cta(dst, part == 0 ? &tmB : &tmB2, B_FULL(s), kb * BK, t * BN + int(rank) * BN_CTA);
              } else if constexpr (NCTA == 2) {
                tma_load_2d_2cta(dst, &tmB, B_FULL(s), kb * BK, t * BN + int(rank) * BN_CTA);
              } else {
                tma_load_2d(dst, &tmB, B_FULL(s), kb * BK, t * BN);
              }
            }
            // pull the same k-block of the tile two ahead into L2 (the shared-memory ring is only a few k-blocks deep)
            if ((a.opt & 1) && t + 2 < tile1) {
              tma_prefetch_l2_2d(&tmB, kb * BK, (t + 2) * BN);
              if constexpr (NT == 3) tma_prefetch_l2_2d(&tmB2, kb * BK, (t + 2) * BN);
            }
          }
          // per-column scale of this tile: its buffer is free once the epilogue released accumulator `buf`
          // two tiles ago (already true by now in steady state: the MMAs of this tile are running)
          const int buf = titer % NBUF;
          t0 = DBG_T0();
          mbar_wait(T_EMPTY(buf), ((titer / NBUF) & 1) ^ 1);
          DBG_ADD(2, t0);
          mbar_expect_tx(S_FULL(buf), BN * 4);
          bulk_load_1d(smem_u32(sScale + buf * BN), a.colscale + int64_t(t) * BN, BN * 4, S_FULL(buf));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      uint32_t kiter = 0, titer = 0, uiter = 0;
      for (int unit = cluster_id; unit < num_units; unit += nclusters, ++uiter) {
        const int p = unit_part(unit);
        const int tile0 = p * a.tiles_per_part;
        const int tile1 = min(a.total_tiles, tile0 + a.tiles_per_part);
        if (a.dbg) {                                             // MMA-issue time stamps per round: slot 15 = round 0, 11 = later
          const long long now = clock64();
          if (uiter == 1) a.dbg[blockIdx.x * 16 + 15] = (unsigned long long)(now - (long long)a.dbg[blockIdx.x * 16 + 14]);
          if (uiter == 0) a.dbg[blockIdx.x * 16 + 14] = (unsigned long long)now;
        }
        long long t0 = DBG_T0();
        if constexpr (ARES) {
          mbar_wait(A_FULL, uiter & 1);
          DBG_ADD(5, t0);
          tc_fence_after();
        }
        for (int t = tile0; t < tile1; ++t, ++titer) {
          const int buf = titer % NBUF;
          t0 = DBG_T0();
          mbar_wait(TE_MMA(buf), ((titer / NBUF) & 1) ^ 1);      // every epilogue warp of the cluster drained this accumulator
          DBG_ADD(4, t0);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + uint32_t(buf * BN);
          for (int kb = 0; kb < a.num_kb; ++kb) {
            const uint32_t a_addr = smem_u32(sA + kb * A_KB_BYTES);
            if constexpr (!ARES) {
              // streamed A: slot order x[_hi], q[_hi] (, x_lo, q_lo); every slot has its own full / empty barrier
              uint32_t sl[SLOTS], phs[SLOTS];
              int sidx[SLOTS];
#pragma unroll
              for (int j = 0; j < SLOTS; ++j) {
                sidx[j] = int((kiter + j) % NST);
                phs[j] = ((kiter + j) / NST) & 1;
                sl[j] = smem_u32(sB + sidx[j] * STAGE_BYTES);
              }
              kiter += SLOTS;
              auto wait_slot = [&](int j) { t0 = DBG_T0(); mbar_wait(B_FULL(sidx[j]), phs[j]); DBG_ADD(3, t0); };
              auto mma4 = [&](uint32_t aa, uint32_t bb, bool first) {
#pragma unroll
                for (int k4 = 0; k4 < BK / UMMA_K; ++k4)
                  tc_mma_bf16_2cta(tmem_d, make_smem_desc(aa + k4 * UMMA_K * 2), make_smem_desc(bb + k4 * UMMA_K * 2), kInstrDesc2,
                                   (first && (kb | k4) == 0) ? 0u : 1u);
              };
              wait_slot(0);
              wait_slot(1);
              tc_fence_after();
              mma4(sl[1], sl[0], true);                                // q[_hi] . x[_hi]
              if constexpr (NT == 3) {
                wait_slot(2);
                tc_fence_after();
                mma4(sl[1], sl[2], false);                             // q_hi . x_lo
                tc_commit_2cta(B_EMPTY(sidx[2]));
                tc_commit_2cta(B_EMPTY(sidx[1]));
                wait_slot(3);
                tc_fence_after();
                mma4(sl[3], sl[0], false);                             // q_lo . x_hi
                tc_commit_2cta(B_EMPTY(sidx[0]));
                tc_commit_2cta(B_EMPTY(sidx[3]));
              } else {
                tc_commit_2cta(B_EMPTY(sidx[0]));
                tc_commit_2cta(B_EMPTY(sidx[1]));
              }
            } else if constexpr (NT == 3) {
              // slots: s0 = x_hi, s1 = x_lo, s2 = q_lo of this k-block
              const int s0 = kiter % NST, s1 = (kiter + 1) % NST, s2 = (kiter + 2) % NST;
              const uint32_t ph0 = (kiter / NST) & 1, ph1 = ((kiter + 1) / NST) & 1, ph2 = ((kiter + 2) / NST) & 1;
              kiter += 3;
              const uint32_t xhi = smem_u32(sB + s0 * STAGE_BYTES), xlo = smem_u32(sB + s1 * STAGE_BYTES), qlo = smem_u32(sB + s2 * STAGE_BYTES);
              t0 = DBG_T0();
              mbar_wait(B_FULL(s0), ph0);
              DBG_ADD(3, t0);
              tc_fence_after();
#pragma unroll
              for (int k4 = 0; k4 < BK / UMMA_K; ++k4)
                tc_mma_bf16_2cta(tmem_d, make_smem_desc(a_addr + k4 * UMMA_K * 2), make_smem_desc(xhi + k4 * UMMA_K * 2), kInstrDesc2,
                                 (kb | k4) != 0 ? 1u : 0u);
              t0 = DBG_T0();
              mbar_wait(B_FULL(s1), ph1);
              DBG_ADD(3, t0);
              tc_fence_after();
#pragma unroll
              for (int k4 = 0; k4 < BK / UMMA_K; ++k4)
                tc_mma_bf16_2cta(tmem_d, make_smem_desc(a_addr + k4 * UMMA_K * 2), make_smem_desc(xlo + k4 * UMMA_K * 2), kInstrDesc2, 1u);
              tc_commit_2cta(B_EMPTY(s1));
              t0 = DBG_T0();
              mbar_wait(B_FULL(s2), ph2);
              DBG_ADD(3, t0);
              tc_fence_after();
#pragma unroll
              for (int k4 = 0; k4 < BK / UMMA_K; ++k4)
                tc_mma_bf16_2cta(tmem_d, make_smem_desc(qlo + k4 * UMMA_K * 2), make_smem_desc(xhi + k4 * UMMA_K * 2), kInstrDesc2, 1u);
              tc_commit_2cta(B_EMPTY(s0));
              tc_commit_2cta(B_EMPTY(s2));
            } else {
              const int s = kiter % NST;
              const uint32_t ph = (kiter / NST) & 1;
              ++kiter;
              t0 = DBG_T0();
              mbar_wait(B_FULL(s), ph);
              DBG_ADD(3, t0);
              tc_fence_after();
              const uint32_t b_addr = smem_u32(sB + s * STAGE_BYTES);
              t0 = DBG_T0();
#pragma unroll
              for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
                if constexpr (NCTA == 2)
                  tc_mma_bf16_2cta(tmem_d, make_smem_desc(a_addr + k4 * UMMA_K * 2), make_smem_desc(b_addr + k4 * UMMA_K * 2),
                                   kInstrDesc2, (kb | k4) != 0 ? 1u : 0u);
                else
                  tc_mma_bf16(tmem_d, make_smem_desc(a_addr + k4 * UMMA_K * 2), make_smem_desc(b_addr + k4 * UMMA_K * 2),
                              kInstrDesc, (kb | k4) != 0 ? 1u : 0u);
              }
              if constexpr (NCTA == 2) tc_commit_2cta(B_EMPTY(s)); else tc_commit(B_EMPTY(s));   // stage reusable once these MMAs retire
              DBG_ADD(6, t0);
            }
          }
          if constexpr (NCTA == 2) tc_commit_2cta(T_FULL(buf)); else tc_commit(T_FULL(buf));    // accumulator complete -> epilogue
        }
        if constexpr (ARES) { if constexpr (NCTA == 2) tc_commit_2cta(A_EMPTY); else tc_commit(A_EMPTY); }
        if (a.dbg && unit + nclusters >= num_units)              // total issue time of this cluster (slot 11)
          a.dbg[blockIdx.x * 16 + 11] = (unsigned long long)(clock64() - (long long)a.dbg[blockIdx.x * 16 + 14]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // Two warps per TMEM lane quarter (warp % 4 selects the lanes a warp may read): warp `half` takes the chunks
    // 2i + half of every tile, so each scheduler always has a second epilogue warp to issue from while the other
    // waits on a dependency, a branch or a TMEM load.  The two threads that serve one query row share ONE candidate
    // list (slots handed out by a shared-memory atomic counter) and meet at a 64-thread named barrier after every
    // chunk pair, where a full list is compacted (the lists to compact are split between the two warps).
    const int quarter = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;
    const int pair_bar = 1 + quarter;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    uint64_t* warp_lists = a.cand + (size_t(blockIdx.x) * BM + quarter * 32) * a.cap;
    uint64_t* mylist = warp_lists + size_t(lane) * a.cap;
    unsigned short* cnt_q = cnt_s + quarter * 96;          // [32 front counts | 32 back counts | 32 sorted-prefix lengths]
    volatile unsigned short* mycnt = cnt_q + half * 32 + lane;
    const int64_t lstep = half ? -1 : 1;                   // warp 0 fills the list upwards from 0, warp 1 downwards from cap-1
    uint64_t* const lbase = half ? mylist + a.cap - 1 : mylist;
    const uint32_t tmem_lane = uint32_t(quarter * 32) << 16;
    const bool dbg_me = DBG_ON && warp == 4 && lane == 0;
    // tile mode: lists are checked once per tile and compacted when longer than tile_limit (<= 256 keys: the narrow
    // sorter is enough in the common case); a tile adds at most BN entries, so tile_limit + BN must fit the list
    const int tile_limit = max(a.kp + 64, 192);
    const bool tile_mode = tile_limit + BN <= a.cap;
    float thr = 0.f;
    int cnt = 0;
    uint32_t* thr_g_warp = a.thr_g;
    uint32_t* lvl_warp = nullptr;
    int cur_p = 0;
    // publish my end's counter, meet the partner warp, compact every list longer than `limit` (lists split between
    // the two warps), pick up the new thresholds
    auto check_and_compact = [&](int limit, long long& tcomp) {
      *mycnt = (unsigned short)cnt;
      pair_sync();
      const int total = cnt + int(*reinterpret_cast<volatile unsigned short*>(cnt_q + (half ^ 1) * 32 + lane));
      const uint32_t full = __ballot_sync(0xffffffffu, total > limit);
      if (full) {
        const long long tc0 = dbg_me ? clock64() : 0ll;
        const uint32_t mine = alternate_bits(full, half);
        warp_compact(warp_lists, a.cap, a.kp, cnt_q, mine, lane, nullptr, 0, 32, thr_g_warp, lvl_warp, cur_p, a.P, !(a.opt & 8));
        pair_sync();
        if ((full >> lane) & 1) {
          cnt = int(*mycnt);                                                // kp (or fewer) at the front, 0 at the back
          if (int(*reinterpret_cast<volatile unsigned short*>(cnt_q + lane)) >= a.kp)
            thr = fmaxf(thr, -key_rank(*reinterpret_cast<volatile uint64_t*>(mylist + a.kp - 1)));   // accept only v > thr
        }
        if (dbg_me) { tcomp += clock64() - tc0; a.dbg[blockIdx.x * 16 + 12] += __popc(full); }
      }
    };
    uint32_t titer = 0;
    int dbg_round = -1;
    for (int unit = cluster_id; unit < num_units; unit += nclusters) {
      ++dbg_round;
      const int qt = unit_qgroup(unit) * NCTA + int(rank), p = unit_part(unit);
      const int q = qt * BM + row;
      thr = q < a.nq ? -INFINITY : INFINITY;                   // padding rows of the last query tile accept nothing
      cnt = 0;                                                 // my end of the list (register; published at every barrier)
      *mycnt = 0;
      if (half == 0) cnt_q[64 + lane] = 0;                     // nothing sorted yet
      pair_sync();
      const int tile0 = p * a.tiles_per_part;
      const int tile1 = min(a.total_tiles, tile0 + a.tiles_per_part);
      thr_g_warp = a.thr_g + qt * BM + quarter * 32;
      lvl_warp = a.lvl ? a.lvl + size_t(qt * BM + quarter * 32) * kLevels * a.P : nullptr;
      cur_p = p;
      import_threshold(thr_g_warp + lane, thr);
      for (int t = tile0; t < tile1; ++t, ++titer) {
        const int buf = titer % NBUF;
        if (((t - tile0) & 7) == 7) import_threshold(thr_g_warp + lane, thr);      // other units' published thresholds
        long long tcomp = 0;
        if (tile_mode) {
          // compaction check once per tile, BEFORE waiting for the accumulator: both TMEM buffers are released, so the
          // MMAs of the next two tiles run while the lists are sorted.  A tile adds at most 256 entries per list.
          check_and_compact(tile_limit, tcomp);
          if (dbg_me) { a.dbg[blockIdx.x * 16 + 9] += (unsigned long long)tcomp; tcomp = 0; }
        }
        long long t0 = dbg_me ? clock64() : 0ll;
        mbar_wait(S_FULL(buf), (titer / NBUF) & 1);
        mbar_wait(T_FULL(buf), (titer / NBUF) & 1);
        if (dbg_me) { a.dbg[blockIdx.x * 16 + 7] += (unsigned long long)(clock64() - t0); t0 = clock64(); }
        tc_fence_after();
        const float* sc = sScale + buf * BN;
        const uint32_t base_idx = uint32_t(t) * BN;
        const uint32_t tbase = tmem_base + tmem_lane + uint32_t(buf * BN);
#pragma unroll 1
        for (int pair = 0; pair < BN / 64; ++pair) {
          const int chunk = pair * 2 + half;
          uint32_t r[32];
          tc_ld32(tbase + uint32_t(chunk * 32), r);
          tc_wait_ld();
          // branch-free pre-filter: score all 32 columns and reduce with 3-input max (max.f32 skips NaN = masked
          // columns).  Only a thread whose 8-column group maximum beats its threshold touches that group, and it
          // extracts the hits by arg-max (select chains, no per-element branches).
          float v[32];
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 s4 = *reinterpret_cast<const float4*>(sc + chunk * 32 + c4 * 4);
#ifndef GEMM_SCALAR_SCORE
            if constexpr (MODE == MODE_COS) {
              // two columns per instruction (mul.rn.f32x2 = FMUL2: same bits as the scalar product).  A scalar FMUL takes a
              // cycle of both halves of the FP32 pipe and so competes with the FMNMX3 of the max reduction below; the
              // packed one stays on the heavy half (profiles/r2_pipe_overlap_probe.md).
              mul2(r[c4 * 4 + 0], r[c4 * 4 + 1], s4.x, s4.y, v[c4 * 4 + 0], v[c4 * 4 + 1]);
              mul2(r[c4 * 4 + 2], r[c4 * 4 + 3], s4.z, s4.w, v[c4 * 4 + 2], v[c4 * 4 + 3]);
              continue;
            }
#endif
            v[c4 * 4 + 0] = score_of<MODE>(__uint_as_float(r[c4 * 4 + 0]), s4.x);
            v[c4 * 4 + 1] = score_of<MODE>(__uint_as_float(r[c4 * 4 + 1]), s4.y);
            v[c4 * 4 + 2] = score_of<MODE>(__uint_as_float(r[c4 * 4 + 2]), s4.z);
            v[c4 * 4 + 3] = score_of<MODE>(__uint_as_float(r[c4 * 4 + 3]), s4.w);
          }
          float gm[4];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const float* w = v + g8 * 8;
            gm[g8] = fmax3(fmax3(w[0], w[1], w[2]), fmax3(w[3], w[4], w[5]), fmaxf(w[6], w[7]));
          }
          const float mx = fmaxf(fmax3(gm[0], gm[1], gm[2]), gm[3]);
          if (mx > thr) {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float* w = v + g8 * 8;
              float gmax = gm[g8];
              while (gmax > thr) {
                int idx = 7;
#pragma unroll
                for (int j = 6; j >= 0; --j) idx = (w[j] == gmax) ? j : idx;       // lowest column among equals first
                lbase[lstep * cnt] = make_key(-gmax, base_idx + uint32_t(chunk * 32 + g8 * 8 + idx));
                ++cnt;
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = (j == idx) ? -INFINITY : w[j];
                gmax = fmax3(fmax3(w[0], w[1], w[2]), fmax3(w[3], w[4], w[5]), fmaxf(w[6], w[7]));
              }
            }
          }
          // large k' (not enough room for a whole tile of candidates): check after every chunk pair instead
          if (!tile_mode) check_and_compact(a.cap - 64, tcomp);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(T_EMPTY(buf));                               // my CTA's scale buffer / accumulator slot is free
          if (NCTA == 2 && rank != 0) mbar_arrive_remote(TE_MMA(buf), 0);
          else mbar_arrive(TE_MMA(buf));                           // the leader's MMA issuer may overwrite the accumulator
        }
        if (dbg_me) {
          a.dbg[blockIdx.x * 16 + 8] += (unsigned long long)(clock64() - t0 - tcomp);
          a.dbg[blockIdx.x * 16 + 9] += (unsigned long long)tcomp;
          a.dbg[blockIdx.x * 16 + 13] += 1;
        }
      }
      // end of unit: emit the best kp keys of every valid query row of this quarter (rows split between the warps)
      const long long te0 = dbg_me ? clock64() : 0ll;
      *mycnt = (unsigned short)cnt;
      pair_sync();
      const int q0 = qt * BM + quarter * 32;
      const int valid = min(32, a.nq - q0);
      if (valid > 0) {
        uint64_t* out = a.partial + (int64_t(q0) * a.P + p) * a.kp;
        const uint32_t mine = half ? 0xffff0000u : 0x0000ffffu;
        warp_compact(warp_lists, a.cap, a.kp, cnt_q, mine, lane, out, int64_t(a.P) * a.kp, valid, thr_g_warp, lvl_warp, cur_p, a.P);
      }
      pair_sync();                                             // lists and counters may be reused by the next unit
      if (dbg_me) a.dbg[blockIdx.x * 16 + 10] += (unsigned long long)(clock64() - te0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();       // nobody exits while the peer may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if constexpr (NCTA == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------ side kernels
// 8 consecutive elements of a row as fp32 (16-byte loads: D % 8 == 0 and 16-byte aligned bases on this path)
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 w = __ldg(reinterpret_cast<const uint4*>(p));
  f[0] = bf16_lo(w.x); f[1] = bf16_hi(w.x); f[2] = bf16_lo(w.y); f[3] = bf16_hi(w.y);
  f[4] = bf16_lo(w.z); f[5] = bf16_hi(w.z); f[6] = bf16_lo(w.w); f[7] = bf16_hi(w.w);
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return uint32_t(__bfloat16_as_ushort(lo)) | (uint32_t(__bfloat16_as_ushort(hi)) << 16);
}

// Per-row state of the B operand, computed ONCE per store (b200ir_index_build) or per call when no index is given:
//   rnorm[i] = 1/|x_i| (0 for a zero row: cos := 0, geometric_metrics.py:16-17), sqnorm[i] = |x_i|^2, NaN past row N
//   (NaN scores vanish in the epilogue's max reduction), *maxsq = max_i |x_i|^2 (ordered bits; bounds the L2
//   certificate), and for fp32 rows the bf16 planes hi = bf16(x), lo = bf16(x - hi) of the split described at the kernel.
// One warp per row.
template <typename T>
__global__ void __launch_bounds__(256)
index_rows_kernel(const T* __restrict__ X, int64_t N, int64_t N_pad, int D, float* __restrict__ rnorm, float* __restrict__ sqnorm,
                  unsigned int* __restrict__ maxsq, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  __shared__ float blockmax[8];
  float ss = 0.f;
  if (row < N) {
    const T* x = X + row * D;
    for (int d = lane * 8; d < D; d += 256) {
      float f[8];
      load8<T>(x + d, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
      if (hi != nullptr) {
        __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          h[e] = __float2bfloat16_rn(f[e]);
          l[e] = __float2bfloat16_rn(f[e] - __bfloat162float(h[e]));
        }
        *reinterpret_cast<uint4*>(hi + row * D + d) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
        *reinterpret_cast<uint4*>(lo + row * D + d) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (rnorm != nullptr && row < N_pad && lane == 0) {
    const float nan = __int_as_float(0x7fc00000);
    rnorm[row] = row < N ? (ss > 0.f ? 1.0f / sqrtf(ss) : 0.f) : nan;
    sqnorm[row] = row < N ? ss : nan;
  }
  if (maxsq != nullptr) {
    if (lane == 0) blockmax[threadIdx.x >> 5] = (row < N && ss == ss) ? ss : 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int w = 0; w < 8; ++w) m = fmaxf(m, blockmax[w]);
      if (m > 0.f) atomicMax(maxsq, __float_as_uint(m));            // non-negative floats order like their bit patterns
    }
  }
}

// The same 8 elements from a row staged in shared memory (re-rank ring of gemm_finalize_kernel)
template <typename T> __device__ __forceinline__ void load8s(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8s<float>(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8s<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(w.x); f[1] = bf16_hi(w.x); f[2] = bf16_lo(w.y); f[3] = bf16_hi(w.y);
  f[4] = bf16_lo(w.z); f[5] = bf16_hi(w.z); f[6] = bf16_lo(w.w); f[7] = bf16_hi(w.w);
}

// Re-rank ring of gemm_finalize_kernel: every warp owns FIN_RING_KB KB of shared memory (at least two rows) into which
// the candidate rows arrive by cp.async.bulk, one mbarrier per slot.  Measured on the headline (10k x 1M x 512 bf16,
// k' = 107, same-box A/B of the re-rank kernel): 16 KB 0.61 ms, 8 KB 0.50 ms, 4 KB 0.45 ms, 2 KB 0.45 ms against 0.57 ms
// for the lane loads behind an L2 prefetch (-DFIN_RING_KB=0 still builds that path): resident warps (8 CTAs per SM at
// 64 registers and <= 25 KB) matter more than rows in flight, so the ring is kept shallow.
#ifndef FIN_RING_KB
#define FIN_RING_KB 4
#endif
template <typename T, int QJ> struct FinRing {
  static constexpr int kSlotBytes = QJ * 256 * int(sizeof(T));            // widest row of the instantiation
  static constexpr int kWant = FIN_RING_KB * 1024;
  static constexpr int kBytes = kWant >= 2 * kSlotBytes ? kWant : 2 * kSlotBytes;
  static constexpr int kSlots = kBytes / kSlotBytes > 16 ? 16 : kBytes / kSlotBytes;   // a power of two, 2..16
};

// A-posteriori exactness certificate (see gemm_finalize_kernel) and the list of queries that failed it
struct Certify {
  float u_eff;                 // |computed dot - exact dot| <= u_eff * |q| * |x| on the tensor-core pass
  const unsigned int* maxsq;   // max_i |x_i|^2 of the store
  int* fb_count;               // number of uncertified queries (zeroed per call), or nullptr: no certificate
  int* fb_list;                // [nq] their ids
};

// One warp per query: merge the P partition lists, re-rank the kp candidates with exact fp32 arithmetic on the
// original rows (direct dot / direct sum of squared differences: no split, no cancellation), write the k winners.
//
// Certificate.  Every row that is NOT among the kp candidates was dropped against kp rows whose tensor-core score was
// at least as good, so its tensor-core rank value is >= a_last, the rank value of the worst candidate kept.  With
// |dot~ - dot| <= u_eff |q||x| its exact score is bounded; if the k-th best EXACT score among the candidates is
// strictly better than that bound, no dropped row can enter or tie the top-k and the result equals the exact scan's.
// Otherwise (near-ties denser than the kp - k margin, duplicates, huge-norm outliers) the query is appended to
// fb_list and re-done by the exact CUDA-core scan.
// QJ: 256-element slices of a row each lane keeps in registers (2: D <= 512, 8: D <= 2048)
template <int E, typename T, int QJ>
__global__ void __launch_bounds__(128)
gemm_finalize_kernel(const uint64_t* __restrict__ partial, const uint32_t* __restrict__ thr_g, const T* __restrict__ Q,
                     const T* __restrict__ X, const float* __restrict__ xsq, int nq, int D, int P, int kp, int k, int mode,
                     int rerank, MetricParams mp, int64_t index_offset, Certify cert,
                     float* __restrict__ out_score, int64_t* __restrict__ out_idx) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const int64_t per_query = int64_t(P) * kp;
  const uint64_t* src = partial + int64_t(q) * per_query;
  // Keys worse than the final published threshold of this query (the k'-th best of some partition) cannot be
  // among the global k' best: drop them while streaming the P lists through the sorter.  Survivors are compacted
  // with ballot prefix sums, so usually one or two sort rounds are enough instead of P * kp / (32 E - kp).
  const uint64_t limit = (uint64_t(thr_g[q]) << 32) | 0xffffffffull;       // 0xffffffff.. = none published: keep all
  __shared__ uint64_t stage_s[4][32 * E];
  uint64_t* stage = stage_s[threadIdx.x >> 5];
#if FIN_RING_KB > 0
  using Ring = FinRing<T, QJ>;
  constexpr int NS = Ring::kSlots;
  extern __shared__ __align__(128) unsigned char ring_s[];
  __shared__ __align__(8) uint64_t ring_bar_s[4][16];
  const unsigned char* ring = ring_s + (threadIdx.x >> 5) * Ring::kBytes;
  const uint32_t ring_a = smem_u32(ring);
  const uint32_t bar_a = smem_u32(&ring_bar_s[threadIdx.x >> 5][0]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(bar_a + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
#endif
  uint64_t r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = kKeyInf;
  int kept = 0;            // sorted survivors of earlier rounds occupy register positions [0, kept)
  int fill = 0;            // survivors staged in shared memory for the next round
  constexpr int kAhead = 4;                                                // list chunks in flight per lane (the loop body
  uint64_t ahead[kAhead];                                                  // branches into the sorter, so ptxas keeps one)
  for (int64_t base = 0; base < per_query; base += 32) {
    const int slot = int(base >> 5) % kAhead;
    if (slot == 0) {
#pragma unroll
      for (int a = 0; a < kAhead; ++a) {
        const int64_t sidx = base + 32 * a + lane;
        ahead[a] = sidx < per_query ? __ldcs(reinterpret_cast<const unsigned long long*>(src) + sidx) : kKeyInf;   // read once: do not push prefetched rows out of L2
      }
    }
    uint64_t key = ahead[0];
#pragma unroll
    for (int a = 1; a < kAhead; ++a) if (slot == a) key = ahead[a];
    const bool keep = key <= limit && key != kKeyInf;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) stage[fill + __popc(m & ((1u << lane) - 1))] = key;
    fill += __popc(m);
    if (fill > 32 * E - kept - 32 || base + 32 >= per_query) {
      __syncwarp();
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = lane * E + e;
        if (i >= kept) r[e] = (i - kept) < fill ? stage[i - kept] : kKeyInf;
      }
      warp_sort_shared<E>(r, lane);
      kept = min(kept + fill, kp);
      fill = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) if (lane * E + e >= kept) r[e] = kKeyInf;
      __syncwarp();
    }
  }
  // tensor-core rank value of the worst candidate kept (only meaningful when the list is full: otherwise nothing was
  // ever dropped for this query and every row of the store is a candidate)
  float a_last = 0.f;
  {
    uint64_t kl = kKeyInf;
#pragma unroll
    for (int e = 0; e < E; ++e) if (e == (kp - 1) % E) kl = r[e];
    kl = shfl_u64(kl, (kp - 1) / E);
    a_last = key_rank(kl);
  }
  const bool list_full = kept >= kp;

  // query row in registers: lane holds elements [8*(lane + 32 j), +8), j < QJ
  const T* qrow = Q + int64_t(q) * D;
  float qf[QJ][8];
  float qss = 0.f;
#pragma unroll
  for (int j = 0; j < QJ; ++j) {
    const int d = (lane + 32 * j) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) qf[j][e] = 0.f;
    if (d < D) load8<T>(qrow + d, qf[j]);
#pragma unroll
    for (int e = 0; e < 8; ++e) qss = fmaf(qf[j][e], qf[j][e], qss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qss += __shfl_xor_sync(0xffffffffu, qss, o);
  const float qn = sqrtf(qss);

  const uint32_t row_bytes = uint32_t(D) * uint32_t(sizeof(T));
  // exact sums of one candidate from its row (global memory, or a ring slot in shared memory): q.x (or sum (q-x)^2 for
  // L2) and, when asked for, |x|^2 - direct fp32 sums, lane <-> element layout and butterfly identical in every build
  auto row_sums = [&](const T* xrow, auto in_smem, auto with_xss, float& xss_out, auto&& after_reads) -> float {
    float dot = 0.f, xss = 0.f, d2 = 0.f;
#pragma unroll
    for (int j = 0; j < QJ; ++j) {
      const int d = (lane + 32 * j) * 8;
      if (d < D) {
        float xf[8];
        if constexpr (decltype(in_smem)::value) load8s<T>(xrow + d, xf); else load8<T>(xrow + d, xf);
        if (mode == MODE_L2) {                                     // warp-uniform: only the sums the metric needs
#pragma unroll
          for (int t = 0; t < 8; ++t) { const float df = qf[j][t] - xf[t]; d2 = fmaf(df, df, d2); }
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            dot = fmaf(xf[t], qf[j][t], dot);
            if constexpr (decltype(with_xss)::value) xss = fmaf(xf[t], xf[t], xss);
          }
        }
      }
    }
    if (mode == MODE_L2) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if constexpr (decltype(with_xss)::value) xss += __shfl_xor_sync(0xffffffffu, xss, o);
      }
    }
    after_reads();
    xss_out = xss;
    return mode == MODE_L2 ? d2 : dot;
  };
  auto cosine_rank = [&](float dot, float xss) -> float {          // geometric_metrics.py:12-18 on the exact sums
    const float xn = sqrtf(xss);
    float cs = (qn != 0.f && xn != 0.f) ? dot / (qn * xn) : 0.f;
    if (mode == MODE_ABSCOS) cs = fabsf(cs);
    return -cs;
  };
  auto epilogue_rank = [&](uint64_t key) -> float {               // no re-rank: the epilogue-domain score
    const float v = -key_rank(key);
    if (mode == MODE_L2) return fmaxf(qss - v, 0.f);               // |q|^2 + |x|^2 - 2 q.x
    return -(v * (qn != 0.f ? 1.0f / qn : 0.f));
  };
#if FIN_RING_KB > 0
  // The sorted candidates move to this warp's staging buffer (candidate c at (c % E) * 32 + c / E: conflict-free for the
  // lane-major copy in and out) and are visited by a ROLLED loop: with the keys in registers the loop had to be unrolled
  // E times (static register indices), ~1500 instructions that 32 resident warps walk at different places - ncu's top
  // stall reason of that build was `no_instruction` (instruction fetch), 3.4 of 10.7 stall cycles per issue.
  // The rows come through a per-warp shared-memory ring: lane 0 asks for the row of candidate c with ONE cp.async.bulk
  // (global -> slot c % NS, completing on that slot's mbarrier) NS candidates before the visit; the visit waits on the
  // barrier and reads the row with 16-byte LDS.  No load holds registers while it is in flight (64 registers: 8 CTAs
  // per SM) and every row crosses DRAM -> SM once (the L2 prefetch of the earlier path fetched a third of them twice).
  auto pos = [](int c) { return (c % E) * 32 + c / E; };
  __syncwarp();
#pragma unroll
  for (int e = 0; e < E; ++e) stage[e * 32 + lane] = r[e];
  __syncwarp();
  auto issue_row = [&](int c) {                                    // lane 0 only
    if (c < kp) {
      const uint64_t key = stage[pos(c)];
      if (key != kKeyInf) {
        const int s = c & (NS - 1);
        mbar_expect_tx(bar_a + 8 * s, row_bytes);
        bulk_load_1d(ring_a + uint32_t(s) * uint32_t(Ring::kSlotBytes), X + int64_t(key_index(key)) * D, row_bytes, bar_a + 8 * s);
      }
    }
  };
  if (rerank && lane == 0) {
#pragma unroll
    for (int c0 = 0; c0 < NS; ++c0) issue_row(c0);
  }
  // The visit leaves the raw sum (q.x, or the L2 sum) next to the row index; the per-candidate scalar tail - |x|, the
  // division, the ordered key - is done once per candidate by the lane that owns it after the loop instead of by all 32
  // lanes inside it, and |x|^2 comes from the prepared index (index_rows_kernel sums the row in the same lane layout and
  // butterfly order, so xsq[i] IS the value the visit would compute, bit for bit): 16 FFMA, a butterfly, a square root and
  // a division less per visit.
  int n_done = 0;
#pragma unroll 1
  for (; n_done < kp; ++n_done) {
    const int c = n_done;
    const uint64_t key = stage[pos(c)];
    if (key == kKeyInf) break;                                     // sorted: nothing valid follows (warp-uniform)
    float v;
    if (rerank) {
      mbar_wait(bar_a + 8 * (c & (NS - 1)), uint32_t(c / NS) & 1u);
      const T* xrow = reinterpret_cast<const T*>(ring + (c & (NS - 1)) * Ring::kSlotBytes);
      float unused;
      v = row_sums(xrow, std::true_type{}, std::false_type{}, unused, [&] {
        __syncwarp();                                              // every lane's reads of the slot fed the butterfly:
        if (lane == 0) issue_row(c + NS);                          // the slot takes candidate c + NS
      });
    } else {
      v = epilogue_rank(key);
    }
    __syncwarp();
    if (lane == 0) stage[pos(c)] = (uint64_t(__float_as_uint(v)) << 32) | key_index(key);
  }
  __syncwarp();
  {
    uint32_t idx[E];
    float v[E], xs[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint64_t raw = stage[e * 32 + lane];
      idx[e] = uint32_t(raw);
      v[e] = __uint_as_float(uint32_t(raw >> 32));
      xs[e] = (rerank && mode != MODE_L2 && lane * E + e < n_done) ? __ldg(xsq + idx[e]) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float rank = (rerank && mode != MODE_L2) ? cosine_rank(v[e], xs[e]) : v[e];
      r[e] = lane * E + e < n_done ? make_key(rank, idx[e]) : kKeyInf;
    }
  }
#else
  // The candidates are visited one after the other (two 16-byte loads per lane and row, then a butterfly), i.e. one DRAM
  // latency per candidate.  Lane L owns the E keys of step L: one step ahead it asks the L2 for its E rows with one bulk
  // prefetch each, so the visit finds them on chip.  One step, not two: with 16 rows per warp in flight a third of the
  // prefetched rows were evicted again before their visit (ncu: 1.13 -> 1.70 GB of DRAM reads); the partition lists are
  // read with the streaming hint for the same reason.
  const int nl = (kp + E - 1) / E;
  auto prefetch_rows = [&](int owner) {
    if (rerank && lane == owner && owner < nl) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (owner * E + e < kp && r[e] != kKeyInf) {
          const T* xrow = X + int64_t(key_index(r[e])) * D;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(xrow), "r"(row_bytes) : "memory");
        }
      }
    }
  };
  prefetch_rows(0);
  for (int L = 0; L < nl; ++L) {
    prefetch_rows(L + 1);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = L * E + e;
      const uint64_t key = shfl_u64(r[e], L);
      if (c >= kp || key == kKeyInf) { if (lane == L && c >= kp) r[e] = kKeyInf; continue; }
      const uint32_t idx = key_index(key);
      float rank;
      if (rerank) {
        float xss;
        const float sum = row_sums(X + int64_t(idx) * D, std::false_type{}, std::true_type{}, xss, [] {});
        rank = mode == MODE_L2 ? sum : cosine_rank(sum, xss);
      } else {
        rank = epilogue_rank(key);
      }
      if (lane == L) r[e] = make_key(rank, idx);
    }
  }
#endif
#pragma unroll
  for (int e = 0; e < E; ++e) if (lane * E + e >= kp) r[e] = kKeyInf;
  warp_sort_shared<E>(r, lane);
  if (rerank && cert.fb_count != nullptr && list_full) {
    uint64_t kk = kKeyInf;
#pragma unroll
    for (int e = 0; e < E; ++e) if (e == (k - 1) % E) kk = r[e];
    kk = shfl_u64(kk, (k - 1) / E);
    const float exact_k = key_rank(kk);                          // k-th best exact rank value (smaller = better)
    bool ok;
    if (mode == MODE_L2) {
      // dropped row: d^2 = |q|^2 + (|x|^2 - 2 q.x) >= |q|^2 + a_last - eps
      const float xmax2 = __uint_as_float(*cert.maxsq);
      const float eps = 2.0f * cert.u_eff * qn * sqrtf(xmax2) + 9.6e-7f * (xmax2 + qss);
      ok = exact_k < (qss + a_last) - eps;
    } else {
      // dropped row: cos <= (-a_last) / |q| + u_eff   (a_last = -dot~/|x|, or -|dot~|/|x|)
      const float bound = qn != 0.f ? (-a_last) / qn + cert.u_eff : INFINITY;
      ok = -exact_k > bound;
    }
    if (!ok && lane == 0) cert.fb_list[atomicAdd(cert.fb_count, 1)] = q;
  }
  emit_topk<E, true>(r, lane, k, mp, index_offset, out_score, out_idx, int64_t(q));
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static bool encode_bf16_rows(CUtensorMap* map, const void* base, int64_t rows, int D, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cuuint64_t(D), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(D) * 2};
  const cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Prepared per-store state (b200ir_index_build, or built per call inside the workspace when no index is passed)
struct IndexLayout {
  int64_t N_pad;
  size_t off_rnorm, off_sqnorm, off_max, off_hi, off_lo, total_bytes;
};

static IndexLayout index_layout(int dtype, int64_t N, int D) {
  IndexLayout L{};
  L.N_pad = ceil_div64(N, BN) * BN;
  size_t off = 0;
  L.off_rnorm = off; off += round_up64(size_t(L.N_pad) * 4, 256);
  L.off_sqnorm = off; off += round_up64(size_t(L.N_pad) * 4, 256);
  L.off_max = off; off += 256;
  L.off_hi = off;
  L.off_lo = off;
  if (dtype == B200IR_F32) {
    off += round_up64(size_t(N) * D * 2, 256);
    L.off_lo = off; off += round_up64(size_t(N) * D * 2, 256);
  }
  L.total_bytes = off;
  return L;
}

struct Plan {
  int num_qtiles, num_qgroups, ncta, total_tiles, P, tiles_per_part, kp, cap, grid, num_kb;
  bool ares;                                    // query tile resident (D <= 512) or streamed
  size_t off_thr, off_lvl, off_cand, off_partial, off_fb, off_qhi, off_qlo, off_index, off_fbws, total_bytes;
};

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = kNumSMs;
  }
  return n;
}

// cluster size of the main kernel: 2 (tcgen05 cta_group::2) unless B200IR_GEMM_NCTA=1 (bf16 stores only)
static int gemm_ncta(int dtype) {
  static const int n = (getenv("B200IR_GEMM_NCTA") && atoi(getenv("B200IR_GEMM_NCTA")) == 1) ? 1 : 2;
  return dtype == B200IR_F32 ? 2 : n;
}

static Plan make_plan(int dtype, int64_t nq, int64_t N, int D, int k, int flags, int sms, bool internal_index, size_t fallback_bytes) {
  Plan pl{};
  pl.num_kb = (D + BK - 1) / BK;
  pl.ares = pl.num_kb <= MAX_KB;
  pl.ncta = pl.ares ? gemm_ncta(dtype) : 2;
  pl.num_qtiles = int(ceil_div64(nq, BM));
  pl.num_qgroups = (pl.num_qtiles + pl.ncta - 1) / pl.ncta;
  const int nclusters_max = sms / pl.ncta;
  pl.total_tiles = int(ceil_div64(N, BN));
  const bool rerank = !(flags & B200IR_FLAG_NO_RERANK);
  pl.cap = 512;
  int kp = k;
  if (rerank) { kp = k + (k / 16 > 4 ? k / 16 : 4); kp = (kp + 3) / 4 * 4; }
  if (kp > pl.cap / 2) kp = pl.cap / 2;
  if (kp < k) kp = k;
  pl.kp = kp;
  // partitions: minimise rounds x (tiles per unit + fixed per-unit cost), prefer fewer partitions
  int64_t best_cost = -1;
  int bestP = 1;
  const int maxP = pl.total_tiles < 64 ? pl.total_tiles : 64;
  for (int P = 1; P <= maxP; ++P) {
    const int tpp = (pl.total_tiles + P - 1) / P;
    const int Pe = (pl.total_tiles + tpp - 1) / tpp;
    const int64_t units = int64_t(pl.num_qgroups) * Pe;
    const int64_t rounds = ceil_div64(units, nclusters_max);
    const int64_t cost = rounds * (tpp + 3);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; bestP = P; }
  }
  pl.tiles_per_part = (pl.total_tiles + bestP - 1) / bestP;
  pl.P = (pl.total_tiles + pl.tiles_per_part - 1) / pl.tiles_per_part;
  const int64_t units = int64_t(pl.num_qgroups) * pl.P;
  pl.grid = int(units < nclusters_max ? units : nclusters_max) * pl.ncta;
  size_t off = 0;
  pl.off_thr = off; off += round_up64(size_t(pl.num_qgroups) * pl.ncta * BM * 4, 256);
  pl.off_lvl = off; off += pl.P <= 32 ? round_up64(size_t(pl.num_qgroups) * pl.ncta * BM * kLevels * pl.P * 4, 256) : 0;
  pl.off_cand = off; off += round_up64(size_t(pl.grid) * BM * pl.cap * 8, 256);
  pl.off_partial = off; off += round_up64(size_t(nq) * pl.P * pl.kp * 8, 256);
  pl.off_fb = off; off += round_up64(size_t(nq) * 4 + 256, 256);                        // [count | pad | list]
  pl.off_qhi = pl.off_qlo = off;
  if (dtype == B200IR_F32) {
    off += round_up64(size_t(nq) * D * 2, 256);
    pl.off_qlo = off; off += round_up64(size_t(nq) * D * 2, 256);
  }
  pl.off_index = off;
  if (internal_index) off += index_layout(dtype, N, D).total_bytes;
  pl.off_fbws = off; off += fallback_bytes;
  pl.total_bytes = off;
  return pl;
}

template <typename T>
static cudaError_t launch_index_rows(const void* X, int64_t N, int64_t N_pad, int D, float* rnorm, float* sqnorm, unsigned int* maxsq,
                                     __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
  index_rows_kernel<T><<<unsigned(ceil_div64(N_pad, 8)), 256, 0, st>>>(static_cast<const T*>(X), N, N_pad, D, rnorm, sqnorm, maxsq, hi, lo);
  return cudaGetLastError();
}

static cudaError_t build_index(int dtype, const void* X, int64_t N, int D, unsigned char* index, cudaStream_t st) {
  const IndexLayout L = index_layout(dtype, N, D);
  cudaError_t e = cudaMemsetAsync(index + L.off_max, 0, 256, st);
  if (e != cudaSuccess) return e;
  float* rn = reinterpret_cast<float*>(index + L.off_rnorm);
  float* sq = reinterpret_cast<float*>(index + L.off_sqnorm);
  unsigned int* mx = reinterpret_cast<unsigned int*>(index + L.off_max);
  if (dtype == B200IR_F32)
    return launch_index_rows<float>(X, N, L.N_pad, D, rn, sq, mx, reinterpret_cast<__nv_bfloat16*>(index + L.off_hi),
                                    reinterpret_cast<__nv_bfloat16*>(index + L.off_lo), st);
  return launch_index_rows<__nv_bfloat16>(X, N, L.N_pad, D, rn, sq, mx, nullptr, nullptr, st);
}

}  // namespace gemm

static int clamp_sms() {
  int sms = gemm::num_sms();
  return sms > kNumSMs ? kNumSMs : sms;
}

bool gemm_path_supported(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags) {
  if (dtype != B200IR_BF16 && dtype != B200IR_F32) return false;
  if (!(metric == B200IR_L2 || metric == B200IR_COS_SIM || metric == B200IR_COS_DIST || metric == B200IR_ANGLE)) return false;
  if (D % 8 != 0 || D > gemm::MAX_D_STREAM || D < 16) return false;
  if (nq < 32 || N < 4 * gemm::BN) return false;      // tiny problems stay on the scan path
  if (k > 240) return false;                          // k' = k + k/16 must fit half a candidate list (cap / 2 = 256)
  return true;
}

size_t gemm_fallback_counter_offset(int dtype, int64_t nq, int64_t N, int D, int k, int flags, bool have_index) {
  return gemm::make_plan(dtype, nq, N, D, k, flags, clamp_sms(), !have_index, 0).off_fb;
}

size_t gemm_index_bytes(int dtype, int64_t N, int D) {
  if (N <= 0 || D <= 0 || D % 8 != 0 || D > gemm::MAX_D_STREAM || D < 16) return 0;
  return gemm::index_layout(dtype, N, D).total_bytes;
}

int gemm_index_build(int dtype, const void* X, int64_t N, int D, unsigned char* index, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(index) & 255)) return B200IR_E_ALIGN;
  ProfileScope ps(PT_PREP, st);
  return int(gemm::build_index(dtype, X, N, D, index, st));
}

size_t gemm_workspace_bytes(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags, bool have_index) {
  (void)metric;
  const size_t fb = make_fallback_plan(dtype, nq, N, D, k).total_bytes;
  return gemm::make_plan(dtype, nq, N, D, k, flags, clamp_sms(), !have_index, fb).total_bytes;
}

int run_gemm_topk(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k, int64_t index_offset,
                  int flags, const MetricParams& mp, float* out_score, int64_t* out_idx, unsigned char* ws,
                  const unsigned char* index, cudaStream_t st) {
  using namespace gemm;
  if ((reinterpret_cast<uintptr_t>(Q) & 15) || (reinterpret_cast<uintptr_t>(X) & 15)) return B200IR_E_ALIGN;
  const bool f32 = dtype == B200IR_F32;
  const FallbackPlan fbp = make_fallback_plan(dtype, nq, N, D, k);
  const Plan pl = make_plan(dtype, nq, N, D, k, flags, clamp_sms(), index == nullptr, fbp.total_bytes);
  const int mode = metric == B200IR_L2 ? MODE_L2 : ((flags & B200IR_FLAG_ABS_SCORE) ? MODE_ABSCOS : MODE_COS);
  const IndexLayout IL = index_layout(dtype, N, D);

  {
    ProfileScope ps(PT_PREP, st);
    if (index == nullptr) {                                        // no prepared index: build the per-store state now
      cudaError_t e = build_index(dtype, X, N, D, ws + pl.off_index, st);
      if (e != cudaSuccess) return int(e);
      index = ws + pl.off_index;
    }
    // per-search state: thresholds + level slots (0xff = none published), fallback counter
    cudaError_t em = cudaMemsetAsync(ws + pl.off_thr, 0xff, pl.off_cand - pl.off_thr, st);
    if (em == cudaSuccess) em = cudaMemsetAsync(ws + pl.off_fb, 0, 256, st);
    if (em != cudaSuccess) return int(em);
    if (f32) {                                                     // bf16 hi / lo planes of the query batch
      cudaError_t e = launch_index_rows<float>(Q, nq, nq, D, nullptr, nullptr, nullptr, reinterpret_cast<__nv_bfloat16*>(ws + pl.off_qhi),
                                               reinterpret_cast<__nv_bfloat16*>(ws + pl.off_qlo), st);
      if (e != cudaSuccess) return int(e);
    }
  }

  CUtensorMap tmA, tmA2, tmB, tmB2;
  const void* a_hi = f32 ? static_cast<const void*>(ws + pl.off_qhi) : Q;
  const void* a_lo = f32 ? static_cast<const void*>(ws + pl.off_qlo) : Q;
  const void* b_hi = f32 ? static_cast<const void*>(index + IL.off_hi) : X;
  const void* b_lo = f32 ? static_cast<const void*>(index + IL.off_lo) : X;
  if (!encode_bf16_rows(&tmA, a_hi, nq, D, BM) || !encode_bf16_rows(&tmA2, a_lo, nq, D, BM) ||
      !encode_bf16_rows(&tmB, b_hi, N, D, BN / pl.ncta) || !encode_bf16_rows(&tmB2, b_lo, N, D, BN / pl.ncta))
    return B200IR_E_DEVICE;

  Args a{};
  a.nq = int(nq); a.N = N; a.D = D; a.num_kb = pl.num_kb; a.num_qtiles = pl.num_qtiles; a.num_qgroups = pl.num_qgroups; a.P = pl.P;
  a.tiles_per_part = pl.tiles_per_part; a.total_tiles = pl.total_tiles; a.kp = pl.kp; a.cap = pl.cap;
  a.colscale = reinterpret_cast<const float*>(index + (mode == MODE_L2 ? IL.off_sqnorm : IL.off_rnorm));
  a.cand = reinterpret_cast<uint64_t*>(ws + pl.off_cand);
  a.partial = reinterpret_cast<uint64_t*>(ws + pl.off_partial);
  a.thr_g = reinterpret_cast<uint32_t*>(ws + pl.off_thr);
  // measurement switches (same-box A/B runs): B200IR_GEMM_OPT bit 0 = L2 prefetch of the tile two ahead (off: 2 % slower),
  // bit 1 = publish level 0 only (per-partition thresholds, the pre-"levels" behaviour); B200IR_GEMM_DEBUG = cycle counters
  static const int opt_flags = getenv("B200IR_GEMM_OPT") ? atoi(getenv("B200IR_GEMM_OPT")) : 0;
  static const bool dbg_on = getenv("B200IR_GEMM_DEBUG") != nullptr;
  a.lvl = (pl.P <= 32 && !(opt_flags & 2)) ? reinterpret_cast<uint32_t*>(ws + pl.off_lvl) : nullptr;
  a.opt = opt_flags;
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, size_t(kNumSMs) * 16 * 8);
    cudaMemsetAsync(dbg_buf, 0, size_t(kNumSMs) * 16 * 8, st);
    a.dbg = dbg_buf;
  }
  {
    ProfileScope ps(PT_GEMM, st);
    cudaError_t e;
    auto launch = [&](auto kern, int ncta) -> cudaError_t {
      const int smem_bytes = pl.ares ? SMEM_TOTAL : SMEM_TOTAL_STREAM;
      cudaError_t le = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (le != cudaSuccess) return le;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(unsigned(pl.grid));
      cfg.blockDim = dim3(THREADS);
      cfg.dynamicSmemBytes = smem_bytes;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = unsigned(ncta);
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, kern, tmA, tmA2, tmB, tmB2, a);
    };
    if (!pl.ares) {
      if (f32) {
        if (mode == MODE_COS) e = launch(gemm_topk_kernel<MODE_COS, 2, 3, false>, 2);
        else if (mode == MODE_ABSCOS) e = launch(gemm_topk_kernel<MODE_ABSCOS, 2, 3, false>, 2);
        else e = launch(gemm_topk_kernel<MODE_L2, 2, 3, false>, 2);
      } else {
        if (mode == MODE_COS) e = launch(gemm_topk_kernel<MODE_COS, 2, 1, false>, 2);
        else if (mode == MODE_ABSCOS) e = launch(gemm_topk_kernel<MODE_ABSCOS, 2, 1, false>, 2);
        else e = launch(gemm_topk_kernel<MODE_L2, 2, 1, false>, 2);
      }
    } else if (f32) {
      if (mode == MODE_COS) e = launch(gemm_topk_kernel<MODE_COS, 2, 3, true>, 2);
      else if (mode == MODE_ABSCOS) e = launch(gemm_topk_kernel<MODE_ABSCOS, 2, 3, true>, 2);
      else e = launch(gemm_topk_kernel<MODE_L2, 2, 3, true>, 2);
    } else if (pl.ncta == 2) {
      if (mode == MODE_COS) e = launch(gemm_topk_kernel<MODE_COS, 2, 1, true>, 2);
      else if (mode == MODE_ABSCOS) e = launch(gemm_topk_kernel<MODE_ABSCOS, 2, 1, true>, 2);
      else e = launch(gemm_topk_kernel<MODE_L2, 2, 1, true>, 2);
    } else {
      if (mode == MODE_COS) e = launch(gemm_topk_kernel<MODE_COS, 1, 1, true>, 1);
      else if (mode == MODE_ABSCOS) e = launch(gemm_topk_kernel<MODE_ABSCOS, 1, 1, true>, 1);
      else e = launch(gemm_topk_kernel<MODE_L2, 1, 1, true>, 1);
    }
    if (e != cudaSuccess) return int(e);
    e = cudaGetLastError();
    if (e != cudaSuccess) return int(e);
  }
  if (dbg_on) {
    static unsigned long long host[kNumSMs * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, dbg_buf, sizeof(host), cudaMemcpyDeviceToHost);
    static const char* names[16] = {"prod_wait_Bempty", "prod_wait_Aempty", "prod_wait_Tempty", "mma_wait_Bfull", "mma_wait_Tempty",
                                    "mma_wait_Afull", "mma_issue", "epi_wait_full", "epi_elements", "epi_compact", "epi_unit_end", "mma_all_rounds",
                                    "epi_compactions(w0)", "epi_tiles", "t_start", "mma_round0"};
    fprintf(stderr, "[b200ir gemm debug] grid=%d ncta=%d P=%d tiles/part=%d kp=%d cap=%d (mean cycles per CTA)\n", pl.grid, pl.ncta, pl.P, pl.tiles_per_part, pl.kp, pl.cap);
    for (int sidx = 0; sidx < 16; ++sidx) {
      if (names[sidx][0] == '-') continue;
      double sum = 0, mx = 0;
      for (int c = 0; c < pl.grid; ++c) { sum += double(host[c * 16 + sidx]); if (double(host[c * 16 + sidx]) > mx) mx = double(host[c * 16 + sidx]); }
      fprintf(stderr, "  %-22s mean %14.0f  max %14.0f\n", names[sidx], sum / pl.grid, mx);
    }
  }
  const int rerank = (flags & B200IR_FLAG_NO_RERANK) ? 0 : 1;
  int* fb_count = reinterpret_cast<int*>(ws + pl.off_fb);
  int* fb_list = reinterpret_cast<int*>(ws + pl.off_fb + 256);
  {
    ProfileScope ps(rerank ? PT_RERANK : PT_FINALIZE, st);
    const int blocks = int(ceil_div64(nq, 4));
    // u_eff: bound on |dot~ - dot| / (|q||x|) of the tensor-core pass.  bf16 store: products are exact in fp32, only the
    // accumulation order differs (D/16 accumulate steps).  fp32 store: the split drops q_lo.x_lo and the residuals of
    // hi + lo (<= 3 * 2^-18 together) and accumulates 3 D / 16 steps.  Both constants carry a >= 4x margin over that
    // analysis; tests/test_gpu_tensor_fp32.py measures the actual worst error against them.
    Certify cert{f32 ? 6.1035156e-5f : 1.5258789e-5f, reinterpret_cast<const unsigned int*>(index + IL.off_max),
                 rerank ? fb_count : nullptr, fb_list};
    const bool wide = D > 512, big = pl.kp > 128;
    auto fin = [&](auto kern, auto Qp, auto Xp) {
      size_t ring_bytes = 0;
#if FIN_RING_KB > 0
      using TT = std::remove_cv_t<std::remove_pointer_t<decltype(Qp)>>;
      ring_bytes = 4 * size_t(wide ? FinRing<TT, 8>::kBytes : FinRing<TT, 2>::kBytes);
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ring_bytes));   // per device: set on every call
#endif
      kern<<<blocks, 128, ring_bytes, st>>>(a.partial, a.thr_g, Qp, Xp, reinterpret_cast<const float*>(index + IL.off_sqnorm), int(nq), D, pl.P, pl.kp,
                                            k, mode, rerank, mp, index_offset, cert, out_score, out_idx);
    };
    if (f32) {
      const float* Qp = static_cast<const float*>(Q);
      const float* Xp = static_cast<const float*>(X);
      if (!wide) { if (!big) fin(gemm_finalize_kernel<8, float, 2>, Qp, Xp); else fin(gemm_finalize_kernel<16, float, 2>, Qp, Xp); }
      else { if (!big) fin(gemm_finalize_kernel<8, float, 8>, Qp, Xp); else fin(gemm_finalize_kernel<16, float, 8>, Qp, Xp); }
    } else {
      const __nv_bfloat16* Qp = static_cast<const __nv_bfloat16*>(Q);
      const __nv_bfloat16* Xp = static_cast<const __nv_bfloat16*>(X);
      if (!wide) { if (!big) fin(gemm_finalize_kernel<8, __nv_bfloat16, 2>, Qp, Xp); else fin(gemm_finalize_kernel<16, __nv_bfloat16, 2>, Qp, Xp); }
      else { if (!big) fin(gemm_finalize_kernel<8, __nv_bfloat16, 8>, Qp, Xp); else fin(gemm_finalize_kernel<16, __nv_bfloat16, 8>, Qp, Xp); }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return int(e);
  }
  if (rerank) {
    // queries whose top-k could not be certified are re-done by the exact CUDA-core scan (device-side count: the
    // launches below are no-ops when it is zero)
    cudaError_t e = run_fallback(fbp, dtype, Q, nq, X, N, D, k, mp, fb_count, fb_list, ws + pl.off_fbws, index_offset, out_score, out_idx, st);
    if (e != cudaSuccess) return int(e);
  }
  return 0;
}

}  // namespace b200ir
