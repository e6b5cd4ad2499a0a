// Selection kernels: merge of per-partition key lists into the final top-k (+ metric transform),
// and the cross-shard merge that follows the all-gather of a row-sharded search.
#pragma once
#include "common.cuh"
#include "profile.h"

namespace b200ir {

// Final step shared by every top-k path: the first k keys of r[] (sorted by rank value) are turned into the
// metric's reference-normalised score, then re-sorted by (score, index): transforms such as 1 - cos, sqrt or
// arccos can round distinct rank values to the same fp32 score, and the reference's stable sort orders equal
// scores by index (app_pipeline.py:171-172, image_search.py:199-219).
// SHARED_SORT: sort through the out-of-line copy of the network (kernels that sort at several places)
template <int E, bool SHARED_SORT = false>
__device__ __forceinline__ void emit_topk(uint64_t (&r)[E], int lane, int k, const MetricParams& mp, int64_t index_offset,
                                          float* __restrict__ out_score, int64_t* __restrict__ out_idx, int64_t q) {
  const bool desc = metric_descending(mp.metric);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    if (i < k && r[e] != kKeyInf) {
      const float sc = rank_to_score(key_rank(r[e]), mp.metric, mp.flags, mp.D);
      r[e] = make_key(desc ? -sc : sc, key_index(r[e]));
    } else {
      r[e] = kKeyInf;
    }
  }
  if constexpr (SHARED_SORT) warp_sort_shared<E>(r, lane); else warp_sort<E>(r, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    if (i < k) {
      float sc;
      int64_t id;
      if (r[e] == kKeyInf) { sc = desc ? -INFINITY : INFINITY; id = -1; }
      else {
        const float v = key_rank(r[e]);
        sc = desc ? 0.0f - v : v;
        id = int64_t(key_index(r[e])) + index_offset;
      }
      out_score[q * k + i] = sc;
      out_idx[q * k + i] = id;
    }
  }
}

// One warp per query: stream P*k sorted-or-not keys through a 32*E-wide bitonic sorter, keeping
// the best k between rounds.  Writes the reference-normalised score and the global index.
// Options of the final merge.
//   qmap / gate: the launch finishes positions [0, min(nq, *gate - gate_base)) of a device-side query list and writes
//                row qmap[q] of the outputs (fallback of the tensor path).
//   last_key:    paging cursor out, the rank key of the k-th winner of each query (kKeyInf: the store is exhausted).
//   nmetrics>0:  multi-list scan; blockIdx.y = requested metric y, which reads list slot[y] ([nl][nq][P][k] keys,
//                list_stride apart), applies metric[y]'s score transform and writes plane y of the outputs.
struct FinalizeOpts {
  const int* qmap = nullptr;
  const int* gate = nullptr;
  int gate_base = 0;
  uint64_t* last_key = nullptr;
  int nmetrics = 0;
  signed char slot[B200IR_NUM_METRICS] = {0};
  signed char metric[B200IR_NUM_METRICS] = {0};
  int64_t list_stride = 0, out_stride = 0;
};

template <int E>
__global__ void __launch_bounds__(128) finalize_topk_kernel(const uint64_t* __restrict__ partial, int nq, int64_t per_query,
                                                           int k, MetricParams mp, int64_t index_offset,
                                                           float* __restrict__ out_score, int64_t* __restrict__ out_idx,
                                                           const FinalizeOpts o) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  if (o.gate != nullptr && q >= *o.gate - o.gate_base) return;
  if (o.nmetrics > 0) {
    const int y = blockIdx.y;
    partial += int64_t(o.slot[y]) * o.list_stride;
    out_score += int64_t(y) * o.out_stride;
    out_idx += int64_t(y) * o.out_stride;
    mp.metric = o.metric[y];
  }
  const uint64_t* src = partial + int64_t(q) * per_query;
  uint64_t r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = kKeyInf;
  int kept = 0;
  int64_t pos = 0;
  do {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = lane * E + e;
      if (i >= kept) {
        const int64_t s = pos + (i - kept);
        r[e] = s < per_query ? src[s] : kKeyInf;
      }
    }
    pos += 32 * E - kept;
    warp_sort<E>(r, lane);
    kept = k;
  } while (pos < per_query);
  if (o.last_key != nullptr) {
#pragma unroll
    for (int e = 0; e < E; ++e) if (lane * E + e == k - 1) o.last_key[q] = r[e];
  }
  emit_topk<E>(r, lane, k, mp, index_offset, out_score, out_idx, o.qmap != nullptr ? int64_t(o.qmap[q]) : int64_t(q));
}

// Cross-shard merge: score/idx lists of R shards -> [nq, k] ordered by (score, global idx).
// Every shard list is sorted, so the best of the shards' k-th scores bounds the global k-th score: entries worse than
// it are dropped while streaming (ties kept), and what survives (about k..2k keys) usually fits one sort round.
template <int E>
__global__ void __launch_bounds__(128) merge_topk_kernel(int descending, const float* __restrict__ score,
                                                        const int64_t* __restrict__ idx, int64_t score_stride,
                                                        int64_t idx_stride, int R, int nq, int k,
                                                        float* __restrict__ out_score, int64_t* __restrict__ out_idx) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  __shared__ key128_t stage_s[4][32 * E];
  key128_t* stage = stage_s[threadIdx.x >> 5];
  const key128_t kInf = ~key128_t(0);
  const int64_t per_query = int64_t(R) * k;
  const int64_t qoff = int64_t(q) * k;
  // Two valid bounds on the global k-th score (ordered domain, smaller = better):
  //   (a) the best of the shards' k-th scores (that shard alone holds k entries at least as good);
  //   (b) the worst of the shards' m-th scores, m = ceil(k / R): together the shards hold R*m >= k entries at least
  //       as good.  For a balanced row-sharded store (b) is the tight one (each shard contributes ~k/R winners).
  const int m = (k + R - 1) / R;
  uint32_t bound = 0xffffffffu, bound_b = 0u;
  bool b_ok = true;
  for (int sh = lane; sh < R; sh += 32) {
    if (idx[int64_t(sh) * idx_stride + qoff + k - 1] >= 0) {
      const float v = score[int64_t(sh) * score_stride + qoff + k - 1];
      bound = min(bound, f32_to_ordered(descending ? -v : v));
    }
    if (idx[int64_t(sh) * idx_stride + qoff + m - 1] >= 0) {
      const float v = score[int64_t(sh) * score_stride + qoff + m - 1];
      bound_b = max(bound_b, f32_to_ordered(descending ? -v : v));
    } else {
      b_ok = false;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bound = min(bound, __shfl_xor_sync(0xffffffffu, bound, o));
    bound_b = max(bound_b, __shfl_xor_sync(0xffffffffu, bound_b, o));
  }
  if (__all_sync(0xffffffffu, b_ok)) bound = min(bound, bound_b);
  key128_t r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = kInf;
  int kept = 0, fill = 0;
  for (int64_t base = 0; base < per_query; base += 32) {
    const int64_t s = base + lane;
    key128_t key = kInf;
    bool keep = false;
    if (s < per_query) {
      const int shard = int(s / k), j = int(s % k);
      const int64_t id = idx[int64_t(shard) * idx_stride + qoff + j];
      if (id >= 0) {
        const float v = score[int64_t(shard) * score_stride + qoff + j];
        const uint32_t o = f32_to_ordered(descending ? -v : v);
        keep = o <= bound;
        key = (key128_t(o) << 64) | key128_t(uint64_t(id));
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) stage[fill + __popc(m & ((1u << lane) - 1))] = key;
    fill += __popc(m);
    if (fill > 32 * E - kept - 32 || base + 32 >= per_query) {
      __syncwarp();
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = lane * E + e;
        if (i >= kept) r[e] = (i - kept) < fill ? stage[i - kept] : kInf;
      }
      warp_sort<E>(r, lane);
      kept = min(kept + fill, k);
      fill = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) if (lane * E + e >= kept) r[e] = kInf;
      __syncwarp();
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    if (i < k) {
      const key128_t key = r[e];
      float sc;
      int64_t id;
      if (key == kInf) { sc = descending ? -INFINITY : INFINITY; id = -1; }
      else {
        const float v = ordered_to_f32(uint32_t(uint64_t(key >> 64)));
        sc = descending ? 0.0f - v : v;
        id = int64_t(uint64_t(key));
      }
      out_score[int64_t(q) * k + i] = sc;
      out_idx[int64_t(q) * k + i] = id;
    }
  }
}

inline cudaError_t launch_finalize(const uint64_t* partial, int64_t nq, int64_t per_query, int k, const MetricParams& mp,
                                   int64_t index_offset, float* out_score, int64_t* out_idx, cudaStream_t st,
                                   const FinalizeOpts& o = FinalizeOpts()) {
  const dim3 grid(unsigned(ceil_div64(nq, 4)), unsigned(o.nmetrics > 0 ? o.nmetrics : 1));
  ProfileScope ps(PT_FINALIZE, st);
  if (k <= 128) finalize_topk_kernel<8><<<grid, 128, 0, st>>>(partial, int(nq), per_query, k, mp, index_offset, out_score, out_idx, o);
  else finalize_topk_kernel<16><<<grid, 128, 0, st>>>(partial, int(nq), per_query, k, mp, index_offset, out_score, out_idx, o);
  return cudaGetLastError();
}

inline cudaError_t launch_merge(int descending, const float* score, const int64_t* idx, int64_t score_stride, int64_t idx_stride,
                                int R, int64_t nq, int k, float* out_score, int64_t* out_idx, cudaStream_t st) {
  const int blocks = int(ceil_div64(nq, 4));
  ProfileScope ps(PT_MERGE, st);
  if (k <= 128) merge_topk_kernel<8><<<blocks, 128, 0, st>>>(descending, score, idx, score_stride, idx_stride, R, int(nq), k, out_score, out_idx);
  else merge_topk_kernel<16><<<blocks, 128, 0, st>>>(descending, score, idx, score_stride, idx_stride, R, int(nq), k, out_score, out_idx);
  return cudaGetLastError();
}

}  // namespace b200ir
