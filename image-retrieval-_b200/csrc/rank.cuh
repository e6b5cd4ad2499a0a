// Small per-query ordering kernels around the scan:
//   sort_topk_rows_kernel   - final (score, index) stable order of a result list longer than one B200IR_MAX_K page
//   rank_candidates_kernel  - the six per-metric orderings of a query's candidate list (image_search.py:199-219) and the
//                             weighted "optimized" score of every candidate (geometric_metrics.py:85-92)
#pragma once
#include "common.cuh"

namespace b200ir {

constexpr int kSortRowsMax = 4096;     // longest paged result list
constexpr int kRankMaxCand = 1024;     // longest candidate list of rank_candidates

// In-place bitonic sort of n (power of two) keys in shared memory by one CTA.
template <typename K>
__device__ __forceinline__ void cta_bitonic_sort(K* keys, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const K a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// One CTA per query: order row q of score / idx [nq][K] by (score, index) - what Python's stable sort gives on a list in
// database order (app_pipeline.py:171-172); padding entries (idx < 0) go last.
__global__ void __launch_bounds__(256) sort_topk_rows_kernel(int descending, float* __restrict__ score, int64_t* __restrict__ idx,
                                                            int K, int n_pow2) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  key128_t* keys = reinterpret_cast<key128_t*>(sort_smem);
  const int64_t base = int64_t(blockIdx.x) * K;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    key128_t key = ~key128_t(0);
    if (i < K) {
      const int64_t id = idx[base + i];
      if (id >= 0) {
        const float v = score[base + i];
        key = (key128_t(f32_to_ordered(descending ? -v : v)) << 64) | key128_t(uint64_t(id));
      }
    }
    keys[i] = key;
  }
  cta_bitonic_sort(keys, n_pow2);
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const key128_t key = keys[i];
    if (key == ~key128_t(0)) { score[base + i] = descending ? -INFINITY : INFINITY; idx[base + i] = -1; }
    else {
      const float v = ordered_to_f32(uint32_t(uint64_t(key >> 64)));
      score[base + i] = descending ? 0.0f - v : v;
      idx[base + i] = int64_t(uint64_t(key));
    }
  }
}

// One CTA per query.  pair_vals [7][nq * kc]: the get_all_metrics values of the pairs (query q, candidate c) as
// b200ir_pair_metrics writes them (cosine_similarity, cosine_distance, angular_distance, l1, l2, linf, magnitude);
// cand_idx [nq][kc]: database rows of the candidates, best cosine first, -1 = padding.
// Orderings y (image_search.py:199-219): 0 cosine desc, 1 l1 asc, 2 l2 asc, 3 linf asc, 4 magnitude asc, 5 optimized desc;
// ties keep candidate order (Python's stable sort of the candidate list).  For each y the first k positions / values /
// rows are written to out_pos / out_val / out_row [6][nq][k] (padding -1 / NaN / -1).
__global__ void __launch_bounds__(128) rank_candidates_kernel(const float* __restrict__ pair_vals, const int64_t* __restrict__ cand_idx,
                                                             int64_t nq, int kc, int n_pow2, float w0, float w1, float w2, float w3,
                                                             float w4, int k, float* __restrict__ out_optimized,
                                                             int32_t* __restrict__ out_pos, float* __restrict__ out_val,
                                                             int64_t* __restrict__ out_row) {
  __shared__ uint64_t keys[kRankMaxCand];
  __shared__ float opt[kRankMaxCand];
  const int64_t q = blockIdx.x;
  const int64_t total = nq * int64_t(kc);
  const int64_t base = q * kc;
  for (int c = threadIdx.x; c < kc; c += blockDim.x) {
    const int64_t p = base + c;
    // geometric_metrics.py:85-92: w_angle * cos - w_l1 * L1n - w_l2 * L2n - w_inf * Linf - w_mag * mag
    const float sim = w0 * pair_vals[p] - w1 * pair_vals[3 * total + p] - w2 * pair_vals[4 * total + p] - w3 * pair_vals[5 * total + p] -
                      w4 * pair_vals[6 * total + p];
    opt[c] = sim;
    if (out_optimized != nullptr) out_optimized[p] = sim;
  }
  const int plane[6] = {0, 3, 4, 5, 6, -1};
  for (int y = 0; y < 6; ++y) {
    __syncthreads();
    const bool desc = (y == 0 || y == 5);
    for (int c = threadIdx.x; c < n_pow2; c += blockDim.x) {
      uint64_t key = kKeyInf;
      if (c < kc && cand_idx[base + c] >= 0) {
        const float v = y == 5 ? opt[c] : pair_vals[plane[y] * total + base + c];
        if (v == v) key = (uint64_t(f32_to_ordered(desc ? -v : v)) << 32) | uint32_t(c);
      }
      keys[c] = key;
    }
    cta_bitonic_sort(keys, n_pow2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const uint64_t key = i < n_pow2 ? keys[i] : kKeyInf;
      const int64_t o = (int64_t(y) * nq + q) * k + i;
      if (key == kKeyInf) { out_pos[o] = -1; out_val[o] = __int_as_float(0x7fc00000); out_row[o] = -1; }
      else {
        const int c = int(uint32_t(key));
        const float v = ordered_to_f32(uint32_t(key >> 32));
        out_pos[o] = c;
        out_val[o] = desc ? 0.0f - v : v;
        out_row[o] = cand_idx[base + c];
      }
    }
  }
}

}  // namespace b200ir
