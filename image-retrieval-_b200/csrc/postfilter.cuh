// Post-filter of a best-first candidate list (image_search.py:115-140): keep scores >= threshold (absolute, or
// min + t * (max - min) over the list for the optimized score, :118-123), drop later entries whose path was already
// taken (:128-137), cut to top_k.  One warp per query; the list is short (3 * top_k in the reference).
#pragma once
#include "common.cuh"

namespace b200ir {

constexpr int kPostMaxCand = 1024;

__global__ void __launch_bounds__(128) threshold_dedupe_kernel(const float* __restrict__ score, const int64_t* __restrict__ idx,
                                                              int nq, int kc, const int64_t* __restrict__ group, int64_t N,
                                                              double threshold, int relative, int top_k,
                                                              float* __restrict__ out_score, int64_t* __restrict__ out_idx,
                                                              int32_t* __restrict__ out_count) {
  __shared__ int64_t g_s[4][kPostMaxCand];
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  int64_t* g = g_s[threadIdx.x >> 5];
  const float* sc = score + int64_t(q) * kc;
  const int64_t* id = idx + int64_t(q) * kc;
  // valid prefix (lists are padded with idx -1 at the end) and the path group of every candidate
  int valid = 0;
  for (int j0 = 0; j0 < kc; j0 += 32) {
    const int j = j0 + lane;
    const int64_t r = j < kc ? id[j] : -1;
    if (j < kc) g[j] = (r >= 0 && group != nullptr && r < N) ? group[r] : r;
    valid += __popc(__ballot_sync(0xffffffffu, r >= 0));
  }
  __syncwarp();
  double thr = threshold;
  if (relative) {                     // best-first list: max = first, min = last valid entry
    const double mx = valid > 0 ? double(sc[0]) : 1.0;
    const double mn = valid > 0 ? double(sc[valid - 1]) : 0.0;
    thr = mn + threshold * (mx - mn);
  }
  int taken = 0;
  for (int j0 = 0; j0 < valid && taken < top_k; j0 += 32) {
    const int j = j0 + lane;
    bool keep = j < valid && double(sc[j]) >= thr;
    if (keep) {
      const int64_t mine = g[j];
      for (int i = 0; i < j; ++i) {
        if (g[i] == mine) { keep = false; break; }
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    const int pos = taken + __popc(m & ((1u << lane) - 1));
    if (keep && pos < top_k) {
      out_score[int64_t(q) * top_k + pos] = sc[j];
      out_idx[int64_t(q) * top_k + pos] = id[j];
    }
    taken += __popc(m);
  }
  taken = min(taken, top_k);
  for (int p = taken + lane; p < top_k; p += 32) {
    out_score[int64_t(q) * top_k + p] = -INFINITY;
    out_idx[int64_t(q) * top_k + p] = -1;
  }
  if (out_count != nullptr && lane == 0) out_count[q] = taken;
}

}  // namespace b200ir
