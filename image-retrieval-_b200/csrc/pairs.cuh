// Explicit pair lists: the seven values of get_all_metrics (geometric_metrics.py:114-129) for every listed
// (row of A, row of B) pair - the arithmetic of ColorMIAnalyzer.calculate_distances (mi_analysis.py:256-297),
// which walks pairs.json instead of the full (i, j) grid.
//
// One warp per pair: both rows are gathered with 16-byte loads (bytes per pair = 2 * D * sizeof(elem), each row
// read once), six fp32 partials per lane, one butterfly reduction.  HBM / L2 gather bound; no shared memory.
#pragma once
#include "common.cuh"

namespace b200ir {

constexpr int kPairOutputs = 7;   // cosine_similarity, cosine_distance, angular_distance, l1, l2, linf, magnitude_difference

struct PairAcc {
  float dot = 0.f, na = 0.f, nb = 0.f, l1 = 0.f, l2 = 0.f, linf = 0.f;
  __device__ __forceinline__ void add(float a, float b) {
    const float d = a - b;
    dot = fmaf(a, b, dot);
    na = fmaf(a, a, na);
    nb = fmaf(b, b, nb);
    l1 += fabsf(d);
    l2 = fmaf(d, d, l2);
    linf = fmaxf(linf, fabsf(d));
  }
};

template <typename T> struct PairVec;
template <> struct PairVec<float> {
  static constexpr int kElems = 4;
  __device__ static __forceinline__ void accumulate(const float* a, const float* b, PairAcc& acc) {
    const float4 va = __ldg(reinterpret_cast<const float4*>(a));
    const float4 vb = __ldg(reinterpret_cast<const float4*>(b));
    acc.add(va.x, vb.x); acc.add(va.y, vb.y); acc.add(va.z, vb.z); acc.add(va.w, vb.w);
  }
};
template <> struct PairVec<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void accumulate(const __nv_bfloat16* a, const __nv_bfloat16* b, PairAcc& acc) {
    const uint4 va = __ldg(reinterpret_cast<const uint4*>(a));
    const uint4 vb = __ldg(reinterpret_cast<const uint4*>(b));
    acc.add(bf16_lo(va.x), bf16_lo(vb.x)); acc.add(bf16_hi(va.x), bf16_hi(vb.x));
    acc.add(bf16_lo(va.y), bf16_lo(vb.y)); acc.add(bf16_hi(va.y), bf16_hi(vb.y));
    acc.add(bf16_lo(va.z), bf16_lo(vb.z)); acc.add(bf16_hi(va.z), bf16_hi(vb.z));
    acc.add(bf16_lo(va.w), bf16_lo(vb.w)); acc.add(bf16_hi(va.w), bf16_hi(vb.w));
  }
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) pair_metrics_kernel(const T* __restrict__ A, int64_t NA, const T* __restrict__ B,
                                                           int64_t NB, int D, const int64_t* __restrict__ ia,
                                                           const int64_t* __restrict__ ib, int64_t P, int ia_group,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t p = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; p < P; p += warps) {
    const int64_t i = ia != nullptr ? ia[p] : p / ia_group, j = ib[p];     // ia == nullptr: candidate lists, ia_group per row of A
    if (i < 0 || i >= NA || j < 0 || j >= NB) {      // unknown row: the pair is reported as NaN, never read
      if (lane < kPairOutputs) out[int64_t(lane) * P + p] = __int_as_float(0x7fc00000);
      continue;
    }
    const T* a = A + i * D;
    const T* b = B + j * D;
    PairAcc acc;
    if (VEC) {
      constexpr int E = PairVec<T>::kElems;
#pragma unroll 4
      for (int d = lane * E; d < D; d += 32 * E) PairVec<T>::accumulate(a + d, b + d, acc);
    } else {
      for (int d = lane; d < D; d += 32) acc.add(to_f32<T>(a[d]), to_f32<T>(b[d]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc.dot += __shfl_xor_sync(0xffffffffu, acc.dot, o);
      acc.na += __shfl_xor_sync(0xffffffffu, acc.na, o);
      acc.nb += __shfl_xor_sync(0xffffffffu, acc.nb, o);
      acc.l1 += __shfl_xor_sync(0xffffffffu, acc.l1, o);
      acc.l2 += __shfl_xor_sync(0xffffffffu, acc.l2, o);
      acc.linf = fmaxf(acc.linf, __shfl_xor_sync(0xffffffffu, acc.linf, o));
    }
    const float norm_a = sqrtf(acc.na), norm_b = sqrtf(acc.nb);
    const float cs = (norm_a == 0.f || norm_b == 0.f) ? 0.f : acc.dot / (norm_a * norm_b);     // geometric_metrics.py:16-18
    float v;
    switch (lane) {
      case 0: v = cs; break;
      case 1: v = 1.0f - cs; break;                                                            // :29-31
      case 2: v = acosf(fminf(1.0f, fmaxf(-1.0f, cs))); break;                                 // :21-26
      case 3: v = acc.l1 / float(D); break;                                                    // :34-39 normalized
      case 4: v = sqrtf(acc.l2) / sqrtf(float(D)); break;                                      // :42-47 normalized
      case 5: v = acc.linf; break;                                                             // :50-52
      default: v = fabsf(norm_a - norm_b); break;                                              // :55-57
    }
    if (lane < kPairOutputs) out[int64_t(lane) * P + p] = v;
  }
}

template <typename T>
inline cudaError_t launch_pair_metrics(const void* A, int64_t NA, const void* B, int64_t NB, int D, const int64_t* ia,
                                       const int64_t* ib, int64_t P, float* out, cudaStream_t st, int ia_group = 1) {
  const int64_t want = ceil_div64(P, 8);
  const int blocks = int(want < int64_t(kNumSMs) * 8 * 4 ? want : int64_t(kNumSMs) * 8 * 4);   // grid-stride past 8 CTAs/SM x 4 waves
  const bool vec = D % PairVec<T>::kElems == 0 && reinterpret_cast<uintptr_t>(A) % 16 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0;
  if (vec) pair_metrics_kernel<T, true><<<blocks, 256, 0, st>>>(static_cast<const T*>(A), NA, static_cast<const T*>(B), NB, D, ia, ib, P, ia_group, out);
  else pair_metrics_kernel<T, false><<<blocks, 256, 0, st>>>(static_cast<const T*>(A), NA, static_cast<const T*>(B), NB, D, ia, ib, P, ia_group, out);
  return cudaGetLastError();
}

}  // namespace b200ir
