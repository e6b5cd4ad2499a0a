// Shared device helpers: ranking keys, warp bitonic sort, cp.async, small utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/b200ir.h"

namespace b200ir {

constexpr int kNumSMs = 148;           // B200: 2 dies x 74 SMs
constexpr uint64_t kKeyInf = ~0ull;    // "no candidate" key, sorts last

// ---------------------------------------------------------------------------------------------
// Ranking keys.  Every metric is turned into a rank value r (fp32, smaller = better):
//   L1: sum|d|   L2: sum d^2   Linf: max|d|   cosine family: -cos   optimized: -similarity
// and packed with the shard-local row index into one 64-bit key so that an unsigned compare is
// "(r, index) lexicographic" == Python's stable sort on r (ties -> lower index first,
// app_pipeline.py:171-172).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f + 0.0f);     // canonicalise -0.0 -> +0.0 (they compare equal in Python)
#else
  union { float f; uint32_t u; } c; c.f = f + 0.0f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float r, uint32_t local_idx) {
  return (uint64_t(f32_to_ordered(r)) << 32) | local_idx;
}
__device__ __forceinline__ float key_rank(uint64_t k) { return ordered_to_f32(uint32_t(k >> 32)); }
__device__ __forceinline__ uint32_t key_index(uint64_t k) { return uint32_t(k); }

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, uint32_t(v), m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, uint32_t(v >> 32), m);
  return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, uint32_t(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, uint32_t(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
}

typedef unsigned __int128 key128_t;   // (ordered score << 64) | global index, for the cross-shard merge
__device__ __forceinline__ key128_t shfl_xor_key(key128_t v, int m) {
  const uint64_t lo = shfl_xor_u64(uint64_t(v), m);
  const uint64_t hi = shfl_xor_u64(uint64_t(v >> 64), m);
  return (key128_t(hi) << 64) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_key(uint64_t v, int m) { return shfl_xor_u64(v, m); }
__device__ __forceinline__ uint32_t shfl_xor_key(uint32_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Warp-wide ascending bitonic sort of 32*E keys held E per lane in blocked order
// (global position of k[e] in lane l is l*E + e).  All 32 lanes must call.
template <int E, typename K>
__device__ __forceinline__ void warp_sort(K (&k)[E], int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= E) {
        const int lstride = stride / E;
        const bool lower = (lane & lstride) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = lane * E + e;
          const bool up = (i & size) == 0;
          const K other = shfl_xor_key(k[e], lstride);
          const K mn = k[e] < other ? k[e] : other;
          const K mx = k[e] < other ? other : k[e];
          k[e] = (lower == up) ? mn : mx;
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & stride) == 0) {
            const int i = lane * E + e;
            const bool up = (i & size) == 0;
            const K a = k[e], b = k[e | stride];
            const bool sw = (a > b) == up;
            k[e] = sw ? b : a;
            k[e | stride] = sw ? a : b;
          }
        }
      }
    }
  }
}

// The same sort as ONE out-of-line function (keys travel in registers, by value): a kernel that sorts at several places
// otherwise carries one ~30 KB copy of the network per call site and becomes instruction-fetch bound.
template <int E> struct KeyRegs { uint64_t k[E]; };
template <int E>
__device__ __noinline__ KeyRegs<E> warp_sort_call(KeyRegs<E> v, int lane) {
  warp_sort<E>(v.k, lane);
  return v;
}
template <int E>
__device__ __forceinline__ void warp_sort_shared(uint64_t (&k)[E], int lane) {
  KeyRegs<E> v;
#pragma unroll
  for (int e = 0; e < E; ++e) v.k[e] = k[e];
  v = warp_sort_call<E>(v, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) k[e] = v.k[e];
}

// Final pass of the sort above on its own: `k` (32*E keys, blocked) holds a bitonic sequence, result ascending.
template <int E, typename K>
__device__ __forceinline__ void warp_bitonic_merge(K (&k)[E], int lane) {
#pragma unroll
  for (int stride = 16 * E; stride > 0; stride >>= 1) {
    if (stride >= E) {
      const int lstride = stride / E;
      const bool lower = (lane & lstride) == 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const K other = shfl_xor_key(k[e], lstride);
        const K mn = k[e] < other ? k[e] : other;
        const K mx = k[e] < other ? other : k[e];
        k[e] = lower ? mn : mx;
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & stride) == 0) {
          const K a = k[e], b = k[e | stride];
          const bool sw = a > b;
          k[e] = sw ? b : a;
          k[e | stride] = sw ? a : b;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cp.async (LDGSTS) helpers: 16-byte copies with zero-fill of the bytes past `src_bytes`.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

// Final value of a winner from its rank value r (see b200ir.h for the per-metric definitions).
struct MetricParams {
  int metric;
  int flags;
  int D;
  float w[5];   // w_angle, w_l1, w_l2, w_inf, w_mag (OPTIMIZED)
};

__device__ __forceinline__ float rank_to_score(float r, int metric, int flags, int D) {
  switch (metric) {
    case B200IR_L1:       return (flags & B200IR_FLAG_RAW) ? r : r / float(D);
    case B200IR_L2:       { float s = sqrtf(r); return (flags & B200IR_FLAG_RAW) ? s : s / sqrtf(float(D)); }
    case B200IR_LINF:     return r;
    case B200IR_COS_SIM:  return 0.0f - r;
    case B200IR_COS_DIST: return 1.0f + r;                                   // 1 - cos, r = -cos
    case B200IR_ANGLE:    return acosf(fminf(1.0f, fmaxf(-1.0f, -r)));
    case B200IR_MAG_DIFF: return r;
    case B200IR_OPTIMIZED:return 0.0f - r;
    default:              return r;
  }
}
__host__ __device__ __forceinline__ bool metric_descending(int metric) {
  return metric == B200IR_COS_SIM || metric == B200IR_OPTIMIZED;
}

}  // namespace b200ir
