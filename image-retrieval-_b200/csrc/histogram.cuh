// 8x8x8 colour histogram of uint8 RGB images (K10 of SURVEY.md section 2.2): shared-memory
// privatised (one 512-bin copy per warp) atomics, 48-byte (16-pixel) vector loads per thread and
// run-length aggregation of equal consecutive bins, optional OpenCV-exact 8-bit RGB->HSV.
#pragma once
#include "common.cuh"

namespace b200ir {

constexpr int kHistThreads = 256;
constexpr int kHistWarps = kHistThreads / 32;
constexpr int kHistBins = 512;

struct HsvTables { int sdiv[256]; int hdiv[256]; };
__constant__ HsvTables c_hsv_tables;

// OpenCV RGB2HSV_b (hsv_shift = 12, hrange = 180): see oracle/histogram.py for the restatement.
__device__ __forceinline__ int hsv_bin(int r, int g, int b, const int* sdiv, const int* hdiv) {
  const int v = max(r, max(g, b));
  const int mn = min(r, min(g, b));
  const int d = v - mn;
  const int s = (d * sdiv[v] + (1 << 11)) >> 12;
  const int h0 = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
  int h = (h0 * hdiv[d] + (1 << 11)) >> 12;     // arithmetic shift, like the C++ original
  h += h < 0 ? 180 : 0;
  return (((h * 365) >> 13) << 6) | ((s >> 5) << 3) | (v >> 5);     // (h*365)>>13 == h*8/180 for h in [0,180)
}
__device__ __forceinline__ int rgb_bin(int r, int g, int b) { return (r >> 5) * 64 + (g >> 5) * 8 + (b >> 5); }

template <bool HSV>
__global__ void __launch_bounds__(kHistThreads) histogram_kernel(const uint8_t* __restrict__ img, int64_t pixels_per_image,
                                                                 int slices, int vector_ok, uint32_t* __restrict__ out) {
  __shared__ uint32_t hist[kHistWarps][kHistBins];
  __shared__ int s_sdiv[HSV ? 256 : 1];
  __shared__ int s_hdiv[HSV ? 256 : 1];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kHistWarps * kHistBins; i += kHistThreads) (&hist[0][0])[i] = 0;
  if constexpr (HSV) {
    for (int i = tid; i < 256; i += kHistThreads) { s_sdiv[i] = c_hsv_tables.sdiv[i]; s_hdiv[i] = c_hsv_tables.hdiv[i]; }
  }
  __syncthreads();

  const int64_t image = blockIdx.x / slices;
  const int slice = blockIdx.x % slices;
  const uint8_t* base = img + image * pixels_per_image * 3;
  uint32_t* myhist = hist[warp];
  int cur_bin = -1;
  uint32_t run = 0;
  auto add_bin = [&](int bin) {
    if (bin == cur_bin) { ++run; }
    else { if (run) atomicAdd(&myhist[cur_bin], run); cur_bin = bin; run = 1; }
  };
  auto add_pixel = [&](int r, int g, int b) { add_bin(HSV ? hsv_bin(r, g, b, s_sdiv, s_hdiv) : rgb_bin(r, g, b)); };

  if (vector_ok) {
    // 16 pixels = 48 bytes = three 128-bit loads per thread per step
    const int64_t groups = pixels_per_image / 16;
    const int64_t per_slice = ceil_div64(groups, slices);
    const int64_t g_begin = slice * per_slice, g_end = min(groups, g_begin + per_slice);
    const uint4* v = reinterpret_cast<const uint4*>(base);
    for (int64_t gi = g_begin + tid; gi < g_end; gi += kHistThreads) {
      uint32_t w[12];
      const uint4 a = __ldg(v + gi * 3), b4 = __ldg(v + gi * 3 + 1), c = __ldg(v + gi * 3 + 2);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b4.x; w[5] = b4.y; w[6] = b4.z; w[7] = b4.w;
      w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int b0 = p * 3, b1 = b0 + 1, b2 = b0 + 2;
        if constexpr (HSV) {
          const int r = int(__byte_perm(w[b0 >> 2], 0, 0x4440 | (b0 & 3)));      // one PRMT per byte
          const int g = int(__byte_perm(w[b1 >> 2], 0, 0x4440 | (b1 & 3)));
          const int bl = int(__byte_perm(w[b2 >> 2], 0, 0x4440 | (b2 & 3)));
          add_pixel(r, g, bl);
        } else {
          // RGB bin straight from the packed pixel: align its three bytes to bits 0..23 (one funnel shift), keep the top
          // three bits of each byte (0x00E0E0E0) and gather them with one multiply: x * (2^22 + 2^11 + 1) puts the
          // r / g / b fields at bits 27..29 / 24..26 / 21..23 (all other partial products fall outside 21..29).
          const uint32_t lo = w[b0 >> 2], hi = (b0 >> 2) + 1 < 12 ? w[(b0 >> 2) + 1] : 0u;
          const uint32_t win = (b0 & 3) ? __funnelshift_r(lo, hi, 8 * (b0 & 3)) : lo;
          add_bin(int(((win & 0x00E0E0E0u) * 0x00400801u) >> 21));
        }
      }
    }
  } else {
    const int64_t per_slice = ceil_div64(pixels_per_image, slices);
    const int64_t p_begin = slice * per_slice, p_end = min(pixels_per_image, p_begin + per_slice);
    for (int64_t pi = p_begin + tid; pi < p_end; pi += kHistThreads) {
      const uint8_t* px = base + pi * 3;
      add_pixel(px[0], px[1], px[2]);
    }
  }
  if (run) atomicAdd(&myhist[cur_bin], run);
  __syncthreads();
  uint32_t* dst = out + image * kHistBins;
  for (int bin = tid; bin < kHistBins; bin += kHistThreads) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < kHistWarps; ++w) s += hist[w][bin];
    if (slices == 1) dst[bin] = s;
    else if (s) atomicAdd(&dst[bin], s);
  }
}

// counts -> fp32 raw / unit-norm vector + magnitude (ImageEmbeddingSystem.py:88-94)
__global__ void __launch_bounds__(128) counts_to_embedding_kernel(const uint32_t* __restrict__ counts, int nb,
                                                                  float* __restrict__ raw, float* __restrict__ unit,
                                                                  float* __restrict__ mag) {
  __shared__ float red[4];
  const int64_t row = blockIdx.x;
  const uint32_t* c = counts + row * nb;
  float ss = 0.f;
  for (int i = threadIdx.x; i < nb; i += 128) { const float v = float(c[i]); ss = fmaf(v, v, ss); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float m = sqrtf(red[0] + red[1] + red[2] + red[3]);
  if (threadIdx.x == 0 && mag) mag[row] = m;
  for (int i = threadIdx.x; i < nb; i += 128) {
    const float v = float(c[i]);
    if (raw) raw[row * nb + i] = v;
    if (unit) unit[row * nb + i] = v / m;
  }
}

inline cudaError_t init_hsv_tables() {
  static bool done_dev[64] = {};
  int dev = 0;
  cudaError_t e0 = cudaGetDevice(&dev);
  if (e0 != cudaSuccess) return e0;
  bool& done = done_dev[dev & 63];
  if (done) return cudaSuccess;
  HsvTables t;
  t.sdiv[0] = 0; t.hdiv[0] = 0;
  for (int i = 1; i < 256; ++i) {
    t.sdiv[i] = int(rint((255 << 12) / double(i)));
    t.hdiv[i] = int(rint((180 << 12) / (6.0 * i)));
  }
  cudaError_t e = cudaMemcpyToSymbol(c_hsv_tables, &t, sizeof(t));
  if (e == cudaSuccess) done = true;
  return e;
}

}  // namespace b200ir
