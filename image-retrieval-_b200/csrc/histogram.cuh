// 8x8x8 colour histogram of uint8 RGB images (K10 of SURVEY.md section 2.2): shared-memory
// privatised (one 512-bin copy per warp) atomics, 48-byte (16-pixel) vector loads per thread and
// run-length aggregation of equal consecutive bins; the OpenCV-exact 8-bit RGB->HSV variant is its own kernel below.
#pragma once
#include "common.cuh"
#include "tma_util.h"

namespace b200ir {

constexpr int kHistThreads = 256;
constexpr int kHistWarps = kHistThreads / 32;
constexpr int kHistBins = 512;

struct HsvTables { int sdiv[256]; int hdiv[256]; };
__constant__ HsvTables c_hsv_tables;

__device__ __forceinline__ int rgb_bin(int r, int g, int b) { return (r >> 5) * 64 + (g >> 5) * 8 + (b >> 5); }

__global__ void __launch_bounds__(kHistThreads) histogram_kernel(const uint8_t* __restrict__ img, int64_t pixels_per_image,
                                                                 int slices, int vector_ok, uint32_t* __restrict__ out) {
  __shared__ uint32_t hist[kHistWarps][kHistBins];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kHistWarps * kHistBins; i += kHistThreads) (&hist[0][0])[i] = 0;
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
  auto add_pixel = [&](int r, int g, int b) { add_bin(rgb_bin(r, g, b)); };

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
        const int b0 = p * 3;
        {
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

// ------------------------------------------------------------------------------------------------------------------
// HSV histogram, round 2.  ncu on the round-1 kernel (profiles/r1_ncu_secondary_kernels.md): 46 lane-instructions per
// pixel, issue slots 80 % busy, ALU pipe 69 %, shared-memory wavefronts at 76-84 % of peak with 26 M of 36.5 M being bank
// conflicts of the two division-table look-ups.  This kernel keeps OpenCV's fixed-point arithmetic bit for bit but does
// it in the FMA pipe on fp32 values that are exact integers:
//   * one PRMT per channel builds the float 2^23 + byte (0x4B0000bb); min / max are 3-input FMNMX3;
//   * d*sdiv[v] + 2048 and hnum*hdiv[d] + 2048 are < 2^21 in magnitude, so ONE fma.rn produces them exactly, and
//     fma.rm(x, 2^-shift, 2^23) is floor(x / 2^shift) in the low mantissa bits (arithmetic shift, negative hue included);
//   * the hue numerator is min(c_b, c_g + 2048 (v - g), c_r + 2048 (v - r)) - the three OpenCV candidates are ordered
//     c_r <= c_g <= c_b for every colour, so penalising the channels that are not the maximum selects the right one with
//     no compare / select (r before g before b on ties, as OpenCV);
//   * hue + 180 -> slot = floor((h + 180) 2913 / 65536) = 8 + floor(h 8 / 180) for h >= 0 and 6 or 7 for the negative
//     hues OpenCV wraps by + 180: slots 6..14, folded onto the 8 hue bins when the CTA flushes;
//   * all float work on TWO pixels per instruction (add / fma.f32x2: one issue slot for both);
//   * the bit patterns 0x4B000000 + field are combined into the shared-memory address by three IMADs (mod 2^32);
//   * both tables are replicated per lane (entry v of lane l at word 32 v + l): every look-up is one conflict-free
//     wavefront and its address is ONE IMAD of the float's bit pattern.
// 25.5 issue slots per pixel (9 ALU, 22 FMA-pipe lane-cycles, 2 LDS, <= 1 RED).  Verified on all 2^24 colours
// (tests/test_gpu_parity.py::test_histogram_every_colour).
constexpr int kHsvComputeWarps = 14;                 // 224 x 224 pixels = 448 threads x 7 steps x 16 pixels: no ragged last step
constexpr int kHsvComputeThreads = kHsvComputeWarps * 32;
constexpr int kHsvThreads = kHsvComputeThreads + 32;  // + one warp that folds / stores / clears finished histograms
constexpr int kHsvHists = kHsvComputeWarps / 2;       // two warps share a set of counters
constexpr int kHsvSlotLo = 6;
constexpr int kHsvSlots = 9;                          // slots 6..14
constexpr int kHsvWarpBins = kHsvSlots * 64;
struct HsvSmem {
  float sdiv[256 * 32];
  float hdiv[256 * 32];
  uint32_t hist[2][kHsvHists][kHsvWarpBins];          // double-buffered: image i + 1 is counted while image i is flushed
  unsigned long long counted[2];                      // mbarrier: the 14 counting warps are done with buffer b
  unsigned long long cleared[2];                      // mbarrier: the flush warp has stored and cleared buffer b
};

using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2p(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fma2_floor(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
// + 1 on a shared counter: SASS ATOMS.POPC.INC (lanes of a warp that hit the same counter are counted, not replayed), one
// issue slot per pixel - a run-length scan of the 16-pixel group costs 4 slots per pixel, pairing neighbours 2.5
__device__ __forceinline__ void red_shared_inc(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory"); }
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ void hsv_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

struct HsvConsts {
  f32x2 m23, c2048, c2m17, c1_32, kv, c4, c2, cp, c2m12, mu, cc, kh;
  uint32_t tbl_s, tbl_h, hbase;
};

// shared-memory addresses of the histogram counters of two pixels given as 0x4B0000bb floats
// V: measurement variants (tests/tools/hsv_probe.cu): bit 0 = no shared atomics, bit 1 = no table look-ups
template <int VAR = 0>
__device__ __forceinline__ void hsv_addr2(float r0, float g0, float b0, float r1, float g1, float b1, const HsvConsts& k,
                                          uint32_t& a0, uint32_t& a1) {
  const float v0 = fmax3(r0, g0, b0), v1 = fmax3(r1, g1, b1);
  const float n0 = fmin3(r0, g0, b0), n1 = fmin3(r1, g1, b1);
  const f32x2 V = pk2(v0, v1), R = pk2(r0, r1), G = pk2(g0, g1), B = pk2(b0, b1);
  const f32x2 d = sub2p(V, pk2(n0, n1));                                  // v - min, exact
  uint32_t db0, db1;
  unpk2(add2(d, k.m23), db0, db1);                                        // 0x4B000000 + d
  float sd0, sd1, hd0, hd1;
  if constexpr (VAR & 2) {
    sd0 = __uint_as_float(__float_as_uint(v0) * 128u + k.tbl_s); sd1 = __uint_as_float(__float_as_uint(v1) * 128u + k.tbl_s);
    hd0 = __uint_as_float(db0 * 128u + k.tbl_h); hd1 = __uint_as_float(db1 * 128u + k.tbl_h);
  } else {
    sd0 = lds_f32(__float_as_uint(v0) * 128u + k.tbl_s); sd1 = lds_f32(__float_as_uint(v1) * 128u + k.tbl_s);
    hd0 = lds_f32(db0 * 128u + k.tbl_h); hd1 = lds_f32(db1 * 128u + k.tbl_h);
  }
  const f32x2 S = fma2_floor(fma2(d, pk2(sd0, sd1), k.c2048), k.c2m17, k.m23);      // 2^23 + ((d sdiv[v] + 2^11) >> 17)
  const f32x2 Vq = fma2_floor(V, k.c1_32, k.kv);                                      // 2^23 + (v >> 5)
  const f32x2 cb = fma2(k.c4, d, sub2p(R, G));                                        // r - g + 4 d
  const f32x2 cg = fma2(k.cp, sub2p(V, G), fma2(k.c2, d, sub2p(B, R)));               // b - r + 2 d  (+ 2048 (v - g))
  const f32x2 cr = fma2(k.cp, sub2p(V, R), sub2p(G, B));                              // g - b        (+ 2048 (v - r))
  uint32_t cb0, cb1, cg0, cg1, cr0, cr1;
  unpk2(cb, cb0, cb1); unpk2(cg, cg0, cg1); unpk2(cr, cr0, cr1);
  const float h0 = fmin3(__uint_as_float(cr0), __uint_as_float(cg0), __uint_as_float(cb0));
  const float h1 = fmin3(__uint_as_float(cr1), __uint_as_float(cg1), __uint_as_float(cb1));
  const f32x2 U = fma2_floor(fma2(pk2(h0, h1), pk2(hd0, hd1), k.c2048), k.c2m12, k.mu);   // 2^23 + 180 + h
  const f32x2 Hs = fma2_floor(U, k.cc, k.kh);                                          // 2^23 + slot
  uint32_t hs0, hs1, s0, s1, q0, q1;
  unpk2(Hs, hs0, hs1); unpk2(S, s0, s1); unpk2(Vq, q0, q1);
  // counter (slot, s, v) of a warp lives at word (slot - 6) + 9 s + 72 v: all three fields reach the bank index (v and s
  // are heavily skewed - a max and a ratio - and alone they send a third of the pixels to 4 of the 32 banks)
  a0 = q0 * 288u + (s0 * 36u + (hs0 * 4u + k.hbase));
  a1 = q1 * 288u + (s1 * 36u + (hs1 * 4u + k.hbase));
}

template <int VAR = 0>
__global__ void __launch_bounds__(kHsvThreads, 2) hsv_histogram_kernel(const uint8_t* __restrict__ img, int64_t pixels_per_image,
                                                                       int slices, int vector_ok, int64_t items,
                                                                       uint32_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char hsv_smem_raw[];
  HsvSmem& sm = *reinterpret_cast<HsvSmem*>(hsv_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 256 * 32; i += kHsvThreads) {
    sm.sdiv[i] = float(c_hsv_tables.sdiv[i >> 5]);
    sm.hdiv[i] = float(c_hsv_tables.hdiv[i >> 5]);
  }
  HsvConsts k;
  const float M = 8388608.f;
  k.m23 = pk2(M, M); k.c2048 = pk2(2048.f, 2048.f); k.c2m17 = pk2(0x1p-17f, 0x1p-17f); k.c1_32 = pk2(0x1p-5f, 0x1p-5f);
  k.kv = pk2(M - 262144.f, M - 262144.f); k.c4 = pk2(4.f, 4.f); k.c2 = pk2(2.f, 2.f); k.cp = k.c2048;
  k.c2m12 = pk2(0x1p-12f, 0x1p-12f); k.mu = pk2(M + 180.f, M + 180.f);
  k.cc = pk2(2913.f / 65536.f, 2913.f / 65536.f); k.kh = pk2(M - 372864.f, M - 372864.f);     // 2^23 * 2913 / 65536 = 372864
  k.tbl_s = uint32_t(__cvta_generic_to_shared(sm.sdiv)) + uint32_t(lane) * 4u - 0x80000000u;  // 0x4B000000 * 128 = 2^31 (mod 2^32)
  k.tbl_h = uint32_t(__cvta_generic_to_shared(sm.hdiv)) + uint32_t(lane) * 4u - 0x80000000u;
  const uint32_t hist_bytes = uint32_t(sizeof(uint32_t) * kHsvHists * kHsvWarpBins);
  k.hbase = uint32_t(__cvta_generic_to_shared(sm.hist[0][(warp >> 1) % kHsvHists])) - uint32_t(kHsvSlotLo * 4) - 0x4B000000u * 328u;
  // a shuffle from the own lane: one register each that ptxas cannot take apart again (it re-associated the folded
  // constants into an extra IADD per look-up)
  k.tbl_s = __shfl_sync(0xffffffffu, k.tbl_s, lane);
  k.tbl_h = __shfl_sync(0xffffffffu, k.tbl_h, lane);
  k.hbase = __shfl_sync(0xffffffffu, k.hbase, lane);
  const uint32_t hbase0 = k.hbase;

  for (int i = tid; i < 2 * kHsvHists * kHsvWarpBins; i += kHsvThreads) (&sm.hist[0][0][0])[i] = 0;
  const uint32_t bar_counted = uint32_t(__cvta_generic_to_shared(&sm.counted[0]));
  const uint32_t bar_cleared = uint32_t(__cvta_generic_to_shared(&sm.cleared[0]));
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { tma::mbar_init(bar_counted + 8 * b, kHsvComputeWarps); tma::mbar_init(bar_cleared + 8 * b, 1); }
    tma::mbar_fence_init();
  }
  __syncthreads();                                                       // tables, zeroed counters and mbarriers ready

  // No CTA-wide barrier in the loop (ncu on the single-buffer version: 17 % of the resident warp time at __syncthreads):
  // every counting warp arrives on counted[b] when it is done with image i and goes on to image i + 1 in the other
  // buffer; the flush warp waits for the 14 arrivals, folds / stores / clears buffer b and arrives on cleared[b], which
  // a counting warp only looks at two images later.
  if (warp == kHsvComputeWarps) {
    int it = 0;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = it & 1;
      tma::mbar_wait(bar_counted + 8 * b, (it >> 1) & 1);
      uint32_t* dst = out + (item / slices) * kHistBins;
      for (int bin = lane; bin < kHistBins; bin += 32) {                 // fold the slots onto the 8 hue bins; leave zeros behind
        const int hb = bin >> 6, rest = 9 * ((bin >> 3) & 7) + 72 * (bin & 7);
        uint32_t sum = 0;
        if (hb + 8 - kHsvSlotLo < kHsvSlots) {
#pragma unroll
          for (int h = 0; h < kHsvHists; ++h) { uint32_t& c = sm.hist[b][h][hb + 8 - kHsvSlotLo + rest]; sum += c; c = 0; }
        }
        if (hb >= kHsvSlotLo) {                                          // negative hues (OpenCV adds 180): slots 6 and 7
#pragma unroll
          for (int h = 0; h < kHsvHists; ++h) { uint32_t& c = sm.hist[b][h][hb - kHsvSlotLo + rest]; sum += c; c = 0; }
        }
        if (slices == 1) dst[bin] = sum;
        else if (sum) atomicAdd(&dst[bin], sum);
      }
      __syncwarp();
      if (lane == 0) hsv_mbar_arrive(bar_cleared + 8 * b);
    }
    return;
  }

  int it = 0;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++it) {
    const int b = it & 1;
    if (it >= 2) tma::mbar_wait(bar_cleared + 8 * b, ((it >> 1) - 1) & 1);
    k.hbase = hbase0 + uint32_t(b) * hist_bytes;
    const int64_t image = item / slices;
    const int slice = int(item % slices);
    const uint8_t* base = img + image * pixels_per_image * 3;
    if (vector_ok) {
      const int64_t groups = pixels_per_image / 16;
      const int64_t per_slice = ceil_div64(groups, slices);
      const int64_t g_begin = slice * per_slice, g_end = min(groups, g_begin + per_slice);
      const uint4* v = reinterpret_cast<const uint4*>(base);
      int64_t gi = g_begin + tid;
      uint32_t acc = 0;
      uint4 na, nb, nc;
      if (gi < g_end) { na = __ldg(v + gi * 3); nb = __ldg(v + gi * 3 + 1); nc = __ldg(v + gi * 3 + 2); }
      while (gi < g_end) {
        uint32_t w[12];
        w[0] = na.x; w[1] = na.y; w[2] = na.z; w[3] = na.w; w[4] = nb.x; w[5] = nb.y; w[6] = nb.z; w[7] = nb.w;
        w[8] = nc.x; w[9] = nc.y; w[10] = nc.z; w[11] = nc.w;
        gi += kHsvComputeThreads;
        if (gi < g_end) { na = __ldg(v + gi * 3); nb = __ldg(v + gi * 3 + 1); nc = __ldg(v + gi * 3 + 2); }   // next step's pixels in flight
#pragma unroll
        for (int p = 0; p < 16; p += 2) {
          float ch[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int byte = p * 3 + j;
            ch[j] = __uint_as_float(__byte_perm(w[byte >> 2], 0x4B000000u, 0x7440 | (byte & 3)));    // 2^23 + byte
          }
          uint32_t a0, a1;
          hsv_addr2<VAR>(ch[0], ch[1], ch[2], ch[3], ch[4], ch[5], k, a0, a1);
          if constexpr (VAR & 1) { acc ^= a0 + a1; }
          else { red_shared_inc(a0); red_shared_inc(a1); }
        }
      }
      if constexpr (VAR & 1) { if (acc == 0x12345u) red_shared_inc(hbase0 + 0x4B000000u * 328u + uint32_t(kHsvSlotLo * 4)); }
    } else {
      const int64_t per_slice = ceil_div64(pixels_per_image, slices);
      const int64_t p_begin = slice * per_slice, p_end = min(pixels_per_image, p_begin + per_slice);
      for (int64_t pi = p_begin + tid; pi < p_end; pi += kHsvComputeThreads) {
        const uint8_t* px = base + pi * 3;
        const float r = __uint_as_float(0x4B000000u | px[0]), g = __uint_as_float(0x4B000000u | px[1]),
                    b2 = __uint_as_float(0x4B000000u | px[2]);
        uint32_t a0, a1;
        hsv_addr2(r, g, b2, r, g, b2, k, a0, a1);
        red_shared_inc(a0);
      }
    }
    __syncwarp();
    if (lane == 0) hsv_mbar_arrive(bar_counted + 8 * b);
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
