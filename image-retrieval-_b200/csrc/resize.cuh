// Image front-end of the embedding producer: bicubic resize + centre crop of uint8 RGB images, bit-exact with the
// PIL path the reference's CLIPProcessor takes (ImageEmbeddingSystem.py:82-83; Pillow libImaging/Resample.c:
// separable, horizontal pass first into uint8, 22-bit fixed-point coefficients, clip8((2^21 + sum) >> 22)).
//
// One fused kernel: a CTA owns `rows_per_block` output rows of one image.  It streams the input rows its vertical
// windows touch through a small row buffer (16-byte loads, kRowBatch rows per barrier pair), resamples each row
// horizontally for the cropped columns only into a shared uint8 tile, then runs the vertical pass out of that tile.
// The intermediate image never reaches HBM; source bytes are read once per CTA (neighbouring CTAs re-read the
// ~2*support overlap rows, normally out of L2).  Coefficient tables are built on the host in double, exactly as
// Pillow does, and live in the caller's workspace.
#pragma once
#include "common.cuh"

#include <math.h>
#include <vector>

namespace b200ir {

constexpr int kResizePrecisionBits = 32 - 8 - 2;
constexpr int kResizeThreads = 256;
constexpr int kRowBatch = 8;
#ifndef RESIZE_RPB
#define RESIZE_RPB 16
#endif
constexpr int kResizeRowsPerBlock = RESIZE_RPB;   // output rows per CTA (largest tried first)
constexpr int kResizeSmemBudget = 160 * 1024;

struct ResampleTable {
  int ksize = 0;
  std::vector<int32_t> bounds;   // [count][2]: first input coordinate, tap count
  std::vector<int32_t> kk;       // [count][ksize]
  int in_lo = 0, in_hi = 0;      // input span [in_lo, in_hi) touched by the table
};

inline double bicubic_weight(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Coefficients of output coordinates [first, first + count) when `in_size` samples are resampled to `out_size`.
inline ResampleTable make_resample_table(int in_size, int out_size, int first, int count) {
  ResampleTable t;
  t.bounds.resize(size_t(count) * 2);
  if (in_size == out_size) {                      // Pillow skips the pass: identity taps
    t.ksize = 1;
    t.kk.assign(size_t(count), 1 << kResizePrecisionBits);
    for (int i = 0; i < count; ++i) { t.bounds[2 * i] = first + i; t.bounds[2 * i + 1] = 1; }
    t.in_lo = first; t.in_hi = first + count;
    return t;
  }
  const double scale = double(in_size) / double(out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  t.ksize = int(ceil(support)) * 2 + 1;
  t.kk.assign(size_t(count) * t.ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(size_t(t.ksize));
  t.in_lo = in_size; t.in_hi = 0;
  for (int i = 0; i < count; ++i) {
    const int xx = first + i;
    const double center = (xx + 0.5) * scale;
    int xmin = int(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = int(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_weight((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      t.kk[size_t(i) * t.ksize + x] = v < 0 ? int(-0.5 + v * (1 << kResizePrecisionBits)) : int(0.5 + v * (1 << kResizePrecisionBits));
    }
    t.bounds[2 * i] = xmin;
    t.bounds[2 * i + 1] = xmax;
    if (xmin < t.in_lo) t.in_lo = xmin;
    if (xmin + xmax > t.in_hi) t.in_hi = xmin + xmax;
  }
  return t;
}

struct ResizePlan {
  ResampleTable tx, ty;
  int rows_per_block = 0, max_rows = 0, row_buf_bytes = 0;
  size_t smem_bytes = 0, table_bytes = 0;
  size_t off_xb = 0, off_xk = 0, off_yb = 0, off_yk = 0;
  bool ok = false;
};

inline ResizePlan make_resize_plan(int H, int W, int rh, int rw, int top, int left, int ch, int cw) {
  ResizePlan p;
  p.tx = make_resample_table(W, rw, left, cw);
  p.ty = make_resample_table(H, rh, top, ch);
  p.row_buf_bytes = int(round_up64(int64_t(p.tx.in_hi - p.tx.in_lo) * 3 + 16 + 48, 16));   // +48: fixed-trip tap loops read past the span
  const int64_t out_row = int64_t(cw) * 3;
  // the largest group of output rows whose input-row window (tile) + row buffers fit the shared-memory budget
  for (int rpb = kResizeRowsPerBlock; rpb >= 1; rpb >>= 1) {
    int max_rows = 0;
    for (int y0 = 0; y0 < ch; y0 += rpb) {
      const int y1 = (y0 + rpb < ch ? y0 + rpb : ch) - 1;
      const int rows = p.ty.bounds[2 * y1] + p.ty.bounds[2 * y1 + 1] - p.ty.bounds[2 * y0];
      if (rows > max_rows) max_rows = rows;
    }
    // tile rows + 3: the 4-tap groups of the vertical pass read up to 3 rows past a window (against zero coefficients);
    // vertical table per output row: first row, tap count, then 3 signed-byte limbs per group of 4 taps
    const size_t smem = size_t(2 * kRowBatch) * p.row_buf_bytes + size_t(max_rows + 3) * size_t(round_up64(out_row, 4)) +
                        size_t(rpb) * (2 + 3 * ((p.ty.ksize + 3) / 4)) * 4;
    if (smem <= size_t(kResizeSmemBudget)) {
      p.rows_per_block = rpb; p.max_rows = max_rows; p.smem_bytes = smem; p.ok = true;
      break;
    }
  }
  size_t off = 0;
  auto take = [&off](size_t n) { size_t o = off; off += size_t(round_up64(int64_t(n), 256)); return o; };
  p.off_xb = take(p.tx.bounds.size() * 4);
  p.off_xk = take(p.tx.kk.size() * 4);
  p.off_yb = take(p.ty.bounds.size() * 4);
  p.off_yk = take(p.ty.kk.size() * 4);
  p.table_bytes = off;
  return p;
}

struct ResizeArgs {
  const uint8_t* img; uint8_t* out;
  int64_t img_stride;              // bytes per input image
  int H, W, cw, ch, kx, ky, rows_per_block, row_buf_bytes, x_lo, x_hi, tile_pitch, max_rows, word_out;
  const int32_t *xb, *xk, *yb, *yk;
};

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kResizePrecisionBits;
  return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// A 22-bit signed coefficient as three signed-byte limbs, k = l0 + 2^8 l1 + 2^16 l2 (balanced digits), so that four 8-bit
// samples meet four taps in one dp4a per limb: sum(p k) = s0 + 2^8 s1 + 2^16 s2 exactly (every s fits 20 bits).
__device__ __forceinline__ void coeff_limbs(int k, int& l0, int& l1, int& l2) {
  l0 = ((k + 128) & 255) - 128;
  const int k1 = (k - l0) >> 8;
  l1 = ((k1 + 128) & 255) - 128;
  l2 = (k1 - l1) >> 8;
}
__device__ __forceinline__ int dp4a_u8s8(uint32_t samples, int limbs, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(samples), "r"(limbs), "r"(acc));
  return d;
}
__device__ __forceinline__ int pack_limb(int a, int b, int c, int d) {
  return (a & 255) | ((b & 255) << 8) | ((c & 255) << 16) | ((d & 255) << 24);
}

// KREG > 0: horizontal taps of "this thread's" output column live in KREG registers (kx <= KREG, cw <= kResizeThreads;
// table entries past the tap count are zero and the row buffers are padded, so the tap loop has a fixed trip count).
// KREG == 0: any geometry, taps read through the read-only cache.
template <int KREG>
__global__ void __launch_bounds__(kResizeThreads) resize_crop_kernel(ResizeArgs a) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint8_t* rowbuf_base = rs_smem;                                // 2 x kRowBatch x row_buf_bytes (double-buffered)
  uint8_t* tile = rs_smem + 2 * kRowBatch * a.row_buf_bytes;     // max_rows x tile_pitch (horizontally resampled rows)
  const int kgy = (a.ky + 3) >> 2;                               // groups of four vertical taps
  const int kstride = 2 + 3 * kgy;
  int32_t* ksm = reinterpret_cast<int32_t*>(tile + (a.max_rows + 3) * a.tile_pitch);   // rows_per_block x (2 + 3 kgy): vertical limbs
  const int tid = threadIdx.x;
  const int y0 = blockIdx.x * a.rows_per_block;
  const int y1 = min(a.ch, y0 + a.rows_per_block);
  const int r0 = a.yb[2 * y0];
  const int r1 = a.yb[2 * (y1 - 1)] + a.yb[2 * (y1 - 1) + 1];
  const uint8_t* img = a.img + int64_t(blockIdx.y) * a.img_stride;
  const int out_row = a.cw * 3;
  const int64_t seg_off = int64_t(a.x_lo) * 3;
  const int seg_bytes = (a.x_hi - a.x_lo) * 3;

  for (int i = tid; i < (y1 - y0) * kstride; i += kResizeThreads) {
    const int yy = i / kstride, j = i - yy * kstride;
    if (j < 2) ksm[i] = a.yb[2 * (y0 + yy) + j];
    else {
      const int g = (j - 2) / 3, l = (j - 2) - 3 * g;
      int lim[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int tap = 4 * g + t;
        int l3[3];
        coeff_limbs(tap < a.ky ? a.yk[(y0 + yy) * a.ky + tap] : 0, l3[0], l3[1], l3[2]);
        lim[t] = l == 0 ? l3[0] : (l == 1 ? l3[1] : l3[2]);
      }
      ksm[i] = pack_limb(lim[0], lim[1], lim[2], lim[3]);
    }
  }
  constexpr int KG = KREG > 0 ? (KREG + 3) / 4 : 1;              // groups of four horizontal taps
  int klim[KG][3];                                               // limbs of "this thread's" taps, four taps per register
  int my_off = 0;
  if (KREG > 0 && tid < a.cw) {
#pragma unroll
    for (int g = 0; g < KG; ++g) {
      int l0[4], l1[4], l2[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) coeff_limbs(4 * g + t < a.kx ? a.xk[tid * a.kx + 4 * g + t] : 0, l0[t], l1[t], l2[t]);
      klim[g][0] = pack_limb(l0[0], l0[1], l0[2], l0[3]);
      klim[g][1] = pack_limb(l1[0], l1[1], l1[2], l1[3]);
      klim[g][2] = pack_limb(l2[0], l2[1], l2[2], l2[3]);
    }
    my_off = (a.xb[2 * tid] - a.x_lo) * 3;
  }

  // Input rows reach shared memory by cp.async, one batch ahead: batch i + 1 is in flight while batch i is resampled
  // (one barrier per batch; the single-buffer version waited for every batch's loads with nothing else to do).
  auto stage_rows = [&](int rb, uint8_t* rowbuf) {
    const int nrows = min(kRowBatch, r1 - rb);
    // the needed byte span of up to kRowBatch input rows, one warp per row; 16-byte copies from the aligned address
    // below the span
    for (int r = tid >> 5; r < nrows; r += kResizeThreads / 32) {
      const uint8_t* src = img + (int64_t(rb + r) * a.W) * 3 + seg_off;
      const int mis = int(reinterpret_cast<uintptr_t>(src) & 15);
      const uint4* src16 = reinterpret_cast<const uint4*>(src - mis);
      uint4* dst16 = reinterpret_cast<uint4*>(rowbuf + r * a.row_buf_bytes);
      const int n16 = (mis + seg_bytes + 15) >> 4;
      // the last 16-byte load may run past the image; it stays inside the allocation except for the final row of the
      // final image, which is read bytewise
      const bool tail_safe = rb + r + 1 < a.H || blockIdx.y + 1 < gridDim.y;
      const int n_fast = tail_safe ? n16 : n16 - 1;
      for (int i = tid & 31; i < n_fast; i += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(uint32_t(__cvta_generic_to_shared(dst16 + i))), "l"(src16 + i) : "memory");
      if (!tail_safe && (tid & 31) == 0) {
        uint8_t* d = reinterpret_cast<uint8_t*>(dst16 + n_fast);
        const uint8_t* s = reinterpret_cast<const uint8_t*>(src16 + n_fast);
        const int valid = mis + seg_bytes - n_fast * 16;
        for (int b = 0; b < 16; ++b) d[b] = b < valid ? s[b] : 0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_rows(r0, rowbuf_base);
  int parity = 0;
  for (int rb = r0; rb < r1; rb += kRowBatch, parity ^= 1) {
    const int nrows = min(kRowBatch, r1 - rb);
    uint8_t* rowbuf = rowbuf_base + parity * kRowBatch * a.row_buf_bytes;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                             // batch `rb` has landed; everybody is done with the other buffer
    if (rb + kRowBatch < r1) stage_rows(rb + kRowBatch, rowbuf_base + (parity ^ 1) * kRowBatch * a.row_buf_bytes);
    if (KREG > 0) {
      if (tid < a.cw) {
        // one input row of "this thread's" output column; `off` = byte offset of its window in the staged row
        auto hrow = [&](int r, int off) {
          // the 12 KG bytes of this pixel's window: aligned 32-bit shared loads, re-aligned with one PRMT per word
          // (byte loads would make the kernel LSU-bound: 3 shared loads per tap)
          const uint32_t* pw = reinterpret_cast<const uint32_t*>(rowbuf + r * a.row_buf_bytes + (off & ~3));
          const uint32_t sel = 0x3210u + 0x1111u * uint32_t(off & 3);
          constexpr int NW = 3 * KG;                               // 12 bytes = 4 taps x 3 channels per group
          uint32_t w[NW + 1];
#pragma unroll
          for (int i = 0; i <= NW; ++i) w[i] = pw[i];
#pragma unroll
          for (int i = 0; i < NW; ++i) w[i] = __byte_perm(w[i], w[i + 1], sel);
          // per group and channel: gather the channel's four samples (two PRMTs), then one dp4a per coefficient limb
          // (round 1: one PRMT + one IMAD per sample and tap; taps past kx and bytes past the window meet zero limbs)
          const int half = 1 << (kResizePrecisionBits - 1);
          int s0[3] = {half, half, half}, s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
#pragma unroll
          for (int g = 0; g < KG; ++g) {
            const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
            const uint32_t c0 = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);     // bytes 0, 3, 6, 9
            const uint32_t c1 = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);     // bytes 1, 4, 7, 10
            const uint32_t c2 = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);     // bytes 2, 5, 8, 11
            s0[0] = dp4a_u8s8(c0, klim[g][0], s0[0]); s1[0] = dp4a_u8s8(c0, klim[g][1], s1[0]); s2[0] = dp4a_u8s8(c0, klim[g][2], s2[0]);
            s0[1] = dp4a_u8s8(c1, klim[g][0], s0[1]); s1[1] = dp4a_u8s8(c1, klim[g][1], s1[1]); s2[1] = dp4a_u8s8(c1, klim[g][2], s2[1]);
            s0[2] = dp4a_u8s8(c2, klim[g][0], s0[2]); s1[2] = dp4a_u8s8(c2, klim[g][1], s1[2]); s2[2] = dp4a_u8s8(c2, klim[g][2], s2[2]);
          }
          const int acc0 = s0[0] + (s1[0] << 8) + (s2[0] << 16);
          const int acc1 = s0[1] + (s1[1] << 8) + (s2[1] << 16);
          const int acc2 = s0[2] + (s1[2] << 8) + (s2[2] << 16);
          uint8_t* t = tile + (rb - r0 + r) * a.tile_pitch + tid * 3;
          t[0] = clip8(acc0); t[1] = clip8(acc1); t[2] = clip8(acc2);
        };
        if (((a.W * 3) & 15) == 0) {
          // the row pitch is a multiple of 16 bytes (every width that is a multiple of 16 pixels): the staging misalignment
          // and with it the window offset / PRMT selector are the same for every row - computed once, two rows per trip
          const int off = int((reinterpret_cast<uintptr_t>(img) + uintptr_t(seg_off)) & 15) + my_off;
          int r = 0;
          for (; r + 1 < nrows; r += 2) { hrow(r, off); hrow(r + 1, off); }
          if (r < nrows) hrow(r, off);
        } else {
          for (int r = 0; r < nrows; ++r) {
            const uint8_t* src = img + (int64_t(rb + r) * a.W) * 3 + seg_off;
            hrow(r, int(reinterpret_cast<uintptr_t>(src) & 15) + my_off);
          }
        }
      }
    } else {
      for (int o = tid; o < nrows * a.cw; o += kResizeThreads) {      // one thread = one pixel (3 channels share the taps)
        const int r = o / a.cw, xx = o - r * a.cw;
        const uint8_t* src = img + (int64_t(rb + r) * a.W) * 3 + seg_off;
        const int mis = int(reinterpret_cast<uintptr_t>(src) & 15);
        const int xmin = a.xb[2 * xx], n = a.xb[2 * xx + 1];
        const uint8_t* px = rowbuf + r * a.row_buf_bytes + mis + (xmin - a.x_lo) * 3;
        const int32_t* k = a.xk + xx * a.kx;
        int acc0 = 1 << (kResizePrecisionBits - 1), acc1 = acc0, acc2 = acc0;
        for (int j = 0; j < n; ++j) {
          const int w = __ldg(k + j);
          acc0 += int(px[j * 3]) * w;
          acc1 += int(px[j * 3 + 1]) * w;
          acc2 += int(px[j * 3 + 2]) * w;
        }
        uint8_t* t = tile + (rb - r0 + r) * a.tile_pitch + xx * 3;
        t[0] = clip8(acc0); t[1] = clip8(acc1); t[2] = clip8(acc2);
      }
    }
  }
  __syncthreads();                                               // the tile is complete
  // vertical pass out of the tile: one thread = 4 consecutive bytes of an output row, taps broadcast from shared memory
  uint8_t* out = a.out + (int64_t(blockIdx.y) * a.ch + y0) * out_row;
  if (a.word_out) {
    const int words = out_row >> 2, pitch4 = a.tile_pitch >> 2;
    const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(tile);
    // work items (output row yy, word wd) dealt round-robin over all 256 threads: a 224-pixel row has 168 words, so one
    // thread per word would leave a third of the CTA idle for the whole pass.  (wd, yy) advance without a division.
    const int step_y = kResizeThreads / words, step_w = kResizeThreads - step_y * words;
    int wd = tid % words, yy = tid / words;
    for (; yy < y1 - y0; ) {
      {
        const int32_t* k = ksm + yy * kstride;
        const int n = k[1];
        const uint32_t* px = tile32 + (k[0] - r0) * pitch4 + wd;
        int s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
        for (int g = 0; 4 * g < n; ++g) {
          // four rows x four columns: transpose (8 PRMTs) so that each column's four samples share a register
          const uint32_t r0w = px[(4 * g) * pitch4], r1w = px[(4 * g + 1) * pitch4], r2w = px[(4 * g + 2) * pitch4],
                         r3w = px[(4 * g + 3) * pitch4];
          const uint32_t t0 = __byte_perm(r0w, r1w, 0x5140), t1 = __byte_perm(r2w, r3w, 0x5140);
          const uint32_t t2 = __byte_perm(r0w, r1w, 0x7362), t3 = __byte_perm(r2w, r3w, 0x7362);
          const uint32_t col[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632), __byte_perm(t2, t3, 0x5410),
                                   __byte_perm(t2, t3, 0x7632)};
          const int l0 = k[2 + 3 * g], l1 = k[3 + 3 * g], l2 = k[4 + 3 * g];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            s0[c] = dp4a_u8s8(col[c], l0, s0[c]); s1[c] = dp4a_u8s8(col[c], l1, s1[c]); s2[c] = dp4a_u8s8(col[c], l2, s2[c]);
          }
        }
        const int half = 1 << (kResizePrecisionBits - 1);
        uint32_t outw = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) outw |= uint32_t(clip8(half + s0[c] + (s1[c] << 8) + (s2[c] << 16))) << (8 * c);
        reinterpret_cast<uint32_t*>(out)[yy * words + wd] = outw;
      }
      wd += step_w; yy += step_y;
      if (wd >= words) { wd -= words; ++yy; }
    }
  } else {
    for (int o = tid; o < (y1 - y0) * out_row; o += kResizeThreads) {
      const int yy = o / out_row, e = o - yy * out_row;
      const int32_t* k = ksm + yy * kstride;
      const int n = k[1];
      const uint8_t* px = tile + (k[0] - r0) * a.tile_pitch + e;
      const int32_t* kt = a.yk + (y0 + yy) * a.ky;
      int acc = 1 << (kResizePrecisionBits - 1);
      for (int j = 0; j < n; ++j) acc += int(px[j * a.tile_pitch]) * __ldg(kt + j);
      out[o] = clip8(acc);
    }
  }
}

template <int KREG>
inline cudaError_t launch_resize_crop(const ResizeArgs& a, int groups, int nb, size_t smem, cudaStream_t st) {
  // per launch: the attribute belongs to the current device / context (a process may use several GPUs) and is cheap to set
  cudaError_t e = cudaFuncSetAttribute(resize_crop_kernel<KREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, kResizeSmemBudget);
  if (e != cudaSuccess) return e;
  resize_crop_kernel<KREG><<<dim3(groups, nb), kResizeThreads, smem, st>>>(a);
  return cudaGetLastError();
}

inline cudaError_t run_resize_crop(const ResizePlan& p, const uint8_t* img, int64_t B, int H, int W, int ch, int cw,
                                   uint8_t* out, unsigned char* ws, cudaStream_t st) {
  cudaError_t e;
  // pageable -> device async copies are staged before the call returns, so the host vectors may die with the plan
  if ((e = cudaMemcpyAsync(ws + p.off_xb, p.tx.bounds.data(), p.tx.bounds.size() * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(ws + p.off_xk, p.tx.kk.data(), p.tx.kk.size() * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(ws + p.off_yb, p.ty.bounds.data(), p.ty.bounds.size() * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(ws + p.off_yk, p.ty.kk.data(), p.ty.kk.size() * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  ResizeArgs a;
  a.img = img; a.out = out;
  a.img_stride = int64_t(H) * W * 3;
  a.H = H; a.W = W; a.cw = cw; a.ch = ch; a.kx = p.tx.ksize; a.ky = p.ty.ksize;
  a.rows_per_block = p.rows_per_block; a.row_buf_bytes = p.row_buf_bytes;
  a.x_lo = p.tx.in_lo; a.x_hi = p.tx.in_hi;
  a.tile_pitch = int(round_up64(int64_t(cw) * 3, 4));
  a.xb = reinterpret_cast<const int32_t*>(ws + p.off_xb);
  a.xk = reinterpret_cast<const int32_t*>(ws + p.off_xk);
  a.yb = reinterpret_cast<const int32_t*>(ws + p.off_yb);
  a.yk = reinterpret_cast<const int32_t*>(ws + p.off_yk);
  a.max_rows = p.max_rows;
  a.word_out = (cw * 3) % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0 ? 1 : 0;
  const int kreg = cw <= kResizeThreads ? (a.kx <= 5 ? 5 : a.kx <= 7 ? 7 : a.kx <= 9 ? 9 : a.kx <= 11 ? 11 : a.kx <= 13 ? 13 : a.kx <= 15 ? 15 : 0) : 0;
  const int groups = int(ceil_div64(ch, p.rows_per_block));
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {          // gridDim.y limit
    const int nb = int(B - b0 < 65535 ? B - b0 : 65535);
    ResizeArgs ab = a;
    ab.img = img + b0 * a.img_stride;
    ab.out = out + b0 * int64_t(ch) * cw * 3;
    switch (kreg) {
      case 5: e = launch_resize_crop<5>(ab, groups, nb, p.smem_bytes, st); break;
      case 7: e = launch_resize_crop<7>(ab, groups, nb, p.smem_bytes, st); break;
      case 9: e = launch_resize_crop<9>(ab, groups, nb, p.smem_bytes, st); break;
      case 11: e = launch_resize_crop<11>(ab, groups, nb, p.smem_bytes, st); break;
      case 13: e = launch_resize_crop<13>(ab, groups, nb, p.smem_bytes, st); break;
      case 15: e = launch_resize_crop<15>(ab, groups, nb, p.smem_bytes, st); break;
      default: e = launch_resize_crop<0>(ab, groups, nb, p.smem_bytes, st); break;
    }
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace b200ir
