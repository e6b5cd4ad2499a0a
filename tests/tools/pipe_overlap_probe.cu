// Probe (not part of the library): do FMA-pipe (FFMA2 / FFMA / IMAD) and ALU-pipe (PRMT / FMNMX3 / LOP3) instructions of
// one warp scheduler overlap on sm_100a, or do their pipe cycles add up?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o image-retrieval-_b200/build/pipe_probe tests/tools/pipe_overlap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
using u64 = unsigned long long;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned s) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned imad(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// MODE bits: 1 = 8 FFMA2, 2 = 8 PRMT, 4 = 8 FMNMX3, 8 = 8 scalar FFMA, 16 = 8 IMAD (per iteration, 8 independent chains each)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, unsigned sel, unsigned mulv) {
  u64 p[8]; unsigned q[8]; float f[8]; float g[8]; unsigned im[8];
  for (int i = 0; i < 8; ++i) { p[i] = threadIdx.x * 77u + i; q[i] = threadIdx.x * 31u + i; f[i] = threadIdx.x * 0.5f + i; g[i] = f[i] + 1.f; im[i] = q[i] ^ 5u; }
  u64 mm = 0x3f8000013f800001ull, cc = 0x3a8000003a800000ull;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE & 1) p[i] = ffma2(p[i], mm, cc);
      if (MODE & 2) q[i] = prmt(q[i], q[(i + 1) & 7], sel);
      if (MODE & 4) f[i] = fmin3(f[i], f[(i + 3) & 7], f[(i + 5) & 7]);
      if (MODE & 8) g[i] = ffma(g[i], 1.0001f, 1e-4f);
      if (MODE & 16) im[i] = imad(im[i], mulv, im[(i + 1) & 7]);
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += float(p[i] & 0xffff) + float(q[i] & 0xff) + f[i] + g[i] + float(im[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* out, int ninst) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, iters, 0x5140u + rep, 3u + rep);       // 32 warps per SM = 8 per scheduler
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  // cycles per scheduler per iteration-of-8-warps: ms * 1.965e6 / iters / 8 warps
  const double cyc = best * 1.965e6 / iters / 8.0;
  printf("%-34s %.3f ms   %.1f cycles per warp-iteration (%d instr) -> %.2f cycles per instruction\n", name, best, cyc, ninst, cyc / ninst);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  run<1>("8 FFMA2", out, 8);
  run<8>("8 FFMA", out, 8);
  run<16>("8 IMAD", out, 8);
  run<2>("8 PRMT", out, 8);
  run<4>("8 FMNMX3", out, 8);
  run<6>("8 PRMT + 8 FMNMX3", out, 16);
  run<3>("8 FFMA2 + 8 PRMT", out, 16);
  run<5>("8 FFMA2 + 8 FMNMX3", out, 16);
  run<7>("8 FFMA2 + 8 PRMT + 8 FMNMX3", out, 24);
  run<10>("8 FFMA + 8 PRMT", out, 16);
  run<18>("8 IMAD + 8 PRMT", out, 16);
  run<17>("8 FFMA2 + 8 IMAD", out, 16);
  run<23>("8 FFMA2 + 8 IMAD + 8 PRMT + 8 FMNMX3", out, 32);
  run<9>("8 FFMA2 + 8 FFMA", out, 16);
  run<24>("8 FFMA + 8 IMAD", out, 16);
  run<12>("8 FFMA + 8 FMNMX3", out, 16);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
