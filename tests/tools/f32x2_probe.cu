// Probe (not part of the library): does sm_100a have packed fp32 FMA/ADD (fma.rn.f32x2 / add.rn.f32x2) and what is its
// throughput relative to scalar FFMA?  nvcc -gencode arch=compute_100a,code=sm_100a -o f32x2_probe f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float m = 1.0001f, c = 1e-4f;
  if (MODE == 0) {
    for (int i = 0; i < iters; ++i) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  } else {
    unsigned long long p0, p1, p2, p3, mm, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    for (int i = 0; i < iters; ++i) { p0 = ffma2(p0, mm, cc); p1 = ffma2(p1, mm, cc); p2 = ffma2(p2, mm, cc); p3 = ffma2(p3, mm, cc); }
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 100000;
  for (int mode = 0; mode < 2; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters); else k<1><<<148 * 8, 256>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = double(148) * 8 * 256 * 8.0 * iters;
    printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", mode ? "fma.rn.f32x2" : "scalar fmaf ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
  }
  return 0;
}
