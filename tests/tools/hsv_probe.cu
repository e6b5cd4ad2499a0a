// Probe (not part of the library): where does the HSV histogram kernel's time go?  Times the kernel with its shared
// atomics and / or its table look-ups compiled out.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -I image-retrieval-_b200/csrc \
//        -o image-retrieval-_b200/build/hsv_probe tests/tools/hsv_probe.cu
#include <cstdio>
#include <cstdlib>
#include "histogram.cuh"
using namespace b200ir;
template <int V>
float run(const uint8_t* img, int64_t B, uint32_t* out) {
  cudaFuncSetAttribute(hsv_histogram_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(HsvSmem)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    hsv_histogram_kernel<V><<<296, kHsvThreads, sizeof(HsvSmem)>>>(img, 224 * 224, 1, 1, B, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}
int main() {
  const int64_t B = 8192, bytes = B * 224 * 224 * 3;
  uint8_t* h = (uint8_t*)malloc(bytes);
  unsigned x = 12345;
  for (int64_t i = 0; i < bytes; ++i) { x = x * 1664525u + 1013904223u; h[i] = uint8_t(x >> 24); }
  uint8_t* d; uint32_t* out;
  cudaMalloc(&d, bytes); cudaMalloc(&out, B * 512 * 4);
  cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
  init_hsv_tables();
  printf("full              %.3f ms\n", run<0>(d, B, out));
  printf("no atomics        %.3f ms\n", run<1>(d, B, out));
  printf("no tables         %.3f ms\n", run<2>(d, B, out));
  printf("no atomics/tables %.3f ms\n", run<3>(d, B, out));
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
