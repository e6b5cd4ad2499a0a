// Launch accounting + optional CUDA-event timing of individual kernels (used by bench.py to get
// the dominant kernel's duration on the launching stream, and to count launches per step).
#pragma once
#include <cuda_runtime.h>

namespace b200ir {

enum ProfileTag { PT_PREP = 0, PT_SCAN = 1, PT_GEMM = 2, PT_FINALIZE = 3, PT_RERANK = 4, PT_MERGE = 5, PT_HIST = 6,
                  PT_MISC = 7, PT_RESIZE = 8, PT_COUNT = 9 };

void profile_begin(int tag, cudaStream_t st);   // counts the launch; records an event when enabled
void profile_end(int tag, cudaStream_t st);

struct ProfileScope {
  int tag; cudaStream_t st;
  ProfileScope(int t, cudaStream_t s) : tag(t), st(s) { profile_begin(tag, st); }
  ~ProfileScope() { profile_end(tag, st); }
};

}  // namespace b200ir
