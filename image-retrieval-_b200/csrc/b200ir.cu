// C ABI of the retrieval hot path (declared in include/b200ir.h).  Argument validation, path
// selection and kernel launches only: all arithmetic lives in the .cuh kernels.
#include "common.cuh"
#include "scan_plan.h"
#include "select.cuh"
#include "histogram.cuh"
#include "pairs.cuh"
#include "resize.cuh"
#include "postfilter.cuh"
#include "rank.cuh"
#include "gemm_topk.h"
#include "profile.h"

#include <atomic>
#include <mutex>
#include <vector>

using namespace b200ir;

namespace b200ir {
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_profile_on{0};
static std::mutex g_profile_mu;
struct EventPair { cudaEvent_t a, b; };
static std::vector<EventPair> g_events[PT_COUNT];
static cudaEvent_t g_open[PT_COUNT];

void profile_begin(int tag, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_profile_mu);
  if (g_events[tag].size() >= 8192) { g_open[tag] = nullptr; return; }
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) { g_open[tag] = nullptr; return; }
  cudaEventRecord(e, st);
  g_open[tag] = e;
}

void profile_end(int tag, cudaStream_t st) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_profile_mu);
  if (!g_open[tag]) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) { cudaEventDestroy(g_open[tag]); g_open[tag] = nullptr; return; }
  cudaEventRecord(e, st);
  g_events[tag].push_back({g_open[tag], e});
  g_open[tag] = nullptr;
}
}  // namespace b200ir

namespace {

inline bool valid_metric(int m) { return m >= 0 && m < B200IR_NUM_METRICS; }
inline bool valid_dtype(int d) { return d == B200IR_F32 || d == B200IR_BF16; }
inline int elem_size(int d) { return d == B200IR_F32 ? 4 : 2; }

MetricParams make_params(int metric, int flags, int D, const float* w) {
  MetricParams mp{};
  mp.metric = metric; mp.flags = flags; mp.D = D;
  mp.w[0] = 1.f; mp.w[1] = mp.w[2] = mp.w[3] = mp.w[4] = 0.f;      // geometric_metrics.py:78-82 defaults
  if (w && metric == B200IR_OPTIMIZED) for (int i = 0; i < 5; ++i) mp.w[i] = w[i];
  return mp;
}

__global__ void fill_empty_topk_kernel(float* score, int64_t* idx, int64_t n, float v) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) { score[i] = v; idx[i] = -1; }
}

template <typename T>
__global__ void row_sqnorms_kernel(const T* __restrict__ X, int64_t N, int D, float* __restrict__ out) {
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const T* x = X + row * D;
  float ss = 0.f;
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int d = d0 + lane;
    const float v = d < D ? to_f32<T>(x[d]) : 0.f;
    float p = v * v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    ss += p;
  }
  if (lane == 0) out[row] = ss;
}

bool use_tensor_path(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags) {
  if (flags & B200IR_FLAG_NO_TENSOR) return false;
  return gemm_path_supported(metric, dtype, nq, N, D, k, flags);
}

}  // namespace

extern "C" {

int b200ir_version(void) { return B200IR_VERSION; }

const char* b200ir_error_string(int status) {
  switch (status) {
    case 0: return "ok";
    case B200IR_E_ARG: return "b200ir: invalid argument (null pointer, negative size or unknown enum)";
    case B200IR_E_K: return "b200ir: k out of range (1..256)";
    case B200IR_E_WORKSPACE: return "b200ir: workspace missing or too small";
    case B200IR_E_ALIGN: return "b200ir: pointer not aligned for its element type";
    case B200IR_E_DEVICE: return "b200ir: device is not an sm_100 (B200) GPU";
    case B200IR_E_SHAPE: return "b200ir: shape not supported";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "b200ir: unknown error";
}

long long b200ir_launch_count(void) { return g_launches.load(); }

void b200ir_profile_enable(int on) { g_profile_on.store(on ? 1 : 0); }

int b200ir_profile_read(int tag, float* total_ms, int* launches) {
  if (tag < 0 || tag >= PT_COUNT || !total_ms || !launches) return B200IR_E_ARG;
  std::lock_guard<std::mutex> lk(g_profile_mu);
  float sum = 0.f;
  int n = 0;
  for (auto& ev : g_events[tag]) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(ev.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev.a, ev.b);
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
    if (e != cudaSuccess) { g_events[tag].clear(); return int(e); }
    sum += ms;
    ++n;
  }
  g_events[tag].clear();
  *total_ms = sum;
  *launches = n;
  return 0;
}

int b200ir_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int b200ir_row_sqnorms(const void* X, int dtype, int64_t N, int D, float* out, void* stream) {
  if (!valid_dtype(dtype) || N < 0 || D <= 0) return B200IR_E_ARG;
  if (N == 0) return 0;
  if (!X || !out) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(X) % elem_size(dtype)) return B200IR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = int(ceil_div64(N, 8));
  ProfileScope ps(PT_MISC, st);
  if (dtype == B200IR_F32) row_sqnorms_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(X), N, D, out);
  else row_sqnorms_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(X), N, D, out);
  return int(cudaGetLastError());
}

size_t b200ir_topk_workspace_bytes(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags) {
  if (!valid_metric(metric) || !valid_dtype(dtype) || nq <= 0 || N <= 0 || D <= 0 || k < 1 || k > B200IR_MAX_K) return 0;
  if (use_tensor_path(metric, dtype, nq, N, D, k, flags))
    return gemm_workspace_bytes(metric, dtype, nq, N, D, k, flags, (flags & B200IR_FLAG_HAVE_INDEX) != 0);
  return make_scan_plan(metric, dtype, nq, N, D, k, false).total_bytes;
}

size_t b200ir_topk_fallback_counter_offset(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags) {
  if (!valid_metric(metric) || !valid_dtype(dtype) || nq <= 0 || N <= 0 || D <= 0 || k < 1 || k > B200IR_MAX_K) return ~size_t(0);
  if (!use_tensor_path(metric, dtype, nq, N, D, k, flags) || (flags & B200IR_FLAG_NO_RERANK)) return ~size_t(0);
  return gemm_fallback_counter_offset(dtype, nq, N, D, k, flags, (flags & B200IR_FLAG_HAVE_INDEX) != 0);
}

size_t b200ir_index_bytes(int dtype, int64_t N, int D) {
  if (!valid_dtype(dtype) || N <= 0 || D <= 0) return 0;
  return gemm_index_bytes(dtype, N, D);
}

int b200ir_index_build(int dtype, const void* X, int64_t N, int D, void* index, size_t index_bytes, void* stream) {
  if (!valid_dtype(dtype) || N <= 0 || D <= 0 || !X || !index) return B200IR_E_ARG;
  const size_t need = gemm_index_bytes(dtype, N, D);
  if (need == 0) return B200IR_E_SHAPE;
  if (index_bytes < need) return B200IR_E_WORKSPACE;
  return gemm_index_build(dtype, X, N, D, static_cast<unsigned char*>(index), static_cast<cudaStream_t>(stream));
}

static int topk_impl(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                     int k, int64_t index_offset, int flags, const float* weights_host,
                     float* out_score, int64_t* out_idx, const void* index, size_t index_bytes,
                     void* workspace, size_t workspace_bytes, void* stream);

int b200ir_topk(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                int k, int64_t index_offset, int flags, const float* weights_host,
                float* out_score, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
  return topk_impl(metric, dtype, Q, nq, X, N, D, k, index_offset, flags & ~B200IR_FLAG_HAVE_INDEX, weights_host, out_score, out_idx,
                   nullptr, 0, workspace, workspace_bytes, stream);
}

int b200ir_topk_indexed(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                        int k, int64_t index_offset, int flags, const float* weights_host,
                        float* out_score, int64_t* out_idx, const void* index, size_t index_bytes,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!index || reinterpret_cast<uintptr_t>(index) % 256) return B200IR_E_ARG;
  if (!valid_dtype(dtype) || N <= 0 || D <= 0) return B200IR_E_ARG;
  const size_t need = gemm_index_bytes(dtype, N, D);
  if (need == 0 || index_bytes < need) return B200IR_E_WORKSPACE;
  return topk_impl(metric, dtype, Q, nq, X, N, D, k, index_offset, flags | B200IR_FLAG_HAVE_INDEX, weights_host, out_score, out_idx,
                   index, index_bytes, workspace, workspace_bytes, stream);
}

}  // extern "C"

static int topk_impl(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                     int k, int64_t index_offset, int flags, const float* weights_host,
                     float* out_score, int64_t* out_idx, const void* index, size_t index_bytes,
                     void* workspace, size_t workspace_bytes, void* stream) {
  (void)index_bytes;
  if (!valid_metric(metric) || !valid_dtype(dtype) || nq < 0 || N < 0 || D <= 0) return B200IR_E_ARG;
  if (k < 1 || k > B200IR_MAX_K) return B200IR_E_K;
  if (N >= (int64_t(1) << 32)) return B200IR_E_SHAPE;      // shard-local ids are 32-bit inside the keys
  if (nq == 0) return 0;
  if (!Q || !out_score || !out_idx || (N > 0 && !X)) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(Q) % elem_size(dtype) || reinterpret_cast<uintptr_t>(X) % elem_size(dtype)) return B200IR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 0) {   // empty store: every slot is padding (app_pipeline.py:147-149 returns [])
    const int64_t n = nq * k;
    ProfileScope ps(PT_MISC, st);
    fill_empty_topk_kernel<<<int(ceil_div64(n, 256)), 256, 0, st>>>(out_score, out_idx, n, metric_descending(metric) ? -INFINITY : INFINITY);
    return int(cudaGetLastError());
  }
  const size_t need = b200ir_topk_workspace_bytes(metric, dtype, nq, N, D, k, flags);
  if (!workspace || workspace_bytes < need) return B200IR_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return B200IR_E_ALIGN;
  const MetricParams mp = make_params(metric, flags, D, weights_host);
  unsigned char* ws = static_cast<unsigned char*>(workspace);

  if (use_tensor_path(metric, dtype, nq, N, D, k, flags)) {
    return run_gemm_topk(metric, dtype, Q, nq, X, N, D, k, index_offset, flags, mp, out_score, out_idx, ws,
                         static_cast<const unsigned char*>(index), st);
  }
  const ScanPlan pl = make_scan_plan(metric, dtype, nq, N, D, k, false);
  cudaError_t e = run_scan(pl, dtype, Q, nq, X, N, D, k, mp, ws, nullptr, st);
  if (e != cudaSuccess) return int(e);
  e = launch_finalize(reinterpret_cast<const uint64_t*>(ws + pl.off_partial), nq, int64_t(pl.P) * k, k, mp,
                      index_offset, out_score, out_idx, st);
  return int(e);
}

extern "C" {

int b200ir_topk_paged(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k,
                      int64_t index_offset, int flags, const float* weights_host, const uint64_t* after, uint64_t* last,
                      float* out_score, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
  if (!valid_metric(metric) || !valid_dtype(dtype) || nq < 0 || N <= 0 || D <= 0) return B200IR_E_ARG;
  if (k < 1 || k > B200IR_MAX_K) return B200IR_E_K;
  if (N >= (int64_t(1) << 32)) return B200IR_E_SHAPE;
  if (nq == 0) return 0;
  if (!Q || !X || !out_score || !out_idx) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(Q) % elem_size(dtype) || reinterpret_cast<uintptr_t>(X) % elem_size(dtype)) return B200IR_E_ALIGN;
  const ScanPlan pl = make_scan_plan(metric, dtype, nq, N, D, k, false);
  if (!workspace || workspace_bytes < pl.total_bytes) return B200IR_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return B200IR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const MetricParams mp = make_params(metric, flags, D, weights_host);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  cudaError_t e = run_scan(pl, dtype, Q, nq, X, N, D, k, mp, ws, nullptr, st, after, 0);
  if (e != cudaSuccess) return int(e);
  FinalizeOpts fo;
  fo.last_key = last;
  return int(launch_finalize(reinterpret_cast<const uint64_t*>(ws + pl.off_partial), nq, int64_t(pl.P) * k, k, mp, index_offset,
                             out_score, out_idx, st, fo));
}

int b200ir_sort_topk_rows(int descending, float* score, int64_t* idx, int64_t nq, int K, void* stream) {
  if (nq < 0 || K < 1 || K > kSortRowsMax || nq > 0x7fffffff) return B200IR_E_ARG;
  if (nq == 0) return 0;
  if (!score || !idx) return B200IR_E_ARG;
  int n = 1;
  while (n < K) n <<= 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = size_t(n) * sizeof(key128_t);
  cudaError_t e = cudaFuncSetAttribute(sort_topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSortRowsMax * sizeof(key128_t)));
  if (e != cudaSuccess) return int(e);
  ProfileScope ps(PT_MERGE, st);
  sort_topk_rows_kernel<<<unsigned(nq), 256, smem, st>>>(descending ? 1 : 0, score, idx, K, n);
  return int(cudaGetLastError());
}

static int multi_kind_mask(const int* metrics_host, int nmetrics) {
  if (!metrics_host || nmetrics < 1 || nmetrics > B200IR_NUM_METRICS) return -1;
  int mask = 0;
  for (int i = 0; i < nmetrics; ++i) {
    if (!valid_metric(metrics_host[i])) return -1;
    mask |= 1 << rank_kind_of(metrics_host[i]);
  }
  return mask;
}

size_t b200ir_topk_multi_workspace_bytes(const int* metrics_host, int nmetrics, int dtype, int64_t nq, int64_t N, int D, int k) {
  const int mask = multi_kind_mask(metrics_host, nmetrics);
  if (mask <= 0 || !valid_dtype(dtype) || nq <= 0 || N <= 0 || D <= 0 || k < 1 || k > B200IR_MAX_K) return 0;
  return make_scan_plan(B200IR_OPTIMIZED, dtype, nq, N, D, k, false, __builtin_popcount(mask)).total_bytes;
}

int b200ir_topk_multi(const int* metrics_host, int nmetrics, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                      int k, int64_t index_offset, int flags, const float* weights_host, float* out_score, int64_t* out_idx,
                      void* workspace, size_t workspace_bytes, void* stream) {
  const int mask = multi_kind_mask(metrics_host, nmetrics);
  if (mask <= 0 || !valid_dtype(dtype) || nq < 0 || N < 0 || D <= 0) return B200IR_E_ARG;
  if (k < 1 || k > B200IR_MAX_K) return B200IR_E_K;
  if (N >= (int64_t(1) << 32)) return B200IR_E_SHAPE;
  if (nq == 0) return 0;
  if (!Q || !out_score || !out_idx || (N > 0 && !X)) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(Q) % elem_size(dtype) || reinterpret_cast<uintptr_t>(X) % elem_size(dtype)) return B200IR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 0) {
    ProfileScope ps(PT_MISC, st);
    for (int y = 0; y < nmetrics; ++y) {
      const int64_t n = nq * k;
      fill_empty_topk_kernel<<<int(ceil_div64(n, 256)), 256, 0, st>>>(out_score + y * n, out_idx + y * n, n,
                                                                     metric_descending(metrics_host[y]) ? -INFINITY : INFINITY);
    }
    return int(cudaGetLastError());
  }
  const int nl = __builtin_popcount(mask);
  const ScanPlan pl = make_scan_plan(B200IR_OPTIMIZED, dtype, nq, N, D, k, false, nl);
  if (!workspace || workspace_bytes < pl.total_bytes) return B200IR_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return B200IR_E_ALIGN;
  // weights feed the optimized ranking (geometric_metrics.py:78-82 defaults when NULL)
  MetricParams mp = make_params(B200IR_OPTIMIZED, flags, D, weights_host);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  cudaError_t e = run_scan(pl, dtype, Q, nq, X, N, D, k, mp, ws, nullptr, st, nullptr, mask);
  if (e != cudaSuccess) return int(e);
  FinalizeOpts fo;
  fo.nmetrics = nmetrics;
  int slot_of_kind[RK_COUNT], next = 0;
  for (int m = 0; m < RK_COUNT; ++m) slot_of_kind[m] = (mask >> m) & 1 ? next++ : -1;
  for (int y = 0; y < nmetrics; ++y) {
    fo.slot[y] = (signed char)slot_of_kind[rank_kind_of(metrics_host[y])];
    fo.metric[y] = (signed char)metrics_host[y];
  }
  fo.list_stride = nq * int64_t(pl.P) * k;
  fo.out_stride = nq * int64_t(k);
  return int(launch_finalize(reinterpret_cast<const uint64_t*>(ws + pl.off_partial), nq, int64_t(pl.P) * k, k, mp, index_offset,
                             out_score, out_idx, st, fo));
}

int b200ir_candidate_metrics(int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, const int64_t* cand_idx, int kc,
                             float* out, void* stream) {
  if (!valid_dtype(dtype) || nq < 0 || N < 0 || D <= 0 || kc < 1) return B200IR_E_ARG;
  if (nq == 0) return 0;
  if (!Q || !cand_idx || !out || (N > 0 && !X)) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(Q) % elem_size(dtype) || reinterpret_cast<uintptr_t>(X) % elem_size(dtype)) return B200IR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope ps(PT_MISC, st);
  if (dtype == B200IR_F32) return int(launch_pair_metrics<float>(Q, nq, X, N, D, nullptr, cand_idx, nq * kc, out, st, kc));
  return int(launch_pair_metrics<__nv_bfloat16>(Q, nq, X, N, D, nullptr, cand_idx, nq * kc, out, st, kc));
}

int b200ir_rank_candidates(const float* pair_vals, const int64_t* cand_idx, int64_t nq, int kc, const float* weights_host, int k,
                           float* out_optimized, int32_t* out_pos, float* out_val, int64_t* out_row, void* stream) {
  if (nq < 0 || kc < 1 || kc > kRankMaxCand || k < 1 || k > kc || nq > 0x7fffffff) return B200IR_E_ARG;
  if (nq == 0) return 0;
  if (!pair_vals || !cand_idx || !out_pos || !out_val || !out_row) return B200IR_E_ARG;
  float w[5] = {1.f, 0.f, 0.f, 0.f, 0.f};                       // geometric_metrics.py:78-82
  if (weights_host) for (int i = 0; i < 5; ++i) w[i] = weights_host[i];
  int n = 1;
  while (n < kc) n <<= 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope ps(PT_MISC, st);
  rank_candidates_kernel<<<unsigned(nq), 128, 0, st>>>(pair_vals, cand_idx, nq, kc, n, w[0], w[1], w[2], w[3], w[4], k, out_optimized,
                                                      out_pos, out_val, out_row);
  return int(cudaGetLastError());
}

size_t b200ir_pairwise_workspace_bytes(int metric, int dtype, int64_t nq, int64_t N, int D) {
  if (!valid_metric(metric) || !valid_dtype(dtype) || nq <= 0 || N <= 0 || D <= 0) return 0;
  return make_scan_plan(metric, dtype, nq, N, D, 1, true).total_bytes;
}

int b200ir_pairwise(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                    int flags, const float* weights_host, float* out, void* workspace, size_t workspace_bytes,
                    void* stream) {
  if (!valid_metric(metric) || !valid_dtype(dtype) || nq < 0 || N < 0 || D <= 0) return B200IR_E_ARG;
  if (N >= (int64_t(1) << 32)) return B200IR_E_SHAPE;
  if (nq == 0 || N == 0) return 0;
  if (!Q || !X || !out) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(Q) % elem_size(dtype) || reinterpret_cast<uintptr_t>(X) % elem_size(dtype)) return B200IR_E_ALIGN;
  const ScanPlan pl = make_scan_plan(metric, dtype, nq, N, D, 1, true);
  if (!workspace || workspace_bytes < pl.total_bytes) return B200IR_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return B200IR_E_ALIGN;
  const MetricParams mp = make_params(metric, flags, D, weights_host);
  return int(run_scan(pl, dtype, Q, nq, X, N, D, 1, mp, static_cast<unsigned char*>(workspace), out,
                      static_cast<cudaStream_t>(stream)));
}

int b200ir_topk_merge(int descending, const float* score, const int64_t* idx, int R, int64_t nq, int k,
                      float* out_score, int64_t* out_idx, void* stream) {
  if (R < 1 || nq < 0) return B200IR_E_ARG;
  if (k < 1 || k > B200IR_MAX_K) return B200IR_E_K;
  if (nq == 0) return 0;
  if (!score || !idx || !out_score || !out_idx) return B200IR_E_ARG;
  return int(launch_merge(descending ? 1 : 0, score, idx, nq * k, nq * k, R, nq, k, out_score, out_idx, static_cast<cudaStream_t>(stream)));
}

int b200ir_topk_merge_strided(int descending, const float* score, const int64_t* idx, int64_t score_shard_stride,
                              int64_t idx_shard_stride, int R, int64_t nq, int k, float* out_score, int64_t* out_idx,
                              void* stream) {
  if (R < 1 || nq < 0 || score_shard_stride < 0 || idx_shard_stride < 0) return B200IR_E_ARG;
  if (k < 1 || k > B200IR_MAX_K) return B200IR_E_K;
  if (nq == 0) return 0;
  if (!score || !idx || !out_score || !out_idx) return B200IR_E_ARG;
  return int(launch_merge(descending ? 1 : 0, score, idx, score_shard_stride, idx_shard_stride, R, nq, k, out_score, out_idx,
                          static_cast<cudaStream_t>(stream)));
}

size_t b200ir_allpairs_eval_workspace_bytes(int64_t N, int D, int nthr) {
  if (N <= 0 || D <= 0 || nthr < 0) return 0;
  return allpairs_eval_workspace_bytes(N, D, nthr);
}

int b200ir_allpairs_eval(const float* X, const int32_t* cat, const int32_t* col, int64_t N, int D, int nbins,
                         const float* lo_host, const float* hi_host, const double* thresholds_host, int nthr,
                         uint64_t* hist, uint64_t* thr_counts, void* workspace, size_t workspace_bytes, void* stream) {
  return b200ir_allpairs_eval_part(X, cat, col, N, D, nbins, lo_host, hi_host, thresholds_host, nthr, 0, 1, hist, thr_counts, workspace,
                                   workspace_bytes, stream);
}

int b200ir_allpairs_eval_part(const float* X, const int32_t* cat, const int32_t* col, int64_t N, int D, int nbins,
                              const float* lo_host, const float* hi_host, const double* thresholds_host, int nthr, int part, int nparts,
                              uint64_t* hist, uint64_t* thr_counts, void* workspace, size_t workspace_bytes, void* stream) {
  if (nparts < 1 || part < 0 || part >= nparts) return B200IR_E_ARG;
  if (N < 0 || D <= 0 || nbins < 1 || nbins > 1024 || nthr < 0 || nthr > 1024) return B200IR_E_ARG;
  if (N >= (int64_t(1) << 31)) return B200IR_E_SHAPE;
  if (!hist || !thr_counts || !lo_host || !hi_host || (nthr > 0 && !thresholds_host)) return B200IR_E_ARG;
  for (int m = 0; m < 5; ++m) if (!(hi_host[m] > lo_host[m])) return B200IR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N < 2) {
    cudaError_t e = cudaMemsetAsync(hist, 0, size_t(5) * 4 * nbins * 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(thr_counts, 0, size_t(5) * 2 * (nthr + 1) * 8, st);
    return int(e);
  }
  if (!X || !cat || !col) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(X) % 4) return B200IR_E_ALIGN;
  const size_t need = allpairs_eval_workspace_bytes(N, D, nthr);
  if (!workspace || workspace_bytes < need) return B200IR_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return B200IR_E_ALIGN;
  return int(run_allpairs_eval(X, cat, col, N, D, nbins, lo_host, hi_host, thresholds_host, nthr,
                               reinterpret_cast<unsigned long long*>(hist), reinterpret_cast<unsigned long long*>(thr_counts),
                               static_cast<unsigned char*>(workspace), st, part, nparts));
}

int b200ir_pair_metrics(int dtype, const void* A, int64_t NA, const void* B, int64_t NB, int D,
                        const int64_t* ia, const int64_t* ib, int64_t P, float* out, void* stream) {
  if (!valid_dtype(dtype) || NA < 0 || NB < 0 || D <= 0 || P < 0) return B200IR_E_ARG;
  if (P == 0) return 0;
  if (!ia || !ib || !out || (NA > 0 && !A) || (NB > 0 && !B)) return B200IR_E_ARG;
  if (reinterpret_cast<uintptr_t>(A) % elem_size(dtype) || reinterpret_cast<uintptr_t>(B) % elem_size(dtype)) return B200IR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope ps(PT_MISC, st);
  if (dtype == B200IR_F32) return int(launch_pair_metrics<float>(A, NA, B, NB, D, ia, ib, P, out, st));
  return int(launch_pair_metrics<__nv_bfloat16>(A, NA, B, NB, D, ia, ib, P, out, st));
}

int b200ir_histogram(int colorspace, const uint8_t* img, int64_t B, int H, int W, int bins_per_channel,
                     uint32_t* out_counts, void* stream) {
  if ((colorspace != B200IR_RGB && colorspace != B200IR_HSV) || B < 0 || H <= 0 || W <= 0) return B200IR_E_ARG;
  if (bins_per_channel != 8) return B200IR_E_SHAPE;
  if (B == 0) return 0;
  if (!img || !out_counts) return B200IR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t pixels = int64_t(H) * W;
  int slices = int(ceil_div64(pixels, 65536));
  slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
  if (B * slices > 0x7fffffff) return B200IR_E_SHAPE;
  const int vector_ok = ((reinterpret_cast<uintptr_t>(img) & 15) == 0 && pixels % 16 == 0) ? 1 : 0;
  cudaError_t e;
  if (slices > 1) {
    e = cudaMemsetAsync(out_counts, 0, size_t(B) * kHistBins * sizeof(uint32_t), st);
    if (e != cudaSuccess) return int(e);
  }
  if (colorspace == B200IR_HSV) {
    e = init_hsv_tables();
    if (e != cudaSuccess) return int(e);
  }
  ProfileScope ps(PT_HIST, st);
  if (colorspace == B200IR_HSV) {
    // persistent: two CTAs per SM, each with its own lane-replicated division tables (100 KB of shared memory)
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return int(e);
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return int(e);
    e = cudaFuncSetAttribute(hsv_histogram_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(HsvSmem)));
    if (e != cudaSuccess) return int(e);
    const int64_t items = B * slices;
    const unsigned grid = unsigned(items < int64_t(2 * sms) ? items : int64_t(2 * sms));
    hsv_histogram_kernel<0><<<grid, kHsvThreads, sizeof(HsvSmem), st>>>(img, pixels, slices, vector_ok, items, out_counts);
  } else {
    histogram_kernel<<<unsigned(B * slices), kHistThreads, 0, st>>>(img, pixels, slices, vector_ok, out_counts);
  }
  return int(cudaGetLastError());
}

int b200ir_threshold_dedupe(const float* score, const int64_t* idx, int64_t nq, int kc, const int64_t* group, int64_t N,
                            double threshold, int relative, int top_k, float* out_score, int64_t* out_idx,
                            int32_t* out_count, void* stream) {
  if (nq < 0 || kc < 1 || kc > kPostMaxCand || top_k < 1 || N < 0 || nq > 0x7fffffff / 32) return B200IR_E_ARG;
  if (nq == 0) return 0;
  if (!score || !idx || !out_score || !out_idx) return B200IR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope ps(PT_MISC, st);
  threshold_dedupe_kernel<<<int(ceil_div64(nq, 4)), 128, 0, st>>>(score, idx, int(nq), kc, group, N, threshold, relative ? 1 : 0,
                                                                 top_k, out_score, out_idx, out_count);
  return int(cudaGetLastError());
}

static bool resize_args_ok(int H, int W, int rh, int rw, int top, int left, int ch, int cw) {
  return H > 0 && W > 0 && rh > 0 && rw > 0 && ch > 0 && cw > 0 && top >= 0 && left >= 0 && top + ch <= rh && left + cw <= rw &&
         int64_t(W) * 3 < (int64_t(1) << 30) && int64_t(cw) * 3 < (1 << 20);
}

size_t b200ir_resize_crop_workspace_bytes(int H, int W, int resized_h, int resized_w, int crop_top, int crop_left,
                                          int crop_h, int crop_w) {
  if (!resize_args_ok(H, W, resized_h, resized_w, crop_top, crop_left, crop_h, crop_w)) return 0;
  const ResizePlan p = make_resize_plan(H, W, resized_h, resized_w, crop_top, crop_left, crop_h, crop_w);
  return p.ok ? p.table_bytes : 0;
}

int b200ir_resize_crop(const uint8_t* img, int64_t B, int H, int W, int resized_h, int resized_w, int crop_top,
                       int crop_left, int crop_h, int crop_w, uint8_t* out, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (B < 0 || !resize_args_ok(H, W, resized_h, resized_w, crop_top, crop_left, crop_h, crop_w)) return B200IR_E_ARG;
  if (B == 0) return 0;
  if (!img || !out) return B200IR_E_ARG;
  const ResizePlan p = make_resize_plan(H, W, resized_h, resized_w, crop_top, crop_left, crop_h, crop_w);
  if (!p.ok) return B200IR_E_SHAPE;                 // down-scale factor too large for the shared-memory tile
  if (!workspace || workspace_bytes < p.table_bytes) return B200IR_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return B200IR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope ps(PT_RESIZE, st);
  return int(run_resize_crop(p, img, B, H, W, crop_h, crop_w, out, static_cast<unsigned char*>(workspace), st));
}

int b200ir_counts_to_embedding(const uint32_t* counts, int64_t B, int nb, float* raw_out, float* unit_out,
                               float* mag_out, void* stream) {
  if (B < 0 || nb <= 0) return B200IR_E_ARG;
  if (B == 0) return 0;
  if (!counts) return B200IR_E_ARG;
  ProfileScope ps(PT_MISC, static_cast<cudaStream_t>(stream));
  counts_to_embedding_kernel<<<unsigned(B), 128, 0, static_cast<cudaStream_t>(stream)>>>(counts, nb, raw_out, unit_out, mag_out);
  return int(cudaGetLastError());
}

}  // extern "C"
