/*
 * b200ir.h - C ABI of the B200-native brute-force retrieval hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json's north_star names.  The
 * reference (MeltingCrystals/Image-Retrieval-) is pure Python and has no FFI of its own
 * (SURVEY.md section 8b): the entry points below are what a ctypes binding for that
 * path binds, one per reference loop they replace (cited per function).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - Every pointer is a DEVICE pointer owned by the caller unless the name ends in
 *     `_host`.  Work is enqueued on the caller's `stream` (a cudaStream_t passed as
 *     void*) and is asynchronous; no hidden allocation: scratch is an explicit
 *     caller-provided workspace sized by the matching *_workspace_bytes() query.
 *   - Return value: 0 = ok, < 0 = argument error (B200IR_E_*), > 0 = cudaError_t.
 *     Nothing throws.  b200ir_error_string() names either kind.
 *   - Output contract of every top-k: row i is sorted best-first, ties broken by
 *     ascending global index (Python stable sort + slice: app_pipeline.py:171-172,
 *     image_search.py:199-219); slots past the number of database rows hold
 *     (+inf | -inf, -1).
 *   - There is no CPU fallback anywhere behind this ABI.
 */
#ifndef B200IR_H
#define B200IR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200IR_VERSION 100

/* metric ids (geometric_metrics.py:11-94) */
#define B200IR_L1         0   /* sum|a-b| (/D unless RAW)                  :34-39 */
#define B200IR_L2         1   /* sqrt(sum (a-b)^2) (/sqrt(D) unless RAW)    :42-47 */
#define B200IR_LINF       2   /* max|a-b|                                   :50-52 */
#define B200IR_COS_SIM    3   /* dot/(|a||b|), 0 when a norm is 0           :12-18 */
#define B200IR_COS_DIST   4   /* 1 - cos                                    :29-31 */
#define B200IR_ANGLE      5   /* arccos(clip(cos,-1,1))                     :21-26 */
#define B200IR_MAG_DIFF   6   /* | |a| - |b| |                              :55-57 */
#define B200IR_OPTIMIZED  7   /* w_angle*cos - w_l1*L1n - w_l2*L2n - w_inf*Linf - w_mag*mag  :60-94 */
#define B200IR_NUM_METRICS 8

/* element types of Q and X (same type for both) */
#define B200IR_F32  0
#define B200IR_BF16 1

/* flags */
#define B200IR_FLAG_RAW        1   /* normalized=False for L1 / L2 (geometric_metrics.py:37,45) */
#define B200IR_FLAG_ABS_SCORE  2   /* rank COS_SIM / OPTIMIZED by |score| (app_pipeline.py:167)  */
#define B200IR_FLAG_NO_TENSOR  4   /* force the CUDA-core scan even where the tcgen05 path applies */
#define B200IR_FLAG_NO_RERANK  8   /* tcgen05 path, bf16 stores: skip the exact fp32 re-rank of the candidates */
#define B200IR_FLAG_HAVE_INDEX 16  /* b200ir_topk_workspace_bytes: size the workspace for b200ir_topk_indexed */

/* colour spaces of b200ir_histogram */
#define B200IR_RGB 0
#define B200IR_HSV 1   /* OpenCV 8-bit RGB2HSV (H in [0,180)), bit-exact */

/* argument errors */
#define B200IR_E_ARG        (-1)   /* null pointer / negative size / unknown enum */
#define B200IR_E_K          (-2)   /* k out of range (1..B200IR_MAX_K) */
#define B200IR_E_WORKSPACE  (-3)   /* workspace missing or too small */
#define B200IR_E_ALIGN      (-4)   /* pointer not aligned for its element type */
#define B200IR_E_DEVICE     (-5)   /* not an sm_100 device / kernel image missing */
#define B200IR_E_SHAPE      (-6)   /* shape not supported by the requested path */
#define B200IR_MAX_K 256

int b200ir_version(void);
const char* b200ir_error_string(int status);

/* 1 if the current device can run the library (compute capability 10.x), else 0. */
int b200ir_device_ok(void);

/*
 * Row squared norms, out[i] = sum_j X[i][j]^2 in fp32.
 * Replaces the np.linalg.norm calls repeated per pair at geometric_metrics.py:14-15,57
 * and app_pipeline.py:161-163.
 */
int b200ir_row_sqnorms(const void* X, int dtype, int64_t N, int D, float* out, void* stream);

/*
 * Fused distance + top-k scan of nq queries against N database rows.
 * Replaces the per-pair scan loops + list.sort + slice of
 *   app_pipeline.py:156-172 (search_images), :296-328 (search_with_multiple_metrics),
 *   image_search.py:98-115 and :173-219 (candidate scoring + per-metric sorts).
 *
 *   Q [nq, D], X [N, D] row-major, contiguous, element type `dtype`.
 *   index_offset: added to every returned index (base of this shard of a row-sharded DB).
 *   weights_host: 5 floats {w_angle, w_l1, w_l2, w_inf, w_mag} (HOST memory) for
 *                 B200IR_OPTIMIZED, ignored (may be NULL) otherwise.
 *   out_score [nq, k] fp32: the metric's reference-normalised value of each winner.
 *   out_idx   [nq, k] int64: global row ids.
 * Ranking: distances ascending, COS_SIM / OPTIMIZED descending; COS_DIST and ANGLE are
 * ranked by descending cosine (monotone), L2 by the squared sum.
 */
size_t b200ir_topk_workspace_bytes(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags);
int b200ir_topk(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                int k, int64_t index_offset, int flags, const float* weights_host,
                float* out_score, int64_t* out_idx,
                void* workspace, size_t workspace_bytes, void* stream);

/*
 * Exactness of the tensor-core path.  L2 / cosine-family searches of bf16 and fp32 stores (D % 8 == 0, D <= 512,
 * nq >= 32, N >= 1024, k <= 240) select kp > k candidates per query on tcgen05 (fp32 stores: three-term bf16 split),
 * re-rank them with exact fp32 arithmetic on the original rows and CERTIFY the result: if a dropped row could still
 * reach the k-th exact score within the proven error of the tensor-core pass, the query is re-done by the exact
 * CUDA-core scan in the same call.  Results therefore equal the scan's; the int32 at this byte offset of the
 * workspace holds the number of queries the last call re-did (device memory; (size_t)-1: the shape does not take the
 * tensor-core path or B200IR_FLAG_NO_RERANK asked for the uncertified candidate scores).
 */
size_t b200ir_topk_fallback_counter_offset(int metric, int dtype, int64_t nq, int64_t N, int D, int k, int flags);

/*
 * Prepared per-store state for repeated searches of a static store (the store plays the part of
 * app_pipeline.py:18 `self.embeddings` / the Milvus collection of ImageEmbeddingSystem.py:41-61, which the
 * reference also builds once and searches many times): row norms (the reference recomputes |e| for every pair,
 * app_pipeline.py:161-163), the largest norm, and for fp32 stores the two bf16 planes of the error-compensated
 * split the tensor-core pass multiplies.  b200ir_index_bytes() is 0 when the shape has no tensor-core path
 * (D % 8 != 0 or D > 512): use plain b200ir_topk then.  The index is only valid while X is unchanged.
 * b200ir_topk_indexed == b200ir_topk with that state supplied instead of being rebuilt inside the workspace on
 * every call; results are identical.  Metrics without a tensor-core path ignore the index.
 */
size_t b200ir_index_bytes(int dtype, int64_t N, int D);
int b200ir_index_build(int dtype, const void* X, int64_t N, int D, void* index, size_t index_bytes, void* stream);
int b200ir_topk_indexed(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                        int k, int64_t index_offset, int flags, const float* weights_host,
                        float* out_score, int64_t* out_idx, const void* index, size_t index_bytes,
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * Result pages beyond B200IR_MAX_K.  The reference returns results[:top_k] for ANY top_k (app_pipeline.py:172) and
 * fetches 3 * top_k / 5 * top_k candidates (image_search.py:92, :169).  b200ir_topk_paged is b200ir_topk's exact
 * CUDA-core scan with a cursor: `after` [nq] (NULL for the first page) holds each query's last rank key of the
 * previous page and only rows that sort strictly after it are candidates; `last` [nq] receives the cursor of this
 * page (all-ones when the store is exhausted).  Workspace: b200ir_topk_workspace_bytes(..., flags | NO_TENSOR).
 * Pages are concatenated by the caller; b200ir_sort_topk_rows then orders each row of score / idx [nq][K] (K <= 4096)
 * by (score, index) in place - distinct rank values can round to one fp32 score across a page boundary, and the
 * reference's stable sort orders equal scores by index.
 */
int b200ir_topk_paged(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, int k,
                      int64_t index_offset, int flags, const float* weights_host, const uint64_t* after, uint64_t* last,
                      float* out_score, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);
int b200ir_sort_topk_rows(int descending, float* score, int64_t* idx, int64_t nq, int K, void* stream);

/*
 * Several metrics from ONE pass over the store: replaces the three scans + three sorts of
 * EnhancedImageSearchApp.search_with_multiple_metrics (app_pipeline.py:296-328).  The scan keeps one candidate list per
 * requested ranking (dot, sum|d|, sum d^2, max|d| and the norms are accumulated together; cosine similarity /
 * distance / angle share the descending-cosine list) and writes out_score / out_idx [nmetrics][nq][k], plane y holding
 * metric metrics_host[y] with b200ir_topk's contract.  weights_host feeds B200IR_OPTIMIZED (defaults when NULL).
 */
size_t b200ir_topk_multi_workspace_bytes(const int* metrics_host, int nmetrics, int dtype, int64_t nq, int64_t N, int D, int k);
int b200ir_topk_multi(const int* metrics_host, int nmetrics, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                      int k, int64_t index_offset, int flags, const float* weights_host, float* out_score, int64_t* out_idx,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Re-ranking of candidate lists (image_search.py:98-115 and :173-219).  pair_vals [7][nq * kc] are the get_all_metrics
 * values of the pairs (query q, candidate c) as b200ir_pair_metrics writes them; cand_idx [nq][kc] the candidates'
 * rows, -1 = padding (kc <= 1024).  out_optimized [nq][kc] (may be NULL) receives w_angle*cos - w_l1*L1 - w_l2*L2 -
 * w_inf*Linf - w_mag*mag of every candidate (geometric_metrics.py:85-92); for each ordering y = 0 cosine desc, 1 l1,
 * 2 l2, 3 linf, 4 magnitude (ascending), 5 optimized desc, the first k entries of the stably sorted candidate list
 * (ties keep candidate order, like list.sort) go to out_pos (position in the list), out_val (the sort value) and
 * out_row (database row), each [6][nq][k], padding -1 / NaN / -1.
 */
/* get_all_metrics of every (query q, candidate cand_idx[q][c]) pair: b200ir_pair_metrics with the query index implied;
 * out [7][nq * kc], NaN for padding candidates (cand_idx < 0). */
int b200ir_candidate_metrics(int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D, const int64_t* cand_idx, int kc,
                             float* out, void* stream);
int b200ir_rank_candidates(const float* pair_vals, const int64_t* cand_idx, int64_t nq, int kc, const float* weights_host, int k,
                           float* out_optimized, int32_t* out_pos, float* out_val, int64_t* out_row, void* stream);

/*
 * Full (nq, N) metric matrix, same arithmetic as b200ir_topk's CUDA-core scan.
 * Replaces the pair loops of mi_analysis.py:277-292 / get_all_metrics
 * (geometric_metrics.py:114-129) for evaluation-sized inputs.
 */
size_t b200ir_pairwise_workspace_bytes(int metric, int dtype, int64_t nq, int64_t N, int D);
int b200ir_pairwise(int metric, int dtype, const void* Q, int64_t nq, const void* X, int64_t N, int D,
                    int flags, const float* weights_host, float* out,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * Merge R per-shard top-k lists (after the all-gather of a row-sharded search):
 * score [R, nq, k], idx [R, nq, k] -> out_score [nq, k], out_idx [nq, k], ordered by
 * (score, global index); entries with idx < 0 are padding.  SURVEY.md section 8e.
 */
int b200ir_topk_merge(int descending, const float* score, const int64_t* idx, int R, int64_t nq, int k,
                      float* out_score, int64_t* out_idx, void* stream);

/* Same merge on lists that sit `*_shard_stride` ELEMENTS apart per shard (the receive buffer of one packed all-gather). */
int b200ir_topk_merge_strided(int descending, const float* score, const int64_t* idx, int64_t score_shard_stride,
                              int64_t idx_shard_stride, int R, int64_t nq, int k, float* out_score, int64_t* out_idx,
                              void* stream);

/*
 * All-pairs evaluation: the reference's analysis workload (mi_analysis.py:256-297 distances per relationship type,
 * :704-713 per-metric densities, :774-796 precision / recall threshold counts) over EVERY unordered pair i < j of the
 * N rows of X [N, D] fp32, without materialising the N x N matrices.
 *   cat[N], col[N]: object category / colour of each row; relationship type of a pair =
 *     0 same_object_same_color, 1 same_object_diff_color, 2 diff_object_same_color, 3 diff_object_diff_color
 *     (mi_analysis.py:176-181).
 *   metrics, in this order (mi_analysis.py:183-189): cosine_distance, l1_distance, l2_distance (both normalised),
 *     linf_distance, magnitude_difference - all from one pass, same arithmetic as b200ir_pairwise.
 *   hist [5][4][nbins] uint64: counts of metric m per relationship type in nbins equal bins of
 *     [lo_host[m], hi_host[m]) (values outside go to the end bins); nbins <= 1024.
 *   thr_counts [5][2][nthr + 1] uint64: for the two same-object labels (0 = same colour, 1 = different colour) the
 *     number of pairs whose FIRST index t with d <= thresholds_host[t] is t (slot nthr: none); thresholds ascending
 *     doubles (np.linspace(0, 1, 100) in the reference).  Prefix sums give tp / fp / fn of :783-796.
 */
size_t b200ir_allpairs_eval_workspace_bytes(int64_t N, int D, int nthr);
int b200ir_allpairs_eval(const float* X, const int32_t* cat, const int32_t* col, int64_t N, int D, int nbins,
                         const float* lo_host, const float* hi_host, const double* thresholds_host, int nthr,
                         uint64_t* hist, uint64_t* thr_counts, void* workspace, size_t workspace_bytes, void* stream);

/* One part of the same evaluation for a store replicated on `nparts` GPUs: part `part` counts the pairs (i, j > i) of
 * its cyclic share of the rows i (groups of 8 rows dealt round-robin: the pair grid is triangular).  The parts' hist /
 * thr_counts add up to b200ir_allpairs_eval's (one all-reduce of the integer count tensors). */
int b200ir_allpairs_eval_part(const float* X, const int32_t* cat, const int32_t* col, int64_t N, int D, int nbins,
                              const float* lo_host, const float* hi_host, const double* thresholds_host, int nthr, int part, int nparts,
                              uint64_t* hist, uint64_t* thr_counts, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Explicit pair lists: replaces the per-pair get_all_metrics calls of ColorMIAnalyzer.calculate_distances
 * (mi_analysis.py:256-297 over pairs.json; geometric_metrics.py:114-129).  For p in [0, P): the pair
 * (A[ia[p]], B[ib[p]]) -> out[m][p], out [7][P] fp32, m in get_all_metrics key order:
 *   0 cosine_similarity, 1 cosine_distance, 2 angular_distance, 3 l1_distance (/D), 4 l2_distance (/sqrt(D)),
 *   5 linf_distance, 6 magnitude_difference.  A and B may be the same matrix.  A pair naming a row outside
 *   [0, NA) x [0, NB) is written as NaN (the reference logs a warning and skips it, :278-280).
 */
#define B200IR_PAIR_OUTPUTS 7
int b200ir_pair_metrics(int dtype, const void* A, int64_t NA, const void* B, int64_t NB, int D,
                        const int64_t* ia, const int64_t* ib, int64_t P, float* out, void* stream);

/*
 * Post-filter of best-first candidate lists, image_search.py:115-140: score [nq, kc] descending with idx [nq, kc]
 * (padding idx -1 at the end, as b200ir_topk writes it).  Keeps score >= threshold (relative != 0: threshold becomes
 * min + threshold * (max - min) over the query's list, the rule for the optimized score :118-123), drops an entry when
 * an earlier one has the same path (group[row], one id per distinct path; NULL: every row is its own path, :128-137)
 * and writes the first top_k survivors to out_score / out_idx [nq, top_k] (padding -inf / -1) and their number to
 * out_count [nq] (may be NULL).  kc <= 1024.
 */
int b200ir_threshold_dedupe(const float* score, const int64_t* idx, int64_t nq, int kc, const int64_t* group, int64_t N,
                            double threshold, int relative, int top_k, float* out_score, int64_t* out_idx,
                            int32_t* out_count, void* stream);

/*
 * Image front-end of the embedding producer (ImageEmbeddingSystem.py:82-83, app_pipeline.py:127-131: PIL image ->
 * CLIPProcessor = resize shorter edge to 224 with PIL BICUBIC, centre crop 224 x 224).  img [B, H, W, 3] uint8 is
 * resized to resized_h x resized_w with Pillow's 8-bit bicubic arithmetic (bit-exact: horizontal pass first, 22-bit
 * fixed-point taps) and the window [crop_top, crop_top + crop_h) x [crop_left, crop_left + crop_w) of the result is
 * written to out [B, crop_h, crop_w, 3]; only that window is computed.  The workspace holds the tap tables.
 * B200IR_E_SHAPE when the vertical down-scale factor is too large for the on-chip tile (> ~80x).
 */
size_t b200ir_resize_crop_workspace_bytes(int H, int W, int resized_h, int resized_w, int crop_top, int crop_left,
                                          int crop_h, int crop_w);
int b200ir_resize_crop(const uint8_t* img, int64_t B, int H, int W, int resized_h, int resized_w, int crop_top,
                       int crop_left, int crop_h, int crop_w, uint8_t* out, void* workspace, size_t workspace_bytes,
                       void* stream);

/*
 * 512-bin joint colour histogram of uint8 images, img [B, H, W, 3] interleaved RGB,
 * out_counts [B, bins^3] uint32, bin = (c0bin*bins + c1bin)*bins + c2bin with
 * c*bin = c>>5 (H: h*8/180).  bins_per_channel must be 8.  The embedding producer
 * named by north_star (no reference code exists: SURVEY.md section 0, 8c).
 */
int b200ir_histogram(int colorspace, const uint8_t* img, int64_t B, int H, int W, int bins_per_channel,
                     uint32_t* out_counts, void* stream);

/*
 * counts [B, nb] uint32 -> unit-norm fp32 vectors [B, nb] and magnitudes [B]
 * (ImageEmbeddingSystem.py:88-94: embedding / |embedding|, magnitude).  unit_out or
 * mag_out may be NULL; raw_out (may be NULL) receives the un-normalised fp32 counts.
 */
int b200ir_counts_to_embedding(const uint32_t* counts, int64_t B, int nb,
                               float* raw_out, float* unit_out, float* mag_out, void* stream);

/*
 * Measurement hooks (bench.py).  b200ir_launch_count: kernels launched by this library since load.
 * b200ir_profile_enable(1) brackets every kernel launch with CUDA events on the launching stream;
 * b200ir_profile_read(tag, &ms, &n) synchronises on and drains the events of one kernel class
 * (0 prep, 1 scan, 2 tcgen05 gemm+topk, 3 finalize, 4 rerank, 5 merge, 6 histogram, 7 misc, 8 resize).
 */
long long b200ir_launch_count(void);
void b200ir_profile_enable(int on);
int b200ir_profile_read(int tag, float* total_ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* B200IR_H */
