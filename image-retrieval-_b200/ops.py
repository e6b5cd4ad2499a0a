"""Batched operators over the C ABI (include/b200ir.h): the new, vectorised face of the hot path.

PyTorch is plumbing here (device memory, streams); every number is produced by libb200ir.so.
Tensors are passed as raw device pointers on torch's current CUDA stream.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (ANGLE, BF16, COS_DIST, COS_SIM, F32, FLAG_ABS_SCORE, FLAG_HAVE_INDEX, FLAG_NO_RERANK, FLAG_NO_TENSOR,
                   FLAG_RAW, HSV, L1, L2, LINF, MAG_DIFF, MAX_CANDIDATES, MAX_K, MAX_K_PAGED, OPTIMIZED, RGB, B200IRError)

METRIC_IDS = {
    "l1": L1, "l1_distance": L1,
    "l2": L2, "l2_distance": L2,
    "linf": LINF, "linf_distance": LINF,
    "cosine_similarity": COS_SIM, "cosine": COS_SIM,
    "cosine_distance": COS_DIST,
    "angular_distance": ANGLE, "angle": ANGLE,
    "magnitude_difference": MAG_DIFF,
    "optimized_similarity": OPTIMIZED,
}
DESCENDING = {COS_SIM, OPTIMIZED}
WEIGHT_KEYS = ("w_angle", "w_l1", "w_l2", "w_inf", "w_mag")

_workspaces = {}


def metric_id(metric):
    if isinstance(metric, int):
        return metric
    try:
        return METRIC_IDS[metric]
    except KeyError:
        raise ValueError(f"unknown metric {metric!r}; one of {sorted(METRIC_IDS)}") from None


def device():
    if not torch.cuda.is_available():
        raise B200IRError("no CUDA device: this package has no CPU fallback (B200 / sm_100a required)")
    lib = _lib.load()
    if not lib.b200ir_device_ok():
        raise B200IRError("current CUDA device is not compute capability 10.x (B200): libb200ir.so only carries sm_100a code")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def as_device_matrix(a, dtype=None):
    """numpy / torch (host or device) -> contiguous 2-D CUDA tensor of fp32 or bf16."""
    dev = device()
    if isinstance(a, np.ndarray):
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    elif isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.as_tensor(np.asarray(a, dtype=np.float32))
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError(f"expected a vector or a matrix, got shape {tuple(t.shape)}")
    want = dtype if dtype is not None else (t.dtype if t.dtype in (torch.float32, torch.bfloat16) else torch.float32)
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    if t.dtype != want:
        t = t.to(want)
    return t.contiguous()


def _dtype_id(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise ValueError(f"unsupported element type {t.dtype}")


def _workspace(nbytes, dev):
    """Scratch buffer of the C ABI calls, cached per (device, stream).  It grows on demand and is given back when a call
    needs less than a quarter of a large (> 256 MB) buffer, so one big search does not pin its scratch for ever;
    release_workspaces() drops everything (e.g. between workloads)."""
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes or (ws.numel() > (256 << 20) and nbytes * 4 < ws.numel()):
        _workspaces.pop(key, None)
        ws = None
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
        if len(_workspaces) >= 16:                 # streams come and go: do not accumulate one buffer per dead stream
            _workspaces.clear()
        _workspaces[key] = ws
    return ws


def release_workspaces():
    """Free every cached scratch buffer (they are re-created on the next call)."""
    global _last_topk
    _workspaces.clear()
    _last_topk = None


def _weights(params):
    if params is None:
        return None
    w = (ctypes.c_float * 5)()
    defaults = (1.0, 0.0, 0.0, 0.0, 0.0)       # geometric_metrics.py:78-82
    for i, key in enumerate(WEIGHT_KEYS):
        w[i] = float(params.get(key, defaults[i]))
    return w


def _flags(normalized, abs_score, flags):
    f = int(flags)
    if not normalized:
        f |= FLAG_RAW
    if abs_score:
        f |= FLAG_ABS_SCORE
    return f


def row_sqnorms(X):
    X = as_device_matrix(X)
    out = torch.empty(X.shape[0], dtype=torch.float32, device=X.device)
    lib = _lib.load()
    _lib.check(lib.b200ir_row_sqnorms(_ptr(X), _dtype_id(X), X.shape[0], X.shape[1], _ptr(out), _stream()), "row_sqnorms")
    return out


class PreparedIndex:
    """A static (N, D) store plus the per-store state the tensor-core searches reuse (b200ir_index_build: row norms,
    max norm, bf16 hi / lo planes of an fp32 store).  Pass it to topk() in place of the matrix.  Rebuild it (or call
    refresh()) after changing rows in place; the state is ignored by metrics / shapes without a tensor-core path."""

    def __init__(self, X):
        self.X = as_device_matrix(X)
        self.state = None
        self.refresh()

    def refresh(self):
        lib = _lib.load()
        N, D = self.X.shape
        need = lib.b200ir_index_bytes(_dtype_id(self.X), N, D) if N > 0 else 0
        if need and self.X.data_ptr() % 16 == 0:
            if self.state is None or self.state.numel() < need:
                self.state = torch.empty(int(need), dtype=torch.uint8, device=self.X.device)
            _lib.check(lib.b200ir_index_build(_dtype_id(self.X), _ptr(self.X), N, D, _ptr(self.state), self.state.numel(), _stream()),
                       "index_build")
        else:
            self.state = None
        return self

    @property
    def shape(self):
        return self.X.shape

    @property
    def dtype(self):
        return self.X.dtype

    @property
    def device(self):
        return self.X.device


def prepare_index(X):
    return X if isinstance(X, PreparedIndex) else PreparedIndex(X)


def topk(Q, X, metric, k, *, index_offset=0, normalized=True, abs_score=False, params=None, flags=0, out=None):
    """Fused distance + top-k.  Returns (scores (nq,k) fp32, indices (nq,k) int64) on the device,
    best first, ties by ascending index; slots past N hold (+-inf, -1).  X: matrix or PreparedIndex."""
    m = metric_id(metric)
    state = None
    if isinstance(X, PreparedIndex):
        X, state = X.X, X.state
    else:
        X = as_device_matrix(X)
    Q = as_device_matrix(Q, dtype=X.dtype)
    if Q.shape[1] != X.shape[1]:
        raise ValueError(f"dimension mismatch: queries {Q.shape[1]} vs database {X.shape[1]}")
    if not 1 <= k <= MAX_K_PAGED:
        raise ValueError(f"k must be in 1..{MAX_K_PAGED} (one page holds {MAX_K}; longer lists are served page by page)")
    lib = _lib.load()
    nq, D = Q.shape
    N = X.shape[0]
    f = _flags(normalized, abs_score, flags)
    if k > MAX_K:
        if out is not None:
            raise ValueError("out= is not supported for paged result lists (k > %d)" % MAX_K)
        return _topk_paged(lib, m, Q, X, k, int(index_offset), f, params)
    if out is None:
        scores = torch.empty((nq, k), dtype=torch.float32, device=X.device)
        idx = torch.empty((nq, k), dtype=torch.int64, device=X.device)
    else:
        scores, idx = out
        if scores.shape != (nq, k) or idx.shape != (nq, k) or scores.dtype != torch.float32 or idx.dtype != torch.int64 \
                or not scores.is_contiguous() or not idx.is_contiguous():
            raise ValueError("out=(scores, idx) must be contiguous (nq, k) fp32 / int64 tensors")
    if state is not None and nq > 0:
        need = lib.b200ir_topk_workspace_bytes(m, _dtype_id(X), nq, N, D, k, f | FLAG_HAVE_INDEX)
        ws = _workspace(need, X.device)
        st = lib.b200ir_topk_indexed(m, _dtype_id(X), _ptr(Q), nq, _ptr(X), N, D, k, int(index_offset), f, _weights(params),
                                     _ptr(scores), _ptr(idx), _ptr(state), state.numel(), _ptr(ws), ws.numel(), _stream())
    else:
        need = lib.b200ir_topk_workspace_bytes(m, _dtype_id(X), nq, N, D, k, f)
        ws = _workspace(need, X.device)
        st = lib.b200ir_topk(m, _dtype_id(X), _ptr(Q), nq, _ptr(X), N, D, k, int(index_offset), f, _weights(params),
                             _ptr(scores), _ptr(idx), _ptr(ws), ws.numel(), _stream())
    _lib.check(st, "topk")
    global _last_topk
    _last_topk = (ws, m, _dtype_id(X), nq, N, D, k, f | (FLAG_HAVE_INDEX if state is not None and nq > 0 else 0))
    return scores, idx


_last_topk = None


def _topk_paged(lib, m, Q, X, k, index_offset, f, params):
    """k > MAX_K: exact result pages of MAX_K through the CUDA-core scan with a per-query cursor (every page continues
    strictly after the previous page's last rank key), then one (score, index) ordering of the concatenated rows."""
    nq, D = Q.shape
    N = X.shape[0]
    scores = torch.empty((nq, k), dtype=torch.float32, device=X.device)
    idx = torch.empty((nq, k), dtype=torch.int64, device=X.device)
    if nq == 0:
        return scores, idx
    if N == 0:
        scores.fill_(float("-inf") if m in DESCENDING else float("inf"))
        idx.fill_(-1)
        return scores, idx
    cursors = [torch.empty(nq, dtype=torch.int64, device=X.device) for _ in range(2)]
    need = lib.b200ir_topk_workspace_bytes(m, _dtype_id(X), nq, N, D, MAX_K, f | FLAG_NO_TENSOR)
    ws = _workspace(need, X.device)
    done, page = 0, 0
    while done < k:
        kk = min(MAX_K, k - done)
        ps = torch.empty((nq, kk), dtype=torch.float32, device=X.device)
        pi = torch.empty((nq, kk), dtype=torch.int64, device=X.device)
        after = _ptr(cursors[(page + 1) % 2]) if page > 0 else None
        st = lib.b200ir_topk_paged(m, _dtype_id(X), _ptr(Q), nq, _ptr(X), N, D, kk, index_offset, f, _weights(params), after,
                                   _ptr(cursors[page % 2]), _ptr(ps), _ptr(pi), _ptr(ws), ws.numel(), _stream())
        _lib.check(st, "topk_paged")
        scores[:, done:done + kk] = ps
        idx[:, done:done + kk] = pi
        done += kk
        page += 1
    _lib.check(lib.b200ir_sort_topk_rows(1 if m in DESCENDING else 0, _ptr(scores), _ptr(idx), nq, k, _stream()), "sort_topk_rows")
    return scores, idx


def topk_multi(Q, X, metrics, k, *, index_offset=0, normalized=True, abs_score=False, params=None):
    """Top-k under SEVERAL metrics from ONE pass over the store (b200ir_topk_multi; replaces the three scans of
    app_pipeline.py:296-328).  Returns (scores (M, nq, k), indices (M, nq, k)), plane y = metrics[y]."""
    ids = [metric_id(mm) for mm in metrics]
    if isinstance(X, PreparedIndex):
        X = X.X
    else:
        X = as_device_matrix(X)
    Q = as_device_matrix(Q, dtype=X.dtype)
    if Q.shape[1] != X.shape[1]:
        raise ValueError(f"dimension mismatch: queries {Q.shape[1]} vs database {X.shape[1]}")
    if not 1 <= k <= MAX_K:
        raise ValueError(f"topk_multi: k must be in 1..{MAX_K}")
    lib = _lib.load()
    nq, D = Q.shape
    N = X.shape[0]
    M = len(ids)
    arr = (ctypes.c_int * M)(*ids)
    scores = torch.empty((M, nq, k), dtype=torch.float32, device=X.device)
    idx = torch.empty((M, nq, k), dtype=torch.int64, device=X.device)
    if nq == 0:
        return scores, idx
    need = lib.b200ir_topk_multi_workspace_bytes(arr, M, _dtype_id(X), nq, max(N, 1), D, k)
    ws = _workspace(need, X.device)
    st = lib.b200ir_topk_multi(arr, M, _dtype_id(X), _ptr(Q), nq, _ptr(X), N, D, k, int(index_offset), _flags(normalized, abs_score, 0),
                               _weights(params), _ptr(scores), _ptr(idx), _ptr(ws), ws.numel(), _stream())
    _lib.check(st, "topk_multi")
    return scores, idx


RANK_ORDERINGS = ("cosine_similarity", "l1_distance", "l2_distance", "linf_distance", "magnitude_difference", "optimized_similarity")


def rank_candidates(Q, X, cand_idx, k, params=None):
    """Re-ranking of per-query candidate lists (image_search.py:98-115, :173-219): the seven get_all_metrics values of
    every (query, candidate) pair from one launch, the weighted optimized score, and the six stable per-metric
    orderings from a second one.  cand_idx (nq, kc) int64 rows, -1 = padding.  Returns a dict of device tensors:
    metrics (7, nq, kc) in PAIR_METRICS order, optimized (nq, kc), pos / val / row (6, nq, k) in RANK_ORDERINGS order."""
    X = X.X if isinstance(X, PreparedIndex) else as_device_matrix(X)
    Q = as_device_matrix(Q, dtype=X.dtype)
    cand_idx = cand_idx.contiguous()
    nq, kc = cand_idx.shape
    if Q.shape[0] != nq:
        raise ValueError("rank_candidates: one candidate list per query")
    if not 1 <= kc <= MAX_CANDIDATES:
        raise ValueError(f"rank_candidates: candidate lists hold 1..{MAX_CANDIDATES} rows")
    k = max(1, min(int(k), kc))
    lib = _lib.load()
    vals = torch.empty((len(PAIR_METRICS), nq, kc), dtype=torch.float32, device=X.device)
    _lib.check(lib.b200ir_candidate_metrics(_dtype_id(X), _ptr(Q), nq, _ptr(X), X.shape[0], X.shape[1], _ptr(cand_idx), kc, _ptr(vals),
                                            _stream()), "candidate_metrics")
    opt = torch.empty((nq, kc), dtype=torch.float32, device=X.device)
    pos = torch.empty((6, nq, k), dtype=torch.int32, device=X.device)
    val = torch.empty((6, nq, k), dtype=torch.float32, device=X.device)
    row = torch.empty((6, nq, k), dtype=torch.int64, device=X.device)
    _lib.check(lib.b200ir_rank_candidates(_ptr(vals), _ptr(cand_idx), nq, kc, _weights(params), k, _ptr(opt), _ptr(pos), _ptr(val),
                                          _ptr(row), _stream()), "rank_candidates")
    return {"metrics": vals, "optimized": opt, "pos": pos, "val": val, "row": row}


def last_fallback_count():
    """Queries of the most recent topk() call that failed the tensor path's exactness certificate and were re-done by
    the exact scan (None when that call did not take the tensor path).  Synchronises; for tests and bench reports."""
    if _last_topk is None:
        return None
    ws, m, dt, nq, N, D, k, f = _last_topk
    off = _lib.load().b200ir_topk_fallback_counter_offset(m, dt, nq, N, D, k, f)
    if off == ctypes.c_size_t(-1).value or nq == 0 or N == 0:
        return None
    return int(ws[off:off + 4].view(torch.int32).item())


def pairwise(Q, X, metric, *, normalized=True, abs_score=False, params=None, flags=0):
    """(nq, N) fp32 matrix of `metric` (evaluation-sized inputs)."""
    m = metric_id(metric)
    X = as_device_matrix(X)
    Q = as_device_matrix(Q, dtype=X.dtype)
    if Q.shape[1] != X.shape[1]:
        raise ValueError(f"dimension mismatch: queries {Q.shape[1]} vs database {X.shape[1]}")
    lib = _lib.load()
    nq, D = Q.shape
    N = X.shape[0]
    out = torch.empty((nq, N), dtype=torch.float32, device=X.device)
    if nq == 0 or N == 0:
        return out
    need = lib.b200ir_pairwise_workspace_bytes(m, _dtype_id(X), nq, N, D)
    ws = _workspace(need, X.device)
    st = lib.b200ir_pairwise(m, _dtype_id(X), _ptr(Q), nq, _ptr(X), N, D, _flags(normalized, abs_score, flags),
                             _weights(params), _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(st, "pairwise")
    return out


def topk_merge(scores, idx, descending):
    """Merge per-shard lists: scores/idx (R, nq, k) -> (nq, k) ordered by (score, global index)."""
    device()
    scores = scores.contiguous()
    idx = idx.contiguous()
    R, nq, k = scores.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    lib = _lib.load()
    _lib.check(lib.b200ir_topk_merge(1 if descending else 0, _ptr(scores), _ptr(idx), R, nq, k, _ptr(out_s), _ptr(out_i),
                                     _stream()), "topk_merge")
    return out_s, out_i


def topk_merge_packed(gathered, R, nq, k, score_bytes, descending):
    """Merge straight out of the receive buffer of ONE packed all-gather: `gathered` is R records of
    [nq*k fp32 scores | pad | nq*k int64 ids], `score_bytes` = offset of the ids inside a record."""
    device()
    rec = gathered.numel() // R
    g = gathered.view(R, rec)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=gathered.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=gathered.device)
    lib = _lib.load()
    st = lib.b200ir_topk_merge_strided(1 if descending else 0, ctypes.c_void_p(g.data_ptr()),
                                       ctypes.c_void_p(g.data_ptr() + score_bytes), rec // 4, rec // 8, R, nq, k,
                                       _ptr(out_s), _ptr(out_i), _stream())
    _lib.check(st, "topk_merge_strided")
    return out_s, out_i


EVAL_METRICS = ("cosine_distance", "l1_distance", "l2_distance", "linf_distance", "magnitude_difference")
RELATIONSHIP_TYPES = ("same_object_same_color", "same_object_diff_color", "diff_object_same_color", "diff_object_diff_color")


def allpairs_eval(X, category, color, ranges, nbins=1024, thresholds=None, part=0, nparts=1):
    """All-pairs evaluation counts (mi_analysis.py:256-297, :704-713, :774-796) over every pair i < j.
    Returns (hist (5, 4, nbins) int64, thr_counts (5, 2, nthr + 1) int64) device tensors; see include/b200ir.h.
    part / nparts: only this part's cyclic share of the rows i (the parts' counts add up; sharded.allpairs_eval)."""
    X = as_device_matrix(X, dtype=torch.float32)
    N, D = X.shape
    cat = torch.as_tensor(category, dtype=torch.int32).to(X.device).contiguous()
    col = torch.as_tensor(color, dtype=torch.int32).to(X.device).contiguous()
    if cat.numel() != N or col.numel() != N:
        raise ValueError("category / color must have one entry per row")
    thresholds = np.linspace(0, 1, 100) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    nthr = len(thresholds)
    lo = (ctypes.c_float * 5)(*[float(ranges[m][0]) for m in EVAL_METRICS])
    hi = (ctypes.c_float * 5)(*[float(ranges[m][1]) for m in EVAL_METRICS])
    thr = (ctypes.c_double * max(nthr, 1))(*[float(t) for t in thresholds])
    hist = torch.empty((5, 4, nbins), dtype=torch.int64, device=X.device)
    thr_counts = torch.empty((5, 2, nthr + 1), dtype=torch.int64, device=X.device)
    lib = _lib.load()
    need = lib.b200ir_allpairs_eval_workspace_bytes(N, D, nthr)
    ws = _workspace(need, X.device)
    st = lib.b200ir_allpairs_eval_part(_ptr(X), _ptr(cat), _ptr(col), N, D, nbins, lo, hi, thr, nthr, int(part), int(nparts), _ptr(hist),
                                       _ptr(thr_counts), _ptr(ws), ws.numel(), _stream())
    _lib.check(st, "allpairs_eval")
    torch.cuda.current_stream().synchronize()          # thresholds are read from host memory by an async copy
    return hist, thr_counts


PAIR_METRICS = ("cosine_similarity", "cosine_distance", "angular_distance", "l1_distance", "l2_distance", "linf_distance",
                "magnitude_difference")                                       # get_all_metrics key order


def pair_metrics(A, B, ia, ib):
    """get_all_metrics (geometric_metrics.py:114-129) for the listed pairs (A[ia[p]], B[ib[p]]): (7, P) fp32 device
    tensor in PAIR_METRICS order.  B=None pairs rows of A with each other.  Unknown rows give NaN columns."""
    A = as_device_matrix(A)
    B = A if B is None else as_device_matrix(B, dtype=A.dtype)
    if B.dtype != A.dtype or B.shape[1] != A.shape[1]:
        raise ValueError("pair_metrics: A and B must share dtype and dimension")
    ia = torch.as_tensor(ia, dtype=torch.int64).to(A.device).contiguous().view(-1)
    ib = torch.as_tensor(ib, dtype=torch.int64).to(A.device).contiguous().view(-1)
    if ia.numel() != ib.numel():
        raise ValueError("pair_metrics: index lists differ in length")
    P = ia.numel()
    out = torch.empty((len(PAIR_METRICS), P), dtype=torch.float32, device=A.device)
    lib = _lib.load()
    _lib.check(lib.b200ir_pair_metrics(_dtype_id(A), _ptr(A), A.shape[0], _ptr(B), B.shape[0], A.shape[1], _ptr(ia), _ptr(ib), P,
                                       _ptr(out), _stream()), "pair_metrics")
    return out


def threshold_dedupe(scores, idx, top_k, threshold, relative=False, group=None):
    """Batched threshold + de-duplicate-by-path + cut of best-first candidate lists (image_search.py:115-140).
    scores / idx (nq, kc) as returned by topk for a descending metric; group (N,) int64 path ids or None.
    Returns (scores (nq, top_k), idx (nq, top_k) padded with -inf / -1, count (nq,) int32)."""
    dev = device()
    scores = scores.contiguous()
    idx = idx.contiguous()
    nq, kc = scores.shape
    N = 0
    if group is not None:
        group = torch.as_tensor(group, dtype=torch.int64).to(dev).contiguous()
        N = group.numel()
    out_s = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, top_k), dtype=torch.int64, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.b200ir_threshold_dedupe(_ptr(scores), _ptr(idx), nq, kc, _ptr(group) if group is not None else None, N,
                                           float(threshold), 1 if relative else 0, int(top_k), _ptr(out_s), _ptr(out_i), _ptr(cnt),
                                           _stream()), "threshold_dedupe")
    return out_s, out_i, cnt


def shortest_edge_size(H, W, size=224):
    """(new_h, new_w) when the shorter edge goes to `size` (CLIPProcessor's resize rule: new_long = int(size * long / short))."""
    short, long = (W, H) if W <= H else (H, W)
    new_long = int(size * long / short)
    return (new_long, size) if W <= H else (size, new_long)


def resize_crop(images, size=224, resized=None, crop=None):
    """(B,H,W,3) uint8 -> (B,size,size,3) uint8 on the device: PIL-exact bicubic resize of the shorter edge to `size`,
    then centre crop (the reference's CLIPProcessor front-end, ImageEmbeddingSystem.py:82-83).  `resized=(h, w)` and
    `crop=(top, left, h, w)` override the rule."""
    dev = device()
    if isinstance(images, np.ndarray):
        images = torch.from_numpy(np.ascontiguousarray(images))
    if images.dtype != torch.uint8:
        raise ValueError("images must be uint8")
    if images.dim() == 3:
        images = images.unsqueeze(0)
    if images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError(f"expected (B,H,W,3) uint8, got {tuple(images.shape)}")
    if images.device != dev:
        images = images.to(dev, non_blocking=True)
    images = images.contiguous()
    B, H, W, _ = images.shape
    rh, rw = shortest_edge_size(H, W, size) if resized is None else resized
    top, left, ch, cw = ((rh - size) // 2, (rw - size) // 2, size, size) if crop is None else crop
    lib = _lib.load()
    need = lib.b200ir_resize_crop_workspace_bytes(H, W, rh, rw, top, left, ch, cw)
    if need == 0:
        raise ValueError(f"resize_crop: unsupported geometry {(H, W)} -> {(rh, rw)} crop {(top, left, ch, cw)}")
    ws = _workspace(need, dev)
    out = torch.empty((B, ch, cw, 3), dtype=torch.uint8, device=dev)
    _lib.check(lib.b200ir_resize_crop(_ptr(images), B, H, W, rh, rw, top, left, ch, cw, _ptr(out), _ptr(ws), ws.numel(), _stream()),
               "resize_crop")
    return out


def histogram(images, colorspace="rgb"):
    """(B,H,W,3) uint8 RGB -> (B,512) int32 counts on the device (8x8x8 joint bins)."""
    dev = device()
    if isinstance(images, np.ndarray):
        images = torch.from_numpy(np.ascontiguousarray(images))
    if images.dtype != torch.uint8:
        raise ValueError("images must be uint8")
    if images.dim() == 3:
        images = images.unsqueeze(0)
    if images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError(f"expected (B,H,W,3) uint8, got {tuple(images.shape)}")
    if images.device != dev:
        images = images.to(dev, non_blocking=True)
    images = images.contiguous()
    B, H, W, _ = images.shape
    counts = torch.empty((B, 512), dtype=torch.int32, device=dev)
    cs = {"rgb": RGB, "hsv": HSV}[colorspace]
    lib = _lib.load()
    _lib.check(lib.b200ir_histogram(cs, _ptr(images), B, H, W, 8, _ptr(counts), _stream()), "histogram")
    return counts


def counts_to_embedding(counts):
    """(B,nb) int32 counts -> (raw fp32 (B,nb), unit-norm fp32 (B,nb), magnitude fp32 (B,))."""
    device()
    counts = counts.contiguous()
    B, nb = counts.shape
    raw = torch.empty((B, nb), dtype=torch.float32, device=counts.device)
    unit = torch.empty_like(raw)
    mag = torch.empty(B, dtype=torch.float32, device=counts.device)
    lib = _lib.load()
    _lib.check(lib.b200ir_counts_to_embedding(_ptr(counts), B, nb, _ptr(raw), _ptr(unit), _ptr(mag), _stream()),
               "counts_to_embedding")
    return raw, unit, mag
