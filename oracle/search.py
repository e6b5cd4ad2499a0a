"""Oracle (test infrastructure): brute-force scan + top-k semantics of the reference.

Top-k rule = Python's stable `list.sort` followed by a slice
(/root/reference/src/app_pipeline.py:171-172, :303-304, :315-316, :327-328;
image_search.py:115, :199-219): among equal keys the earlier (lower-index) entry
comes first, also with reverse=True.
"""
import numpy as np

from . import metrics as M


def topk(scores, k, descending):
    """Stable top-k per row of an (nq, N) score matrix.

    Returns (values (nq, kk), indices (nq, kk) int64), kk = min(k, N), best first, ties by
    ascending index - what `sorted(..., reverse=descending)[:k]` gives on a list in DB order.
    NaN scores are not part of the contract.
    """
    scores = np.asarray(scores)
    nq, N = scores.shape
    kk = min(k, N)
    key = -scores if descending else scores
    idx = np.argsort(key, axis=1, kind="stable")[:, :kk].astype(np.int64)
    vals = np.take_along_axis(scores, idx, axis=1)
    return vals, idx


def topk_search(Q, X, metric, k, dtype=np.float64, **kw):
    """Batched restatement of one scan: distances in `dtype`, then stable top-k."""
    S = M.pairwise(Q, X, metric, dtype=dtype, **kw)
    return topk(S, k, M.DESCENDING[metric])


def merge_topk(vals, idx, k, descending):
    """Merge R per-shard lists (R, nq, k) -> (nq, k), order (score, global index)."""
    R, nq, kk = vals.shape
    v = np.transpose(vals, (1, 0, 2)).reshape(nq, R * kk)
    i = np.transpose(idx, (1, 0, 2)).reshape(nq, R * kk)
    key = -v if descending else v
    out_v = np.empty((nq, min(k, R * kk)), dtype=vals.dtype)
    out_i = np.empty((nq, min(k, R * kk)), dtype=np.int64)
    for r in range(nq):
        valid = i[r] >= 0
        order = np.lexsort((i[r][valid], key[r][valid]))[:k]
        n = len(order)
        out_v[r, :n] = v[r][valid][order]
        out_i[r, :n] = i[r][valid][order]
        out_v[r, n:] = -np.inf if descending else np.inf
        out_i[r, n:] = -1
    return out_v, out_i


# ------------------------------------------------------------------ reference entry points
def search_images(embeddings, query_embedding, top_k=10, use_optimized_similarity=False, params=None):
    """app_pipeline.py:143-172 with the CLIP text encoder replaced by a given query vector.

    `embeddings` is an insertion-ordered {path: vector} dict.  score = abs(similarity)
    (:167), stable sort descending (:171), slice (:172); [] for an empty store (:147-149).
    """
    if not embeddings:
        return []
    results = []
    for path, embedding in embeddings.items():
        if use_optimized_similarity:
            similarity = M.optimized_similarity(query_embedding, embedding, params or {"w_angle": 1.0})
        else:
            similarity = np.dot(query_embedding, embedding) / (
                np.linalg.norm(query_embedding) * np.linalg.norm(embedding))
        results.append({"path": path, "score": abs(similarity)})
    results.sort(key=lambda x: x["score"], reverse=True)
    return results[:top_k]


def search_with_multiple_metrics(embeddings, query_embedding, top_k=5):
    """app_pipeline.py:278-372 (three scans, three stable sorts, overlap statistics)."""
    if not embeddings:
        return {"analysis": {"intersections": {}, "unique_contributions": {}}}
    out = {}
    for name, fn, sign in (("cosine_similarity", M.cosine_similarity, 1.0),
                           ("l1_distance", M.l1_distance, -1.0),
                           ("l2_distance", M.l2_distance, -1.0)):
        rows = []
        for path, emb in embeddings.items():
            v = fn(query_embedding, emb)
            rows.append({"path": path, name: v, "score": sign * v})
        rows.sort(key=lambda x: x["score"], reverse=True)
        out[name] = rows[:top_k]
    out["analysis"] = overlap_analysis(out, top_k)
    return out


def overlap_analysis(results_by_metric, top_k):
    """app_pipeline.py:331-370."""
    cp = set(r["path"] for r in results_by_metric["cosine_similarity"])
    l1 = set(r["path"] for r in results_by_metric["l1_distance"])
    l2 = set(r["path"] for r in results_by_metric["l2_distance"])

    def inter(a, b):
        return {"intersection_size": len(a & b),
                "intersection_ratio": len(a & b) / top_k if top_k > 0 else 0}

    allp = cp | l1 | l2

    def uniq(a, b, c):
        return {"unique_count": len(a - b - c),
                "unique_ratio": len(a - b - c) / len(allp) if allp else 0}

    return {"intersections": {"cosine_vs_l1": inter(cp, l1), "cosine_vs_l2": inter(cp, l2),
                              "l1_vs_l2": inter(l1, l2)},
            "unique_contributions": {"cosine_similarity": uniq(cp, l1, l2),
                                     "l1_distance": uniq(l1, cp, l2),
                                     "l2_distance": uniq(l2, cp, l1)}}


def threshold_and_dedupe(matches, top_k, score_threshold, use_optimized_similarity):
    """image_search.py:115-140: stable sort desc, threshold (absolute, or min+t*(max-min) for the
    optimized score), de-duplicate by path keeping first, cut to top_k."""
    matches = sorted(matches, key=lambda x: x["score"], reverse=True)
    if use_optimized_similarity:
        mn = min(m["score"] for m in matches) if matches else 0
        mx = max(m["score"] for m in matches) if matches else 1
        thr = mn + score_threshold * (mx - mn)
        filtered = [m for m in matches if m["score"] >= thr]
    else:
        filtered = [m for m in matches if m["score"] >= score_threshold]
    seen, unique = set(), []
    for m in filtered:
        if m["path"] not in seen:
            seen.add(m["path"])
            unique.append(m)
            if len(unique) >= top_k:
                break
    return unique[:top_k]
