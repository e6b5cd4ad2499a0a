"""Oracle (test infrastructure): the reference's evaluation workload on explicit pair lists.

Follows /root/reference/src/mi_analysis.py: relationship types :176-181, per-pair metrics via get_all_metrics
:256-297 (metric list :183-189), precision / recall over thresholds :774-796.  The reference works on sampled
pair lists; the all-pairs variant here enumerates every unordered pair i < j.
"""
import numpy as np

from . import metrics as M

METRICS = ("cosine_distance", "l1_distance", "l2_distance", "linf_distance", "magnitude_difference")
_PAIRWISE = {"cosine_distance": "cosine_distance", "l1_distance": "l1", "l2_distance": "l2", "linf_distance": "linf",
             "magnitude_difference": "magnitude_difference"}
RELATIONSHIP_TYPES = ("same_object_same_color", "same_object_diff_color", "diff_object_same_color", "diff_object_diff_color")


def relationship(cat, col):
    """(N, N) matrix of relationship-type ids (mi_analysis.py:176-181 order)."""
    cat = np.asarray(cat)
    col = np.asarray(col)
    return np.where(cat[:, None] == cat[None, :], 0, 2) + np.where(col[:, None] == col[None, :], 0, 1)


def metric_matrices(X, dtype=np.float64):
    return {m: M.pairwise(X, X, _PAIRWISE[m], dtype=dtype) for m in METRICS}


def bin_counts(values, rel, ranges, nbins, thresholds):
    """Counts from explicit (N, N) metric matrices (only i < j is used).  Binning arithmetic is the kernel's:
    fp32 floor((v - lo) * (nbins / (hi - lo))) clamped to [0, nbins); thresholds compared in fp64."""
    N = rel.shape[0]
    iu = np.triu_indices(N, 1)
    r = rel[iu]
    hist = np.zeros((len(METRICS), 4, nbins), dtype=np.int64)
    thr = np.zeros((len(METRICS), 2, len(thresholds) + 1), dtype=np.int64)
    thresholds = np.asarray(thresholds, dtype=np.float64)
    for mi, m in enumerate(METRICS):
        v = np.asarray(values[m], dtype=np.float32)[iu]
        lo, hi = np.float32(ranges[m][0]), np.float32(ranges[m][1])
        inv_w = np.float32(nbins) / (hi - lo)
        b = np.floor((v - lo) * inv_w).astype(np.int64).clip(0, nbins - 1)
        for t in range(4):
            hist[mi, t] = np.bincount(b[r == t], minlength=nbins)
        first = np.searchsorted(thresholds, v.astype(np.float64), side="left")      # first t with v <= thresholds[t]
        for lab in (0, 1):
            thr[mi, lab] = np.bincount(first[r == lab], minlength=len(thresholds) + 1)
    return hist, thr


def pr_curve_reference(distances, labels, thresholds):
    """mi_analysis.py:774-796, verbatim loop: (tp, fp, fn) per threshold."""
    out = []
    for threshold in thresholds:
        predictions = [1 if d <= threshold else 0 for d in distances]
        tp = sum(pred == 1 and label == 1 for pred, label in zip(predictions, labels))
        fp = sum(pred == 1 and label == 0 for pred, label in zip(predictions, labels))
        fn = sum(pred == 0 and label == 1 for pred, label in zip(predictions, labels))
        out.append((tp, fp, fn))
    return np.array(out, dtype=np.int64)


def pr_from_counts(thr_counts_metric):
    """(2, nthr + 1) first-threshold-index counts -> (nthr, 3) tp / fp / fn."""
    c0, c1 = thr_counts_metric[0], thr_counts_metric[1]
    tp = np.cumsum(c1[:-1])
    fp = np.cumsum(c0[:-1])
    fn = c1.sum() - tp
    return np.stack([tp, fp, fn], axis=1)


def calculate_distances(embeddings, pairs, metric_names=METRICS, relationship_types=RELATIONSHIP_TYPES):
    """mi_analysis.py:256-297 restated: distances[metric][rel_type] = list of get_all_metrics values over the listed
    (path1, path2) pairs, skipping pairs with a missing embedding (:278-280)."""
    distances = {m: {r: [] for r in relationship_types} for m in metric_names}
    for rel_type in relationship_types:
        for p1, p2 in pairs.get(rel_type, []):
            if p1 not in embeddings or p2 not in embeddings:
                continue
            allm = M.get_all_metrics(embeddings[p1], embeddings[p2])
            for m in metric_names:
                distances[m][rel_type].append(allm[m])
    return distances


def generate_relationship_pairs(metadata, categories, colors):
    """imageProcessing.py:296-387 restated (loop structure kept): metadata rows -> four pair lists."""
    from collections import defaultdict
    pairs = {r: [] for r in RELATIONSHIP_TYPES}
    if len(metadata) < 2:
        return pairs
    groups = defaultdict(list)
    for meta in metadata:
        groups[(meta["category"], meta["color"])].append(meta["path"])
    for (category, color), paths in list(groups.items()):
        if len(paths) >= 2:
            for i in range(len(paths)):
                for j in range(i + 1, len(paths)):
                    pairs["same_object_same_color"].append((paths[i], paths[j]))
    for category in categories:
        cc = [color for (cat, color), paths in groups.items() if cat == category and paths]
        if len(cc) >= 2:
            for i1, c1 in enumerate(cc):
                for c2 in cc[i1 + 1:]:
                    for a in groups[(category, c1)]:
                        for b in groups[(category, c2)]:
                            pairs["same_object_diff_color"].append((a, b))
    for color in colors:
        cats = [cat for (cat, col), paths in groups.items() if col == color and paths]
        if len(cats) >= 2:
            for i1, k1 in enumerate(cats):
                for k2 in cats[i1 + 1:]:
                    for a in groups[(k1, color)]:
                        for b in groups[(k2, color)]:
                            pairs["diff_object_same_color"].append((a, b))
    return pairs    # diff_object_diff_color iterates a python set (:358-360): its order is unspecified, compared as a set
