"""Distance / threshold arithmetic of the reference's evaluation workload
(/root/reference/src/mi_analysis.py:256-297 calculate_distances, :704-713 per-metric densities, :774-796
precision / recall over thresholds) on the B200 hot path.  Only the counting arithmetic lives here: mutual
information, KDE plots and the weight grid search are CPU statistics on top of these counts and stay out of scope.
"""
import numpy as np

from . import ops

METRIC_NAMES = list(ops.EVAL_METRICS)                 # mi_analysis.py:183-189
RELATIONSHIP_TYPES = list(ops.RELATIONSHIP_TYPES)     # mi_analysis.py:176-181


class AllPairsEvaluator:
    """All-pairs variant of ColorMIAnalyzer.calculate_distances + the PR counting loop."""

    def __init__(self, embeddings, category, color, ranges=None, nbins=1024, thresholds=None):
        self.X = ops.as_device_matrix(embeddings)
        self.category, self.color = category, color
        self.nbins = nbins
        self.thresholds = np.linspace(0, 1, 100) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
        self.ranges = ranges or {"cosine_distance": (0.0, 2.0), "l1_distance": (0.0, 2.0), "l2_distance": (0.0, 2.0),
                                 "linf_distance": (0.0, 2.0), "magnitude_difference": (0.0, 2.0)}
        self.hist = None
        self.thr_counts = None

    def calculate_distances(self):
        """Per (metric, relationship type) distance histograms over every pair i < j (device tensors kept on self)."""
        self.hist, self.thr_counts = ops.allpairs_eval(self.X, self.category, self.color, self.ranges, self.nbins, self.thresholds)
        return self.hist

    def densities(self):
        """{metric: {relationship: (bin_centres, density)}} - histogram analogue of the reference's KDE curves."""
        if self.hist is None:
            self.calculate_distances()
        h = self.hist.cpu().numpy().astype(np.float64)
        out = {}
        for mi, m in enumerate(METRIC_NAMES):
            lo, hi = self.ranges[m]
            w = (hi - lo) / self.nbins
            centres = lo + (np.arange(self.nbins) + 0.5) * w
            out[m] = {r: (centres, h[mi, ri] / max(h[mi, ri].sum(), 1.0) / w) for ri, r in enumerate(RELATIONSHIP_TYPES)}
        return out

    def precision_recall(self, metric="cosine_distance"):
        """(thresholds, precision, recall) with label 1 = same_object_diff_color, 0 = same_object_same_color and
        prediction d <= threshold, exactly the counts of mi_analysis.py:783-796."""
        if self.thr_counts is None:
            self.calculate_distances()
        c = self.thr_counts[METRIC_NAMES.index(metric)].cpu().numpy()
        tp = np.cumsum(c[1][:-1])
        fp = np.cumsum(c[0][:-1])
        fn = c[1].sum() - tp
        with np.errstate(divide="ignore", invalid="ignore"):
            precision = np.where(tp + fp > 0, tp / np.maximum(tp + fp, 1), 0.0)
            recall = np.where(tp + fn > 0, tp / np.maximum(tp + fn, 1), 0.0)
        return self.thresholds, precision, recall
