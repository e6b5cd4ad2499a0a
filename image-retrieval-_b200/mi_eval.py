"""Distance / threshold arithmetic of the reference's evaluation workload
(/root/reference/src/mi_analysis.py:256-297 calculate_distances, :704-713 per-metric densities, :774-796
precision / recall over thresholds) on the B200 hot path.  Only the counting arithmetic lives here: mutual
information, KDE plots and the weight grid search are CPU statistics on top of these counts and stay out of scope.
"""
import logging
import os
from pathlib import Path

import numpy as np

from . import ops

logger = logging.getLogger(__name__)

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


def load_embeddings_file(embeddings_file):
    """{path: vector} from the reference's embedding artefacts: an .npz holding a pickled dict under 'embeddings'
    (app_pipeline.py:34-58 / mi_analysis.py:240-247) or a .npy holding the dict itself (:248-250)."""
    data = np.load(embeddings_file, allow_pickle=True)
    if isinstance(data, np.lib.npyio.NpzFile):
        if "embeddings" not in data:
            raise KeyError(f"No 'embeddings' array found in {embeddings_file}")
        return data["embeddings"].item()
    return data.item()


def generate_relationship_pairs(metadata, categories=None, colors=None):
    """The four pair lists of ColorDatasetManager.generate_relationship_pairs (imageProcessing.py:296-387) from metadata
    rows {'path', 'category', 'color'}, in the reference's enumeration order."""
    rows = metadata.to_dict("records") if hasattr(metadata, "to_dict") else list(metadata)
    pairs = {r: [] for r in RELATIONSHIP_TYPES}
    if len(rows) < 2:
        return pairs
    groups = {}
    for meta in rows:
        groups.setdefault((meta["category"], meta["color"]), []).append(meta["path"])
    categories = list(dict.fromkeys(c for c, _ in groups)) if categories is None else list(categories)
    colors = list(dict.fromkeys(c for _, c in groups)) if colors is None else list(colors)
    for paths in groups.values():
        pairs["same_object_same_color"] += [(paths[i], paths[j]) for i in range(len(paths)) for j in range(i + 1, len(paths))]
    for category in categories:
        cc = [col for (cat, col) in groups if cat == category]
        for a, c1 in enumerate(cc):
            for c2 in cc[a + 1:]:
                pairs["same_object_diff_color"] += [(p1, p2) for p1 in groups[(category, c1)] for p2 in groups[(category, c2)]]
    for color in colors:
        cats = [cat for (cat, col) in groups if col == color]
        for a, k1 in enumerate(cats):
            for k2 in cats[a + 1:]:
                pairs["diff_object_same_color"] += [(p1, p2) for p1 in groups[(k1, color)] for p2 in groups[(k2, color)]]
    cat_list = list(dict.fromkeys(cat for cat, _ in groups))
    for a, k1 in enumerate(cat_list):
        for k2 in cat_list[a + 1:]:
            for c1 in [col for (c, col) in groups if c == k1]:
                for c2 in [col for (c, col) in groups if c == k2]:
                    if c1 != c2:
                        pairs["diff_object_diff_color"] += [(p1, p2) for p1 in groups[(k1, c1)] for p2 in groups[(k2, c2)]]
    return pairs


def save_pairs(base_dir, pairs):
    """pairs.json as ColorDatasetManager.create_dataset writes it (imageProcessing.py:421-434): paths relative to base_dir."""
    import json
    base_str = str(base_dir) + os.sep
    out = {r: [(p1[len(base_str):] if p1.startswith(base_str) else p1, p2[len(base_str):] if p2.startswith(base_str) else p2)
               for p1, p2 in lst] for r, lst in pairs.items()}
    with open(os.path.join(str(base_dir), "pairs.json"), "w") as f:
        json.dump(out, f)


class ColorMIAnalyzer:
    """Data-loading and distance stage of the reference's ColorMIAnalyzer (mi_analysis.py:155-297): metadata.csv +
    pairs.json + embeddings file -> distances[metric][relationship_type].  The per-pair get_all_metrics loop
    (:277-291) is ONE b200ir_pair_metrics launch over all listed pairs; the MI / KDE / grid-search statistics that
    consume these lists stay CPU-side and out of scope."""

    def __init__(self, base_dir="color_dataset", bin_count=20, bin_strategy="uniform"):
        from .geometric_metrics import GeometricSimilarityMetrics
        self.base_dir = Path(base_dir)
        self.bin_count = bin_count
        self.bin_strategy = bin_strategy
        self.metrics = GeometricSimilarityMetrics()
        self.relationship_types = list(RELATIONSHIP_TYPES)
        self.metric_names = list(METRIC_NAMES)
        self.embeddings = {}
        self.metadata = None
        self.pairs = {}
        self.distances = {}
        self.mi_results = {}
        self.optimal_weights = {}

    def load_dataset(self, embeddings_file):
        """(success, message) with the reference's messages (mi_analysis.py:199-254)."""
        import json
        metadata_path = self.base_dir / "metadata.csv"
        if not metadata_path.exists():
            return False, f"Metadata file not found: {metadata_path}"
        import pandas as pd
        self.metadata = pd.read_csv(metadata_path)
        pairs_path = self.base_dir / "pairs.json"
        if not pairs_path.exists():
            return False, f"Pairs file not found: {pairs_path}"
        with open(pairs_path, "r") as f:
            raw_pairs = json.load(f)
        for rel_type, rel_pairs in raw_pairs.items():
            self.pairs[rel_type] = [(p1 if os.path.isabs(p1) else os.path.join(self.base_dir, p1),
                                     p2 if os.path.isabs(p2) else os.path.join(self.base_dir, p2)) for p1, p2 in rel_pairs]
        try:
            self.embeddings = load_embeddings_file(embeddings_file)
            return True, "Dataset loaded successfully"
        except KeyError as e:
            return False, str(e.args[0])
        except Exception as e:
            return False, f"Error loading embeddings: {str(e)}"

    def _pair_indices(self, rel_pairs, row_of):
        ia, ib = [], []
        for p1, p2 in rel_pairs:
            if p1 not in row_of or p2 not in row_of:        # :278-280: warn and skip
                logger.warning(f"Embeddings not found for {p1} or {p2}")
                continue
            ia.append(row_of[p1]); ib.append(row_of[p2])
        return ia, ib

    def calculate_distances(self):
        """distances[metric][rel_type] = fp32 array, one value per listed pair whose two embeddings exist, in list order."""
        self.distances = {m: {r: [] for r in self.relationship_types} for m in self.metric_names}
        if not self.embeddings:
            return
        paths = list(self.embeddings)
        row_of = {p: i for i, p in enumerate(paths)}
        X = ops.as_device_matrix(np.stack([np.asarray(self.embeddings[p], dtype=np.float32) for p in paths]))
        spans, ia, ib = {}, [], []
        for rel_type in self.relationship_types:
            if rel_type not in self.pairs:
                logger.warning(f"No pairs found for relationship type: {rel_type}")
                continue
            a, b = self._pair_indices(self.pairs[rel_type], row_of)
            spans[rel_type] = (len(ia), len(ia) + len(a))
            ia += a; ib += b
        if not ia:
            return
        vals = ops.pair_metrics(X, None, ia, ib).cpu().numpy()
        for m in self.metric_names:
            row = vals[ops.PAIR_METRICS.index(m)]
            for rel_type, (b, e) in spans.items():
                self.distances[m][rel_type] = row[b:e]

    def precision_recall(self, metric="cosine_distance", thresholds=None):
        """(thresholds, precision, recall) of mi_analysis.py:741-796 from the stored distance lists."""
        thresholds = np.linspace(0, 1, 100) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
        pos = np.sort(np.asarray(self.distances[metric]["same_object_diff_color"], dtype=np.float64))
        neg = np.sort(np.asarray(self.distances[metric]["same_object_same_color"], dtype=np.float64))
        tp = np.searchsorted(pos, thresholds, side="right")
        fp = np.searchsorted(neg, thresholds, side="right")
        fn = len(pos) - tp
        with np.errstate(divide="ignore", invalid="ignore"):
            precision = np.where(tp + fp > 0, tp / np.maximum(tp + fp, 1), 0.0)
            recall = np.where(tp + fn > 0, tp / np.maximum(tp + fn, 1), 0.0)
        return thresholds, precision, recall
