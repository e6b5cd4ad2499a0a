"""EnhancedImageSearchApp / SimpleSearcher - the reference's brute-force search surface
(/root/reference/src/app_pipeline.py:14-390) on the B200 hot path.

What is kept: the `embeddings` dict, `search_images(query, top_k, use_optimized_similarity)`,
`search_with_multiple_metrics(query, top_k)`, `SimpleSearcher.set_similarity_params`, result
shapes (lists of {'path','score'} dicts, dict of lists + 'analysis'), empty-store behaviour.
What changes underneath: the per-embedding Python loop + list.sort + slice becomes one fused
distance + top-k kernel launch over an HBM-resident matrix (ops.topk).  CLIP, tkinter and the
MI plots are out of scope: `query` may be an embedding vector, an RGB
uint8 image (embedded with the colour-histogram kernel) or - if a `text_encoder` callable was
given - a string.
"""
import logging
import os
from pathlib import Path

import numpy as np
import torch

from . import ops
from .config import EMBEDDING_DIM

logger = logging.getLogger(__name__)


class _TrackedDict(dict):
    """dict that counts mutations so the device copy of the store can be rebuilt lazily."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.version = 0

    def _bump(self):
        self.version += 1

    def __setitem__(self, k, v):
        super().__setitem__(k, v); self._bump()

    def __delitem__(self, k):
        super().__delitem__(k); self._bump()

    def update(self, *a, **kw):
        super().update(*a, **kw); self._bump()

    def clear(self):
        super().clear(); self._bump()

    def pop(self, *a):
        r = super().pop(*a); self._bump(); return r

    def popitem(self):
        r = super().popitem(); self._bump(); return r

    def setdefault(self, k, d=None):
        r = super().setdefault(k, d); self._bump(); return r

    def __ior__(self, other):
        super().update(other); self._bump(); return self

    def invalidate(self):
        """Call after editing a stored row IN PLACE (e.g. app.embeddings[p][:] = v): the device copy of the store is
        rebuilt on the next search.  Assignments, deletions, update() and |= are tracked automatically."""
        self._bump()


class EnhancedImageSearchApp:
    """Enhanced image search application with geometric metrics."""

    def __init__(self, text_encoder=None, dtype=torch.float32):
        self._embeddings = _TrackedDict()
        self.searcher = SimpleSearcher()
        self.text_encoder = text_encoder
        self.dtype = dtype
        self._paths = []
        self._matrix = None
        self._synced = None       # (id(dict), version) of the dict the device matrix mirrors

    # the reference exposes a plain attribute; assignments of ordinary dicts are wrapped
    @property
    def embeddings(self):
        return self._embeddings

    @embeddings.setter
    def embeddings(self, value):
        self._embeddings = value if isinstance(value, _TrackedDict) else _TrackedDict(value)
        self._synced = None

    # ------------------------------------------------------------------ store
    def set_embeddings(self, paths, matrix):
        """Fast path: adopt an (N, D) matrix (host or device) without building a dict of rows."""
        self._paths = [str(p) for p in paths]
        self._matrix = ops.as_device_matrix(matrix, dtype=self.dtype)
        if len(self._paths) != self._matrix.shape[0]:
            raise ValueError("set_embeddings: len(paths) != rows")
        self._embeddings = _LazyRows(self._paths, self._matrix)
        self._synced = (id(self._embeddings), self._embeddings.version)

    def _store(self):
        e = self._embeddings
        key = (id(e), e.version)
        if self._synced != key:
            self._paths = list(e.keys())
            rows = np.stack([np.asarray(v, dtype=np.float32).reshape(-1) for v in e.values()]) if e else \
                np.zeros((0, EMBEDDING_DIM), np.float32)
            self._matrix = ops.as_device_matrix(rows, dtype=self.dtype)
            self._synced = key
        return self._paths, self._matrix

    def process_image_arrays(self, paths, images, colorspace="rgb"):
        """Embed (B,H,W,3) uint8 RGB images with the colour-histogram kernel and store them."""
        counts = ops.histogram(images, colorspace)
        raw, _unit, _mag = ops.counts_to_embedding(counts)
        host = raw.cpu().numpy()
        for p, v in zip(paths, host):
            self._embeddings[str(p)] = v
        return len(paths)

    # the reference's on-disk cache format: np.savez(file, embeddings={path: vector}) (a pickled dict,
    # color_analysis_workflow.py:145, app_pipeline.py:124), searched for at these places (app_pipeline.py:34-42)
    EMBEDDING_CACHE_PATHS = (
        "color_embeddings.npz", "color_analysis/color_embeddings.npz", "../color_embeddings.npz", "embeddings.npz",
        "color_dataset/embeddings.npz", os.path.expanduser("~/Desktop/color_embeddings.npz"),
        os.path.expanduser("~/Desktop/color_analysis/color_embeddings.npz"))

    @staticmethod
    def load_embeddings_npz(path):
        """{path: vector} dict from the reference's .npz cache (app_pipeline.py:54-58, mi_analysis.py:239-248)."""
        data = np.load(path, allow_pickle=True)
        if 'embeddings' not in data:
            raise KeyError(f"{path} has no 'embeddings' entry")
        return data['embeddings'].item()

    def save_embeddings_npz(self, path):
        """Write the store in the reference's cache format (app_pipeline.py:124)."""
        np.savez(path, embeddings=dict(self._embeddings.items()))

    def process_images(self, image_paths, colorspace="rgb", embeddings_file=None):
        """app_pipeline.py:29-90: adopt pre-computed embeddings from an .npz cache when one matches the selected
        images (exact path first, then file name), else compute embeddings - here colour-histogram embeddings of
        the image files (PIL decode on the host, histogram on the device; CLIP is out of scope).  Unreadable
        files are skipped with a warning, like the reference's per-image try/except."""
        logger.info(f"Processing {len(image_paths)} images...")
        candidates = [embeddings_file] if embeddings_file else list(self.EMBEDDING_CACHE_PATHS)
        cache = next((c for c in candidates if c and os.path.exists(c)), None)
        if cache:
            try:
                stored = self.load_embeddings_npz(cache)
                by_name = {}
                for sp in stored:
                    by_name.setdefault(Path(sp).name, sp)          # first stored path wins, as in the reference loop
                matched = {}
                for image_path in image_paths:
                    sp = str(image_path)
                    if sp in stored:
                        matched[sp] = stored[sp]
                    elif Path(image_path).name in by_name:
                        matched[sp] = stored[by_name[Path(image_path).name]]
                if matched:
                    self._embeddings.update(matched)
                    logger.info(f"Successfully matched {len(matched)}/{len(image_paths)} images with embeddings")
                    return len(matched)
                logger.warning("No matching embeddings found for selected images")
            except Exception as e:  # noqa: BLE001 - mirror of the reference's broad handler
                logger.warning(f"Failed to load pre-computed embeddings: {e}")
        from PIL import Image
        done = 0
        for path in image_paths:
            try:
                with Image.open(path) as im:
                    arr = np.asarray(im.convert("RGB"), dtype=np.uint8)
                done += self.process_image_arrays([path], arr[None], colorspace)
            except Exception as e:  # noqa: BLE001
                logger.warning(f"Skipping {path} due to error: {e}")
        return done

    def _generate_dummy_embeddings(self, image_paths):
        """Generate dummy embeddings as fallback (app_pipeline.py:136-141)."""
        logger.info("Generating dummy embeddings...")
        for path in image_paths:
            self._embeddings[str(path)] = np.random.randn(EMBEDDING_DIM)
        logger.info(f"Generated {len(self._embeddings)} dummy embeddings")

    # ------------------------------------------------------------------ queries
    def _get_query_embedding(self, query):
        """Reference: CLIP text tower, random vector on any failure (app_pipeline.py:174-191)."""
        if isinstance(query, str):
            if self.text_encoder is not None:
                try:
                    return np.asarray(self.text_encoder(query), dtype=np.float32).reshape(-1)
                except Exception as e:  # noqa: BLE001
                    logger.warning(f"Error generating query embedding: {e}, using random")
            else:
                logger.warning("No text encoder configured (CLIP is out of scope), using random")
            return np.random.randn(EMBEDDING_DIM)
        if isinstance(query, np.ndarray) and query.dtype == np.uint8 and query.ndim == 3:
            raw, _, _ = ops.counts_to_embedding(ops.histogram(query[None]))
            return raw[0]
        return query

    # ------------------------------------------------------------------ search
    def search_images(self, query, top_k=10, use_optimized_similarity=False):
        """Search images (app_pipeline.py:143-172): score = |similarity|, stable sort descending, top_k."""
        logger.info(f"Searching (optimized: {use_optimized_similarity})")
        if not self._embeddings:
            logger.warning("No embeddings available for search")
            return []
        paths, X = self._store()
        q = self._get_query_embedding(query)
        if top_k <= 0:
            return []
        k = min(int(top_k), len(paths))               # results[:top_k] for any top_k (:172); lists beyond one page are paged
        if k > ops.MAX_K_PAGED:
            raise ValueError(f"search_images: top_k up to {ops.MAX_K_PAGED} rows is supported, got {top_k}")
        if use_optimized_similarity:
            s, i = ops.topk(q, X, "optimized_similarity", k, abs_score=True, params=self.searcher.similarity_params)
        else:
            s, i = ops.topk(q, X, "cosine_similarity", k, abs_score=True)
        s, i = s[0].cpu().numpy(), i[0].cpu().numpy()
        return [{'path': paths[j], 'score': sc} for sc, j in zip(s, i) if j >= 0]

    def search_with_multiple_metrics(self, query, top_k=5):
        """Search with multiple geometric metrics (app_pipeline.py:278-372)."""
        if not self._embeddings:
            return {'analysis': {'intersections': {}, 'unique_contributions': {}}}
        paths, X = self._store()
        q = self._get_query_embedding(query)
        k = max(1, min(int(top_k), len(paths)))
        names = (('cosine_similarity', 1.0), ('l1_distance', -1.0), ('l2_distance', -1.0))
        if k <= ops.MAX_K:
            # the reference's three scans + three sorts (:296-328) as ONE pass that keeps three candidate lists
            S, I = ops.topk_multi(q, X, [n for n, _ in names], k)
            S, I = S[:, 0].cpu().numpy(), I[:, 0].cpu().numpy()
        else:
            if k > ops.MAX_K_PAGED:
                raise ValueError(f"search_with_multiple_metrics: top_k up to {ops.MAX_K_PAGED} rows is supported, got {top_k}")
            pages = [ops.topk(q, X, n, k) for n, _ in names]
            S = np.stack([p[0][0].cpu().numpy() for p in pages])
            I = np.stack([p[1][0].cpu().numpy() for p in pages])
        results_by_metric = {}
        for y, (name, sign) in enumerate(names):
            results_by_metric[name] = [{'path': paths[j], name: v, 'score': sign * v}
                                       for v, j in zip(S[y], I[y]) if j >= 0][:max(0, int(top_k))]
        results_by_metric['analysis'] = _overlap_analysis(results_by_metric, top_k)
        return results_by_metric

    # ------------------------------------------------------------------ out of scope
    def run_mi_analysis(self, *a, **kw):
        raise NotImplementedError("MI analysis / plotting is outside the retrieval hot path (SURVEY.md section 2.1)")

    run_enhanced_mi_analysis = run_mi_analysis


def _overlap_analysis(results_by_metric, top_k):
    """Set intersections / unique contributions (app_pipeline.py:331-370)."""
    cosine_paths = set(r['path'] for r in results_by_metric['cosine_similarity'])
    l1_paths = set(r['path'] for r in results_by_metric['l1_distance'])
    l2_paths = set(r['path'] for r in results_by_metric['l2_distance'])

    def inter(a, b):
        return {'intersection_size': len(a & b), 'intersection_ratio': len(a & b) / top_k if top_k > 0 else 0}

    all_paths = cosine_paths | l1_paths | l2_paths

    def uniq(a, b, c):
        return {'unique_count': len(a - b - c), 'unique_ratio': len(a - b - c) / len(all_paths) if all_paths else 0}

    return {
        'intersections': {'cosine_vs_l1': inter(cosine_paths, l1_paths), 'cosine_vs_l2': inter(cosine_paths, l2_paths),
                          'l1_vs_l2': inter(l1_paths, l2_paths)},
        'unique_contributions': {'cosine_similarity': uniq(cosine_paths, l1_paths, l2_paths),
                                 'l1_distance': uniq(l1_paths, cosine_paths, l2_paths),
                                 'l2_distance': uniq(l2_paths, cosine_paths, l1_paths)},
    }


class _LazyRows(_TrackedDict):
    """Read-only dict view {path: row} over a device matrix (rows fetched on demand)."""

    def __init__(self, paths, matrix):
        super().__init__()
        self._index = {p: i for i, p in enumerate(paths)}
        self._m = matrix
        for p in paths:
            dict.__setitem__(self, p, None)

    def __getitem__(self, k):
        return self._m[self._index[k]].float().cpu().numpy()

    def __setitem__(self, k, v):
        raise TypeError("a store adopted with set_embeddings() is read-only; assign app.embeddings = {...} to edit it")

    def items(self):
        return ((k, self[k]) for k in self.keys())

    def values(self):
        return (self[k] for k in self.keys())


class SimpleSearcher:
    """Simple searcher class for compatibility (app_pipeline.py:375-390)."""

    def __init__(self):
        self.similarity_params = {
            'w_angle': 1.0,
            'w_l1': 0.0,
            'w_l2': 0.0,
            'w_inf': 0.0,
            'w_mag': 0.0
        }

    def set_similarity_params(self, params):
        """Set similarity parameters."""
        self.similarity_params.update(params)
        logger.info(f"Updated similarity parameters: {self.similarity_params}")
