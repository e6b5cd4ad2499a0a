"""GeometricSimilarityMetrics - the reference's metric surface
(/root/reference/src/geometric_metrics.py:8-149) backed by the B200 kernels.

Scalar methods keep the reference's names, argument meaning, return types and error behaviour
(they never raise on zero vectors; zero norm -> 0.0).  Each scalar call is a 1x1 instance of the
batched operators in `ops` - use `pairwise` / `topk` (new, batched) for real workloads.
"""
from typing import Dict, List

import numpy as np

from . import ops


def _vec(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(1, -1))


def _one(metric, vec1, vec2, **kw):
    return np.float32(ops.pairwise(_vec(vec1), _vec(vec2), metric, **kw).item())


def _seven(vec1, vec2):
    """The seven get_all_metrics values of one pair from ONE kernel launch and one device->host copy
    (b200ir_pair_metrics), as np.float32 in ops.PAIR_METRICS order."""
    a, b = _vec(vec1), _vec(vec2)
    if a.shape[1] != b.shape[1]:
        raise ValueError(f"operands could not be broadcast together with shapes ({a.shape[1]},) ({b.shape[1]},)")
    return ops.pair_metrics(a, b, [0], [0]).cpu().numpy()[:, 0]


class GeometricSimilarityMetrics:
    """Class implementing various geometric similarity metrics for embeddings."""

    # ------------------------------------------------------------------ batched (new)
    @staticmethod
    def pairwise(queries, database, metric, **kw):
        """(nq, N) matrix of `metric` on the device (torch tensor)."""
        return ops.pairwise(queries, database, metric, **kw)

    @staticmethod
    def topk(queries, database, metric, k, **kw):
        """Fused scan + top-k: (scores (nq,k), indices (nq,k)) device tensors."""
        return ops.topk(queries, database, metric, k, **kw)

    # ------------------------------------------------------------------ scalar (reference surface)
    @staticmethod
    def cosine_similarity(vec1: np.ndarray, vec2: np.ndarray) -> float:
        """Computes the cosine similarity between two vectors (geometric_metrics.py:12-18)."""
        c = _seven(vec1, vec2)[0]             # zero norm -> 0.0 inside the kernel (:16-17)
        return 0.0 if c == 0 else c           # the reference's zero-norm branch returns the Python float 0.0

    @staticmethod
    def angular_distance(vec1: np.ndarray, vec2: np.ndarray) -> float:
        """Computes the angular distance in radians (geometric_metrics.py:21-26)."""
        cos_sim = GeometricSimilarityMetrics.cosine_similarity(vec1, vec2)
        cos_sim = np.clip(cos_sim, -1.0, 1.0)
        return np.arccos(cos_sim)

    @staticmethod
    def cosine_distance(vec1: np.ndarray, vec2: np.ndarray) -> float:
        """Computes cosine distance (1 - cosine similarity) (geometric_metrics.py:29-31)."""
        return 1.0 - GeometricSimilarityMetrics.cosine_similarity(vec1, vec2)

    @staticmethod
    def l1_distance(vec1: np.ndarray, vec2: np.ndarray, normalized: bool = True) -> float:
        """Computes the L1 (Manhattan) distance (geometric_metrics.py:34-39)."""
        distance = _one("l1", vec1, vec2, normalized=False)
        if normalized:
            distance /= len(vec1)
        return distance

    @staticmethod
    def l2_distance(vec1: np.ndarray, vec2: np.ndarray, normalized: bool = True) -> float:
        """Computes the L2 (Euclidean) distance (geometric_metrics.py:42-47); like the reference the
        normalised result is a float64 (fp32 value / np.sqrt(D))."""
        distance = _one("l2", vec1, vec2, normalized=False)
        if normalized:
            distance /= np.sqrt(len(vec1))
        return distance

    @staticmethod
    def linf_distance(vec1: np.ndarray, vec2: np.ndarray) -> float:
        """Computes the L-infinity (Chebyshev) distance (geometric_metrics.py:50-52)."""
        return _one("linf", vec1, vec2)

    @staticmethod
    def magnitude_difference(vec1: np.ndarray, vec2: np.ndarray) -> float:
        """Computes the absolute difference in vector magnitudes (geometric_metrics.py:55-57)."""
        return _one("magnitude_difference", vec1, vec2)

    @staticmethod
    def optimized_similarity(vec1: np.ndarray, vec2: np.ndarray, params: Dict[str, float]) -> float:
        """Weighted combination w_angle*cos - w_l1*L1 - w_l2*L2 - w_inf*Linf - w_mag*mag
        (geometric_metrics.py:60-94); one fused kernel pass, fp32, returned as float64."""
        return np.float64(_one("optimized_similarity", vec1, vec2, params=dict(params)))

    @staticmethod
    def optimized_distance(vec1: np.ndarray, vec2: np.ndarray, params: Dict[str, float]) -> float:
        """Negated optimized_similarity (geometric_metrics.py:97-111)."""
        return -GeometricSimilarityMetrics.optimized_similarity(vec1, vec2, params)

    @staticmethod
    def get_all_metrics(vec1: np.ndarray, vec2: np.ndarray) -> Dict[str, float]:
        """All seven metrics between two vectors (geometric_metrics.py:114-129)."""
        v = _seven(vec1, vec2)                # one launch for all seven values (the reference recomputes cos 3x, norms 8x)
        return {
            'cosine_similarity': v[0],
            'cosine_distance': v[1],
            'angular_distance': v[2],
            'l1_distance': v[3],
            'l2_distance': np.float64(v[4]),   # the reference's l2 is fp32 / np.sqrt(int) -> float64 (:46)
            'linf_distance': v[5],
            'magnitude_difference': v[6]
        }

    @staticmethod
    def create_parameter_grid(granularity: int = 5) -> Dict[str, List[float]]:
        """Grid of weight values for optimisation (geometric_metrics.py:132-149)."""
        values = np.linspace(0.0, 1.0, granularity)
        return {
            'w_angle': list(values),
            'w_l1': list(values),
            'w_l2': list(values),
            'w_inf': list(values),
            'w_mag': list(values)
        }
