"""Oracle (test infrastructure): metric arithmetic of the reference, on the CPU.

Scalar functions follow /root/reference/src/geometric_metrics.py line by line
(citations on each function).  They are written against NumPy only, like the
reference, so that fp32 inputs give the same fp32 (or, for L2, fp64) outputs.

Batched functions (`pairwise`, `pairwise_f64`) are the vectorised restatement
used at sizes where a Python pair loop is too slow; tests/test_oracle.py checks
them against the scalar functions (bit-exact for L1/L2/Linf, a few ulps for the
BLAS-backed dot products).
"""
import numpy as np

METRICS = ("l1", "l2", "linf", "cosine_similarity", "cosine_distance", "angular_distance")
# direction of "best": distances ascend, similarity descends
DESCENDING = {"l1": False, "l2": False, "linf": False, "cosine_similarity": True,
              "cosine_distance": False, "angular_distance": False,
              "magnitude_difference": False, "optimized_similarity": True, "abs_cosine": True}


# ----------------------------------------------------------------------------- scalar
def cosine_similarity(vec1, vec2):
    """geometric_metrics.py:12-18 - dot/(|a||b|); python 0.0 when a norm is 0."""
    norm1 = np.linalg.norm(vec1)
    norm2 = np.linalg.norm(vec2)
    if norm1 == 0 or norm2 == 0:
        return 0.0
    return np.dot(vec1, vec2) / (norm1 * norm2)


def angular_distance(vec1, vec2):
    """geometric_metrics.py:21-26 - arccos(clip(cos, -1, 1)) in radians."""
    return np.arccos(np.clip(cosine_similarity(vec1, vec2), -1.0, 1.0))


def cosine_distance(vec1, vec2):
    """geometric_metrics.py:29-31 - 1 - cos."""
    return 1.0 - cosine_similarity(vec1, vec2)


def l1_distance(vec1, vec2, normalized=True):
    """geometric_metrics.py:34-39 - sum|a-b|, divided by D when normalized."""
    distance = np.sum(np.abs(vec1 - vec2))
    if normalized:
        distance /= len(vec1)
    return distance


def l2_distance(vec1, vec2, normalized=True):
    """geometric_metrics.py:42-47 - sqrt(sum (a-b)^2), / sqrt(D) when normalized
    (the divide promotes an fp32 result to fp64, as in the reference)."""
    distance = np.sqrt(np.sum((vec1 - vec2) ** 2))
    if normalized:
        distance /= np.sqrt(len(vec1))
    return distance


def linf_distance(vec1, vec2):
    """geometric_metrics.py:50-52 - max|a-b|, never normalised."""
    return np.max(np.abs(vec1 - vec2))


def magnitude_difference(vec1, vec2):
    """geometric_metrics.py:55-57 - | |a| - |b| |."""
    return abs(np.linalg.norm(vec1) - np.linalg.norm(vec2))


def optimized_similarity(vec1, vec2, params):
    """geometric_metrics.py:60-94 - w_angle*cos - w_l1*L1n - w_l2*L2n - w_inf*Linf - w_mag*mag."""
    w_angle = params.get("w_angle", 1.0)
    w_l1 = params.get("w_l1", 0.0)
    w_l2 = params.get("w_l2", 0.0)
    w_inf = params.get("w_inf", 0.0)
    w_mag = params.get("w_mag", 0.0)
    return (w_angle * cosine_similarity(vec1, vec2)
            - w_l1 * l1_distance(vec1, vec2)
            - w_l2 * l2_distance(vec1, vec2)
            - w_inf * linf_distance(vec1, vec2)
            - w_mag * magnitude_difference(vec1, vec2))


def optimized_distance(vec1, vec2, params):
    """geometric_metrics.py:97-111 - negated optimized_similarity."""
    return -optimized_similarity(vec1, vec2, params)


def get_all_metrics(vec1, vec2):
    """geometric_metrics.py:114-129 - the seven-key dict."""
    return {
        "cosine_similarity": cosine_similarity(vec1, vec2),
        "cosine_distance": cosine_distance(vec1, vec2),
        "angular_distance": angular_distance(vec1, vec2),
        "l1_distance": l1_distance(vec1, vec2),
        "l2_distance": l2_distance(vec1, vec2),
        "linf_distance": linf_distance(vec1, vec2),
        "magnitude_difference": magnitude_difference(vec1, vec2),
    }


def create_parameter_grid(granularity=5):
    """geometric_metrics.py:132-149."""
    values = np.linspace(0.0, 1.0, granularity)
    return {k: list(values) for k in ("w_angle", "w_l1", "w_l2", "w_inf", "w_mag")}


SCALAR = {
    "l1": l1_distance, "l2": l2_distance, "linf": linf_distance,
    "cosine_similarity": cosine_similarity, "cosine_distance": cosine_distance,
    "angular_distance": angular_distance, "magnitude_difference": magnitude_difference,
}


def pair_loop(Q, X, metric, **kw):
    """Reference-faithful pair loop (app_pipeline.py:156-168 shape): out[i, j] = metric(Q[i], X[j])."""
    f = SCALAR[metric]
    out = np.empty((len(Q), len(X)), dtype=np.float64)
    for i, q in enumerate(Q):
        for j, x in enumerate(X):
            out[i, j] = f(q, x, **kw)
    return out


# ----------------------------------------------------------------------------- batched
def _chunks(n, step):
    for s in range(0, n, step):
        yield s, min(n, s + step)


def pairwise(Q, X, metric, dtype=np.float64, normalized=True, params=None, chunk=2048):
    """Vectorised restatement: (nq, N) matrix of `metric`, arithmetic carried out in `dtype`.

    dtype=np.float64 is the "truth" the GPU results are compared with (tolerance stated in
    the tests); dtype=np.float32 reproduces the reference's own fp32 arithmetic up to the
    summation order of the BLAS-backed dot products.
    """
    Q = np.ascontiguousarray(Q, dtype=dtype)
    X = np.ascontiguousarray(X, dtype=dtype)
    nq, D = Q.shape
    N = X.shape[0]
    out = np.empty((nq, N), dtype=dtype)
    if metric in ("cosine_similarity", "cosine_distance", "angular_distance", "abs_cosine",
                  "magnitude_difference", "optimized_similarity"):
        qn = np.sqrt(np.einsum("ij,ij->i", Q, Q))
        xn = np.sqrt(np.einsum("ij,ij->i", X, X))
    if metric == "optimized_similarity":
        p = params or {}
        cos = pairwise(Q, X, "cosine_similarity", dtype, chunk=chunk)
        out[:] = p.get("w_angle", 1.0) * cos
        for key, name in (("w_l1", "l1"), ("w_l2", "l2"), ("w_inf", "linf"), ("w_mag", "magnitude_difference")):
            w = p.get(key, 0.0)
            if w != 0.0:
                out -= w * pairwise(Q, X, name, dtype, chunk=chunk)
        return out
    if metric == "magnitude_difference":
        return np.abs(qn[:, None] - xn[None, :]).astype(dtype)
    for s, e in _chunks(N, chunk):
        Xc = X[s:e]
        if metric in ("l1", "l2", "linf"):
            for qs, qe in _chunks(nq, 16):
                diff = Q[qs:qe, None, :] - Xc[None, :, :]
                if metric == "l1":
                    d = np.abs(diff).sum(-1)
                    if normalized:
                        d = d / dtype(D)
                elif metric == "l2":
                    d = np.sqrt((diff ** 2).sum(-1))
                    if normalized:
                        d = d / np.sqrt(dtype(D))
                else:
                    d = np.abs(diff).max(-1)
                out[qs:qe, s:e] = d
        else:
            dot = Q @ Xc.T
            den = qn[:, None] * xn[None, s:e]
            with np.errstate(divide="ignore", invalid="ignore"):
                cos = np.where(den == 0, dtype(0), dot / den)   # zero norm -> 0.0 (geometric_metrics.py:16-17)
            if metric == "cosine_similarity":
                out[:, s:e] = cos
            elif metric == "abs_cosine":
                out[:, s:e] = np.abs(cos)                       # app_pipeline.py:167
            elif metric == "cosine_distance":
                out[:, s:e] = dtype(1.0) - cos
            else:
                out[:, s:e] = np.arccos(np.clip(cos, -1.0, 1.0))
    return out


def pairwise_f64(Q, X, metric, **kw):
    return pairwise(np.asarray(Q, dtype=np.float64), np.asarray(X, dtype=np.float64), metric, np.float64, **kw)


def bf16_round(a):
    """Round fp32 -> bf16 (round-to-nearest-even) -> fp32, the value a bf16 database row holds."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)
