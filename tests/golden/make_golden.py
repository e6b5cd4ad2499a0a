"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE in the build container.

Run here (where /root/reference exists):  python tests/golden/make_golden.py
The reference's geometric_metrics.py (NumPy only) imports cleanly; app_pipeline.py
and image_search.py do not (matplotlib / pymilvus absent), so their search loops
are exercised by feeding the reference's own metric functions through the exact
statements of app_pipeline.py:156-172 / :296-328 (list of dicts, list.sort, slice).

Outputs (committed, small):
  metrics_golden.npz  - vectors + every reference metric, pair by pair
  search_golden.npz   - DB/query vectors + reference top-k paths/scores
  hist_golden.npz     - images + cv2.calcHist counts (RGB, HSV)  [OpenCV, not the reference]
  resize_golden.npz   - PIL Image.resize(BICUBIC) + centre crop of seeded images (the CLIPProcessor front-end)  [Pillow]
  pairs_golden.json   - metadata rows -> the reference's generate_relationship_pairs output, and the per-pair
                        get_all_metrics values of mi_analysis.py:277-291 for seeded embeddings of those paths
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
from geometric_metrics import GeometricSimilarityMetrics as G  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["cosine_similarity", "cosine_distance", "angular_distance", "l1_distance",
         "l2_distance", "linf_distance", "magnitude_difference"]


def metrics_golden():
    rng = np.random.default_rng(20261018)
    cases = {}
    for D in (1, 3, 7, 64, 512, 2048):
        Q = rng.standard_normal((5, D)).astype(np.float32)
        X = rng.standard_normal((9, D)).astype(np.float32)
        X[0] = 0.0                         # zero vector: cos -> 0.0, angle -> pi/2
        X[1] = Q[0]                        # exact duplicate: distance 0
        X[2] = Q[1] + 1e-3 * X[2]          # near duplicate (cancellation case for the GEMM form)
        X[3] = -Q[2]                       # antipodal: cos -1, angle pi
        X[4] = np.maximum(X[4], 0)
        cases[D] = (Q, X)
    out = {}
    params = {"w_angle": 1.0, "w_l1": 1.0, "w_l2": 1.0, "w_inf": 0.0, "w_mag": 0.5}   # results.json:16-22
    for D, (Q, X) in cases.items():
        out[f"Q_{D}"] = Q
        out[f"X_{D}"] = X
        for n in NAMES:
            out[f"{n}_{D}"] = np.array([[getattr(G, n)(q, x) for x in X] for q in Q], dtype=np.float64)
        out[f"l1_raw_{D}"] = np.array([[G.l1_distance(q, x, normalized=False) for x in X] for q in Q], dtype=np.float64)
        out[f"l2_raw_{D}"] = np.array([[G.l2_distance(q, x, normalized=False) for x in X] for q in Q], dtype=np.float64)
        out[f"optimized_similarity_{D}"] = np.array(
            [[G.optimized_similarity(q, x, params) for x in X] for q in Q], dtype=np.float64)
        out[f"optimized_default_{D}"] = np.array(
            [[G.optimized_similarity(q, x, {}) for x in X] for q in Q], dtype=np.float64)
    out["grid5"] = np.array(G.create_parameter_grid(5)["w_l1"])
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), **out)


def search_golden():
    rng = np.random.default_rng(7)
    N, D, nq = 300, 64, 6
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[17] = X[5]
    X[200] = X[5]          # exact ties -> lower index first
    X[40] = -X[41]
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    Q[0] = X[5] * 2.0
    embeddings = {f"img_{i:04d}.jpg": X[i] for i in range(N)}
    out = {"X": X, "Q": Q}
    for qi, q in enumerate(Q):
        # app_pipeline.py:156-172 (plain cosine branch)
        results = []
        for path, e in embeddings.items():
            sim = np.dot(q, e) / (np.linalg.norm(q) * np.linalg.norm(e))
            results.append({"path": path, "score": abs(sim)})
        results.sort(key=lambda x: x["score"], reverse=True)
        top = results[:10]
        out[f"search_images_idx_{qi}"] = np.array([int(r["path"][4:8]) for r in top])
        out[f"search_images_score_{qi}"] = np.array([r["score"] for r in top], dtype=np.float64)
        # app_pipeline.py:296-328
        for name, fn, sign in (("cosine_similarity", G.cosine_similarity, 1.0),
                               ("l1_distance", G.l1_distance, -1.0), ("l2_distance", G.l2_distance, -1.0)):
            rows = []
            for path, e in embeddings.items():
                v = fn(q, e)
                rows.append({"path": path, name: v, "score": sign * v})
            rows.sort(key=lambda x: x["score"], reverse=True)
            out[f"multi_{name}_idx_{qi}"] = np.array([int(r["path"][4:8]) for r in rows[:5]])
            out[f"multi_{name}_val_{qi}"] = np.array([r[name] for r in rows[:5]], dtype=np.float64)
        # image_search.py:199-219 orderings for the remaining metrics, over all candidates
        for name, fn, rev in (("linf_distance", G.linf_distance, False),
                              ("magnitude_difference", G.magnitude_difference, False),
                              ("angular_distance", G.angular_distance, False)):
            cands = [{"i": i, name: fn(q, X[i])} for i in range(N)]
            cands = sorted(cands, key=lambda x: x[name], reverse=rev)[:5]
            out[f"multi_{name}_idx_{qi}"] = np.array([c["i"] for c in cands])
            out[f"multi_{name}_val_{qi}"] = np.array([c[name] for c in cands], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "search_golden.npz"), **out)


def hist_golden():
    import cv2
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import synth
    imgs = np.concatenate([synth.images_uniform(2, 32, 48, 11), synth.images_palette(3, 32, 48, 12)])
    flat = np.zeros((1, 32, 48, 3), np.uint8)
    flat[...] = (255, 0, 128)
    imgs = np.concatenate([imgs, flat])
    rgb, hsv = [], []
    for im in imgs:
        im = np.ascontiguousarray(im)
        rgb.append(cv2.calcHist([im], [0, 1, 2], None, [8, 8, 8], [0, 256] * 3).reshape(-1))
        hv = cv2.cvtColor(im, cv2.COLOR_RGB2HSV)
        hsv.append(cv2.calcHist([hv], [0, 1, 2], None, [8, 8, 8], [0, 180, 0, 256, 0, 256]).reshape(-1))
    np.savez_compressed(os.path.join(HERE, "hist_golden.npz"), images=imgs,
                        rgb=np.array(rgb).astype(np.uint32), hsv=np.array(hsv).astype(np.uint32))


RESIZE_CASES = [  # (H, W, size, seed, generator)
    (60, 80, 32, 21, "uniform"), (80, 60, 32, 22, "palette"), (33, 47, 32, 23, "uniform"), (32, 32, 32, 24, "uniform"),
    (20, 31, 32, 25, "palette"), (300, 32, 32, 26, "uniform"), (32, 301, 32, 27, "uniform"), (1000, 1300, 32, 28, "palette"),
    (480, 640, 224, 29, "palette")]


def resize_golden():
    """PIL itself: Image.fromarray(img).resize((new_w, new_h), BICUBIC), then the processor's centre crop.  The result
    equals transformers' CLIPImageProcessorPil(do_rescale=False, do_normalize=False) (checked when available)."""
    from PIL import Image
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import synth
    out = {"cases": np.array([c[:4] for c in RESIZE_CASES], dtype=np.int64), "generators": np.array([c[4] for c in RESIZE_CASES])}
    for n, (H, W, size, seed, gen) in enumerate(RESIZE_CASES):
        img = getattr(synth, f"images_{gen}")(1, H, W, seed)[0]
        short, long = (W, H) if W <= H else (H, W)
        new_long = int(size * long / short)
        nh, nw = (new_long, size) if W <= H else (size, new_long)
        r = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
        top, left = (nh - size) // 2, (nw - size) // 2
        crop = r[top:top + size, left:left + size]
        try:
            from transformers import CLIPImageProcessorPil
            p = CLIPImageProcessorPil(do_rescale=False, do_normalize=False, do_convert_rgb=False, size={"shortest_edge": size},
                                      crop_size={"height": size, "width": size})
            t = np.moveaxis(np.asarray(p(images=Image.fromarray(img), return_tensors="np")["pixel_values"][0]), 0, -1)
            assert np.array_equal(t, crop), (H, W, size)
        except ImportError:
            pass
        out[f"out_{n}"] = crop
    np.savez_compressed(os.path.join(HERE, "resize_golden.npz"), **out)


def pairs_golden():
    """ColorDatasetManager.generate_relationship_pairs (imageProcessing.py:296-387) run on hand-made metadata, then the
    calculate_distances loop body (mi_analysis.py:277-291; mi_analysis itself needs matplotlib and cannot be imported)."""
    import json
    import tempfile
    from imageProcessing import ColorDatasetManager
    rng = np.random.default_rng(77)
    with tempfile.TemporaryDirectory() as tmp:
        mgr = ColorDatasetManager(base_dir=os.path.join(tmp, "ds"))
        base = str(mgr.base_dir)
        meta = []
        for category in ("dog", "car", "boat", "chair"):
            for color in ("brown", "white", "black"):
                if (category, color) in (("boat", "black"), ("chair", "brown"), ("chair", "white")):
                    continue                                    # ragged: some cells empty, chair has one colour
                for i in range(int(rng.integers(1, 4))):
                    meta.append({"path": os.path.join(base, category, color, f"{i + 1}.jpg"), "category": category, "color": color})
        mgr.metadata = meta
        pairs = mgr.generate_relationship_pairs()
        rel = lambda p: p[len(base) + 1:]
        paths = [m["path"] for m in meta]
        emb = {p: rng.standard_normal(64).astype(np.float32) for p in paths}
        emb[paths[3]] = np.zeros(64, np.float32)
        dist = {}
        for rel_type, lst in pairs.items():
            for n in NAMES:
                dist.setdefault(n, {})[rel_type] = [float(G.get_all_metrics(emb[a], emb[b])[n]) for a, b in lst]
        out = {"metadata": [{"path": rel(m["path"]), "category": m["category"], "color": m["color"]} for m in meta],
               "categories": mgr.categories, "colors": mgr.colors,
               "pairs": {r: [[rel(a), rel(b)] for a, b in lst] for r, lst in pairs.items()},
               "embeddings": {rel(p): [float(x) for x in v] for p, v in emb.items()},
               "distances": dist}
    with open(os.path.join(HERE, "pairs_golden.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    pairs_golden()
    resize_golden()
    if "--new-only" in sys.argv:
        sys.exit(0)
    metrics_golden()
    search_golden()
    hist_golden()
    print("golden vectors written to", HERE)
