"""Pin the oracle (CPU restatement) against vectors produced by executing the reference
(tests/golden/make_golden.py) and against OpenCV for the histogram definition."""
import os

import numpy as np
import pytest

from oracle import histogram as H
from oracle import metrics as M
from oracle import search as S
from oracle import synth

DIMS = (1, 3, 7, 64, 512, 2048)
PARAMS = {"w_angle": 1.0, "w_l1": 1.0, "w_l2": 1.0, "w_inf": 0.0, "w_mag": 0.5}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics_golden.npz"))


@pytest.mark.parametrize("D", DIMS)
def test_scalar_restatement_bit_exact(gold, D):
    Q, X = gold[f"Q_{D}"], gold[f"X_{D}"]
    table = {"cosine_similarity": M.cosine_similarity, "cosine_distance": M.cosine_distance,
             "angular_distance": M.angular_distance, "l1_distance": M.l1_distance,
             "l2_distance": M.l2_distance, "linf_distance": M.linf_distance,
             "magnitude_difference": M.magnitude_difference}
    for name, fn in table.items():
        got = np.array([[fn(q, x) for x in X] for q in Q], dtype=np.float64)
        assert np.array_equal(got, gold[f"{name}_{D}"]), name
    got = np.array([[M.optimized_similarity(q, x, PARAMS) for x in X] for q in Q], dtype=np.float64)
    assert np.array_equal(got, gold[f"optimized_similarity_{D}"])
    got = np.array([[M.optimized_similarity(q, x, {}) for x in X] for q in Q], dtype=np.float64)
    assert np.array_equal(got, gold[f"optimized_default_{D}"])
    got = np.array([[M.l1_distance(q, x, normalized=False) for x in X] for q in Q], dtype=np.float64)
    assert np.array_equal(got, gold[f"l1_raw_{D}"])
    got = np.array([[M.l2_distance(q, x, normalized=False) for x in X] for q in Q], dtype=np.float64)
    assert np.array_equal(got, gold[f"l2_raw_{D}"])


def test_types_and_edge_values(gold):
    Q, X = gold["Q_512"], gold["X_512"]
    assert isinstance(M.cosine_similarity(Q[0], X[0]), float) and M.cosine_similarity(Q[0], X[0]) == 0.0
    assert M.angular_distance(Q[0], X[0]) == np.pi / 2
    assert type(M.l2_distance(Q[0], X[5])) is np.float64          # fp32 / sqrt(int) promotes
    assert type(M.l1_distance(Q[0], X[5])) is np.float32
    assert M.l1_distance(Q[0], X[1]) == 0 and M.linf_distance(Q[0], X[1]) == 0
    assert np.array_equal(np.array(M.create_parameter_grid(5)["w_l1"]), gold["grid5"])
    d = M.get_all_metrics(Q[1], X[6])
    assert list(d) == ["cosine_similarity", "cosine_distance", "angular_distance", "l1_distance",
                       "l2_distance", "linf_distance", "magnitude_difference"]


@pytest.mark.parametrize("D", DIMS)
def test_batched_matches_reference(gold, D):
    Q, X = gold[f"Q_{D}"], gold[f"X_{D}"]
    # fp32 batched: bit-exact where NumPy's row reduction == 1-D reduction
    assert np.array_equal(M.pairwise(Q, X, "l1", np.float32).astype(np.float64), gold[f"l1_distance_{D}"])
    assert np.array_equal(M.pairwise(Q, X, "linf", np.float32).astype(np.float64), gold[f"linf_distance_{D}"])
    # fp64 "truth" vs the reference's fp32 arithmetic: fp32 rounding only
    for name, key, atol in (("l1", "l1_distance", 0), ("l2", "l2_distance", 0), ("linf", "linf_distance", 0),
                            ("cosine_similarity", "cosine_similarity", 2e-6),
                            ("cosine_distance", "cosine_distance", 2e-6),
                            ("magnitude_difference", "magnitude_difference", 1e-5)):
        t = M.pairwise_f64(Q, X, name)
        np.testing.assert_allclose(gold[f"{key}_{D}"], t, rtol=2e-5, atol=atol + 1e-30, err_msg=name)
    t = M.pairwise_f64(Q, X, "angular_distance")
    np.testing.assert_allclose(gold[f"angular_distance_{D}"], t, rtol=1e-5, atol=1e-3)   # arccos near 0/pi
    t = M.pairwise_f64(Q, X, "optimized_similarity", params=PARAMS)
    np.testing.assert_allclose(gold[f"optimized_similarity_{D}"], t, rtol=1e-4, atol=1e-5)


def test_bf16_round():
    import torch
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    assert np.array_equal(M.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_search_semantics(golden_dir):
    g = np.load(os.path.join(golden_dir, "search_golden.npz"))
    X, Q = g["X"], g["Q"]
    emb = {f"img_{i:04d}.jpg": X[i] for i in range(len(X))}
    for qi, q in enumerate(Q):
        res = S.search_images(emb, q, top_k=10)
        assert [int(r["path"][4:8]) for r in res] == list(g[f"search_images_idx_{qi}"])
        assert np.array_equal(np.array([r["score"] for r in res], dtype=np.float64), g[f"search_images_score_{qi}"])
        multi = S.search_with_multiple_metrics(emb, q, top_k=5)
        for name in ("cosine_similarity", "l1_distance", "l2_distance"):
            assert [int(r["path"][4:8]) for r in multi[name]] == list(g[f"multi_{name}_idx_{qi}"])
        # batched stable top-k == list.sort + slice (ties -> lower index)
        for name, mname in (("l1_distance", "l1"), ("l2_distance", "l2"), ("linf_distance", "linf"),
                            ("cosine_similarity", "cosine_similarity")):
            v, i = S.topk_search(q[None], X, mname, 5, dtype=np.float32)
            assert list(i[0]) == list(g[f"multi_{name}_idx_{qi}"]), name
    # duplicates 5, 17, 200 of the scaled query 0 -> tie order 5, 17, 200
    assert list(g["multi_l1_distance_idx_0"][:0]) == []
    v, i = S.topk_search(Q[0][None], X, "cosine_similarity", 3)
    assert set(i[0]) == {5, 17, 200}
    assert S.search_images({}, Q[0]) == []


def test_topk_ties_and_merge():
    s = np.array([[1.0, 2.0, 2.0, 1.0]])
    assert list(S.topk(s, 4, True)[1][0]) == [1, 2, 0, 3]
    assert list(S.topk(s, 4, False)[1][0]) == [0, 3, 1, 2]
    assert S.topk(s, 10, False)[1].shape == (1, 4)
    rng = np.random.default_rng(3)
    sc = rng.integers(0, 5, size=(4, 64)).astype(np.float32)
    for desc in (False, True):
        fv, fi = S.topk(sc, 7, desc)
        parts_v, parts_i = [], []
        for r in range(4):
            v, i = S.topk(sc[:, r * 16:(r + 1) * 16], 7, desc)
            parts_v.append(v)
            parts_i.append(i + r * 16)
        mv, mi = S.merge_topk(np.stack(parts_v), np.stack(parts_i), 7, desc)
        assert np.array_equal(mi, fi) and np.array_equal(mv, fv)


def test_hsv_restatement_all_colours():
    import cv2
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8),
                          np.arange(0, 256, 1, dtype=np.uint8), indexing="ij")
    rgb = np.stack([r, g, b], -1).reshape(4096, 4096, 3)
    ref = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    assert np.array_equal(H.rgb_to_hsv_u8(rgb), ref)


def test_hsv_fp32_construction_all_colours():
    """The device kernel (csrc/histogram.cuh, hsv_histogram_kernel) does OpenCV's fixed-point conversion with fp32 operations on
    exact integers: 2^23-biased bytes, one exact fma for d*sdiv + 2^11 / hnum*hdiv + 2^11, floor by fma.rm against 2^23, the
    hue numerator as a penalised 3-input min, slots 6..14 folded onto 8 hue bins.  Restated here operation by operation
    in float64 (every intermediate is checked to be exactly representable in fp32, so the fp32 kernel computes the same
    numbers) and compared with the integer oracle on all 2^24 colours."""
    M = 2.0 ** 23
    cc = 2913.0 / 65536.0
    kh = M - M * cc
    assert kh == float(np.float32(kh)) and M * cc == 372864.0

    def exact32(x):
        assert np.array_equal(x.astype(np.float32).astype(np.float64), x)
        return x

    for r0 in range(0, 256, 16):
        c = np.arange(r0 << 16, (r0 + 16) << 16, dtype=np.uint32)
        r = ((c >> 16) & 255).astype(np.float64) + M
        g = ((c >> 8) & 255).astype(np.float64) + M
        b = (c & 255).astype(np.float64) + M
        v = np.maximum(np.maximum(r, g), b)
        d = v - np.minimum(np.minimum(r, g), b)
        sd = H.SDIV[(v - M).astype(np.int64)].astype(np.float64)
        hd = H.HDIV[d.astype(np.int64)].astype(np.float64)
        s1 = np.floor(exact32(d * sd + 2048.0) * 2.0 ** -17 + M)                   # fma.rn, then fma.rm
        v1 = np.floor(v * 2.0 ** -5 + (M - 2.0 ** 18))
        cb = exact32(4 * d + exact32(r - g))
        cg = exact32(2048.0 * exact32(v - g) + exact32(2 * d + exact32(b - r)))
        cr = exact32(2048.0 * exact32(v - r) + exact32(g - b))
        hn = np.minimum(np.minimum(cr, cg), cb)
        u1 = np.floor(exact32(hn * hd + 2048.0) * 2.0 ** -12 + (M + 180.0))
        h1 = np.floor(u1 * cc + kh)
        slot, sb, vb = (h1 - M).astype(np.int64), (s1 - M).astype(np.int64), (v1 - M).astype(np.int64)
        assert slot.min() >= 6 and slot.max() <= 14
        img = np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], axis=1).astype(np.uint8)
        assert np.array_equal((slot & 7) * 64 + sb * 8 + vb, H.bin_index(img, "hsv"))


def test_histogram_vs_opencv(golden_dir):
    g = np.load(os.path.join(golden_dir, "hist_golden.npz"))
    imgs = g["images"]
    assert np.array_equal(H.histogram(imgs, "rgb"), g["rgb"])
    assert np.array_equal(H.histogram(imgs, "hsv"), g["hsv"])
    for im in imgs[:3]:
        assert np.array_equal(H.histogram_cv2(im, "rgb").astype(np.uint32), H.histogram(im, "rgb")[0])
        assert np.array_equal(H.histogram_cv2(im, "hsv").astype(np.uint32), H.histogram(im, "hsv")[0])
    assert H.histogram(imgs, "rgb").sum(1).tolist() == [32 * 48] * len(imgs)
    # every H value 0..179 lands in the same bin as calcHist's float LUT
    import cv2
    hv = np.zeros((1, 180, 3), np.uint8)
    hv[0, :, 0] = np.arange(180)
    ref = cv2.calcHist([hv], [0], None, [8], [0, 180]).reshape(-1)
    mine = np.bincount(np.arange(180) * 8 // 180, minlength=8)
    assert np.array_equal(ref.astype(np.int64), mine)
    unit, mag = H.embedding(imgs[:2])
    np.testing.assert_allclose(np.linalg.norm(unit, axis=1), 1.0, rtol=1e-6)


def test_evaluation_oracle_matches_reference_pr_loop(gold):
    """oracle.evaluation: prefix sums of first-threshold-index counts == the reference's PR loop (mi_analysis.py:783-796)."""
    from oracle import evaluation as E
    rng = np.random.default_rng(5)
    X = rng.standard_normal((40, 16)).astype(np.float32)
    cat = np.arange(40) % 3
    col = (np.arange(40) // 3) % 2
    vals = E.metric_matrices(X, np.float32)
    rel = E.relationship(cat, col)
    thresholds = np.linspace(0, 1, 100)
    ranges = {m: (0.0, 4.0) for m in E.METRICS}
    hist, thr = E.bin_counts(vals, rel, ranges, 64, thresholds)
    iu = np.triu_indices(40, 1)
    assert hist.sum() == 5 * len(iu[0])
    r = rel[iu]
    for mi, m in enumerate(E.METRICS):
        d = vals[m][iu]
        sel = r <= 1
        ref = E.pr_curve_reference(list(d[sel]), list((r[sel] == 1).astype(int)), thresholds)
        assert np.array_equal(E.pr_from_counts(thr[mi]), ref), m
    # per-pair values are the reference's own get_all_metrics numbers
    i, j = 3, 17
    ref = M.get_all_metrics(X[i], X[j])
    for m in E.METRICS:
        np.testing.assert_allclose(vals[m][i, j], ref[m], rtol=2e-5, atol=2e-6)


def test_pair_list_distances_match_reference_run(golden_dir):
    """calculate_distances restatement (mi_analysis.py:256-297) vs values produced by the reference's get_all_metrics."""
    import json
    from oracle import evaluation as E
    with open(os.path.join(golden_dir, "pairs_golden.json")) as f:
        g = json.load(f)
    emb = {p: np.asarray(v, np.float32) for p, v in g["embeddings"].items()}
    pairs = {r: [tuple(p) for p in lst] for r, lst in g["pairs"].items()}
    pairs["same_object_same_color"].append(("not/in/store.jpg", g["metadata"][0]["path"]))   # skipped (:278-280)
    d = E.calculate_distances(emb, pairs)
    for m in E.METRICS:
        for r in E.RELATIONSHIP_TYPES:
            assert [float(x) for x in d[m][r]] == g["distances"][m][r], (m, r)


def _resize_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "resize_golden.npz"))
    for n, ((H, W, size, seed), gen) in enumerate(zip(g["cases"], g["generators"])):
        img = getattr(synth, f"images_{gen}")(1, int(H), int(W), int(seed))[0]
        yield img, int(size), g[f"out_{n}"]


def test_resize_oracle_matches_pil_golden(golden_dir):
    """Pillow's 8-bit bicubic resample restated in oracle/resize.py == PIL's own output (make_golden.resize_golden)."""
    from oracle import resize as R
    for img, size, want in _resize_cases(golden_dir):
        assert np.array_equal(R.clip_preprocess_u8(img, size), want), img.shape


def test_resize_oracle_matches_pil_live():
    Image = pytest.importorskip("PIL.Image")
    from oracle import resize as R
    for H, W, size, seed in [(97, 131, 48, 1), (131, 97, 48, 2), (500, 375, 224, 3), (40, 40, 64, 4)]:
        img = synth.images_palette(1, H, W, seed)[0]
        nh, nw = R.shortest_edge_size(H, W, size)
        pil = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
        assert np.array_equal(R.resize_bicubic(img, nh, nw), pil)
