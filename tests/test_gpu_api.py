"""GPU tests of the reference-facing Python surface (same names / shapes as the reference)."""
import os

import numpy as np
import pytest

from oracle import metrics as OM
from oracle import search as OS
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "search_golden.npz"))


def test_scalar_metrics_surface(golden_dir):
    from image_retrieval_b200.geometric_metrics import GeometricSimilarityMetrics as G
    g = np.load(os.path.join(golden_dir, "metrics_golden.npz"))
    Q, X = g["Q_512"], g["X_512"]
    q, x = Q[1], X[6]
    ref = OM.get_all_metrics(q, x)
    got = G.get_all_metrics(q, x)
    assert list(got) == list(ref)
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=2e-5, atol=4e-6, err_msg=k)
    assert type(G.l2_distance(q, x)) is np.float64 and type(G.l1_distance(q, x)) is np.float32
    assert G.cosine_similarity(q, X[0]) == 0.0 and isinstance(G.cosine_similarity(q, X[0]), float)
    assert G.angular_distance(q, X[0]) == np.pi / 2
    np.testing.assert_allclose(G.l1_distance(q, x, normalized=False), g["l1_raw_512"][1, 6], rtol=2e-5)
    p = {"w_angle": 1.0, "w_l1": 1.0, "w_l2": 1.0, "w_inf": 0.0, "w_mag": 0.5}
    np.testing.assert_allclose(G.optimized_similarity(q, x, p), g["optimized_similarity_512"][1, 6], rtol=1e-4)
    np.testing.assert_allclose(G.optimized_distance(q, x, p), -g["optimized_similarity_512"][1, 6], rtol=1e-4)
    np.testing.assert_allclose(G.optimized_similarity(q, x, {}), g["optimized_default_512"][1, 6], rtol=1e-4, atol=4e-6)
    assert np.array_equal(np.array(G.create_parameter_grid(5)["w_l1"]), g["grid5"])


def test_search_images_matches_reference_run(gold):
    from image_retrieval_b200.app_pipeline import EnhancedImageSearchApp
    X, Q = gold["X"], gold["Q"]
    app = EnhancedImageSearchApp()
    assert app.search_images(Q[0]) == []                                   # empty store (app_pipeline.py:147-149)
    for i in range(len(X)):
        app.embeddings[f"img_{i:04d}.jpg"] = X[i]
    for qi, q in enumerate(Q):
        res = app.search_images(q, top_k=10)
        assert [int(r["path"][4:8]) for r in res] == list(gold[f"search_images_idx_{qi}"])
        np.testing.assert_allclose([r["score"] for r in res], gold[f"search_images_score_{qi}"], rtol=1e-5, atol=2e-6)
        multi = app.search_with_multiple_metrics(q, top_k=5)
        for name in ("cosine_similarity", "l1_distance", "l2_distance"):
            assert [int(r["path"][4:8]) for r in multi[name]] == list(gold[f"multi_{name}_idx_{qi}"]), name
            np.testing.assert_allclose([r[name] for r in multi[name]], gold[f"multi_{name}_val_{qi}"], rtol=2e-5, atol=2e-6)
        ref = OS.search_with_multiple_metrics({f"img_{i:04d}.jpg": X[i] for i in range(len(X))}, q, 5)
        assert multi["analysis"] == ref["analysis"]
    # optimized similarity branch
    app.searcher.set_similarity_params({"w_l1": 0.5, "w_mag": 0.25})
    res = app.search_images(Q[1], top_k=7, use_optimized_similarity=True)
    emb = {f"img_{i:04d}.jpg": X[i] for i in range(len(X))}
    ref = OS.search_images(emb, Q[1], 7, True, app.searcher.similarity_params)
    assert [r["path"] for r in res] == [r["path"] for r in ref]
    # fast path store
    app2 = EnhancedImageSearchApp()
    app2.set_embeddings([f"img_{i:04d}.jpg" for i in range(len(X))], X)
    assert [r["path"] for r in app2.search_images(Q[2], 10)] == [r["path"] for r in app.search_images(Q[2], 10)]
    assert app.search_with_multiple_metrics(Q[0], 5).keys() >= {"cosine_similarity", "l1_distance", "l2_distance", "analysis"}
    assert EnhancedImageSearchApp().search_with_multiple_metrics(Q[0]) == {'analysis': {'intersections': {}, 'unique_contributions': {}}}


def test_embedding_system_and_text_searcher(gold):
    from image_retrieval_b200.ImageEmbeddingSystem import ImageEmbeddingSystem
    from image_retrieval_b200.image_search import EnhancedTextImageSearcher
    from oracle import histogram as OH
    imgs = synth.images_palette(60, 40, 40, 77)
    sys_ = ImageEmbeddingSystem()
    assert sys_.process_and_store_images([]) == (0, 0)
    ok, failed = sys_.process_and_store_images([im for im in imgs[:50]] + ["/nonexistent/file.jpg"])
    assert (ok, failed) == (50, 1)
    sys_.store_arrays([f"p{i}" for i in range(50, 60)], imgs[50:])
    unit, mag = sys_.generate_embedding(imgs[3])
    ou, om = OH.embedding(imgs[3:4])
    np.testing.assert_allclose(unit, ou[0], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(mag, om[0], rtol=1e-6)
    with pytest.raises(Exception):
        sys_.generate_embedding("/nonexistent/file.jpg")
    embs = sys_.get_embeddings_with_magnitude(limit=1000)
    assert len(embs) == 60 and len(sys_.get_embeddings(limit=7)) == 7
    rec = sys_.reconstruct_original_embeddings(embs)
    np.testing.assert_allclose(rec[3][1], OH.histogram(imgs[3:4])[0].astype(np.float32), rtol=1e-5, atol=1e-3)

    searcher = EnhancedTextImageSearcher(collection=sys_.collection, text_encoder=lambda t: OH.embedding(imgs[5:6])[0][0])
    with pytest.raises(ValueError):
        searcher.search("   ")
    res = searcher.search("a query", top_k=5)
    assert len(res) <= 5 and all(r["score"] >= 0.25 for r in res)
    U = np.stack([e[1] for e in embs])
    paths = [e[0] for e in embs]
    q = OH.embedding(imgs[5:6])[0][0]
    # reference flow with an exact candidate stage (image_search.py:88-140)
    cos = OM.pairwise(q[None], U, "cosine_similarity", np.float64)[0]
    order = np.argsort(-cos, kind="stable")[:15]
    ref = OS.threshold_and_dedupe([{"path": paths[j], "score": cos[j]} for j in order], 5, 0.25, False)
    assert [r["path"] for r in res] == [r["path"] for r in ref]
    multi = searcher.search_with_multiple_metrics("a query", top_k=4)
    assert set(multi) == {"cosine_similarity", "l1_distance", "l2_distance", "linf_distance", "magnitude_difference",
                          "optimized_similarity", "analysis"}
    cand = np.argsort(-cos, kind="stable")[:20]
    for name, m, desc in (("l1_distance", "l1", False), ("linf_distance", "linf", False), ("cosine_similarity", "cosine_similarity", True)):
        vals = OM.pairwise(q[None], U[cand], m, np.float64)[0]
        o = np.argsort(-vals if desc else vals, kind="stable")[:4]
        assert [r["path"] for r in multi[name]] == [paths[cand[j]] for j in o], name
    cmp = searcher.compare_search_methods("a query", top_k=3)
    assert set(cmp) == {"standard_results", "optimized_results", "metrics"}


def test_allpairs_evaluator_surface():
    from image_retrieval_b200.mi_eval import AllPairsEvaluator, METRIC_NAMES, RELATIONSHIP_TYPES
    from oracle import evaluation as E
    N = 300
    X = synth.gaussian(N, 32, 3, normalize=True)
    cat, col = np.arange(N) % 5, (np.arange(N) // 5) % 3
    ev = AllPairsEvaluator(X, cat, col, nbins=128)
    h = ev.calculate_distances().cpu().numpy()
    assert h.shape == (5, 4, 128) and METRIC_NAMES == list(E.METRICS) and RELATIONSHIP_TYPES == list(E.RELATIONSHIP_TYPES)
    thr, prec, rec = ev.precision_recall("cosine_distance")
    vals = E.metric_matrices(X, np.float64)
    rel = E.relationship(cat, col)
    iu = np.triu_indices(N, 1)
    sel = rel[iu] <= 1
    ref = E.pr_curve_reference(list(vals["cosine_distance"][iu][sel]), list((rel[iu][sel] == 1).astype(int)), thr)
    tp, fp, fn = ref[:, 0], ref[:, 1], ref[:, 2]
    want_p = np.where(tp + fp > 0, tp / np.maximum(tp + fp, 1), 0.0)
    want_r = np.where(tp + fn > 0, tp / np.maximum(tp + fn, 1), 0.0)
    assert np.abs(prec - want_p).max() < 0.02 and np.abs(rec - want_r).max() < 0.02
    d = ev.densities()
    assert set(d) == set(METRIC_NAMES) and set(d["l1_distance"]) == set(RELATIONSHIP_TYPES)


def test_color_mi_analyzer_distances_match_reference_run(tmp_path, golden_dir):
    """ColorMIAnalyzer.load_dataset + calculate_distances (mi_analysis.py:199-297) on the reference's file formats ==
    the values the reference's get_all_metrics produced for the same pair lists (tests/golden/pairs_golden.json)."""
    import json
    import pandas as pd
    from image_retrieval_b200.mi_eval import ColorMIAnalyzer, save_pairs
    with open(os.path.join(golden_dir, "pairs_golden.json")) as f:
        g = json.load(f)
    base = tmp_path / "color_dataset"
    base.mkdir()
    pd.DataFrame([{**m, "path": str(base / m["path"])} for m in g["metadata"]]).to_csv(base / "metadata.csv", index=False)
    pairs = {r: [(str(base / a), str(base / b)) for a, b in lst] for r, lst in g["pairs"].items()}
    pairs["same_object_same_color"].insert(1, (str(base / "nowhere.jpg"), pairs["same_object_same_color"][0][0]))   # skipped
    save_pairs(base, pairs)
    np.savez(tmp_path / "e.npz", embeddings={str(base / p): np.asarray(v, np.float32) for p, v in g["embeddings"].items()})
    an = ColorMIAnalyzer(base_dir=str(base))
    assert an.load_dataset(str(tmp_path / "e.npz")) == (True, "Dataset loaded successfully")
    an.calculate_distances()
    assert list(an.distances) == an.metric_names
    for m in an.metric_names:
        for r in an.relationship_types:
            want = np.asarray(g["distances"][m][r], np.float64)
            got = np.asarray(an.distances[m][r], np.float64)
            assert got.shape == want.shape, (m, r)
            assert np.all(np.abs(got - want) <= 1e-5 * np.abs(want) + 2e-6), (m, r)
    thr, prec, rec = an.precision_recall("cosine_distance")
    from oracle import evaluation as E
    d = g["distances"]["cosine_distance"]
    ref = E.pr_curve_reference(d["same_object_diff_color"] + d["same_object_same_color"],
                               [1] * len(d["same_object_diff_color"]) + [0] * len(d["same_object_same_color"]), thr)
    tp, fp, fn = ref[:, 0], ref[:, 1], ref[:, 2]
    assert np.allclose(prec, np.where(tp + fp > 0, tp / np.maximum(tp + fp, 1), 0.0), atol=0.05)
    assert np.allclose(rec, np.where(tp + fn > 0, tp / np.maximum(tp + fn, 1), 0.0), atol=0.05)


def test_embedding_system_with_processor_front_end():
    """image_size=224: mixed-size images go through the PIL-exact resize + centre crop before the histogram
    (the reference's processor step, ImageEmbeddingSystem.py:82-83)."""
    from image_retrieval_b200.ImageEmbeddingSystem import ImageEmbeddingSystem
    from oracle import histogram as OH
    from oracle import resize as R
    imgs = [synth.images_palette(1, 120, 160, 5)[0], synth.images_palette(1, 300, 200, 6)[0], synth.images_uniform(1, 120, 160, 7)[0],
            synth.images_palette(1, 224, 224, 8)[0]]
    sys_ = ImageEmbeddingSystem(image_size=224)
    assert sys_.process_and_store_images(imgs) == (4, 0)
    assert len(sys_.get_embeddings_with_magnitude()) == 4
    for im in imgs:
        unit, mag = sys_.generate_embedding(im)
        ou, om = OH.embedding(R.clip_preprocess_u8(im, 224)[None])
        np.testing.assert_allclose(unit, ou[0], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(mag, om[0], rtol=1e-6)


def test_search_batch_equals_search_per_query():
    """EnhancedTextImageSearcher.search_batch (device threshold + dedupe) == search() query by query, with duplicate paths."""
    from image_retrieval_b200.image_search import EnhancedTextImageSearcher
    X = synth.gaussian(400, 64, 9, normalize=True)
    paths = [f"img{i % 150}.jpg" for i in range(400)]            # every path stored 2-3 times
    Q = X[:12] + 0.4 * synth.gaussian(12, 64, 10, normalize=True)
    searcher = EnhancedTextImageSearcher(collection=(paths, X))
    for opt in (False, True):
        if opt:
            searcher.set_similarity_params({"w_angle": 1.0, "w_l1": 0.5, "w_l2": 0.0, "w_inf": 0.0, "w_mag": 0.0})
        batch = searcher.search_batch(Q, top_k=6, score_threshold=0.3, use_optimized_similarity=opt)
        for q in range(12):
            single = searcher.search(Q[q], top_k=6, score_threshold=0.3, use_optimized_similarity=opt)
            assert [r["path"] for r in batch[q]] == [r["path"] for r in single], (opt, q)
            np.testing.assert_allclose([r["score"] for r in batch[q]], [r["score"] for r in single], rtol=1e-5, atol=2e-6)
            assert len({r["path"] for r in batch[q]}) == len(batch[q]) <= 6
