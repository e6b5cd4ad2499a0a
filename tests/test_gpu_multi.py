"""GPU tests of the multi-metric single pass (b200ir_topk_multi), candidate re-ranking (b200ir_candidate_metrics +
b200ir_rank_candidates) and result pages beyond one 256-row page (b200ir_topk_paged + b200ir_sort_topk_rows).
Reference semantics: app_pipeline.py:296-328 (three scans, three stable sorts), image_search.py:98-115 and :173-219
(candidate scoring, six orderings), app_pipeline.py:171-172 (results[:top_k] for any top_k).  Run with -m gpu."""
import numpy as np
import pytest

from oracle import metrics as OM
from oracle import search as OS
from oracle import synth
from parity import check_topk

pytestmark = pytest.mark.gpu
PARAMS = {"w_angle": 1.0, "w_l1": 1.0, "w_l2": 1.0, "w_inf": 0.25, "w_mag": 0.5}
ALL = ["cosine_similarity", "cosine_distance", "angular_distance", "l1", "l2", "linf", "magnitude_difference", "optimized_similarity"]


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from image_retrieval_b200 import ops as o
    o.device()
    return o


@pytest.mark.parametrize("nq,N,D,k", [(1, 5000, 512, 5), (3, 20000, 64, 10), (9, 3001, 100, 100), (13, 1500, 33, 256)])
@pytest.mark.parametrize("bf16", [False, True])
def test_topk_multi_equals_one_scan_per_metric(ops, nq, N, D, k, bf16):
    """Every plane of the one-pass result is bit-identical to that metric's own fused scan (same arithmetic, same
    accumulators), so parity with the oracle carries over from the single-metric tests."""
    import torch
    from image_retrieval_b200 import _lib
    Q = synth.gaussian(nq, D, 71)
    X = synth.gaussian(N, D, 72)
    X[5] = 0
    X[6] = Q[0]
    Qd, Xd = torch.from_numpy(Q).cuda(), torch.from_numpy(X).cuda()
    if bf16:
        Qd, Xd = Qd.bfloat16(), Xd.bfloat16()
    lib = _lib.load()
    n0 = lib.b200ir_launch_count()
    S, I = ops.topk_multi(Qd, Xd, ALL, k, params=PARAMS)
    assert lib.b200ir_launch_count() - n0 == 3, "query prep + ONE scan + one merge launch"
    for y, m in enumerate(ALL):
        kw = {"params": PARAMS} if m == "optimized_similarity" else {}
        s, i = ops.topk(Qd, Xd, m, k, flags=ops.FLAG_NO_TENSOR, **kw)
        assert torch.equal(I[y], i), m
        assert torch.equal(S[y], s), m
    if not bf16:
        truth = OM.pairwise_f64(Q, X, "l1")
        check_topk(S[3].cpu().numpy(), I[3].cpu().numpy(), truth, k, False, rtol=1e-5, atol=1e-30)


def test_app_multi_metrics_is_one_scan(ops):
    import torch
    from image_retrieval_b200 import _lib
    from image_retrieval_b200.app_pipeline import EnhancedImageSearchApp
    X = synth.gaussian(4000, 512, 81)
    q = synth.gaussian(1, 512, 82)[0]
    app = EnhancedImageSearchApp()
    app.set_embeddings([f"img_{i:05d}.jpg" for i in range(len(X))], X)
    app.search_with_multiple_metrics(q, 5)
    lib = _lib.load()
    n0 = lib.b200ir_launch_count()
    got = app.search_with_multiple_metrics(q, 5)
    assert lib.b200ir_launch_count() - n0 == 3                       # prep + ONE scan + merge (the reference: 3 scans + 3 sorts)
    ref = OS.search_with_multiple_metrics({f"img_{i:05d}.jpg": X[i] for i in range(len(X))}, q, 5)
    for name in ("cosine_similarity", "l1_distance", "l2_distance"):
        assert [r["path"] for r in got[name]] == [r["path"] for r in ref[name]], name
        np.testing.assert_allclose([r[name] for r in got[name]], [r[name] for r in ref[name]], rtol=2e-5, atol=2e-6)
    assert got["analysis"] == ref["analysis"]
    torch.cuda.synchronize()


@pytest.mark.parametrize("metric", ["l1", "cosine_similarity", "l2", "angular_distance", "linf"])
@pytest.mark.parametrize("nq,N,D,k", [(3, 5000, 64, 700), (2, 900, 128, 900), (1, 3000, 512, 257)])
def test_paged_topk_beyond_one_page(ops, metric, nq, N, D, k):
    Q = synth.gaussian(nq, D, 91)
    X = synth.gaussian(N, D, 92)
    s, i = ops.topk(Q, X, metric, k)
    truth = OM.pairwise_f64(Q, X, metric)
    tol = dict(rtol=1e-5, atol=2e-6) if metric == "cosine_similarity" else (dict(rtol=1e-5, atol=1e-5 * np.pi) if metric == "angular_distance"
                                                                            else dict(rtol=1e-5, atol=1e-30))
    disputed = check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, k, OM.DESCENDING[metric], **tol)
    assert disputed <= max(2, nq * k // 100), disputed


def test_paged_topk_integer_ties_exact(ops):
    """Small-integer vectors: distances are exact, ties abound (also across page boundaries) -> the paged list equals
    the stable-sort oracle exactly, and pages beyond the store are padded."""
    rng = np.random.default_rng(5)
    X = rng.integers(0, 3, size=(1000, 16)).astype(np.float32)
    Q = rng.integers(0, 3, size=(4, 16)).astype(np.float32)
    for metric in ("l1", "l2", "linf"):
        s, i = ops.topk(Q, X, metric, 1200)
        tv, ti = OS.topk(OM.pairwise_f64(Q, X, metric), 1200, False)
        assert np.array_equal(i.cpu().numpy()[:, :1000], ti), metric
        np.testing.assert_allclose(s.cpu().numpy()[:, :1000], tv, rtol=1e-6)
        assert (i.cpu().numpy()[:, 1000:] == -1).all() and np.isinf(s.cpu().numpy()[:, 1000:]).all()


def test_search_images_top_k_beyond_256(ops):
    from image_retrieval_b200.app_pipeline import EnhancedImageSearchApp
    X = synth.gaussian(1500, 64, 101)
    q = synth.gaussian(1, 64, 102)[0]
    app = EnhancedImageSearchApp()
    app.set_embeddings([f"img_{i:05d}.jpg" for i in range(len(X))], X)
    emb = {f"img_{i:05d}.jpg": X[i] for i in range(len(X))}
    got = app.search_images(q, top_k=600)
    ref = OS.search_images(emb, q, 600)
    assert len(got) == 600
    same = sum(a["path"] == b["path"] for a, b in zip(got, ref))
    assert same >= 597                                              # fp32 vs fp64 near-ties may swap neighbours
    np.testing.assert_allclose([r["score"] for r in got], [r["score"] for r in ref], rtol=1e-5, atol=2e-6)
    assert len(app.search_images(q, top_k=5000)) == 1500            # more than the store holds: everything, like results[:top_k]
    with pytest.raises(ValueError):
        big = EnhancedImageSearchApp()
        big.set_embeddings([str(i) for i in range(5000)], synth.gaussian(5000, 8, 1))
        big.search_images(synth.gaussian(1, 8, 2)[0], top_k=4500)


def test_rank_candidates_vs_oracle(ops):
    import torch
    nq, N, D, kc, k = 5, 2000, 96, 40, 7
    Q = synth.gaussian(nq, D, 111)
    X = synth.gaussian(N, D, 112)
    rng = np.random.default_rng(3)
    cand = np.stack([rng.choice(N, kc, replace=False) for _ in range(nq)]).astype(np.int64)
    cand[1, 30:] = -1                                               # ragged list: padding at the end
    cand[2, 5] = cand[2, 4]                                         # the same row twice: exact tie, candidate order decides
    r = ops.rank_candidates(Q, X, torch.from_numpy(cand).cuda(), k, params=PARAMS)
    vals = r["metrics"].cpu().numpy()
    opt = r["optimized"].cpu().numpy()
    pos = r["pos"].cpu().numpy()
    names = ["cosine_similarity", "cosine_distance", "angular_distance", "l1", "l2", "linf", "magnitude_difference"]
    for q in range(nq):
        valid = cand[q] >= 0
        Xc = X[cand[q][valid]]
        table = {}
        for y, m in enumerate(names):
            truth = OM.pairwise_f64(Q[q:q + 1], Xc, m)[0]
            np.testing.assert_allclose(vals[y, q][valid], truth, rtol=2e-5, atol=1e-5 if m == "angular_distance" else 4e-6, err_msg=m)
            assert np.isnan(vals[y, q][~valid]).all()
            table[m] = truth
        t_opt = OM.pairwise_f64(Q[q:q + 1], Xc, "optimized_similarity", params=PARAMS)[0]
        np.testing.assert_allclose(opt[q][valid], t_opt, rtol=1e-4, atol=2e-5)
        orders = (("cosine_similarity", True), ("l1", False), ("l2", False), ("linf", False), ("magnitude_difference", False))
        for y, (m, desc) in enumerate(orders):
            got = pos[y, q]
            # stable order of the GPU's own fp32 values (ties keep candidate order), then consistency with the fp64 truth
            v = vals[names.index(m), q][valid]
            want = np.argsort(-v if desc else v, kind="stable")[:k]
            assert np.array_equal(got[:len(want)], want), (m, q)
        want = np.argsort(-opt[q][valid], kind="stable")[:k]
        assert np.array_equal(pos[5, q][:len(want)], want)
        assert np.array_equal(r["row"].cpu().numpy()[5, q][:len(want)], cand[q][valid][want])


def test_get_all_metrics_is_one_launch(ops):
    from image_retrieval_b200 import _lib
    from image_retrieval_b200.geometric_metrics import GeometricSimilarityMetrics as G
    a, b = synth.gaussian(2, 512, 121)
    lib = _lib.load()
    G.get_all_metrics(a, b)
    n0 = lib.b200ir_launch_count()
    got = G.get_all_metrics(a, b)
    assert lib.b200ir_launch_count() - n0 == 1
    ref = OM.get_all_metrics(a, b)
    for key in ref:
        np.testing.assert_allclose(got[key], ref[key], rtol=2e-5, atol=4e-6, err_msg=key)
    n0 = lib.b200ir_launch_count()
    G.cosine_similarity(a, b)
    assert lib.b200ir_launch_count() - n0 == 1


def test_tracked_dict_ior_and_invalidate(ops):
    from image_retrieval_b200.app_pipeline import EnhancedImageSearchApp
    X = synth.gaussian(50, 32, 131)
    app = EnhancedImageSearchApp()
    for i in range(40):
        app.embeddings[f"a{i}"] = X[i]
    q = X[45]
    assert app.search_images(q, 1)[0]["path"] != "a45"
    app.embeddings |= {f"a{i}": X[i] for i in range(40, 50)}        # dict.__ior__ must invalidate the device copy
    assert app.search_images(q, 1)[0]["path"] == "a45"
    app.embeddings["a3"][:] = X[45] * 2                             # in-place row edit needs an explicit invalidate()
    app.embeddings.invalidate()
    assert app.search_images(q, 2)[0]["path"] == "a3"


def test_path_groups_follow_collection_changes(ops):
    from image_retrieval_b200.image_search import EnhancedTextImageSearcher
    from image_retrieval_b200.store import EmbeddingStore
    X = synth.gaussian(30, 16, 141, normalize=True)
    st = EmbeddingStore()
    for i in range(30):
        st.add(f"p{i}", X[i])
    s = EnhancedTextImageSearcher(collection=st)
    assert s._path_groups(st.paths) is None                         # all paths distinct
    st.set_path(7, "p3")                                            # same length, one duplicate path now
    g = s._path_groups(st.paths)
    assert g is not None and g[7].item() == 3
    res = s.search_batch(X[3:4] + X[7:8], top_k=30, score_threshold=-1.0)[0]
    assert [r["path"] for r in res].count("p3") == 1
    s.collection = ([f"q{i}" for i in range(30)], X)                # tuple collection of the same length: no stale groups
    assert s._path_groups(s._store()[0]) is None


def test_allpairs_eval_parts_add_up(ops):
    """The cyclic row shares of b200ir_allpairs_eval_part (what each GPU of a replicated store counts) sum to the
    unsharded evaluation exactly - integer counts, so the all-reduce across GPUs is exact too."""
    import torch
    N = 1000
    X = synth.gaussian(N, 64, 151)
    cat, col = np.arange(N) % 7, (np.arange(N) // 7) % 3
    ranges = {"cosine_distance": (0.0, 2.0), "l1_distance": (0.0, 2.5), "l2_distance": (0.0, 3.0), "linf_distance": (0.0, 8.0),
              "magnitude_difference": (0.0, 5.0)}
    h, t = ops.allpairs_eval(X, cat, col, ranges, 128)
    for nparts in (2, 3, 8):
        hs, ts = torch.zeros_like(h), torch.zeros_like(t)
        for part in range(nparts):
            hp, tp = ops.allpairs_eval(X, cat, col, ranges, 128, part=part, nparts=nparts)
            hs += hp
            ts += tp
        assert torch.equal(hs, h) and torch.equal(ts, t), nparts
    assert int(h[0].sum()) == N * (N - 1) // 2
