"""GPU parity tests: the CUDA path (through the C ABI) against the oracle.  Run with -m gpu."""
import os

import numpy as np
import pytest

from oracle import histogram as OH
from oracle import metrics as OM
from oracle import search as OS
from oracle import synth
from parity import check_topk

pytestmark = pytest.mark.gpu

METRICS = ["l1", "l2", "linf", "cosine_similarity", "cosine_distance", "angular_distance", "magnitude_difference"]
PARAMS = {"w_angle": 1.0, "w_l1": 1.0, "w_l2": 1.0, "w_inf": 0.0, "w_mag": 0.5}


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from image_retrieval_b200 import ops as o
    o.device()
    return o


def _tol(metric):
    # fp32: 1e-5 relative (north_star); cosine-family values come from a difference-prone dot
    # product, so they also get an absolute floor of 2e-6 (in cosine units).
    if metric in ("cosine_similarity", "cosine_distance", "optimized_similarity"):
        return dict(rtol=1e-5, atol=2e-6)
    if metric == "angular_distance":
        return dict(rtol=1e-5, atol=1e-5 * np.pi)          # arccos is ill-conditioned near 0 / pi
    if metric == "magnitude_difference":
        return dict(rtol=1e-5, atol=1e-5)                  # difference of two norms ~ sqrt(D)
    return dict(rtol=1e-5, atol=1e-30)


def _assert_angles_close(got, truth):
    """arccos amplifies a cosine error e to e / sqrt(1 - c^2) (up to sqrt(2e) at c = +-1): the fp32
    cosine tolerance (1e-5 rel + 2e-6 abs) is propagated through that derivative."""
    c = np.cos(truth)
    e = 1e-5 * np.abs(c) + 2e-6
    tol = 1e-5 * np.abs(truth) + np.minimum(np.sqrt(2 * e) * 1.5, 2 * e / np.sqrt(np.maximum(1 - c * c, 1e-30)))
    assert np.all(np.abs(got - truth) <= tol), np.max(np.abs(got - truth) - tol)


# ------------------------------------------------------------------------------- pairwise values
@pytest.mark.parametrize("D", [1, 3, 7, 33, 64, 512, 2048])
@pytest.mark.parametrize("metric", METRICS + ["optimized_similarity"])
def test_pairwise_fp32_vs_oracle(ops, D, metric):
    Q = synth.gaussian(5, D, 100 + D)
    X = synth.gaussian(300, D, 200 + D)
    X[0] = 0
    X[1] = Q[0]
    kw = {"params": PARAMS} if metric == "optimized_similarity" else {}
    got = ops.pairwise(Q, X, metric, **kw).cpu().numpy()
    truth = OM.pairwise_f64(Q, X, metric, **kw)
    if metric == "angular_distance":
        _assert_angles_close(got, truth)
    else:
        np.testing.assert_allclose(got, truth, **_tol(metric))


def test_pairwise_golden_reference_vectors(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_golden.npz"))
    names = {"l1": "l1_distance", "l2": "l2_distance", "linf": "linf_distance",
             "cosine_similarity": "cosine_similarity", "cosine_distance": "cosine_distance",
             "magnitude_difference": "magnitude_difference"}
    for D in (1, 3, 7, 64, 512, 2048):
        Q, X = g[f"Q_{D}"], g[f"X_{D}"]
        for m, key in names.items():
            got = ops.pairwise(Q, X, m).cpu().numpy()
            np.testing.assert_allclose(got, g[f"{key}_{D}"], rtol=2e-5, atol=1e-5 if m == "magnitude_difference" else 4e-6,
                                       err_msg=f"{m} D={D}")
        got = ops.pairwise(Q, X, "l1", normalized=False).cpu().numpy()
        np.testing.assert_allclose(got, g[f"l1_raw_{D}"], rtol=2e-5)
        got = ops.pairwise(Q, X, "l2", normalized=False).cpu().numpy()
        np.testing.assert_allclose(got, g[f"l2_raw_{D}"], rtol=2e-5)
        got = ops.pairwise(Q, X, "optimized_similarity", params=PARAMS).cpu().numpy()
        np.testing.assert_allclose(got, g[f"optimized_similarity_{D}"], rtol=1e-4, atol=2e-5)
        # zero vector: cos -> 0, angle -> pi/2; duplicate: distance exactly 0
        assert np.all(ops.pairwise(Q, X[:1], "cosine_similarity").cpu().numpy() == 0)
        np.testing.assert_allclose(ops.pairwise(Q, X[:1], "angular_distance").cpu().numpy(), np.pi / 2, rtol=1e-6)
        assert ops.pairwise(Q[:1], X[1:2], "l1").item() == 0 and ops.pairwise(Q[:1], X[1:2], "l2").item() == 0


# ------------------------------------------------------------------------------- top-k
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq,N,D,k", [(1, 1000, 512, 10), (3, 5000, 64, 5), (8, 20000, 128, 100),
                                      (13, 3001, 100, 10), (40, 777, 36, 1), (2, 300, 2048, 256)])
def test_topk_fp32_vs_oracle(ops, metric, nq, N, D, k):
    Q = synth.gaussian(nq, D, 1)
    X = synth.gaussian(N, D, 2)
    s, i = ops.topk(Q, X, metric, k)
    truth = OM.pairwise_f64(Q, X, metric)
    disputed = check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, k, OM.DESCENDING[metric], **_tol(metric))
    # |q|-|x| is a difference of two ~sqrt(D) norms: its smallest values sit within fp32 rounding of
    # each other, so more near-ties are expected there than for the other metrics
    limit = nq * k // 20 if metric == "magnitude_difference" else max(1, nq * k // 200)
    assert disputed <= limit, f"{disputed} disputed ranks"


def test_topk_equals_pairwise_then_stable_sort(ops):
    """Fused top-k == the same kernel's full matrix + the reference's stable sort, bit for bit."""
    Q = synth.gaussian(6, 96, 5)
    X = synth.gaussian(4000, 96, 6)
    for metric in ("l1", "l2", "linf", "cosine_similarity"):
        full = ops.pairwise(Q, X, metric).cpu().numpy()
        tv, ti = OS.topk(full, 20, OM.DESCENDING[metric])
        s, i = ops.topk(Q, X, metric, 20)
        assert np.array_equal(i.cpu().numpy(), ti), metric
        assert np.array_equal(s.cpu().numpy(), tv), metric


@pytest.mark.parametrize("metric", ["l1", "l2", "linf"])
def test_topk_integer_data_exact_with_ties(ops, metric):
    """Small-integer vectors: every distance is exact in fp32, ties abound -> exact equality with
    the stable-sort oracle (lower index wins)."""
    rng = np.random.default_rng(11)
    X = rng.integers(0, 4, size=(5000, 32)).astype(np.float32)
    X[100] = X[7]; X[4000] = X[7]; X[4999] = X[7]
    Q = np.concatenate([X[7:8], rng.integers(0, 4, size=(6, 32)).astype(np.float32)])
    s, i = ops.topk(Q, X, metric, 50)
    tv, ti = OS.topk_search(Q, X, metric, 50, dtype=np.float64)
    assert np.array_equal(i.cpu().numpy(), ti)
    np.testing.assert_allclose(s.cpu().numpy(), tv, rtol=1e-6)
    assert list(i[0, :4].cpu().numpy()) == [7, 100, 4000, 4999]


def test_topk_edge_cases(ops):
    import torch
    Q = synth.gaussian(3, 16, 1)
    X = synth.gaussian(5, 16, 2)
    s, i = ops.topk(Q, X, "l2", 8)                       # k > N: padding
    truth = OM.pairwise_f64(Q, X, "l2")
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 8, False)
    assert np.all(i[:, 5:].cpu().numpy() == -1) and np.all(np.isposinf(s[:, 5:].cpu().numpy()))
    s, i = ops.topk(Q, X, "cosine_similarity", 8)
    assert np.all(np.isneginf(s[:, 5:].cpu().numpy()))
    s, i = ops.topk(Q, np.zeros((0, 16), np.float32), "l1", 4)     # empty store
    assert np.all(i.cpu().numpy() == -1)
    s, i = ops.topk(np.zeros((0, 16), np.float32), X, "l1", 4)     # no queries
    assert s.shape == (0, 4)
    s, i = ops.topk(Q, X, "l1", 3, index_offset=1_000_000_000_000)
    assert np.all(i.cpu().numpy() >= 1_000_000_000_000)
    Xz = X.copy(); Xz[2] = 0                               # zero row: cos 0 exactly (geometric_metrics.py:16-17)
    full = ops.pairwise(Q, Xz, "cosine_similarity").cpu().numpy()
    assert np.all(full[:, 2] == 0)
    with pytest.raises(ValueError):
        ops.topk(Q, X, "l1", 0)
    s, i = ops.topk(Q, X, "l1", 257)                       # more than one 256-row page: served page by page, padded past N
    assert s.shape == (3, 257) and np.all(i[:, 5:].cpu().numpy() == -1)
    with pytest.raises(ValueError):
        ops.topk(Q, X, "l1", 4097)
    with pytest.raises(ValueError):
        ops.topk(Q, synth.gaussian(5, 17, 3), "l1", 2)
    # unaligned view (row stride not a multiple of 16 bytes is copied to contiguous by as_device_matrix)
    big = torch.from_numpy(synth.gaussian(50, 33, 4)).cuda()
    s, i = ops.topk(big[:2, 1:], big[:, 1:], "l1", 3)
    truth = OM.pairwise_f64(big[:2, 1:].cpu().numpy(), big[:, 1:].cpu().numpy(), "l1")
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 3, False)


def test_abs_score_and_optimized_topk(ops):
    Q = synth.gaussian(4, 64, 21)
    X = synth.gaussian(2000, 64, 22)
    s, i = ops.topk(Q, X, "cosine_similarity", 10, abs_score=True)
    truth = np.abs(OM.pairwise_f64(Q, X, "cosine_similarity"))
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 10, True, rtol=1e-5, atol=2e-6)
    s, i = ops.topk(Q, X, "optimized_similarity", 10, params=PARAMS)
    truth = OM.pairwise_f64(Q, X, "optimized_similarity", params=PARAMS)
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 10, True, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("metric", ["l1", "linf", "l2", "cosine_similarity", "angular_distance"])
def test_topk_bf16_database(ops, metric):
    """bf16 rows: the oracle is fed the same bf16-rounded values; arithmetic stays fp32."""
    import torch
    Q = OM.bf16_round(synth.gaussian(9, 256, 31, normalize=True))
    X = OM.bf16_round(synth.gaussian(6000, 256, 32, normalize=True))
    s, i = ops.topk(torch.from_numpy(Q).bfloat16(), torch.from_numpy(X).bfloat16(), metric, 20,
                    flags=ops.FLAG_NO_TENSOR)
    truth = OM.pairwise_f64(Q, X, metric)
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 20, OM.DESCENDING[metric], **_tol(metric))


def test_merge_matches_oracle(ops):
    import torch
    rng = np.random.default_rng(5)
    sc = rng.integers(0, 6, size=(16, 4096)).astype(np.float32)      # many ties across shards
    for desc in (False, True):
        fv, fi = OS.topk(sc, 33, desc)
        pv, pi = [], []
        for r in range(4):
            v, i = OS.topk(sc[:, r * 1024:(r + 1) * 1024], 33, desc)
            pv.append(v); pi.append(i + r * 1024)
        mv, mi = ops.topk_merge(torch.from_numpy(np.stack(pv)).cuda(), torch.from_numpy(np.stack(pi)).cuda(), desc)
        assert np.array_equal(mi.cpu().numpy(), fi) and np.array_equal(mv.cpu().numpy(), fv)
    # padded shard lists (idx -1) are ignored
    v = torch.tensor([[[1.0, np.inf]], [[0.5, 2.0]]]).cuda()
    i = torch.tensor([[[3, -1]], [[10, 11]]]).cuda()
    mv, mi = ops.topk_merge(v, i, False)
    assert mi.cpu().tolist() == [[10, 3]]


@pytest.mark.parametrize("R,k", [(2, 100), (8, 100), (8, 256), (5, 7), (16, 3), (3, 1)])
def test_merge_unbalanced_and_short_shards(ops, R, k):
    """The merge's score bound must hold when one shard owns every winner, when shards are shorter than k
    (padded with idx -1) and when k < R."""
    import torch
    rng = np.random.default_rng(100 + R + k)
    nq = 37
    for desc in (False, True):
        for case in ("skewed", "short", "random"):
            n_r = [k + 50] * R
            if case == "short":
                n_r = [int(x) for x in rng.integers(0, k + 1, size=R)]
                n_r[0] = max(n_r[0], 1)
            cols = []
            for r in range(R):
                a = rng.standard_normal((nq, n_r[r])).astype(np.float32)
                if case == "skewed" and r == R - 1:
                    a += -10.0 if not desc else 10.0                   # this shard owns the whole top-k
                cols.append(a)
            full = np.concatenate(cols, axis=1)
            kk = min(k, full.shape[1])
            fv, fi = OS.topk(full, kk, desc)
            pv = np.full((R, nq, k), -np.inf if desc else np.inf, np.float32)
            pi = np.full((R, nq, k), -1, np.int64)
            off = 0
            for r in range(R):
                if n_r[r]:
                    v, i = OS.topk(cols[r], min(k, n_r[r]), desc)
                    pv[r, :, :v.shape[1]] = v
                    pi[r, :, :v.shape[1]] = i + off
                off += n_r[r]
            mv, mi = ops.topk_merge(torch.from_numpy(pv).cuda(), torch.from_numpy(pi).cuda(), desc)
            mv, mi = mv.cpu().numpy(), mi.cpu().numpy()
            assert np.array_equal(mi[:, :kk], fi) and np.array_equal(mv[:, :kk], fv), (case, desc)
            assert (mi[:, kk:] == -1).all()


def test_sharded_equals_single(ops):
    """Row-sharded search (emulated ranks on one GPU) == single-shard search, exactly."""
    import torch
    from image_retrieval_b200.sharded import shard_range
    Q = synth.gaussian(7, 64, 41)
    X = synth.gaussian(10007, 64, 42)
    for metric in ("l1", "cosine_similarity"):
        s1, i1 = ops.topk(Q, X, metric, 25)
        for R in (2, 4, 8):
            ps, pi = [], []
            for r in range(R):
                b, e = shard_range(len(X), R, r)
                s, i = ops.topk(Q, X[b:e], metric, 25, index_offset=b)
                ps.append(s); pi.append(i)
            ms, mi = ops.topk_merge(torch.stack(ps), torch.stack(pi), OM.DESCENDING[metric])
            assert torch.equal(mi, i1) and torch.equal(ms, s1), (metric, R)


# ------------------------------------------------------------------------------- histograms
@pytest.mark.parametrize("cs", ["rgb", "hsv"])
def test_histogram_bit_exact(ops, cs, golden_dir):
    g = np.load(os.path.join(golden_dir, "hist_golden.npz"))
    got = ops.histogram(g["images"], cs).cpu().numpy().astype(np.uint32)
    assert np.array_equal(got, g[cs])                                   # OpenCV-generated fixture
    for imgs in (synth.images_uniform(5, 224, 224, 1), synth.images_palette(5, 224, 224, 2),
                 synth.images_uniform(3, 37, 53, 3),                    # pixel count not a multiple of 16
                 synth.images_uniform(1, 600, 500, 4)):                 # sliced (multi-CTA) image
        got = ops.histogram(imgs, cs).cpu().numpy().astype(np.uint32)
        assert np.array_equal(got, OH.histogram(imgs, cs))
    flat = np.full((2, 64, 64, 3), 200, np.uint8)                       # worst-case contention
    assert np.array_equal(ops.histogram(flat, cs).cpu().numpy().astype(np.uint32), OH.histogram(flat, cs))


@pytest.mark.parametrize("cs", ["rgb", "hsv"])
def test_histogram_every_colour(ops, cs):
    """All 2^24 RGB colours, 4096 per image (so a mis-binned colour cannot hide behind another one): the kernel's integer
    HSV conversion and binning agree with the oracle (itself checked against OpenCV on all colours) everywhere."""
    c = np.arange(1 << 24, dtype=np.uint32)
    imgs = np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], axis=1).astype(np.uint8).reshape(4096, 64, 64, 3)
    got = ops.histogram(imgs, cs).cpu().numpy().astype(np.uint32)
    want = np.concatenate([OH.histogram(imgs[i:i + 512], cs) for i in range(0, 4096, 512)])
    assert np.array_equal(got, want)


def test_histogram_embedding(ops):
    imgs = synth.images_palette(6, 64, 64, 9)
    raw, unit, mag = ops.counts_to_embedding(ops.histogram(imgs))
    ou, om = OH.embedding(imgs)
    np.testing.assert_allclose(unit.cpu().numpy(), ou, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(mag.cpu().numpy(), om, rtol=1e-6)
    assert np.array_equal(raw.cpu().numpy(), OH.histogram(imgs).astype(np.float32))


def test_config1_histogram_l2_search(ops):
    """BASELINE config 1 in miniature: histogram embeddings, L2 top-10 with self matches."""
    db = synth.images_palette(400, 48, 48, 1002)
    qi = db[:20]
    X = ops.counts_to_embedding(ops.histogram(db))[0]
    Q = ops.counts_to_embedding(ops.histogram(qi))[0]
    s, i = ops.topk(Q, X, "l2", 10)
    Xh = OH.histogram(db).astype(np.float32)
    tv, ti = OS.topk_search(Xh[:20], Xh, "l2", 10, dtype=np.float64)
    assert np.array_equal(i.cpu().numpy(), ti)                          # integer counts: exact
    np.testing.assert_allclose(s.cpu().numpy(), tv, rtol=1e-6)
    assert np.all(i[:, 0].cpu().numpy() == np.arange(20)) and np.all(s[:, 0].cpu().numpy() == 0)


# ------------------------------------------------------------------------------- tcgen05 path
def _bf16_case(nq, N, D, seed, normalize=True):
    Q = OM.bf16_round(synth.gaussian(nq, D, seed, normalize=normalize))
    X = OM.bf16_round(synth.gaussian(N, D, seed + 1, normalize=normalize))
    return Q, X


@pytest.mark.parametrize("metric", ["cosine_similarity", "cosine_distance", "angular_distance", "l2"])
@pytest.mark.parametrize("nq,N,D,k", [(128, 2048, 512, 10), (200, 5000, 512, 100), (33, 70001, 256, 100),
                                      (300, 3000, 64, 5), (64, 1500, 136, 200)])
def test_tensor_path_topk_vs_oracle(ops, metric, nq, N, D, k):
    """bf16 store on the tcgen05 path (default flags: exact fp32 re-rank of the k' candidates)."""
    import torch
    Q, X = _bf16_case(nq, N, D, 50 + D)
    s, i = ops.topk(torch.from_numpy(Q).bfloat16(), torch.from_numpy(X).bfloat16(), metric, k)
    truth = OM.pairwise_f64(Q, X, metric)
    tol = _tol(metric)
    disputed = check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, k, OM.DESCENDING[metric], **tol)
    assert disputed <= max(1, nq * k // 200), f"{disputed} disputed ranks"


def test_tensor_path_matches_scan_path(ops):
    """Same inputs through the tcgen05 path and the CUDA-core scan: identical index lists."""
    import torch
    Q, X = _bf16_case(150, 9000, 512, 77)
    Qt, Xt = torch.from_numpy(Q).bfloat16().cuda(), torch.from_numpy(X).bfloat16().cuda()
    for metric in ("cosine_similarity", "l2"):
        s1, i1 = ops.topk(Qt, Xt, metric, 50)
        s2, i2 = ops.topk(Qt, Xt, metric, 50, flags=ops.FLAG_NO_TENSOR)
        assert (i1 == i2).float().mean().item() > 0.999, metric
        torch.testing.assert_close(s1, s2, rtol=1e-5, atol=2e-6)


def test_tensor_path_near_duplicates_and_no_rerank(ops):
    """Cancellation case of the |q|^2+|x|^2-2qx form: near-duplicate rows.  The re-rank makes the
    reported L2 exact; without it the GEMM-form value is only bounded loosely (stated: 1e-2 abs)."""
    import torch
    rng = np.random.default_rng(3)
    Q, X = _bf16_case(64, 4096, 512, 91, normalize=False)
    X[:64] = Q                                            # exact duplicates of the queries
    X[64:128] = OM.bf16_round(Q + 0.01 * rng.standard_normal(Q.shape).astype(np.float32))
    Qt, Xt = torch.from_numpy(Q).bfloat16().cuda(), torch.from_numpy(X).bfloat16().cuda()
    s, i = ops.topk(Qt, Xt, "l2", 8)
    truth = OM.pairwise_f64(Q, X, "l2")
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 8, False, rtol=1e-5, atol=1e-30)
    assert np.all(i[:, 0].cpu().numpy() == np.arange(64)) and np.all(s[:, 0].cpu().numpy() == 0)
    s2, i2 = ops.topk(Qt, Xt, "l2", 8, flags=ops.FLAG_NO_RERANK)
    assert np.all(i2[:, 0].cpu().numpy() == np.arange(64))
    got = s2.cpu().numpy()
    want = np.take_along_axis(truth, i2.cpu().numpy(), 1)
    assert np.max(np.abs(got - want)) < 1e-2
    # zero rows and zero queries keep the reference's cos := 0 convention
    X[200] = 0
    Q[5] = 0
    s, i = ops.topk(torch.from_numpy(Q).bfloat16(), torch.from_numpy(X).bfloat16(), "cosine_similarity", 4096 // 32)
    truth = OM.pairwise_f64(Q, X, "cosine_similarity")
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 128, True, rtol=1e-5, atol=2e-6)
    assert np.all(s[5].cpu().numpy() == 0) and list(i[5, :3].cpu().numpy()) == [0, 1, 2]


def test_tensor_path_sharded_and_abs(ops):
    import torch
    from image_retrieval_b200.sharded import shard_range
    Q, X = _bf16_case(130, 20000, 512, 17)
    Qt, Xt = torch.from_numpy(Q).bfloat16().cuda(), torch.from_numpy(X).bfloat16().cuda()
    s1, i1 = ops.topk(Qt, Xt, "cosine_similarity", 100)
    ps, pi = [], []
    for r in range(4):
        b, e = shard_range(len(X), 4, r)
        s, i = ops.topk(Qt, Xt[b:e], "cosine_similarity", 100, index_offset=b)
        ps.append(s); pi.append(i)
    ms, mi = ops.topk_merge(torch.stack(ps), torch.stack(pi), True)
    assert torch.equal(mi, i1) and torch.equal(ms, s1)
    s, i = ops.topk(Qt, Xt, "cosine_similarity", 20, abs_score=True)
    truth = np.abs(OM.pairwise_f64(Q, X, "cosine_similarity"))
    check_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 20, True, rtol=1e-5, atol=2e-6)


# ------------------------------------------------------------------------------- BASELINE full sizes
def _device_unit_rows(n, d, seed, dtype):
    import torch
    out = torch.empty((n, d), dtype=dtype, device="cuda")
    for c, s in enumerate(range(0, n, 250_000)):
        g = torch.Generator(device="cuda")
        g.manual_seed(seed * 1000 + c)
        e = min(n, s + 250_000)
        x = torch.randn((e - s, d), generator=g, device="cuda", dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        out[s:e] = x.to(dtype)
    return out


def test_full_size_config2_cosine_top100(ops):
    """BASELINE configs[1] at full size (1M x 512 bf16, 10k queries, top-100): a sample of queries against the
    fp64 oracle, plus size-independent properties: sharded == unsharded, angle/cosine-distance consistent."""
    import torch
    X = _device_unit_rows(1_000_000, 512, 2001, torch.bfloat16)
    Q = _device_unit_rows(10_000, 512, 2002, torch.bfloat16)
    s, i = ops.topk(Q, X, "cosine_similarity", 100)
    sel = [0, 1, 4999, 9999]
    Xh = X.float().cpu().numpy()
    truth = OM.pairwise_f64(Q[sel].float().cpu().numpy(), Xh, "cosine_similarity")
    disputed = check_topk(s[sel].cpu().numpy(), i[sel].cpu().numpy(), truth, 100, True, rtol=1e-5, atol=2e-6)
    assert disputed <= 2
    # sortedness + uniqueness over the whole result
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    srt = torch.sort(i, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # row-sharded (4 shards) + merge == unsharded, exactly
    ps, pi = [], []
    for r in range(4):
        b, e = r * 250_000, (r + 1) * 250_000
        ss, ii = ops.topk(Q[:512], X[b:e], "cosine_similarity", 100, index_offset=b)
        ps.append(ss); pi.append(ii)
    ms, mi = ops.topk_merge(torch.stack(ps), torch.stack(pi), True)
    assert torch.equal(mi, i[:512]) and torch.equal(ms, s[:512])
    # angle selects the same rows (its own order re-sorts fp32-equal angles by index, so compare as sets)
    sa, ia = ops.topk(Q[:256], X, "angular_distance", 100)
    assert torch.equal(torch.sort(ia, dim=1).values, torch.sort(i[:256], dim=1).values)
    torch.testing.assert_close(torch.cos(sa), s[:256], rtol=1e-5, atol=2e-6)
    assert bool((sa[:, 1:] >= sa[:, :-1]).all())


def test_full_size_config3_l1_linf_top10(ops):
    """BASELINE configs[2] at full size (1M x 2048 fp32): L1 / Linf top-10 of an 8-query batch vs the fp64 oracle."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(3001)
    X = torch.relu(torch.randn((1_000_000, 2048), generator=g, device="cuda"))
    Q = torch.relu(torch.randn((8, 2048), generator=g, device="cuda"))
    Xh = X.cpu().numpy()
    Qh = Q.cpu().numpy()
    for metric in ("l1", "linf"):
        s, i = ops.topk(Q, X, metric, 10)
        truth = np.stack([OM.pairwise_f64(Qh[j:j + 1], Xh, metric, chunk=50_000)[0] for j in range(2)])
        check_topk(s[:2].cpu().numpy(), i[:2].cpu().numpy(), truth, 10, False, rtol=1e-5, atol=1e-30)
        # every returned score equals the reference formula on the returned row (all 8 queries)
        rows = Xh[i.cpu().numpy()]                                       # (8, 10, 2048)
        d = np.abs(rows.astype(np.float64) - Qh[:, None, :].astype(np.float64))
        want = d.sum(-1) / 2048 if metric == "l1" else d.max(-1)
        np.testing.assert_allclose(s.cpu().numpy(), want, rtol=1e-5)
        # a second identical call is bit-identical (no atomics-order dependence in the result)
        s2, i2 = ops.topk(Q, X, metric, 10)
        assert torch.equal(i2, i) and torch.equal(s2, s)


# ------------------------------------------------------------------------------- all-pairs evaluation (config 5)
def test_allpairs_eval_counts(ops):
    """Integer outputs: bit-exact against the oracle's binning of the SAME per-pair values (the pairwise kernel shares
    the arithmetic), and within a handful of bin-edge flips of the fp64 oracle; PR counts == the reference's loop."""
    from oracle import evaluation as E
    N, D, nbins = 777, 96, 256
    X = synth.gaussian(N, D, 61)
    X[5] = X[4]                                        # exact duplicate pair: every distance 0
    cat = np.arange(N) % 10
    col = (np.arange(N) // 10) % 3
    ranges = {"cosine_distance": (0.0, 2.0), "l1_distance": (0.0, 2.5), "l2_distance": (0.0, 3.0), "linf_distance": (0.0, 8.0),
              "magnitude_difference": (0.0, 5.0)}
    thresholds = np.linspace(0, 1, 100)
    hist, thr = ops.allpairs_eval(X, cat, col, ranges, nbins, thresholds)
    hist, thr = hist.cpu().numpy(), thr.cpu().numpy()
    rel = E.relationship(cat, col)
    npairs = N * (N - 1) // 2
    assert hist.sum(axis=(1, 2)).tolist() == [npairs] * 5
    # same fp32 values (through the pairwise ABI) -> identical counts
    names = {"cosine_distance": "cosine_distance", "l1_distance": "l1", "l2_distance": "l2", "linf_distance": "linf",
             "magnitude_difference": "magnitude_difference"}
    vals = {m: ops.pairwise(X, X, names[m]).cpu().numpy() for m in E.METRICS}
    # the eval kernel pairs (query i, row j > i); pairwise[i, j] is the same orientation
    h2, t2 = E.bin_counts(vals, rel, ranges, nbins, thresholds)
    assert np.abs(hist - h2).sum() <= 10 and np.abs(thr - t2).sum() <= 10          # different TQ grouping: ulp-level flips only
    # fp64 oracle: only bin-edge flips
    h3, t3 = E.bin_counts(E.metric_matrices(X, np.float64), rel, ranges, nbins, thresholds)
    assert np.abs(hist - h3).sum() <= 2e-4 * 5 * npairs + 10
    assert np.abs(thr - t3).sum() <= 2e-4 * 5 * npairs + 10
    # precision / recall counts against the reference loop on the GPU's own cosine distances
    iu = np.triu_indices(N, 1)
    r = rel[iu]
    sel = r <= 1
    ref = E.pr_curve_reference(list(vals["cosine_distance"][iu][sel]), list((r[sel] == 1).astype(int)), thresholds)
    assert np.abs(E.pr_from_counts(thr[0]) - ref).max() <= 3
    # relationship-type totals
    for t in range(4):
        assert hist[0, t].sum() == int((r == t).sum())


@pytest.mark.parametrize("D", [33, 7, 100])
def test_allpairs_eval_unaligned_rows(ops, D):
    """Rows that are not 16-byte aligned take the cp.async / element-load ring (one stage deep in evaluation mode)."""
    from oracle import evaluation as E
    N, nbins = 300, 64
    X = synth.gaussian(N, D, 80 + D)
    cat, col = np.arange(N) % 7, (np.arange(N) // 7) % 3
    ranges = {m: (0.0, 6.0) for m in E.METRICS}
    thresholds = np.linspace(0, 1, 100)
    hist, thr = ops.allpairs_eval(X, cat, col, ranges, nbins, thresholds)
    hist, thr = hist.cpu().numpy(), thr.cpu().numpy()
    names = {"cosine_distance": "cosine_distance", "l1_distance": "l1", "l2_distance": "l2", "linf_distance": "linf",
             "magnitude_difference": "magnitude_difference"}
    vals = {m: ops.pairwise(X, X, names[m]).cpu().numpy() for m in E.METRICS}
    h2, t2 = E.bin_counts(vals, E.relationship(cat, col), ranges, nbins, thresholds)
    assert hist.sum(axis=(1, 2)).tolist() == [N * (N - 1) // 2] * 5
    assert np.abs(hist - h2).sum() <= 10 and np.abs(thr - t2).sum() <= 10


def test_allpairs_eval_counter_flush_at_scale(ops):
    """20k rows of ONE repeated vector pattern: every pair lands in the same few bins, so the 16-bit shared counters
    would wrap without the periodic flush.  Totals and per-type totals are exact (size-independent properties)."""
    N, D, nbins = 20000, 32, 64
    base = synth.gaussian(4, D, 71)
    X = base[np.arange(N) % 4]                                    # only 4 distinct rows -> <= 10 distinct distances
    cat = np.arange(N) % 10
    col = (np.arange(N) // 10) % 3
    ranges = {m: (0.0, 4.0) for m in ("cosine_distance", "l1_distance", "l2_distance", "linf_distance", "magnitude_difference")}
    hist, thr = ops.allpairs_eval(X, cat, col, ranges, nbins, np.linspace(0, 1, 100))
    hist, thr = hist.cpu().numpy(), thr.cpu().numpy()
    npairs = N * (N - 1) // 2
    assert hist.sum(axis=(1, 2)).tolist() == [npairs] * 5
    same_cat = sum(int(c) * (int(c) - 1) // 2 for c in np.bincount(cat))
    key = cat * 3 + col
    same_both = sum(int(c) * (int(c) - 1) // 2 for c in np.bincount(key))
    same_col = sum(int(c) * (int(c) - 1) // 2 for c in np.bincount(col))
    want = [same_both, same_cat - same_both, same_col - same_both, npairs - same_cat - same_col + same_both]
    for m in range(5):
        assert hist[m].sum(axis=1).tolist() == want, m
        assert thr[m].sum(axis=1).tolist() == want[:2], m
    # identical rows (i = j mod 4): distance exactly 0 -> bin 0 holds at least those pairs
    ident = sum(int(c) * (int(c) - 1) // 2 for c in np.bincount(np.arange(N) % 4))
    assert hist[1, :, 0].sum() >= ident and hist[3, :, 0].sum() >= ident


# ------------------------------------------------------------------------------- explicit pair lists
@pytest.mark.parametrize("D,dtype", [(512, "f32"), (64, "f32"), (7, "f32"), (513, "f32"), (512, "bf16"), (36, "bf16")])
def test_pair_metrics_match_oracle(ops, D, dtype):
    """b200ir_pair_metrics == get_all_metrics pair by pair (geometric_metrics.py:114-129), incl. a zero row,
    a self pair and rows named outside the store (NaN column)."""
    import torch
    rng = np.random.default_rng(D)
    A = synth.gaussian(300, D, 61)
    B = synth.gaussian(200, D, 62)
    A[5] = 0.0
    B[9] = A[17]
    if dtype == "bf16":
        A, B = OM.bf16_round(A), OM.bf16_round(B)
    P = 1000
    ia = rng.integers(0, 300, size=P); ib = rng.integers(0, 200, size=P)
    ia[:3] = (5, 17, 5); ib[:3] = (0, 9, 9)
    ia[10], ib[11] = 300, -1                                     # unknown rows
    tA = torch.from_numpy(A).bfloat16() if dtype == "bf16" else A
    tB = torch.from_numpy(B).bfloat16() if dtype == "bf16" else B
    got = ops.pair_metrics(tA, tB, ia, ib).cpu().numpy()
    assert got.shape == (7, P) and np.isnan(got[:, 10]).all() and np.isnan(got[:, 11]).all()
    for p in [q for q in range(P) if q not in (10, 11)]:
        want = OM.get_all_metrics(A[ia[p]], B[ib[p]])
        for mi, name in enumerate(ops.PAIR_METRICS):
            # absolute floors: cosine family 2e-6, arccos amplification near 0 / pi, |a|-|b| is a difference of ~sqrt(D) norms
            tol = {"cosine_similarity": 2e-6, "cosine_distance": 2e-6, "angular_distance": 2e-3, "magnitude_difference": 1e-5}.get(name, 1e-6)
            assert abs(got[mi, p] - float(want[name])) <= 1e-5 * abs(float(want[name])) + tol, (p, name, got[mi, p], want[name])
    assert got[0, 0] == 0.0 and got[1, 0] == 1.0 and abs(got[2, 0] - np.pi / 2) < 1e-6       # zero vector (:16-17)
    assert got[5, 1] == 0.0 and got[3, 1] == 0.0                                             # identical rows
    same = ops.pair_metrics(tA, None, [1, 2], [2, 1]).cpu().numpy()                          # symmetric within one store
    assert np.array_equal(same[:, 0], same[:, 1])
    assert ops.pair_metrics(tA, None, [], []).shape == (7, 0)


# ------------------------------------------------------------------------------- image front-end
def test_resize_crop_bit_exact_pil_golden(ops, golden_dir):
    """b200ir_resize_crop == PIL BICUBIC resize + centre crop (fixtures written by PIL, tests/golden/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "resize_golden.npz"))
    for n, ((H, W, size, seed), gen) in enumerate(zip(g["cases"], g["generators"])):
        img = getattr(synth, f"images_{gen}")(1, int(H), int(W), int(seed))[0]
        got = ops.resize_crop(img, int(size)).cpu().numpy()
        assert got.shape == (1, size, size, 3)
        assert np.array_equal(got[0], g[f"out_{n}"]), (H, W, size)


@pytest.mark.parametrize("H,W,size", [(480, 640, 224), (640, 480, 224), (224, 224, 224), (100, 150, 224), (333, 517, 224),
                                      (225, 300, 224), (50, 37, 64), (1200, 900, 224), (224, 500, 224), (3000, 2000, 224)])
def test_resize_crop_bit_exact_oracle(ops, H, W, size):
    from oracle import resize as R
    imgs = np.concatenate([synth.images_uniform(2, H, W, H + W), synth.images_palette(1, H, W, H * W)])
    got = ops.resize_crop(imgs, size).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], R.clip_preprocess_u8(imgs[b], size)), b
    # unaligned device view (odd byte offset) and an explicit window of the resized image
    import torch
    flat = torch.zeros(imgs[0].size + 5, dtype=torch.uint8, device="cuda")
    flat[5:] = torch.from_numpy(imgs[0]).cuda().view(-1)
    view = flat[5:].view(1, H, W, 3)
    nh, nw = R.shortest_edge_size(H, W, size)
    win = (nh // 3, nw // 4, min(17, nh - nh // 3), min(29, nw - nw // 4))
    got = ops.resize_crop(view, size, crop=win).cpu().numpy()[0]
    want = R.resize_bicubic(imgs[0], nh, nw)[win[0]:win[0] + win[2], win[1]:win[1] + win[3]]
    assert np.array_equal(got, want)


def test_resize_then_histogram_pipeline(ops):
    """front-end + embedding producer: histogram(resize_crop(img)) == oracle histogram of the oracle-resized image."""
    from oracle import resize as R
    imgs = synth.images_palette(4, 300, 400, 77)
    counts = ops.histogram(ops.resize_crop(imgs, 224), "rgb").cpu().numpy()
    for b in range(4):
        assert np.array_equal(counts[b], OH.histogram(R.clip_preprocess_u8(imgs[b], 224)[None], "rgb")[0])


# ------------------------------------------------------------------------------- post-filter
@pytest.mark.parametrize("relative", [False, True])
@pytest.mark.parametrize("kc,top_k", [(15, 5), (96, 32), (256, 100), (7, 10)])
def test_threshold_dedupe_matches_oracle(ops, kc, top_k, relative):
    """b200ir_threshold_dedupe == image_search.py:115-140 (oracle.search.threshold_and_dedupe) per query."""
    import torch
    rng = np.random.default_rng(kc * 7 + top_k)
    nq, N = 41, 500
    group = rng.integers(0, 120, size=N)                       # ~4 rows per path: many duplicates
    group = np.array([np.flatnonzero(group == g)[0] for g in group], dtype=np.int64)
    sc = np.sort(rng.uniform(-0.2, 1.0, size=(nq, kc)).astype(np.float32), axis=1)[:, ::-1].copy()
    sc[3, 2:6] = sc[3, 2]                                       # ties
    idx = np.stack([rng.permutation(N)[:kc] for _ in range(nq)]).astype(np.int64)
    nvalid = rng.integers(0, kc + 1, size=nq); nvalid[0] = 0; nvalid[1] = kc
    for q in range(nq):
        idx[q, nvalid[q]:] = -1
        sc[q, nvalid[q]:] = -np.inf
    thr = 0.25
    for grp in (group, None):
        fs, fi, cnt = ops.threshold_dedupe(torch.from_numpy(sc).cuda(), torch.from_numpy(idx).cuda(), top_k, thr, relative, grp)
        fs, fi, cnt = fs.cpu().numpy(), fi.cpu().numpy(), cnt.cpu().numpy()
        for q in range(nq):
            matches = [{"path": int(group[idx[q, j]]) if grp is not None else int(idx[q, j]), "score": sc[q, j], "row": int(idx[q, j])}
                       for j in range(nvalid[q])]
            want = OS.threshold_and_dedupe(matches, top_k, thr, relative)
            assert cnt[q] == len(want), (q, cnt[q], len(want))
            assert [int(r) for r in fi[q, :cnt[q]]] == [m["row"] for m in want]
            assert np.array_equal(fs[q, :cnt[q]], np.array([m["score"] for m in want], np.float32))
            assert (fi[q, cnt[q]:] == -1).all() and np.isneginf(fs[q, cnt[q]:]).all()
