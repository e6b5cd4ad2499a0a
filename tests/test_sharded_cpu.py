"""world_size-2 / -3 gloo test of the row-sharded search: shard ranges, index offsets, the all-to-all of
query slices, the merge order and the final all-gather.  The local search and the merge are the oracle here (this is a
CPU test of the host logic; the CUDA operators are covered by -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_retrieval_b200.sharded import ShardedIndex, shard_range
    from oracle import metrics as OM
    from oracle import search as OS
    from oracle import synth

    X = synth.gaussian(1001, 32, 7)
    X[900] = X[3]                      # tie across shards -> lower global index first
    Q = np.concatenate([X[3:4], synth.gaussian(4, 32, 8)])

    def local_topk(Qm, Xm, metric, k, index_offset, **kw):
        v, i = OS.topk_search(Qm, Xm, metric, k, dtype=np.float32)
        pad = k - v.shape[1]
        if pad:
            v = np.pad(v, ((0, 0), (0, pad)), constant_values=np.inf)
            i = np.pad(i, ((0, 0), (0, pad)), constant_values=-1 - index_offset)
        return torch.from_numpy(v.astype(np.float32)), torch.from_numpy(i + index_offset)

    def merge(s, i, desc):
        v, ix = OS.merge_topk(s.numpy(), i.numpy(), s.shape[2], desc)
        return torch.from_numpy(v), torch.from_numpy(ix)

    b, e = shard_range(len(X), world, rank)
    index = ShardedIndex(X[b:e], b, local_topk=local_topk, merge=merge)
    res = {}
    from image_retrieval_b200.sharded import query_slice
    for metric in ("l1", "cosine_similarity"):
        s, i = index.topk(Q, metric, 9)
        res[metric] = (s.numpy(), i.numpy())
        # the slice every rank serves on its own (no final all-gather) is the matching rows of the full result
        ss, si, q0, q1 = index.topk_slice(Q, metric, 9)
        assert (q0, q1) == query_slice(len(Q), world, rank)
        assert np.array_equal(si.numpy()[:q1 - q0], i.numpy()[q0:q1]) and np.array_equal(ss.numpy()[:q1 - q0], s.numpy()[q0:q1])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{f"{m}_{n}": a for m, (s, i) in res.items() for n, a in (("s", s), ("i", i))})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_equals_single(tmp_path, world):
    """2 ranks, and 3 ranks with uneven shards (1001 rows): every rank ends with the unsharded result."""
    from oracle import search as OS
    from oracle import synth
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    X = synth.gaussian(1001, 32, 7)
    X[900] = X[3]
    Q = np.concatenate([X[3:4], synth.gaussian(4, 32, 8)])
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for metric in ("l1", "cosine_similarity"):
        v, i = OS.topk_search(Q, X, metric, 9, dtype=np.float32)
        for r in ranks:
            assert np.array_equal(r[f"{metric}_i"], i), metric
            assert np.array_equal(r[f"{metric}_s"], v.astype(np.float32)), metric
    assert list(ranks[0]["l1_i"][0, :2]) == [3, 900]


def _eval_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_retrieval_b200 import sharded
    from oracle import evaluation as E
    from oracle import synth
    N, nbins = 203, 64
    X = synth.gaussian(N, 24, 17)
    cat, col = np.arange(N) % 5, (np.arange(N) // 5) % 3
    ranges = {m: (0.0, 4.0) for m in E.METRICS}
    thresholds = np.linspace(0, 1, 20)
    vals = E.metric_matrices(X, np.float32)

    def local_eval(part, nparts):
        # this part's rows i: groups of 8 rows dealt round-robin (what b200ir_allpairs_eval_part counts)
        rel = E.relationship(cat, col).copy()
        mine = (np.arange(N) // 8) % nparts == part
        rel[~mine, :] = 9                                   # pairs (i, j > i) of foreign rows i match no relationship type
        h, t = E.bin_counts(vals, rel, ranges, nbins, thresholds)
        return torch.from_numpy(h), torch.from_numpy(t)

    hist, thr = sharded.allpairs_eval(X, cat, col, ranges, nbins, thresholds, local_eval=local_eval)
    np.savez(os.path.join(out_dir, f"eval{rank}.npz"), hist=hist.numpy(), thr=thr.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_allpairs_eval_counts_add_up(tmp_path):
    """3 gloo ranks: the cyclic row shares + one all-reduce of the integer counts equal the unsharded evaluation."""
    from oracle import evaluation as E
    from oracle import synth
    world = 3
    mp.spawn(_eval_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    N, nbins = 203, 64
    X = synth.gaussian(N, 24, 17)
    cat, col = np.arange(N) % 5, (np.arange(N) // 5) % 3
    ranges = {m: (0.0, 4.0) for m in E.METRICS}
    h, t = E.bin_counts(E.metric_matrices(X, np.float32), E.relationship(cat, col), ranges, nbins, np.linspace(0, 1, 20))
    for r in range(world):
        got = np.load(tmp_path / f"eval{r}.npz")
        assert np.array_equal(got["hist"], h) and np.array_equal(got["thr"], t)
