"""Row-sharded multi-GPU search: one process per GPU, each rank owns a
contiguous range of database rows and runs the fused distance + top-k locally with its row base as
`index_offset`.  The per-shard lists then meet in ONE exchange step:

  all-to-all   rank r receives, from every shard, the lists of ITS slice of the query batch
               (nq / R queries: R x smaller than an all-gather of everything to everybody),
  merge        of those R lists per query with the reference's (score, index) order
               (app_pipeline.py:171-172: stable sort -> ties to the lower global index),
  all-gather   of the merged slices when every rank wants the full result (`topk`); a rank that serves
               only its slice of the query stream stops before it (`topk_slice`).
Small batches (nq * k <= 64 K entries, e.g. the 8-query scans) skip all that: one all-gather of a packed
[scores | ids] record per rank and a merge of every query on every rank - a single latency-bound collective.

The scan itself is embarrassingly parallel; no other data-path collective exists.  Bytes per rank at
nq = 10k, k = 100, R = 8: 10.5 MB out + 10.5 MB in for the all-to-all and 1.5 MB / 10.5 MB for the
all-gather (an all-gather of the raw lists moved 12 MB out / 84 MB in and merged 8 x more queries per rank).

The local search and the merge are injectable so that the sharding / collective logic can be
exercised with the gloo backend on CPU (tests/test_sharded_cpu.py injects the oracle); the
defaults are the CUDA operators and there is no CPU fallback in the product path.
"""
import torch
import torch.distributed as dist

from . import ops


def shard_range(n_rows, world_size, rank):
    """Contiguous balanced split: the first n_rows % world_size ranks own one extra row."""
    base, rem = divmod(int(n_rows), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def query_slice(nq, world_size, rank):
    """Queries rank `rank` merges and serves: [q0, q1) of equal slices of ceil(nq / world_size)."""
    sl = -(-int(nq) // int(world_size))
    q0 = min(int(nq), rank * sl)
    return q0, min(int(nq), q0 + sl)


def _default_local(Q, X, metric, k, index_offset, **kw):
    return ops.topk(Q, X, metric, k, index_offset=index_offset, **kw)


def _default_merge(scores, idx, descending):
    return ops.topk_merge(scores, idx, descending)


class ShardedIndex:
    def __init__(self, local_rows, row_begin, group=None, local_topk=None, merge=None, prepare=True):
        # default (CUDA) path: keep the per-shard search state (row norms, split planes) next to the rows;
        # prepare=False skips it (stores that are only searched with L1 / Linf / optimized have no use for it)
        if prepare and local_topk is None and isinstance(local_rows, torch.Tensor) and local_rows.is_cuda:
            local_rows = ops.prepare_index(local_rows)
        self.X = local_rows
        self.row_begin = int(row_begin)
        self.group = group
        self.local_topk = local_topk or _default_local
        self.merge = merge or _default_merge
        self._send = {}

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def _send_buffers(self, nq_pad, k, descending, device):
        """(nq_pad, k) score / id buffers whose rows past the query batch are permanent padding lists."""
        key = (nq_pad, k, descending, str(device))
        buf = self._send.get(key)
        if buf is None:
            s = torch.full((nq_pad, k), float("-inf") if descending else float("inf"), dtype=torch.float32, device=device)
            i = torch.full((nq_pad, k), -1, dtype=torch.int64, device=device)
            if len(self._send) > 8:
                self._send.clear()
            buf = self._send[key] = (s, i)
        return buf

    def topk_slice(self, Q, metric, k, **kw):
        """Global top-k of THIS rank's slice of the query batch: (scores (n, k), indices (n, k), q0, q1) with
        n = q1 - q0 and [q0, q1) = query_slice(nq, world, rank).  Single process: the whole batch."""
        m = ops.metric_id(metric)
        desc = m in ops.DESCENDING
        R, r = self._world()
        default = self.local_topk is _default_local
        if R == 1:
            s, i = self.local_topk(Q, self.X, metric, k, self.row_begin, **kw)
            return s, i, 0, s.shape[0]
        if default:
            Q = ops.as_device_matrix(Q, dtype=self.X.dtype)
        elif getattr(Q, "ndim", 2) == 1:
            Q = Q[None]
        nq = Q.shape[0]
        sl = -(-nq // R)
        nq_pad = sl * R
        if default:
            send_s, send_i = self._send_buffers(nq_pad, k, desc, Q.device)
            ops.topk(Q, self.X, metric, k, index_offset=self.row_begin, out=(send_s[:nq], send_i[:nq]), **kw)
        else:
            s, i = self.local_topk(Q, self.X, metric, k, self.row_begin, **kw)
            send_s, send_i = self._send_buffers(nq_pad, k, desc, s.device)
            send_s[:nq] = s
            send_i[:nq] = i
        recv_s = torch.empty((R, sl, k), dtype=torch.float32, device=send_s.device)
        recv_i = torch.empty((R, sl, k), dtype=torch.int64, device=send_s.device)
        dist.all_to_all_single(recv_s.view(-1), send_s.view(-1), group=self.group)
        dist.all_to_all_single(recv_i.view(-1), send_i.view(-1), group=self.group)
        ms, mi = self.merge(recv_s, recv_i, desc)
        q0, q1 = query_slice(nq, R, r)
        return ms, mi, q0, q1                                      # rows past q1 - q0 are padding lists

    SMALL_PAYLOAD_ENTRIES = 1 << 16      # nq * k up to this: one packed all-gather beats four latency-bound collectives

    def _topk_small(self, Q, metric, m, k, **kw):
        """Few queries (e.g. the 8-query L1 / Linf scans): the lists are tiny, so ONE all-gather of a packed record
        [scores | ids] per rank and a merge of every query on every rank (CUDA operators only)."""
        R, _ = self._world()
        Qd = ops.as_device_matrix(Q, dtype=self.X.dtype)
        nq = Qd.shape[0]
        score_bytes = -(-nq * k * 4 // 16) * 16
        rec = -(-(score_bytes + nq * k * 8) // 16) * 16
        buf = torch.empty(rec, dtype=torch.uint8, device=Qd.device)
        s = buf[:nq * k * 4].view(torch.float32).view(nq, k)
        i = buf[score_bytes:score_bytes + nq * k * 8].view(torch.int64).view(nq, k)
        ops.topk(Qd, self.X, metric, k, index_offset=self.row_begin, out=(s, i), **kw)
        gathered = torch.empty(R * rec, dtype=torch.uint8, device=Qd.device)
        dist.all_gather_into_tensor(gathered, buf, group=self.group)
        return ops.topk_merge_packed(gathered, R, nq, k, score_bytes, m in ops.DESCENDING)

    def topk(self, Q, metric, k, **kw):
        """Global top-k over all shards; every rank returns the same (scores (nq, k), indices (nq, k))."""
        R, _ = self._world()
        if R > 1 and self.local_topk is _default_local and self.merge is _default_merge:
            nq = 1 if getattr(Q, "ndim", 2) == 1 else Q.shape[0]
            if nq * k <= self.SMALL_PAYLOAD_ENTRIES:
                return self._topk_small(Q, metric, ops.metric_id(metric), k, **kw)
        ms, mi, _q0, _q1 = self.topk_slice(Q, metric, k, **kw)
        if R == 1:
            return ms, mi
        nq = 1 if getattr(Q, "ndim", 2) == 1 else Q.shape[0]
        sl = ms.shape[0]
        out_s = torch.empty((R * sl, k), dtype=torch.float32, device=ms.device)
        out_i = torch.empty((R * sl, k), dtype=torch.int64, device=ms.device)
        dist.all_gather_into_tensor(out_s, ms.contiguous(), group=self.group)
        dist.all_gather_into_tensor(out_i, mi.contiguous(), group=self.group)
        return out_s[:nq], out_i[:nq]


def allpairs_eval(X, category, color, ranges, nbins=1024, thresholds=None, group=None, local_eval=None):
    """BASELINE configs[4] across GPUs: the store is replicated, rank r counts the pairs (i, j > i)
    of its cyclic share of the rows i, and ONE all-reduce of the integer count tensors gives every rank the totals.
    `local_eval(part, nparts) -> (hist, thr_counts)` is injectable for the gloo test; default: ops.allpairs_eval."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    R, r = (dist.get_world_size(group), dist.get_rank(group)) if multi else (1, 0)
    if local_eval is None:
        hist, thr = ops.allpairs_eval(X, category, color, ranges, nbins, thresholds, part=r, nparts=R)
    else:
        hist, thr = local_eval(r, R)
    if multi:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(thr, op=dist.ReduceOp.SUM, group=group)
    return hist, thr
