"""Row-sharded multi-GPU search: one process per GPU, each rank owns a
contiguous range of database rows, runs the fused distance + top-k locally with its row base as
`index_offset`, and ONE all-gather of the packed (score, index) lists feeds the final merge on
every rank.  No other data-path collective exists: the scan itself is embarrassingly parallel.

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


def _default_local(Q, X, metric, k, index_offset, **kw):
    return ops.topk(Q, X, metric, k, index_offset=index_offset, **kw)


def _default_merge(scores, idx, descending):
    return ops.topk_merge(scores, idx, descending)


def _round_up(a, b):
    return (a + b - 1) // b * b


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

    def topk(self, Q, metric, k, **kw):
        """Global top-k over all shards; every rank returns the same (scores, indices)."""
        m = ops.metric_id(metric)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if multi and self.local_topk is _default_local and self.merge is _default_merge:
            return self._topk_packed(Q, metric, m, k, **kw)
        s, i = self.local_topk(Q, self.X, metric, k, self.row_begin, **kw)
        if not multi:
            return s, i
        R = dist.get_world_size(self.group)
        nq = s.shape[0]
        # one collective: pack fp32 scores (as int32 bits) and int64 ids into one int64 payload
        payload = torch.empty((2, nq, k), dtype=torch.int64, device=s.device)
        payload[0] = s.contiguous().view(torch.int32).to(torch.int64)
        payload[1] = i
        gathered = torch.empty(R * payload.numel(), dtype=torch.int64, device=s.device)
        dist.all_gather_into_tensor(gathered, payload.view(-1), group=self.group)
        gathered = gathered.view(R, 2, nq, k)
        all_s = gathered[:, 0].to(torch.int32).view(torch.float32).contiguous()
        all_i = gathered[:, 1].contiguous()
        return self.merge(all_s, all_i, m in ops.DESCENDING)

    def _topk_packed(self, Q, metric, m, k, **kw):
        """CUDA fast path: the local search writes scores and ids into ONE byte record, a single all-gather moves
        the records, and the merge kernel reads the receive buffer in place (no pack / unpack kernels)."""
        R = dist.get_world_size(self.group)
        Qd = ops.as_device_matrix(Q, dtype=self.X.dtype if isinstance(self.X, (torch.Tensor, ops.PreparedIndex)) else None)
        nq = Qd.shape[0]
        score_bytes = _round_up(nq * k * 4, 16)
        rec = _round_up(score_bytes + nq * k * 8, 16)
        buf = torch.empty(rec, dtype=torch.uint8, device=Qd.device)
        s = buf[:nq * k * 4].view(torch.float32).view(nq, k)
        i = buf[score_bytes:score_bytes + nq * k * 8].view(torch.int64).view(nq, k)
        ops.topk(Qd, self.X, metric, k, index_offset=self.row_begin, out=(s, i), **kw)
        gathered = torch.empty(R * rec, dtype=torch.uint8, device=Qd.device)
        dist.all_gather_into_tensor(gathered, buf, group=self.group)
        return ops.topk_merge_packed(gathered, R, nq, k, score_bytes, m in ops.DESCENDING)
