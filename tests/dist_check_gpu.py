"""Multi-GPU check (run under torchrun on a GPU box): the row-sharded search over NCCL equals the single-GPU search
exactly, for tensor-core metrics (bf16 and fp32 stores) and scan metrics.  tests/test_gpu_dist.py launches it when the
box has >= 2 GPUs; by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_check_gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_retrieval_b200 import ops  # noqa: E402
from image_retrieval_b200.sharded import ShardedIndex, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    N, D = 200_003, 512
    X = torch.randn((N, D), generator=g, device="cuda")
    Q = torch.randn((300, D), generator=g, device="cuda")
    X[150_000] = X[7]                                   # tie across shards -> lower global index first
    b, e = shard_range(N, world, rank)
    ok = True
    for dtype, metric, k in ((torch.bfloat16, "cosine_similarity", 100), (torch.bfloat16, "l2", 10),
                             (torch.float32, "cosine_similarity", 100), (torch.float32, "l2", 10),      # fp32 store on the tensor path
                             (torch.float32, "l1", 10), (torch.float32, "linf", 7)):
        Xd, Qd = X.to(dtype), Q.to(dtype)
        s1, i1 = ops.topk(Qd, Xd, metric, k)
        s2, i2 = ShardedIndex(Xd[b:e].contiguous(), b).topk(Qd, metric, k)
        same = torch.equal(i1, i2) and torch.equal(s1, s2)
        # 8 queries: the small-payload path (one packed all-gather)
        s3, i3 = ops.topk(Qd[:8], Xd, metric, k)
        s4, i4 = ShardedIndex(Xd[b:e].contiguous(), b).topk(Qd[:8], metric, k)
        same = same and torch.equal(i3, i4) and torch.equal(s3, s4)
        ok = ok and same
        if rank == 0:
            print(f"{metric:18s} {str(dtype):15s} k={k:3d} sharded==single: {same}")
    # all-pairs evaluation (config 5) across GPUs: replicated store, cyclic row shares, one all-reduce of the counts
    from image_retrieval_b200 import sharded
    Xe = X[:4000, :64].contiguous()
    cat, col = (torch.arange(4000) % 10).numpy(), ((torch.arange(4000) // 10) % 3).numpy()
    ranges = {m: (0.0, 4.0) for m in ops.EVAL_METRICS}
    h1, t1 = ops.allpairs_eval(Xe, cat, col, ranges, 256)
    h2, t2 = sharded.allpairs_eval(Xe, cat, col, ranges, 256)
    same = torch.equal(h1, h2) and torch.equal(t1, t2)
    ok = ok and same
    if rank == 0:
        print(f"allpairs_eval sharded==single: {same}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if flag.item() != 1:
        raise SystemExit("MISMATCH")
    if rank == 0:
        print("dist check ok")


if __name__ == "__main__":
    main()
