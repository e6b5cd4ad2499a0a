"""Quick device-resident timing of the fp32 tensor path (three-term bf16 split): python tests/tools/fp32_tensor_bench.py [N] [nq] [k]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from image_retrieval_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
X = torch.randn((N, 512), generator=g, device=dev); X /= X.norm(dim=1, keepdim=True)
Q = torch.randn((nq, 512), generator=g, device=dev); Q /= Q.norm(dim=1, keepdim=True)
idx = ops.prepare_index(X)
for name, target, kw in (("fp32 indexed", idx, {}), ("fp32 no index", X, {}), ("fp32 scan", X, {"flags": ops.FLAG_NO_TENSOR} if nq <= 256 else None),
                         ("bf16 indexed", ops.prepare_index(X.bfloat16()), {})):
    if kw is None:
        continue
    q = Q.bfloat16() if name.startswith("bf16") else Q
    for metric in ("cosine_similarity", "l2"):
        for _ in range(2):
            ops.topk(q, target, metric, k, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.topk(q, target, metric, k, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name:14s} {metric:18s} {ms:8.3f} ms  {nq / ms * 1e3:12.0f} q/s  fallback={ops.last_fallback_count()}", flush=True)
