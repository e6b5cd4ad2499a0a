"""Quick timing of wide-row (D > 512) searches: A-streamed tensor path vs the CUDA-core scan.  python tests/tools/wide_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from image_retrieval_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
for D, N, nq in ((768, 1_000_000, 10_000), (1024, 1_000_000, 10_000), (2048, 500_000, 10_000)):
    X = torch.randn((N, D), generator=g, device=dev); X /= X.norm(dim=1, keepdim=True)
    Q = torch.randn((nq, D), generator=g, device=dev); Q /= Q.norm(dim=1, keepdim=True)
    for name, x, q in (("bf16", X.bfloat16(), Q.bfloat16()), ("fp32", X, Q)):
        idx = ops.prepare_index(x)
        for _ in range(2):
            ops.topk(q, idx, "cosine_similarity", 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.topk(q, idx, "cosine_similarity", 100)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tf = 2.0 * nq * N * D / (ms * 1e-3) / 1e12
        print(f"D={D} N={N} {name}: {ms:8.2f} ms  {nq / ms * 1e3:10.0f} q/s  {tf:7.1f} algorithmic TFLOP/s  fallback={ops.last_fallback_count()}", flush=True)
        del idx
    qs = Q[:64]
    ops.topk(qs, X, "cosine_similarity", 100, flags=ops.FLAG_NO_TENSOR)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.topk(qs, X, "cosine_similarity", 100, flags=ops.FLAG_NO_TENSOR); e1.record(); torch.cuda.synchronize()
    print(f"D={D} N={N} fp32 CUDA-core scan: {64 / e0.elapsed_time(e1) * 1e3:10.0f} q/s", flush=True)
    del X, Q
    torch.cuda.empty_cache()
