"""Timing helper (not a test): tcgen05 top-k kernel time for several k on the headline shape.
usage: python tests/tools/gemm_sweep.py 1 10 100"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from image_retrieval_b200 import _lib, ops  # noqa: E402

ks = [int(a) for a in sys.argv[1:]] or [100]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(2001)
X = torch.nn.functional.normalize(torch.randn((1_000_000, 512), generator=g, device=dev)).bfloat16()
Q = torch.nn.functional.normalize(torch.randn((10_000, 512), generator=g, device=dev)).bfloat16()
lib = _lib.load()
for k in ks:
    for _ in range(3):
        ops.topk(Q, X, "cosine_similarity", k)
    torch.cuda.synchronize()
    lib.b200ir_profile_enable(1)
    for _ in range(5):
        ops.topk(Q, X, "cosine_similarity", k)
    torch.cuda.synchronize()
    out = []
    for tag in (2, 4):
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        lib.b200ir_profile_read(tag, ctypes.byref(ms), ctypes.byref(n))
        out.append(ms.value / max(n.value, 1))
    lib.b200ir_profile_enable(0)
    print(f"k={k} gemm_ms={out[0]:.3f} rerank_ms={out[1]:.3f}", flush=True)
