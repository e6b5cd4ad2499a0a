import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from image_retrieval_b200 import ops
torch.manual_seed(0)
X = torch.randn(3000, 512, device="cuda").bfloat16(); Q = torch.randn(200, 512, device="cuda").bfloat16()
for m in ("cosine_similarity", "l2"):
    s, i = ops.topk(Q, X, m, 100)
Xs = torch.randn(3000, 512, device="cuda"); Qs = torch.randn(64, 512, device="cuda")          # fp32 store: split tensor path
for m in ("cosine_similarity", "l2"):
    s, i = ops.topk(Qs, ops.prepare_index(Xs), m, 10)
Xw = torch.randn(2100, 768, device="cuda").bfloat16(); Qw = torch.randn(40, 768, device="cuda").bfloat16()   # A-streamed variant
s, i = ops.topk(Qw, Xw, "cosine_similarity", 10)
S, I = ops.topk_multi(Qs[:3], Xs, ["cosine_similarity", "l1", "l2"], 5)
s, i = ops.topk(Qs[:2], Xs, "l1", 300)                                                       # paged
Xf = torch.randn(3001, 100, device="cuda"); Qf = torch.randn(13, 100, device="cuda")
for m in ("l1", "linf", "l2", "cosine_similarity", "optimized_similarity"):
    s, i = ops.topk(Qf, Xf, m, 10, params={"w_l1": 0.5} if m.startswith("opt") else None)
Xa = torch.randn(5000, 512, device="cuda"); Qa = torch.randn(8, 512, device="cuda")
s, i = ops.topk(Qa, Xa, "l1", 10)
imgs = torch.randint(0, 256, (5, 64, 64, 3), device="cuda", dtype=torch.uint8)
ops.histogram(imgs, "rgb"); ops.histogram(imgs, "hsv")
ops.allpairs_eval(Xf[:300], np.arange(300) % 5, np.arange(300) % 3, {k: (0.0, 4.0) for k in ops.EVAL_METRICS}, 64)
torch.cuda.synchronize(); print("sanity ok")
