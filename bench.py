#!/usr/bin/env python
"""bench.py - headline benchmark of the retrieval hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (one JSON line on rank 0)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path, same metric
    python bench.py --workload l1_scan|linf_scan|l2_fp32|cos_tensor_fp32|histogram|config1|config5|... # one s8(d) row

The default line also carries `side` (the other s8(d) rows, N = 1 only, each timed for >= 0.25 s so the clock sampler
sees it), `strong` (BASELINE configs[3]: a 10M x 512 store row-sharded over the N ranks, cosine + L1) and `parity`
(at N > 1 against a distributed oracle: every rank's NumPy top-k of its own shard, merged on rank 0).

Headline workload (BASELINE.json configs[1]): cosine top-100, 10k-query bf16 batch against a
1M x 512 bf16 database per GPU.  One "step" = one pass of the hot path over one query batch.
With N > 1 GPUs the database is row-sharded (N x 1M rows, weak scaling: per-GPU shard fixed), each
rank runs the fused scan on its shard and ONE all-gather + merge produces the global top-100 on
every rank; `value` is queries/s normalised to a 1M-row database (queries x total_rows / 1M / s),
which equals plain queries/s at N = 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS_PER_GPU = 1_000_000
DIM = 512
NQ = 10_000
TOPK = 100
PEAKS_FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        p["_source"] = "measured (MEASURED_PEAKS.json)"
        return p
    except Exception:  # noqa: BLE001
        p = dict(PEAKS_FALLBACK)
        p["_source"] = "fallback (B200_PROFILING.md)"
        return p


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (NVML, else nvidia-smi)."""

    def __init__(self, index=0, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(pynvml, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = get_reasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception:  # noqa: BLE001
            self._smi()

    def _smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(int(out[0]))
                self.max_mhz = int(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------ data
def make_unit_rows(torch, n, d, seed, device, dtype, chunk=250_000):
    """Row-normalised N(0,1) rows, generated on the device in fixed chunks (seed = base + chunk id)."""
    out = torch.empty((n, d), dtype=dtype, device=device)
    for c, s in enumerate(range(0, n, chunk)):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000 + c)
        e = min(n, s + chunk)
        x = torch.randn((e - s, d), generator=g, device=device, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        out[s:e] = x.to(dtype)
    return out


def init_nccl(torch, dist, dev):
    """init_process_group + first collective with stdout pointed at stderr: NCCL prints its version banner on stdout
    when NCCL_DEBUG >= VERSION, and stdout must carry exactly one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        t = torch.zeros(1, device=dev)
        dist.all_reduce(t)
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------ CPU legs
def cpu_port_baseline(Qh, Xh, k, budget_s=6.0):
    """Vectorised NumPy port (oracle.search.topk_search: sgemm + FULL stable argsort, the reference's list.sort semantics)
    on a bounded sample of the query batch against the FULL database, all host cores via BLAS.  Also returns the
    (values, indices) of the first 16 queries: the parity spot-check of the timed GPU result."""
    import numpy as np
    from oracle import search as OS
    nq = 16
    t0 = time.perf_counter()
    first_v, first_i = OS.topk_search(Qh[:nq], Xh, "cosine_similarity", k, dtype=np.float32)
    dt = time.perf_counter() - t0
    done = nq
    if dt < budget_s / 3:
        nq2 = int(min(len(Qh) - nq, max(16, nq * (budget_s - dt) / max(dt, 1e-3) * 0.8)))
        t1 = time.perf_counter()
        OS.topk_search(Qh[nq:nq + nq2], Xh, "cosine_similarity", k, dtype=np.float32)
        dt = time.perf_counter() - t1
        done = nq2
    return {"value": done / dt, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{done} of {len(Qh)} queries vs the full {Xh.shape[0]}x{Xh.shape[1]} fp32 database, "
                      f"NumPy sgemm + stable argsort (oracle.search.topk_search), {dt:.1f} s"}, first_v, first_i


def cpu_torch_baseline(Qh, Xh, k, budget_s=12.0):
    """Best-effort vectorised CPU path of BASELINE.md section 3.2: row-normalised torch.mm + torch.topk on all host
    cores, 64-query batches against the FULL database (row norms precomputed once, like our prepared index)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    X = torch.from_numpy(Xh)
    Q = torch.from_numpy(Qh)
    rx = 1.0 / X.norm(dim=1).clamp_min(1e-30)
    nq = 64

    def batch(b):
        q = Q[b * nq % len(Q):][:nq]
        sc = torch.mm(q, X.t())
        sc *= (1.0 / q.norm(dim=1).clamp_min(1e-30))[:, None]
        sc *= rx[None, :]
        return torch.topk(sc, k, dim=1)
    batch(0)
    t0 = time.perf_counter()
    n = 0
    while True:
        batch(n)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 40:
            break
    return {"value": n * nq / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} batches of {nq} queries vs the full {Xh.shape[0]}x{Xh.shape[1]} fp32 database, torch.mm + "
                      f"torch.topk (BASELINE.md 3.2), {torch.get_num_threads()} threads, {dt:.1f} s"}


def _ref_loop_worker(args):
    """The reference's scan (app_pipeline.py:156-172): per-row np.dot / norms, list.sort, slice."""
    import numpy as np
    q, X, k = args
    results = []
    for j in range(X.shape[0]):
        e = X[j]
        sim = np.dot(q, e) / (np.linalg.norm(q) * np.linalg.norm(e))
        results.append({"path": j, "score": abs(sim)})
    results.sort(key=lambda x: x["score"], reverse=True)
    return [r["path"] for r in results[:k]]


_G = {}


def _ref_loop_worker_idx(i):
    return _ref_loop_worker((_G["Q"][i], _G["X"], _G["k"]))


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (Python pair loop of
    app_pipeline.py:156-172, restated in oracle/ because app_pipeline.py itself cannot be imported:
    matplotlib / CLIP missing), one query per host core per step, against a bounded row sample of the
    database; value is converted to the headline unit (queries/s on a 1M x 512 DB) linearly in rows."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import multiprocessing as mp
    import numpy as np
    from oracle import metrics as OM
    cores = os.cpu_count() or 1
    rows = 20_000
    rng = np.random.default_rng(2001)
    X = rng.standard_normal((rows, DIM), dtype=np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X = OM.bf16_round(X)
    Q = rng.standard_normal((cores, DIM), dtype=np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    Q = OM.bf16_round(Q)
    _G.update(Q=Q, X=X, k=TOPK)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_loop_worker_idx, range(cores))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_loop_worker_idx, range(cores))
        dt = (time.perf_counter() - t0) / args.steps
    scale = rows / ROWS_PER_GPU
    value = cores / dt * scale
    sample = (f"{cores} queries/step (one per core, {cores} processes) x {rows} rows of the {ROWS_PER_GPU}-row database; "
              f"queries/s scaled linearly in rows (x{scale:g}); reference loop app_pipeline.py:156-172")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


METRIC = "queries/sec (cosine top-100, 1M x 512 bf16 DB per GPU, 10k-query batch)"


def headline_config(n_gpus):
    return {"workload": "configs[1]: cosine top-100, 10k-query bf16 batch vs 1Mx512 bf16 DB (row-sharded 1M rows/GPU)",
            "db_rows_per_gpu": ROWS_PER_GPU, "db_rows_total": ROWS_PER_GPU * n_gpus, "dim": DIM, "queries": NQ, "k": TOPK,
            "parallelism": f"row-shard x{n_gpus} + all-to-all of query slices + merge + all-gather" if n_gpus > 1 else "single GPU",
            "value_definition": "queries x (db_rows_total / 1M) / s", "l2_policy": "inputs (1 GB shard) larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------ our arm
def roofline_tensor(pk, kern_name, kern_ms, flops, kernels_ms=None, traffic=None):
    """Tensor-pipe roofline of one launch: `frac` is against the BURST cuBLAS figure (a 10-20 ms kernel inside a
    sub-second run is in the burst regime: clocks near max, power below the cap); `frac_sustained` against the
    seconds-long figure, both from MEASURED_PEAKS.json."""
    burst = pk.get("bf16_tflops", PEAKS_FALLBACK["bf16_tflops"])
    sust = pk.get("bf16_tflops_sustained", PEAKS_FALLBACK["bf16_tflops_sustained"])
    ach = flops / (kern_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": kern_name, "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst,
            "frac_burst": ach / burst, "frac_sustained": ach / sust, "peak_sustained": sust, "traffic": traffic,
            "kernel_ms": kern_ms, "kernels_ms": kernels_ms, "algorithmic_flops_per_launch": flops,
            "peak_source": pk["_source"] + ": bf16_tflops (burst) for `frac`, bf16_tflops_sustained for `frac_sustained`"}


def read_kernel_ms(lib, tags):
    import ctypes
    out = {}
    for tag, name in tags:
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        lib.b200ir_profile_read(tag, ctypes.byref(ms), ctypes.byref(n))
        if n.value:
            out[name] = ms.value / n.value
    return out


def drain_profile(lib):
    """Drop the events of every kernel class (a previous workload may have recorded classes it never read)."""
    import ctypes
    for tag in range(16):
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        if lib.b200ir_profile_read(tag, ctypes.byref(ms), ctypes.byref(n)) != 0:
            break


KERNEL_TAGS = ((2, "gemm_topk(tcgen05)"), (1, "scan_topk(cuda-core)"), (3, "finalize"), (4, "rerank"), (0, "prep"), (5, "merge"))


def distributed_oracle_parity(np, torch, dist, world, rank, dev, Q, X, row_begin, got_i, k, nq=16):
    """Parity of the sharded result at any N: every rank runs the NumPy port (oracle) for the first `nq` queries on
    ITS shard, the per-shard lists are gathered and merged on the host with the reference's (score, index) order, and
    rank 0 compares the merged oracle with the GPU result of the timed step."""
    from oracle import search as OS
    v, i = OS.topk_search(Q[:nq].float().cpu().numpy(), X.float().cpu().numpy(), "cosine_similarity", k, dtype=np.float32)
    i = i + row_begin
    if world > 1:
        tv = torch.from_numpy(np.ascontiguousarray(v)).to(dev)
        ti = torch.from_numpy(np.ascontiguousarray(i)).to(dev)
        gv = [torch.empty_like(tv) for _ in range(world)]
        gi = [torch.empty_like(ti) for _ in range(world)]
        dist.all_gather(gv, tv)
        dist.all_gather(gi, ti)
        v, i = OS.merge_topk(np.stack([t.cpu().numpy() for t in gv]), np.stack([t.cpu().numpy() for t in gi]), k, True)
    if rank != 0:
        return None
    got = got_i[:nq].cpu().numpy()
    return {"queries_checked": int(nq), "shards": world,
            "index_sets_equal": bool(all(set(a_) == set(b_) for a_, b_ in zip(got, i))),
            "ranks_equal_frac": float((got == i).mean()),
            "note": "GPU (tcgen05 candidates + exact fp32 re-rank + certificate) vs the NumPy fp32 port run per shard and merged "
                    "with (score, index) order; rank swaps only between fp32-equal scores"}


def strong_config4(args, torch, dist, world, rank, local, dev):
    """BASELINE configs[3] as a STRONG-scaling record: a 10M x 512 store row-sharded over the ranks (total fixed),
    one all-gather + merge per search.  Cosine top-100 of 10k bf16 queries (tensor path) and L1 top-10 of 8 fp32
    queries (HBM scan).  Max over ranks, CUDA events, clocks sampled during each timed loop."""
    from image_retrieval_b200.sharded import ShardedIndex, shard_range
    total = args.strong_rows
    b, e = shard_range(total, world, rank)
    n_local = e - b
    res = {"db_rows_total": total, "rows_per_gpu": n_local, "scaling": "strong"}

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), clk.summary()

    Xb = make_unit_rows(torch, n_local, DIM, 4001 + rank, dev, torch.bfloat16)
    Qb = make_unit_rows(torch, NQ, DIM, 4002, dev, torch.bfloat16)
    idx = ShardedIndex(Xb, b)
    ms, clocks = timed(lambda: idx.topk(Qb, "cosine_similarity", TOPK), max(5, int(300 * world / 100)))
    res["cosine_top100_bf16"] = {"queries": NQ, "ms": ms, "queries_per_s": NQ / (ms * 1e-3), "clocks": clocks,
                                 "tflops_per_gpu": 2.0 * NQ * n_local * DIM / (ms * 1e-3) / 1e12}
    del Xb, idx
    torch.cuda.empty_cache()
    Xf = make_unit_rows(torch, n_local, DIM, 4001 + rank, dev, torch.float32)
    Qf = make_unit_rows(torch, 8, DIM, 4003, dev, torch.float32)
    idx = ShardedIndex(Xf, b, prepare=False)
    ms, clocks = timed(lambda: idx.topk(Qf, "l1", 10), max(20, int(60 * world)))
    res["l1_top10_fp32"] = {"queries": 8, "ms": ms, "queries_per_s": 8 / (ms * 1e-3), "clocks": clocks,
                            "hbm_GBps_per_gpu": n_local * DIM * 4 / (ms * 1e-3) / 1e9}
    del Xf, idx
    torch.cuda.empty_cache()
    return res


SIDE_ROWS = (("l1_scan", dict(dim=512, queries=1)), ("l1_scan", dict(dim=512, queries=8)), ("linf_scan", dict(dim=512, queries=8)),
             ("l1_scan", dict(dim=2048, queries=8)), ("l1_scan", dict(dim=512, queries=1024)), ("cos_tensor_fp32", {}), ("l2_tensor_fp32", {}),
             ("histogram", {}), ("histogram", dict(hsv=True)), ("resize", {}), ("config1", {}), ("config5", {}), ("pairs", {}))


def run_headline(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from image_retrieval_b200 import _lib, ops
    from image_retrieval_b200.sharded import ShardedIndex, query_slice

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(torch, dist, dev)
    lib = _lib.load()
    ops.device()

    X = make_unit_rows(torch, ROWS_PER_GPU, DIM, 2001 + rank, dev, torch.bfloat16)
    Q = make_unit_rows(torch, NQ, DIM, 2002, dev, torch.bfloat16)
    Q_host = Q.cpu().pin_memory()
    # end to end every rank hands ITS slice of the query batch to its host consumer (the ranks hold identical results
    # after the merge): together the R ranks deliver each result row to the host exactly once
    q0, q1 = query_slice(NQ, world, rank)
    out_s_host = torch.empty((q1 - q0, TOPK), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((q1 - q0, TOPK), dtype=torch.int64).pin_memory()
    index = ShardedIndex(X, rank * ROWS_PER_GPU)          # per-store search state (row norms) is built here, once
    flags = ops.FLAG_NO_TENSOR if args.no_tensor else 0

    def step():
        return index.topk(Q, "cosine_similarity", TOPK, flags=flags)

    def step_e2e():
        # the call a user makes, host buffers on both sides: pinned queries -> device, search, results -> pinned host
        q = Q_host.to(dev, non_blocking=True)
        s, i, a0, a1 = index.topk_slice(q, "cosine_similarity", TOPK, flags=flags)     # this rank's slice of the global result
        out_s_host.copy_(s[:a1 - a0], non_blocking=True)
        out_i_host.copy_(i[:a1 - a0], non_blocking=True)
        torch.cuda.current_stream().synchronize()          # results are usable on the host after every step
        return out_s_host, out_i_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, out

    for _ in range(max(args.warmup, 3)):
        step()
    step_e2e()
    torch.cuda.synchronize()

    launches0 = lib.b200ir_launch_count()
    with ClockSampler(local) as clk:
        ms_step, (s_dev, i_dev) = timed(step, args.steps)
    launches = lib.b200ir_launch_count() - launches0
    fallback_queries = ops.last_fallback_count()
    ms_e2e, _ = timed(step_e2e, max(2, min(args.steps, 5)))

    # dominant-kernel duration on its launching stream (separate pass; events perturb nothing else)
    drain_profile(lib)
    lib.b200ir_profile_enable(1)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    kern = read_kernel_ms(lib, KERNEL_TAGS)
    lib.b200ir_profile_enable(0)

    # parity of the timed result, at every N (the oracle runs here and in the CPU leg only)
    parity = None
    if not args.no_cpu:
        parity = distributed_oracle_parity(np, torch, dist, world, rank, dev, Q, X, rank * ROWS_PER_GPU, i_dev, TOPK)
    strong = None
    if not args.no_strong:
        del index, X
        torch.cuda.empty_cache()
        strong = strong_config4(args, torch, dist, world, rank, local, dev)
        X = make_unit_rows(torch, ROWS_PER_GPU, DIM, 2001 + rank, dev, torch.bfloat16)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    pk = peaks()
    dom = max(kern, key=kern.get) if kern else None
    flops = 2.0 * NQ * ROWS_PER_GPU * DIM
    roofline = None
    if dom:
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tr = json.load(f).get(dom)
            if tr:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]      # measured under ncu, per launch
        except Exception:  # noqa: BLE001
            pass
        roofline = roofline_tensor(pk, dom, kern[dom], flops, kern, traffic)

    # CPU legs: the best-effort vectorised path (torch.mm + topk) is the reported cpu_baseline; the NumPy
    # port with the reference's full stable sort rides along
    cpu, cpu_port = None, None
    if not args.no_cpu:
        Xh, Qh = X.float().cpu().numpy(), Q.float().cpu().numpy()
        cpu = cpu_torch_baseline(Qh, Xh, TOPK)
        cpu_port, _v, _i = cpu_port_baseline(Qh, Xh, TOPK)
    side = None
    if world == 1 and not args.no_side:
        del X
        torch.cuda.empty_cache()
        side = {}
        for w, kw in SIDE_ROWS:
            a2 = argparse.Namespace(**{**vars(args), "workload": w, "no_cpu": True, "rows": 0, "dim": 0, "queries": 0, "k": 0,
                                       "hsv": False, **kw})
            try:
                r = measure_side(a2)
                key = w + "".join(f"_{k_}{v_}" for k_, v_ in kw.items())
                side[key] = {k_: r[k_] for k_ in ("metric", "value", "unit", "ms_per_step", "steps", "gpu_launches", "clocks", "roofline",
                                                 "config", "dtype", "kernels_ms", "fallback_queries_per_step") if k_ in r}
            except Exception as ex:  # noqa: BLE001
                side[w] = {"error": repr(ex)[:300]}
            torch.cuda.empty_cache()
    total_rows = ROWS_PER_GPU * world
    scale = total_rows / 1e6
    h2d = NQ * DIM * 2
    d2h = (q1 - q0) * TOPK * (4 + 8)
    line = {
        "metric": METRIC, "value": NQ * scale / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 inputs, fp32 accumulate", "data": "synthetic (row-normalised N(0,1), seeded, generated on device)",
        "config": headline_config(world),
        "e2e": {"value": NQ * scale / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "note": "per rank: pinned host query batch -> device, fused scan (+ all-to-all + merge of this rank's 1/N slice "
                                               "of the queries), that slice's (scores, ids) -> pinned host: together the ranks deliver every "
                                               "result row to the host once; database resident in HBM"},
        "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_port": cpu_port,
        "parity": parity, "fallback_queries_per_step": fallback_queries, "strong": strong, "side": side,
        "path": "cuda-core scan" if (args.no_tensor or "gemm_topk(tcgen05)" not in kern) else "tcgen05 gemm + fused top-k",
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------ other rows
def side_cpu_baseline(kind, unit, fn, units_per_call, sample, budget_s=6.0):
    """Bounded CPU leg of a side workload: the oracle (NumPy port of the reference arithmetic) on a small sample of the
    same synthetic input, repeated until ~budget_s of CPU time is spent."""
    fn()
    t0 = time.perf_counter()
    n = 0
    while True:
        fn()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 50:
            break
    return {"value": n * units_per_call / dt, "unit": unit, "cores": 1 if kind == "loop" else os.cpu_count(), "kind": "port",
            "sample": f"{sample}, {n} repeats, {dt:.1f} s"}


def run_side(args):
    print(json.dumps(measure_side(args)))
    return 0


def measure_side(args):
    """Secondary s8(d) workloads (single GPU): HBM-bound scans, the fp32 tensor path, histograms, evaluation.  Returns
    the JSON line as a dict.  Every timed loop runs for at least ~0.25 s so that the 20 ms clock sampler sees it."""
    import numpy as np
    import torch
    from image_retrieval_b200 import _lib, ops
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    pk = peaks()
    w = args.workload

    def time_fn(fn, tags):
        # calibrate: one call, then enough warm-ups / steps for a >= 0.25 s timed loop (>= 5 clock samples at 20 ms)
        fn()
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); fn(); c1.record(); torch.cuda.synchronize()
        est = max(c0.elapsed_time(c1), 1e-3)
        warm = max(1, min(max(args.warmup, 3), int(300 / est)))
        steps = int(max(1, min(5000, max(args.steps if est < 50 else 1, 250 / est))))
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        n0 = lib.b200ir_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(0) as clk:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = (lib.b200ir_launch_count() - n0) / steps
        drain_profile(lib)
        lib.b200ir_profile_enable(1)
        for _ in range(3 if est < 100 else 1):
            fn()
        torch.cuda.synchronize()
        kms = read_kernel_ms(lib, tags)
        lib.b200ir_profile_enable(0)
        return ms, kms, launches, clk.summary(), steps

    def hbm_roofline(kernel, kms, bytes_alg):
        ach = bytes_alg / (kms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "frac_of_nominal_8TBs": ach / 8000.0, "traffic": None, "kernel_ms": kms, "algorithmic_bytes_per_launch": bytes_alg,
                "peak_source": pk["_source"] + ": hbm_gbs (copy bandwidth, read + write)"}

    if w in ("l1_scan", "linf_scan", "l2_fp32", "cos_fp32"):
        D = args.dim or 2048
        N = args.rows or 1_000_000
        nq = args.queries or 8
        k = args.k or 10
        metric = {"l1_scan": "l1", "linf_scan": "linf", "l2_fp32": "l2", "cos_fp32": "cosine_similarity"}[w]
        g = torch.Generator(device=dev); g.manual_seed(3001)
        X = torch.relu(torch.randn((N, D), generator=g, device=dev))
        Q = torch.relu(torch.randn((nq, D), generator=g, device=dev))
        fl = ops.FLAG_NO_TENSOR if w in ("l2_fp32", "cos_fp32") else 0
        ms, kern, launches, clocks, steps = time_fn(lambda: ops.topk(Q, X, metric, k, flags=fl), ((1, "scan"),))
        kms = kern.get("scan", ms)
        cpu = None
        if not args.no_cpu:
            from oracle import search as OS
            ns = min(N, 100_000)
            Xs, Qs = X[:ns].cpu().numpy(), Q[:64].cpu().numpy()
            cpu = side_cpu_baseline("blas", "queries/s", lambda: OS.topk_search(Qs, Xs, metric, k, dtype=np.float32),
                                    len(Qs) * ns / N, f"{len(Qs)} queries vs the first {ns} rows (NumPy port, scaled linearly to {N} rows)")
        roof = hbm_roofline("scan_topk", kms, N * D * 4)
        if nq > 16:
            # batch regime: more queries than one pass holds -> FP32-issue bound (2 lane-instructions per element pair)
            lane_ops = 2.0 * nq * N * D
            peak = 148 * 128 * 1.965e9
            roof = {"bound": "fp32-alu", "kernel": "scan_topk", "achieved": lane_ops / (kms * 1e-3) / 1e12, "peak": peak / 1e12,
                    "unit": "T lane-instr/s", "frac": lane_ops / (kms * 1e-3) / peak, "kernel_ms": kms, "traffic": None,
                    "note": "batch regime (SURVEY 8d): 2 fp32 lane-instructions per element pair; peak = 148 SMs x 128 lanes x 1.965 GHz"}
        return {"metric": f"queries/sec ({metric} top-{k}, {N}x{D} fp32 DB, {nq}-query batch)", "value": nq / (ms * 1e-3),
                "unit": "queries/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "dtype": "f32", "data": "synthetic relu(N(0,1))",
                "config": {"workload": f"{w}: {metric} top-{k}, {N}x{D} fp32, Q={nq}", "l2_policy": "inputs larger than L2"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    if w in ("cos_tensor_fp32", "l2_tensor_fp32"):
        # fp32 store on the tcgen05 path: three-term bf16 split + exact fp32 re-rank + certificate (prepared index)
        D, N, nq, k = args.dim or DIM, args.rows or ROWS_PER_GPU, args.queries or NQ, args.k or TOPK
        metric = "cosine_similarity" if w == "cos_tensor_fp32" else "l2"
        X = make_unit_rows(torch, N, D, 2101, dev, torch.float32)
        Q = make_unit_rows(torch, nq, D, 2102, dev, torch.float32)
        idx = ops.prepare_index(X)
        ms, kern, launches, clocks, steps = time_fn(lambda: ops.topk(Q, idx, metric, k), KERNEL_TAGS)
        fb = ops.last_fallback_count()
        kms = kern.get("gemm_topk(tcgen05)", ms)
        cpu = None
        if not args.no_cpu:
            cpu = cpu_torch_baseline(Q.cpu().numpy(), X.cpu().numpy(), k, budget_s=6.0)
        roof = roofline_tensor(pk, "gemm_topk(tcgen05, 3-term bf16 split)", kms, 2.0 * nq * N * D, kern)
        roof["tensor_flops_issued_per_launch"] = 3 * roof["algorithmic_flops_per_launch"]
        roof["frac_of_issued_flops_burst"] = 3 * roof["frac_burst"]
        return {"metric": f"queries/sec ({metric} top-{k}, {N}x{D} fp32 DB, {nq}-query batch)", "value": nq / (ms * 1e-3),
                "unit": "queries/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "dtype": "f32 (bf16 hi+lo split on tensor cores, exact fp32 re-rank)",
                "data": "synthetic row-normalised N(0,1)", "fallback_queries_per_step": fb,
                "config": {"workload": f"{w}: {metric} top-{k}, {N}x{D} fp32, Q={nq}", "l2_policy": "inputs larger than L2"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    if w == "histogram":
        B = args.rows or 8192
        g = torch.Generator(device=dev); g.manual_seed(1001)
        imgs = torch.randint(0, 256, (B, 224, 224, 3), generator=g, device=dev, dtype=torch.uint8)
        cs = "hsv" if args.hsv else "rgb"
        ms, kern, launches, clocks, steps = time_fn(lambda: ops.histogram(imgs, cs), ((6, "hist"),))
        kms = kern.get("hist", ms)
        cpu = None
        if not args.no_cpu:
            from oracle import histogram as OH
            hs = imgs[:64].cpu().numpy()
            cpu = side_cpu_baseline("numpy", "images/s", lambda: OH.histogram(hs, cs), 64, f"64 images, NumPy bincount port ({cs})")
        return {"metric": f"images/sec (512-bin {cs} histogram, 224x224x3 uint8)", "value": B / (ms * 1e-3), "unit": "images/s",
                "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "dtype": "u8", "data": "synthetic uniform pixels",
                "config": {"workload": f"histogram {cs}: {B} images 224x224x3", "l2_policy": "inputs larger than L2"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": hbm_roofline("histogram", kms, B * (224 * 224 * 3 + 512 * 4)), "cpu_baseline": cpu}
    if w == "config5":
        # BASELINE configs[4]: all-pairs evaluation (5 metrics, 4 relationship types, density histograms + PR counts)
        N = args.rows or 100_000
        D = args.dim or 512
        g = torch.Generator(device=dev); g.manual_seed(5001)
        X = torch.randn((N, D), generator=g, device=dev)
        ids = torch.arange(N, device=dev) % 30
        cat, col = (ids // 3).int(), (ids % 3).int()           # 10 categories x 3 colours (imageProcessing.py:60-62)
        ranges = {"cosine_distance": (0.0, 2.0), "l1_distance": (0.0, 2.0), "l2_distance": (0.0, 2.5),
                  "linf_distance": (0.0, 8.0), "magnitude_difference": (0.0, 6.0)}
        ms, kern, launches, clocks, steps = time_fn(lambda: ops.allpairs_eval(X, cat, col, ranges, 1024), ((1, "scan"),))
        kms = kern.get("scan", ms)
        pairs = N * (N - 1) / 2
        cpu = None
        if not args.no_cpu:
            from oracle import evaluation as E
            ns = 1500
            Xs, cs_, ks_ = X[:ns].cpu().numpy(), cat[:ns].cpu().numpy(), col[:ns].cpu().numpy()
            thr_np = np.linspace(0, 1, 100)
            cpu = side_cpu_baseline("blas", "pairs/s",
                                    lambda: E.bin_counts(E.metric_matrices(Xs, np.float32), E.relationship(cs_, ks_), ranges, 1024, thr_np),
                                    ns * (ns - 1) / 2, f"all pairs of the first {ns} rows, vectorised NumPy port of the five metrics + binning")
        lane_ops = pairs * D * 5.25                              # dot, |d|, d^2, max|d| + shared x^2: instructions per element pair
        peak = 148 * 128 * 1.965e9
        return {"metric": f"pairs/sec (all-pairs evaluation, {N}x{N}, D={D}, 5 metrics)", "value": pairs / (ms * 1e-3), "unit": "pairs/s",
                "n_gpus": 1, "steps": steps, "ms_per_step": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic N(0,1)",
                "config": {"workload": "configs[4]: all-pairs distance-density + precision-recall counts, five metrics"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "fp32-alu", "kernel": "scan_topk<K_EVAL>", "achieved": lane_ops / (kms * 1e-3) / 1e12,
                             "peak": peak / 1e12, "unit": "T lane-instr/s", "frac": lane_ops / (kms * 1e-3) / peak, "kernel_ms": kms,
                             "traffic": None,
                             "note": "CUDA-core bound: 5.25 fp32 lane-instructions per element pair; peak = 148 SMs x 128 lanes x 1.965 GHz"},
                "cpu_baseline": cpu}
    if w == "resize":
        # image front-end: PIL-exact bicubic resize (shorter edge -> 224) + centre crop, then the 512-bin histogram
        B = args.rows or 2048
        H, W = 480, 640
        g = torch.Generator(device=dev); g.manual_seed(1003)
        imgs = torch.randint(0, 256, (B, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
        ms, kern, launches, clocks, steps = time_fn(lambda: ops.histogram(ops.resize_crop(imgs, 224)), ((8, "resize"),))
        kms = kern.get("resize", ms)
        rh, rw = ops.shortest_edge_size(H, W, 224)
        scale = W / rw
        left = (rw - 224) // 2
        cols = min(W, int((left + 224) * scale + 2 * scale + 1)) - max(0, int(left * scale - 2 * scale))
        bytes_alg = B * (H * cols * 3 + 224 * 224 * 3)            # source window the crop depends on + cropped output
        cpu = None
        if not args.no_cpu:
            try:
                from PIL import Image
                from oracle import histogram as OH
                hs = imgs[:32].cpu().numpy()

                def pil_front_end():
                    out = []
                    for im in hs:
                        r = np.asarray(Image.fromarray(im).resize((rw, rh), resample=Image.BICUBIC))
                        out.append(r[(rh - 224) // 2:(rh - 224) // 2 + 224, left:left + 224])
                    return OH.histogram(np.stack(out), "rgb")
                cpu = side_cpu_baseline("loop", "images/s", pil_front_end, 32, "32 images, PIL resize(BICUBIC) + crop (the reference's "
                                        "processor front-end) + NumPy histogram, one core")
            except ImportError:
                cpu = None
        return {"metric": f"images/sec (bicubic resize {H}x{W} -> 224 crop + 512-bin histogram)", "value": B / (ms * 1e-3),
                "unit": "images/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "dtype": "u8", "data": "synthetic uniform pixels",
                "config": {"workload": f"resize: {B} images {H}x{W}x3 -> 224x224x3 -> histogram", "l2_policy": "inputs larger than L2"},
                "gpu_launches": launches, "clocks": clocks, "roofline": hbm_roofline("resize_crop", kms, bytes_alg), "cpu_baseline": cpu}
    if w == "pairs":
        # explicit pair lists (mi_analysis.py:256-297): P random pairs over an N x D fp32 store, seven values per pair
        N = args.rows or 1_000_000
        D = args.dim or 512
        P = args.queries or 4_000_000
        g = torch.Generator(device=dev); g.manual_seed(5101)
        X = torch.randn((N, D), generator=g, device=dev)
        ia = torch.randint(0, N, (P,), generator=g, device=dev)
        ib = torch.randint(0, N, (P,), generator=g, device=dev)
        ms, kern, launches, clocks, steps = time_fn(lambda: ops.pair_metrics(X, None, ia, ib), ((7, "pairs"),))
        kms = kern.get("pairs", ms)
        cpu = None
        if not args.no_cpu:
            from oracle import metrics as OM
            ns = 2000
            Xa, Xb = X[ia[:ns]].cpu().numpy(), X[ib[:ns]].cpu().numpy()
            cpu = side_cpu_baseline("loop", "pairs/s", lambda: [OM.get_all_metrics(a, b) for a, b in zip(Xa, Xb)], ns,
                                    f"{ns} pairs, get_all_metrics per pair (the reference loop mi_analysis.py:277-291), one core")
        return {"metric": f"pairs/sec (get_all_metrics over an explicit pair list, {N}x{D} fp32 store)", "value": P / (ms * 1e-3),
                "unit": "pairs/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "dtype": "f32", "data": "synthetic N(0,1), uniform random pairs",
                "config": {"workload": f"pairs: {P} pairs over {N}x{D} fp32", "l2_policy": "inputs larger than L2"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": hbm_roofline("pair_metrics", kms, P * (2 * D * 4 + 16 + 28)), "cpu_baseline": cpu}
    if w == "config1":
        def palette_images(b, seed):
            """a few flat colour blocks + small noise per image: peaky histograms with many exact ties"""
            g = torch.Generator(device=dev); g.manual_seed(seed)
            blocks = torch.randint(0, 256, (b, 4, 4, 3), generator=g, device=dev)
            img = blocks.repeat_interleave(56, dim=1).repeat_interleave(56, dim=2)
            noise = torch.randint(-8, 9, (b, 224, 224, 3), generator=g, device=dev)
            return (img + noise).clamp_(0, 255).to(torch.uint8)
        qi = palette_images(1000, 1001)
        di = palette_images(10000, 1002)

        def fn():
            X = ops.counts_to_embedding(ops.histogram(di))[0]
            Qm = ops.counts_to_embedding(ops.histogram(qi))[0]
            return ops.topk(Qm, X, "l2", 10)
        ms, kern, launches, clocks, steps = time_fn(fn, KERNEL_TAGS + ((6, "histogram"),))
        fb = ops.last_fallback_count()
        cpu = None
        if not args.no_cpu:
            from oracle import histogram as OH
            from oracle import search as OS
            qh, dh = qi[:100].cpu().numpy(), di[:1000].cpu().numpy()

            def cpu_pass():
                Xc = OH.histogram(dh, "rgb").astype(np.float32)
                Qc = OH.histogram(qh, "rgb").astype(np.float32)
                return OS.topk_search(Qc, Xc, "l2", 10, dtype=np.float32)
            cpu = side_cpu_baseline("numpy", "queries/s", cpu_pass, 100, "100 query + 1000 database images (a tenth of config 1 on "
                                    "both sides, same 10 database images per query: histogram port + NumPy L2 top-10)")
        return {"metric": "queries/sec (config 1: histogram embeddings of 1k+10k images, L2 top-10 over 10kx512 fp32)",
                "value": 1000 / (ms * 1e-3), "unit": "queries/s", "n_gpus": 1, "steps": steps, "ms_per_step": ms,
                "higher_is_better": True, "dtype": "u8 -> f32", "data": "synthetic palette images",
                "config": {"workload": "configs[0]"}, "gpu_launches": launches, "clocks": clocks, "kernels_ms": kern,
                "fallback_queries_per_step": fb, "cpu_baseline": cpu}
    raise SystemExit(f"unknown workload {w}")


def run_config4(args):
    """BASELINE configs[3] on its own: python bench.py --workload config4 [--strong-rows N] (torchrun for N > 1)."""
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(torch, dist, dev)
    res = strong_config4(args, torch, dist, world, rank, local, dev)
    if rank == 0:
        print(json.dumps({"metric": "queries/sec per metric (config 4: 10M x 512 row-sharded, all-gather top-k merge)",
                          "n_gpus": world, **res}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--no-tensor", action="store_true", help="force the CUDA-core scan path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the oracle parity check")
    ap.add_argument("--no-side", action="store_true", help="headline: skip the side rows")
    ap.add_argument("--no-strong", action="store_true", help="headline: skip the config-4 strong-scaling sub-record")
    ap.add_argument("--strong-rows", type=int, default=10_000_000)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--hsv", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "config4":
        return run_config4(args)
    if args.workload != "headline":
        return run_side(args)
    return run_headline(args)


if __name__ == "__main__":
    sys.exit(main())
