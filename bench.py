#!/usr/bin/env python
"""bench.py - headline benchmark of the retrieval hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (one JSON line on rank 0)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path, same metric
    python bench.py --workload l1_scan|linf_scan|l2_fp32|histogram|config1 ...   # other s8(d) rows

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
def cpu_port_baseline(Qh, Xh, k, budget_s=20.0):
    """Vectorised NumPy port (oracle.search.topk_search: sgemm + stable argsort) on a bounded sample of the
    query batch against the FULL database, all host cores via BLAS."""
    import numpy as np
    from oracle import search as OS
    nq = 16
    t0 = time.perf_counter()
    _first_v, first_i = OS.topk_search(Qh[:nq], Xh, "cosine_similarity", k, dtype=np.float32)
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
                      f"NumPy sgemm + stable argsort (oracle.search.topk_search), {dt:.1f} s"}, first_i


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
            "parallelism": f"row-shard x{n_gpus} + 1 all-gather + merge" if n_gpus > 1 else "single GPU",
            "value_definition": "queries x (db_rows_total / 1M) / s", "l2_policy": "inputs (1 GB shard) larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------ our arm
def run_headline(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from image_retrieval_b200 import _lib, ops
    from image_retrieval_b200.sharded import ShardedIndex

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
    out_s_host = torch.empty((NQ, TOPK), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((NQ, TOPK), dtype=torch.int64).pin_memory()
    index = ShardedIndex(X, rank * ROWS_PER_GPU)
    flags = ops.FLAG_NO_TENSOR if args.no_tensor else 0

    def step():
        return index.topk(Q, "cosine_similarity", TOPK, flags=flags)

    def step_e2e():
        # the call a user makes, host buffers on both sides: pinned queries -> device, search, results -> pinned host
        q = Q_host.to(dev, non_blocking=True)
        s, i = index.topk(q, "cosine_similarity", TOPK, flags=flags)
        out_s_host.copy_(s, non_blocking=True)
        out_i_host.copy_(i, non_blocking=True)
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
    ms_e2e, _ = timed(step_e2e, max(2, min(args.steps, 5)))

    # dominant-kernel duration on its launching stream (separate pass; events perturb nothing else)
    import ctypes
    lib.b200ir_profile_enable(1)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    kern = {}
    for tag, name in ((2, "gemm_topk(tcgen05)"), (1, "scan_topk(cuda-core)"), (3, "finalize"), (4, "rerank"), (0, "prep"), (5, "merge")):
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        lib.b200ir_profile_read(tag, ctypes.byref(ms), ctypes.byref(n))
        if n.value:
            kern[name] = ms.value / n.value
    lib.b200ir_profile_enable(0)

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
        peak = pk.get("bf16_tflops_sustained", PEAKS_FALLBACK["bf16_tflops_sustained"])
        ach = flops / (kern[dom] * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tr = json.load(f).get(dom)
            if tr:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]      # measured under ncu, per launch
        except Exception:  # noqa: BLE001
            pass
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic, "kernel_ms": kern[dom], "kernels_ms": kern,
                    "algorithmic_flops_per_launch": flops, "peak_source": pk["_source"] + ", sustained bf16 (kernel timed inside back-to-back steps)"}

    # CPU leg (the only place the oracle runs in this arm): the vectorised port is timed on a bounded query sample, and
    # its answer for the first 16 queries doubles as a parity spot-check of the timed GPU result
    cpu, parity = None, None
    if not args.no_cpu:
        cpu, ref_i = cpu_port_baseline(Q.float().cpu().numpy(), X.float().cpu().numpy(), TOPK)
        if world == 1:
            got = i_dev[:len(ref_i)].cpu().numpy()
            parity = {"queries_checked": int(len(ref_i)),
                      "index_sets_equal": bool(all(set(a_) == set(b_) for a_, b_ in zip(got, ref_i))),
                      "ranks_equal_frac": float((got == ref_i).mean()),
                      "note": "GPU (fp32 re-rank) vs NumPy fp32 port; rank swaps only between fp32-equal scores"}
    total_rows = ROWS_PER_GPU * world
    scale = total_rows / 1e6
    h2d = NQ * DIM * 2
    d2h = NQ * TOPK * (4 + 8)
    line = {
        "metric": METRIC, "value": NQ * scale / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 inputs, fp32 accumulate", "data": "synthetic (row-normalised N(0,1), seeded, generated on device)",
        "config": headline_config(world),
        "e2e": {"value": NQ * scale / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "note": "pinned host query batch -> device, fused scan, (scores, ids) -> host; database resident in HBM"},
        "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
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
    """Secondary s8(d) workloads (single GPU): HBM-bound scans, exact fp32 L2, histograms."""
    import ctypes
    import numpy as np
    import torch
    from image_retrieval_b200 import _lib, ops
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    pk = peaks()
    w = args.workload

    def time_fn(fn, tag):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        n0 = lib.b200ir_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(0) as clk:
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        launches = lib.b200ir_launch_count() - n0
        lib.b200ir_profile_enable(1)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        kms, n = ctypes.c_float(0), ctypes.c_int(0)
        lib.b200ir_profile_read(tag, ctypes.byref(kms), ctypes.byref(n))
        lib.b200ir_profile_enable(0)
        return ms, kms.value / max(n.value, 1), launches, clk.summary()

    if w in ("l1_scan", "linf_scan", "l2_fp32", "cos_fp32"):
        D = args.dim or 2048
        N = args.rows or 1_000_000
        nq = args.queries or 8
        k = args.k or 10
        metric = {"l1_scan": "l1", "linf_scan": "linf", "l2_fp32": "l2", "cos_fp32": "cosine_similarity"}[w]
        g = torch.Generator(device=dev); g.manual_seed(3001)
        X = torch.relu(torch.randn((N, D), generator=g, device=dev))
        Q = torch.relu(torch.randn((nq, D), generator=g, device=dev))
        ms, kms, launches, clocks = time_fn(lambda: ops.topk(Q, X, metric, k), 1)
        bytes_alg = N * D * 4
        ach = bytes_alg / (kms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu:
            from oracle import search as OS
            ns = min(N, 100_000)
            Xs, Qs = X[:ns].cpu().numpy(), Q.cpu().numpy()
            cpu = side_cpu_baseline("blas", "queries/s", lambda: OS.topk_search(Qs, Xs, metric, k, dtype=np.float32),
                                    nq * ns / N, f"{nq} queries vs the first {ns} rows (NumPy port, scaled linearly to {N} rows)")
        line = {"metric": f"queries/sec ({metric} top-{k}, {N}x{D} fp32 DB, {nq}-query batch)", "value": nq / (ms * 1e-3),
                "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "dtype": "f32", "data": "synthetic relu(N(0,1))",
                "config": {"workload": f"{w}: {metric} top-{k}, {N}x{D} fp32, Q={nq}", "l2_policy": "inputs larger than L2"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "scan_topk", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": ach / pk["hbm_gbs"], "traffic": None, "kernel_ms": kms,
                             "algorithmic_bytes_per_launch": bytes_alg, "peak_source": pk["_source"]}, "cpu_baseline": cpu}
        print(json.dumps(line))
    elif w == "histogram":
        B = args.rows or 8192
        g = torch.Generator(device=dev); g.manual_seed(1001)
        imgs = torch.randint(0, 256, (B, 224, 224, 3), generator=g, device=dev, dtype=torch.uint8)
        cs = "hsv" if args.hsv else "rgb"
        ms, kms, launches, clocks = time_fn(lambda: ops.histogram(imgs, cs), 6)
        bytes_alg = B * (224 * 224 * 3 + 512 * 4)
        ach = bytes_alg / (kms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu:
            from oracle import histogram as OH
            hs = imgs[:64].cpu().numpy()
            cpu = side_cpu_baseline("numpy", "images/s", lambda: OH.histogram(hs, cs), 64, f"64 images, NumPy bincount port ({cs})")
        line = {"metric": f"images/sec (512-bin {cs} histogram, 224x224x3 uint8)", "value": B / (ms * 1e-3), "unit": "images/s",
                "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "dtype": "u8", "data": "synthetic uniform pixels",
                "config": {"workload": f"histogram {cs}: {B} images 224x224x3", "l2_policy": "inputs larger than L2"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "histogram", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": ach / pk["hbm_gbs"], "traffic": None, "kernel_ms": kms,
                             "algorithmic_bytes_per_launch": bytes_alg, "peak_source": pk["_source"]}, "cpu_baseline": cpu}
        print(json.dumps(line))
    elif w == "config5":
        # BASELINE configs[4]: all-pairs evaluation (5 metrics, 4 relationship types, density histograms + PR counts)
        N = args.rows or 100_000
        D = args.dim or 512
        g = torch.Generator(device=dev); g.manual_seed(5001)
        X = torch.randn((N, D), generator=g, device=dev)
        ids = torch.arange(N, device=dev) % 30
        cat, col = (ids // 3).int(), (ids % 3).int()           # 10 categories x 3 colours (imageProcessing.py:60-62)
        ranges = {"cosine_distance": (0.0, 2.0), "l1_distance": (0.0, 2.0), "l2_distance": (0.0, 2.5),
                  "linf_distance": (0.0, 8.0), "magnitude_difference": (0.0, 6.0)}
        ms, kms, launches, clocks = time_fn(lambda: ops.allpairs_eval(X, cat, col, ranges, 1024), 1)
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
        line = {"metric": f"pairs/sec (all-pairs evaluation, {N}x{N}, D={D}, 5 metrics)", "value": pairs / (ms * 1e-3), "unit": "pairs/s",
                "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic N(0,1)",
                "config": {"workload": "configs[4]: all-pairs distance-density + precision-recall counts, five metrics"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "fp32-alu", "kernel": "scan_topk<K_EVAL>", "achieved": lane_ops / (kms * 1e-3) / 1e12,
                             "peak": peak / 1e12, "unit": "T lane-instr/s", "frac": lane_ops / (kms * 1e-3) / peak, "kernel_ms": kms,
                             "note": "CUDA-core bound: 5.25 fp32 lane-instructions per element pair; peak = 148 SMs x 128 lanes x 1.965 GHz"},
                "cpu_baseline": cpu}
        print(json.dumps(line))
    elif w == "resize":
        # image front-end: PIL-exact bicubic resize (shorter edge -> 224) + centre crop, then the 512-bin histogram
        B = args.rows or 2048
        H, W = 480, 640
        g = torch.Generator(device=dev); g.manual_seed(1003)
        imgs = torch.randint(0, 256, (B, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
        ms, kms, launches, clocks = time_fn(lambda: ops.histogram(ops.resize_crop(imgs, 224)), 8)
        rh, rw = ops.shortest_edge_size(H, W, 224)
        scale = W / rw
        left = (rw - 224) // 2
        cols = min(W, int((left + 224) * scale + 2 * scale + 1)) - max(0, int(left * scale - 2 * scale))
        bytes_alg = B * (H * cols * 3 + 224 * 224 * 3)            # source window the crop depends on + cropped output
        ach = bytes_alg / (kms * 1e-3) / 1e9
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
        line = {"metric": f"images/sec (bicubic resize {H}x{W} -> 224 crop + 512-bin histogram)", "value": B / (ms * 1e-3),
                "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "dtype": "u8", "data": "synthetic uniform pixels",
                "config": {"workload": f"resize: {B} images {H}x{W}x3 -> 224x224x3 -> histogram", "l2_policy": "inputs larger than L2"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "resize_crop", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": ach / pk["hbm_gbs"], "traffic": None, "kernel_ms": kms,
                             "algorithmic_bytes_per_launch": bytes_alg, "peak_source": pk["_source"]}, "cpu_baseline": cpu}
        print(json.dumps(line))
    elif w == "pairs":
        # explicit pair lists (mi_analysis.py:256-297): P random pairs over an N x D fp32 store, seven values per pair
        N = args.rows or 1_000_000
        D = args.dim or 512
        P = args.queries or 4_000_000
        g = torch.Generator(device=dev); g.manual_seed(5101)
        X = torch.randn((N, D), generator=g, device=dev)
        ia = torch.randint(0, N, (P,), generator=g, device=dev)
        ib = torch.randint(0, N, (P,), generator=g, device=dev)
        ms, kms, launches, clocks = time_fn(lambda: ops.pair_metrics(X, None, ia, ib), 7)
        bytes_alg = P * (2 * D * 4 + 16 + 28)
        ach = bytes_alg / (kms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu:
            from oracle import metrics as OM
            ns = 2000
            Xa, Xb = X[ia[:ns]].cpu().numpy(), X[ib[:ns]].cpu().numpy()
            cpu = side_cpu_baseline("loop", "pairs/s", lambda: [OM.get_all_metrics(a, b) for a, b in zip(Xa, Xb)], ns,
                                    f"{ns} pairs, get_all_metrics per pair (the reference loop mi_analysis.py:277-291), one core")
        line = {"metric": f"pairs/sec (get_all_metrics over an explicit pair list, {N}x{D} fp32 store)", "value": P / (ms * 1e-3),
                "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "dtype": "f32", "data": "synthetic N(0,1), uniform random pairs",
                "config": {"workload": f"pairs: {P} pairs over {N}x{D} fp32", "l2_policy": "inputs larger than L2"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "pair_metrics", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": ach / pk["hbm_gbs"], "traffic": None, "kernel_ms": kms,
                             "algorithmic_bytes_per_launch": bytes_alg, "peak_source": pk["_source"]}, "cpu_baseline": cpu}
        print(json.dumps(line))
    elif w == "config1":
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
        ms, kms, launches, clocks = time_fn(fn, 1)
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
        line = {"metric": "queries/sec (config 1: histogram embeddings of 1k+10k images, L2 top-10 over 10kx512 fp32)",
                "value": 1000 / (ms * 1e-3), "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": ms,
                "higher_is_better": True, "dtype": "u8 -> f32", "data": "synthetic palette images",
                "config": {"workload": "configs[0]"}, "gpu_launches": int(launches), "clocks": clocks, "scan_kernel_ms": kms,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    else:
        raise SystemExit(f"unknown workload {w}")
    return 0


def run_config4(args):
    """BASELINE configs[3]: 10M x 512 database row-sharded over the ranks (strong scaling: total rows fixed), all five
    metrics, one all-gather + merge per search.  bf16 rows for the tensor-core metrics (10k queries, k=100), fp32 rows
    for L1 / Linf (8 queries, k=10).  Prints one JSON line with per-metric queries/s (max over ranks)."""
    import torch
    import torch.distributed as dist
    from image_retrieval_b200 import ops
    from image_retrieval_b200.sharded import ShardedIndex, shard_range
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(torch, dist, dev)
    total = args.rows or 10_000_000
    b, e = shard_range(total, world, rank)
    n_local = e - b
    res = {}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    Xb = make_unit_rows(torch, n_local, DIM, 4001 + rank, dev, torch.bfloat16)
    Qb = make_unit_rows(torch, NQ, DIM, 4002, dev, torch.bfloat16)
    idx = ShardedIndex(Xb, b)
    for metric in ("cosine_similarity", "angular_distance", "l2"):
        ms = timed(lambda: idx.topk(Qb, metric, TOPK))
        res[metric] = {"queries": NQ, "k": TOPK, "dtype": "bf16", "ms": ms, "queries_per_s": NQ / (ms * 1e-3)}
    del Xb, idx
    torch.cuda.empty_cache()
    Xf = make_unit_rows(torch, n_local, DIM, 4001 + rank, dev, torch.float32)
    Qf = make_unit_rows(torch, 8, DIM, 4003, dev, torch.float32)
    idx = ShardedIndex(Xf, b)
    for metric in ("l1", "linf"):
        ms = timed(lambda: idx.topk(Qf, metric, 10))
        res[metric] = {"queries": 8, "k": 10, "dtype": "f32", "ms": ms, "queries_per_s": 8 / (ms * 1e-3),
                       "hbm_GBps_per_gpu": n_local * DIM * 4 / (ms * 1e-3) / 1e9}
    if rank == 0:
        print(json.dumps({"metric": "queries/sec per metric (config 4: 10M x 512 row-sharded, all-gather top-k merge)",
                          "n_gpus": world, "steps": args.steps, "db_rows_total": total, "rows_per_gpu": n_local,
                          "scaling": "strong", "results": res}))
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
