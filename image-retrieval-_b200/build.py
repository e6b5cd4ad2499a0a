"""Build libb200ir.so (sm_100a) in-tree with nvcc: one object per translation unit, in parallel.

Usage: python image-retrieval-_b200/build.py [--force] [--verbose]
The shared library links the CUDA runtime statically and resolves the driver API
(cuTensorMapEncodeTiled) at run time, so it also LOADS on a machine without a GPU driver.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200ir.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]

KINDS = ["K_L1", "K_L2", "K_LINF", "K_DOT", "K_MULTI", "K_MULTI6"]


def units():
    u = [("b200ir", "b200ir.cu", []), ("scan_run", "scan_run.cu", []), ("gemm_topk", "gemm_topk.cu", [])]
    for kind in KINDS:
        for bf in (0, 1):
            u.append((f"scan_{kind}_{'bf16' if bf else 'f32'}", "scan_inst.cu", [f"-DSCAN_KIND={kind}", f"-DSCAN_BF16={bf}"]))
    u.append(("scan_eval_f32", "scan_inst.cu", ["-DSCAN_KIND=K_EVAL", "-DSCAN_BF16=0", "-DSCAN_EVAL=1"]))
    return [x for x in u if os.path.exists(os.path.join(CSRC, x[1]))]


def source_hash():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def compile_one(unit, verbose):
    name, src, defs = unit
    obj = os.path.join(OBJ, name + ".o")
    cmd = [NVCC] + FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(OBJ, name + ".log"), "w") as f:
        f.write(log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    h = source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h:
        return LIB
    us = units()
    with cf.ThreadPoolExecutor(max_workers=min(len(us), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda u: compile_one(u, verbose), us))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(h)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
