"""ctypes binding of libb200ir.so (include/b200ir.h).  Fails loudly: no fallback of any kind."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200ir.so")

# metric ids / dtypes / flags: keep in sync with include/b200ir.h
L1, L2, LINF, COS_SIM, COS_DIST, ANGLE, MAG_DIFF, OPTIMIZED = range(8)
F32, BF16 = 0, 1
FLAG_RAW, FLAG_ABS_SCORE, FLAG_NO_TENSOR, FLAG_NO_RERANK, FLAG_HAVE_INDEX = 1, 2, 4, 8, 16
RGB, HSV = 0, 1
MAX_K = 256          # one result page
MAX_K_PAGED = 4096    # longest list ops.topk serves through b200ir_topk_paged + b200ir_sort_topk_rows
MAX_CANDIDATES = 1024 # longest candidate list of b200ir_rank_candidates

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p

_SIGNATURES = {
    "b200ir_version": (ctypes.c_int, []),
    "b200ir_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "b200ir_device_ok": (ctypes.c_int, []),
    "b200ir_row_sqnorms": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_vp, c_vp]),
    "b200ir_topk_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, c_i64, c_i64, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_int]),
    "b200ir_topk": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_vp, c_i64, ctypes.c_int, ctypes.c_int,
                                   c_i64, ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_vp, c_vp, c_vp,
                                   ctypes.c_size_t, c_vp]),
    "b200ir_topk_fallback_counter_offset": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, c_i64, c_i64, ctypes.c_int,
                                                              ctypes.c_int, ctypes.c_int]),
    "b200ir_index_bytes": (ctypes.c_size_t, [ctypes.c_int, c_i64, ctypes.c_int]),
    "b200ir_index_build": (ctypes.c_int, [ctypes.c_int, c_vp, c_i64, ctypes.c_int, c_vp, ctypes.c_size_t, c_vp]),
    "b200ir_topk_indexed": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_vp, c_i64, ctypes.c_int, ctypes.c_int,
                                           c_i64, ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_vp, c_vp, c_vp, ctypes.c_size_t,
                                           c_vp, ctypes.c_size_t, c_vp]),
    "b200ir_topk_paged": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_i64,
                                         ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "b200ir_sort_topk_rows": (ctypes.c_int, [ctypes.c_int, c_vp, c_vp, c_i64, ctypes.c_int, c_vp]),
    "b200ir_topk_multi_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, c_i64, c_i64,
                                                            ctypes.c_int, ctypes.c_int]),
    "b200ir_topk_multi": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_vp, c_i64, ctypes.c_int,
                                         ctypes.c_int, c_i64, ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_vp, c_vp, c_vp,
                                         ctypes.c_size_t, c_vp]),
    "b200ir_candidate_metrics": (ctypes.c_int, [ctypes.c_int, c_vp, c_i64, c_vp, c_i64, ctypes.c_int, c_vp, ctypes.c_int, c_vp, c_vp]),
    "b200ir_rank_candidates": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int, c_vp, c_vp,
                                              c_vp, c_vp, c_vp]),
    "b200ir_pairwise_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, c_i64, c_i64, ctypes.c_int]),
    "b200ir_pairwise": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_vp, c_i64, ctypes.c_int,
                                       ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "b200ir_topk_merge": (ctypes.c_int, [ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_vp, c_vp, c_vp]),
    "b200ir_topk_merge_strided": (ctypes.c_int, [ctypes.c_int, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, c_i64, ctypes.c_int, c_vp,
                                                   c_vp, c_vp]),
    "b200ir_allpairs_eval_workspace_bytes": (ctypes.c_size_t, [c_i64, ctypes.c_int, ctypes.c_int]),
    "b200ir_allpairs_eval": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                              ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double), ctypes.c_int,
                                              c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "b200ir_allpairs_eval_part": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                                   ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double), ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "b200ir_pair_metrics": (ctypes.c_int, [ctypes.c_int, c_vp, c_i64, c_vp, c_i64, ctypes.c_int, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "b200ir_threshold_dedupe": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, c_vp, c_i64, ctypes.c_double, ctypes.c_int,
                                               ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "b200ir_resize_crop_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 8),
    "b200ir_resize_crop": (ctypes.c_int, [c_vp, c_i64] + [ctypes.c_int] * 8 + [c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "b200ir_histogram": (ctypes.c_int, [ctypes.c_int, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "b200ir_counts_to_embedding": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "b200ir_launch_count": (ctypes.c_longlong, []),
    "b200ir_profile_enable": (None, [ctypes.c_int]),
    "b200ir_profile_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


class B200IRError(RuntimeError):
    pass


def load():
    """Load libb200ir.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200IRError(
            f"{LIB_PATH} not found: build it with `python image-retrieval-_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().b200ir_error_string(status)
        raise B200IRError(f"{what} failed: [{status}] {msg.decode() if msg else '?'}")
