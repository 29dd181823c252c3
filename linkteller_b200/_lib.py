"""ctypes binding of liblinkteller_b200.so (include/linkteller_b200.h).

This is the stub a maintainer of the reference would add.  There is NO fallback:
if the library is missing or a call fails, a RuntimeError is raised — the product path never
routes through the oracle or through torch ops.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblinkteller_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "linkteller_b200.h")

_lib = None

P = C.c_void_p
I32 = C.c_int32
I64 = C.c_int64
INT = C.c_int
F32 = C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPE)
_SIGS = {
    "lt_last_error": [],
    "lt_version": [],
    "lt_device_info": [P, P, P, P],
    "lt_csr_from_edges": [P, I64, I32, INT, P, P, P, I64, P, P],
    "lt_coo_to_csr": [P, P, I64, I32, P, P, P],
    "lt_csr_check": [P, P, I32, I64, P, P],
    "lt_csr_row_sums": [P, P, I32, INT, P, P],
    "lt_adj_normalize": [P, P, P, I32, INT, INT, P, P, P, P, I64, P, P],
    "lt_gemm_f32": [P, I64, P, P, I64, INT, P, INT, P, I64, I32, I32, I32, INT, P],
    "lt_spmm_csr_f32": [P, P, P, I32, P, I32, P, I64, I32, P, INT, P, I64, P],
    "lt_gcn_forward": [P, P, P, I32, P, I64, I32, P, P, P, P, I32, I32, P, I64, P],
    "lt_plan_create": [P, P, P, I32, P, I32, I32, I32, P, P],
    "lt_plan_destroy": [P, P],
    "lt_plan_info": [P, I32, I32, P, P],
    "lt_plan_rows": [P],
    "lt_plan_copy_rows": [P, P, P],
    "lt_influence_rows": [P, P, I64, P, I64, P, P, I32, I32, F32, P, I64, P],
    "lt_attack_efficient": [P, P, P, I32, P, I64, I32, P, P, P, I32, I32, P, I32, I32, I32, F32, P, I64, P, P],
    "lt_gcn3_forward": [P, P, P, I32, P, I64, I32, P, P, P, P, P, P, I32, I32, I32, P, I64, P],
    "lt_attack_efficient3": [P, P, P, I32, P, I64, I32, P, P, P, P, P, I32, I32, I32, P, I32, I32, I32, F32, P, I64, I64,
                             P, P],
    "lt_pair_labels": [P, P, I32, P, I32, P, P, P],
    "lt_pair_scores": [P, I64, I32, P, P],
    "lt_auc_ap": [P, P, I64, P, P, P],
    "lt_topk_tp": [P, P, I64, P, I32, P, P, P],
    "lt_rank_curve": [P, P, I64, P, P, P, I64, P, P],
    "lt_lsa2_corr": [P, I64, I32, P, I32, INT, P, I64, P],
    "lt_lsa2_pairs": [P, I64, I32, I32, INT, P, P, I64, P, P],
    "lt_pair_gather": [P, I64, I32, I32, P, P, I64, P, P],
    "lt_dp_edgerand_apply": [P, P, P, I32, P, P, I64, P, P, P, I64, P, P],
    "lt_dp_lapgraph_select": [P, P, P, I32, P, I64, P, P, P, I64, P, P],
    "lt_pair_influence": [P, P, P, I32, P, I64, I32, P, P, P, I32, I32, P, I32, P, P, I64, F32, P, P],
    "lt_comm_create": [INT, INT, I64, I64, P],
    "lt_comm_get_handle": [P, P],
    "lt_comm_open": [P, P],
    "lt_comm_buffer": [P],
    "lt_comm_timed_out": [P, P],
    "lt_comm_destroy": [P],
    "lt_attack_efficient_comm": [P, P, P, I32, P, I64, I32, P, P, P, I32, I32, P, I32, I32, I32, F32, P, P, P],
    "lt_set_option": [C.c_char_p, INT],
    "lt_get_option": [C.c_char_p, P],
    "lt_launch_count": [],
    "lt_prof_enable": [INT],
    "lt_prof_reset": [],
    "lt_prof_report": [C.c_char_p, I64],
    "lt_exclusive_scan_i32": [P, P, I64, P],
    "lt_radix_sort_pairs_u64": [P, P, I64, INT, INT, P],
}
_RESTYPE = {"lt_last_error": C.c_char_p, "lt_plan_rows": P, "lt_launch_count": I64, "lt_comm_buffer": P}


class LibraryMissing(RuntimeError):
    pass


def header_symbols():
    """Every function name include/linkteller_b200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lt_[a-z0-9_]+)\s*\(", txt)))


def load():
    """Load the shared library (once).  Raises LibraryMissing when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found — build it with `python -m linkteller_b200.build` "
            "(there is no CPU or torch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().lt_last_error().decode(errors="replace")
        raise RuntimeError(f"linkteller_b200 {what} failed (status {status}): {msg}")
