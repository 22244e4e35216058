"""ctypes binding of libppr_b200.so (the C-ABI declared in include/pprb200.h).

There is no fallback: if the shared library has not been built (``python -c 'import __graft_entry__ as g;
g.build()'``) loading raises, and compute entry points fail with PPRB200_ERR_CUDA when no sm_100 GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libppr_b200.so"

_lib = None

c_void_p, c_int32, c_uint32, c_uint64, c_int64, c_double, c_int = (
    C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_int64, C.c_double, C.c_int)


class Stats(C.Structure):
    """pprb200_stats (include/pprb200.h)."""
    _fields_ = [
        ("iterations_run", C.c_uint32),
        ("n_gpus", C.c_uint32),
        ("node_iterations", C.c_uint64),
        ("nonsink_node_iterations", C.c_uint64),
        ("edge_reads", C.c_uint64),
        ("merged_entries", C.c_uint64),
        ("candidates", C.c_uint64),
        ("truncations", C.c_uint64),
        ("boundary_ties", C.c_uint64),
        ("algorithmic_bytes", C.c_uint64),
        ("walk_steps", C.c_uint64),
        ("walks", C.c_uint64),
        ("overflow_requeues", C.c_uint64),
        ("walk_algorithmic_bytes", C.c_uint64),
        ("max_diff", C.c_double * 2),
        ("prep_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("kernel_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("total_ms", C.c_double),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == "max_diff" else v
        return d


class PprB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"pprb200 error {code}: {message}")
        self.code = code
        self.message = message


def ptr(a):
    """numpy array (or None) -> void*"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("PPRB200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(path))
    lib.pprb200_version.restype = C.c_char_p
    lib.pprb200_last_error.restype = C.c_char_p
    lib.pprb200_device_count.restype = c_int
    lib.pprb200_find_partitions.argtypes = [c_void_p, c_void_p, c_int32, c_void_p]
    lib.pprb200_find_partitions_device.argtypes = [c_void_p, c_void_p, c_int32, c_void_p]
    lib.pprb200_grank.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_uint32, c_uint32, c_uint32, c_double, c_double,
                                  c_uint32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pprb200_mccompletepathv2.argtypes = [c_void_p, c_void_p, c_int32, c_uint32, c_uint32, c_uint32, c_double, c_uint64,
                                             c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pprb200_session_create.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_uint32, c_uint32, c_int32, c_int32,
                                           c_void_p, C.POINTER(c_void_p)]
    lib.pprb200_session_destroy.argtypes = [c_void_p]
    lib.pprb200_session_destroy.restype = None
    lib.pprb200_session_grank.argtypes = [c_void_p, c_uint32, c_uint32, c_uint32, c_double, c_double]
    lib.pprb200_session_mc.argtypes = [c_void_p, c_uint32, c_uint32, c_uint32, c_double, c_uint64, c_uint32]
    lib.pprb200_session_fetch.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pprb200_session_stats.argtypes = [c_void_p, c_void_p]
    lib.pprb200_session_kernel_time.argtypes = [c_void_p, c_int, C.POINTER(c_uint32), C.POINTER(c_double)]
    lib.pprb200_session_launches.argtypes = [c_void_p, C.POINTER(c_uint64)]
    lib.pprb200_session_launches.restype = c_int
    lib.pprb200_session_ipc_export.argtypes = [c_void_p, c_void_p]
    lib.pprb200_session_ipc_attach.argtypes = [c_void_p, c_void_p]
    lib.pprb200_ppr_exact.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_uint32, c_uint32, c_double, c_double, c_void_p,
                                      c_void_p, C.POINTER(c_double)]
    lib.pprb200_debug_host_plan.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_uint32, c_int32, c_int32, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.pprb200_shard_owner.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_uint32, c_int32, c_void_p]
    lib.pprb200_debug_need_mask.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_uint32, c_int32, c_void_p, c_void_p]
    lib.pprb200_debug_need_mask.restype = c_int
    lib.pprb200_gen_rmat.argtypes = [c_uint32, c_uint32, c_uint64, c_double, c_double, c_double, c_void_p, c_void_p]
    lib.pprb200_gen_ba.argtypes = [c_int32, c_uint32, c_uint64, c_void_p, c_void_p, C.POINTER(c_int64)]
    for name in ("pprb200_find_partitions", "pprb200_find_partitions_device", "pprb200_grank", "pprb200_mccompletepathv2", "pprb200_session_create",
                 "pprb200_session_grank", "pprb200_session_mc", "pprb200_session_fetch", "pprb200_session_stats",
                 "pprb200_session_kernel_time", "pprb200_gen_rmat", "pprb200_gen_ba", "pprb200_session_ipc_export",
                 "pprb200_session_ipc_attach", "pprb200_shard_owner"):
        getattr(lib, name).restype = c_int
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise PprB200Error(rc, load().pprb200_last_error().decode())


def exported_symbols():
    """Names declared in include/pprb200.h (parsed), for the `library exports every symbol` test."""
    import re
    text = (_HERE.parent / "include" / "pprb200.h").read_text()
    return sorted(set(re.findall(r"\b(pprb200_[a-z0-9_]+)\s*\(", text)))
