"""ctypes binding of libgrandplus_b200.so (include/grandplus_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing it is built with nvcc
(``build.build()``); if that is impossible the import raises.  Nothing here imports ``oracle``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libgrandplus_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_vp = ctypes.c_void_p

GP_OK = 0
GP_SCRATCH_AUTO, GP_SCRATCH_SMEM, GP_SCRATCH_HBM = 0, 1, 2


class GPError(RuntimeError):
    """A C-ABI call failed; carries the library's status code and message."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libgrandplus_b200 error {status}: {message}")
        self.status = status


class PushConfig(ctypes.Structure):
    _fields_ = [("scratch_mode", ctypes.c_int32), ("block_threads", ctypes.c_int32),
                ("ctas_per_sm", ctypes.c_int32), ("max_scratch_bytes", ctypes.c_int64)]


class PushStats(ctypes.Structure):
    _fields_ = [("edges_pushed", ctypes.c_int64), ("frontier_total", ctypes.c_int64),
                ("support_total", ctypes.c_int64), ("sources", ctypes.c_int64), ("ctas", ctypes.c_int64),
                ("scratch_bytes", ctypes.c_int64), ("scratch_mode", ctypes.c_int32),
                ("kernel_launches", ctypes.c_int32), ("cluster_sources", ctypes.c_int64),
                ("redo_sources", ctypes.c_int64), ("cluster_size", ctypes.c_int32), ("table_slots", ctypes.c_int32),
                ("bucket_count", ctypes.c_int32), ("reserved", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AggregateArgs(ctypes.Structure):
    _fields_ = [("table", c_vp), ("n_table_rows", ctypes.c_int64), ("F", ctypes.c_int32),
                ("ld_table", ctypes.c_int64), ("row_ptr", c_vp), ("slot_rows", c_vp), ("slot_K", ctypes.c_int32),
                ("nbr", c_vp), ("score", c_vp), ("B", ctypes.c_int64), ("n_entries", ctypes.c_int64),
                ("p", ctypes.c_double), ("training", ctypes.c_int32), ("n_aug", ctypes.c_int32),
                ("seed", ctypes.c_uint64), ("offset", ctypes.c_uint64), ("mask_in", c_vp), ("mask_out", c_vp),
                ("eps", ctypes.c_float), ("out", c_vp), ("ld_out", ctypes.c_int64), ("denom_out", c_vp)]


class AggregateBwdArgs(ctypes.Structure):
    _fields_ = [("grad_out", c_vp), ("ld_grad_out", ctypes.c_int64), ("denom", c_vp), ("row_ptr", c_vp),
                ("nbr", c_vp), ("score", c_vp), ("B", ctypes.c_int64), ("n_entries", ctypes.c_int64),
                ("F", ctypes.c_int32), ("p", ctypes.c_double), ("training", ctypes.c_int32),
                ("n_aug", ctypes.c_int32), ("mask_in", c_vp), ("grad_table", c_vp),
                ("ld_grad_table", ctypes.c_int64), ("n_table_rows", ctypes.c_int64)]


# name -> (restype, argtypes); every symbol include/grandplus_b200.h declares
SIGNATURES = {
    "gp_last_error": (ctypes.c_char_p, []),
    "gp_abi_version": (ctypes.c_int, []),
    "gp_device_count": (ctypes.c_int, []),
    "gp_graph_create": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int,
                                       ctypes.POINTER(c_vp)]),
    "gp_graph_create_device": (ctypes.c_int, [c_vp, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_int,
                                              ctypes.POINTER(c_vp)]),
    "gp_graph_destroy": (None, [c_vp]),
    "gp_graph_num_nodes": (ctypes.c_int64, [c_vp]),
    "gp_graph_num_edges": (ctypes.c_int64, [c_vp]),
    "gp_graph_configure": (ctypes.c_int, [c_vp, ctypes.POINTER(PushConfig)]),
    "gp_gfpush": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int32, ctypes.c_double, ctypes.c_int32,
                                 c_vp, c_vp, c_vp]),
    "gp_gfpush_device": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_vp, ctypes.c_int32, ctypes.c_double,
                                        ctypes.c_int32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gp_gfpush_phase_cycles": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_uint64 * 8), ctypes.c_int]),
    "gp_gfpush_last_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(PushStats)]),
    "gp_gfpush_cumulative_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(PushStats), ctypes.c_int]),
    "gp_aggregate_fwd": (ctypes.c_int, [ctypes.POINTER(AggregateArgs), c_vp]),
    "gp_aggregate_bwd": (ctypes.c_int, [ctypes.POINTER(AggregateBwdArgs), c_vp]),
    "gp_segments_from_sorted_index": (ctypes.c_int, [c_vp, ctypes.c_int64, ctypes.c_int64, c_vp, c_vp, c_vp]),
    "gp_narrow_index": (ctypes.c_int, [c_vp, ctypes.c_int64, ctypes.c_int64, c_vp, c_vp, c_vp]),
    "gp_set_tuning": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64]),
    "gp_emb_dropout_fwd": (ctypes.c_int, [c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, c_vp, c_vp, c_vp, ctypes.c_int64,
                                          ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_float, c_vp, ctypes.c_int64,
                                          c_vp, c_vp]),
    "gp_emb_dropout_bwd": (ctypes.c_int, [c_vp, ctypes.c_int64, ctypes.c_int32, c_vp, c_vp, c_vp, c_vp, ctypes.c_int64,
                                          ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64, c_vp, ctypes.c_int64, c_vp]),
    "gp_emb_dropout_mask": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                           c_vp, c_vp]),
    "gp_lazy_adam_rows": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int32, c_vp, ctypes.c_int64, c_vp,
                                         ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_float, c_vp]),
    "gp_dropnode_mask": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_uint64,
                                        ctypes.c_uint64, c_vp, c_vp]),
}

_lib = None


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """dlopen the library (building it first if absent).  Raises -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        # content-stamped: a no-op when the objects match the sources, a rebuild when they are stale
        from . import build as _build
        try:
            _build.build()
        except Exception:
            if not os.path.exists(LIB_PATH):
                raise
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing; run `python __graft_entry__.py build`")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gp_abi_version() != 1:
        raise ImportError(f"{LIB_NAME} has ABI version {lib.gp_abi_version()}, this binding expects 1")
    _lib = lib
    # GP_TUNING="key=value,key=value": performance knobs for sweeps (gp_set_tuning; results never depend on them)
    for item in filter(None, os.environ.get("GP_TUNING", "").split(",")):
        k, v = item.split("=")
        if lib.gp_set_tuning(k.strip().encode(), int(v)) != GP_OK:
            raise GPError(-1, lib.gp_last_error().decode("utf-8", "replace"))
    return lib


def check(status: int) -> None:
    if status != GP_OK:
        msg = load().gp_last_error()
        raise GPError(status, msg.decode("utf-8", "replace") if msg else "unknown error")


def set_tuning(key: str, value: int) -> None:
    check(load().gp_set_tuning(key.encode(), int(value)))


def require_cuda() -> None:
    """Fail loudly when there is no device: there is no CPU path behind this package."""
    if load().gp_device_count() <= 0:
        raise GPError(-2, "no CUDA device visible; grandplus_b200 has no CPU fallback")
