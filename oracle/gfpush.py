"""oracle.gfpush -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of ``oracle/gfpush_oracle.c`` (our C restatement of
/root/reference/precompute/graph.h:53-131) and a loader for the reference's own
``precompute.propagation`` pybind11 module compiled unmodified into ``oracle/_ref``
(``make -C oracle ref``; only possible where /root/reference exists, the binary then
travels with the repo snapshot).

Parity status: PINNED against the reference module (tests/test_oracle.py, tests/golden/).
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sys
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgp_oracle.so")
_REF_DIR = os.path.join(_HERE, "_ref")


class _Stats(ctypes.Structure):
    _fields_ = [
        ("edges_pushed", ctypes.c_int64),
        ("frontier_total", ctypes.c_int64),
        ("support_total", ctypes.c_int64),
        ("max_frontier", ctypes.c_int64),
        ("max_support", ctypes.c_int64),
    ]


@dataclass
class PushStats:
    edges_pushed: int
    frontier_total: int
    support_total: int
    max_frontier: int
    max_support: int


def build(force: bool = False) -> str:
    """Compile gfpush_oracle.c (gcc) if the library is missing or stale."""
    src = os.path.join(_HERE, "gfpush_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f64p = ctypes.POINTER(ctypes.c_double)
        lib.gp_oracle_gfpush.restype = ctypes.c_int
        lib.gp_oracle_gfpush.argtypes = [i32p, i32p, ctypes.c_int32, i32p, ctypes.c_int64, f64p,
                                         ctypes.c_int32, ctypes.c_double, ctypes.c_int32, i32p, i32p,
                                         f64p, ctypes.POINTER(_Stats), ctypes.c_int32]
        lib.gp_oracle_reserve_row.restype = ctypes.c_int
        lib.gp_oracle_reserve_row.argtypes = [i32p, i32p, ctypes.c_int32, ctypes.c_int32, f64p,
                                              ctypes.c_int32, ctypes.c_double, f64p,
                                              ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(_Stats)]
        _lib = lib
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _csr(indptr, indices):
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    return indptr, indices


def coef_for(prop_mode: str, order: int, alpha: float = 0.2) -> np.ndarray:
    """The caller's weight vector, /root/reference/model.py:255-267 (normalised to sum 1)."""
    if prop_mode == "avg":
        coef = list(np.ones(order + 1, dtype=np.float64))
    elif prop_mode == "ppr":
        coef = [alpha]
        for _ in range(order):
            coef.append(coef[-1] * (1 - alpha))
    elif prop_mode == "single":
        coef = list(np.zeros(order + 1, dtype=np.float64))
        coef[-1] = 1.0
    else:
        raise ValueError(f"Unknown propagation mode: {prop_mode}")
    return np.asarray(coef) / np.sum(coef)


def gfpush(indptr, indices, node_idx, coef, rmax, K, nthreads: int = 0):
    """Oracle GFPush + top-k.  Returns (row_idx, col_idx, value, PushStats), the first
    three laid out exactly like the reference's outputs (graph.h:117-126): int32/int32/
    float64 [S*K], zero where a slot is unfilled; within a row sorted by descending value."""
    lib = _load()
    indptr, indices = _csr(indptr, indices)
    node_idx = np.ascontiguousarray(node_idx, dtype=np.int32)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    S = node_idx.shape[0]
    row = np.zeros(S * K, dtype=np.int32)
    col = np.zeros(S * K, dtype=np.int32)
    val = np.zeros(S * K, dtype=np.float64)
    st = _Stats()
    rc = lib.gp_oracle_gfpush(_p(indptr, ctypes.c_int32), _p(indices, ctypes.c_int32),
                              indptr.shape[0] - 1, _p(node_idx, ctypes.c_int32), S,
                              _p(coef, ctypes.c_double), coef.shape[0], float(rmax), int(K),
                              _p(row, ctypes.c_int32), _p(col, ctypes.c_int32), _p(val, ctypes.c_double),
                              ctypes.byref(st), int(nthreads))
    if rc != 0:
        raise MemoryError("gp_oracle_gfpush failed")
    return row, col, val, PushStats(st.edges_pushed, st.frontier_total, st.support_total,
                                    st.max_frontier, st.max_support)


def reserve_row(indptr, indices, src, coef, rmax):
    """Dense un-truncated reserve vector of one source: (values float64[N], seen uint8[N], PushStats)."""
    lib = _load()
    indptr, indices = _csr(indptr, indices)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    n = indptr.shape[0] - 1
    out = np.zeros(n, dtype=np.float64)
    seen = np.zeros(n, dtype=np.uint8)
    st = _Stats()
    rc = lib.gp_oracle_reserve_row(_p(indptr, ctypes.c_int32), _p(indices, ctypes.c_int32), n, int(src),
                                   _p(coef, ctypes.c_double), coef.shape[0], float(rmax),
                                   _p(out, ctypes.c_double), _p(seen, ctypes.c_uint8), ctypes.byref(st))
    if rc != 0:
        raise MemoryError("gp_oracle_reserve_row failed")
    return out, seen, PushStats(st.edges_pushed, st.frontier_total, st.support_total,
                                st.max_frontier, st.max_support)


# ---------------------------------------------------------------------------------------------
# The reference's own module (oracle/_ref), when it has been built.
# ---------------------------------------------------------------------------------------------
def reference_available() -> bool:
    d = os.path.join(_REF_DIR, "precompute")
    return os.path.isdir(d) and any(f.startswith("propagation") and f.endswith(".so") for f in os.listdir(d))


def load_reference():
    """Import the reference's ``precompute.propagation`` (unmodified propagation.cpp,
    /root/reference/precompute/propagation.cpp:8-12) from oracle/_ref.  Raises if absent."""
    d = os.path.join(_REF_DIR, "precompute")
    if not reference_available():
        raise ImportError("oracle/_ref is not built (make -C oracle ref needs /root/reference)")
    name = "_gp_reference_propagation"
    if name in sys.modules:
        return sys.modules[name]
    so = [f for f in os.listdir(d) if f.startswith("propagation") and f.endswith(".so")][0]
    # The module's init symbol is PyInit_propagation, so it must be loaded under that name.
    spec = importlib.util.spec_from_file_location("propagation", os.path.join(d, so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def reference_gfpush(indptr, indices, node_idx, coef, rmax, K, nthreads=None):
    """Run the reference exactly as /root/reference/model.py:249-268 does.
    Returns (row_idx, col_idx, value) as the reference filled them.
    ``nthreads``: the reference's constructor calls omp_set_num_threads(40) (graph.h:41,46); when given, the same
    call is made again with this value before gfpush_omp -- the unmodified binary with NUMTHREAD = nthreads
    (BASELINE.md 3 asks for the cpu_count setting beside the literal 40)."""
    prop = load_reference()
    indptr = np.array(indptr, dtype=np.int32)
    indices = np.array(indices, dtype=np.int32)
    graph = prop.Graph(indptr, indices, 0)
    if nthreads:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(nthreads))
    node_idx = np.ascontiguousarray(node_idx, dtype=np.int32)
    S = node_idx.shape[0]
    row = np.zeros(S * K, dtype=np.int32)
    col = np.zeros(S * K, dtype=np.int32)
    val = np.zeros(S * K, dtype=np.float64)
    graph.gfpush_omp(node_idx, row, col, val, np.ascontiguousarray(coef, dtype=np.float64), float(rmax), int(K))
    # keep the borrowed CSR buffers alive until the call has returned (graph.h:34-36)
    del graph
    return row, col, val


def rows_as_sets(col, val, K):
    """[S*K] slot arrays -> list of (cols sorted, vals in that order) keeping only v > 0."""
    col = np.asarray(col).reshape(-1, K)
    val = np.asarray(val).reshape(-1, K)
    out = []
    for c, v in zip(col, val):
        keep = v > 0
        c, v = c[keep], v[keep]
        o = np.argsort(c, kind="stable")
        out.append((c[o], v[o]))
    return out
