"""The reference-side binding (bindings/pybind11/propagation.cpp): a pybind11 module named `propagation` over the C ABI,
i.e. what replaces /root/reference/precompute/propagation.cpp:8-12.  The CPU test builds and imports it; the GPU test runs
the reference's own call sequence (model.py:249-268) through `from precompute import propagation` against the golden."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import GOLDEN, check_topk_rows, load_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "bindings", "_build")


def _import_binding():
    from grandplus_b200 import build
    build.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "bindings", "pybind11")], stdout=subprocess.DEVNULL)
    if BUILD not in sys.path:
        sys.path.insert(0, BUILD)
    for name in [m for m in sys.modules if m == "precompute" or m.startswith("precompute.")]:
        del sys.modules[name]
    from precompute import propagation          # exactly model.py:9
    return propagation


def test_pybind11_binding_builds_and_imports():
    propagation = _import_binding()
    assert propagation.__file__.startswith(BUILD)
    assert hasattr(propagation, "Graph") and hasattr(propagation.Graph, "gfpush_omp")
    from grandplus_b200 import _lib
    if _lib.load().gp_device_count() == 0:      # no CPU fallback behind the binding either
        with pytest.raises((ValueError, RuntimeError)):
            propagation.Graph(np.array([0, 1, 2], np.int32), np.array([0, 1], np.int32), 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", [("cora", "ppr"), ("citeseer", "avg"), ("pubmed", "single")])
def test_pybind11_binding_runs_the_reference_call_sequence(name, mode):
    propagation = _import_binding()
    indptr, indices = load_graph(name)
    z = np.load(os.path.join(GOLDEN, f"gfpush_{name}_{mode}.npz"))
    K, rmax, coef = int(z["K"]), float(z["rmax"]), z["coef"]
    idx_train_unlabel = z["node_idx"].astype(np.int64)                     # model.py:247 passes int64
    indptr32 = np.array(indptr, dtype=np.int32)                            # model.py:249
    indices32 = np.array(indices, dtype=np.int32)                          # model.py:250
    graph = propagation.Graph(indptr32, indices32, 42)                     # model.py:251
    row_idx = np.zeros((idx_train_unlabel.shape[0] * K), dtype=np.int32)   # model.py:252
    col_idx = np.zeros((idx_train_unlabel.shape[0] * K), dtype=np.int32)   # model.py:253
    mat_value = np.zeros((idx_train_unlabel.shape[0] * K), dtype=np.float64)   # model.py:254
    graph.gfpush_omp(idx_train_unlabel, row_idx, col_idx, mat_value, coef, rmax, K)   # model.py:268
    worst = check_topk_rows(indptr, indices, z["node_idx"], coef, rmax, K, col_idx, mat_value, row=row_idx)
    assert worst < 1e-11
    import scipy.sparse as sp
    n = indptr.shape[0] - 1
    topk_adj = sp.coo_matrix((mat_value, (row_idx, col_idx)), (n, n)).tocsr()           # model.py:270-272
    ref_adj = sp.coo_matrix((z["value"], (z["row_idx"], z["col_idx"])), (n, n)).tocsr()
    assert abs(topk_adj.sum() - ref_adj.sum()) <= 1e-9 * ref_adj.sum()
    with pytest.raises(ValueError):                                        # wrong output dtype: refused, not silently lost
        graph.gfpush_omp(idx_train_unlabel, row_idx, col_idx, mat_value.astype(np.float32), coef, rmax, K)
