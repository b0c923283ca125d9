"""Drop-in for the reference's pybind11 module ``precompute.propagation``
(/root/reference/precompute/propagation.cpp:8-12; class ``Graph``,
/root/reference/precompute/graph.h:17-133).

Same class name, method name, positional order, dtypes and in-place output convention as the
reference, so ``model.py:249-268`` runs unchanged against it::

    graph = propagation.Graph(indptr, indices, seed)
    graph.gfpush_omp(idx_train_unlabel, row_idx, col_idx, mat_value, coef, rmax, top_k)

What differs, on purpose:
  * the work runs on the B200 through libgrandplus_b200.so (``gp_gfpush``); there is no CPU path;
  * the CSR is copied to the device at construction, so the borrowed-pointer lifetime hazard of
    graph.h:34-36 does not exist;
  * arguments are validated (dtype, contiguity, length, id range) and a ``ValueError`` /
    ``GPError`` is raised where the reference would silently overrun (graph.h:59-71);
  * the GIL is released for the duration of the call (ctypes does that; propagation.cpp holds it).
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib

__all__ = ["Graph"]


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


class Graph:
    """``Graph(indptr: int32[N+1], indices: int32[nnz], seed: int)`` -- graph.h:32-47."""

    def __init__(self, indptr, indices, seed=0, device=None):
        lib = _lib.load()
        _lib.require_cuda()
        # pybind's array_t<int> force-casts other integer dtypes (graph.h:32); do the same, explicitly
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        if indptr.ndim != 1 or indices.ndim != 1 or indptr.shape[0] < 2:
            raise ValueError("indptr must be 1-D with at least 2 entries and indices 1-D")
        if device is None:
            device = _current_device()
        self.num_nodes = int(indptr.shape[0] - 1)
        self.device = int(device)
        self.seed = int(seed)  # stored and never read, like graph.h:30,40
        handle = ctypes.c_void_p()
        _lib.check(lib.gp_graph_create(_ptr(indptr), self.num_nodes, _ptr(indices), int(indices.shape[0]),
                                       ctypes.c_int32(self.seed & 0x7FFFFFFF), self.device, ctypes.byref(handle)))
        self._h = handle
        self._lib = lib

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.gp_graph_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_device_csr(cls, indptr_t, indices_t):
        """Build from int32 CUDA tensors already on the device (kept alive by the object)."""
        import torch
        lib = _lib.load()
        _lib.require_cuda()
        if indptr_t.dtype != torch.int32 or indices_t.dtype != torch.int32 or not indptr_t.is_cuda:
            raise ValueError("from_device_csr needs int32 CUDA tensors")
        self = cls.__new__(cls)
        self._keep = (indptr_t.contiguous(), indices_t.contiguous())
        self.num_nodes = int(indptr_t.numel() - 1)
        self.device = indptr_t.device.index
        self.seed = 0
        handle = ctypes.c_void_p()
        _lib.check(lib.gp_graph_create_device(ctypes.c_void_p(self._keep[0].data_ptr()), self.num_nodes,
                                              ctypes.c_void_p(self._keep[1].data_ptr()), int(indices_t.numel()),
                                              self.device, ctypes.byref(handle)))
        self._h = handle
        self._lib = lib
        return self

    # -- configuration (no reference counterpart: the reference hard-codes 40 threads, graph.h:41) --
    def configure(self, scratch_mode=0, block_threads=0, ctas_per_sm=0, max_scratch_bytes=0):
        cfg = _lib.PushConfig(int(scratch_mode), int(block_threads), int(ctas_per_sm), int(max_scratch_bytes))
        _lib.check(self._lib.gp_graph_configure(self._h, ctypes.byref(cfg)))

    # -- the reference's one method ---------------------------------------------------------
    def gfpush_omp(self, node_idx, row_idx, col_idx, value, coef, rmax, K):
        """graph.h:53-131.  Fills ``row_idx``/``col_idx`` (int32) and ``value`` (float64), each
        ``[len(node_idx)*K]``, IN PLACE: source ``it`` owns slots ``it*K .. it*K+K-1``; unfilled
        slots are (0, 0, 0.0) as in the reference's caller-zeroed arrays (model.py:252-254)."""
        node_idx = np.ascontiguousarray(node_idx)
        if node_idx.dtype != np.int32:  # model.py:247,268 passes int64; pybind force-casts a copy
            if node_idx.size and (node_idx.min() < 0 or node_idx.max() >= self.num_nodes):
                raise ValueError("node_idx contains an id outside the graph")
            node_idx = node_idx.astype(np.int32)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        S, K = int(node_idx.shape[0]), int(K)
        for name, arr, dt in (("row_idx", row_idx, np.int32), ("col_idx", col_idx, np.int32),
                              ("value", value, np.float64)):
            # the reference writes through pybind temporaries when the dtype is wrong and the result
            # is silently lost; refuse instead
            if not isinstance(arr, np.ndarray) or arr.dtype != dt or not arr.flags.c_contiguous or not arr.flags.writeable:
                raise ValueError(f"{name} must be a writeable C-contiguous numpy array of dtype {np.dtype(dt).name}")
            if arr.size < S * K:
                raise ValueError(f"{name} has {arr.size} slots, needs len(node_idx)*K = {S * K}")
        _lib.check(self._lib.gp_gfpush(self._h, _ptr(node_idx), S, _ptr(coef), int(coef.shape[0]), float(rmax), K,
                                       _ptr(row_idx), _ptr(col_idx), _ptr(value)))

    # -- device-resident variant (Pi stays on the GPU) ------------------
    def gfpush_device(self, node_idx_t, coef, rmax, K, want_fp32=True, stream=None, check=False):
        """node_idx_t: int32 CUDA tensor [S].  Returns (row, col, val64, val32|None) CUDA tensors
        [S, K].  Asynchronous on the current torch stream.  Device-side refusals (a source id outside
        the graph, a list that outgrew its bound) are sticky on the handle and raise at the next
        ``check_errors()`` / ``last_stats()``; ``check=True`` waits for the push and raises at once."""
        import torch
        if node_idx_t.dtype != torch.int32 or not node_idx_t.is_cuda:
            raise ValueError("node_idx_t must be an int32 CUDA tensor")
        node_idx_t = node_idx_t.contiguous()
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        S, K = int(node_idx_t.numel()), int(K)
        dev = node_idx_t.device
        row = torch.empty((S, K), dtype=torch.int32, device=dev)
        col = torch.empty((S, K), dtype=torch.int32, device=dev)
        val = torch.empty((S, K), dtype=torch.float64, device=dev)
        val32 = torch.empty((S, K), dtype=torch.float32, device=dev) if want_fp32 else None
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        _lib.check(self._lib.gp_gfpush_device(
            self._h, ctypes.c_void_p(node_idx_t.data_ptr()), S, _ptr(coef), int(coef.shape[0]), float(rmax), K,
            ctypes.c_void_p(row.data_ptr()), ctypes.c_void_p(col.data_ptr()), ctypes.c_void_p(val.data_ptr()),
            ctypes.c_void_p(val32.data_ptr() if val32 is not None else 0), ctypes.c_void_p(st)))
        if check:
            self.check_errors()
        return row, col, val, val32

    def check_errors(self) -> None:
        """Waits for the handle's last push and raises GPError if the device refused anything since
        the last check (gp_gfpush_device itself returns before the kernels run)."""
        self.last_stats()

    def last_stats(self) -> dict:
        st = _lib.PushStats()
        _lib.check(self._lib.gp_gfpush_last_stats(self._h, ctypes.byref(st)))
        return st.as_dict()


    def phase_cycles(self, reset=False) -> dict:
        """SM cycles per phase, summed over the persistent CTAs (profiling hook).  The cluster kernel reports
        fetch / expand / settle / exchange / topk / resident, the per-CTA kernels fetch / expand / settle / reserve merge /
        topk / the widest level's expand and settle / resident."""
        out = (ctypes.c_uint64 * 8)()
        _lib.check(self._lib.gp_gfpush_phase_cycles(self._h, ctypes.byref(out), int(bool(reset))))
        names = ("fetch", "expand", "settle", "merge_or_exchange", "topk", "wide_expand", "wide_settle", "resident")
        return dict(zip(names, [int(x) for x in out]))

    def cumulative_stats(self, reset=False) -> dict:
        st = _lib.PushStats()
        _lib.check(self._lib.gp_gfpush_cumulative_stats(self._h, ctypes.byref(st), int(bool(reset))))
        return st.as_dict()


def _current_device() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0
