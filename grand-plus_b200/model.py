"""Host-side mirror of the reference's aggregation interface, backed by the fused CUDA kernel.

Mirrors (same names, argument order and meaning):
  * ``Grand_Plus.random_prop(feats, mat_scores, mat_idx, dropnode_rate)``
    -- /root/reference/model.py:80-87 and /root/reference/model_mag.py:80-86
  * ``MLP.emb(attr_idx, node_idx, attr_data)`` / ``Grand_Plus.emb`` -- /root/reference/model_mag.py:48-55,88-90

and adds the fused entry points the edited call sites use:
  * :func:`random_prop_fused` -- replaces the host gather + H2D copy (model.py:314) AND random_prop
    (model.py:322) with one launch over device-resident features;
  * :class:`PiMatrix` -- device-resident top-k propagation matrix built directly from the GFPush
    output slots, replacing ``coo_matrix(...).tocsr()`` and the per-batch scipy row slice
    (model.py:270-272, 310-313).

torch is plumbing here (device memory, streams, autograd glue); every reduction runs in
libgrandplus_b200.so.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

EPS_RANDOM_PROP = 1e-12  # model.py:87
EPS_EMB = 1e-10          # model_mag.py:54


def _vp(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("grandplus_b200 has no CPU path: all tensors must be CUDA tensors")


class DeviceFeatures:
    """Feature matrix X resident in HBM, rows padded to a 16-byte multiple so every row start is
    128-bit aligned.  ``data`` is [N, ld] fp32 with ld = ceil(F/4)*4; columns >= F are 0."""

    def __init__(self, features, device=None):
        x = torch.as_tensor(features)
        if x.dim() != 2:
            raise ValueError("features must be [N, F]")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.F = int(x.shape[1])
        ld = (self.F + 3) // 4 * 4
        if x.is_cuda and x.dtype == torch.float32 and ld == self.F and x.is_contiguous():
            self.data = x
        else:
            self.data = torch.zeros((x.shape[0], ld), dtype=torch.float32, device=device)
            self.data[:, : self.F] = x.to(device=device, dtype=torch.float32, non_blocking=True)
        self.N = int(x.shape[0])
        self.ld = ld

    @property
    def device(self):
        return self.data.device


def segments_from_sorted_index(mat_idx: torch.Tensor, check: bool = True):
    """``mat_idx`` (int64, ascending) -> (row_ptr int32 [B+1], B) with B = mat_idx[-1]+1 -- the
    ``dim_size`` of model.py:84.  One host sync, as in the reference (torch_scatter reads it too)."""
    _need_cuda(mat_idx)
    lib = _lib.load()
    if mat_idx.dtype != torch.int64:
        mat_idx = mat_idx.to(torch.int64)
    mat_idx = mat_idx.contiguous()
    n = int(mat_idx.numel())
    if n == 0:
        raise ValueError("empty batch: mat_idx[-1] is undefined (the reference fails here too)")
    B = int(mat_idx[-1].item()) + 1
    row_ptr = torch.empty(B + 1, dtype=torch.int32, device=mat_idx.device)
    flags = torch.zeros(2, dtype=torch.int32, device=mat_idx.device)
    _lib.check(lib.gp_segments_from_sorted_index(_vp(mat_idx), n, B, _vp(row_ptr), _vp(flags), _stream(mat_idx.device)))
    if check and int(flags[0].item()) != 0:
        raise ValueError("mat_idx must be ascending and non-negative (model.py:312 produces it that way)")
    return row_ptr, B


def narrow_index(idx: torch.Tensor, n_rows: int) -> torch.Tensor:
    """int64 -> int32 row ids with a range check against n_rows."""
    _need_cuda(idx)
    if idx.dtype == torch.int32:   # already narrow: still range-checked (the kernels index the table with it)
        idx = idx.contiguous()
        if idx.numel() and (int(idx.min().item()) < 0 or int(idx.max().item()) >= int(n_rows)):
            raise IndexError(f"index out of range for a table of {n_rows} rows")
        return idx
    lib = _lib.load()
    idx = idx.to(torch.int64).contiguous()
    out = torch.empty(idx.numel(), dtype=torch.int32, device=idx.device)
    flags = torch.zeros(2, dtype=torch.int32, device=idx.device)
    _lib.check(lib.gp_narrow_index(_vp(idx), int(idx.numel()), int(n_rows), _vp(out), _vp(flags), _stream(idx.device)))
    if int(flags[0].item()) != 0:
        raise IndexError(f"index out of range for a table of {n_rows} rows")
    return out


MAX_AUG_PER_LAUNCH = 4   # one Philox block (four 32-bit words) per entry decides four augmentations


def _chunk_offset(offset: int, c: int) -> int:
    """Philox offset of the c-th group of four augmentations (c = 0: the caller's offset itself)."""
    return (int(offset) + c * 0x9E3779B97F4A7C15) & (2**64 - 1)


def _per_aug_group(A: int, mask, launch):
    """The reference's --sample is unbounded (model.py:321); a launch handles four augmentations (one Philox block per
    entry), so more are produced in groups of four with independent Philox offsets and concatenated.
    ``launch(a, c, mask_group)`` -> (out [a,B,F], mask [a,nz] or None)."""
    if A <= MAX_AUG_PER_LAUNCH:
        return launch(A, 0, mask)
    outs, masks = [], []
    for c, a0 in enumerate(range(0, A, MAX_AUG_PER_LAUNCH)):
        a = min(MAX_AUG_PER_LAUNCH, A - a0)
        o, m = launch(a, c, None if mask is None else mask[a0:a0 + a].contiguous())
        outs.append(o); masks.append(m)
    return torch.cat(outs, 0), (torch.cat(masks, 0) if masks[0] is not None else None)



def dropnode_mask(n_entries: int, n_aug: int, p: float, seed: int, offset: int, device) -> torch.Tensor:
    """The keep-mask gp_aggregate_fwd draws for (seed, offset): uint8 [n_aug, n_entries]."""
    lib = _lib.load()
    mask = torch.empty((n_aug, n_entries), dtype=torch.uint8, device=device)
    for c, a0 in enumerate(range(0, n_aug, MAX_AUG_PER_LAUNCH)):
        a = min(MAX_AUG_PER_LAUNCH, n_aug - a0)
        _lib.check(lib.gp_dropnode_mask(int(n_entries), int(a), float(p), int(seed) & (2**64 - 1),
                                        _chunk_offset(offset, c), _vp(mask[a0:a0 + a]), _stream(mask.device)))
    return mask


def _launch_fwd(table, F_cols, ld_table, row_ptr, slot_rows, slot_K, nbr, score, B, n_entries, p, training, n_aug,
                seed, offset, mask_in, want_mask, eps, want_denom, out=None):
    lib = _lib.load()
    dev = table.device
    # rows padded to a 16-byte multiple so the kernel can use 128-bit stores; callers get the [.., :F] view
    ld_out = (int(F_cols) + 3) // 4 * 4
    if out is None:
        out = torch.empty((n_aug, B, ld_out), dtype=torch.float32, device=dev)
    else:   # caller-provided [n_aug, B, ld] fp32 buffer (may be a row range of a larger table)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape[-1] % 4 == 0
        assert out.numel() == n_aug * B * out.shape[-1] and out.shape[-1] >= ld_out
        ld_out = int(out.shape[-1])
        out = out.view(n_aug, B, ld_out)
    mask_out = torch.empty((n_aug, n_entries), dtype=torch.uint8, device=dev) if want_mask else None
    denom = torch.empty((n_aug, B), dtype=torch.float32, device=dev) if want_denom else None
    a = _lib.AggregateArgs()
    a.table = table.data_ptr(); a.n_table_rows = int(table.shape[0]); a.F = int(F_cols); a.ld_table = int(ld_table)
    a.row_ptr = 0 if row_ptr is None else row_ptr.data_ptr()
    a.slot_rows = 0 if slot_rows is None else slot_rows.data_ptr()
    a.slot_K = int(slot_K)
    a.nbr = 0 if nbr is None else nbr.data_ptr()
    a.score = score.data_ptr(); a.B = int(B); a.n_entries = int(n_entries)
    a.p = float(p); a.training = int(bool(training)); a.n_aug = int(n_aug)
    a.seed = int(seed) & (2**64 - 1); a.offset = int(offset) & (2**64 - 1)
    a.mask_in = 0 if mask_in is None else mask_in.data_ptr()
    a.mask_out = 0 if mask_out is None else mask_out.data_ptr()
    a.eps = float(eps); a.out = out.data_ptr(); a.ld_out = ld_out
    a.denom_out = 0 if denom is None else denom.data_ptr()
    _lib.check(lib.gp_aggregate_fwd(ctypes.byref(a), _stream(dev)))
    return (out if ld_out == F_cols else out[:, :, :F_cols]), mask_out, denom


class _AggregateFn(torch.autograd.Function):
    """out[a,b,:] = sum_j m_aj table[nbr_j,:] / (sum_j m_aj + eps); gradient w.r.t. ``table`` only
    (no gradient reaches scores or indices in the reference either)."""

    @staticmethod
    def forward(ctx, table, F_cols, row_ptr, nbr, score, B, p, training, n_aug, seed, offset, mask_in, eps,
                return_mask):
        needs_grad = table.requires_grad
        use_mask = bool(training) and p > 0.0
        want_mask = (use_mask and needs_grad and mask_in is None) or return_mask
        n_entries = int(score.numel())
        out, mask_out, denom = _launch_fwd(table, F_cols, table.stride(0), row_ptr, None, 0, nbr, score, B, n_entries,
                                           p, training, n_aug, seed, offset, mask_in, want_mask, eps, needs_grad)
        if needs_grad:
            ctx.save_for_backward(row_ptr, nbr, score, denom, mask_in if mask_in is not None else mask_out)
            ctx.meta = (tuple(table.shape), table.stride(0), F_cols, B, p, training, n_aug)
        ctx.mark_non_differentiable(*([mask_out] if mask_out is not None else []))
        return out, mask_out

    @staticmethod
    def backward(ctx, grad_out, _grad_mask):
        row_ptr, nbr, score, denom, mask = ctx.saved_tensors
        shape, ld, F_cols, B, p, training, n_aug = ctx.meta
        lib = _lib.load()
        grad_out = grad_out.contiguous()
        dev = grad_out.device
        if nbr is None:
            grad_table = torch.empty(shape, dtype=torch.float32, device=dev)
            if shape[1] != F_cols:
                grad_table.zero_()
        else:
            grad_table = torch.zeros(shape, dtype=torch.float32, device=dev)
        a = _lib.AggregateBwdArgs()
        a.grad_out = grad_out.data_ptr(); a.ld_grad_out = int(grad_out.stride(1)); a.denom = denom.data_ptr()
        a.row_ptr = row_ptr.data_ptr(); a.nbr = 0 if nbr is None else nbr.data_ptr(); a.score = score.data_ptr()
        a.B = int(B); a.n_entries = int(score.numel()); a.F = int(F_cols); a.p = float(p)
        a.training = int(bool(training)); a.n_aug = int(n_aug)
        a.mask_in = 0 if mask is None else mask.data_ptr()
        a.grad_table = grad_table.data_ptr(); a.ld_grad_table = int(grad_table.stride(0))
        a.n_table_rows = int(shape[0])
        _lib.check(lib.gp_aggregate_bwd(ctypes.byref(a), _stream(dev)))
        return (grad_table,) + (None,) * 13


class _SeedState:
    """Counter-based DropNode randomness: (seed, offset) fully determines every mask, so a run is
    reproducible and a mask can be regenerated (``dropnode_mask``) without storing it."""

    def __init__(self, seed: Optional[int] = None):
        self.seed = int(torch.initial_seed() if seed is None else seed) & (2**63 - 1)
        self.offset = 0

    def next(self) -> int:
        self.offset += 1
        return self.offset


_default_seed_state: Optional[_SeedState] = None


def _seed_state() -> _SeedState:
    global _default_seed_state
    if _default_seed_state is None or _default_seed_state.seed != (torch.initial_seed() & (2**63 - 1)):
        _default_seed_state = _SeedState()
    return _default_seed_state


def _finish(out, mask, n_aug_given, return_mask):
    res = out[0] if n_aug_given is None else out
    return (res, mask) if return_mask else res


def random_prop(feats, mat_scores, mat_idx, dropnode_rate, training=True, n_aug=None, seed=None, offset=None,
                mask=None, return_mask=False, check_index=True):
    """model.py:80-87 with pre-gathered ``feats`` [nz,F] (the reference's exact signature plus
    keyword extras).  Returns [B,F]; with ``n_aug=A`` returns [A,B,F] (A independent masks, one
    read of feats).  ``mask`` (uint8 [A,nz]) imports a mask, ``return_mask`` exports the one drawn."""
    _need_cuda(feats, mat_scores, mat_idx, mask)
    if feats.dtype != torch.float32:
        feats = feats.float()
    if feats.stride(-1) != 1:
        feats = feats.contiguous()
    score = mat_scores.to(torch.float32).contiguous()
    row_ptr, B = segments_from_sorted_index(mat_idx, check=check_index)
    A = 1 if n_aug is None else int(n_aug)
    if seed is None or offset is None:
        st = _seed_state()
        seed, offset = st.seed, st.next()
    if mask is not None:
        mask = mask.to(torch.uint8).reshape(A, -1).contiguous()
    out, m = _per_aug_group(A, mask, lambda a, c, mk: _AggregateFn.apply(
        feats, int(feats.shape[1]), row_ptr, None, score, B, float(dropnode_rate), bool(training), a, seed,
        _chunk_offset(offset, c), mk, EPS_RANDOM_PROP, bool(return_mask)))
    return _finish(out, m, n_aug, return_mask)


def random_prop_fused(features: DeviceFeatures, neighbor_idx, mat_scores, mat_idx=None, dropnode_rate=0.5,
                      training=True, n_aug=None, row_ptr=None, seed=None, offset=None, mask=None,
                      return_mask=False):
    """Fused replacement of ``features[neighbor_idx].to(device)`` (model.py:314) + ``random_prop``
    (model.py:322): gathers rows of the device-resident X inside the kernel.  Give either
    ``mat_idx`` (int64 ascending, as model.py:312 produces) or a ready ``row_ptr`` (int32 [B+1])."""
    _need_cuda(features.data, neighbor_idx, mat_scores, mat_idx, row_ptr, mask)
    nbr = narrow_index(neighbor_idx, features.N)
    score = mat_scores.to(torch.float32).contiguous()
    if row_ptr is None:
        row_ptr, B = segments_from_sorted_index(mat_idx)
    else:
        B = int(row_ptr.numel()) - 1
    A = 1 if n_aug is None else int(n_aug)
    if seed is None or offset is None:
        st = _seed_state()
        seed, offset = st.seed, st.next()
    if mask is not None:
        mask = mask.to(torch.uint8).reshape(A, -1).contiguous()
    out, m = _per_aug_group(A, mask, lambda a, c, mk: _AggregateFn.apply(
        features.data, features.F, row_ptr, nbr, score, B, float(dropnode_rate), bool(training), a, seed,
        _chunk_offset(offset, c), mk, EPS_RANDOM_PROP, bool(return_mask)))
    return _finish(out, m, n_aug, return_mask)


class _EmbFn(torch.autograd.Function):
    """MLP.emb (model_mag.py:48-55): out[n,:] = sum_a w_a drop(E[idx_a,:]) / (sum_a w_a + 1e-10), gradient w.r.t. the
    embedding table only.  ``sparse_grad``: the gradient is a coalesced sparse COO tensor over the rows the batch touched
    (what :class:`grandplus_b200.optim.SparseRowAdam` consumes) instead of a dense [n_rows, H] tensor."""

    @staticmethod
    def forward(ctx, weight, nbr, row_ptr, score, B, p, seed, offset, sparse_grad):
        lib = _lib.load()
        dev = weight.device
        H = int(weight.shape[1])
        needs_grad = weight.requires_grad
        if p > 0.0:
            ld_out = (H + 3) // 4 * 4
            out = torch.empty((B, ld_out), dtype=torch.float32, device=dev)
            denom = torch.empty((B,), dtype=torch.float32, device=dev)
            _lib.check(lib.gp_emb_dropout_fwd(_vp(weight), int(weight.shape[0]), int(weight.stride(0)), H, _vp(row_ptr), _vp(nbr),
                                              _vp(score), int(B), float(p), int(seed) & (2**64 - 1), int(offset) & (2**64 - 1),
                                              EPS_EMB, _vp(out), ld_out, _vp(denom), _stream(dev)))
            out = out if ld_out == H else out[:, :H]
        else:
            o3, _, denom = _launch_fwd(weight, H, weight.stride(0), row_ptr, None, 0, nbr, score, B, int(score.numel()), 0.0,
                                       False, 1, 0, 0, None, False, EPS_EMB, needs_grad)
            out = o3[0]
            denom = None if denom is None else denom[0]
        if needs_grad:
            ctx.save_for_backward(nbr, row_ptr, score, denom)
            ctx.meta = (tuple(weight.shape), H, int(B), float(p), int(seed), int(offset), bool(sparse_grad))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        nbr, row_ptr, score, denom = ctx.saved_tensors
        shape, H, B, p, seed, offset, sparse_grad = ctx.meta
        lib = _lib.load()
        dev = grad_out.device
        grad_out = grad_out.contiguous()
        if sparse_grad:
            rows, slot = torch.unique(nbr, return_inverse=True)     # the distinct table rows of the batch, ascending
            slot = slot.to(torch.int32).contiguous()
            target = torch.zeros((int(rows.numel()), H), dtype=torch.float32, device=dev)
        else:
            rows, slot = None, nbr
            target = torch.zeros(shape, dtype=torch.float32, device=dev)
        if p > 0.0:
            _lib.check(lib.gp_emb_dropout_bwd(_vp(grad_out), int(grad_out.stride(0)), H, _vp(row_ptr), _vp(slot), _vp(score),
                                              _vp(denom), B, p, seed & (2**64 - 1), offset & (2**64 - 1), _vp(target),
                                              int(target.stride(0)), _stream(dev)))
        else:
            a = _lib.AggregateBwdArgs()
            a.grad_out = grad_out.data_ptr(); a.ld_grad_out = int(grad_out.stride(0)); a.denom = denom.data_ptr()
            a.row_ptr = row_ptr.data_ptr(); a.nbr = slot.data_ptr(); a.score = score.data_ptr()
            a.B = B; a.n_entries = int(score.numel()); a.F = H; a.p = 0.0; a.training = 0; a.n_aug = 1; a.mask_in = 0
            a.grad_table = target.data_ptr(); a.ld_grad_table = int(target.stride(0)); a.n_table_rows = int(target.shape[0])
            _lib.check(lib.gp_aggregate_bwd(ctypes.byref(a), _stream(dev)))
        if sparse_grad:
            grad = torch.sparse_coo_tensor(rows.to(torch.int64)[None, :], target, size=shape, is_coalesced=True,
                                           check_invariants=False)
        else:
            grad = target
        return (grad,) + (None,) * 8


def emb(weight, attr_idx, node_idx, attr_data, input_droprate=0.0, training=False, sparse_grad=False, seed=None,
        offset=None):
    """model_mag.py:48-55 with the embedding table resident on the GPU (the reference keeps it on the CPU and copies the
    gathered rows every batch, model_mag.py:49).  Differentiable w.r.t. ``weight``.  A non-zero ``input_droprate`` in
    training mode is the element-wise dropout of model_mag.py:50, drawn from counter-based Philox inside the fused
    kernel (forward and backward regenerate the same mask from (seed, offset); :func:`emb_dropout_mask` exports it).
    ``sparse_grad=True`` keeps the gradient on the touched rows."""
    _need_cuda(weight, node_idx, attr_data)
    dev = weight.device
    attr_idx = attr_idx.to(dev)
    score = attr_data.to(torch.float32).contiguous()
    row_ptr, B = segments_from_sorted_index(node_idx)
    nbr = narrow_index(attr_idx, int(weight.shape[0]))
    w = weight if weight.stride(-1) == 1 else weight.contiguous()
    p = float(input_droprate) if training else 0.0
    if p > 0.0 and (seed is None or offset is None):
        st = _seed_state()
        seed, offset = st.seed, st.next()
    return _EmbFn.apply(w, nbr, row_ptr, score, B, p, seed or 0, offset or 0, bool(sparse_grad))


def emb_dropout_mask(nza: int, H: int, p: float, seed: int, offset: int, device) -> torch.Tensor:
    """The element keep-mask :func:`emb` draws for (seed, offset): uint8 [nza, H]."""
    lib = _lib.load()
    mask = torch.empty((nza, H), dtype=torch.uint8, device=device)
    _lib.check(lib.gp_emb_dropout_mask(int(nza), int(H), float(p), int(seed) & (2**64 - 1), int(offset) & (2**64 - 1),
                                       _vp(mask), _stream(mask.device)))
    return mask


def aggregate_slots(features: DeviceFeatures, col, val32, slot_rows=None, dropnode_rate=0.5, training=True,
                    n_aug=None, seed=None, offset=None, mask=None, return_mask=False):
    """Aggregate straight from GFPush's [S,K] output slots (col int32, val fp32) on the device:
    output row b reads slots ``slot_rows[b]*K .. +K`` (b itself when slot_rows is None); zero pads
    and rows with a negative slot id contribute nothing."""
    _need_cuda(features.data, col, val32, slot_rows, mask)
    S, K = int(col.shape[0]), int(col.shape[1])
    B = S if slot_rows is None else int(slot_rows.numel())
    A = 1 if n_aug is None else int(n_aug)
    if seed is None or offset is None:
        st = _seed_state()
        seed, offset = st.seed, st.next()
    if mask is not None:
        mask = mask.to(torch.uint8).reshape(A, -1).contiguous()
    out, m = _per_aug_group(A, mask, lambda a, c, mk: _launch_fwd(
        features.data, features.F, features.ld, None, slot_rows, K, col, val32, B, S * K, float(dropnode_rate),
        bool(training), a, seed, _chunk_offset(offset, c), mk, bool(return_mask), EPS_RANDOM_PROP, False)[:2])
    return _finish(out, m, n_aug, return_mask)


class PiMatrix:
    """Device-resident top-k propagation matrix.

    Built straight from GFPush's [S,K] output slots (row ``it`` <-> source ``node_idx[it]``), it
    replaces ``sp.coo_matrix(...).tocsr()`` (model.py:270-272) and the per-batch scipy slice +
    ``.nonzero()`` (model.py:310-313): a batch is a list of node ids, mapped to slot rows on the
    device, and the kernel walks the K slots of each row directly (zero pads skipped)."""

    def __init__(self, node_idx, col, val32, n_nodes: int):
        _need_cuda(node_idx, col, val32)
        self.K = int(col.shape[1])
        self.S = int(col.shape[0])
        self.col = col.contiguous()
        self.val = val32.to(torch.float32).contiguous()
        self.node_idx = node_idx.to(torch.int64)
        self.row_of_node = torch.full((n_nodes,), -1, dtype=torch.int32, device=col.device)
        self.row_of_node[self.node_idx] = torch.arange(self.S, dtype=torch.int32, device=col.device)

    @classmethod
    def from_graph(cls, graph, node_idx, coef, rmax, K, check=True):
        """Run GFPush on the device and keep the result there.  ``check`` (default) waits for the push and raises if the
        device refused a source id or a list outgrew its bound, like the host-buffer path does."""
        dev = torch.device("cuda", graph.device)
        nid = torch.as_tensor(node_idx).to(device=dev, dtype=torch.int32)
        _row, col, _val, val32 = graph.gfpush_device(nid, coef, rmax, K, want_fp32=True, check=check)
        return cls(nid, col, val32, graph.num_nodes)

    # -- cache on disk: the reference recomputes GFPush for every (seed1, seed2) run ---------
    @staticmethod
    def cache_key(indptr, indices, node_idx, coef, rmax, K) -> str:
        """Hash of everything Pi depends on (graph, sources, coef, rmax, K)."""
        import hashlib
        import numpy as np
        h = hashlib.sha256()
        for a, dt in ((indptr, np.int32), (indices, np.int32), (node_idx, np.int64), (coef, np.float64)):
            a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
            h.update(np.ascontiguousarray(a.astype(dt, copy=False)).tobytes())
        h.update(repr((float(rmax), int(K))).encode())
        return h.hexdigest()[:32]

    def save(self, path: str, key: str = "") -> None:
        import numpy as np
        np.savez(path, node_idx=self.node_idx.cpu().numpy(), col=self.col.cpu().numpy(), val=self.val.cpu().numpy(),
                 n_nodes=np.int64(self.row_of_node.numel()), key=np.array(key))

    @classmethod
    def load(cls, path: str, device=None, key: str = None):
        """The saved matrix, or None when ``key`` is given and does not match the file's."""
        import numpy as np
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        if key is not None and str(z["key"]) != key:
            return None
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        return cls(torch.from_numpy(z["node_idx"]).to(device), torch.from_numpy(z["col"]).to(device),
                   torch.from_numpy(z["val"]).to(device), int(z["n_nodes"]))

    def slot_rows(self, batch_nodes) -> torch.Tensor:
        b = torch.as_tensor(batch_nodes).to(device=self.col.device, dtype=torch.int64)
        rows = self.row_of_node[b]
        return rows.contiguous()

    def aggregate(self, features: DeviceFeatures, batch_nodes, dropnode_rate=0.5, training=True, n_aug=None,
                  seed=None, offset=None, mask=None, return_mask=False, slot_rows=None, validate=False):
        """[B,F] (or [A,B,F]) aggregated features of ``batch_nodes`` -- model.py:310-322 in one launch."""
        rows = self.slot_rows(batch_nodes) if slot_rows is None else slot_rows
        if validate and bool((rows < 0).any().item()):
            raise IndexError("batch contains a node that is not a GFPush source")
        return aggregate_slots(features, self.col, self.val, rows, dropnode_rate, training, n_aug, seed, offset,
                               mask, return_mask)

    def to_scipy(self, n_nodes=None):
        """The host CSR the reference builds (model.py:270-272), for interop and tests."""
        import numpy as np
        import scipy.sparse as sp
        n = int(self.row_of_node.numel()) if n_nodes is None else n_nodes
        col = self.col.cpu().numpy().reshape(-1)
        val = self.val.double().cpu().numpy().reshape(-1)
        row = np.repeat(self.node_idx.cpu().numpy(), self.K)
        row = np.where(val > 0, row, 0)
        return sp.coo_matrix((val, (row, col)), (n, n)).tocsr()


class Grand_Plus(nn.Module):
    """Same constructor idea and methods as the reference's ``Grand_Plus`` (model.py:70-87): wraps an
    MLP (kept in PyTorch, north_star) and exposes ``random_prop`` with the reference's signature."""

    def __init__(self, mlp: nn.Module, dropnode_rate: float = 0.5, seed: Optional[int] = None):
        super().__init__()
        self.mlp = mlp
        self.dropnode_rate = dropnode_rate
        self._seeds = _SeedState(seed)

    def forward(self, X):
        return self.mlp(X)

    def random_prop(self, feats, mat_scores, mat_idx, dropnode_rate, **kw):
        kw.setdefault("seed", self._seeds.seed)
        kw.setdefault("offset", self._seeds.next())
        return random_prop(feats, mat_scores, mat_idx, dropnode_rate, training=self.training, **kw)

    def random_prop_fused(self, features, neighbor_idx, mat_scores, mat_idx=None, dropnode_rate=None, **kw):
        kw.setdefault("seed", self._seeds.seed)
        kw.setdefault("offset", self._seeds.next())
        p = self.dropnode_rate if dropnode_rate is None else dropnode_rate
        return random_prop_fused(features, neighbor_idx, mat_scores, mat_idx, p, training=self.training, **kw)

    def emb(self, attr_idx, node_idx, attr_data, cuda=True):
        """model_mag.py:88-90; the MLP must expose ``embeds`` (nn.Embedding on the GPU) and
        ``input_droprate`` like the reference's MAG MLP (model_mag.py:21-36)."""
        return emb(self.mlp.embeds.weight, attr_idx, node_idx, attr_data,
                   getattr(self.mlp, "input_droprate", 0.0), self.training)
