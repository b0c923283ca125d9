"""Exact (non-GFPush) propagation for inference -- the reference's ``predict`` / ``get_local_logits``
(/root/reference/model.py:169-224, /root/reference/model_mag.py:180-245) with the ``order`` rounds of
``D^-1 A H`` over ALL nodes on the GPU instead of scipy on the host.

One round ``H <- diag(c / deg) (A H)`` is the same gather - scale - reduce the per-batch aggregation does
(``out[i,:] = sum_j w_ij H[j,:] / (sum_j w_ij + eps)``), so it runs on ``gp_aggregate_fwd`` with the CSR of
``adj`` as the entry list: rows are owned by warps, no atomics, H is read once per edge with 128-bit loads
and stays in HBM between rounds.  SURVEY.md 8(f) rank 2.
"""
from __future__ import annotations

import numpy as np
import torch

from . import model as _m

__all__ = ["DeviceAdjacency", "propagate_exact", "get_local_logits", "predict"]


HUB_DEGREE = 1024   # rows with more entries than this are reduced in chunks ...
HUB_CHUNK = 512     # ... of this many entries, then combined


class DeviceAdjacency:
    """``adj`` (model.py:243: adjacency + I, any scipy sparse format or (indptr, indices[, data])) as a
    device CSR: ``indptr`` int32 [N+1], ``indices`` int32 [nnz], ``data`` fp32 [nnz] (ones when absent).

    Power-law graphs have rows of 10^4..10^5 entries (75 K on the Reddit-shape graph, 241 K on the
    Amazon2M-shape one).  One warp walking such a row is a long serial tail and a long fp32 summation, so rows
    above ``HUB_DEGREE`` are split ONCE, here, into chunks of ``HUB_CHUNK`` entries:
      pass A reduces every chunk into an extra table row  N + c  (normalised by the chunk's weight W_c),
      pass B reduces every row; a hub row's entry list is its chunk rows with scores W_c (+eps), so
             sum_c W_c * (S_c / W_c) / sum_c W_c  ==  S / W.
    Both passes are the same kernel (gp_aggregate_fwd); all index structures are static per graph."""

    def __init__(self, adj, device=None, hub_degree: int = HUB_DEGREE, hub_chunk: int = HUB_CHUNK):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if isinstance(adj, (tuple, list)):
            indptr, indices = adj[0], adj[1]
            data = adj[2] if len(adj) > 2 else None
        else:
            csr = adj.tocsr()
            indptr, indices, data = csr.indptr, csr.indices, csr.data
        to = lambda a, dt: torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(device=device, dtype=dt).contiguous()  # noqa: E731
        self.indptr = to(indptr, torch.int32)
        self.indices = to(indices, torch.int32)
        self.data = torch.ones(self.indices.numel(), dtype=torch.float32, device=device) if data is None else to(data, torch.float32)
        self.N = int(self.indptr.numel()) - 1
        self.nnz = int(self.indices.numel())
        if self.N < 1 or int(self.indptr[-1].item()) != self.nnz:
            raise ValueError("malformed CSR: indptr[-1] must equal nnz")
        self.n_chunks = 0
        self._split(int(hub_degree), int(hub_chunk))

    @property
    def device(self):
        return self.indptr.device

    def _split(self, hub_degree, hub_chunk):
        dev = self.device
        ptr = self.indptr.to(torch.int64)
        deg = ptr[1:] - ptr[:-1]
        hub = torch.nonzero(deg > hub_degree).flatten()
        if hub.numel() == 0:
            return
        hdeg = deg[hub]
        nch = (hdeg + hub_chunk - 1) // hub_chunk                      # chunks per hub row
        first = torch.cumsum(nch, 0) - nch                              # first chunk id of each hub
        nc = int(nch.sum().item())
        owner = torch.repeat_interleave(torch.arange(hub.numel(), device=dev), nch)   # chunk -> hub ordinal
        k = torch.arange(nc, device=dev) - first[owner]                 # chunk number within its row
        c_start = ptr[hub][owner] + k * hub_chunk                       # entry range of the chunk in the CSR
        c_end = torch.minimum(c_start + hub_chunk, ptr[hub + 1][owner])
        c_len = c_end - c_start
        # pass A: the hub rows' entries, compacted, one row per chunk
        a_ptr = torch.zeros(nc + 1, dtype=torch.int64, device=dev)
        a_ptr[1:] = torch.cumsum(c_len, 0)
        src = torch.repeat_interleave(c_start - a_ptr[:-1], c_len) + torch.arange(int(a_ptr[-1].item()), device=dev)
        self.a_ptr = a_ptr.to(torch.int32)
        self.a_idx = self.indices[src].contiguous()
        self.a_dat = self.data[src].contiguous()
        w = torch.zeros(nc, dtype=torch.float64, device=dev).index_add_(0, torch.repeat_interleave(torch.arange(nc, device=dev), c_len),
                                                                         self.a_dat.double())
        # pass B: ordinary rows keep their entries, hub rows list their chunk rows N + c with score W_c
        is_hub = torch.zeros(self.N, dtype=torch.bool, device=dev)
        is_hub[hub] = True
        b_len = torch.where(is_hub, torch.zeros_like(deg), deg)
        b_len[hub] = nch
        b_ptr = torch.zeros(self.N + 1, dtype=torch.int64, device=dev)
        b_ptr[1:] = torch.cumsum(b_len, 0)
        tot = int(b_ptr[-1].item())
        b_idx = torch.empty(tot, dtype=torch.int32, device=dev)
        b_dat = torch.empty(tot, dtype=torch.float32, device=dev)
        keep = torch.repeat_interleave(~is_hub, deg)                    # CSR entries of ordinary rows, in order
        row_of = torch.repeat_interleave(torch.arange(self.N, device=dev), b_len)
        ordinary = ~is_hub[row_of]
        b_idx[ordinary] = self.indices[keep]
        b_dat[ordinary] = self.data[keep]
        b_idx[~ordinary] = (self.N + torch.arange(nc, device=dev)).to(torch.int32)   # chunks are numbered in hub order
        b_dat[~ordinary] = (w + _m.EPS_RANDOM_PROP).to(torch.float32)
        self.b_ptr, self.b_idx, self.b_dat = b_ptr.to(torch.int32), b_idx, b_dat
        self.n_chunks = nc


def _round(adj: DeviceAdjacency, H: torch.Tensor, F: int, out: torch.Tensor = None) -> torch.Tensor:
    """``diag(1/deg) (A H)`` with deg = adj.sum(1) (model.py:189,191).  H and the result are [N + n_chunks, ld] fp32
    buffers (ld % 4 == 0) whose first N rows hold the matrix; the extra rows are scratch for the hub chunks."""
    rows, ld = adj.N + adj.n_chunks, int(H.shape[1])
    assert H.shape[0] == rows and H.is_contiguous()
    if out is None:
        out = torch.empty((rows, ld), dtype=torch.float32, device=H.device)
    args = (0.0, False, 1, 0, 0, None, False, _m.EPS_RANDOM_PROP, False)
    if adj.n_chunks:
        _m._launch_fwd(H, F, ld, adj.a_ptr, None, 0, adj.a_idx, adj.a_dat, adj.n_chunks, int(adj.a_idx.numel()), *args,
                       out=H[adj.N:])        # chunk sums go beside the iterate they were computed from
        _m._launch_fwd(H, F, ld, adj.b_ptr, None, 0, adj.b_idx, adj.b_dat, adj.N, int(adj.b_idx.numel()), *args,
                       out=out[:adj.N])
    else:
        _m._launch_fwd(H, F, ld, adj.indptr, None, 0, adj.indices, adj.data, adj.N, adj.nnz, *args, out=out[:adj.N])
    return out


def propagate_exact(adj, features, order: int, alpha: float = 0.2, mode: str = "ppr") -> torch.Tensor:
    """The propagated feature matrix ``predict`` feeds to the MLP (model.py:184-212), [N, F] fp32 on the GPU.

    ppr:    sum_{i<=order} alpha (1-alpha)^i (D^-1 A)^i X      avg: mean_{i<=order} (D^-1 A)^i X
    single: (D^-1 A)^order X
    (The reference iterates in fp64 on the host and accumulates into fp32; here both are fp32 on the device.)"""
    if not isinstance(adj, DeviceAdjacency):
        adj = DeviceAdjacency(adj)
    X = features if isinstance(features, _m.DeviceFeatures) else _m.DeviceFeatures(features, adj.device)
    if X.N != adj.N:
        raise ValueError(f"features has {X.N} rows, adj has {adj.N}")
    F, N = X.F, adj.N
    H = torch.empty((N + adj.n_chunks, X.ld), dtype=torch.float32, device=adj.device)
    H[:N] = X.data
    nxt = torch.empty_like(H)
    if mode == "ppr":
        H[:N] *= float(alpha)                     # model.py:185
        prop = H[:N].clone()                      # :186
        for _ in range(int(order)):
            _round(adj, H, F, nxt)
            H, nxt = nxt, H
            H[:N] *= 1.0 - float(alpha)           # :188,190  (1-alpha)/deg
            prop += H[:N]                         # :191
    elif mode == "avg":
        prop = H[:N].clone()
        for _ in range(int(order)):
            _round(adj, H, F, nxt)
            H, nxt = nxt, H
            prop += H[:N]
        prop /= float(order + 1)                  # :200
    elif mode == "single":
        for _ in range(int(order)):
            _round(adj, H, F, nxt)
            H, nxt = nxt, H
        prop = H[:N].clone()
    else:
        raise ValueError(f"Unknown propagation mode: {mode}")   # model.py:211
    return prop[:, :F]


def get_local_logits(model, attr_mat: torch.Tensor, batch_size: int = 10000) -> torch.Tensor:
    """model.py:169-178 with ``attr_mat`` already on the device: the MLP over row chunks, no host round trip."""
    logits = []
    with torch.no_grad():
        for i in range(0, attr_mat.shape[0], batch_size):
            logits.append(model(attr_mat[i:i + batch_size].contiguous()))
    return torch.cat(logits, 0)


def predict(args, adj, features_np, model, idx_test, labels_org, mode="ppr", batch_size_logits=10000):
    """Drop-in for model.py:181-224: same arguments, same printed and returned test accuracy."""
    model.eval()
    dev = next(model.parameters()).device
    feat = propagate_exact(DeviceAdjacency(adj, dev), _m.DeviceFeatures(features_np, dev), args.order, args.alpha, mode)
    logits = get_local_logits(model.mlp, feat, batch_size_logits)
    preds = logits.argmax(1)
    idx = torch.as_tensor(np.asarray(idx_test), device=dev, dtype=torch.long)
    labels = torch.as_tensor(labels_org).to(dev)
    acc_test = float((preds[idx] == labels[idx]).double().sum().item()) / len(idx_test)
    print(acc_test)
    return acc_test
