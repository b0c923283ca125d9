"""oracle.aggregate -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's per-batch Pi.X aggregation:

* :func:`random_prop`  <- ``Grand_Plus.random_prop`` (/root/reference/model.py:80-87,
  identical copy at /root/reference/model_mag.py:80-86)
* :func:`emb`          <- ``MLP.emb`` (/root/reference/model_mag.py:48-55)
* :func:`batch_slice`  <- the batch assembly around them (/root/reference/model.py:310-316)

``torch_scatter`` 2.0.6 (requirements.txt:7) is a third-party dependency that is not in
/root/reference and not installable here.  Its published ``scatter(src, index, dim=0,
dim_size, reduce='sum')`` is ``zeros(dim_size).scatter_add_(0, broadcast(index), src)``;
that is what is restated (a sequential fp32 accumulation in entry order).

Parity status: PINNED against tests/golden/random_prop.npz and tests/golden/emb.npz, which
tests/golden/make_golden.py produced by executing the reference's own ``random_prop`` /
``emb`` source with a scatter_add_ stand-in for torch_scatter (see that script's header).
``F.dropout``'s mask layout is a PyTorch implementation detail, so every function here
takes the keep-mask as an input ("given identical masks", BASELINE.json north_star).
"""
from __future__ import annotations

import numpy as np


def dropout_scores(scores: np.ndarray, p: float, training: bool, mask: np.ndarray | None) -> np.ndarray:
    """F.dropout on the per-entry scores (model.py:82): kept entries become score * 1/(1-p),
    dropped entries 0, identity in eval mode.  fp32 like ATen (noise = mask / (1-p); x * noise)."""
    scores = np.asarray(scores, dtype=np.float32)
    if not training or p == 0.0:
        return scores.copy()
    if p >= 1.0:
        return np.zeros_like(scores)
    assert mask is not None, "training-mode parity needs the keep-mask"
    noise = mask.astype(np.float32) / np.float32(1.0 - p)
    return (scores * noise).astype(np.float32)


def _segment_sum(src: np.ndarray, idx: np.ndarray, dim_size: int, dtype) -> np.ndarray:
    out = np.zeros((dim_size,) + src.shape[1:], dtype=dtype)
    np.add.at(out, idx, src.astype(dtype))       # sequential, in entry order, like scatter_add_ on CPU
    return out


def random_prop(feats, scores, idx, p, training, mask=None, dtype=np.float32):
    """model.py:80-87.  feats [nz,F] fp32, scores [nz] fp32, idx [nz] int64 ascending.
    dim_size = idx[-1] + 1 (model.py:84).  dtype=np.float64 gives the tie-breaker version."""
    feats = np.asarray(feats, dtype=np.float32)
    idx = np.asarray(idx, dtype=np.int64)
    m = dropout_scores(scores, p, training, mask)                         # model.py:82
    dim_size = int(idx[-1]) + 1
    if dtype == np.float32:
        weighted = (feats * m[:, None]).astype(np.float32)                # model.py:83 temporary
    else:
        weighted = feats.astype(dtype) * m.astype(dtype)[:, None]
    num = _segment_sum(weighted, idx, dim_size, dtype)                    # model.py:83-84
    den = _segment_sum(m[:, None], idx, dim_size, dtype)                  # model.py:85-86
    return (num / (den + dtype(1e-12))).astype(dtype)                     # model.py:87


def emb(table, attr_idx, node_idx, attr_data, dtype=np.float32, elem_mask=None, input_droprate=0.0):
    """model_mag.py:48-55.  table [n_attr,H] fp32; attr_idx/node_idx [nza] int64 (node_idx ascending);
    attr_data [nza] fp32.  elem_mask [nza,H] (1 keep) is the input-dropout mask in training mode."""
    table = np.asarray(table, dtype=np.float32)
    node_idx = np.asarray(node_idx, dtype=np.int64)
    attr_data = np.asarray(attr_data, dtype=np.float32)
    E = table[np.asarray(attr_idx, dtype=np.int64)]                       # model_mag.py:49
    if elem_mask is not None and input_droprate > 0.0:                    # model_mag.py:50
        E = (E * (elem_mask.astype(np.float32) / np.float32(1.0 - input_droprate))).astype(np.float32)
    dim_size = int(node_idx[-1]) + 1                                      # model_mag.py:51
    if dtype == np.float32:
        weighted = (E * attr_data[:, None]).astype(np.float32)
    else:
        weighted = E.astype(dtype) * attr_data.astype(dtype)[:, None]
    num = _segment_sum(weighted, node_idx, dim_size, dtype)               # model_mag.py:52
    den = _segment_sum(attr_data[:, None], node_idx, dim_size, dtype)     # model_mag.py:53
    return (num / (den + dtype(1e-10))).astype(dtype)                     # model_mag.py:54


def topk_adj_from_slots(row_idx, col_idx, value, n):
    """model.py:270-272: COO -> CSR with duplicate summation (pads collapse into (0,0))."""
    import scipy.sparse as sp
    return sp.coo_matrix((value, (row_idx, col_idx)), (n, n)).tocsr()


def batch_slice(topk_adj, batch_index):
    """model.py:310-316 minus the device copies: (source_idx int64, neighbor_idx, mat_scores fp32)."""
    sub = topk_adj[batch_index]
    source_idx, neighbor_idx = sub.nonzero()
    return source_idx.astype(np.int64), neighbor_idx.astype(np.int64), sub.data.astype(np.float32)


def random_prop_backward_feats(grad_out, scores, idx, p, training, mask=None):
    """d out / d feats for the MAG path (model_mag.py:356 does not detach):
    grad_feats[j,:] = m_j / (sum_row m + 1e-12) * grad_out[idx_j,:]  (fp64 reference)."""
    idx = np.asarray(idx, dtype=np.int64)
    m = dropout_scores(scores, p, training, mask).astype(np.float64)
    den = _segment_sum(m[:, None], idx, int(idx[-1]) + 1, np.float64) + 1e-12
    return (m / den[idx, 0])[:, None] * np.asarray(grad_out, dtype=np.float64)[idx]


def emb_backward_table(grad_node, n_attr, attr_idx, node_idx, attr_data, elem_mask=None, input_droprate=0.0):
    """d out / d table for MLP.emb: dense [n_attr,H] fp64 gradient (elem_mask [nza,H]: the input-dropout mask of
    model_mag.py:50 in training mode, through which the gradient flows scaled by 1/(1-p))."""
    node_idx = np.asarray(node_idx, dtype=np.int64)
    w = np.asarray(attr_data, dtype=np.float64)
    den = _segment_sum(w[:, None], node_idx, int(node_idx[-1]) + 1, np.float64) + 1e-10
    contrib = (w / den[node_idx, 0])[:, None] * np.asarray(grad_node, dtype=np.float64)[node_idx]
    if elem_mask is not None and input_droprate > 0.0:
        contrib = contrib * (np.asarray(elem_mask, dtype=np.float64) / (1.0 - input_droprate))
    out = np.zeros((n_attr, contrib.shape[1]), dtype=np.float64)
    np.add.at(out, np.asarray(attr_idx, dtype=np.int64), contrib)
    return out
