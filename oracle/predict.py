"""oracle.predict -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/scipy restatement of the exact propagation inside the reference's ``predict``
(/root/reference/model.py:184-212; the MAG copy /root/reference/model_mag.py:212-234 is the same
arithmetic on embeddings).  Line by line, including the dtypes the reference ends up with:
``deg_row_inv`` is float64, so the iterate becomes float64 after the first round while the
accumulator ``features_np_prop`` stays in the dtype of the input (in-place ``+=``).

Parity status: PINNED against tests/golden/predict.npz, produced by tests/golden/make_golden.py by
calling the reference's own ``predict`` with ``get_local_logits`` intercepted to capture the matrix it
is handed.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def propagate_exact(adj, features_np: np.ndarray, order: int, alpha: float, mode: str) -> np.ndarray:
    adj = sp.csr_matrix(adj)
    nprop = int(order)
    if mode == "ppr":
        features_np = alpha * features_np                                        # model.py:185
        prop = features_np.copy()                                                # :186
        deg_row = np.asarray(adj.sum(1)).ravel()                                 # :187
        inv = np.asarray((1 - alpha) / np.maximum(deg_row, 1e-12))               # :188
        for _ in range(nprop):
            features_np = np.multiply(inv[:, None], adj.dot(features_np))       # :190
            prop += features_np                                                  # :191
        return prop
    if mode == "avg":
        prop = features_np.copy()                                                # :194
        deg_row = np.asarray(adj.sum(1)).ravel()
        inv = 1 / np.maximum(deg_row, 1e-12)                                     # :196
        for _ in range(nprop):
            features_np = np.multiply(inv[:, None], adj.dot(features_np))       # :198
            prop += features_np
        return prop / (nprop + 1)                                                # :200
    if mode == "single":
        deg_row = np.asarray(adj.sum(1)).ravel()
        inv = 1 / np.maximum(deg_row, 1e-12)
        for _ in range(nprop):
            features_np = np.multiply(inv[:, None], adj.dot(features_np))       # :206
        return features_np
    raise ValueError(f"Unknown propagation mode: {mode}")                        # :211
