"""Optimizer step of the MAG embedding table on the touched rows only.

The reference builds ``nn.Embedding(2 784 240, hidden)`` with the default ``sparse=False`` (/root/reference/model_mag.py:27)
and hands it to ``torch.optim.Adam(model.parameters(), lr, weight_decay)`` (model_mag.py:312-313): every step zero-fills and
reduces a dense 713 MB gradient and walks the whole table (model_mag.py:367-369), although a batch touches a few thousand
rows.  :class:`SparseRowAdam` keeps the gradient on those rows and still produces the reference's parameters: under dense
Adam a row that receives no gradient keeps moving while its moments decay, so every row remembers the step it was last
brought up to date and the skipped zero-gradient steps are replayed (in registers, ``gp_lazy_adam_rows``) the next time
the row is read or updated.  ``weight_decay`` must be 0, as in ``scripts/run_mag.sh:7``; with a non-zero decay every row
moves every step and the dense optimizer is the right tool.

    opt = SparseRowAdam(model.mlp.embeds.weight, lr=args.lr)
    ...
    opt.prepare(attr_idx)                                    # before the forward pass reads the rows
    out = gm.emb(weight, attr_idx, node_idx, attr_data, sparse_grad=True)
    loss.backward(); opt.step()                              # weight.grad is a coalesced sparse COO tensor
    ...
    opt.flush()                                              # before inference over the whole table / saving
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _vp(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class SparseRowAdam:
    def __init__(self, weight: torch.Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        if weight_decay != 0.0:
            raise ValueError("SparseRowAdam reproduces dense Adam only for weight_decay = 0 (scripts/run_mag.sh:7)")
        if not weight.is_cuda or weight.dtype != torch.float32 or weight.dim() != 2 or weight.stride(1) != 1:
            raise ValueError("weight must be a 2-D fp32 CUDA tensor with unit column stride")
        self.weight = weight
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = torch.zeros_like(weight)
        self.exp_avg_sq = torch.zeros_like(weight)
        self.last_step = torch.zeros(weight.shape[0], dtype=torch.int32, device=weight.device)
        self.step_count = 0          # dense-Adam steps taken so far
        self._lib = _lib.load()

    def _launch(self, rows, grad_rows):
        w = self.weight
        R = int(w.shape[0]) if rows is None else int(rows.numel())
        st = ctypes.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)
        _lib.check(self._lib.gp_lazy_adam_rows(
            _vp(w.data), _vp(self.exp_avg), _vp(self.exp_avg_sq), _vp(self.last_step), int(w.stride(0)), int(w.shape[1]),
            _vp(rows), R, _vp(grad_rows), 0 if grad_rows is None else int(grad_rows.stride(0)), int(self.step_count),
            self.lr, self.betas[0], self.betas[1], self.eps, st))

    @torch.no_grad()
    def prepare(self, rows: torch.Tensor) -> None:
        """Bring the rows a forward pass is about to read up to date with the steps taken so far."""
        rows = torch.unique(rows.to(device=self.weight.device, dtype=torch.int64))
        self._launch(rows.contiguous(), None)

    @torch.no_grad()
    def flush(self) -> None:
        """Bring every row up to date (before inference over the whole table, or saving the model)."""
        self._launch(None, None)

    def zero_grad(self) -> None:
        self.weight.grad = None

    @torch.no_grad()
    def step(self, rows: torch.Tensor = None, grad_rows: torch.Tensor = None) -> None:
        """One Adam step.  By default reads ``weight.grad`` (a sparse COO tensor from ``emb(..., sparse_grad=True)``);
        (rows, grad_rows) can be given directly, e.g. after :func:`grandplus_b200.dist.allreduce_sparse_rows`."""
        if rows is None:
            g = self.weight.grad
            if g is None:
                raise RuntimeError("weight.grad is empty")
            if not g.is_sparse:
                raise RuntimeError("SparseRowAdam needs a sparse gradient: call emb(..., sparse_grad=True)")
            g = g if g.is_coalesced() else g.coalesce()
            rows, grad_rows = g._indices()[0], g._values()
        rows = rows.to(torch.int64).contiguous()
        grad_rows = grad_rows.to(torch.float32).contiguous()
        self._launch(rows, grad_rows)
        self.step_count += 1
