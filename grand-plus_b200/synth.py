"""Synthetic inputs of the BASELINE.json shapes: Chung-Lu power-law graphs with the
self-loops of /root/reference/model.py:243, dense N(0,1) features, source samples.

One torch implementation that runs on the CPU (tests) or on the GPU (bench); all randomness comes
from a counter-based integer hash, so a (shape, seed) names the same edge draws on either device
up to float rounding of the inverse-CDF lookup.
"""
from __future__ import annotations

import numpy as np
import torch

# named shapes: (nodes, undirected edge draws, feature width) -- BASELINE.json configs
SHAPES = {
    "reddit": (232_965, 11_606_919, 602),
    "amazon2m": (2_449_029, 61_859_140, 100),
    "mag": (10_541_560, 265_219_994, 64),
    "small": (50_000, 600_000, 64),
}

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic)."""
    z = x + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def uniform01(n: int, seed: int, stream: int, device) -> torch.Tensor:
    """n float64 uniforms in [0,1) from counters (seed, stream, i)."""
    i = torch.arange(n, dtype=torch.int64, device=device)
    h = splitmix64(i + _s64(seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03))
    return _lsr(h, 11).to(torch.float64) * (1.0 / (1 << 53))


def powerlaw_csr(n: int, n_draws: int, gamma: float = 2.5, seed: int = 0, device="cpu", chunk: int = 1 << 26,
                 relabel: bool = True):
    """Chung-Lu graph: weights w_i = (i+1)^(-1/(gamma-1)), `n_draws` endpoint pairs by inverse CDF,
    random relabelling, self pairs dropped, symmetrised, de-duplicated, + I, sorted.
    Returns (indptr int32 [n+1], indices int32 [nnz]) torch tensors on `device`."""
    dev = torch.device(device)
    w = torch.arange(1, n + 1, dtype=torch.float64, device=dev).pow_(-1.0 / (gamma - 1.0))
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    del w
    perm = torch.argsort(splitmix64(torch.arange(n, dtype=torch.int64, device=dev) + _s64(seed * 77 + 12345)))
    if not relabel:  # ids stay sorted by expected degree (locality experiment)
        perm = torch.arange(n, dtype=torch.int64, device=dev)
    keys = []
    for lo in range(0, n_draws, chunk):
        m = min(chunk, n_draws - lo)
        i = torch.arange(lo, lo + m, dtype=torch.int64, device=dev)
        base = _s64(seed * 0x9E3779B97F4A7C15)
        ua = _lsr(splitmix64(i * 2 + base), 11).to(torch.float64) * (1.0 / (1 << 53))
        ub = _lsr(splitmix64(i * 2 + 1 + base), 11).to(torch.float64) * (1.0 / (1 << 53))
        a = perm[torch.searchsorted(cdf, ua).clamp_(max=n - 1)]
        b = perm[torch.searchsorted(cdf, ub).clamp_(max=n - 1)]
        keep = a != b
        a, b = a[keep], b[keep]
        keys.append(a * n + b)
        keys.append(b * n + a)
        del i, ua, ub, a, b, keep
    d = torch.arange(n, dtype=torch.int64, device=dev)
    keys.append(d * n + d)  # model.py:243: adj + I
    del cdf, perm
    key = torch.unique(torch.cat(keys))
    del keys
    rows = torch.div(key, n, rounding_mode="floor")
    cols = (key - rows * n).to(torch.int32)
    del key
    counts = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    if int(indptr[-1]) >= 2**31:
        raise ValueError("graph exceeds int32 CSR")
    return indptr.to(torch.int32), cols


def features(n: int, f: int, seed: int = 1, device="cpu") -> torch.Tensor:
    """X ~ N(0,1) fp32 [n, f] (Box-Muller on hashed uniforms: same values on CPU and GPU)."""
    dev = torch.device(device)
    tot = n * f
    u1 = uniform01(tot, seed, 1, dev).clamp_(min=2.0 ** -53)
    u2 = uniform01(tot, seed, 2, dev)
    x = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * np.pi * u2)
    return x.to(torch.float32).reshape(n, f)


def sources(n: int, s: int, seed: int = 1, device="cpu") -> torch.Tensor:
    """S distinct source nodes (int32), a seeded sample without replacement."""
    dev = torch.device(device)
    order = torch.argsort(splitmix64(torch.arange(n, dtype=torch.int64, device=dev) + _s64(seed * 1315423911)))
    return order[:s].to(torch.int32)
