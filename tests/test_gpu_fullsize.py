"""GPU tests at BASELINE.json's FULL sizes (configs[2], the Reddit-shape graph bench.py times): the oracle on a
sample of rows, and size-independent properties on all of them -- agreement between the residue-table tiers,
idempotence, mass bounds, the work bound (L-1)/rmax, linearity of the aggregation in X."""
import numpy as np
import pytest

from oracle import gfpush as og
from tests.helpers import check_topk_rows

pytestmark = pytest.mark.gpu

N, DRAWS, F = 232_965, 11_606_919, 602          # bench.WORKLOADS["reddit"]
ORDER, ALPHA, RMAX, K = 6, 0.05, 1e-5, 32       # scripts/run_reddit.sh:7, K per BASELINE.json


@pytest.fixture(scope="module")
def reddit_shape():
    import torch
    from grandplus_b200 import synth
    from grandplus_b200.precompute import propagation
    indptr, indices = synth.powerlaw_csr(N, DRAWS, seed=0, device="cuda")
    graph = propagation.Graph.from_device_csr(indptr, indices)
    src = synth.sources(N, 4096, seed=7, device="cuda")
    hub = (indptr[1:] - indptr[:-1]).argmax().to(torch.int32)
    src[0] = hub                                  # a 75 K-degree source: level 0 is expanded by the whole CTA
    return indptr, indices, graph, src.contiguous()


def _rows(col, val):
    return og.rows_as_sets(col.cpu().numpy().reshape(-1), val.cpu().numpy().reshape(-1), K)


def test_gfpush_full_size_tiers_agree_and_match_oracle(reddit_shape):
    import torch
    from grandplus_b200 import _lib
    indptr, indices, graph, src = reddit_shape
    coef = og.coef_for("ppr", ORDER, ALPHA)
    out = {}
    try:
        for tier, mode in (("table", 2), ("slabs", 0)):
            _lib.set_tuning("push_smem_hash", mode)
            graph.cumulative_stats(reset=True)
            row, col, val, _ = graph.gfpush_device(src, coef, RMAX, K, want_fp32=True)
            torch.cuda.synchronize()
            out[tier] = (row, col, val, graph.cumulative_stats())
    finally:
        _lib.set_tuning("push_smem_hash", 1)
    (_, ca, va, sa), (_, cb, vb, sb) = out["table"], out["slabs"]
    # work counters are integers of the algorithm: both tiers count the same pushes, frontiers and supports
    for k in ("edges_pushed", "frontier_total", "support_total", "sources"):
        assert abs(sa[k] - sb[k]) <= 1e-6 * sb[k], (k, sa[k], sb[k])
    # the work bound of SURVEY 7: a level pushes at most 1/rmax edges (+ the source's own degree at level 0)
    assert sa["edges_pushed"] <= len(src) * ((ORDER - 1) / RMAX + 80_000)
    same = 0
    for (ac, av), (bc, bv) in zip(_rows(ca, va), _rows(cb, vb)):
        assert len(ac) == len(bc)
        if np.array_equal(ac, bc):
            same += 1
            np.testing.assert_allclose(av, bv, rtol=1e-11, atol=0)
        else:   # only a tie at the cut may differ
            assert abs(av.min() - bv.min()) <= 1e-9 * bv.min()
    assert same >= 0.85 * len(src)   # exact ties at the cut are common (symmetric neighbourhoods): ~8 % of rows
    # mass: a row never sums above 1 (coef sums to 1, pushes only lose mass) and keeps at least coef[0] on the source
    sums = va.reshape(-1, K).sum(1)
    assert float(sums.max()) <= 1.0 + 1e-12 and float(sums.min()) >= coef[0] * (1 - 1e-12)
    # the oracle itself on a sample of rows (hub source included)
    ip, ix = indptr.cpu().numpy(), indices.cpu().numpy()
    worst = check_topk_rows(ip, ix, src.cpu().numpy(), coef, RMAX, K, ca.cpu().numpy().reshape(-1),
                            va.cpu().numpy().reshape(-1), row=out["table"][0].cpu().numpy().reshape(-1), max_rows=12)
    assert worst < 1e-11


def test_gfpush_full_size_idempotent(reddit_shape):
    import torch
    _, _, graph, src = reddit_shape
    coef = og.coef_for("ppr", ORDER, ALPHA)
    _, c1, v1, _ = graph.gfpush_device(src, coef, RMAX, K, want_fp32=True)
    perm = torch.randperm(len(src), device="cuda")
    _, c2, v2, _ = graph.gfpush_device(src[perm].contiguous(), coef, RMAX, K, want_fp32=True)
    a, b = _rows(c1, v1), _rows(c2[torch.argsort(perm)], v2[torch.argsort(perm)])
    diff = 0
    for (ac, av), (bc, bv) in zip(a, b):
        if np.array_equal(ac, bc):
            np.testing.assert_allclose(av, bv, rtol=1e-12, atol=0)
        else:
            diff += 1
    assert diff <= 0.15 * len(src)


def test_aggregation_full_size_linearity_and_mask_reproducibility(reddit_shape):
    """16 384 rows x 32 slots over the full [232 965, 602] table: out(aX + bY) = a out(X) + b out(Y) under one mask,
    and the same (seed, offset) gives the same bits."""
    import torch
    from grandplus_b200 import model as gm, synth
    _, _, graph, _ = reddit_shape
    coef = og.coef_for("ppr", ORDER, ALPHA)
    src = synth.sources(N, 16_384, seed=9, device="cuda")
    _, col, _, val32 = graph.gfpush_device(src, coef, RMAX, K, want_fp32=True)
    X = synth.features(N, F, seed=1, device="cuda")
    Y = synth.features(N, F, seed=2, device="cuda")
    fx, fy, fz = gm.DeviceFeatures(X), gm.DeviceFeatures(Y), gm.DeviceFeatures(2.0 * X - 3.0 * Y)
    kw = dict(dropnode_rate=0.5, training=True, n_aug=2, seed=11, offset=5)
    ox = gm.aggregate_slots(fx, col, val32, None, **kw)
    oy = gm.aggregate_slots(fy, col, val32, None, **kw)
    oz = gm.aggregate_slots(fz, col, val32, None, **kw)
    assert ox.shape == (2, 16_384, F)
    err = (oz - (2.0 * ox - 3.0 * oy)).abs().max()
    assert float(err) <= 1e-5 * float((2.0 * ox - 3.0 * oy).abs().max())
    assert torch.equal(ox, gm.aggregate_slots(fx, col, val32, None, **kw))
    assert not torch.equal(ox, gm.aggregate_slots(fx, col, val32, None, **{**kw, "offset": 6}))
