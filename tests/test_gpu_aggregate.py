"""GPU parity tests, part 2: fused Pi.X aggregation (random_prop / random_prop_fused / emb /
PiMatrix) through the C ABI against the reference's golden vectors and the numpy oracle."""
import os

import numpy as np
import pytest

from oracle import aggregate as oa
from oracle import gfpush as og
from tests.helpers import GOLDEN, load_graph

pytestmark = pytest.mark.gpu

# aggregated features: north_star asks 1e-5 relative in fp32 given identical masks.  Entries are
# signed, so the bound is relative to the row's L1 mass  sum_j |m_j x_j| / sum_j m_j.
RTOL = 1e-5


def _cases(npz):
    z = np.load(os.path.join(GOLDEN, npz))
    return z, sorted({k.split("/")[0] for k in z.files})


def _scale(feats, m, idx):
    num = oa._segment_sum(np.abs(feats.astype(np.float64)) * m.astype(np.float64)[:, None], idx, int(idx[-1]) + 1, np.float64)
    den = oa._segment_sum(m.astype(np.float64)[:, None], idx, int(idx[-1]) + 1, np.float64) + 1e-12
    return num / den


def _assert_close(got, want64, scale):
    err = np.abs(got.astype(np.float64) - want64)
    tol = RTOL * np.maximum(scale, 1e-30)
    assert np.all(err <= tol + 1e-30), f"max err/tol = {np.max(err / (tol + 1e-30)):.3f}"


def test_random_prop_matches_reference_golden():
    import torch
    from grandplus_b200 import model as gm
    z, tags = _cases("random_prop.npz")
    for tag in tags:
        feats, scores, idx = z[f"{tag}/feats"], z[f"{tag}/scores"], z[f"{tag}/idx"]
        p, training, mask = float(z[f"{tag}/p"]), bool(z[f"{tag}/training"]), z[f"{tag}/mask"]
        out = gm.random_prop(torch.from_numpy(feats).cuda(), torch.from_numpy(scores).cuda(),
                             torch.from_numpy(idx).cuda(), p, training=training,
                             mask=torch.from_numpy(mask).cuda())
        got = out.cpu().numpy()
        assert got.shape == z[f"{tag}/out"].shape
        m = oa.dropout_scores(scores, p, training, mask)
        want64 = oa.random_prop(feats, scores, idx, p, training, mask, dtype=np.float64)
        sc = _scale(feats, m, idx)
        _assert_close(got, want64, sc)
        _assert_close(z[f"{tag}/out"], want64, sc)          # the reference's own fp32 result obeys the same bound
        dropped_rows = oa._segment_sum(m[:, None], idx, int(idx[-1]) + 1, np.float64)[:, 0] == 0
        assert np.all(got[dropped_rows] == 0)                # an all-dropped row is exactly 0


def test_emb_matches_reference_golden():
    import torch
    from grandplus_b200 import model as gm
    z, tags = _cases("emb.npz")
    for tag in tags:
        table = torch.from_numpy(z[f"{tag}/table"]).cuda()
        out = gm.emb(table, torch.from_numpy(z[f"{tag}/attr_idx"]), torch.from_numpy(z[f"{tag}/node_idx"]).cuda(),
                     torch.from_numpy(z[f"{tag}/attr_data"]).cuda())
        want64 = oa.emb(z[f"{tag}/table"], z[f"{tag}/attr_idx"], z[f"{tag}/node_idx"], z[f"{tag}/attr_data"], dtype=np.float64)
        sc = _scale(z[f"{tag}/table"][z[f"{tag}/attr_idx"]], z[f"{tag}/attr_data"], z[f"{tag}/node_idx"])
        _assert_close(out.cpu().numpy(), want64, sc)


# kernel variants reachable through gp_set_tuning: the default (register-staged, 64-bit loads), 128-bit loads, and the
# TMA-staged cp.async.bulk kernel with two buffer depths -- every one must give the oracle's rows
AGG_VARIANTS = {"default": {}, "vec4": {"agg_max_vec": 4}, "vec1": {"agg_max_vec": 1}, "bulk": {"agg_kernel": 2},
                "bulk_nbuf2": {"agg_kernel": 2, "agg_nbuf": 2}, "chunk1": {"agg_max_chunk": 1},
                "one_warp_per_item": {"agg_waves": 0}, "waves3": {"agg_waves": 3}}
AGG_DEFAULTS = {"agg_kernel": 0, "agg_nbuf": 0, "agg_max_vec": 2, "agg_max_chunk": 4, "agg_smem_kb": 96, "agg_waves": 1}


@pytest.fixture(params=sorted(AGG_VARIANTS))
def agg_variant(request):
    from grandplus_b200 import _lib
    for k, v in AGG_VARIANTS[request.param].items():
        _lib.set_tuning(k, v)
    yield request.param
    for k, v in AGG_DEFAULTS.items():
        _lib.set_tuning(k, v)


@pytest.mark.parametrize("Fdim", [1, 7, 64, 100, 500, 602, 1433, 2050])
@pytest.mark.parametrize("n_aug", [1, 2])
def test_fused_gather_matches_oracle(Fdim, n_aug, agg_variant):
    """random_prop_fused == oracle(random_prop(features[nbr], ...)) for every vector width / tiling / kernel variant."""
    import torch
    from grandplus_b200 import model as gm
    rng = np.random.default_rng(Fdim * 10 + n_aug)
    N, B, kmax = 3000, 97, 32
    X = rng.standard_normal((N, Fdim)).astype(np.float32)
    counts = rng.integers(0, kmax + 1, size=B); counts[-1] = max(counts[-1], 1); counts[3] = 0
    idx = np.repeat(np.arange(B), counts).astype(np.int64)
    nz = len(idx)
    nbr = rng.integers(0, N, size=nz).astype(np.int64)
    scores = (rng.random(nz) ** 4 + 1e-6).astype(np.float32)
    feats = gm.DeviceFeatures(X)
    out, mask = gm.random_prop_fused(feats, torch.from_numpy(nbr).cuda(), torch.from_numpy(scores).cuda(),
                                     torch.from_numpy(idx).cuda(), 0.5, training=True, n_aug=n_aug,
                                     seed=11, offset=5, return_mask=True)
    out, mask = out.cpu().numpy(), mask.cpu().numpy()
    assert out.shape == (n_aug, B, Fdim) and mask.shape == (n_aug, nz)
    assert 0.35 < mask.mean() < 0.65
    if n_aug == 2:
        assert (mask[0] != mask[1]).mean() > 0.3            # independent masks per augmentation
    regen = gm.dropnode_mask(nz, n_aug, 0.5, 11, 5, "cuda").cpu().numpy()
    np.testing.assert_array_equal(mask, regen)               # counter-based: reproducible from (seed, offset)
    for a in range(n_aug):
        m = oa.dropout_scores(scores, 0.5, True, mask[a])
        want64 = oa.random_prop(X[nbr], scores, idx, 0.5, True, mask[a], dtype=np.float64)
        _assert_close(out[a], want64, _scale(X[nbr], m, idx))
    assert np.all(out[:, 3] == 0)                            # empty row


@pytest.mark.parametrize("n_aug", [5, 6, 9])
def test_more_than_four_augmentations_run_in_groups(n_aug):
    """The reference's --sample is unbounded (model.py:321): above four augmentations the launches are grouped, every
    augmentation gets an independent mask, the masks are reproducible, and every augmentation matches the oracle --
    through random_prop_fused, random_prop (with autograd) and aggregate_slots."""
    import torch
    from grandplus_b200 import model as gm
    rng = np.random.default_rng(n_aug)
    N, B, Fdim, K = 2000, 64, 37, 16
    X = rng.standard_normal((N, Fdim)).astype(np.float32)
    idx = np.repeat(np.arange(B), K).astype(np.int64)
    nz = len(idx)
    nbr = rng.integers(0, N, size=nz).astype(np.int64)
    scores = (rng.random(nz) + 1e-3).astype(np.float32)
    feats = gm.DeviceFeatures(X)
    out, mask = gm.random_prop_fused(feats, torch.from_numpy(nbr).cuda(), torch.from_numpy(scores).cuda(),
                                     torch.from_numpy(idx).cuda(), 0.5, training=True, n_aug=n_aug, seed=3, offset=9,
                                     return_mask=True)
    out, mask = out.cpu().numpy(), mask.cpu().numpy()
    assert out.shape == (n_aug, B, Fdim) and mask.shape == (n_aug, nz)
    np.testing.assert_array_equal(mask, gm.dropnode_mask(nz, n_aug, 0.5, 3, 9, "cuda").cpu().numpy())
    for a in range(n_aug):
        for b in range(a):
            assert (mask[a] != mask[b]).mean() > 0.3          # pairwise independent, across the groups too
        m = oa.dropout_scores(scores, 0.5, True, mask[a])
        want64 = oa.random_prop(X[nbr], scores, idx, 0.5, True, mask[a], dtype=np.float64)
        _assert_close(out[a], want64, _scale(X[nbr], m, idx))
    # the reference's own signature on pre-gathered rows, importing the mask; gradient reaches every augmentation
    f = torch.from_numpy(X[nbr]).cuda().requires_grad_(True)
    o2 = gm.random_prop(f, torch.from_numpy(scores).cuda(), torch.from_numpy(idx).cuda(), 0.5, training=True, n_aug=n_aug,
                        mask=torch.from_numpy(mask).cuda())
    np.testing.assert_allclose(o2.detach().cpu().numpy(), out, rtol=1e-5, atol=1e-6)
    w = torch.arange(1, n_aug + 1, device="cuda", dtype=torch.float32)[:, None, None]
    (o2 * w).sum().backward()
    g = f.grad.cpu().numpy()
    want_g = np.zeros_like(g)
    for a in range(n_aug):
        m = oa.dropout_scores(scores, 0.5, True, mask[a]).astype(np.float64)
        den = np.zeros(B); np.add.at(den, idx, m); den += 1e-12
        want_g += (a + 1) * (m / den[idx])[:, None]
    np.testing.assert_allclose(g, want_g, rtol=2e-4, atol=1e-6)
    # GFPush slot layout
    col = torch.from_numpy(nbr.reshape(B, K).astype(np.int32)).cuda()
    val = torch.from_numpy(scores.reshape(B, K)).cuda()
    o3 = gm.aggregate_slots(feats, col, val, None, 0.5, True, n_aug=n_aug, seed=3, offset=9)
    np.testing.assert_allclose(o3.cpu().numpy(), out, rtol=1e-5, atol=1e-6)


def test_eval_mode_ignores_mask_and_matches_oracle():
    import torch
    from grandplus_b200 import model as gm
    rng = np.random.default_rng(0)
    N, B, Fdim = 500, 40, 100
    X = rng.standard_normal((N, Fdim)).astype(np.float32)
    idx = np.repeat(np.arange(B), 16).astype(np.int64)
    nbr = rng.integers(0, N, size=len(idx))
    scores = rng.random(len(idx)).astype(np.float32)
    out = gm.random_prop_fused(gm.DeviceFeatures(X), torch.from_numpy(nbr).cuda(), torch.from_numpy(scores).cuda(),
                               torch.from_numpy(idx).cuda(), 0.5, training=False)
    want64 = oa.random_prop(X[nbr], scores, idx, 0.5, False, None, dtype=np.float64)
    _assert_close(out.cpu().numpy(), want64, _scale(X[nbr], scores, idx))


def test_linearity_and_scale_invariance():
    """Size-independent properties: linear in X; invariant to a common scaling of a row's scores."""
    import torch
    from grandplus_b200 import model as gm
    rng = np.random.default_rng(1)
    N, B, Fdim, K = 20000, 4096, 602, 32
    X1 = rng.standard_normal((N, Fdim)).astype(np.float32)
    X2 = rng.standard_normal((N, Fdim)).astype(np.float32)
    idx = torch.from_numpy(np.repeat(np.arange(B), K)).cuda()
    nbr = torch.from_numpy(rng.integers(0, N, size=B * K)).cuda()
    s = torch.from_numpy(rng.random(B * K).astype(np.float32) + 0.01).cuda()
    kw = dict(dropnode_rate=0.5, training=True, seed=3, offset=9)
    a = gm.random_prop_fused(gm.DeviceFeatures(X1), nbr, s, idx, **kw)
    b = gm.random_prop_fused(gm.DeviceFeatures(X2), nbr, s, idx, **kw)
    ab = gm.random_prop_fused(gm.DeviceFeatures(2 * X1 + X2), nbr, s, idx, **kw)
    assert torch.allclose(ab, 2 * a + b, rtol=1e-4, atol=1e-4)
    a4 = gm.random_prop_fused(gm.DeviceFeatures(X1), nbr, 4 * s, idx, **kw)
    assert torch.allclose(a4, a, rtol=1e-5, atol=1e-6)
    ones = gm.random_prop_fused(gm.DeviceFeatures(np.ones((N, 8), np.float32)), nbr, s, idx, **kw)
    kept = ones.abs().sum(1) > 0
    assert torch.allclose(ones[kept], torch.ones_like(ones[kept]), rtol=1e-5)   # rows are convex combinations


def test_pimatrix_matches_reference_batch_path():
    """GFPush on the device -> PiMatrix -> aggregate == the reference's coo->csr, slice, nonzero,
    gather, random_prop chain (model.py:270-272, 310-322) evaluated by the oracle."""
    import torch
    from grandplus_b200 import model as gm
    from grandplus_b200.precompute import propagation
    indptr, indices = load_graph("cora")
    n = indptr.shape[0] - 1
    coef = og.coef_for("ppr", 20, 0.2)
    rng = np.random.default_rng(2)
    src = np.sort(rng.choice(n, 600, replace=False)).astype(np.int32)
    src[0] = 0                                              # node 0 a source: keeps the pad hazard of SURVEY 8b away
    src = np.unique(src).astype(np.int32)
    X = rng.standard_normal((n, 1433)).astype(np.float32)
    g = propagation.Graph(indptr, indices, 0)
    pi = gm.PiMatrix.from_graph(g, src, coef, 1e-7, 32)
    feats = gm.DeviceFeatures(X)
    # reference chain on the host from the same Pi
    K = 32
    col = pi.col.cpu().numpy().ravel(); val = pi.val.double().cpu().numpy().ravel()
    row = np.where(val > 0, np.repeat(src, K), 0)
    adj = oa.topk_adj_from_slots(row, col, val, n)
    batch = rng.choice(src, 150, replace=False)
    sidx, nidx, sc = oa.batch_slice(adj, batch)
    for training in (False, True):
        out, mask = pi.aggregate(feats, batch, 0.5, training=training, seed=21, offset=1, return_mask=True)
        out = out.cpu().numpy()
        # map the slot-ordered mask onto the reference's column-sorted entries
        mask = mask.cpu().numpy()[0].reshape(-1, K)
        rows = pi.slot_rows(batch).cpu().numpy()
        m_ref = np.zeros(len(sidx), np.uint8)
        colk = pi.col.cpu().numpy()
        for e, (b_i, c) in enumerate(zip(sidx, nidx)):
            r = rows[b_i]
            hit = np.nonzero((colk[r] == c) & (pi.val.cpu().numpy()[r] > 0))[0]
            assert len(hit) == 1
            m_ref[e] = mask[r, hit[0]]
        want64 = oa.random_prop(X[nidx], sc, sidx, 0.5, training, m_ref, dtype=np.float64)
        m = oa.dropout_scores(sc, 0.5, training, m_ref)
        _assert_close(out, want64, _scale(X[nidx], m, sidx))


def test_backward_matches_oracle():
    """model_mag.py:356 keeps autograd through random_prop and emb."""
    import torch
    from grandplus_b200 import model as gm
    rng = np.random.default_rng(4)
    n_attr, H, n_nodes, B = 400, 64, 120, 30
    table = torch.from_numpy(rng.standard_normal((n_attr, H)).astype(np.float32)).cuda().requires_grad_(True)
    cnt = rng.integers(1, 9, size=n_nodes)
    node_idx = np.repeat(np.arange(n_nodes), cnt).astype(np.int64)
    attr_idx = rng.integers(0, n_attr, size=len(node_idx)).astype(np.int64)
    attr_data = (rng.random(len(node_idx)) + 0.1).astype(np.float32)
    src_cnt = np.full(B, n_nodes // B)
    mat_idx = np.repeat(np.arange(B), src_cnt).astype(np.int64)
    scores = (rng.random(n_nodes) + 0.05).astype(np.float32)
    node_emb = gm.emb(table, torch.from_numpy(attr_idx), torch.from_numpy(node_idx).cuda(), torch.from_numpy(attr_data).cuda())
    out, mask = gm.random_prop(node_emb, torch.from_numpy(scores).cuda(), torch.from_numpy(mat_idx).cuda(), 0.5,
                               training=True, seed=8, offset=2, return_mask=True)
    gout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(torch.from_numpy(gout).cuda())
    mask = mask.cpu().numpy()[0]
    g_nodes = oa.random_prop_backward_feats(gout, scores, mat_idx, 0.5, True, mask)
    g_table = oa.emb_backward_table(g_nodes, n_attr, attr_idx, node_idx, attr_data)
    got = table.grad.cpu().numpy().astype(np.float64)
    assert np.max(np.abs(got - g_table)) <= 2e-5 * np.max(np.abs(g_table))
    # and the forward value feeding it
    want = oa.random_prop(oa.emb(table.detach().cpu().numpy(), attr_idx, node_idx, attr_data), scores, mat_idx, 0.5, True, mask,
                          dtype=np.float64)
    assert np.max(np.abs(out.detach().cpu().numpy() - want)) <= 2e-5 * np.max(np.abs(want))


def test_rejects_cpu_tensors_and_unsorted_index():
    import torch
    from grandplus_b200 import model as gm
    with pytest.raises(RuntimeError):
        gm.random_prop(torch.zeros(4, 3), torch.ones(4), torch.tensor([0, 0, 1, 1]), 0.5)
    with pytest.raises(ValueError):
        gm.random_prop(torch.zeros(4, 3).cuda(), torch.ones(4).cuda(), torch.tensor([0, 1, 0, 1]).cuda(), 0.5)
    with pytest.raises(IndexError):
        gm.random_prop_fused(gm.DeviceFeatures(np.zeros((5, 4), np.float32)), torch.tensor([0, 7]).cuda(),
                             torch.ones(2).cuda(), torch.tensor([0, 0]).cuda(), 0.5)
    with pytest.raises(IndexError):   # ids that are already int32 are range-checked too
        gm.random_prop_fused(gm.DeviceFeatures(np.zeros((5, 4), np.float32)), torch.tensor([0, 7], dtype=torch.int32).cuda(),
                             torch.ones(2).cuda(), torch.tensor([0, 0]).cuda(), 0.5)
    with pytest.raises(IndexError):
        gm.random_prop_fused(gm.DeviceFeatures(np.zeros((5, 4), np.float32)), torch.tensor([-1, 2], dtype=torch.int32).cuda(),
                             torch.ones(2).cuda(), torch.tensor([0, 0]).cuda(), 0.5)


def test_pimatrix_disk_cache_roundtrip(tmp_path):
    """Pi saved and reloaded aggregates to the same rows; a different key is a cache miss."""
    import torch
    from grandplus_b200 import model as gm
    from grandplus_b200.precompute import propagation
    from tests.helpers import load_graph
    from oracle import gfpush as og
    indptr, indices = load_graph("cora")
    n = indptr.shape[0] - 1
    coef = og.coef_for("ppr", 6, 0.2)
    src = np.arange(0, n, 5, dtype=np.int32)
    g = propagation.Graph(indptr, indices, 0)
    pi = gm.PiMatrix.from_graph(g, src, coef, 1e-6, 16)
    key = gm.PiMatrix.cache_key(indptr, indices, src, coef, 1e-6, 16)
    assert key != gm.PiMatrix.cache_key(indptr, indices, src, coef, 1e-5, 16)
    path = str(tmp_path / "pi_cache")
    pi.save(path, key)
    assert gm.PiMatrix.load(path, key="something else") is None
    pi2 = gm.PiMatrix.load(path, key=key)
    X = gm.DeviceFeatures(np.random.default_rng(0).standard_normal((n, 33)).astype(np.float32))
    a = pi.aggregate(X, src[:100], 0.5, training=True, n_aug=2, seed=3, offset=4)
    b = pi2.aggregate(X, src[:100], 0.5, training=True, n_aug=2, seed=3, offset=4)
    assert torch.equal(a, b)
