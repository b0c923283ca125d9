"""GPU parity tests for the MAG-Scholar-C side of the path: MLP.emb with the
element-wise input dropout fused into the kernel, the row-sparse embedding gradient, and the lazy Adam step that must
reproduce torch.optim.Adam over the dense table (/root/reference/model_mag.py:27,48-55,312-313,367-369)."""
import numpy as np
import pytest

from oracle import aggregate as oa

pytestmark = pytest.mark.gpu


def _batch(rng, n_attr, B, kmax):
    counts = rng.integers(1, kmax + 1, size=B)
    node_idx = np.repeat(np.arange(B), counts).astype(np.int64)
    nza = len(node_idx)
    attr_idx = rng.integers(0, n_attr, size=nza).astype(np.int64)
    attr_data = (rng.random(nza) + 0.1).astype(np.float32)
    return node_idx, attr_idx, attr_data


@pytest.mark.parametrize("H", [64, 7, 130])
@pytest.mark.parametrize("p", [0.3, 0.0])
def test_emb_input_dropout_forward_backward_match_oracle(H, p):
    import torch
    from grandplus_b200 import model as gm
    rng = np.random.default_rng(H * 7 + int(p * 10))
    n_attr, B = 900, 61
    node_idx, attr_idx, attr_data = _batch(rng, n_attr, B, 40)
    nza = len(node_idx)
    table = rng.standard_normal((n_attr, H)).astype(np.float32)
    gout = rng.standard_normal((B, H)).astype(np.float32)
    outs = {}
    for sparse in (False, True):
        w = torch.from_numpy(table).cuda().requires_grad_(True)
        out = gm.emb(w, torch.from_numpy(attr_idx), torch.from_numpy(node_idx).cuda(), torch.from_numpy(attr_data).cuda(),
                     input_droprate=p, training=True, sparse_grad=sparse, seed=5, offset=9)
        out.backward(torch.from_numpy(gout).cuda())
        g = w.grad
        if sparse:
            assert g.is_sparse
            g = g.coalesce()                # (autograd's accumulation drops the flag; the indices are distinct already)
            assert g._nnz() == len(np.unique(attr_idx))
            rows = g._indices()[0].cpu().numpy()
            np.testing.assert_array_equal(rows, np.unique(attr_idx))          # exactly the touched rows, ascending
            g = g.to_dense()
        outs[sparse] = (out.detach().cpu().numpy(), g.cpu().numpy())
    mask = gm.emb_dropout_mask(nza, H, p, 5, 9, "cuda").cpu().numpy() if p > 0 else None
    if p > 0:
        assert abs(mask.mean() - (1 - p)) < 0.02
        assert not np.array_equal(mask, gm.emb_dropout_mask(nza, H, p, 5, 10, "cuda").cpu().numpy())
    want = oa.emb(table, attr_idx, node_idx, attr_data, dtype=np.float64, elem_mask=mask, input_droprate=p)
    E = np.abs(table[attr_idx]).astype(np.float64) * (1.0 if mask is None else mask / (1 - p))
    scale = oa._segment_sum(E * attr_data[:, None].astype(np.float64), node_idx, B, np.float64) / \
        (oa._segment_sum(attr_data[:, None].astype(np.float64), node_idx, B, np.float64) + 1e-10)
    want_g = oa.emb_backward_table(gout, n_attr, attr_idx, node_idx, attr_data, elem_mask=mask, input_droprate=p)
    gscale = oa.emb_backward_table(np.abs(gout), n_attr, attr_idx, node_idx, attr_data, elem_mask=mask, input_droprate=p)
    for sparse in (False, True):
        out, g = outs[sparse]
        assert np.all(np.abs(out - want) <= 1e-5 * np.maximum(scale, 1e-30) + 1e-30)     # north_star: 1e-5 relative in fp32
        assert np.all(np.abs(g - want_g) <= 2e-5 * gscale + 1e-30)
    np.testing.assert_array_equal(outs[False][0], outs[True][0])
    # eval mode ignores the dropout rate (model_mag.py:50: training=self.training)
    w = torch.from_numpy(table).cuda()
    ev = gm.emb(w, torch.from_numpy(attr_idx), torch.from_numpy(node_idx).cuda(), torch.from_numpy(attr_data).cuda(),
                input_droprate=p, training=False)
    want_ev = oa.emb(table, attr_idx, node_idx, attr_data, dtype=np.float64)
    assert np.all(np.abs(ev.cpu().numpy() - want_ev) <= 1e-5 * np.maximum(
        oa._segment_sum(np.abs(table[attr_idx]).astype(np.float64) * attr_data[:, None], node_idx, B, np.float64) /
        (oa._segment_sum(attr_data[:, None].astype(np.float64), node_idx, B, np.float64) + 1e-10), 1e-30))


def test_sparse_row_adam_reproduces_dense_adam():
    """30 steps, each touching a random handful of rows: the rows a step reads (after prepare) and the whole table (after
    flush) must equal torch.optim.Adam run over the dense table with zero gradients elsewhere -- the reference's optimizer
    (model_mag.py:312-313) -- including rows that were touched once and then drift for 20 steps."""
    import torch
    from grandplus_b200.optim import SparseRowAdam
    torch.manual_seed(3)
    n, H, lr = 400, 24, 0.01                                   # scripts/run_mag.sh:7: lr 0.01, weight_decay 0
    w0 = torch.randn(n, H, device="cuda")
    dense = w0.clone().requires_grad_(True)
    ref = torch.optim.Adam([dense], lr=lr, weight_decay=0.0)
    ours = w0.clone()
    opt = SparseRowAdam(ours, lr=lr)
    rng = np.random.default_rng(8)
    for step in range(30):
        rows = np.unique(rng.integers(0, n, size=int(rng.integers(1, 25))))
        if step in (7, 8):
            rows = np.array([5], dtype=np.int64)               # a row hit on consecutive steps
        rows_t = torch.from_numpy(rows).cuda()
        g = torch.randn(len(rows), H, device="cuda")
        opt.prepare(rows_t)
        np.testing.assert_allclose(ours[rows_t].cpu().numpy(), dense.detach()[rows_t].cpu().numpy(), rtol=3e-6, atol=3e-7)
        dense.grad = torch.zeros_like(dense)
        dense.grad[rows_t] = g
        ref.step()
        opt.step(rows_t, g)
    assert opt.step_count == 30
    stale = (ours - dense.detach()).abs().max().item()
    assert stale > 1e-5                                        # untouched rows really lag until they are caught up ...
    opt.flush()
    np.testing.assert_allclose(ours.cpu().numpy(), dense.detach().cpu().numpy(), rtol=3e-6, atol=3e-7)   # ... and then agree
    never = np.setdiff1d(np.arange(n), np.unique(np.concatenate([[5]])))   # rows with zero moments did not move at all
    untouched = (opt.exp_avg.abs().sum(1) == 0).cpu().numpy()
    np.testing.assert_array_equal(ours.cpu().numpy()[untouched], w0.cpu().numpy()[untouched])


def test_mag_step_with_sparse_gradient_matches_dense_step():
    """One model_mag.py training step of the path (emb -> random_prop -> loss -> backward -> Adam) with the sparse
    gradient + SparseRowAdam against the same step with a dense gradient + torch.optim.Adam."""
    import torch
    from grandplus_b200 import model as gm
    from grandplus_b200.optim import SparseRowAdam
    rng = np.random.default_rng(21)
    n_attr, H, n_nbr, B, K = 5000, 64, 300, 40, 8
    node_idx, attr_idx, attr_data = _batch(rng, n_attr, n_nbr, 30)
    src = np.repeat(np.arange(B), K).astype(np.int64)
    nbr_of = rng.integers(0, n_nbr, size=B * K)
    scores = (rng.random(B * K) + 0.05).astype(np.float32)
    W0 = (rng.standard_normal((n_attr, H)) * 0.1).astype(np.float32)
    target = torch.from_numpy(rng.standard_normal((B, H)).astype(np.float32)).cuda()
    res = {}
    for sparse in (False, True):
        w = torch.from_numpy(W0).cuda().requires_grad_(True)
        opt = SparseRowAdam(w, lr=0.01) if sparse else torch.optim.Adam([w], lr=0.01)
        for it in range(3):
            if sparse:
                opt.prepare(torch.from_numpy(attr_idx).cuda())
            batch_emb = gm.emb(w, torch.from_numpy(attr_idx), torch.from_numpy(node_idx).cuda(),
                               torch.from_numpy(attr_data).cuda(), sparse_grad=sparse)
            feats = batch_emb[torch.from_numpy(nbr_of).cuda()]                # the batch's neighbour rows (model_mag.py:341)
            out = gm.random_prop(feats, torch.from_numpy(scores).cuda(), torch.from_numpy(src).cuda(), 0.5, training=True,
                                 seed=3, offset=it + 1)
            loss = ((out - target) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
        if sparse:
            opt.flush()
        res[sparse] = w.detach().cpu().numpy()
    np.testing.assert_allclose(res[True], res[False], rtol=2e-5, atol=2e-6)
