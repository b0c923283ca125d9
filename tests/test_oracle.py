"""CPU tests: the oracle (oracle/) against the golden vectors the reference produced
(tests/golden/make_golden.py) and, where oracle/_ref exists, against the reference live."""
import os

import numpy as np
import pytest

from oracle import aggregate as oa
from oracle import gfpush as og
from tests.helpers import GOLDEN, check_topk_rows, load_graph

DATASETS = ["cora", "citeseer", "pubmed"]
MODES = ["ppr", "avg", "single"]


def _golden_rows(z):
    K = int(z["K"])
    return K, og.rows_as_sets(z["col_idx"], z["value"], K)


@pytest.mark.parametrize("name", DATASETS)
@pytest.mark.parametrize("mode", MODES)
def test_oracle_matches_reference_golden(name, mode):
    """The reference's own outputs (unmodified propagation.cpp) pin the C restatement."""
    indptr, indices = load_graph(name)
    z = np.load(os.path.join(GOLDEN, f"gfpush_{name}_{mode}.npz"))
    K = int(z["K"])
    # golden (reference) rows obey the oracle's reserve vectors, modulo ties
    worst = check_topk_rows(indptr, indices, z["node_idx"], z["coef"], float(z["rmax"]), K,
                            z["col_idx"], z["value"])
    assert worst < 1e-12
    # and the oracle's own top-k equals the reference's wherever there is no tie
    row, col, val, st = og.gfpush(indptr, indices, z["node_idx"], z["coef"], float(z["rmax"]), K)
    same = 0
    for (gc, gv), (oc, ov) in zip(og.rows_as_sets(z["col_idx"], z["value"], K), og.rows_as_sets(col, val, K)):
        assert len(gc) == len(oc)
        if np.array_equal(gc, oc):
            same += 1
            np.testing.assert_allclose(ov, gv, rtol=1e-12, atol=0)
    # `single`/`avg` rows are full of exact ties (equal-mass neighbourhoods), which nth_element breaks arbitrarily
    assert same >= (0.5 if mode == "ppr" else 0.2) * len(z["node_idx"])
    assert st.edges_pushed > 0


@pytest.mark.parametrize("name", ["path8", "star33", "isolated", "dangling"])
def test_oracle_tiny_graphs(name):
    z = np.load(os.path.join(GOLDEN, f"tiny_{name}.npz"))
    tags = sorted({k.split("/")[0] for k in z.files if "/" in k})
    assert len(tags) == 5
    for tag in tags:
        K = int(z[f"{tag}/K"])
        check_topk_rows(z["indptr"], z["indices"], z["node_idx"], z[f"{tag}/coef"], float(z[f"{tag}/rmax"]),
                        K, z[f"{tag}/col_idx"], z[f"{tag}/value"], row=z[f"{tag}/row_idx"])
        row, col, val, _ = og.gfpush(z["indptr"], z["indices"], z["node_idx"], z[f"{tag}/coef"],
                                     float(z[f"{tag}/rmax"]), K)
        check_topk_rows(z["indptr"], z["indices"], z["node_idx"], z[f"{tag}/coef"], float(z[f"{tag}/rmax"]),
                        K, col, val, row=row)


def test_oracle_dangling_mass_returns_to_source():
    """graph.h:91-93: a degree-0 node hands its residue back to the source."""
    z = np.load(os.path.join(GOLDEN, "tiny_dangling.npz"))
    coef = og.coef_for("avg", 3)
    dense, seen, _ = og.reserve_row(z["indptr"], z["indices"], 2, coef, 0.0)
    # 2 -> 3 (dangling) -> back to 2 -> 3
    np.testing.assert_allclose(dense[[2, 3]], [0.5, 0.5])
    assert seen.sum() == 2


def test_oracle_mass_and_threshold_properties():
    indptr, indices = load_graph("cora")
    coef = og.coef_for("ppr", 20, 0.2)
    total_exact, _, _ = og.reserve_row(indptr, indices, 5, coef, 0.0)
    assert abs(total_exact.sum() - 1.0) < 1e-12          # rmax = 0: nothing is dropped
    approx, _, st = og.reserve_row(indptr, indices, 5, coef, 1e-4)
    assert approx.sum() <= 1.0 + 1e-12
    assert np.all(approx <= total_exact + 1e-15)           # thresholding only loses mass
    assert st.edges_pushed <= (len(coef) - 1) / 1e-4       # pushed edges/level <= 1/rmax


@pytest.mark.skipif(not og.reference_available(), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("mode,order,alpha,rmax,K", [("ppr", 6, 0.5, 1e-5, 16), ("avg", 4, 0.2, 1e-5, 16),
                                                     ("single", 2, 0.2, 1e-5, 16), ("ppr", 8, 0.1, 0.0, 8)])
def test_oracle_vs_live_reference(mode, order, alpha, rmax, K):
    indptr, indices = load_graph("pubmed")
    src = np.random.default_rng(3).choice(indptr.shape[0] - 1, 64, replace=False).astype(np.int32)
    coef = og.coef_for(mode, order, alpha)
    row, col, val = og.reference_gfpush(indptr, indices, src, coef, rmax, K)
    assert check_topk_rows(indptr, indices, src, coef, rmax, K, col, val, row=row) < 1e-12


def test_reference_zero_pad_convention():
    """graph.h:117-126 leaves (0,0,0.0) in unfilled slots; the oracle must too."""
    z = np.load(os.path.join(GOLDEN, "tiny_isolated.npz"))
    tag = "ppr_o6_r0_k64"
    row, col, val, _ = og.gfpush(z["indptr"], z["indices"], z["node_idx"], z[f"{tag}/coef"], 0.0, 64)
    assert (val > 0).sum() == (z[f"{tag}/value"] > 0).sum()
    assert np.all(row[val == 0] == 0) and np.all(col[val == 0] == 0)


# ------------------------------------------------------------------ part 2: aggregation oracle
def _cases(npz):
    z = np.load(os.path.join(GOLDEN, npz))
    return z, sorted({k.split("/")[0] for k in z.files})


def test_random_prop_oracle_matches_reference_golden():
    z, tags = _cases("random_prop.npz")
    assert len(tags) == 6
    for tag in tags:
        got = oa.random_prop(z[f"{tag}/feats"], z[f"{tag}/scores"], z[f"{tag}/idx"], float(z[f"{tag}/p"]),
                             bool(z[f"{tag}/training"]), z[f"{tag}/mask"])
        # same fp32 op order as the reference's scatter_add_ on CPU -> bit-exact
        np.testing.assert_array_equal(got, z[f"{tag}/out"])
        got64 = oa.random_prop(z[f"{tag}/feats"], z[f"{tag}/scores"], z[f"{tag}/idx"], float(z[f"{tag}/p"]),
                               bool(z[f"{tag}/training"]), z[f"{tag}/mask"], dtype=np.float64)
        np.testing.assert_allclose(got64, z[f"{tag}/out"], rtol=2e-5, atol=2e-6)


def test_random_prop_all_dropped_row_is_zero():
    feats = np.ones((4, 3), np.float32)
    out = oa.random_prop(feats, np.full(4, 0.25, np.float32), np.array([0, 0, 1, 1]), 0.5, True,
                         np.array([0, 0, 1, 0], np.uint8))
    assert np.all(out[0] == 0) and np.allclose(out[1], 1.0)


def test_emb_oracle_matches_reference_golden():
    z, tags = _cases("emb.npz")
    for tag in tags:
        got = oa.emb(z[f"{tag}/table"], z[f"{tag}/attr_idx"], z[f"{tag}/node_idx"], z[f"{tag}/attr_data"])
        np.testing.assert_array_equal(got, z[f"{tag}/out"])


def test_batch_slice_matches_scipy_contract():
    """model.py:270-272,310-313: pads collapse into (0,0); rows come back column-sorted."""
    z = np.load(os.path.join(GOLDEN, "gfpush_cora_ppr.npz"))
    indptr, _ = load_graph("cora")
    n = indptr.shape[0] - 1
    adj = oa.topk_adj_from_slots(z["row_idx"], z["col_idx"], z["value"], n)
    batch = z["node_idx"][:10]
    s, nb, sc = oa.batch_slice(adj, batch)
    assert s[0] == 0 and s[-1] == 9 and np.all(np.diff(s) >= 0)
    assert sc.dtype == np.float32 and len(sc) == len(nb)


@pytest.mark.parametrize("name", ["cora", "citeseer"])
def test_predict_oracle_matches_reference_golden(name):
    """oracle.predict.propagate_exact against what the reference's own predict() hands to its MLP."""
    import scipy.sparse as sp
    from oracle import predict as op
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "predict.npz"))
    g = np.load(os.path.join(GOLDEN, f"graph_{name}.npz"))
    n = len(g["indptr"]) - 1
    adj = sp.csr_matrix((np.ones(len(g["indices"])), g["indices"], g["indptr"]), shape=(n, n))
    X = z[f"{name}/X"]
    tags = sorted({k.split("/")[1] for k in z.files if k.startswith(name + "/") and k.count("/") == 2})
    assert len(tags) == 3
    for tag in tags:
        mode = tag.split("_")[0]
        got = op.propagate_exact(adj, X.copy(), int(z[f"{name}/{tag}/order"]), float(z[f"{name}/{tag}/alpha"]), mode)
        np.testing.assert_array_equal(got.astype(np.float32), z[f"{name}/{tag}/feat"])   # same arithmetic: bit-exact
