"""GPU parity tests, part 1: GFPush + top-k through the C ABI (Graph.gfpush_omp -> gp_gfpush)
against the reference's golden vectors and the oracle.  Run with -m gpu on the B200 box."""
import os

import numpy as np
import pytest

from oracle import gfpush as og
from tests.helpers import GOLDEN, check_topk_rows, load_graph

pytestmark = pytest.mark.gpu

DATASETS = ["cora", "citeseer", "pubmed"]
MODES = ["ppr", "avg", "single"]
SMEM, HBM = 1, 2


def _graph(indptr, indices, **cfg):
    from grandplus_b200.precompute import propagation
    g = propagation.Graph(np.array(indptr, dtype=np.int32), np.array(indices, dtype=np.int32), 0)
    if cfg:
        g.configure(**cfg)
    return g


def _run(g, node_idx, coef, rmax, K):
    S = len(node_idx)
    row = np.zeros(S * K, dtype=np.int32)
    col = np.zeros(S * K, dtype=np.int32)
    val = np.zeros(S * K, dtype=np.float64)
    g.gfpush_omp(node_idx, row, col, val, coef, rmax, K)  # exactly model.py:268
    return row, col, val


@pytest.mark.parametrize("name", DATASETS)
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("scratch", [SMEM, HBM])
def test_gfpush_matches_reference_golden(name, mode, scratch):
    indptr, indices = load_graph(name)
    z = np.load(os.path.join(GOLDEN, f"gfpush_{name}_{mode}.npz"))
    K, rmax = int(z["K"]), float(z["rmax"])
    g = _graph(indptr, indices, scratch_mode=scratch)
    row, col, val = _run(g, z["node_idx"].astype(np.int64), z["coef"], rmax, K)  # int64 like model.py:247
    assert g.last_stats()["scratch_mode"] == scratch
    worst = check_topk_rows(indptr, indices, z["node_idx"], z["coef"], rmax, K, col, val, row=row)
    assert worst < 1e-11, worst
    # against the reference's own rows: identical sets wherever the reference row has no tie at the cut
    same = 0
    for (gc, gv), (rc, rv) in zip(og.rows_as_sets(col, val, K), og.rows_as_sets(z["col_idx"], z["value"], K)):
        assert len(gc) == len(rc)
        if np.array_equal(gc, rc):
            same += 1
            np.testing.assert_allclose(gv, rv, rtol=1e-11, atol=0)   # north_star asks 1e-5
    assert same >= (0.5 if mode == "ppr" else 0.2) * len(z["node_idx"])


@pytest.mark.parametrize("name", ["path8", "star33", "isolated", "dangling"])
@pytest.mark.parametrize("scratch", [SMEM, HBM])
def test_gfpush_tiny_graphs(name, scratch):
    z = np.load(os.path.join(GOLDEN, f"tiny_{name}.npz"))
    g = _graph(z["indptr"], z["indices"], scratch_mode=scratch)
    for tag in sorted({k.split("/")[0] for k in z.files if "/" in k}):
        K, rmax, coef = int(z[f"{tag}/K"]), float(z[f"{tag}/rmax"]), z[f"{tag}/coef"]
        row, col, val = _run(g, z["node_idx"], coef, rmax, K)
        check_topk_rows(z["indptr"], z["indices"], z["node_idx"], coef, rmax, K, col, val, row=row)
        # same number of filled slots as the reference, row by row
        ref_filled = (z[f"{tag}/value"].reshape(-1, K) > 0).sum(1)
        np.testing.assert_array_equal((val.reshape(-1, K) > 0).sum(1), ref_filled)


@pytest.mark.parametrize("block", [256, 512, 1024])
@pytest.mark.parametrize("scratch", [SMEM, HBM])
def test_gfpush_all_sources_cora_block_sizes(block, scratch):
    """Every node of Cora as a source (S=N), ppr at scripts/run_cora.sh:7 parameters."""
    indptr, indices = load_graph("cora")
    n = indptr.shape[0] - 1
    coef = og.coef_for("ppr", 20, 0.2)
    g = _graph(indptr, indices, scratch_mode=scratch, block_threads=block)
    src = np.arange(n, dtype=np.int32)
    row, col, val = _run(g, src, coef, 1e-7, 32)
    st = g.last_stats()
    assert st["sources"] == n
    _, _, _, ost = og.gfpush(indptr, indices, src, coef, 1e-7, 32)
    # work counters are integers of the algorithm: exact unless a threshold comparison flipped
    assert abs(st["edges_pushed"] - ost.edges_pushed) <= 1e-6 * ost.edges_pushed
    assert abs(st["frontier_total"] - ost.frontier_total) <= 1e-6 * ost.frontier_total
    check_topk_rows(indptr, indices, src, coef, 1e-7, 32, col, val, row=row, max_rows=200)


def test_gfpush_idempotent_and_order_independent():
    indptr, indices = load_graph("pubmed")
    coef = og.coef_for("ppr", 6, 0.5)
    rng = np.random.default_rng(5)
    src = rng.choice(indptr.shape[0] - 1, 512, replace=False).astype(np.int32)
    g = _graph(indptr, indices)
    a = og.rows_as_sets(*_run(g, src, coef, 1e-5, 16)[1:], 16)
    b = og.rows_as_sets(*_run(g, src, coef, 1e-5, 16)[1:], 16)
    perm = rng.permutation(len(src))
    c = og.rows_as_sets(*_run(g, src[perm], coef, 1e-5, 16)[1:], 16)
    for i in range(len(src)):
        for other in (b[i], c[int(np.nonzero(perm == i)[0][0])]):
            if np.array_equal(a[i][0], other[0]):
                np.testing.assert_allclose(a[i][1], other[1], rtol=1e-12)
            else:  # only a tie at the cut may differ
                assert len(a[i][0]) == len(other[0])
                assert abs(a[i][1].min() - other[1].min()) <= 1e-12 * a[i][1].min()


def test_gfpush_mass_conservation_full_rows():
    """rmax = 0 and K >= support: nothing is dropped, every row sums to 1 (coef sums to 1)."""
    z = np.load(os.path.join(GOLDEN, "tiny_star33.npz"))
    g = _graph(z["indptr"], z["indices"])
    coef = og.coef_for("ppr", 8, 0.3)
    _, _, val = _run(g, z["node_idx"], coef, 0.0, 64)
    np.testing.assert_allclose(val.reshape(-1, 64).sum(1), 1.0, rtol=0, atol=1e-13)


@pytest.mark.parametrize("scratch", [HBM])
def test_gfpush_powerlaw_synthetic(scratch):
    """A Chung-Lu graph with hubs (max degree in the thousands): CTA-wide hub expansion, HBM slabs."""
    from grandplus_b200 import synth
    indptr, indices = synth.powerlaw_csr(50_000, 600_000, seed=3)
    indptr, indices = indptr.numpy(), indices.numpy()
    deg = np.diff(indptr)
    src = np.concatenate([[int(np.argmax(deg))], synth.sources(50_000, 255, seed=2).numpy()]).astype(np.int32)
    for mode, order, alpha, rmax, K in [("ppr", 6, 0.05, 1e-5, 32), ("avg", 4, 0.2, 1e-6, 64), ("single", 2, 0.2, 1e-7, 64)]:
        coef = og.coef_for(mode, order, alpha)
        g = _graph(indptr, indices, scratch_mode=scratch)
        row, col, val = _run(g, src, coef, rmax, K)
        worst = check_topk_rows(indptr, indices, src, coef, rmax, K, col, val, row=row, max_rows=64)
        assert worst < 1e-11
        _, _, _, ost = og.gfpush(indptr, indices, src, coef, rmax, K)
        st = g.last_stats()
        assert abs(st["edges_pushed"] - ost.edges_pushed) <= 1e-6 * max(ost.edges_pushed, 1)
        if st["bucket_count"]:   # the bucket kernel's candidate merge never materialises the support (a source that falls back to the full merge counts)
            assert st["support_total"] <= ost.support_total
        else:
            assert abs(st["support_total"] - ost.support_total) <= 1e-6 * ost.support_total


def test_gfpush_device_resident_outputs_match_host_path():
    import torch
    indptr, indices = load_graph("citeseer")
    coef = og.coef_for("ppr", 10, 0.4)
    src = np.arange(0, 3327, 13, dtype=np.int32)
    g = _graph(indptr, indices)
    row, col, val = _run(g, src, coef, 1e-7, 32)
    drow, dcol, dval, dval32 = g.gfpush_device(torch.from_numpy(src).cuda(), coef, 1e-7, 32)
    torch.cuda.synchronize()
    a = og.rows_as_sets(col, val, 32)
    b = og.rows_as_sets(dcol.cpu().numpy().ravel(), dval.cpu().numpy().ravel(), 32)
    for (ac, av), (bc, bv) in zip(a, b):
        if np.array_equal(ac, bc):
            np.testing.assert_allclose(av, bv, rtol=1e-12)
    np.testing.assert_array_equal(dval32.cpu().numpy(), dval.cpu().numpy().astype(np.float32))
    assert np.all(drow.cpu().numpy()[dval.cpu().numpy() > 0] == np.repeat(src, 32).reshape(-1, 32)[dval.cpu().numpy() > 0])


def test_gfpush_rejects_bad_arguments():
    from grandplus_b200._lib import GPError
    indptr, indices = load_graph("cora")
    g = _graph(indptr, indices)
    coef = og.coef_for("ppr", 4, 0.2)
    S, K = 4, 8
    row = np.zeros(S * K, np.int32); col = np.zeros(S * K, np.int32); val = np.zeros(S * K, np.float64)
    with pytest.raises(ValueError):
        g.gfpush_omp(np.arange(S), row, col, val.astype(np.float32), coef, 1e-5, K)   # wrong dtype
    with pytest.raises(ValueError):
        g.gfpush_omp(np.arange(S), row[:-1], col, val, coef, 1e-5, K)                   # too short
    with pytest.raises(ValueError):
        g.gfpush_omp(np.array([0, 1, 99999, 2]), row, col, val, coef, 1e-5, K)          # id out of range
    with pytest.raises(GPError):
        g.gfpush_omp(np.arange(S), np.zeros(S * 4096, np.int32), np.zeros(S * 4096, np.int32),
                     np.zeros(S * 4096, np.float64), coef, 1e-5, 4096)                  # K too large
    bad_ptr = indptr.copy(); bad_ptr[5] = bad_ptr[4] - 1
    with pytest.raises(GPError):
        _graph(bad_ptr, indices)
    # empty source list is a no-op, like the reference's empty omp loop
    g.gfpush_omp(np.zeros(0, np.int64), row[:0], col[:0], val[:0], coef, 1e-5, K)


@pytest.mark.skipif(not og.reference_available(), reason="oracle/_ref not shipped to this box")
def test_gfpush_vs_live_reference_all_pubmed_sources():
    """Full Pubmed (BASELINE config 2), every node a source, against the reference module itself."""
    indptr, indices = load_graph("pubmed")
    n = indptr.shape[0] - 1
    coef = og.coef_for("ppr", 6, 0.5)
    src = np.arange(n, dtype=np.int32)
    g = _graph(indptr, indices)
    row, col, val = _run(g, src, coef, 1e-5, 16)
    rrow, rcol, rval = og.reference_gfpush(indptr, indices, src, coef, 1e-5, 16)
    mine, ref = og.rows_as_sets(col, val, 16), og.rows_as_sets(rcol, rval, 16)
    same, worst = 0, 0.0
    for (gc, gv), (rc, rv) in zip(mine, ref):
        assert len(gc) == len(rc)
        if np.array_equal(gc, rc):
            same += 1
            worst = max(worst, float(np.max(np.abs(gv - rv) / rv)))
        else:
            assert abs(gv.min() - rv.min()) <= 1e-9 * rv.min()   # differs only by a tie at the cut
    assert worst < 1e-11
    assert same > 0.8 * n


class _tuning:
    """Scoped gp_set_tuning: restores the defaults on exit."""
    DEFAULTS = {"push_bucket": 1, "push_bucket_merge": 0, "push_bucket_nb": 0, "push_bucket_block": 0, "push_bucket_fill": 5, "push_cluster": 0, "push_cluster_probe": 128, "push_hub_deg": 0, "push_max_clusters": 0, "push_smem_hash": 1,
                "push_smem_probe": 2, "push_max_ctas": 0}

    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        from grandplus_b200 import _lib
        for k, v in self.kv.items():
            _lib.set_tuning(k, v)

    def __exit__(self, *exc):
        from grandplus_b200 import _lib
        for k in self.kv:
            _lib.set_tuning(k, self.DEFAULTS[k])


# Every tier of HBM-mode GFPush: plain slabs / the shared-memory table in front of the slabs (with a probe limit of 1
# and 2 most nodes spill to the slab, so both residencies and their mix are exercised) / the cluster kernel with one
# CTA per source and with clusters of 2..16 CTAs exchanging pushed edges through L2, with hub entries shared by the
# whole cluster, and with a probe limit so small that many sources are handed over to the slab kernel.
TIERS = {
    "slab": dict(push_cluster=0, push_smem_hash=0, push_bucket=0),
    "bucket": dict(push_cluster=0, push_smem_hash=0, push_bucket=1, push_bucket_merge=1),   # auto: takes the slab kernel's place
    "bucket_forced": dict(push_cluster=0, push_bucket=2, push_bucket_merge=1),
    "bucket_cand": dict(push_cluster=0, push_smem_hash=0, push_bucket=1, push_bucket_merge=0),   # merges only the top-k candidates
    "bucket_cand_forced": dict(push_cluster=0, push_bucket=2, push_bucket_merge=0),
    "bucket_cand_nb8": dict(push_cluster=0, push_bucket=2, push_bucket_merge=0, push_bucket_nb=8),
    # two / three sources per SM: 512-thread CTAs with 8 192-slot tables, 256-thread CTAs with 4 096-slot tables
    "bucket_b512": dict(push_cluster=0, push_bucket=2, push_bucket_merge=1, push_bucket_block=512),
    "bucket_cand_b512": dict(push_cluster=0, push_bucket=2, push_bucket_merge=0, push_bucket_block=512),
    "bucket_b256": dict(push_cluster=0, push_bucket=2, push_bucket_merge=1, push_bucket_block=256),
    "bucket_cand_b256": dict(push_cluster=0, push_bucket=2, push_bucket_merge=0, push_bucket_block=256),
    "bucket_cand_b1024": dict(push_cluster=0, push_bucket=2, push_bucket_merge=0, push_bucket_block=1024),
    "smem": dict(push_cluster=0, push_smem_hash=2),
    "smem_probe1": dict(push_cluster=0, push_smem_hash=2, push_smem_probe=1),
    "smem_probe2": dict(push_cluster=0, push_smem_hash=2, push_smem_probe=2),
    "cluster_auto": dict(push_cluster=1),
    "cluster_g1": dict(push_cluster=-1),
    "cluster_g2": dict(push_cluster=2),
    "cluster_g4": dict(push_cluster=4),
    "cluster_g8": dict(push_cluster=8),
    "cluster_g16": dict(push_cluster=16),
    "cluster_g4_hubs": dict(push_cluster=4, push_hub_deg=8),
    "cluster_g1_redo": dict(push_cluster=-1, push_cluster_probe=3),
    "cluster_g2_redo": dict(push_cluster=2, push_cluster_probe=2),
    "cluster_g4_few": dict(push_cluster=4, push_max_clusters=3),
}


@pytest.mark.parametrize("tier", sorted(TIERS))
def test_gfpush_tiers_give_the_oracle_rows_and_counters(tier):
    """Every tier, and the hand-over between tiers, must give the oracle's rows and the oracle's work counters."""
    from grandplus_b200 import synth
    kv = TIERS[tier]
    indptr, indices = synth.powerlaw_csr(60_000, 700_000, seed=5)
    indptr, indices = indptr.numpy(), indices.numpy()
    src = synth.sources(60_000, 400, seed=4).numpy()
    coef = og.coef_for("ppr", 6, 0.05)
    with _tuning(**kv):
        g = _graph(indptr, indices, scratch_mode=HBM)
        g.cumulative_stats(reset=True)
        row, col, val = _run(g, src, coef, 1e-5, 32)
        st = g.cumulative_stats()
        # the tables must be clean afterwards: a second run on the same handle gives the same rows
        row2, col2, val2 = _run(g, src, coef, 1e-5, 32)
        st2 = g.cumulative_stats()
    worst = check_topk_rows(indptr, indices, src, coef, 1e-5, 32, col, val, row=row, max_rows=120)
    assert worst < 1e-11
    _, _, _, ost = og.gfpush(indptr, indices, src, coef, 1e-5, 32)
    assert abs(st["edges_pushed"] - ost.edges_pushed) <= 1e-6 * ost.edges_pushed      # handed-over work is not double counted
    assert abs(st["frontier_total"] - ost.frontier_total) <= 1e-6 * ost.frontier_total
    if "bucket_cand" in tier:
        assert st["support_total"] <= 0.5 * ost.support_total   # the support is not materialised (only by sources that fall back to the full merge)
    else:
        assert abs(st["support_total"] - ost.support_total) <= 1e-6 * ost.support_total
    assert st["sources"] == len(src)
    if "bucket" in tier:
        assert st["cluster_sources"] == len(src) and st["redo_sources"] == 0, st
    elif kv.get("push_cluster") == 0:
        assert st["cluster_sources"] == 0 and st["redo_sources"] == 0
    else:
        assert st["cluster_sources"] + st["redo_sources"] == len(src)
        if tier == "cluster_g1_redo":   # one CTA's table at a probe limit of 3 buckets: (nearly) every source is handed over
            assert 0 < st["redo_sources"] <= len(src), st
        elif "redo" in tier:
            assert 0 < st["redo_sources"] < len(src), st
        elif tier == "cluster_g1":   # mean support 13.7 K: a few sources outgrow one CTA's 16 384 slots
            assert st["redo_sources"] <= 0.25 * len(src), st
        else:
            assert st["redo_sources"] == 0, st
    assert st2["sources"] == 2 * len(src)
    assert abs(st2["edges_pushed"] - 2 * ost.edges_pushed) <= 2e-6 * ost.edges_pushed
    for (ac, av), (bc, bv) in zip(og.rows_as_sets(col, val, 32), og.rows_as_sets(col2, val2, 32)):
        if np.array_equal(ac, bc):
            np.testing.assert_allclose(av, bv, rtol=1e-12)


@pytest.mark.parametrize("probe", [1, 3, 16])
@pytest.mark.parametrize("name,mode", [("cora", "ppr"), ("pubmed", "ppr"), ("citeseer", "single")])
def test_gfpush_smem_hash_matches_reference_golden(name, mode, probe):
    """Real graphs through the shared-memory table tier in HBM mode (Cora's ppr support is the whole component)."""
    indptr, indices = load_graph(name)
    z = np.load(os.path.join(GOLDEN, f"gfpush_{name}_{mode}.npz"))
    K, rmax = int(z["K"]), float(z["rmax"])
    with _tuning(push_cluster=0, push_smem_hash=2, push_smem_probe=probe):
        g = _graph(indptr, indices, scratch_mode=HBM)
        row, col, val = _run(g, z["node_idx"].astype(np.int64), z["coef"], rmax, K)
    worst = check_topk_rows(indptr, indices, z["node_idx"], z["coef"], rmax, K, col, val, row=row)
    assert worst < 1e-11, worst


@pytest.mark.parametrize("merge", [0, 1])
@pytest.mark.parametrize("name,mode", [("cora", "ppr"), ("cora", "single"), ("citeseer", "avg"), ("pubmed", "ppr"), ("pubmed", "single")])
def test_gfpush_bucket_kernel_matches_reference_golden(name, mode, merge):
    """Real graphs through the hash-bucket kernel (forced: these supports fit the shared-memory table of the default path)."""
    indptr, indices = load_graph(name)
    z = np.load(os.path.join(GOLDEN, f"gfpush_{name}_{mode}.npz"))
    K, rmax = int(z["K"]), float(z["rmax"])
    with _tuning(push_bucket=2, push_cluster=0, push_bucket_merge=merge):
        g = _graph(indptr, indices, scratch_mode=HBM)
        g.cumulative_stats(reset=True)
        row, col, val = _run(g, z["node_idx"].astype(np.int64), z["coef"], rmax, K)
        st = g.cumulative_stats()
        nb = g.last_stats()["bucket_count"]
        assert nb >= 2 and nb & (nb - 1) == 0          # hash buckets: a power of two
    assert st["cluster_sources"] == len(z["node_idx"]) and st["redo_sources"] == 0
    worst = check_topk_rows(indptr, indices, z["node_idx"], z["coef"], rmax, K, col, val, row=row)
    assert worst < 1e-11, worst
    _, _, _, ost = og.gfpush(indptr, indices, z["node_idx"], z["coef"], rmax, K)
    assert abs(st["edges_pushed"] - ost.edges_pushed) <= 1e-6 * max(ost.edges_pushed, 1)
    assert abs(st["frontier_total"] - ost.frontier_total) <= 1e-6 * ost.frontier_total
    if merge == 1:   # (the candidate merge does not materialise the support)
        assert abs(st["support_total"] - ost.support_total) <= 1e-6 * ost.support_total
    else:
        assert st["support_total"] <= ost.support_total


def test_gfpush_bucket_kernel_pilot_sizes_the_buckets():
    """The first large call on a handle runs two sources per SM as a pilot, reads the largest support back and sizes the hash
    buckets of the rest (and of later calls) from it; both parts write the same rows the oracle computes."""
    indptr, indices = load_graph("pubmed")
    n = indptr.shape[0] - 1
    z = np.load(os.path.join(GOLDEN, "gfpush_pubmed_ppr.npz"))
    K, rmax, coef = int(z["K"]), float(z["rmax"]), z["coef"]
    rng = np.random.default_rng(5)
    src = rng.integers(0, n, 3000).astype(np.int64)
    with _tuning(push_bucket=2, push_cluster=0):
        g = _graph(indptr, indices, scratch_mode=HBM)
        g.cumulative_stats(reset=True)
        row, col, val = _run(g, src, coef, rmax, K)
        st, last = g.cumulative_stats(), g.last_stats()
        assert last["kernel_launches"] == 3                       # pilot + the rest + the (empty) slab pass
        assert st["cluster_sources"] == len(src) and st["redo_sources"] == 0
        nb_first = last["bucket_count"]
        row2, col2, val2 = _run(g, src[:2500], coef, rmax, K)
        last2 = g.last_stats()
        assert last2["kernel_launches"] == 2 and last2["bucket_count"] == nb_first   # the measurement is kept on the handle
    check_topk_rows(indptr, indices, src, coef, rmax, K, col, val, row=row, max_rows=96)
    pick = np.r_[0:40, 280:320, 2960:3000]                        # pilot sources, the seam, the tail
    check_topk_rows(indptr, indices, src[pick], coef, rmax, K, col.reshape(-1, K)[pick].ravel(), val.reshape(-1, K)[pick].ravel(),
                    row=row.reshape(-1, K)[pick].ravel())
    # (the order of the slots within a row is not defined, and fp64 sums depend on the order the pairs arrive in)
    np.testing.assert_allclose(np.sort(val.reshape(-1, K)[:2500], axis=1), np.sort(val2.reshape(-1, K), axis=1), rtol=1e-11, atol=0)


@pytest.mark.parametrize("merge", [0, 1])
@pytest.mark.parametrize("name", ["path8", "star33", "isolated", "dangling"])
def test_gfpush_bucket_kernel_tiny_graphs(name, merge):
    """Dangling nodes, K > support, degree-1 nodes, zero-valued reserves (`single`): graph.h:91-93,113,121 on the bucket kernel."""
    z = np.load(os.path.join(GOLDEN, f"tiny_{name}.npz"))
    reps = 6
    src = np.tile(z["node_idx"], reps)
    with _tuning(push_bucket=2, push_cluster=0, push_bucket_merge=merge):
        g = _graph(z["indptr"], z["indices"], scratch_mode=HBM)
        for tag in sorted({k.split("/")[0] for k in z.files if "/" in k}):
            K, rmax, coef = int(z[f"{tag}/K"]), float(z[f"{tag}/rmax"]), z[f"{tag}/coef"]
            g.cumulative_stats(reset=True)
            row, col, val = _run(g, src, coef, rmax, K)
            st = g.cumulative_stats()
            assert st["cluster_sources"] == len(src), (tag, st)
            check_topk_rows(z["indptr"], z["indices"], src, coef, rmax, K, col, val, row=row)
            ref_filled = np.tile((z[f"{tag}/value"].reshape(-1, K) > 0).sum(1), reps)
            np.testing.assert_array_equal((val.reshape(-1, K) > 0).sum(1), ref_filled)


@pytest.mark.parametrize("name,mode", [("cora", "ppr"), ("citeseer", "avg"), ("pubmed", "ppr"), ("pubmed", "single")])
@pytest.mark.parametrize("cluster", [-1, 2, 16])
def test_gfpush_cluster_kernel_matches_reference_golden(name, mode, cluster):
    """Real graphs through the cluster kernel (forced: these graphs are small enough that auto picks the dense mode)."""
    indptr, indices = load_graph(name)
    z = np.load(os.path.join(GOLDEN, f"gfpush_{name}_{mode}.npz"))
    K, rmax = int(z["K"]), float(z["rmax"])
    with _tuning(push_cluster=cluster):
        g = _graph(indptr, indices, scratch_mode=HBM)
        g.cumulative_stats(reset=True)
        row, col, val = _run(g, z["node_idx"].astype(np.int64), z["coef"], rmax, K)
        st = g.cumulative_stats()
        assert g.last_stats()["cluster_size"] == abs(cluster)
    assert st["cluster_sources"] == len(z["node_idx"]) and st["redo_sources"] == 0
    worst = check_topk_rows(indptr, indices, z["node_idx"], z["coef"], rmax, K, col, val, row=row)
    assert worst < 1e-11, worst
    _, _, _, ost = og.gfpush(indptr, indices, z["node_idx"], z["coef"], rmax, K)
    assert abs(st["edges_pushed"] - ost.edges_pushed) <= 1e-6 * max(ost.edges_pushed, 1)
    assert abs(st["frontier_total"] - ost.frontier_total) <= 1e-6 * ost.frontier_total
    assert abs(st["support_total"] - ost.support_total) <= 1e-6 * ost.support_total


@pytest.mark.parametrize("name", ["path8", "star33", "isolated", "dangling"])
@pytest.mark.parametrize("cluster", [-1, 2, 8])
def test_gfpush_cluster_kernel_tiny_graphs(name, cluster):
    """Dangling nodes, K > support, degree-1 nodes: the edge cases of graph.h:91-93,113,121 on the cluster kernel."""
    z = np.load(os.path.join(GOLDEN, f"tiny_{name}.npz"))
    reps = 6
    src = np.tile(z["node_idx"], reps)
    with _tuning(push_cluster=cluster, push_hub_deg=4):
        g = _graph(z["indptr"], z["indices"], scratch_mode=HBM)
        for tag in sorted({k.split("/")[0] for k in z.files if "/" in k}):
            K, rmax, coef = int(z[f"{tag}/K"]), float(z[f"{tag}/rmax"]), z[f"{tag}/coef"]
            g.cumulative_stats(reset=True)
            row, col, val = _run(g, src, coef, rmax, K)
            st = g.cumulative_stats()
            assert st["cluster_sources"] == len(src), (tag, st)
            check_topk_rows(z["indptr"], z["indices"], src, coef, rmax, K, col, val, row=row)
            ref_filled = np.tile((z[f"{tag}/value"].reshape(-1, K) > 0).sum(1), reps)
            np.testing.assert_array_equal((val.reshape(-1, K) > 0).sum(1), ref_filled)


def test_gfpush_device_path_surfaces_device_side_errors():
    """gp_gfpush_device returns before the kernels run: a source id outside the graph must still raise, either at once
    (check=True) or at the next check_errors(), and must not poison later calls."""
    import torch
    from grandplus_b200._lib import GPError
    indptr, indices = load_graph("cora")
    coef = og.coef_for("ppr", 4, 0.2)
    for scratch in (SMEM, HBM):
        g = _graph(indptr, indices, scratch_mode=scratch)
        bad = torch.tensor([0, 5, 99999, 7], dtype=torch.int32, device="cuda")
        with pytest.raises(GPError):
            g.gfpush_device(bad, coef, 1e-5, 8, check=True)
        row, col, val, _ = g.gfpush_device(bad, coef, 1e-5, 8)        # asynchronous: returns ...
        g.gfpush_device(bad[:2].contiguous(), coef, 1e-5, 8)          # ... and the flag survives a later good call
        with pytest.raises(GPError):
            g.check_errors()
        assert float(val[2].abs().sum()) == 0.0                       # the refused row reads (0, 0, 0.0)
        good = torch.tensor([0, 5, 7], dtype=torch.int32, device="cuda")
        g.gfpush_device(good, coef, 1e-5, 8, check=True)              # clean again


def test_gfpush_two_streams_on_one_handle_are_ordered():
    """A push on another stream must wait for the previous push of the handle (they share the control block, the
    coefficient array and the scratch): host-buffer call, device call on a side stream, device call on the default
    stream, no synchronisation in between -- all three must give the oracle's rows."""
    import torch
    from grandplus_b200 import synth
    indptr, indices = synth.powerlaw_csr(60_000, 700_000, seed=5)
    indptr, indices = indptr.numpy(), indices.numpy()
    coef_a, coef_b = og.coef_for("ppr", 6, 0.05), og.coef_for("avg", 3, 0.2)
    src = synth.sources(60_000, 600, seed=9)
    g = _graph(indptr, indices)
    side = torch.cuda.Stream()
    d_src = src.cuda()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        ra = g.gfpush_device(d_src, coef_a, 1e-5, 32)
    rb = g.gfpush_device(d_src, coef_b, 1e-5, 32)
    with torch.cuda.stream(side):
        rc = g.gfpush_device(d_src, coef_a, 1e-5, 32)
    torch.cuda.synchronize()
    g.check_errors()
    for (row, col, val, _), coef in ((ra, coef_a), (rb, coef_b), (rc, coef_a)):
        worst = check_topk_rows(indptr, indices, src.numpy(), coef, 1e-5, 32, col.cpu().numpy().ravel(),
                                val.cpu().numpy().ravel(), row=row.cpu().numpy().ravel(), max_rows=60)
        assert worst < 1e-11
