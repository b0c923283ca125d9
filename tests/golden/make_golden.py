#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

What it runs, unmodified:
  * the reference's pybind11 module (precompute/propagation.cpp + graph.h) compiled by
    oracle/Makefile into oracle/_ref, called exactly as /root/reference/model.py:249-268;
  * the reference's graph loading lines (/root/reference/utils/data_loader.py:118-120)
    and the self-loop line (/root/reference/model.py:243) on the Planetoid pickles that
    ship in /root/reference/dataset/citation;
  * ``Grand_Plus.random_prop`` imported from /root/reference/model.py and ``MLP.emb``
    imported from /root/reference/model_mag.py, executed on CPU.  torch_scatter 2.0.6 is a
    third-party dependency that is not vendored and not installable here, so a stand-in
    module implementing its documented ``scatter(src, index, dim, dim_size, reduce='sum')``
    (= zeros(dim_size).scatter_add_(dim, broadcast(index), src)) is injected before the
    import; ``Tensor.cuda`` is patched to the identity for the generating run only.

Nothing here is used at test time except the .npz files it writes.
"""
import os
import pickle as pkl
import sys
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import gfpush as og  # noqa: E402

# (order, alpha, rmax, K) from scripts/run_{cora,citeseer,pubmed}.sh lines 7, 11, 15
PARAMS = {
    "cora": {"ppr": (20, 0.2, 1e-7, 32), "avg": (4, 0.2, 1e-7, 32), "single": (2, 0.2, 1e-7, 32)},
    "citeseer": {"ppr": (10, 0.4, 1e-7, 32), "avg": (2, 0.2, 1e-7, 32), "single": (2, 0.2, 1e-7, 32)},
    "pubmed": {"ppr": (6, 0.5, 1e-5, 16), "avg": (4, 0.2, 1e-5, 16), "single": (2, 0.2, 1e-5, 16)},
}
N_SOURCES = 96


def load_planetoid_graph(name):
    import networkx as nx
    with open(os.path.join(REF, "dataset", "citation", f"ind.{name}.graph"), "rb") as f:
        graph = pkl.load(f, encoding="latin1")
    adj = nx.adjacency_matrix(nx.from_dict_of_lists(graph))          # data_loader.py:118
    adj = adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)  # data_loader.py:120
    adj = adj + sp.eye(adj.shape[0])                                   # model.py:243
    adj = sp.csr_matrix(adj)
    adj.sort_indices()
    return np.array(adj.indptr, dtype=np.int32), np.array(adj.indices, dtype=np.int32)  # model.py:249-250


def pick_sources(indptr, rng):
    n = indptr.shape[0] - 1
    deg = np.diff(indptr)
    forced = [0, int(np.argmax(deg)), int(np.argmin(deg)), n - 1]
    rest = rng.choice(n, size=N_SOURCES, replace=False)
    out = []
    for v in forced + list(rest):
        if v not in out:
            out.append(int(v))
    return np.asarray(out[:N_SOURCES], dtype=np.int32)


def tiny_graphs():
    """Hand-checkable graphs: path, star, isolated+self-loop, a true dangling
    node (degree 0: only reachable when the caller skips `+I`), a directed chain into it."""
    g = {}

    def csr(n, edges, self_loops=True, symmetric=True):
        a = sp.lil_matrix((n, n))
        for u, v in edges:
            a[u, v] = 1
            if symmetric:
                a[v, u] = 1
        a = sp.csr_matrix(a)
        if self_loops:
            a = sp.csr_matrix(a + sp.eye(n))
        a.sort_indices()
        return np.array(a.indptr, dtype=np.int32), np.array(a.indices, dtype=np.int32)

    g["path8"] = csr(8, [(i, i + 1) for i in range(7)])
    g["star33"] = csr(33, [(0, i) for i in range(1, 33)])
    g["isolated"] = csr(6, [(0, 1), (1, 2), (2, 0)])          # nodes 3,4,5: self-loop only (deg 1)
    g["dangling"] = csr(5, [(0, 1), (1, 2), (2, 3), (0, 4)], self_loops=False, symmetric=False)  # 3,4: deg 0
    return g


def gen_gfpush():
    rng = np.random.default_rng(20221017)
    for name, modes in PARAMS.items():
        indptr, indices = load_planetoid_graph(name)
        np.savez_compressed(os.path.join(HERE, f"graph_{name}.npz"), indptr=indptr, indices=indices)
        src = pick_sources(indptr, rng)
        for mode, (order, alpha, rmax, K) in modes.items():
            coef = og.coef_for(mode, order, alpha)
            row, col, val = og.reference_gfpush(indptr, indices, src, coef, rmax, K)
            np.savez_compressed(os.path.join(HERE, f"gfpush_{name}_{mode}.npz"), node_idx=src, coef=coef,
                                rmax=np.float64(rmax), K=np.int32(K), row_idx=row, col_idx=col, value=val)
            print(f"{name:9s} {mode:6s} N={indptr.shape[0]-1} nnz={indices.shape[0]} S={src.shape[0]} "
                  f"filled={(val > 0).sum()}")
    for name, (indptr, indices) in tiny_graphs().items():
        n = indptr.shape[0] - 1
        src = np.arange(n, dtype=np.int32)
        cases = {}
        for mode, order, alpha, rmax, K in [("ppr", 4, 0.2, 1e-3, 4), ("ppr", 6, 0.5, 0.0, 64),
                                           ("avg", 3, 0.2, 1e-2, 3), ("single", 2, 0.2, 1e-7, 5),
                                           ("single", 1, 0.2, 0.05, 2)]:
            coef = og.coef_for(mode, order, alpha)
            row, col, val = og.reference_gfpush(indptr, indices, src, coef, rmax, K)
            tag = f"{mode}_o{order}_r{rmax:g}_k{K}"
            cases[f"{tag}/coef"] = coef
            cases[f"{tag}/rmax"] = np.float64(rmax)
            cases[f"{tag}/K"] = np.int32(K)
            cases[f"{tag}/row_idx"] = row
            cases[f"{tag}/col_idx"] = col
            cases[f"{tag}/value"] = val
        np.savez_compressed(os.path.join(HERE, f"tiny_{name}.npz"), indptr=indptr, indices=indices,
                            node_idx=src, **cases)
        print(f"tiny {name}: n={n} nnz={indices.shape[0]} cases={len(cases)//6}")


def _install_scatter_standin():
    import torch

    def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
        assert reduce == "sum" and out is None
        if dim_size is None:
            dim_size = int(index.max()) + 1
        size = list(src.shape)
        size[dim] = int(dim_size)
        index = index.expand_as(src) if index.dim() == src.dim() else index
        return torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, index, src)

    mod = types.ModuleType("torch_scatter")
    mod.scatter = scatter
    sys.modules["torch_scatter"] = mod


def gen_aggregate():
    import torch
    import torch.nn.functional as F
    _install_scatter_standin()
    # `from precompute import propagation` at model.py:9 -> serve the compiled reference module
    pre = types.ModuleType("precompute")
    pre.propagation = og.load_reference()
    sys.modules["precompute"] = pre
    sys.modules["precompute.propagation"] = pre.propagation
    sys.path.insert(0, REF)
    import model as ref_model          # /root/reference/model.py
    import model_mag as ref_model_mag  # /root/reference/model_mag.py
    torch.Tensor.cuda = lambda self, *a, **k: self  # generating run only: the host has no GPU

    rng = np.random.default_rng(7)
    out = {}
    # (tag, B rows, max entries per row, F, p, training)
    cases = [("eval_f100", 37, 16, 100, 0.5, False), ("train_f100", 37, 16, 100, 0.5, True),
             ("train_f602", 12, 32, 602, 0.5, True), ("train_f1433", 6, 24, 1433, 0.5, True),
             ("train_p09_f64", 40, 8, 64, 0.9, True), ("train_p0_f7", 11, 5, 7, 0.0, True)]
    for tag, B, kmax, Fdim, p, training in cases:
        counts = rng.integers(1, kmax + 1, size=B)
        idx = np.repeat(np.arange(B), counts).astype(np.int64)   # ascending, last row non-empty
        nz = idx.shape[0]
        feats = rng.standard_normal((nz, Fdim)).astype(np.float32)
        scores = (rng.random(nz).astype(np.float32) ** 3 + 1e-4).astype(np.float32)
        m = ref_model.Grand_Plus(Fdim, 3, 8, 2, False, 0.0, 0.0, dropnode_rate=p)
        m.train(training)
        t_feats, t_scores, t_idx = torch.from_numpy(feats), torch.from_numpy(scores), torch.from_numpy(idx)
        torch.manual_seed(1234)
        dropped = F.dropout(t_scores, p=p, training=training)     # the mask random_prop will draw
        mask = (dropped != 0).numpy().astype(np.uint8)
        torch.manual_seed(1234)
        res = m.random_prop(t_feats, t_scores, t_idx, p)         # /root/reference/model.py:80-87
        out[f"{tag}/feats"] = feats
        out[f"{tag}/scores"] = scores
        out[f"{tag}/idx"] = idx
        out[f"{tag}/p"] = np.float64(p)
        out[f"{tag}/training"] = np.int32(training)
        out[f"{tag}/mask"] = mask
        out[f"{tag}/out"] = res.numpy()
        print(f"random_prop {tag}: nz={nz} kept={int(mask.sum())} out={tuple(res.shape)}")
    np.savez_compressed(os.path.join(HERE, "random_prop.npz"), **out)

    # MLP.emb (model_mag.py:48-55), eval mode and input_droprate = 0 as scripts/run_mag.sh:7 sets it
    out = {}
    for tag, n_nodes, amax, n_attr, H in [("h64", 50, 12, 1000, 64), ("h16", 13, 40, 300, 16)]:
        counts = rng.integers(1, amax + 1, size=n_nodes)
        node_idx = np.repeat(np.arange(n_nodes), counts).astype(np.int64)
        nza = node_idx.shape[0]
        attr_idx = rng.integers(0, n_attr, size=nza).astype(np.int64)
        attr_data = (rng.random(nza).astype(np.float32) + 0.1).astype(np.float32)
        torch.manual_seed(99)
        mlp = ref_model_mag.MLP(n_attr, H, H, 1, False, 0.0, 0.0, False)
        mlp.eval()
        table = mlp.embeds.weight.detach().numpy().copy()
        res = mlp.emb(torch.from_numpy(attr_idx), torch.from_numpy(node_idx), torch.from_numpy(attr_data))
        out[f"{tag}/table"] = table
        out[f"{tag}/attr_idx"] = attr_idx
        out[f"{tag}/node_idx"] = node_idx
        out[f"{tag}/attr_data"] = attr_data
        out[f"{tag}/out"] = res.detach().numpy()
        print(f"emb {tag}: nza={nza} out={tuple(res.shape)}")
    np.savez_compressed(os.path.join(HERE, "emb.npz"), **out)


def gen_predict():
    """The matrix the reference's own ``predict`` (model.py:181-224) hands to the MLP, captured by
    intercepting ``get_local_logits``; real Cora / Citeseer graphs (+I), random 12-column features."""
    import torch
    _install_scatter_standin()
    pre = types.ModuleType("precompute")
    pre.propagation = og.load_reference()
    sys.modules.setdefault("precompute", pre)
    sys.modules.setdefault("precompute.propagation", pre.propagation)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import model as ref_model
    captured = []

    def fake_logits(mlp, feat, batch_size=10000):
        captured.append(np.array(feat))
        return np.zeros((feat.shape[0], 2), dtype=np.float32)

    ref_model.get_local_logits = fake_logits
    rng = np.random.default_rng(11)
    out = {}
    for name in ("cora", "citeseer"):
        z = np.load(os.path.join(HERE, f"graph_{name}.npz"))
        n = len(z["indptr"]) - 1
        adj = sp.csr_matrix((np.ones(len(z["indices"])), z["indices"], z["indptr"]), shape=(n, n))
        X = rng.standard_normal((n, 12)).astype(np.float32)
        out[f"{name}/X"] = X
        for mode, order, alpha in (("ppr", 6, 0.2), ("avg", 4, 0.2), ("single", 2, 0.2)):
            args = types.SimpleNamespace(order=order, alpha=alpha)
            net = types.SimpleNamespace(eval=lambda: None, mlp=None)
            captured.clear()
            ref_model.predict(args, adj, X.copy(), net, np.arange(4), torch.zeros(n, dtype=torch.long), mode=mode)
            tag = f"{name}/{mode}_o{order}_a{alpha:g}"
            out[f"{tag}/order"] = np.int32(order)
            out[f"{tag}/alpha"] = np.float64(alpha)
            out[f"{tag}/feat"] = captured[0].astype(np.float32)
            print(f"predict {tag}: feat {captured[0].shape} {captured[0].dtype}")
    np.savez_compressed(os.path.join(HERE, "predict.npz"), **out)


if __name__ == "__main__":
    if not og.reference_available():
        raise SystemExit("build the reference first: make -C oracle ref")
    if "--only-predict" in sys.argv:
        gen_predict()
        raise SystemExit(0)
    gen_gfpush()
    gen_aggregate()
    gen_predict()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"{f}: {os.path.getsize(os.path.join(HERE, f))/1024:.1f} KiB")
