import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from grandplus_b200 import _lib, synth
from grandplus_b200.precompute import propagation
from oracle import gfpush as og
n, draws, order, alpha, rmax, k, S = 2_449_029, 61_859_140, 6, 0.2, 1e-6, 64, 768
indptr, indices = synth.powerlaw_csr(n, draws, seed=0, device="cuda")
graph = propagation.Graph.from_device_csr(indptr, indices)
src = synth.sources(n, S, seed=7, device="cuda")
deg = indptr[1:] - indptr[:-1]
src[0] = deg.argmax().to(torch.int32)
src = src.contiguous()
coef = og.coef_for("ppr", order, alpha)
D = dict(push_cluster=0, push_smem_hash=1, push_bucket=1, push_bucket_merge=0, push_bucket_block=0, push_bucket_nb=0)
def run(**kv):
    for a, b in {**D, **kv}.items(): _lib.set_tuning(a, b)
    graph.cumulative_stats(reset=True)
    row, col, val, _ = graph.gfpush_device(src, coef, rmax, k, want_fp32=True, check=True)
    st = graph.cumulative_stats(); ls = graph.last_stats()
    return col.cpu().numpy().reshape(S, k), val.cpu().numpy().reshape(S, k), st, ls
cb, vb, _, _ = run(push_smem_hash=0, push_bucket=0)
for name, kv in (("cand_b512", dict(push_smem_hash=0, push_bucket=2, push_bucket_block=512)),
                 ("full_b512", dict(push_smem_hash=0, push_bucket=2, push_bucket_block=512, push_bucket_merge=1)),
                 ("cand_b512_nb64", dict(push_smem_hash=0, push_bucket=2, push_bucket_block=512, push_bucket_nb=64)), ("cand_b512_nb32", dict(push_smem_hash=0, push_bucket=2, push_bucket_block=512, push_bucket_nb=32)), ("full_b512_nb32", dict(push_smem_hash=0, push_bucket=2, push_bucket_block=512, push_bucket_nb=32, push_bucket_merge=1)), ("cand_b512_nb16", dict(push_smem_hash=0, push_bucket=2, push_bucket_block=512, push_bucket_nb=16)),
                 ("cand_b1024", dict(push_smem_hash=0, push_bucket=2, push_bucket_block=1024))):
    ca, va, st, ls = run(**kv)
    bad = 0; detail = []
    for i in range(S):
        a = dict(zip(ca[i], va[i])); b = dict(zip(cb[i], vb[i]))
        miss = [x for x in b if x not in a]
        if miss:
            cut = vb[i].min()
            if max(b[x] for x in miss) > cut * (1 + 1e-9):
                bad += 1
                if len(detail) < 3: detail.append((i, int(src[i]), len(miss), float(max(b[x] for x in miss) / cut), max(abs(a[x] - b[x]) / b[x] for x in a if x in b)))
    print(name, "rows with a missed node above the cut:", bad, "nb", ls["bucket_count"], "redo", st["redo_sources"], "support_total", st["support_total"], detail, flush=True)
