import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grandplus_b200 import _lib, synth
from grandplus_b200.precompute import propagation
from oracle import gfpush as og
from tests.helpers import check_topk_rows
indptr, indices = synth.powerlaw_csr(20_000, 200_000, seed=5)
indptr, indices = indptr.numpy(), indices.numpy()
src = synth.sources(20_000, 64, seed=4).numpy()
coef = og.coef_for("ppr", 5, 0.1)
"""Small GFPush run for compute-sanitizer (memcheck / racecheck) over every residue-table mode:
    compute-sanitizer --tool racecheck python tools/sanitize_gfpush.py"""
for kv in (dict(push_cluster=0, push_smem_hash=2, push_smem_probe=4), dict(push_cluster=0, push_smem_hash=2, push_smem_probe=1),
           dict(push_cluster=0, push_smem_hash=0, push_bucket=0), dict(scratch=1),
           dict(push_cluster=-1), dict(push_cluster=2), dict(push_cluster=4, push_hub_deg=8), dict(push_cluster=16),
           dict(push_cluster=2, push_cluster_probe=1), dict(push_cluster=0, push_bucket=2, push_bucket_merge=1),
           dict(push_cluster=0, push_bucket=2, push_bucket_merge=0), dict(push_cluster=0, push_bucket=2, push_bucket_merge=0, push_bucket_nb=8),
           dict(push_cluster=0, push_bucket=2, push_bucket_merge=0, push_bucket_block=512), dict(push_cluster=0, push_bucket=2, push_bucket_merge=1, push_bucket_block=256)):
    if os.environ.get("SANITIZE_ONLY") and not any(os.environ["SANITIZE_ONLY"] in k for k in kv): continue   # e.g. SANITIZE_ONLY=bucket
    scratch = kv.pop("scratch", 2)
    for k, v in {**dict(push_bucket=1, push_bucket_merge=0, push_bucket_nb=0, push_bucket_block=0, push_cluster_probe=128, push_hub_deg=0), **kv}.items(): _lib.set_tuning(k, v)
    g = propagation.Graph(indptr, indices, 0); g.configure(scratch_mode=scratch)
    S, K = len(src), 16
    row = np.zeros(S*K, np.int32); col = np.zeros(S*K, np.int32); val = np.zeros(S*K, np.float64)
    g.gfpush_omp(src, row, col, val, coef, 1e-4, K)
    print(kv, 'scratch', scratch, check_topk_rows(indptr, indices, src, coef, 1e-4, K, col, val, row=row, max_rows=32), g.last_stats()['support_total'])
# a low-degree graph (Cora, mean degree 3.9): the bucket kernel's packing loop for short push-list entries, at every geometry
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "graph_cora.npz"))
ip2, ix2 = z["indptr"].astype(np.int32), z["indices"].astype(np.int32)
src2 = np.arange(0, 2708, 43, dtype=np.int32)
coef2 = og.coef_for("ppr", 8, 0.2)
for block in (1024, 512, 256):
    if os.environ.get("SANITIZE_ONLY") and "bucket" not in os.environ["SANITIZE_ONLY"]: continue
    for k, v in dict(push_cluster=0, push_bucket=2, push_bucket_merge=0, push_bucket_nb=0, push_bucket_block=block).items(): _lib.set_tuning(k, v)
    g = propagation.Graph(ip2, ix2, 0); g.configure(scratch_mode=2)
    S, K = len(src2), 16
    row = np.zeros(S*K, np.int32); col = np.zeros(S*K, np.int32); val = np.zeros(S*K, np.float64)
    g.gfpush_omp(src2, row, col, val, coef2, 1e-6, K)
    print('cora bucket block', block, check_topk_rows(ip2, ix2, src2, coef2, 1e-6, K, col, val, row=row, max_rows=32), g.last_stats()['bucket_count'])
