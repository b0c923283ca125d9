#!/usr/bin/env python
"""Exact propagation (predict, model.py:181-212) on the GPU: one D^-1 A H round over all nodes, timed with CUDA
events, reported as algorithmic GB/s (nnz*F*4 gathered + nnz*8 entry metadata + (N+1)*4 + N*F*4 written)
against the measured HBM peak.   python tools/bench_predict.py [reddit|amazon2m|pubmed|cora] [rounds]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from grandplus_b200 import model as gm, predict as gp, synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "reddit"
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    w = bench.WORKLOADS[name]
    indptr, indices, n = bench.build_workload(name, dev)
    adj = gp.DeviceAdjacency((indptr, indices))
    X = gm.DeviceFeatures(synth.features(n, w["F"], seed=1, device=dev))
    peak, src = bench.load_peaks()
    H = torch.zeros((n + adj.n_chunks, X.ld), dtype=torch.float32, device=dev)
    H[:n] = X.data
    nxt = torch.empty_like(H)
    for _ in range(2):
        gp._round(adj, H, X.F, nxt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        gp._round(adj, H, X.F, nxt)
        H, nxt = nxt, H
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / rounds
    byts = adj.nnz * X.F * 4 + adj.nnz * 8 + (n + 1) * 4 + n * X.F * 4
    t0 = time.time()
    out = gp.propagate_exact(adj, X, w["order"], w["alpha"], w["mode"])
    torch.cuda.synchronize()
    t_all = time.time() - t0
    print(json.dumps({"workload": name, "nodes": n, "nnz": adj.nnz, "F": X.F, "ms_per_round": ms,
                      "algorithmic_GBps": byts / ms / 1e6, "frac_of_hbm_peak": byts / ms / 1e6 / peak, "peak": peak,
                      "peak_source": src, "hub_chunks": adj.n_chunks, "max_degree": int((indptr[1:] - indptr[:-1]).max().item()),
                      "predict_propagation_s": t_all, "order": w["order"], "mode": w["mode"],
                      "checksum": float(out.double().sum().item())}))


if __name__ == "__main__":
    main()
