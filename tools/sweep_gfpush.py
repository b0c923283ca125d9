#!/usr/bin/env python
"""Sweep GFPush tier settings on one workload (GPU box).  Prints rows/s per configuration.

    python tools/sweep_gfpush.py reddit "push_cluster=0" "push_cluster=2" "push_cluster=4,push_hub_deg=128" ...
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from grandplus_b200 import _lib  # noqa: E402
from grandplus_b200.precompute import propagation  # noqa: E402

DEFAULTS = {"push_bucket": 1, "push_bucket_merge": 0, "push_bucket_nb": 0, "push_bucket_block": 0, "push_bucket_fill": 5, "push_cluster": 0, "push_cluster_probe": 128, "push_hub_deg": 0, "push_max_clusters": 0, "push_smem_hash": 1,
            "push_smem_probe": 2, "push_max_ctas": 0}


def main():
    name = sys.argv[1]
    configs = sys.argv[2:] or ["push_cluster=0", "push_cluster=1"]
    steps = int(os.environ.get("SWEEP_STEPS", "4"))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if os.environ.get("SWEEP_L2_GRAN"):   # cudaLimitMaxL2FetchGranularity = 0x05
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        torch.zeros(1, device=dev)
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["SWEEP_L2_GRAN"])))
        val = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        print("L2 fetch granularity set rc", rc, "now", val.value, flush=True)
    w = dict(bench.WORKLOADS[name])
    if os.environ.get("SWEEP_RMAX"):
        w["rmax"] = float(os.environ["SWEEP_RMAX"])
    S = int(os.environ.get("SWEEP_SOURCES", w["S"]))
    indptr, indices, n = bench.build_workload(name, dev)
    S = min(S, n)
    coef = bench.coef_for(w["mode"], w["order"], w["alpha"])
    batches = bench.source_batches(n, S, steps + 2, 0, 1, dev)
    graph = propagation.Graph.from_device_csr(indptr, indices)
    if os.environ.get("SWEEP_SCRATCH") or os.environ.get("SWEEP_BLOCK") or os.environ.get("SWEEP_PER_SM"):
        graph.configure(scratch_mode=int(os.environ.get("SWEEP_SCRATCH", "0")), block_threads=int(os.environ.get("SWEEP_BLOCK", "0")),
                        ctas_per_sm=int(os.environ.get("SWEEP_PER_SM", "0")))
    for cfg in configs:
        kv = dict(DEFAULTS)
        for item in cfg.split(","):
            if item:
                k, v = item.split("=")
                kv[k] = int(v)
        for k, v in kv.items():
            _lib.set_tuning(k, v)
        for i in range(2):
            graph.gfpush_device(batches[i], coef, w["rmax"], w["K"], want_fp32=True)
        torch.cuda.synchronize()
        graph.cumulative_stats(reset=True)
        graph.phase_cycles(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            graph.gfpush_device(batches[2 + i], coef, w["rmax"], w["K"], want_fp32=True)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 1e3
        st = graph.cumulative_stats(reset=True)
        ls = graph.last_stats()
        print(f"{name} S={S} {cfg:44s} rows/s={S * steps / t:12.0f}  edges/s={st['edges_pushed'] / t / 1e9:7.2f}G  "
              f"G={ls['cluster_size']} nb={ls['bucket_count']} ctas={ls['ctas']} cluster={st['cluster_sources']} redo={st['redo_sources']} "
              f"sup/src={st['support_total'] / max(st['sources'], 1):.0f} scratch={ls['scratch_bytes'] / 1e6:.0f}MB "
              f"[E={st['edges_pushed']} F={st['frontier_total']} S={st['support_total']}]", flush=True)
        ph = graph.phase_cycles(reset=True)
        if ph["resident"]:
            # CTA time per source: resident cycles of all CTAs / sources (a cluster of G CTAs spends G x its latency)
            us = ph["resident"] / 1965.0 / max(st["sources"], 1)
            print(f"    {us:.0f} us of CTA time per source; % by phase: " +
                  " ".join(f"{n}={100.0 * v / ph['resident']:.1f}" for n, v in ph.items() if n not in ("resident", "wide_expand", "wide_settle")), flush=True)
    for k, v in DEFAULTS.items():
        _lib.set_tuning(k, v)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"sweep wall {time.time() - t0:.1f}s")
