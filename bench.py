#!/usr/bin/env python
"""bench.py -- GRAND+ propagation hot path on B200: GFPush source rows/s and Pi.X aggregation GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload reddit] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of S sources of the workload graph:
  (1) GFPush + top-k of the S sources (gp_gfpush_device; Pi rows stay in HBM), then
  (2) the DropNode-masked Pi.X aggregation of those S rows, `sample`=2 augmentations in one launch
      (gp_aggregate_fwd, slot layout, p=0.5 -- /root/reference/model.py:321-322).
`value` is whole-job source rows/s with inputs resident in HBM; `e2e` is the same metric through the
reference-facing calls with HOST (pinned) buffers: Graph.gfpush_omp(host arrays) [H2D node ids, D2H
row/col/value], upload of the batch's (neighbor, score) arrays as /root/reference/model.py:314-316
does, the fused aggregation, and a D2H read of the result checksum.  Multi-GPU: sources are sharded
over ranks on a replicated CSR with no data-path collective (weak scaling: S per rank is fixed).

`--impl reference` times the reference's own CPU implementation (oracle/_ref = its propagation.cpp
compiled unmodified; falls back to the oracle port) on the host cores, on a bounded source sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "gfpush_source_rows_per_s"
UNIT = "rows/s"

# BASELINE.json configs.  Parameters: scripts/run_<dataset>.sh line 7 (ppr); K for reddit per BASELINE.
WORKLOADS = {
    "cora": dict(config=0, real="cora", F=1433, mode="ppr", order=20, alpha=0.2, rmax=1e-7, K=32, S=2708),
    "pubmed": dict(config=1, real="pubmed", F=500, mode="ppr", order=6, alpha=0.5, rmax=1e-5, K=16, S=19717),
    "reddit": dict(config=2, n=232_965, draws=11_606_919, F=602, mode="ppr", order=6, alpha=0.05, rmax=1e-5, K=32,
                   S=16384),
    "amazon2m": dict(config=3, n=2_449_029, draws=61_859_140, F=100, mode="ppr", order=6, alpha=0.2, rmax=1e-6,
                     K=64, S=16384),
    # MAG-Scholar-C shape (scripts/run_mag.sh:7); F = the hidden width: on the model_mag path the table the
    # aggregation reads is the [N, hidden] output of MLP.emb, not the 2.78M-column sparse attribute matrix
    "mag": dict(config=4, n=10_541_560, draws=265_219_994, F=64, mode="ppr", order=10, alpha=0.2, rmax=1e-5, K=32,
                S=16384),
    "small": dict(config=-1, n=50_000, draws=600_000, F=64, mode="ppr", order=6, alpha=0.05, rmax=1e-5, K=32, S=2048),
}
DROPNODE_P = 0.5   # run_model.py:56 default
N_AUG = 2          # run_model.py --sample default (model.py:321)


def coef_for(mode, order, alpha):
    """model.py:255-267."""
    if mode == "avg":
        c = np.ones(order + 1)
    elif mode == "ppr":
        c = alpha * (1 - alpha) ** np.arange(order + 1)
    else:
        c = np.zeros(order + 1); c[-1] = 1.0
    return (c / c.sum()).astype(np.float64)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload, kernel):
    """dram bytes per launch from the committed ncu --set full capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p))[workload][kernel]
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []      # (arrival time, csv line)
        self.proc = None
        self.windows = []      # [t0, t1] of the timed regions

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.time(), line.strip()))
        except Exception:
            pass

    def wait_first_sample(self, timeout=5.0):
        t_end = time.time() + timeout
        while not self.samples and time.time() < t_end:
            time.sleep(0.01)

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        time.sleep(0.06)   # let the last periodic sample arrive
        if self.proc:
            self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for t, s in self.samples:
            if not any(t0 <= t <= t1 + 0.05 for t0, t1 in self.windows):
                continue
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": min(sm) if sm else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "device-timed steps + end-to-end steps (both under load)"}


# ----------------------------------------------------------------------------------------------
def build_workload(name, device):
    """(indptr, indices) int32 torch tensors on `device`, features tensor [n,F] on `device`."""
    import torch
    from grandplus_b200 import synth
    w = WORKLOADS[name]
    if "real" in w:
        z = np.load(os.path.join(ROOT, "tests", "golden", f"graph_{w['real']}.npz"))
        indptr = torch.from_numpy(z["indptr"]).to(device)
        indices = torch.from_numpy(z["indices"]).to(device)
    else:
        indptr, indices = synth.powerlaw_csr(w["n"], w["draws"], seed=0, device=device,
                                             relabel=os.environ.get("GP_BENCH_NO_RELABEL") is None)
    n = int(indptr.numel() - 1)
    return indptr, indices, n


def source_batches(n, S, count, rank, world, device):
    """`count` batches of S distinct sources for this rank: consecutive slices of one seeded
    permutation of the nodes (wrapping), disjoint across ranks within a step."""
    from grandplus_b200 import synth
    perm = synth.sources(n, n, seed=1, device=device)
    out = []
    for i in range(count):
        start = ((i * world + rank) * S) % n
        idx = (start + np.arange(S)) % n
        import torch
        out.append(perm[torch.as_tensor(idx, device=perm.device)].contiguous())
    return out


def algorithmic_bytes_gfpush(stats, S_total, K):
    """E_push*(4+8) + F_tot*(8+8) + K*16 per source."""
    return stats["edges_pushed"] * 12 + stats["frontier_total"] * 16 + S_total * K * 16


# Second roof of the GFPush kernels (the HBM-bandwidth roofline is reported first; these say what actually binds):
#  * residues in shared memory (MODE 1 / MODE 2 / cluster kernel): one pushed edge = one find-or-claim + fp64 add in a
#    shared-memory table; measured 0.6 per clock per SM at a table load of 0.8 (profiles/r01_smem_hash_microbench.txt);
#  * residues on the HBM slabs (MODE 0): one pushed edge = one fp64 atomic on a footprint far beyond L2, measured
#    20.9 G/s (profiles/r01_random_access_microbench.txt).
SMEM_PUSH_PER_CLK_PER_SM = 0.6
DRAM_ATOMICS_PER_S = 20.9e9


def second_roof(stats_last, edges_per_s, sm_count, sm_mhz):
    """Mode-appropriate roof in pushed edges/s; frac is capped at 1 (a roof is never exceeded, only mis-measured)."""
    on_chip = stats_last.get("scratch_mode") == 1 or stats_last.get("table_slots", 0) > 0
    if on_chip:
        roof = SMEM_PUSH_PER_CLK_PER_SM * sm_count * sm_mhz * 1e6
        kind = "shared-memory find-or-claim + fp64 add (0.6 / clk / SM, profiles/r01_smem_hash_microbench.txt)"
    else:
        roof = DRAM_ATOMICS_PER_S
        kind = "fp64 atomics on an HBM-resident footprint (profiles/r01_random_access_microbench.txt)"
    return {"kind": kind, "achieved_edges_per_s": edges_per_s, "roof_edges_per_s": roof, "frac": min(edges_per_s / roof, 1.0)}


def measure_workload(name, dev, rank, world, steps, warmup, sources=0, configure=None, e2e=False, sampler=None):
    """GFPush + aggregation of `steps` batches of S sources of workload `name` on this rank (weak scaling: every rank
    takes its own batches).  Returns a dict with the device-timed step, the work counters, both rooflines and -- when
    `e2e` -- the same metric through the host-buffer API."""
    import torch
    import torch.distributed as dist
    from grandplus_b200 import dist as gd
    from grandplus_b200 import model as gm
    from grandplus_b200 import synth
    from grandplus_b200.precompute import propagation

    w = WORKLOADS[name]
    K = w["K"]
    coef = coef_for(w["mode"], w["order"], w["alpha"])
    t_setup = time.time()
    indptr, indices, n = build_workload(name, dev)
    S = min(sources or w["S"], n)
    graph = propagation.Graph.from_device_csr(indptr, indices)
    if configure:
        graph.configure(**configure)
    feats = gm.DeviceFeatures(synth.features(n, w["F"], seed=1, device=dev))
    total = warmup + steps
    batches = source_batches(n, S, total, rank, world, dev)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for i in range(total):
        if i == warmup:
            torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
            graph.check_errors()
            graph.cumulative_stats(reset=True)
            graph.phase_cycles(reset=True)
            t0 = time.perf_counter()
            w0 = time.time()
        timed = i >= warmup
        if timed:
            ev[i - warmup][0].record()
        _row, col, _val, val32 = graph.gfpush_device(batches[i], coef, w["rmax"], K, want_fp32=True)
        if timed:
            ev[i - warmup][1].record()
        out = gm.aggregate_slots(feats, col, val32, None, DROPNODE_P, True, n_aug=N_AUG, seed=1234, offset=i)
        if timed:
            ev[i - warmup][2].record()
        del out, col, val32, _row, _val
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if sampler is not None:
        sampler.mark(w0, time.time())
    graph.check_errors()   # the device path is asynchronous: a refused source or an overflow must not go unnoticed
    t_push = sum(e[0].elapsed_time(e[1]) for e in ev) / 1e3
    t_agg = sum(e[1].elapsed_time(e[2]) for e in ev) / 1e3
    t_dev = ev[0][0].elapsed_time(ev[-1][2]) / 1e3   # CUDA events on the launching stream, first to last timed launch
    stats = graph.cumulative_stats(reset=True)
    last = graph.last_stats()
    phases = graph.phase_cycles(reset=True)
    step_time = gd.max_over_ranks(t_dev, device=dev)   # a multi-GPU step takes as long as its slowest rank
    wall = gd.max_over_ranks(wall, device=dev)

    # algorithmic bytes of the aggregation (untimed post-pass: the masks are counter-based, so the kept-entry count of
    # every timed step can be regenerated exactly)
    agg_bytes = 0
    slots = S * K
    for i in range(warmup, total):
        _row, col, _val, val32 = graph.gfpush_device(batches[i], coef, w["rmax"], K, want_fp32=True)
        mask = gm.dropnode_mask(slots, N_AUG, DROPNODE_P, 1234, i, dev)
        kept_any = ((mask.sum(0) > 0) & (val32.reshape(-1) > 0)).sum().item()
        agg_bytes += kept_any * w["F"] * 4 + slots * 8 + (S + 1) * 4 + N_AUG * S * w["F"] * 4
    graph.cumulative_stats(reset=True)
    graph.phase_cycles(reset=True)
    push_bytes = algorithmic_bytes_gfpush(stats, S * steps, K)
    peak, peak_src = load_peaks()
    props = torch.cuda.get_device_properties(dev)
    sm_mhz = 1965.0
    try:
        sm_mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        pass
    kernel = ("gfpush_cluster_kernel" if last.get("cluster_size", 0) > 0 else
              "gfpush_bucket_kernel" if last.get("bucket_count", 0) > 0 else "gfpush_kernel")
    edges_per_s = stats["edges_pushed"] / t_push
    resident = max(phases.get("resident", 0), 1)
    roof_push = {"kernel": kernel, "bound": "hbm", "achieved": push_bytes / t_push / 1e9, "peak": peak,
                 "unit": "GB/s", "frac": push_bytes / t_push / 1e9 / peak,
                 "traffic": load_traffic(name, kernel), "peak_source": peak_src,
                 "ms_per_launch": t_push / steps * 1e3, "share_of_step": t_push / t_dev,
                 "edges_per_s": edges_per_s, "edges_per_source": stats["edges_pushed"] / (S * steps),
                 "algorithmic_bytes_per_launch": push_bytes / steps,
                 "frontier_per_source": stats["frontier_total"] / (S * steps),
                 "support_per_source": stats["support_total"] / (S * steps),
                 "cluster_size": last.get("cluster_size", 0), "cluster_sources": stats.get("cluster_sources"),
                 "handed_over_sources": stats.get("redo_sources"), "persistent_ctas": last["ctas"],
                 "scratch_mode": {1: "smem", 2: "hbm"}.get(last["scratch_mode"]),
                 # SM cycles per phase summed over the persistent CTAs, as fractions of their residency: attributes
                 # box-to-box differences (gp_gfpush_phase_cycles)
                 "phase_share": {k: v / resident for k, v in phases.items() if k != "resident"},
                 "cta_us_per_source": phases.get("resident", 0) / sm_mhz / max(stats["sources"], 1),
                 "table_slots": last.get("table_slots", 0), "bucket_count": last.get("bucket_count", 0),
                 "second_roof": second_roof(last, edges_per_s, props.multi_processor_count, sm_mhz)}
    roof_agg = {"kernel": "aggregate_fwd_kernel", "bound": "hbm", "achieved": agg_bytes / t_agg / 1e9, "peak": peak,
                "unit": "GB/s", "frac": agg_bytes / t_agg / 1e9 / peak,
                "traffic": load_traffic(name, "aggregate_fwd_kernel"), "peak_source": peak_src,
                "ms_per_launch": t_agg / steps * 1e3, "share_of_step": t_agg / t_dev,
                "algorithmic_bytes_per_launch": agg_bytes / steps, "rows_per_s": S * steps / t_agg}
    res = {"name": name, "w": w, "n": n, "nnz": int(indices.numel()), "S": S, "K": K, "steps": steps, "warmup": warmup,
           "value": S * steps * world / step_time, "step_time": step_time, "wall": wall, "t_push": t_push, "t_agg": t_agg,
           "t_dev": t_dev, "stats": stats, "last": last, "roof_push": roof_push, "roof_agg": roof_agg, "setup_s": t_setup,
           "x_mb": feats.data.numel() * 4 / 1e6, "csr_mb": (indices.numel() + indptr.numel()) * 4 / 1e6,
           "launches": (int(last.get("kernel_launches") or 1) + 1) * steps}
    if e2e:
        res["e2e"] = measure_e2e(graph, feats, batches, coef, w, S, K, dev, world, steps, warmup, sampler)
        res["indptr_host"], res["indices_host"] = indptr.cpu().numpy(), indices.cpu().numpy()
    del graph, feats, batches, indptr, indices
    torch.cuda.empty_cache()
    return res


def measure_e2e(graph, feats, batches, coef, w, S, K, dev, world, steps, warmup, sampler):
    """The same metric through the reference-facing calls with pinned HOST buffers: Graph.gfpush_omp(host arrays)
    [H2D node ids, D2H row/col/value], upload of the batch's (neighbour, score) arrays as
    /root/reference/model.py:314-316 does, the fused aggregation, and a D2H read of the result checksum."""
    import torch
    import torch.distributed as dist
    from grandplus_b200 import dist as gd
    from grandplus_b200 import model as gm

    def barrier():
        if world > 1:
            dist.barrier()

    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)  # noqa: E731
    h_node = pin((S,), torch.int32)
    h_row, h_col, h_val = pin((S * K,), torch.int32), pin((S * K,), torch.int32), pin((S * K,), torch.float64)
    total = len(batches)
    host_batches = [b.cpu() for b in batches]
    torch.cuda.synchronize(); barrier()
    e2e_steps = max(2, min(steps, 10))

    def one_step(i, stage=None):
        def tick(name):
            if stage is not None:
                torch.cuda.synchronize()
                now = time.perf_counter()
                stage[name] = stage.get(name, 0.0) + now - tick.t
                tick.t = now
        tick.t = time.perf_counter()
        h_node.copy_(host_batches[(warmup + i) % total])
        tick("host_copy_node_ids")
        graph.gfpush_omp(h_node.numpy(), h_row.numpy(), h_col.numpy(), h_val.numpy(), coef, w["rmax"], K)
        tick("gfpush_omp_h2d_kernel_d2h")
        d_col = h_col.to(dev, non_blocking=True).reshape(S, K)              # model.py:314-316: the batch's
        d_val = h_val.to(dev, non_blocking=True).reshape(S, K).float()      # neighbour ids and scores go H2D
        tick("h2d_col_score")
        out = gm.aggregate_slots(feats, d_col, d_val, None, DROPNODE_P, True, n_aug=N_AUG, seed=1234, offset=i)
        tick("aggregate")
        chk = float(out.sum().item())                                       # D2H read of the step's result
        tick("checksum_d2h")
        return chk

    for i in range(2 + e2e_steps):
        if i == 2:
            torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            w0 = time.time()
        chk = one_step(i)
    torch.cuda.synchronize(); barrier()
    e2e_time = gd.max_over_ranks(time.perf_counter() - t0, device=dev)
    if sampler is not None:
        sampler.mark(w0, time.time())
    # untimed diagnostic pass with a synchronise between the stages: where an end-to-end step spends its time
    stage = {}
    for i in range(3):
        one_step(i, stage)
    return {"value": S * e2e_steps * world / e2e_time, "unit": UNIT, "h2d_bytes_per_step": S * 4 + S * K * 12,
            "d2h_bytes_per_step": S * K * 16 + 4, "steps": e2e_steps, "ms_per_step": e2e_time / e2e_steps * 1e3,
            "api": "Graph.gfpush_omp(host arrays) + aggregate_slots(H2D col/score) + checksum D2H", "checksum": chk,
            "stage_ms": {k: v / 3 * 1e3 for k, v in stage.items()}}


def measure_strong(name, dev, rank, world, total_sources, steps=3):
    """Strong scaling of the product multi-GPU entry point: a FIXED set of sources is sharded over the ranks
    (dist.gfpush_sharded) and every rank ends up with all rows (all_gather_into_tensor over NCCL inside the timed
    region) -- the reference's contract of one [S*K] result set (/root/reference/model.py:252-268)."""
    import torch
    import torch.distributed as dist
    from grandplus_b200 import dist as gd
    from grandplus_b200 import synth
    from grandplus_b200.precompute import propagation
    w = WORKLOADS[name]
    coef = coef_for(w["mode"], w["order"], w["alpha"])
    indptr, indices, n = build_workload(name, dev)
    graph = propagation.Graph.from_device_csr(indptr, indices)
    S = min(total_sources, n)
    src = synth.sources(n, S, seed=11, device=dev)
    t_all, t_push, t_gather = [], [], []
    for i in range(2 + steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        col, val, val32, (lo, hi) = gd.gfpush_sharded(graph, src, coef, w["rmax"], w["K"], gather=False, check=False)
        e[1].record()
        if world > 1:
            col, val, val32 = gd.all_gather_rows([col, val, val32], S)
        e[2].record()
        torch.cuda.synchronize()
        if i >= 2:
            t_all.append(e[0].elapsed_time(e[2]) / 1e3); t_push.append(e[0].elapsed_time(e[1]) / 1e3)
            t_gather.append(e[1].elapsed_time(e[2]) / 1e3)
        assert col.shape[0] == S
    graph.check_errors()
    t = gd.max_over_ranks(sum(t_all), device=dev)
    tp = gd.max_over_ranks(sum(t_push), device=dev)
    tp_min = -gd.max_over_ranks(-sum(t_push), device=dev)
    tg = gd.max_over_ranks(sum(t_gather), device=dev)
    del graph
    torch.cuda.empty_cache()
    return {"workload": name, "total_sources": S, "steps": steps, "rows_per_s": S * steps / t, "ms_per_step": t / steps * 1e3,
            "gfpush_ms_slowest_rank": tp / steps * 1e3, "gfpush_ms_fastest_rank": tp_min / steps * 1e3,
            "collective": "all_gather_into_tensor of [S,K] x (int32 col + f64 val + f32 val) over NCCL",
            "collective_ms": tg / steps * 1e3, "collective_share": tg / t,
            "gathered_bytes_per_step": S * w["K"] * 16}


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    configure = None
    if args.scratch or args.block or args.ctas_per_sm:
        configure = dict(scratch_mode=args.scratch, block_threads=args.block, ctas_per_sm=args.ctas_per_sm)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first_sample()
    r = measure_workload(args.workload, dev, rank, world, args.steps, args.warmup, sources=args.sources,
                         configure=configure, e2e=True, sampler=sampler)
    clocks = sampler.stop()
    w, n, S, K = r["w"], r["n"], r["S"], r["K"]
    stats, roof_push, roof_agg = r["stats"], r["roof_push"], r["roof_agg"]
    dominant = roof_push if r["t_push"] >= r["t_agg"] else roof_agg
    strong = None
    if world > 1 or args.scaling == "strong":
        strong = measure_strong(args.workload, dev, rank, world, args.strong_sources)
    headline_strong = args.scaling == "strong" and strong is not None
    line = {
        "metric": METRIC, "value": strong["rows_per_s"] if headline_strong else r["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": strong["ms_per_step"] if headline_strong else r["step_time"] / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if headline_strong else "weak", "vs_baseline": None,
        "dtype": "f64 (GFPush residues) / f32 (aggregation)", "data": "synthetic" if "real" not in w else
        f"real {w['real']} graph (Planetoid, tests/golden) + synthetic N(0,1) features",
        "config": {"workload": f"BASELINE configs[{w['config']}] {args.workload}", "nodes": n,
                   "csr_nnz": r["nnz"], "features": w["F"], "prop_mode": w["mode"], "order": w["order"],
                   "alpha": w["alpha"], "rmax": w["rmax"], "top_k": K, "sources_per_step_per_gpu": S,
                   "dropnode_rate": DROPNODE_P, "augmentations": N_AUG, "parallelism": f"source-sharded x{world}, CSR+X replicated",
                   "l2": "inputs larger than L2 (X %.0f MB, CSR %.0f MB, per-CTA scratch %.0f MB); new sources every step"
                         % (r["x_mb"], r["csr_mb"], r["last"]["scratch_bytes"] / 1e6),
                   "scratch_mode": roof_push["scratch_mode"], "persistent_ctas": roof_push["persistent_ctas"],
                   "cluster_size": roof_push["cluster_size"]},
        "roofline": dominant, "roofline_gfpush": roof_push, "roofline_aggregate": roof_agg,
        "aggregation_gb_per_s": roof_agg["achieved"],
        "e2e": r["e2e"], "gpu_launches": r["launches"], "clocks": clocks,
        "setup_s": r["setup_s"], "impl": "ours", "wall_ms_per_step": r["wall"] / args.steps * 1e3,
    }
    if strong is not None:
        line["strong_scaling"] = strong
    indptr_host, indices_host = r.pop("indptr_host"), r.pop("indices_host")
    # the other BASELINE configs, measured beside the headline at every N (weak scaling, same step definition)
    if not args.no_side:
        if rank == 0 and world == 1 and args.workload != "pubmed":
            line["config1_pubmed"] = pubmed_side_measurement(dev)
        for key, wl in (("config3_amazon2m", "amazon2m"), ("config4_mag", "mag")):
            if wl == args.workload:
                continue
            side_sampler = ClockSampler(local)   # (the side measurements run after the headline: their own clock samples)
            side_sampler.start()
            side_sampler.wait_first_sample()
            sr = measure_workload(wl, dev, rank, world, steps=3, warmup=3, sampler=side_sampler)
            side_clocks = side_sampler.stop()
            side_clocks["window"] = "the side measurement's device-timed steps"
            line[key] = {"workload": f"BASELINE configs[{sr['w']['config']}] {wl}", "value": sr["value"], "unit": UNIT,
                         "ms_per_step": sr["step_time"] / sr["steps"] * 1e3, "steps": sr["steps"], "warmup": sr["warmup"],
                         "sources_per_step_per_gpu": sr["S"], "nodes": sr["n"], "csr_nnz": sr["nnz"], "rmax": sr["w"]["rmax"],
                         "order": sr["w"]["order"], "top_k": sr["K"], "features": sr["w"]["F"], "scaling": "weak",
                         "roofline": sr["roof_push"], "roofline_aggregate": sr["roof_agg"], "setup_s": sr["setup_s"],
                         "clocks": side_clocks}
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.workload, indptr_host, indices_host, n, budget_s=args.cpu_budget)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit_line(line)


def pubmed_side_measurement(dev, steps=5):
    """BASELINE configs[1] -- Pubmed, ppr/avg/single (scripts/run_pubmed.sh:7,11,15) on one B200 -- reported
    beside the headline workload.  The Pubmed graph and X are L2-resident, so L2 is flushed (a 512 MB write)
    between timed steps.  All 19 717 nodes are the sources of every step."""
    import torch
    from grandplus_b200 import model as gm
    from grandplus_b200 import synth
    from grandplus_b200.precompute import propagation
    z = np.load(os.path.join(ROOT, "tests", "golden", "graph_pubmed.npz"))
    indptr = torch.from_numpy(z["indptr"]).to(dev)
    indices = torch.from_numpy(z["indices"]).to(dev)
    n = int(indptr.numel() - 1)
    graph = propagation.Graph.from_device_csr(indptr, indices)
    feats = gm.DeviceFeatures(synth.features(n, 500, seed=1, device=dev))
    src = synth.sources(n, n, seed=1, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out = {"nodes": n, "csr_nnz": int(indices.numel()), "sources_per_step": n, "top_k": 16, "rmax": 1e-5,
           "l2": "flushed between steps (512 MB write)", "steps": steps}
    for mode, order, alpha in (("ppr", 6, 0.5), ("avg", 4, 0.2), ("single", 2, 0.2)):
        coef = coef_for(mode, order, alpha)
        t_push = t_agg = 0.0
        for i in range(3 + steps):
            flush.fill_(i & 0xFF)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            _r, col, _v, val32 = graph.gfpush_device(src, coef, 1e-5, 16, want_fp32=True)
            e[1].record()
            res = gm.aggregate_slots(feats, col, val32, None, DROPNODE_P, True, n_aug=N_AUG, seed=1234, offset=i)
            e[2].record()
            torch.cuda.synchronize()
            if i >= 3:
                t_push += e[0].elapsed_time(e[1]) / 1e3
                t_agg += e[1].elapsed_time(e[2]) / 1e3
            del res, col, val32, _r, _v
        out[mode] = {"order": order, "alpha": alpha, "gfpush_rows_per_s": n * steps / t_push,
                     "aggregate_rows_per_s": n * steps / t_agg, "rows_per_s": n * steps / (t_push + t_agg)}
    return out


# ----------------------------------------------------------------------------------------------
def _cpu_step(name, indptr, indices, n, src, X, use_ref, threads=None):
    """One pass of the reference CPU path over `src`: GFPush (the reference's own module -- 40 OpenMP threads as
    graph.h:41 hard-codes, or `threads` when given: the same binary with omp_set_num_threads called after its
    constructor; or the oracle port) + the host gather / scatter aggregation of those rows for N_AUG augmentations
    (model.py:80-87,314 restated with index_add_)."""
    import torch
    from oracle import gfpush as og
    w = WORKLOADS[name]
    coef = coef_for(w["mode"], w["order"], w["alpha"])
    K = w["K"]
    t0 = time.perf_counter()
    if use_ref:
        row, col, val = og.reference_gfpush(indptr, indices, src, coef, w["rmax"], K, nthreads=threads)
    else:
        row, col, val, _ = og.gfpush(indptr, indices, src, coef, w["rmax"], K, nthreads=threads or 0)
    t1 = time.perf_counter()
    S = len(src)
    idx = torch.arange(S).repeat_interleave(K)
    colt = torch.from_numpy(col.astype(np.int64))
    sc = torch.from_numpy(val.astype(np.float32))
    live = sc > 0
    idx, colt, sc = idx[live], colt[live], sc[live]
    for _ in range(N_AUG):
        m = torch.nn.functional.dropout(sc, DROPNODE_P, True)
        f = X[colt]                                              # model.py:314 host gather
        num = torch.zeros(S, X.shape[1]).index_add_(0, idx, f * m[:, None])
        den = torch.zeros(S, 1).index_add_(0, idx, m[:, None])
        _ = num / (den + 1e-12)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def _cpu_sources_per_step(per_source_s, steps_total, budget_s, n, floor=4096):
    """Sources per CPU step: BASELINE.md 3 asks for a 4 096-source sample on the large graphs; more when the budget
    allows, fewer only when 4 096 per step would push the whole run beyond ~5 minutes."""
    S = int(budget_s / steps_total / per_source_s)
    if S < floor:
        S = min(floor, max(16, int(300.0 / steps_total / per_source_s)))
    return int(max(16, min(n, S)))


def cpu_baseline(name, indptr, indices, n, budget_s=15.0):
    """The reference's CPU path on this box's host cores, on a bounded sample of the same workload, with both thread
    settings BASELINE.md 3 asks for: the literal 40 of graph.h:41 and NUMTHREAD = cpu_count."""
    import torch
    from oracle import gfpush as og
    w = WORKLOADS[name]
    use_ref = og.reference_available()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(1)
    X = torch.randn(n, w["F"])
    probe = rng.choice(n, size=min(n, 16), replace=False).astype(np.int32)
    tp, ta = _cpu_step(name, indptr, indices, n, probe, X, use_ref)
    per = max((tp + ta) / len(probe), 1e-6)
    S = _cpu_sources_per_step(per, 2, budget_s, n)
    src = rng.choice(n, size=S, replace=False).astype(np.int32)
    tp40, ta = _cpu_step(name, indptr, indices, n, src, X, use_ref)
    tpc, ta2 = _cpu_step(name, indptr, indices, n, src, X, use_ref, threads=cores)
    ta = min(ta, ta2)
    tp = min(tp40, tpc)
    return {"value": S / (tp + ta), "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{S} sources of the same graph in one call (GFPush {tp:.2f} s + aggregation {ta:.2f} s)",
            "gfpush_rows_per_s": S / tp, "aggregate_rows_per_s": S / ta,
            "gfpush_rows_per_s_40_threads": S / tp40, "gfpush_rows_per_s_cpu_count_threads": S / tpc,
            "threads": f"GFPush: better of the reference's hard-coded 40 OpenMP threads (graph.h:41) and {cores} = cpu_count "
                       f"(same binary, omp_set_num_threads after its constructor); torch aggregation uses all {cores} cores"
            if use_ref else f"{cores} OpenMP threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from grandplus_b200 import synth
    from oracle import gfpush as og
    w = WORKLOADS[args.workload]
    if "real" in w:
        z = np.load(os.path.join(ROOT, "tests", "golden", f"graph_{w['real']}.npz"))
        indptr, indices = z["indptr"], z["indices"]
    else:
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        ip, ix = synth.powerlaw_csr(w["n"], w["draws"], seed=0, device=dev)
        indptr, indices = ip.cpu().numpy(), ix.cpu().numpy()
    n = len(indptr) - 1
    use_ref = og.reference_available()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    X = torch.randn(n, w["F"])
    rng = np.random.default_rng(1)
    probe = rng.choice(n, size=min(n, 16), replace=False).astype(np.int32)
    # thread setting: the better of the reference's literal 40 and cpu_count on this box (both reported)
    t40, _ = _cpu_step(args.workload, indptr, indices, n, probe, X, use_ref)
    probe2 = rng.choice(n, size=min(n, 256), replace=False).astype(np.int32)
    t40, ta40 = _cpu_step(args.workload, indptr, indices, n, probe2, X, use_ref)
    tcc, tacc = _cpu_step(args.workload, indptr, indices, n, probe2, X, use_ref, threads=cores)
    threads = cores if tcc < t40 else None
    per = max((min(t40, tcc) + min(ta40, tacc)) / len(probe2), 1e-6)
    total = args.steps + args.warmup
    S = _cpu_sources_per_step(per, total, args.ref_budget, n)
    times = []
    for i in range(total):
        src = rng.choice(n, size=S, replace=False).astype(np.int32)
        tp, ta = _cpu_step(args.workload, indptr, indices, n, src, X, use_ref, threads=threads)
        if i >= args.warmup:
            times.append((tp, ta))
    t = sum(a + b for a, b in times)
    value = S * args.steps / t
    kind = "reference" if use_ref else "port"
    sample = f"{S} sources per step of the same graph, {args.steps} steps"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (GFPush residues) / f32 (aggregation)",
            "data": "synthetic" if "real" not in w else f"real {w['real']} graph + synthetic features",
            "config": {"workload": f"BASELINE configs[{w['config']}] {args.workload}", "nodes": n, "csr_nnz": int(len(indices)),
                       "features": w["F"], "prop_mode": w["mode"], "order": w["order"], "alpha": w["alpha"],
                       "rmax": w["rmax"], "top_k": w["K"], "dropnode_rate": DROPNODE_P, "augmentations": N_AUG,
                       "sources_per_step": S},
            "impl": "reference",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "openmp_threads": threads or 40,
                             "threads": "the better of the reference's hard-coded 40 (graph.h:41) and cpu_count on a 256-source probe: "
                                        f"40 -> {len(probe2) / t40:.0f} rows/s, {cores} -> {len(probe2) / tcc:.0f} rows/s (GFPush alone)",
                             "gfpush_rows_per_s": S * args.steps / sum(a for a, _ in times),
                             "aggregate_rows_per_s": S * args.steps / sum(b for _, b in times)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)


_REAL_STDOUT = None


def emit_line(line: dict) -> None:
    """The contract is ONE JSON line on stdout; libraries (NCCL's version banner, ...) write to fd 1
    too, so fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reddit", choices=sorted(WORKLOADS))
    ap.add_argument("--sources", type=int, default=0, help="sources per step per GPU (default: workload's)")
    ap.add_argument("--scratch", type=int, default=0, help="0 auto, 1 smem, 2 hbm")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=60.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: S sources per step per GPU (default); strong: the headline is a fixed source set sharded "
                         "over the ranks with the rows all-gathered (also reported beside the weak line when N > 1)")
    ap.add_argument("--strong-sources", type=int, default=1 << 17, help="source set of the strong-scaling measurement")
    ap.add_argument("--no-side", action="store_true", help="skip the Pubmed / Amazon2M-shape / MAG-shape side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
