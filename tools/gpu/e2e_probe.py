"""Where does Graph.gfpush_omp (host buffers) spend its time?  Reddit-shape, 16 384 sources, K = 32."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from grandplus_b200.precompute import propagation
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["reddit"]
indptr, indices, n = bench.build_workload("reddit", dev)
graph = propagation.Graph.from_device_csr(indptr, indices)
coef = bench.coef_for(w["mode"], w["order"], w["alpha"])
S, K = 16384, 32
batches = bench.source_batches(n, S, 8, 0, 1, dev)
pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
h_row, h_col, h_val = pin((S * K,), torch.int32), pin((S * K,), torch.int32), pin((S * K,), torch.float64)
p_row, p_col, p_val = np.zeros(S * K, np.int32), np.zeros(S * K, np.int32), np.zeros(S * K, np.float64)
hb = [b.cpu().numpy() for b in batches]
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
i = [0]
def dev_only():
    i[0] += 1; graph.gfpush_device(batches[i[0] % 8], coef, w["rmax"], K)
def pinned():
    i[0] += 1; graph.gfpush_omp(hb[i[0] % 8], h_row.numpy(), h_col.numpy(), h_val.numpy(), coef, w["rmax"], K)
def pageable():
    i[0] += 1; graph.gfpush_omp(hb[i[0] % 8], p_row, p_col, p_val, coef, w["rmax"], K)
def dev_then_copy():
    i[0] += 1
    r, c, v, _ = graph.gfpush_device(batches[i[0] % 8], coef, w["rmax"], K, want_fp32=False)
    h_row.copy_(r.reshape(-1), non_blocking=True); h_col.copy_(c.reshape(-1), non_blocking=True); h_val.copy_(v.reshape(-1), non_blocking=True)
    torch.cuda.synchronize()
print("gfpush_device (kernel only)        %.2f ms" % t(dev_only))
print("gfpush_omp, pinned host arrays     %.2f ms" % t(pinned))
print("gfpush_omp, pageable numpy arrays  %.2f ms" % t(pageable))
print("gfpush_device + 3 torch D2H copies %.2f ms" % t(dev_then_copy))
print("gfpush_omp, pinned host arrays     %.2f ms" % t(pinned))
for _ in range(3):
    pinned(); print({k: graph.last_stats()[k] for k in ("bucket_count", "table_slots", "kernel_launches", "redo_sources", "ctas", "scratch_bytes")})
