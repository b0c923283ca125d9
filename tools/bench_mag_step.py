#!/usr/bin/env python
"""The MAG-path hot ops of one training batch (model_mag.py:337-369): MLP.emb over the attribute nonzeros of the batch's
neighbour rows (forward + backward into the embedding table) and random_prop with autograd, at the reference's batch shape
(scripts/run_mag.sh: batch_size 20 + unlabel_batch_size 20 rows, top_k 32, hidden 64, 2 augmentations) and at a
bandwidth-sized batch.  GPU: the fused kernels with the [2 784 240, 64] table resident in HBM.  CPU: the reference's
formulation restated with index_add_ on the host cores (embedding gather, scatter-mean, scatter-sum, dense grad).

    python tools/bench_mag_step.py [attrs_per_row]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from grandplus_b200 import model as gm  # noqa: E402

N_ATTR, H, K = 2_784_240, 64, 32


def make_batch(B, attrs, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    nz = B * K                                              # neighbour rows of the batch (Pi entries)
    nza = nz * attrs
    node_idx = torch.arange(nz).repeat_interleave(attrs)    # ascending, model_mag.py:345-347
    attr_idx = torch.randint(0, N_ATTR, (nza,), generator=g)
    attr_data = torch.ones(nza)                             # MAG attributes are not binarised but are ones in practice
    mat_idx = torch.arange(B).repeat_interleave(K)
    scores = torch.rand(nz, generator=g) + 0.01
    return [t.to(dev) for t in (attr_idx, node_idx, attr_data, scores, mat_idx)]


def gpu_step(weight, batch, n_aug=2, sparse=False):
    attr_idx, node_idx, attr_data, scores, mat_idx = batch
    emb = gm.emb(weight, attr_idx, node_idx, attr_data, sparse_grad=sparse)   # model_mag.py:355
    loss = 0.0
    for a in range(n_aug):                                                    # model_mag.py:354-357
        out = gm.random_prop(emb, scores, mat_idx, 0.5, training=True, seed=3, offset=a)
        loss = loss + out.square().mean()
    loss.backward()                                                           # through both reductions into weight.grad
    return loss


def cpu_step(weight, batch, n_aug=2):
    attr_idx, node_idx, attr_data, scores, mat_idx = batch
    nz, B = int(node_idx[-1]) + 1, int(mat_idx[-1]) + 1
    e = weight[attr_idx] * attr_data[:, None]
    num = torch.zeros(nz, H).index_add_(0, node_idx, e)
    den = torch.zeros(nz, 1).index_add_(0, node_idx, attr_data[:, None])
    emb = num / (den + 1e-10)
    loss = 0.0
    for a in range(n_aug):
        m = torch.nn.functional.dropout(scores, 0.5, True)
        o = torch.zeros(B, H).index_add_(0, mat_idx, emb * m[:, None]) / (torch.zeros(B, 1).index_add_(0, mat_idx, m[:, None]) + 1e-12)
        loss = loss + o.square().mean()
    loss.backward()
    return loss


def main():
    attrs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    w_gpu = torch.nn.Parameter(torch.randn(N_ATTR, H, device=dev) * 0.01)
    w_cpu = torch.nn.Parameter(w_gpu.detach().cpu().clone())
    torch.set_num_threads(os.cpu_count() or 1)
    res = {"attrs_per_row": attrs, "table": [N_ATTR, H], "cores": os.cpu_count()}
    for tag, B, reps in (("reference batch (40 rows)", 40, 50), ("4096 rows", 4096, 10)):
        batch = make_batch(B, attrs, dev)
        for _ in range(3):
            w_gpu.grad = None
            gpu_step(w_gpu, batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            w_gpu.grad = None          # the dense grad is re-created (zero-filled) every step, as autograd does
            gpu_step(w_gpu, batch)
        e1.record(); torch.cuda.synchronize()
        t_gpu = e0.elapsed_time(e1) / reps
        # the whole optimisation step of the table: dense gradient + torch.optim.Adam(fused) over 2.78 M rows (the reference's
        # semantics on the GPU) against the row-sparse gradient + SparseRowAdam (the same parameters, lazily)
        from grandplus_b200.optim import SparseRowAdam
        dense_opt = torch.optim.Adam([w_gpu], lr=0.01, fused=True)
        timings = {}
        for mode in ("dense", "sparse"):
            w_s = w_gpu if mode == "dense" else torch.nn.Parameter(w_gpu.detach().clone())
            opt = dense_opt if mode == "dense" else SparseRowAdam(w_s, lr=0.01)
            for i in range(3 + reps):
                if i == 3:
                    torch.cuda.synchronize(); e0.record()
                if mode == "sparse":
                    opt.prepare(batch[0])
                w_s.grad = None
                gpu_step(w_s, batch, sparse=(mode == "sparse"))
                opt.step()
            e1.record(); torch.cuda.synchronize()
            timings[mode] = e0.elapsed_time(e1) / reps
            del opt
            if mode == "sparse":
                del w_s
        dense_opt = None
        cb = [t.cpu() for t in batch]
        n_cpu = 3 if B <= 64 else 1
        t0 = time.perf_counter()
        for _ in range(n_cpu):
            w_cpu.grad = None
            cpu_step(w_cpu, cb)
        t_cpu = (time.perf_counter() - t0) / n_cpu * 1e3
        nza = int(batch[0].numel())
        res[tag] = {"attr_nonzeros": nza, "gpu_ms_fwd_bwd": t_gpu, "cpu_ms_fwd_bwd": t_cpu, "speedup": t_cpu / t_gpu,
                    "gathered_MB": nza * H * 4 / 1e6, "gpu_ms_step_dense_grad_fused_adam": timings["dense"],
                    "gpu_ms_step_sparse_grad_lazy_adam": timings["sparse"]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
