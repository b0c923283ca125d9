"""Multi-GPU product path on real devices (needs >= 2 GPUs; skipped otherwise): dist.gfpush_sharded over NCCL must give
every rank the rows a single GPU computes -- the reference's contract of one [S*K] result set (model.py:252-268)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
from grandplus_b200 import dist as gd, synth
from grandplus_b200.precompute import propagation
from oracle import gfpush as og
indptr, indices = synth.powerlaw_csr(60_000, 700_000, seed=5, device="cuda")
graph = propagation.Graph.from_device_csr(indptr, indices)
src = synth.sources(60_000, 1001, seed=4, device="cuda")          # odd count: ragged shards
coef = og.coef_for("ppr", 6, 0.05)
col, val, val32, (lo, hi) = gd.gfpush_sharded(graph, src, coef, 1e-5, 32, gather=True)
graph.check_errors()
assert col.shape == (1001, 32) and val.shape == (1001, 32) and val32.shape == (1001, 32)
_r, c1, v1, _ = graph.gfpush_device(src, coef, 1e-5, 32, check=True)    # the same sources on this GPU alone
a = og.rows_as_sets(col.cpu().numpy().ravel(), val.cpu().numpy().ravel(), 32)
b = og.rows_as_sets(c1.cpu().numpy().ravel(), v1.cpu().numpy().ravel(), 32)
same = 0
for (ac, av), (bc, bv) in zip(a, b):
    assert len(ac) == len(bc)
    if np.array_equal(ac, bc):
        same += 1
        np.testing.assert_allclose(av, bv, rtol=1e-11, atol=0)
    else:
        assert abs(av.min() - bv.min()) <= 1e-9 * bv.min()        # only a tie at the cut may differ
assert same >= 0.85 * 1001, same
# every rank holds identical gathered rows
chk = torch.stack([val.sum(), col.double().sum()])
lst = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(lst, chk)
assert all(torch.equal(x, lst[0]) for x in lst)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("world", [2])
def test_gfpush_sharded_over_nccl_equals_single_gpu(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        assert f"ok {r}" in o
