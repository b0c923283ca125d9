"""CPU tests: the C-ABI library loads and exports every declared symbol (no compute without a GPU),
the product path refuses to run without CUDA, host-side helpers, the synthetic generator, and the
source/batch sharding logic under a world_size-2 gloo group."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "grandplus_b200.h")).read()
    return sorted(set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_header_symbol():
    from grandplus_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/grandplus_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.gp_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", out), f"{s} is not a defined text symbol"


def test_library_is_sm100a_only():
    from grandplus_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_cuda():
    import torch
    from grandplus_b200 import _lib
    from grandplus_b200.precompute import propagation
    if _lib.load().gp_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.GPError):
        propagation.Graph(np.array([0, 1, 2], np.int32), np.array([0, 1], np.int32), 0)
    from grandplus_b200 import model as gm
    with pytest.raises(RuntimeError):
        gm.random_prop(torch.zeros(4, 3), torch.ones(4), torch.tensor([0, 0, 1, 1]), 0.5)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "grand-plus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "libgp_oracle" not in text and "oracle/_" not in text, f"{f} loads oracle binaries"


def test_tuning_keys_are_validated_on_the_host():
    """gp_set_tuning is pure host code: every documented key is accepted with its default, out-of-range values and
    unknown keys are refused (the kernels never see an unchecked geometry)."""
    from grandplus_b200 import _lib
    defaults = {"push_bucket": 1, "push_bucket_block": 0, "push_bucket_fill": 5, "push_bucket_nb": 0, "push_bucket_merge": 0,
                "push_cluster": 0, "push_cluster_probe": 128, "push_hub_deg": 0, "push_max_clusters": 0, "push_max_ctas": 0,
                "push_smem_hash": 1, "push_smem_probe": 2}
    for k, v in defaults.items():
        _lib.set_tuning(k, v)
    for k, bad in (("push_bucket_block", 384), ("push_bucket_block", 2048), ("push_bucket_fill", 8), ("push_bucket_fill", 2),
                   ("push_bucket", 3), ("push_bucket_nb", 257), ("push_bucket_merge", 2), ("push_cluster", 3),
                   ("push_smem_probe", 0), ("no_such_key", 1)):
        with pytest.raises(_lib.GPError):
            _lib.set_tuning(k, bad)
    for geometry in (256, 512, 1024, 0):
        _lib.set_tuning("push_bucket_block", geometry)
    header = open(os.path.join(ROOT, "include", "grandplus_b200.h")).read()
    for k in defaults:
        assert f'"{k}"' in header, f"tuning key {k} is not documented in the header"


def test_philox_known_answer():
    """Philox4x32-10 KATs from the Random123 distribution (kat_vectors): the DropNode masks are
    exactly reproducible from (seed, offset)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def philox(c, k):
        c = list(c); k = list(k)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        return c

    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_synthetic_powerlaw_graph_shape():
    from grandplus_b200 import synth
    indptr, indices = synth.powerlaw_csr(5000, 40000, seed=1)
    indptr, indices = indptr.numpy(), indices.numpy()
    n = 5000
    assert indptr[0] == 0 and indptr[-1] == len(indices) and np.all(np.diff(indptr) >= 1)
    rows = np.repeat(np.arange(n), np.diff(indptr))
    import scipy.sparse as sp
    a = sp.csr_matrix((np.ones(len(indices)), indices, indptr), (n, n))
    assert (a != a.T).nnz == 0                              # symmetric
    assert np.all(a.diagonal() == 1)                        # + I (model.py:243)
    assert np.all(np.diff(indices)[np.diff(rows) == 0] > 0)  # sorted, de-duplicated
    deg = np.diff(indptr)
    assert deg.max() > 20 * np.median(deg)                  # heavy tail
    i2, x2 = synth.powerlaw_csr(5000, 40000, seed=1)
    assert np.array_equal(i2.numpy(), indptr) and np.array_equal(x2.numpy(), indices)   # deterministic
    src = synth.sources(n, 100, seed=1).numpy()
    assert len(np.unique(src)) == 100


def test_shard_ranges_cover_and_balance():
    from grandplus_b200 import dist as gd
    for total in (0, 1, 7, 1000, 12345):
        for world in (1, 2, 3, 8):
            spans = [gd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from grandplus_b200 import dist as gd
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
S, K = 11, 4
lo, hi = gd.shard_range(S, rank, 2)
col = torch.arange(lo * K, hi * K, dtype=torch.int32).reshape(-1, K)
val = col.to(torch.float64) * 0.5
gcol, gval = gd.all_gather_rows([col, val], S, group=None)
assert gcol.shape == (S, K) and torch.equal(gcol.reshape(-1), torch.arange(S * K, dtype=torch.int32))
assert torch.equal(gval, gcol.to(torch.float64) * 0.5)
t = gd.max_over_ranks(1.0 + rank)
assert t == 2.0
# row-sparse gradient all-reduce (SURVEY 8e, MAG backward): ranks touch overlapping rows, ragged counts
rows = torch.tensor([2, 5, 9], dtype=torch.int64) if rank == 0 else torch.tensor([5, 7], dtype=torch.int64)
vals = torch.arange(rows.numel() * 3, dtype=torch.float32).reshape(-1, 3) + 10 * rank
ur, uv = gd.allreduce_sparse_rows(rows, vals)
assert ur.tolist() == [2, 5, 7, 9]
want = {2: [0., 1., 2.], 5: [3. + 10., 4. + 11., 5. + 12.], 7: [13., 14., 15.], 9: [6., 7., 8.]}
for r, v in zip(ur.tolist(), uv.tolist()):
    assert v == want[r], (r, v)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_world_size_2_gloo_shard_and_gather(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"ok {r}" in o


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` runs on the host cores alone: one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "1", "--warmup", "0", "--ref-budget", "2"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "gfpush_source_rows_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
