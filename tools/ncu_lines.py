#!/usr/bin/env python
"""Per-SOURCE-LINE view of an ncu capture: executed instructions and stall samples of one kernel,
attributed to .cu/.cuh lines through nvdisasm's line table of the in-tree library (-lineinfo).

    python tools/ncu_lines.py X.ncu-rep <mangled-kernel-substring> [top N] [demangled-substring]

The SASS page of `ncu --page source --csv` has no line column; offsets within the kernel match
nvdisasm's /*offset*/ comments of the same build, so the .so must be the one that was profiled.
An instruction is charged to the innermost inlined line.  The mangled substring selects the function in
the cubin (e.g. gfpush_hash_kernelILi1024ELb0); the optional demangled one selects the kernel inside the
report when it holds several (default: the first).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "grand-plus_b200", "libgrandplus_b200.so")


def line_table(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, active = None, None, False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                active = kernel_sub in m.group(1) and not table
                cur_fn = m.group(1)
                continue
            if not active:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)), "inlined" in m.group(3))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
            if m and cur_line:
                table[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
        if table:
            break
    return table


def main():
    rep, sub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 30
    table = line_table(sub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the report may hold several kernels: take the block whose name matches
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    dsub = sys.argv[4] if len(sys.argv) > 4 else ""
    blk = next(b for b in blocks if dsub in b["name"])
    hdr = blk["rows"][0]
    ix = {k: i for i, k in enumerate(hdr)}
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    base = None
    per = defaultdict(lambda: {"inst": 0, "samples": 0, "stalls": defaultdict(int)})
    tot_i = tot_s = 0
    for r in blk["rows"][1:]:
        if len(r) != len(hdr):
            continue
        addr = int(r[ix["Address"]], 16)
        if base is None:
            base = addr
        ent = table.get(addr - base)
        key = ent[0][:2] if ent else ("?", 0)
        n_i = int(r[ix["Instructions Executed"]] or 0)
        n_s = int(r[ix["# Samples"]] or 0)
        per[key]["inst"] += n_i
        per[key]["samples"] += n_s
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                per[key]["stalls"][c[6:]] += v
        tot_i += n_i
        tot_s += n_s
    src_cache = {}

    def src(fn, ln):
        if fn not in src_cache:
            p = os.path.join(ROOT, "grand-plus_b200", "csrc", fn)
            src_cache[fn] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[fn]
        return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""

    print(f"kernel: {blk['name']}\nwarp instructions executed: {tot_i}   stall samples: {tot_s}\n")
    for title, keyf in (("by stall samples", lambda kv: -kv[1]["samples"]), ("by instructions executed", lambda kv: -kv[1]["inst"])):
        print(f"--- top source lines {title}")
        for (fn, ln), d in sorted(per.items(), key=keyf)[:top]:
            st = ", ".join(f"{k}:{100.0 * v / max(d['samples'], 1):.0f}%" for k, v in sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3])
            print(f"{fn:>18s}:{ln:<4d} samples {100.0 * d['samples'] / max(tot_s, 1):5.1f}%  inst {100.0 * d['inst'] / max(tot_i, 1):5.1f}%  [{st}]  {src(fn, ln)}")
        print()


if __name__ == "__main__":
    main()
