#!/usr/bin/env python
"""Per-source-line stall samples: joins `ncu --page source --csv` (SASS view) with `nvdisasm -g` line info.
usage: tools/ncu_lines.py REP.ncu-rep LIB.so KERNEL_SUBSTRING [top_n] [SECTION_SUBSTRING]
SECTION_SUBSTRING picks the template instance in the cubin (mangled, e.g. hnsw_search_kernelILi6E); default = KERNEL."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sect = sys.argv[5] if len(sys.argv) > 5 else kern
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    inst = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    base = int(inst[0][0], 16)
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    sass, start = None, None
    for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):          # one cubin per translation unit
        cand = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()
        hit = [i for i, l in enumerate(cand) if l.startswith(".text.") and sect in l]
        if hit:
            sass, start = cand, hit[0]
            break
    if sass is None:
        sys.exit(f"no .text section matching {sect} in {so}")
    off2line = {}
    cur = ("?", 0)
    for l in sass[start + 1:]:
        if l.startswith("\t.section") or l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])
    tot_s = tot_i = 0.0
    for r in inst:
        off = int(r[0], 16) - base
        key = off2line.get(off, ("?", 0))
        s = float(r[ix["# Samples"]] or 0)
        n = float(r[ix["Instructions Executed"]] or 0)
        t = float(r[ix["Thread Instructions Executed"]] or 0)
        agg[key][0] += s; agg[key][1] += n; agg[key][2] += t
        tot_s += s; tot_i += n
    src_cache = {}

    def src(f, ln):
        for root in ("typesense_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                L = src_cache[p]
                return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
        return ""
    print(f"kernel {kern}: {tot_s:.0f} samples, {tot_i:.3g} warp instructions")
    print(f"{'samples%':>8} {'instr%':>7} {'thr/inst':>8}  location")
    for (f, ln), (s, n, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*s/tot_s:8.2f} {100*n/tot_i:7.2f} {t/max(n,1):8.1f}  {f}:{ln}  {src(f, ln)}")


if __name__ == "__main__":
    main()
