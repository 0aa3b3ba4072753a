#!/usr/bin/env python3
"""Per-kernel SASS summary of a built library: instruction count, local-memory loads/stores (LDL/STL), global loads, and a
hash of the instruction stream (addresses stripped) — to show that a change left a kernel's code untouched, or what it did
to its stack traffic, without a GPU.   usage: tools/sass_funcs.py lib.so [substring]"""
import hashlib
import re
import subprocess
import sys


def funcs(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append(re.sub(r"\s+", " ", m.group(1)))
    return out


if __name__ == "__main__":
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, ins in funcs(sys.argv[1]).items():
        if sub not in name:
            continue
        cnt = lambda p: sum(1 for i in ins if re.search(p, i))
        print("%-90s n=%6d LDL=%4d STL=%4d LDG=%4d sha=%s" % (name[:90], len(ins), cnt(r"\bLDL"), cnt(r"\bSTL"), cnt(r"\bLDG"),
                                                              hashlib.sha1("\n".join(ins).encode()).hexdigest()[:12]))
