#!/usr/bin/env python3
"""GPU timing of tsgpu_art_walk_batch on the vocabulary of tools/bench_art_cpu.py: a batch of 4096 x (prefix, 1-typo,
2-typo) searches, wall clock around the C-ABI call (copies included), checked against the host walk on a sample.
The walk mode is read once per process:   python tools/bench_art_gpu.py            (one thread per search)
                                           TSGPU_ART_MODE=frontier python tools/bench_art_gpu.py   (one thread per node visit)"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol          # noqa: E402
import test_art_mirror as T     # noqa: E402
import test_zz_art_gpu as Z     # noqa: E402
from typesense_b200 import capi  # noqa: E402


def main():
    n_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    n_queries = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    rng = np.random.default_rng(1)
    alpha = np.array(list("abcdefghijklmnopqrstuvwxyz"))
    lens = rng.integers(3, 11, n_tokens * 2)
    toks = sorted({"".join(rng.choice(alpha, int(l))) for l in lens})[:n_tokens]
    n = len(toks)
    if not os.path.exists(T.SO):
        import subprocess
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", os.path.join(ROOT, "tests", "cpp", "art_mirror_capi.cpp"), "-o", T.SO])
    L = C.CDLL(T.SO)
    vp = C.c_void_p
    L.am_build.restype = vp
    L.am_build.argtypes = [C.c_char_p, C.POINTER(C.c_int64), ol.u32p, C.c_uint32]
    L.am_walk.restype = C.c_size_t
    L.am_walk.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_size_t, C.POINTER(C.c_int)]
    ms = np.zeros(n, np.int64)
    df = np.ones(n, np.uint32)
    h = L.am_build("\n".join(toks).encode(), ms.ctypes.data_as(C.POINTER(C.c_int64)), ol.p32(df), n)
    root, arrs = Z.flat_arrays(L, h)
    gi = capi.GpuIndex(16, 0)
    gi.load_art(0, root, *arrs)
    res = {"tokens": n, "searches_per_kind": n_queries, "mode": os.environ.get("TSGPU_ART_MODE", "dfs"), "kinds": {}}
    hits = np.zeros(1 << 16, np.int32)
    so = C.c_int(0)
    for kind, cost, pre in (("prefix0", 0, 1), ("typo1", 1, 0), ("typo2", 2, 0)):
        terms = []
        for _ in range(n_queries):
            w = toks[int(rng.integers(0, n))]
            if kind == "prefix0":
                w = w[:max(2, len(w) // 2)]
            else:
                i = int(rng.integers(0, len(w)))
                w = w[:i] + str(rng.choice(alpha)) + w[i + 1:]
            terms.append(w.encode())
        cap = 2048
        gi.art_walk(0, terms[:64], [cost] * 64, [cost] * 64, [pre] * 64, cap)          # warm-up (buffers, first launch)
        t0 = time.perf_counter()
        got, flags = gi.art_walk(0, terms, [cost] * n_queries, [cost] * n_queries, [pre] * n_queries, cap)
        dt = time.perf_counter() - t0
        bad = 0
        for i in range(0, n_queries, max(1, n_queries // 64)):
            k = L.am_walk(h, 0, terms[i], cost, cost, pre, hits.ctypes.data_as(C.POINTER(C.c_int32)), len(hits), C.byref(so))
            want = hits[:k].tolist()
            bad += not ((flags[i] == 0 and got[i] == want) or (flags[i] == 4 and k > cap))
        res["kinds"][kind] = {"ms_per_batch": round(dt * 1e3, 2), "us_per_search": round(dt / n_queries * 1e6, 2), "flagged_for_host": int((flags != 0).sum()),
                              "sample_mismatches": int(bad)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
