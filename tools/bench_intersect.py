#!/usr/bin/env python
"""BASELINE configs[0] / SURVEY §8(d) cfg 1: the reference's own DISABLED_BenchmarkIntersection shape
(test/posting_list_test.cpp:1452-1549): three posting lists of ~100 K / 50 K / 25 K ids drawn from [0, 1 M), block size
1024, 3-way AND. Times, on this machine:
  reference  posting_list_t::intersect          (oracle/_ref: the reference's sources compiled in place), single thread
  reference  sorted_array uncompress + ArrayUtils::and_scalar chain (the test's second measurement)
  port       oracle/ts_oracle.cpp tso_intersect
  tsgpu      tsgpu_intersect through the C-ABI (host lists in, host ids out) — only when a CUDA device is present
Prints one JSON object. Not part of bench.py (whose workload is the metric's hybrid10m configuration)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol                      # noqa: E402
from typesense_b200 import structs as S      # noqa: E402


def best_of(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def run(cpu_reps=20):
    rng = np.random.default_rng(42)
    n_range = 1_000_000
    lists = [np.unique(rng.integers(0, n_range, n)).astype(np.uint32) for n in (100_000, 50_000, 25_000)]
    algo_bytes = 4 * sum(len(l) for l in lists)
    out = {"config": {"workload": "intersect3", "ids": [len(l) for l in lists], "range": n_range, "block": 1024, "offsets": [0, 1, 3]},
           "cores": 1}
    ol.build_oracle()
    L = ol.oracle()
    ptrs = (S.u32p * 3)(*[ol.p32(a) for a in lists])
    lens = (C.c_size_t * 3)(*[len(a) for a in lists])
    res = np.zeros(len(lists[2]), np.uint32)
    t, n = best_of(lambda: L.tso_intersect(3, ptrs, lens, ol.p32(res), len(res)), cpu_reps)
    expect = res[:n].copy()
    out["port_us"] = 1e6 * t
    out["result_len"] = int(n)
    if ol.have_ref():
        R = ol.ref()
        flat = S.FlatField.from_postings([[(int(i), [0, 1, 3]) for i in l] for l in lists])
        pls = ol.ref_plists_of(flat, [0, 1, 2], 1024)
        hs = (C.c_void_p * 3)(*[p.h for p in pls])
        t, n2 = best_of(lambda: R.ref_plist_intersect(hs, 3, ol.p32(res), len(res)), 20)
        assert n2 == n and (res[:n2] == expect).all()
        out["reference_posting_list_us"] = 1e6 * t
        buf = np.zeros(len(lists[1]), np.uint32)
        buf2 = np.zeros(len(lists[2]), np.uint32)

        def chain():
            m = R.ref_and_scalar(ol.p32(lists[0]), len(lists[0]), ol.p32(lists[1]), len(lists[1]), ol.p32(buf))
            return R.ref_and_scalar(ol.p32(buf), m, ol.p32(lists[2]), len(lists[2]), ol.p32(buf2))
        t, n3 = best_of(chain, 20)
        assert n3 == n
        out["reference_and_scalar_chain_us"] = 1e6 * t
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        from typesense_b200 import capi
        flat = S.FlatField.from_postings([[(int(i), [0, 1, 3]) for i in l] for l in lists])
        gi = capi.GpuIndex(n_range, 0)
        gi.load_field(flat)
        got = gi.intersect(0, [0, 1, 2], len(lists[2]))
        assert got.tolist() == expect.tolist()
        t, _ = best_of(lambda: gi.intersect(0, [0, 1, 2], len(lists[2])), 50)
        st = gi.stats()
        out["tsgpu_call_us"] = 1e6 * t
        out["tsgpu_device_ms"] = st["ms_total"]
        out["tsgpu_algorithmic_GBps"] = algo_bytes / t / 1e9
        gi.close()
    return out


def main():
    print(json.dumps(run()))


if __name__ == "__main__":
    main()
