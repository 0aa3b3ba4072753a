#!/usr/bin/env python
"""Device HNSW build at scale: time, counters, recall@k of the built graph against brute force.
usage: tools/bench_build_gpu.py [n] [dim] [max_batch] [efc]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from typesense_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
efc = int(sys.argv[4]) if len(sys.argv) > 4 else 200
dev = "cuda:0"
ncl = max(8, n // 2000)
vec, _ = synth.make_vectors_clustered(n, dim, ncl, seed=1234, device=dev, spread=0.35)
torch.cuda.synchronize()
gi = capi.GpuIndex(n, 0)
t0 = time.time()
info = gi.build_hnsw(vec, 16, efc, 100, max_batch=max_batch, keep_device_vectors=True)
dt = time.time() - t0
R = 256
q = synth.make_vectors_clustered(R, dim, ncl, seed=555, device=dev, centers_seed=1234, spread=0.35)[0]
best = torch.full((R, 100), -2.0, device=dev)
ex = torch.zeros(R, 100, dtype=torch.int64, device=dev)
step = 1 << 20
for s0 in range(0, n, step):
    sims = q @ vec[s0:s0 + step].T
    cs = torch.cat([best, sims], 1)
    ci = torch.cat([ex, torch.arange(s0, min(n, s0 + step), device=dev)[None, :].expand(R, -1)], 1)
    best, pos = torch.topk(cs, 100, dim=1)
    ex = torch.gather(ci, 1, pos)
ex = ex.cpu().numpy()
qn = q.cpu().numpy()
out = {"n": n, "dim": dim, "max_batch": max_batch, "efc": efc, "build_s": dt, "inserts_per_s": n / dt, "info": info}
for ef in (100, 200):
    t1 = time.time()
    d, l, c = gi.knn(qn, 100, ef)
    out[f"recall_at_100_ef{ef}"] = float(np.mean([len(set(l[i][:c[i]].tolist()) & set(ex[i].tolist())) / 100 for i in range(R)]))
    st = gi.stats()
    out[f"ef{ef}_dist_per_query"] = st["knn_dist"] / R
    out[f"ef{ef}_exp_per_query"] = st["knn_expanded"] / R
    out[f"ef{ef}_ms"] = st["ms_knn"]
print(json.dumps(out))
