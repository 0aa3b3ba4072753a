"""N>1 path on CPU: two gloo ranks each run their slice of a query batch (through the oracle here — there is no GPU in
this container; on the GPU box the same plumbing carries the CUDA results) and rank 0 gathers the top-k. The gathered
result must equal the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as ol
    from typesense_b200 import shard, structs as S, synth
    n_docs = 3000
    fd = synth.make_string_field(n_docs, 200, 3, 9, seed=5)
    pts = synth.make_points(n_docs, 6, hi=100)
    oi = ol.OracleIndex(n_docs, [fd.flat], [pts])
    toks = synth.sample_queries(fd, 37, 2, 9)          # 37: uneven split
    def make(lo, hi):
        qs = [S.Query([S.Combo([[int(t)] for t in toks[i]], 2)], topk=50,
                      sort=((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))) for i in range(lo, hi)]
        return S.KwBatch(qs, [0])
    lo, hi = shard.shard_range(len(toks), world, rank)
    kv, cnt, found = oi.keyword_search(make(lo, hi), 64)
    got = shard.gather_topk(kv, cnt, found, len(toks))
    if rank == 0:
        fkv, fcnt, ffound = oi.keyword_search(make(0, len(toks)), 64)
        ok = (got[1] == fcnt).all() and (got[2] == ffound).all()
        for q in range(len(toks)):
            ok = ok and (got[0]["key"][q, :fcnt[q]] == fkv["key"][q, :fcnt[q]]).all() and (got[0]["scores"][q, :fcnt[q]] == fkv["scores"][q, :fcnt[q]]).all()
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    sys.path.insert(0, ROOT)
    from typesense_b200 import shard
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, w, i) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
