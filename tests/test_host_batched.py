"""The C++ host layer's batched multi_search (tsgpu_host.hpp: multi_search_batched — replay passes, lazily batched candidate
walks, hybrid tail through tsgpu_hybrid_fuse_batch) end to end from query STRINGS, against the oracle's one-call hybrid search
on the RESOLVED form of the same queries. What has to line up: the ART mirror built from the vocabulary, the cost-0 / typo
control flow of fuzzy_search_fields (a misspelt token has no cost-0 candidate, is struck off, and comes back corrected at
cost 1 -> total_cost 2), persistent filters on both sides, the keyword rounds, and the rank fusion tail.
CPU: the host layer on the oracle-backed double. GPU: the same host layer on libtsgpu.so (device walks, device rounds)."""
import numpy as np
import pytest

import hostlib
import oracle_lib as ol
from typesense_b200 import hostapi, structs as S, synth


def collection(n_docs=6000, vocab=400, dim=32):
    fd = synth.make_string_field(n_docs, vocab, 4, 10, seed=11)
    pts = synth.make_points(n_docs, 5)
    vec = synth.make_vectors_clustered(n_docs, dim, 12, seed=3, spread=0.5, latent=6, center_latent=6)[0].numpy()
    g = ol.hnsw_build(vec, 16, 80, 100)
    rng = np.random.default_rng(2)
    cat = rng.integers(0, 4, n_docs)
    filters = [np.nonzero(cat == c)[0].astype(np.uint32) for c in range(4)]
    return fd, pts, vec, g, filters, synth.vocab_words(vocab)


def queries(fd, words, n, seed, typo_frac=0.3):
    rng = np.random.default_rng(seed)
    toks = synth.sample_queries(fd, n, 3, seed)
    taken = set(words)
    qs, fixed_cost = [], []
    for row in toks:
        q = [words[int(t)] for t in row]
        cost = 0
        if rng.random() < typo_frac:
            j = int(rng.integers(0, 3))
            q[j] = synth.misspell(q[j], rng, taken)
            cost = 2                                             # next_suggestion2: 2 * typo cost (+1 only for a prefix-extended candidate)
        qs.append(q); fixed_cost.append(cost)
    return toks, qs, fixed_cost


def run_and_check(lib_path, device_walk):
    fd, pts, vec, g, filters, words = collection()
    n_docs = len(pts)
    hi = hostapi.HostIndex(n_docs, 0, lib_path)
    hi.add_field_flat("title", words, fd.flat)
    hi.add_sort_column("points", pts)
    handles = [hi.add_filter(f) for f in filters]
    hi.device_index().load_hnsw(g)
    oi = ol.OracleIndex(n_docs, [fd.flat], [pts], g)
    toks, qs, cost = queries(fd, words, 48, 9)
    rng = np.random.default_rng(4)
    qf = np.asarray([int(rng.integers(0, 4)) if i % 2 else -1 for i in range(len(qs))])
    qv = synth.make_vectors_clustered(len(qs), vec.shape[1], 12, seed=77, spread=0.5, latent=6, center_latent=6, centers_seed=3)[0].numpy()
    opt = hostapi.Options(device_art_walk=1 if device_walk else 0, n_threads=3, vec_fetch_size=100)
    kv, cnt, found, st = hi.multi_search("title", "points", qs, 250, np.asarray([handles[f] if f >= 0 else -1 for f in qf], np.int32), qv, opt)
    st_first = dict(st)
    assert st["fuse_queries"] == len(qs) and st["kw_queries"] >= len(qs)
    # the resolved form: one combination of the (corrected) tokens per query
    sort = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
    oq = []
    for i, row in enumerate(toks):
        q = S.Query([S.Combo([[int(t)] for t in row], 3, total_cost=cost[i])], topk=250, sort=sort, num_query_tokens=3)
        if qf[i] >= 0:
            q.filter = int(qf[i])
        oq.append(q)
    okv, ocnt, ofound = oi.hybrid_search(S.KwBatch(oq, [0], filters), qv, S.vec_params(k=0, ef=10, alpha=0.3, fetch_size=100), 250, 1)
    # a query whose tokens have no common document inside its filter goes on to typo candidates and token dropping (as in the
    # reference): the single resolved combination describes the flow only where it matches
    _, _, kw_only_found = oi.keyword_search(S.KwBatch(oq, [0], filters), 250)
    compared = 0
    for i in range(len(qs)):
        if kw_only_found[i] == 0:
            continue
        compared += 1
        assert cnt[i] == ocnt[i] and found[i] == ofound[i], (i, cnt[i], ocnt[i], found[i], ofound[i])
        assert kv["key"][i, :cnt[i]].tolist() == okv["key"][i, :ocnt[i]].tolist(), i
        assert kv["scores"][i, :cnt[i]].tolist() == okv["scores"][i, :ocnt[i]].tolist(), i
    assert compared >= len(qs) // 2
    # keyword only, same flow
    kv, cnt, found, st = hi.multi_search("title", "points", qs, 250, None, None, opt)
    kq = [S.Query(q.combos, topk=250, sort=sort, num_query_tokens=3) for q in oq]
    okv, ocnt, ofound = oi.keyword_search(S.KwBatch(kq, [0]), 250)
    for i in range(len(qs)):
        assert cnt[i] == ocnt[i] and found[i] == ofound[i] and kv["key"][i, :cnt[i]].tolist() == okv["key"][i, :ocnt[i]].tolist(), i
    hi.close()
    return st_first


def test_batched_multi_search_on_the_oracle_double():
    st = run_and_check(hostlib.build_host_cpu(), device_walk=False)
    assert st["passes"] >= 2


@pytest.mark.gpu
def test_batched_multi_search_on_the_gpu():
    import os
    if os.environ.get("TSGPU_TEST_DOUBLE") == "1":
        pytest.skip("links the real libtsgpu.so")
    st = run_and_check(hostapi.build_gpu_lib(), device_walk=True)
    assert st["walk_batches"] >= 1 and st["walks"] >= 5
