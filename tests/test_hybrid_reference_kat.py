"""Rank-fusion expectations of the reference's own hybrid tests, replayed on the oracle (CPU; the GPU path is compared with the oracle on
random hybrid batches in test_gpu_parity.py). The reference embeds text with a model that is not available offline; only the ORDER of the vector distances
matters to these assertions, so vectors with that order stand in. What is pinned: text ranks (ties share a rank), vector
ranks, the 0.7 / 0.3 weights, the float arithmetic of the fused score and the final order.
  test/collection_test.cpp:4779-4850            HybridSearchRankFusionTest
  test/collection_vector_search_test.cpp:5674-5753   TestRankFusionOrdering"""
import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import structs as S

SORT = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_SEQ_ID, -1, 1, 0), (S.SORT_NONE, -1, 1, 0))   # no default_sorting_field
FLAGS = S.FLAG_PRIORITIZE_EXACT_MATCH | S.FLAG_PRIORITIZE_NUM_MATCHING_FIELDS


def unit(v):
    v = np.asarray(v, np.float32)
    return v / np.linalg.norm(v)


def fused(kv, cnt):
    """[(seq_id, rank_fusion_score as float)] of query 0 in result order."""
    out = []
    for i in range(int(cnt[0])):
        msi = int(kv["match_score_index"][0, i])
        bits = np.int64(kv["scores"][0, i][msi]).astype(np.int32)
        out.append((int(kv["key"][0, i]), float(bits.view(np.float32))))
    return out


def field_of(docs):
    """plain string field: docs = list of token lists"""
    vocab, per_tok = {}, []
    for sid, toks in enumerate(docs):
        t2o = {}
        for i, t in enumerate(toks):
            t2o.setdefault(t, []).append(i + 1)
        t2o[toks[-1]].append(0)
        for t, offs in t2o.items():
            if t not in vocab:
                vocab[t] = len(per_tok)
                per_tok.append([])
            per_tok[vocab[t]].append((sid, offs))
    return vocab, S.FlatField.from_postings(per_tok)


def cases():
    q = unit([1, 0.2, 0, 0])
    # HybridSearchRankFusionTest: "butter" with prefix search -> candidates butter (cost 0), butterfly / butterball (prefix,
    # cost 1: next_suggestion2 adds 1 for a prefix-found candidate). Vector order butter < butterball < butterfly.
    vocab, flat = field_of([["butter"], ["butterball"], ["butterfly"]])
    vecs = np.stack([q, unit([1, 0.5, 0, 0]), unit([1, 1.5, 0, 0])])
    combos = [S.Combo([[vocab["butter"]]], 1, total_cost=0), S.Combo([[vocab["butterfly"]]], 1, total_cost=1),
              S.Combo([[vocab["butterball"]]], 1, total_cost=1)]
    yield ("HybridSearchRankFusionTest", flat, vecs, q, combos,
           [(0, 1.0 / 1.0 * 0.7 + 1.0 / 1.0 * 0.3), (1, 1.0 / 2.0 * 0.7 + 1.0 / 2.0 * 0.3), (2, 1.0 / 2.0 * 0.7 + 1.0 / 3.0 * 0.3)])
    # TestRankFusionOrdering: "apple" matches all three with the same text score (one shared text rank); vector order
    # green apple < apple pie < red apple
    vocab, flat = field_of([["red", "apple"], ["green", "apple"], ["apple", "pie"]])
    vecs = np.stack([unit([1, 2.0, 0, 0]), q, unit([1, 0.8, 0, 0])])
    yield ("TestRankFusionOrdering", flat, vecs, q, [S.Combo([[vocab["apple"]]], 1, total_cost=0)],
           [(1, 0.7 + 0.3 * 1.0 / 1.0), (2, 0.7 + 0.3 * 1.0 / 2.0), (0, 0.7 + 0.3 * 1.0 / 3.0)])


def check(run):
    for name, flat, vecs, q, combos, expect in cases():
        graph = ol.hnsw_build(vecs, 16, 200, 100)
        b = S.KwBatch([S.Query(combos, topk=250, sort=SORT, flags=FLAGS, num_query_tokens=1, field_weight=[15])], [0])
        kv, cnt, found = run(flat, graph, b, q[None, :].copy(), S.vec_params(k=0, ef=10, alpha=0.3, fetch_size=10))
        got = fused(kv, cnt)
        assert [k for k, _ in got] == [k for k, _ in expect], name
        for (_, g), (_, e) in zip(got, expect):
            assert g == pytest.approx(np.float32(e), rel=3e-7), name          # ASSERT_FLOAT_EQ in the reference
        assert int(found[0]) == 3, name


def test_rank_fusion_reference_kats_oracle():
    def run(flat, graph, b, qv, vp):
        oi = ol.OracleIndex(3, [flat], [], graph)
        return oi.hybrid_search(b, qv, vp, 256)
    check(run)


def test_distance_threshold_reference_kat_oracle():
    """DistanceThresholdTest, test/collection_vector_search_test.cpp:1548-1598: cosine (the default vec_dist), wildcard
    query + vector query; both documents without a threshold (nearest first), only the near one with
    distance_threshold:0.01."""
    docs = np.asarray([[0.1, 0.2, 0.3], [0.6, 0.7, 0.8]], np.float32)
    vecs = np.stack([unit(v) for v in docs])                      # hnsw_index_t::normalize_vector at index time
    q = unit([0.3, 0.4, 0.5])
    graph = ol.hnsw_build(vecs, 16, 200, 100, metric=1)
    oi = ol.OracleIndex(2, [], [], graph)
    sort = ((S.SORT_VECTOR_DISTANCE, -1, -1, 0), (S.SORT_SEQ_ID, -1, 1, 0), (S.SORT_NONE, -1, 1, 0))
    b = S.KwBatch([S.Query([], topk=250, sort=sort)], [])
    kv, cnt, found = oi.vector_search(b, q[None, :].copy(), S.vec_params(k=0, ef=10, fetch_size=20), 256)
    assert int(found[0]) == 2 and [int(kv["key"][0, i]) for i in range(int(cnt[0]))] == [1, 0]
    kv, cnt, found = oi.vector_search(b, q[None, :].copy(), S.vec_params(k=0, ef=10, fetch_size=20, distance_threshold=0.01), 256)
    assert int(found[0]) == 1 and [int(kv["key"][0, i]) for i in range(int(cnt[0]))] == [1]


def test_hybrid_flat_cutoff_reference_kat_oracle():
    """HybridSearchWithFilteringAndFlatSearchCutoff, test/collection_vector_search_test.cpp:5199-5263: the query word matches
    nothing, the filter (age:>0) holds all four documents, flat_search_cutoff:100 sends the vector part down the brute-force
    path — all four documents come back. (Vectors are stand-ins; only the control flow is asserted, as in the reference.)"""
    names = [["nike", "running", "shoes", "for", "men"], ["nike", "running", "sneakers"], ["adidas", "shoes"], ["puma"]]
    vocab, flat = field_of(names)
    rng = np.random.default_rng(5)
    vecs = np.stack([unit(rng.normal(size=8)) for _ in range(4)])
    graph = ol.hnsw_build(vecs, 16, 200, 100)
    oi = ol.OracleIndex(4, [flat], [], graph)
    q = S.Query([], topk=250, sort=SORT, flags=FLAGS, filter=0)        # "footwear": no candidate, so no combination to run
    b = S.KwBatch([q], [0], [np.arange(4, dtype=np.uint32)])
    kv, cnt, found = oi.hybrid_search(b, unit(rng.normal(size=8))[None, :].copy(), S.vec_params(k=0, ef=10, alpha=0.3, flat_search_cutoff=100, fetch_size=10), 256)
    assert int(cnt[0]) == 4 and int(found[0]) == 4 and sorted(int(kv["key"][0, i]) for i in range(4)) == [0, 1, 2, 3]


def test_rerank_hybrid_matches_scores_follow_the_two_rankings():
    """Index::compute_aux_scores (src/index.cpp:8793-8923; no reference test sets rerank_hybrid_matches): after it every result carries a text match
    score and a distance, and its score is 1/keyword_rank * (1 - alpha) + 1/semantic_rank * alpha with the keyword ranking by (text_match_score, key)
    descending and the semantic one by distance, stable on the keyword order. Recomputed here from the returned records."""
    import struct
    from typesense_b200 import synth
    n, dim = 3000, 16
    fd = synth.make_string_field(n, 60, 3, 8, seed=2)
    vec = synth.make_vectors(n, dim, 4).numpy()
    g = ol.hnsw_build(vec, 8, 40, 100)
    oi = ol.OracleIndex(n, [fd.flat], [], g)
    toks = synth.sample_queries(fd, 12, 2, 6)
    qs = []
    for row in toks:
        q = S.Query([S.Combo([[int(t)] for t in row], 2)], topk=40, sort=SORT, num_query_tokens=2)
        q.flags = FLAGS | S.FLAG_RERANK_HYBRID_MATCHES
        qs.append(q)
    qv = synth.make_vectors(12, dim, 8).numpy()
    alpha = 0.3
    kv, cnt, found = oi.hybrid_search(S.KwBatch(qs, [0]), qv, S.vec_params(k=20, ef=30, alpha=alpha, fetch_size=10), 64)

    def f2i(x):
        i = struct.unpack("<i", struct.pack("<f", np.float32(x)))[0]
        return i ^ 0x7FFFFFFF if i < 0 else i
    checked = 0
    for q in range(len(qs)):
        c = int(cnt[q])
        if c < 3:
            continue
        recs = [(int(kv["text_match_score"][q, i]), int(kv["key"][q, i]), float(kv["vector_distance"][q, i]), int(kv["scores"][q, i, 0])) for i in range(c)]
        assert all(r[2] != -1.0 for r in recs) and any(r[0] != 0 for r in recs)
        kw = sorted(recs, key=lambda r: (-r[0], -r[1]))
        krank = {r[1]: i + 1 for i, r in enumerate(kw)}
        sem = sorted(kw, key=lambda r: r[2])                       # Python's sort is stable, like std::stable_sort
        srank = {r[1]: i + 1 for i, r in enumerate(sem)}
        for r in recs:
            a = float(np.float32(alpha))                           # vector_query.alpha is a float; the expression is evaluated in double
            expect = f2i(np.float32((1.0 / krank[r[1]]) * (1.0 - a) + (1.0 / srank[r[1]]) * a))
            assert r[3] == expect, (q, r, krank[r[1]], srank[r[1]])
            checked += 1
        assert [r[3] for r in recs] == sorted([r[3] for r in recs], reverse=True)
    assert checked > 50
