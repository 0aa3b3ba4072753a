"""GPU parity tests (run with -m gpu on the B200 box): every call goes through the C-ABI of libtsgpu.so and is
compared with the CPU oracle on the same seeded inputs — bit-exact for ids / integer scores, 1e-4 relative for float
distances (they are in fact bit-equal because both sides use the same summation order)."""
import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import capi, structs as S, synth
from test_oracle_ref import random_batch, small_collection  # noqa: F401

pytestmark = pytest.mark.gpu


def assert_kv_equal(kv, cnt, found, okv, ocnt, ofound, check_query_index=True, vd_tol=None):
    assert cnt.tolist() == ocnt.tolist()
    assert found.tolist() == ofound.tolist()
    for q in range(len(cnt)):
        n = int(cnt[q])
        assert kv["key"][q, :n].tolist() == okv["key"][q, :n].tolist(), f"query {q}: ids"
        assert kv["scores"][q, :n].tolist() == okv["scores"][q, :n].tolist(), f"query {q}: scores"
        assert kv["text_match_score"][q, :n].tolist() == okv["text_match_score"][q, :n].tolist(), f"query {q}: text score"
        assert kv["match_score_index"][q, :n].tolist() == okv["match_score_index"][q, :n].tolist(), f"query {q}: msi"
        if check_query_index:
            assert kv["query_index"][q, :n].tolist() == okv["query_index"][q, :n].tolist(), f"query {q}: query_index"
        if vd_tol is None:
            assert kv["vector_distance"][q, :n].tolist() == okv["vector_distance"][q, :n].tolist(), f"query {q}: vdist"
        else:
            assert np.allclose(kv["vector_distance"][q, :n], okv["vector_distance"][q, :n], rtol=vd_tol, atol=1e-6)


@pytest.fixture(scope="module")
def coll(small_collection):
    n_docs, fds, pts = small_collection
    flats = [fd.flat for fd in fds]
    gi = capi.GpuIndex(n_docs, 0)
    for f in flats:
        gi.load_field(f)
    gi.load_sort_column(pts)
    oi = ol.OracleIndex(n_docs, flats, [pts])
    yield n_docs, fds, flats, pts, gi, oi
    gi.close()


@pytest.mark.parametrize("seed", range(4))
def test_keyword_search_random(coll, seed):
    n_docs, fds, flats, pts, gi, oi = coll
    rng = np.random.default_rng(900 + seed)
    filters = [np.unique(rng.integers(0, n_docs, 1500)).astype(np.uint32), np.arange(0, n_docs, 7, dtype=np.uint32),
               np.zeros(0, np.uint32)]
    b = random_batch(rng, fds, 60, filters)
    kv, cnt, found = gi.keyword_search(b, 256)
    okv, ocnt, ofound = oi.keyword_search(b, 256)
    assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)
    assert int(cnt.sum()) > 500


def test_keyword_persistent_filter_and_big_k(coll):
    n_docs, fds, flats, pts, gi, oi = coll
    rng = np.random.default_rng(5)
    fil = np.unique(rng.integers(0, n_docs, 2500)).astype(np.uint32)
    h = gi.filter_create(fil)
    toks = synth.sample_queries(fds[0], 30, 2, 3)
    qs, qso = [], []
    for i, row in enumerate(toks):
        combos = [S.Combo([[int(t), S.NO_LIST, int(t) if fds[2].flat.df(int(t)) else S.NO_LIST] for t in row], 2)]
        kw = dict(topk=int(rng.choice([3, 100, 600, 1024])), sort=((S.SORT_NUMERIC, 0, -1, 1), (S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_SEQ_ID, -1, 1, 0)))
        qs.append(S.Query(combos, filter=h if i % 2 else -1, **kw))
        qso.append(S.Query(combos, filter=0 if i % 2 else -1, **kw))
    b = S.KwBatch(qs, [0, 1, 2])
    bo = S.KwBatch(qso, [0, 1, 2], [fil])
    kv, cnt, found = gi.keyword_search(b, 1024)
    okv, ocnt, ofound = oi.keyword_search(bo, 1024)
    assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)


def test_keyword_single_token_large_lists():
    # one frequent token: every posting matches; exercises the streaming top-K with threshold pruning across units
    n_docs = 300000
    fd = synth.make_string_field(n_docs, 50, 2, 6, seed=21)
    pts = synth.make_points(n_docs, 4, hi=1000)
    gi = capi.GpuIndex(n_docs, 0)
    gi.load_field(fd.flat)
    gi.load_sort_column(pts)
    oi = ol.OracleIndex(n_docs, [fd.flat], [pts])
    qs = [S.Query([S.Combo([[t]], 1)], topk=k, sort=((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0)))
          for t, k in [(0, 250), (1, 10), (3, 1000), (7, 1)]]
    qs.append(S.Query([S.Combo([[0], [1]], 2), S.Combo([[0], [2]], 2, total_cost=1), S.Combo([[1], [2]], 2, total_cost=2)], topk=250, num_query_tokens=2))
    b = S.KwBatch(qs, [0])
    kv, cnt, found = gi.keyword_search(b, 1024)
    okv, ocnt, ofound = oi.keyword_search(b, 1024)
    assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)
    st = gi.stats()
    assert st["kw_matches"] > 100000 and st["launches_total"] >= 2
    gi.close()


def test_intersect_and_phrase(coll):
    n_docs, fds, flats, pts, gi, oi = coll
    from test_oracle_ref import tso_intersect, KAT
    # reference golden vectors through the C-ABI
    for case in KAT["posting_intersect"]:
        flat = S.FlatField.from_postings([[(i, [1]) for i in l] for l in case["lists"]])
        g2 = capi.GpuIndex(64, 0)
        g2.load_field(flat)
        assert g2.intersect(0, list(range(len(case["lists"]))), 64).tolist() == case["expect"]
        g2.close()
    rng = np.random.default_rng(12)
    L = ol.oracle()
    for fi in (0, 1, 2):
        fl = flats[fi]
        for _ in range(15):
            k = int(rng.integers(1, 4))
            lists = rng.integers(0, 40, k).astype(np.uint32)
            if any(fl.df(int(l)) == 0 for l in lists):
                continue
            exp = tso_intersect([fl.ids[int(fl.list_off[l]):int(fl.list_off[l + 1])] for l in lists])
            got = gi.intersect(fi, lists.tolist(), n_docs)
            assert got.tolist() == exp
            if len(exp) and k > 1:
                ids = np.asarray(exp, np.uint32)
                out = np.zeros(len(ids), np.uint32)
                oix = ol.OracleIndex(n_docs, [fl], [])
                n = L.tso_phrase_matches(oix.h, 0, ol.p32(np.ascontiguousarray(lists)), k, ol.p32(ids), len(ids), ol.p32(out))
                assert gi.phrase_matches(fi, lists.tolist(), ids).tolist() == out[:n].tolist()


def test_contains_atleast_one(coll):
    """posting_t::contains_atleast_one: the reference's own cases (test/posting_list_test.cpp:823-859) on a loaded field, then
    random target sets against the oracle (pinned on the reference's compiled code in tests/test_oracle_ref.py)."""
    from test_oracle_ref import CONTAINS_KAT
    lists = [np.asarray(k[0], np.uint32) for k in CONTAINS_KAT[:1]] + [np.asarray(CONTAINS_KAT[3][0], np.uint32)]
    flat = S.FlatField.from_postings([[(int(d), [1, 0]) for d in l] for l in lists], False)
    gk = capi.GpuIndex(4000, 0)
    fid = gk.load_field(flat)
    for lst, targets, expect in CONTAINS_KAT:
        li = 0 if len(lst) > 100 else 1
        assert gk.contains_atleast_one(fid, li, targets) == expect
    gk.close()
    n_docs, fds, flats, pts, gi, oi = coll
    rng = np.random.default_rng(3)
    L = ol.oracle()
    for f, flat in enumerate(flats):
        df = np.diff(flat.list_off.astype(np.int64))
        for _ in range(60):
            l = int(rng.choice(np.nonzero(df > 0)[0]))
            tg = np.unique(rng.integers(0, n_docs, int(rng.integers(1, 50)))).astype(np.uint32)
            a = np.ascontiguousarray(flat.ids[int(flat.list_off[l]):int(flat.list_off[l + 1])])
            assert gi.contains_atleast_one(f, l, tg) == bool(L.tso_contains_atleast_one(ol.p32(a), len(a), ol.p32(tg), len(tg)))


def test_exact_and_prefix_matches(coll):
    """tsgpu_exact_matches / tsgpu_prefix_matches vs the oracle (itself pinned on the reference's compiled code)."""
    from test_hostsim import idset_cases
    n_docs, fds, flats, pts, gi, oi = coll
    L = ol.oracle()
    rng = np.random.default_rng(2024)
    hits = {"exact": 0, "prefix": 0, "phrase": 0}
    for fi in (0, 1, 2):
        oix = ol.OracleIndex(n_docs, [flats[fi]], [])
        for lists, ids in idset_cases(rng, fds[fi], 120):
            k = len(lists)
            for name, ofn, gfn in (("exact", L.tso_exact_matches, gi.exact_matches), ("prefix", L.tso_prefix_matches, gi.prefix_matches),
                                   ("phrase", L.tso_phrase_matches, gi.phrase_matches)):
                out = np.zeros(len(ids), np.uint32)
                n = ofn(oix.h, 0, ol.p32(lists), k, ol.p32(ids), len(ids), ol.p32(out))
                assert gfn(fi, lists.tolist(), ids).tolist() == out[:n].tolist(), (fi, name, lists.tolist())
                hits[name] += n
    assert min(hits.values()) > 20, hits


def test_ids_setop(coll):
    """tsgpu_ids_setop: the reference's array_utils_test.cpp vectors, then random strictly-ascending arrays vs the oracle."""
    from test_oracle_ref import KAT
    n_docs, fds, flats, pts, gi, oi = coll
    for case in KAT["array_utils"]:
        assert gi.ids_setop(case["op"], case["a"], case["b"]).tolist() == case["expect"], case["src"]
    L = ol.oracle()
    rng = np.random.default_rng(8)
    names = ("tso_and_scalar", "tso_or_scalar", "tso_exclude_scalar")
    for it in range(30):
        hi = int(rng.choice([40, 500, n_docs]))
        a = np.unique(rng.integers(0, hi, int(rng.integers(0, 3000)))).astype(np.uint32)
        b = np.unique(rng.integers(0, hi, int(rng.integers(0, 3000)))).astype(np.uint32)
        a0 = a if len(a) else np.zeros(1, np.uint32)
        b0 = b if len(b) else np.zeros(1, np.uint32)
        for op in range(3):
            out = np.zeros(len(a) + len(b) + 1, np.uint32)
            n = getattr(L, names[op])(ol.p32(a0), len(a), ol.p32(b0), len(b), ol.p32(out))
            assert gi.ids_setop(op, a, b).tolist() == out[:n].tolist(), (it, op)
    with pytest.raises(capi.TsgpuError):
        gi.ids_setop(0, [3, 3, 4], [3])                   # not strictly ascending
    with pytest.raises(capi.TsgpuError):
        gi.ids_setop(1, [1, n_docs], [3])                 # id out of range


@pytest.fixture(scope="module")
def vec_coll():
    n, dim = 6000, 128
    vec = synth.make_vectors(n, dim, 5).numpy()
    g = ol.hnsw_build(vec, 16, 100, 100)
    fd = synth.make_string_field(n, 200, 3, 8, seed=31)
    pts = synth.make_points(n, 8, hi=100)
    gi = capi.GpuIndex(n, 0)
    gi.load_field(fd.flat)
    gi.load_sort_column(pts)
    gi.load_hnsw(g)
    oi = ol.OracleIndex(n, [fd.flat], [pts], g)
    yield n, dim, vec, g, fd, pts, gi, oi
    gi.close()


def test_knn_matches_oracle(vec_coll):
    n, dim, vec, g, fd, pts, gi, oi = vec_coll
    qv = synth.make_vectors(200, dim, 77).numpy()
    for k, ef in [(10, 10), (100, 10), (5, 200)]:
        d, l, cnt = gi.knn(qv, k, ef)
        od, olab, ocnt, _ = oi.knn(qv, k, ef)
        assert cnt.tolist() == ocnt.tolist()
        assert l.tolist() == olab.tolist()
        assert np.allclose(d, od, rtol=1e-4, atol=1e-6)
        assert (d == od).all(), "same summation order => bit-equal distances"
    # filtered
    rng = np.random.default_rng(1)
    filters = [np.unique(rng.integers(0, n, 600)).astype(np.uint32), np.arange(0, n, 3, dtype=np.uint32)]
    qf = rng.integers(-1, 2, len(qv)).astype(np.int32)
    d, l, cnt = gi.knn(qv, 20, 40, qf, filters)
    od, olab, ocnt, _ = oi.knn(qv, 20, 40, qf, filters)
    assert cnt.tolist() == ocnt.tolist() and l.tolist() == olab.tolist() and (d == od).all()


def test_hnsw_load_rejects_malformed_graph_and_keeps_the_old_one(vec_coll):
    """tsgpu_index_load_hnsw validates what the walk kernels follow blindly (ids, counts, offsets) and swaps only on success."""
    import copy
    n, dim, vec, g, fd, pts, gi, oi = vec_coll
    qv = synth.make_vectors(16, dim, 5).numpy()
    before = gi.knn(qv, 10, 20)
    L0 = 2 * g.M + 1

    def broken(mut):
        b = copy.copy(g)
        for f in ("links0", "upper_off", "links_up", "levels"):
            setattr(b, f, getattr(g, f).copy())
        mut(b)
        return b

    def bad_id(b): b.links0[5 * L0 + 1] = n + 7
    def bad_count(b): b.links0[9 * L0] = 2 * g.M + 1
    def bad_levels(b): b.levels[int(np.argmax(g.levels == 0))] = 1
    def bad_entry(b): b.entry_point = n
    def bad_entry_level(b): b.entry_point = int(np.argmax(g.levels == 0))
    def bad_upper(b):
        assert len(b.links_up)
        b.links_up[1] = n

    for mut in (bad_id, bad_count, bad_levels, bad_entry, bad_entry_level, bad_upper):
        with pytest.raises(capi.TsgpuError):
            gi.load_hnsw(broken(mut))
        after = gi.knn(qv, 10, 20)
        assert all((a == b).all() for a, b in zip(before, after)), mut.__name__


def test_knn_generic_dim_and_reference_kat():
    from test_oracle_ref import KAT
    k = KAT["vector_cosine"]
    L = ol.oracle()
    docs = np.asarray(k["docs"], np.float32)
    q = np.asarray(k["query"], np.float32)
    nd = np.zeros_like(docs)
    for i in range(len(docs)):
        L.tso_normalize(docs[i].ctypes.data_as(S.f32p), nd[i].ctypes.data_as(S.f32p), 4)
    nq = np.zeros_like(q)
    L.tso_normalize(q.ctypes.data_as(S.f32p), nq.ctypes.data_as(S.f32p), 4)
    g = ol.hnsw_build(nd, 16, 200, 100, metric=1)
    gi = capi.GpuIndex(3, 0)
    gi.load_hnsw(g)
    d, l, n = gi.knn(nq[None, :], 10, 10)
    assert n[0] == 3 and l[0][:3].tolist() == k["expect_ids"]
    assert np.allclose(np.abs(d[0][:3]), k["expect_distances"], rtol=0, atol=2e-7)
    d, l, n = gi.knn(nq[None, :], 10, 10, np.asarray([0], np.int32), [np.asarray(k["filtered"]["filter_ids"], np.uint32)])
    assert l[0][:n[0]].tolist() == k["filtered"]["expect_ids"]
    gi.close()
    # odd dimension through the generic path
    vec = synth.make_vectors(1500, 50, 3).numpy()
    g = ol.hnsw_build(vec, 8, 60, 100)
    gi = capi.GpuIndex(1500, 0)
    gi.load_hnsw(g)
    oi = ol.OracleIndex(1500, [], [], g)
    qv = synth.make_vectors(40, 50, 4).numpy()
    d, l, n = gi.knn(qv, 7, 30)
    od, olab, on, _ = oi.knn(qv, 7, 30)
    assert l.tolist() == olab.tolist() and (d == od).all()
    ids = np.arange(0, 1500, 11, dtype=np.uint32)
    fd = gi.flat_distances(qv[0], ids)
    od = np.zeros(len(ids), np.float32)
    L.tso_flat_distances(oi.hs, qv[0].ctypes.data_as(S.f32p), ol.p32(ids), len(ids), od.ctypes.data_as(S.f32p))
    assert (fd == od).all()
    gi.close()


def test_knn_selective_filters_long_walks():
    """Selective filters make hnswlib visit most of the graph before `ef` allowed results exist (the functor gates only the
    result heap): the visited set leaves shared memory (tier 2) and, with a handful of allowed ids, outgrows the per-slot
    scratch — those walks are handed to the retry launch and must still return the oracle's answer, without disturbing the
    other queries of the batch (one long walk used to fail the whole call)."""
    import torch
    n, dim = 90000, 32
    vec = synth.make_vectors_clustered(n, dim, 45, seed=5, device="cuda", spread=0.5, latent=8, center_latent=8)[0]
    lv, l0, uo, lu, ml, ep = synth.build_graph_bulk(vec, 16, 100)
    g = S.HnswGraph(vec.cpu().numpy(), lv.cpu().numpy(), l0.cpu().numpy().astype(np.uint32), uo.cpu().numpy().astype(np.uint64),
                    lu.cpu().numpy().astype(np.uint32), 16, ml, ep)
    gi = capi.GpuIndex(n, 0)
    gi.load_hnsw(g)
    oi = ol.OracleIndex(n, [], [], g)
    rng = np.random.default_rng(8)
    qv = synth.make_vectors_clustered(48, dim, 45, seed=77, device="cpu", spread=0.5, latent=8, center_latent=8, centers_seed=5)[0].numpy()
    filters = [np.unique(rng.integers(0, n, 25)).astype(np.uint32),            # ~25 allowed ids: the walk covers the graph (retry)
               np.unique(rng.integers(0, n, n // 40)).astype(np.uint32),       # 2.5 %: thousands of visited nodes (tier 2)
               np.arange(0, n, 2, dtype=np.uint32)]
    qf = (np.arange(len(qv)) % 4 - 1).astype(np.int32)                         # -1 (none), 0, 1, 2
    d, l, cnt = gi.knn(qv, 10, 30, qf, filters)
    st = gi.stats()
    od, olab, ocnt, _ = oi.knn(qv, 10, 30, qf, filters)
    assert cnt.tolist() == ocnt.tolist() and l.tolist() == olab.tolist() and (d == od).all()
    assert st["knn_tier2_walks"] > 0 and st["knn_retried"] > 0, st
    gi.close()


def _vec_queries(rng, fd, n, filters, sort, with_combos):
    toks = synth.sample_queries(fd, n, 2, int(rng.integers(0, 1 << 30)))
    qs = []
    for i, row in enumerate(toks):
        combos = [S.Combo([[int(t)] for t in row], 2)] if with_combos else []
        if with_combos and i % 3 == 0:
            combos.append(S.Combo([[int(row[0])], [int(rng.integers(0, 30))]], 2, total_cost=1))
        q = S.Query(combos, topk=int(rng.choice([10, 250])), sort=sort, num_query_tokens=2)
        if filters and i % 2:
            q.filter = int(rng.integers(0, len(filters)))
        if i % 5 == 0:
            q.excl = np.unique(rng.integers(0, 6000, 30)).tolist()
        qs.append(q)
    return S.KwBatch(qs, [0], filters)


def test_vector_search_matches_oracle(vec_coll):
    n, dim, vec, g, fd, pts, gi, oi = vec_coll
    rng = np.random.default_rng(41)
    filters = [np.unique(rng.integers(0, n, 900)).astype(np.uint32), np.arange(0, n, 40, dtype=np.uint32)]
    sort = ((S.SORT_VECTOR_DISTANCE, -1, -1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
    b = _vec_queries(rng, fd, 60, filters, sort, False)
    qv = synth.make_vectors(60, dim, 43).numpy()
    for vp in [S.vec_params(k=0, ef=10, fetch_size=20), S.vec_params(k=50, ef=80, flat_search_cutoff=500, fetch_size=10),
               S.vec_params(k=30, ef=30, distance_threshold=0.9, fetch_size=10)]:
        kv, cnt, found = gi.vector_search(b, qv, vp, 256)
        okv, ocnt, ofound = oi.vector_search(b, qv, vp, 256)
        assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)


def test_hybrid_search_matches_oracle(vec_coll):
    n, dim, vec, g, fd, pts, gi, oi = vec_coll
    rng = np.random.default_rng(51)
    filters = [np.unique(rng.integers(0, n, 1200)).astype(np.uint32), np.arange(0, n, 50, dtype=np.uint32)]
    sort = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
    b = _vec_queries(rng, fd, 80, filters, sort, True)
    qv = synth.make_vectors(80, dim, 53).numpy()
    for vp in [S.vec_params(k=0, ef=10, alpha=0.3, fetch_size=10), S.vec_params(k=40, ef=60, alpha=0.7, flat_search_cutoff=300, fetch_size=10)]:
        kv, cnt, found = gi.hybrid_search(b, qv, vp, 256)
        okv, ocnt, ofound = oi.hybrid_search(b, qv, vp, 256)
        assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)
    # rerank_hybrid_matches: Index::compute_aux_scores after the fusion (vector-only results get a text score, keyword-only ones a
    # distance, everything is re-ranked and re-scored), on half of the queries, both metrics of field count
    b3 = _vec_queries(rng, fd, 60, filters, sort, True)
    for i in range(0, b3.n_queries, 2):
        b3.q_flags[i] |= 0x40
    for vp in [S.vec_params(k=0, ef=10, alpha=0.3, fetch_size=10), S.vec_params(k=30, ef=40, alpha=0.8, fetch_size=10)]:
        kv, cnt, found = gi.hybrid_search(b3, qv[:60], vp, 256)
        okv, ocnt, ofound = oi.hybrid_search(b3, qv[:60], vp, 256)
        assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)
    # sort clause with _vector_distance as tie-breaker and a full Topster (the add()-on-sorted-array path)
    sort2 = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_VECTOR_DISTANCE, -1, -1, 0), (S.SORT_SEQ_ID, -1, 1, 0))
    b2 = _vec_queries(rng, fd, 40, [], sort2, True)
    for i in range(b2.n_queries):
        b2.q_topk[i] = 10
    kv, cnt, found = gi.hybrid_search(b2, qv[:40], S.vec_params(k=25, ef=25, fetch_size=10), 256)
    okv, ocnt, ofound = oi.hybrid_search(b2, qv[:40], S.vec_params(k=25, ef=25, fetch_size=10), 256)
    assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)


def test_wildcard_search_matches_oracle(coll):
    # Index::search_wildcard: q=* over the filter ids / all docs, exclusion list, sort clauses incl. ASC + missing-first
    n_docs, fds, flats, pts, gi, oi = coll
    rng = np.random.default_rng(61)
    filters = [np.unique(rng.integers(0, n_docs, 1700)).astype(np.uint32), np.arange(3, n_docs, 11, dtype=np.uint32), np.zeros(0, np.uint32)]
    h = gi.filter_create(filters[0])
    qs, qso = [], []
    for i in range(24):
        sort = [((S.SORT_NUMERIC, 0, -1, 1), (S.SORT_SEQ_ID, -1, 1, 0), (S.SORT_NONE, -1, 1, 0)),
                ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_SEQ_ID, -1, -1, 0)),
                ((S.SORT_SEQ_ID, -1, 1, 0), (S.SORT_NONE, -1, 1, 0), (S.SORT_NONE, -1, 1, 0))][i % 3]
        kw = dict(topk=int(rng.choice([1, 10, 250, 1000])), sort=sort, excl=np.unique(rng.integers(0, n_docs, 25)).tolist() if i % 4 == 0 else ())
        f_inline = [-1, 0, 1, 2][i % 4]
        qs.append(S.Query([], filter=(h if f_inline == 0 and i % 8 == 1 else f_inline), **kw))
        qso.append(S.Query([], filter=f_inline, **kw))
    b, bo = S.KwBatch(qs, [0], filters), S.KwBatch(qso, [0], filters)
    kv, cnt, found = gi.wildcard_search(b, 1024)
    okv, ocnt, ofound = oi.wildcard_search(bo, 1024)
    assert_kv_equal(kv, cnt, found, okv, ocnt, ofound, check_query_index=False)
    assert int(found.max()) == n_docs or int(found.max()) >= n_docs - 25


def test_edge_cases(coll):
    n_docs, fds, flats, pts, gi, oi = coll
    # empty batch
    b = S.KwBatch([], [0])
    kv, cnt, found = gi.keyword_search(b, 8)
    assert len(cnt) == 0
    # combination whose tokens exist in no field / query without combinations / K = 1 / stride smaller than K
    qs = [S.Query([S.Combo([[S.NO_LIST, S.NO_LIST, S.NO_LIST]], 1)], topk=5),
          S.Query([], topk=5),
          S.Query([S.Combo([[0, S.NO_LIST, 0]], 1)], topk=1),
          S.Query([S.Combo([[1, 1, 1], [S.NO_LIST, S.NO_LIST, S.NO_LIST]], 2)], topk=300),     # a required token found nowhere is skipped
          S.Query([S.Combo([[2, 2, 2]] * 16, 16)], topk=250, num_query_tokens=16)]                 # TSGPU_MAX_TOKENS rows... 16*3 > 32 lists
    b = S.KwBatch(qs[:4], [0, 1, 2])
    kv, cnt, found = gi.keyword_search(b, 4)
    okv, ocnt, ofound = oi.keyword_search(b, 4)
    assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)
    assert cnt[0] == 0 and cnt[1] == 0 and cnt[2] == 1 and cnt[3] == 4
    with pytest.raises(capi.TsgpuError):
        gi.keyword_search(S.KwBatch(qs[4:], [0, 1, 2]), 4)
    with pytest.raises(capi.TsgpuError):
        gi.keyword_search(S.KwBatch([S.Query([S.Combo([[0]], 1)], topk=2000)], [0]), 4)
    # ten tokens (Match window limit) + duplicates of the same token
    row = [3, 3, 3]
    q10 = S.Query([S.Combo([[t, S.NO_LIST, S.NO_LIST] for t in (0, 1, 0, 2, 1, 0, 3, 0, 1, 0)], 10)], topk=50, num_query_tokens=10)
    b = S.KwBatch([q10], [0, 1, 2])
    kv, cnt, found = gi.keyword_search(b, 64)
    okv, ocnt, ofound = oi.keyword_search(b, 64)
    assert_kv_equal(kv, cnt, found, okv, ocnt, ofound)


def test_large_scale_properties_and_sample_parity():
    """1 M docs / 100 k vocabulary: size-independent properties on every query, oracle parity on a sample."""
    n_docs = 1_000_000
    fd = synth.make_string_field(n_docs, 100_000, 4, 12, seed=7, device="cuda")
    pts = synth.make_points(n_docs, 13)
    gi = capi.GpuIndex(n_docs, 0)
    gi.load_field(fd.flat)
    gi.load_sort_column(pts)
    toks = synth.sample_queries(fd, 512, 3, 99)
    sort = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
    rng = np.random.default_rng(2)
    fil = np.nonzero(rng.random(n_docs) < 0.1)[0].astype(np.uint32)
    qs = []
    for i, r in enumerate(toks):
        combos = [S.Combo([[int(t)] for t in r], 3)]
        if i % 3 == 0:
            combos.append(S.Combo([[int(r[0])], [int(r[1])], [int(min(r[2] + 1, 99_999))]], 3, total_cost=2))
        qs.append(S.Query(combos, topk=250, sort=sort, num_query_tokens=3, filter=0 if i % 2 else -1))
    b = S.KwBatch(qs, [0], [fil])
    kv, cnt, found = gi.keyword_search(b, 250)
    kv2, cnt2, found2 = gi.keyword_search(b, 250)
    assert (cnt == cnt2).all() and (found == found2).all() and (kv["key"] == kv2["key"]).all()       # idempotent
    fset = set(fil.tolist())
    for q in range(len(qs)):
        n = int(cnt[q])
        assert n == min(250, int(found[q]))
        keys = kv["key"][q, :n].astype(np.int64)
        sc = kv["scores"][q, :n]
        assert len(set(keys.tolist())) == n                                                          # de-duplicated
        tup = [(int(a), int(b_), int(c), int(k)) for (a, b_, c), k in zip(sc, keys)]
        assert tup == sorted(tup, reverse=True)                                                      # Topster::sort order
        assert (sc[:, 1] == pts[keys]).all()                                                         # sort column gathered right
        if q % 2:
            assert all(int(k) in fset for k in keys)                                                 # filter respected
        ids0 = fd.flat.ids[int(fd.flat.list_off[toks[q][0]]):int(fd.flat.list_off[toks[q][0] + 1])]
        assert np.isin(keys, ids0).all() or q % 3 == 0
    oi = ol.OracleIndex(n_docs, [fd.flat], [pts])
    sample = S.KwBatch(qs[:96], [0], [fil])
    okv, ocnt, ofound = oi.keyword_search(sample, 250, threads=16)
    assert_kv_equal(kv[:96], cnt[:96], found[:96], okv, ocnt, ofound)
    gi.close()
