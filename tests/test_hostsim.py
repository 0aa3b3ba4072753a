"""CPU check of the product's __host__ __device__ headers (score_device.cuh / postings_device.cuh / postings_pack.h)
against the oracle: tests/hostsim compiles them with g++ and replays the per-thread call sequence of
kw_search_kernel. This is test infrastructure — the product itself has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import structs as S
from typesense_b200 import synth
from test_oracle_ref import random_batch, small_collection  # noqa: F401  (fixture)

HS_DIR = os.path.join(os.path.dirname(__file__), "hostsim")
HS_SO = os.path.join(HS_DIR, "libhostsim.so")


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HS_DIR, "hostsim.cpp")
    csrc = os.path.join(ol.ROOT, "typesense_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("score_device.cuh", "postings_device.cuh", "postings_pack.h")]
    if not os.path.exists(HS_SO) or any(os.path.getmtime(d) > os.path.getmtime(HS_SO) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++", src, "-o", HS_SO])
    L = C.CDLL(HS_SO)
    L.hs_keyword_combo.restype = C.c_size_t
    L.hs_keyword_combo.argtypes = [C.POINTER(S.FieldStruct), C.c_uint32, C.POINTER(S.KwBatchStruct), C.c_uint32, C.c_uint32,
                                   S.u32p, S.u64p, C.c_size_t]
    L.hs_phrase_match_doc.argtypes = [C.c_uint32, S.u32p, S.u32p]
    L.hs_set_reg_score.argtypes = [C.c_int]
    L.hs_reg_score_hits.restype = C.c_long
    L.hs_score_plain_both.argtypes = [C.c_uint32, C.c_uint32, S.u32p, S.u32p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.hs_ip_kat_data.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.hs_sort_scores.argtypes = [S.u8p, S.i8p, S.u8p, C.POINTER(C.POINTER(C.c_int64)), C.c_uint32, C.c_int64, C.c_float,
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.hs_idset_matches.restype = C.c_size_t
    L.hs_idset_matches.argtypes = [C.POINTER(S.FieldStruct), S.u32p, C.c_uint32, S.u32p, C.c_size_t, C.c_int, S.u32p]
    L.hs_probe_ids.argtypes = [S.u32p, C.c_uint64, S.u32p, C.c_uint64, S.u32p]
    return L


def hs_combo(L, flats, b, q, c, cap=1 << 20):
    arr = (S.FieldStruct * len(flats))(*[f.struct() for f in flats])
    ids = np.zeros(cap, np.uint32)
    sc = np.zeros(cap, np.uint64)
    s = b.struct()
    n = L.hs_keyword_combo(arr, len(flats), C.byref(s), q, c, ol.p32(ids), sc.ctypes.data_as(S.u64p), cap)
    assert n <= cap
    return ids[:n].copy(), sc[:n].copy()


@pytest.mark.parametrize("seed", range(4))
def test_device_scoring_matches_oracle(hs, small_collection, seed):
    n_docs, fds, pts = small_collection
    rng = np.random.default_rng(500 + seed)
    filters = [np.unique(rng.integers(0, n_docs, 1500)).astype(np.uint32), np.arange(0, n_docs, 7, dtype=np.uint32)]
    flats = [fd.flat for fd in fds]
    ix = ol.OracleIndex(n_docs, flats, [pts])
    b = random_batch(rng, fds, 40, filters)
    total = 0
    for q in range(b.n_queries):
        for c in range(int(b.q_combo_off[q]), int(b.q_combo_off[q + 1])):
            ids, sc, _ = ix.keyword_combo(b, q, c)
            hids, hsc = hs_combo(hs, flats, b, q, c)
            assert ids.tolist() == hids.tolist(), (q, c)
            assert sc.tolist() == hsc.tolist(), (q, c)
            total += len(ids)
    assert total > 200


def test_device_scoring_synonym_and_wide_positions(hs):
    # synonym rescale branch (src/index.cpp:7038-7061) + positions beyond 255 / 65535 wrap
    rng = np.random.default_rng(9)
    n_docs = 300
    lists = []
    for t in range(6):
        pl = []
        for d in np.unique(rng.integers(0, n_docs, 150)):
            k = int(rng.integers(1, 4))
            pos = np.sort(rng.choice([1, 2, 3, 4, 5, 9, 40, 254, 255, 256, 300, 65534, 65535, 65536, 70000], k, replace=False)).tolist()
            if rng.random() < 0.3:
                pos.append(0)
            pl.append((int(d), pos))
        lists.append(pl)
    flat = S.FlatField.from_postings(lists)
    ix = ol.OracleIndex(n_docs, [flat], [])
    qs = []
    for i in range(30):
        nt = int(rng.integers(1, 5))
        toks = rng.choice(6, nt, replace=(nt > 6)).tolist()
        syn = bool(i % 2)
        qs.append(S.Query([S.Combo([[t] for t in toks], nt, total_cost=int(rng.integers(0, 5)),
                                   flags=(S.CFLAG_SYNONYM | (S.CFLAG_DEMOTE_SYNONYM if i % 4 == 1 else 0)) if syn else 0,
                                   syn_orig_num_tokens=int(rng.integers(1, 4)) if syn else -1,
                                   orig_num_tokens=int(rng.integers(1, 4)) if syn else -1)],
                          flags=int(rng.integers(0, 8)), match_type=int(rng.integers(0, 3)), num_query_tokens=nt))
    b = S.KwBatch(qs, [0])
    n = 0
    for q in range(b.n_queries):
        ids, sc, _ = ix.keyword_combo(b, q, q)
        hids, hsc = hs_combo(hs, [flat], b, q, q)
        assert ids.tolist() == hids.tolist()
        assert sc.tolist() == hsc.tolist(), q
        n += len(ids)
    assert n > 100


def test_device_block_probe_random(hs):
    # dense and sparse lists, bit widths 0..32, via a 2-token AND through the packed-block probe
    rng = np.random.default_rng(3)
    for trial in range(12):
        hi = int(rng.choice([200, 5000, 1 << 20, (1 << 32) - 1]))
        a = np.unique(rng.integers(0, hi, int(rng.integers(1, 3000)), dtype=np.uint64)).astype(np.uint32)
        bb = np.unique(np.concatenate([rng.choice(a, min(len(a), 200)), rng.integers(0, hi, 500, dtype=np.uint64).astype(np.uint32)]))
        flat = S.FlatField.from_postings([[(int(i), [1]) for i in a], [(int(i), [2]) for i in bb]])
        b = S.KwBatch([S.Query([S.Combo([[0], [1]], 2)])], [0])
        ids, _ = hs_combo(hs, [flat], b, 0, 0)
        assert ids.tolist() == np.intersect1d(a, bb).tolist()


def test_device_probe_fuzz(hs):
    # every id of the list must be found at its index, every absent id must miss; all bit widths and block fills
    rng = np.random.default_rng(17)
    for trial in range(300):
        n = int(rng.choice([1, 2, 3, 127, 128, 129, 255, 256, 257, 1000, 5000]))
        kind = trial % 5
        if kind == 0:
            ids = np.sort(rng.choice(max(n, int(n * rng.uniform(1.0, 1.5))), n, replace=False))
        elif kind == 1:
            ids = np.unique(rng.integers(0, (1 << 32) - 1, n, dtype=np.uint64))
        elif kind == 2:   # clustered: dense runs separated by big gaps
            ids = np.unique(np.concatenate([np.arange(s, s + int(rng.integers(1, 300))) for s in rng.integers(0, 1 << 31, max(1, n // 100), dtype=np.uint64)]))
        elif kind == 3:   # skewed inside blocks
            ids = np.unique((rng.random(n) ** 4 * (1 << 24)).astype(np.uint64))
        else:
            ids = np.unique(rng.integers(0, max(2, 8 * n), n, dtype=np.uint64))
        ids = ids.astype(np.uint32)
        absent = rng.integers(0, int(ids[-1]) + 3, 500, dtype=np.uint64).astype(np.uint32)
        q = np.concatenate([ids, absent, np.asarray([0, int(ids[0]), int(ids[-1]), min(int(ids[-1]) + 1, (1 << 32) - 1)], np.uint32)])
        out = np.zeros(len(q), np.uint32)
        hs.hs_probe_ids(ol.p32(np.ascontiguousarray(ids)), len(ids), ol.p32(np.ascontiguousarray(q)), len(q), ol.p32(out))
        pos = np.searchsorted(ids, q)
        hit = (pos < len(ids)) & (ids[np.minimum(pos, len(ids) - 1)] == q)
        exp = np.where(hit, pos, 0xFFFFFFFF).astype(np.uint32)
        assert (out == exp).all(), (trial, kind, n)


def idset_cases(rng, fd, n_cases, kmax=4):
    """(lists, candidate ids) pairs: the first k tokens of a doc (prefix/exact hits), a random window (phrase hits) or
    random tokens; candidates = intersection of the lists, as at the reference's call sites."""
    n_docs = len(fd.doc_off) - 1
    for _ in range(n_cases):
        d = int(rng.integers(0, n_docs))
        a, e = int(fd.doc_off[d]), int(fd.doc_off[d + 1])
        if e == a:
            continue
        k = int(rng.integers(1, min(kmax, e - a) + 1))
        mode = int(rng.integers(0, 3))
        s = a if mode == 0 else int(rng.integers(a, e - k + 1))
        lists = fd.doc_tok[s:s + k].astype(np.uint32)
        if mode == 2:
            lists = lists[rng.permutation(k)]
        cand = None
        for l in lists:
            ids = fd.flat.ids[int(fd.flat.list_off[l]):int(fd.flat.list_off[l + 1])]
            cand = ids if cand is None else np.intersect1d(cand, ids)
        if cand is None or len(cand) == 0:
            continue
        yield np.ascontiguousarray(lists), np.ascontiguousarray(cand, np.uint32)


def test_device_idset_modes_match_oracle(hs, small_collection):
    n_docs, fds, _ = small_collection
    L = ol.oracle()
    rng = np.random.default_rng(4242)
    hits = {1: 0, 2: 0, 3: 0}
    for fi in (1, 2, 0):
        fd = fds[fi]
        ix = ol.OracleIndex(n_docs, [fd.flat], [])
        fs = fd.flat.struct()
        for lists, ids in idset_cases(rng, fd, 150):
            k = len(lists)
            for mode, fn in ((1, L.tso_phrase_matches), (2, L.tso_exact_matches), (3, L.tso_prefix_matches)):
                out = np.zeros(len(ids), np.uint32)
                n = fn(ix.h, 0, ol.p32(lists), k, ol.p32(ids), len(ids), ol.p32(out))
                hout = np.zeros(len(ids), np.uint32)
                hn = hs.hs_idset_matches(C.byref(fs), ol.p32(lists), k, ol.p32(ids), len(ids), mode, ol.p32(hout))
                assert out[:n].tolist() == hout[:hn].tolist(), (fi, mode, lists.tolist())
                hits[mode] += n
    assert min(hits.values()) > 20, hits


def test_ip_distance_reference_kat(hs):
    """test/collection_vector_search_test.cpp:5094-5196 (TestDistanceThresholdWithIP): inner-product distances 1 - dot on
    data regenerated with the reference's own RNG recipe; the rank_score literals prove the stream is the same."""
    vec = np.zeros((5, 5), np.float32)
    rank = np.zeros(5, np.int32)
    hs.hs_ip_kat_data(vec.ctypes.data_as(C.POINTER(C.c_float)), rank.ctypes.data_as(C.POINTER(C.c_int)))
    assert sorted(rank.tolist()) == sorted([93, 51, 94, 80, 18])
    L = ol.oracle()
    f32p = C.POINTER(C.c_float)

    def dist(q):
        q = np.asarray(q, np.float32)
        return [float(L.tso_ip_distance(q.ctypes.data_as(f32p), np.ascontiguousarray(vec[i]).ctypes.data_as(f32p), 5)) for i in range(5)]
    d = dist([0.11731103425347378, -0.6694758317235057, -0.6211945774857595, -0.27966758971688255, -0.4683744007950299])
    by_rank = {int(rank[i]): d[i] for i in range(5)}
    assert by_rank[93] == pytest.approx(0.2189185470342636, rel=1e-5) and by_rank[51] == pytest.approx(0.7371898889541626, rel=1e-5)
    assert all(by_rank[r] > 1.0 for r in (94, 80, 18))          # beyond distance_threshold:1 in the reference test
    d = dist([-100] * 5)
    expect = {1: -45.23314666748047, 2: -38.66290283203125, 4: -36.0988655090332, 3: 9.637892723083496, 0: 288.0364685058594}
    for i, e in expect.items():
        assert d[i] == pytest.approx(e, rel=1e-5), i
    assert sorted(range(5), key=lambda i: d[i]) == [1, 2, 4, 3, 0]


def hostsim_backend(L, coll):
    """refflow backend that runs every combination of a one-query batch through the product's device functions compiled
    for the host (probe, scoring, sort keys) and keeps the top-K on the CPU — the kernels' per-document logic without a GPU."""
    pts = np.ascontiguousarray(coll.points, np.int64)
    cols = (C.POINTER(C.c_int64) * 3)()

    def run(b, K):
        assert b.n_queries == 1
        best = {}
        for c in range(int(b.q_combo_off[0]), int(b.q_combo_off[1])):
            ids, sc = hs_combo(L, coll.flats, b, 0, c)
            for sid, text in zip(ids.tolist(), sc.tolist()):
                for i in range(3):
                    col = int(b.q_sort_col[i])
                    cols[i] = pts.ctypes.data_as(C.POINTER(C.c_int64)) if int(b.q_sort_type[i]) == S.SORT_NUMERIC and col == 0 else None
                out = (C.c_int64 * 3)()
                msi = C.c_int(0)
                L.hs_sort_scores(b.q_sort_type.ctypes.data_as(S.u8p), b.q_sort_order.ctypes.data_as(S.i8p),
                                 b.q_sort_missing_first.ctypes.data_as(S.u8p), cols, sid, np.int64(np.uint64(text).astype(np.int64)), 0.0,
                                 out, C.byref(msi))
                tup = (int(out[0]), int(out[1]), int(out[2]))
                if sid not in best or tup >= best[sid]:
                    best[sid] = tup
        order = sorted(best.items(), key=lambda kv_: (kv_[1], kv_[0]), reverse=True)[:K]
        kv = np.zeros((1, max(K, 1)), S.KV_DTYPE)
        for i, (sid, tup) in enumerate(order):
            kv["key"][0, i] = sid
            kv["scores"][0, i] = tup
        return kv, np.asarray([len(order)], np.uint32), np.asarray([len(best)], np.uint32)
    return run


def test_device_functions_reproduce_reference_scenarios(hs):
    """The reference's end-to-end expectations (incl. its literal text_match values) through score_device.cuh /
    postings_device.cuh compiled for the host: what the CUDA threads compute per document, checked here without a GPU."""
    import test_reference_scenarios as trs

    def mk(coll):
        return hostsim_backend(hs, coll), (lambda: None)
    trs.multi_field_scenarios(mk)
    trs.exact_match_scenario(mk)
    trs.match_ranking_scenarios(mk)
    trs.relevance2_scenarios(mk)
    trs.relevance36_scenarios(mk)
    trs.repeating_token_scenario(mk)
    trs.text_match_literals_scenario(mk)
    coll = trs.refflow.Collection.from_jsonl(os.path.join(trs.GOLD, "documents.jsonl"))
    trs.scenarios(hostsim_backend(hs, coll), coll)
    trs.more_scenarios(hostsim_backend(hs, coll), coll)


def test_register_resident_plain_scoring_equals_score_field_plain(hs):
    """score_field_plain_small() (the opt-in REGSCORE kernel's scoring: no sort, no run-time indexed arrays) against
    score_field_plain() on random well-formed plain-field documents: 1..4 rows, rows missing from the field, duplicate
    query tokens, documents whose last token is / is not a query token, phrases in and out of order, every switch that
    reaches the field score (exact match, token position, synonym rescaling and demotion)."""
    rng = np.random.default_rng(20260922)
    out = np.zeros(2, np.int64)
    n_multi = n_exact = 0
    for it in range(60000):
        n_rows = int(rng.integers(1, 5))
        L = max(n_rows + 1, int(rng.choice([3, 6, 12, 40, 300, 3000, 65535])))
        universe = np.arange(1, L + 1) if L <= 300 else np.unique(np.concatenate([rng.integers(1, L + 1, 60), [L]]))
        rows = [[] for _ in range(n_rows)]
        taken = set()
        if rng.random() < 0.5:             # the query tokens as a (possibly permuted / gapped) phrase somewhere in the doc
            gap = int(rng.choice([1, 1, 1, 2, 5, 11, 12]))
            start = int(rng.integers(1, max(2, L - n_rows * gap)))
            order = rng.permutation(n_rows) if rng.random() < 0.4 else np.arange(n_rows)
            for i in range(n_rows):
                pos = start + i * gap
                if pos <= min(L, 65535) and pos not in taken:
                    rows[int(order[i])].append(pos)
                    taken.add(pos)
        for pos in universe.tolist():      # every other position belongs to one row or to a token outside the query
            if pos not in taken and rng.random() < 0.45:
                rows[int(rng.integers(0, n_rows))].append(pos)
                taken.add(pos)
        for r in range(n_rows):            # a matched row holds at least one position
            if not rows[r]:
                pos = next(x for x in range(1, 70000) if x not in taken)
                rows[r].append(pos)
                taken.add(pos)
            rows[r].sort()
        verbatim = rng.random() < 0.08
        if verbatim:                        # the document IS the query (or the query plus a few more tokens)
            rows = [[i + 1] for i in range(n_rows)]
            if rng.random() < 0.3:
                rows[int(rng.integers(0, n_rows))].append(n_rows + 3)
        if n_rows >= 2 and rng.random() < 0.15:          # the same token twice in the query: identical lists
            a, b2 = rng.choice(n_rows, 2, replace=False)
            rows[int(b2)] = list(rows[int(a)])
        doc_last = max(max(r) for r in rows)
        if rng.random() < 0.6:                           # the doc's last token is a query token: trailing 0 on every row holding it
            for r in rows:
                if r[-1] == doc_last:
                    r.append(0)
        present = (1 << n_rows) - 1 if verbatim else int(rng.integers(1, 1 << n_rows))
        tok_off = np.zeros(n_rows + 1, np.uint32)
        tok_off[1:] = np.cumsum([len(r) for r in rows])
        raw = np.asarray([x for r in rows for x in r], np.uint32)
        syn = rng.random() < 0.25
        nq = int(rng.integers(1, 5))
        params = np.asarray([int(rng.integers(0, 5)), nq, int(rng.choice([-1, 1, 2, 3, 4])) if syn else -1, int(rng.choice([-1, 1, 2, 3])) if syn else -1,
                             1 if syn else 0, int(rng.integers(0, 2)) if syn else 0, int(rng.integers(0, 2)), int(rng.integers(0, 2))], np.int32)
        hs.hs_score_plain_both(n_rows, present, ol.p32(tok_off), ol.p32(raw), params.ctypes.data_as(C.POINTER(C.c_int32)),
                               out.ctypes.data_as(C.POINTER(C.c_int64)))
        assert out[0] == out[1], (it, rows, present, params.tolist(), out.tolist())
        n_multi += bin(present).count("1") > 1
        n_exact += (int(out[0]) >> 12) & 1
    assert n_multi > 20000 and n_exact > 500, (n_multi, n_exact)


def test_reference_scenarios_with_register_resident_scoring(hs, small_collection):
    """The scenario replays and the random-combination parity above, once more with the REGSCORE branch switched on."""
    hs.hs_set_reg_score(1)
    before = hs.hs_reg_score_hits()
    try:
        for seed in range(4):
            test_device_scoring_matches_oracle(hs, small_collection, seed)
        test_device_scoring_synonym_and_wide_positions(hs)          # positions beyond 65535: the branch must stand aside
        test_device_functions_reproduce_reference_scenarios(hs)
    finally:
        hs.hs_set_reg_score(0)
    assert hs.hs_reg_score_hits() - before > 2000           # the branch really ran
