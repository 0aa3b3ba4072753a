"""facet counting over result ids (Index::do_facets hash-index branch, src/index.cpp:1674-1780, then
Collection::search's (count, id) order). CPU: the oracle's restatement on literal cases worked out from the reference's
loop. GPU: tsgpu_facet_counts / tsgpu_facet_counts_last / tsgpu_all_result_ids_last against the oracle."""
import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import capi, structs as S, synth


def make_facet(n_docs, n_values, seed, max_per_doc=4, missing=0.2):
    rng = np.random.default_rng(seed)
    off = [0]
    vals = []
    w = np.arange(1, n_values + 1, dtype=np.float64) ** -1.1
    cdf = np.cumsum(w / w.sum())
    for d in range(n_docs):
        k = 0 if rng.random() < missing else int(rng.integers(1, max_per_doc + 1))
        v = np.minimum(np.searchsorted(cdf, rng.random(k)), n_values - 1).tolist()       # duplicates inside a doc happen (array facets)
        vals.extend(v)
        off.append(len(vals))
    return np.asarray(off, np.uint64), np.asarray(vals if vals else [0], np.uint32)


def test_oracle_facet_counts_literal():
    # docs: 0 -> [2, 2, 5]   1 -> []   2 -> [5]   3 -> [7, 2]   4 -> [5, 5]
    off = np.asarray([0, 3, 3, 4, 6, 8], np.uint64)
    vals = np.asarray([2, 2, 5, 5, 7, 2, 5, 5], np.uint32)
    out, dis = ol.facet_counts(5, 8, off, vals, [0, 2, 3, 4], 10)
    # value 5: docs 0, 2, 4 -> count 3, last doc 4 at position 0; value 2: docs 0, 3 -> count 2, last doc 3 at position 1;
    # value 7: doc 3 -> count 1 at position 0. Order: (count, id) descending.
    assert dis == 3
    assert [(int(e["value_id"]), int(e["count"]), int(e["doc_id"]), int(e["array_pos"])) for e in out] == [(5, 3, 4, 0), (2, 2, 3, 1), (7, 1, 3, 0)]
    out, dis = ol.facet_counts(5, 8, off, vals, [0, 2, 3, 4], 10, sample_mod=2)            # estimate_facets: results 0 and 2 of the list
    assert [(int(e["value_id"]), int(e["count"])) for e in out] == [(2, 2), (7, 1), (5, 1)]
    out, dis = ol.facet_counts(5, 8, off, vals, [1], 10)
    assert len(out) == 0 and dis == 0


@pytest.mark.gpu
def test_gpu_facet_counts_match_oracle():
    n_docs, n_values = 50000, 3000
    off, vals = make_facet(n_docs, n_values, 1)
    gi = capi.GpuIndex(n_docs, 0)
    f = gi.load_facet(n_values, off, vals)
    rng = np.random.default_rng(2)
    for n_ids, top_n, mod in ((0, 10, 0), (1, 10, 0), (700, 10, 0), (20000, 100, 0), (20000, 250, 3), (n_docs, 1024, 0)):
        ids = np.sort(rng.choice(n_docs, n_ids, replace=False)).astype(np.uint32) if n_ids < n_docs else np.arange(n_docs, dtype=np.uint32)
        got, dis = gi.facet_counts(f, ids, top_n, mod)
        exp, edis = ol.facet_counts(n_docs, n_values, off, vals, ids, top_n, mod)
        assert dis == edis and got.tolist() == exp.tolist(), (n_ids, top_n, mod)
    gi.close()


@pytest.mark.gpu
def test_gpu_all_result_ids_and_facets_of_a_search_batch():
    """TSGPU_QFLAG_KEEP_ALL_IDS: the all_result_ids of every query of a keyword batch stay on the device; they must be the union
    over the query's combinations of (AND over tokens) minus nothing (no filter / exclusion here), their size must equal
    `found`, and the facet counts over them must equal the oracle's over the same ids."""
    n_docs, n_values = 30000, 500
    fd = synth.make_string_field(n_docs, 400, 4, 10, seed=5)
    pts = synth.make_points(n_docs, 9)
    off, vals = make_facet(n_docs, n_values, 7)
    gi = capi.GpuIndex(n_docs, 0)
    gi.load_field(fd.flat)
    gi.load_sort_column(pts)
    f = gi.load_facet(n_values, off, vals)
    oi = ol.OracleIndex(n_docs, [fd.flat], [pts])
    toks = synth.sample_queries(fd, 40, 2, 3)
    rng = np.random.default_rng(4)
    qs = []
    for i, row in enumerate(toks):
        combos = [S.Combo([[int(t)] for t in row], 2)]
        if i % 3 == 0:
            combos.append(S.Combo([[int(row[0])], [int(rng.integers(0, 50))]], 2, total_cost=1))
        q = S.Query(combos, topk=50, sort=((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0)), num_query_tokens=2)
        if i % 4 != 3:
            q.flags |= capi.QFLAG_KEEP_ALL_IDS
        qs.append(q)
    b = S.KwBatch(qs, [0])
    kv, cnt, found = gi.keyword_search(b, 50)
    okv, ocnt, ofound = oi.keyword_search(S.KwBatch([S.Query(q.combos, topk=50, sort=q.sort, num_query_tokens=2) for q in qs], [0]), 50)
    assert found.tolist() == ofound.tolist() and cnt.tolist() == ocnt.tolist()
    fc, fn, fdis = gi.facet_counts_last(f, len(qs), 20)
    lo = fd.flat.list_off.astype(np.int64)
    for qi, q in enumerate(qs):
        if not (q.flags & capi.QFLAG_KEEP_ALL_IDS):
            assert fn[qi] == 0
            with pytest.raises(capi.TsgpuError):
                gi.all_result_ids_last(qi, 10)
            continue
        expect = np.zeros(0, np.uint32)
        for c in q.combos:
            ids = None
            for r in c.rows:
                l = fd.flat.ids[lo[r[0]]:lo[r[0] + 1]]
                ids = l if ids is None else np.intersect1d(ids, l)
            expect = np.union1d(expect, ids)
        got = gi.all_result_ids_last(qi, n_docs)
        assert got.tolist() == expect.tolist() and len(got) == found[qi], qi
        exp, edis = ol.facet_counts(n_docs, n_values, off, vals, expect, 20)
        assert fdis[qi] == edis and fc[qi, :fn[qi]].tolist() == exp.tolist(), qi
    gi.close()
