"""numeric / bool filter leaves and the AND / OR / NOT tree evaluated on the device (tsgpu_filter_numeric,
tsgpu_filter_combine), against the comparator semantics of the reference's numeric index (src/num_tree.cpp: a doc without a
value is in no leaf; `!=` = every doc minus the equal ones, src/filter_result_iterator.cpp apply_not_equals) restated in numpy,
and then USED: a keyword search under the device-built filter must equal the oracle's search under the same id list."""
import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import capi, structs as S, synth

pytestmark = pytest.mark.gpu
MISSING = np.iinfo(np.int64).min


def expect_ids(col, op, v1, v2=0):
    has = col != MISSING
    m = {"=": has & (col == v1), "!=": ~(has & (col == v1)), "<": has & (col < v1), "<=": has & (col <= v1), ">": has & (col > v1),
         ">=": has & (col >= v1), "range": has & (col >= v1) & (col <= v2)}[op]
    return np.nonzero(m)[0].astype(np.uint32)


def test_numeric_leaves_tree_and_search_under_a_device_filter():
    n_docs = 40000
    fd = synth.make_string_field(n_docs, 300, 4, 10, seed=21)
    pts = synth.make_points(n_docs, 3, hi=1000, missing_frac=0.1)
    flag = (np.random.default_rng(5).integers(0, 2, n_docs)).astype(np.int64)          # a bool field
    gi = capi.GpuIndex(n_docs, 0)
    gi.load_field(fd.flat)
    c_pts = gi.load_sort_column(pts)
    c_flag = gi.load_sort_column(flag)
    handles = {}
    for op, v1, v2 in [("=", 17, 0), ("!=", 17, 0), ("<", 100, 0), ("<=", 100, 0), (">", 900, 0), (">=", 900, 0), ("range", 250, 260), ("=", 5000, 0)]:
        h, n = gi.filter_numeric(c_pts, op, v1, v2)
        e = expect_ids(pts, op, v1, v2)
        assert n == len(e) and gi.filter_ids(h, n_docs).tolist() == e.tolist(), (op, v1)
        handles[(op, v1)] = (h, e)
    hb, nb = gi.filter_numeric(c_flag, "=", 1)
    eb = expect_ids(flag, "=", 1)
    assert gi.filter_ids(hb, n_docs).tolist() == eb.tolist()
    # points:<100 && flag:true ; points:>900 || points:[250..260] ; flag:true && !(points:<=100)
    (h1, e1), (h2, e2), (h3, e3), (h4, e4) = handles[("<", 100)], handles[(">", 900)], handles[("range", 250)], handles[("<=", 100)]
    ha, na = gi.filter_combine(capi_set("and"), h1, hb)
    assert gi.filter_ids(ha, n_docs).tolist() == np.intersect1d(e1, eb).tolist() and na == len(np.intersect1d(e1, eb))
    ho, no = gi.filter_combine(capi_set("or"), h2, h3)
    assert gi.filter_ids(ho, n_docs).tolist() == np.union1d(e2, e3).tolist()
    hx, nx = gi.filter_combine(capi_set("exclude"), hb, h4)
    assert gi.filter_ids(hx, n_docs).tolist() == np.setdiff1d(eb, e4).tolist()
    # a search under the device-built filter == the oracle's search under the same ids given inline
    oi = ol.OracleIndex(n_docs, [fd.flat], [pts, flag])
    toks = synth.sample_queries(fd, 30, 2, 4)
    sort = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
    qs_dev, qs_ora = [], []
    for row in toks:
        q = S.Query([S.Combo([[int(t)] for t in row], 2)], topk=50, sort=sort, num_query_tokens=2)
        q2 = S.Query([S.Combo([[int(t)] for t in row], 2)], topk=50, sort=sort, num_query_tokens=2)
        q.filter = 0; q2.filter = 0
        qs_dev.append(q); qs_ora.append(q2)
    bd = S.KwBatch(qs_dev, [0], [np.zeros(0, np.uint32)]).with_filter_handles([ho])
    kv, cnt, found = gi.keyword_search(bd, 50)
    okv, ocnt, ofound = oi.keyword_search(S.KwBatch(qs_ora, [0], [np.union1d(e2, e3).astype(np.uint32)]), 50)
    assert cnt.tolist() == ocnt.tolist() and found.tolist() == ofound.tolist()
    for q in range(len(qs_dev)):
        assert kv["key"][q, :cnt[q]].tolist() == okv["key"][q, :ocnt[q]].tolist()
    gi.close()


def capi_set(name):
    return {"and": 0, "or": 1, "exclude": 2}[name]
