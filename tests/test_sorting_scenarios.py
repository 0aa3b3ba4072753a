"""Sort-clause scenarios of the reference's test/collection_sorting_test.cpp replayed on the CPU oracle, the host-compiled
device functions and (with -m gpu) libtsgpu: numeric ASC/DESC with `_text_match` appended (src/collection.cpp:1736-1812),
large and equal int64 values, float columns through float_to_int64_t (src/index.cpp:266-274, 1164), two numeric clauses.
Fixtures: tests/golden/multi_field_documents.jsonl and float_documents.jsonl are byte copies of the reference's test data."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import refflow
from typesense_b200 import structs as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TM = (S.SORT_TEXT_MATCH, -1, 1, 0)
NONE = (S.SORT_NONE, -1, 1, 0)


def num(col, desc):
    return (S.SORT_NUMERIC, col, 1 if desc else -1, 0)


def f2i(x):
    return int(ol.oracle().tso_float_to_int64(float(np.float32(x))))


class Coll(refflow.Collection):
    """refflow.Collection with any number of numeric sort columns"""

    def __init__(self, docs, fields, cols):
        super().__init__([dict(d, points=0) for d in docs], fields)
        self.cols = [np.asarray(c, np.int64) for c in cols]


def run(backend_of, docs, fields, cols, q, sort, expect_ids, id_of=lambda d, i: str(d.get("id", i))):
    coll = Coll(docs, fields, cols)
    backend, close = backend_of(coll)
    got, found = refflow.search(backend, coll, q, sort)
    close()
    assert [id_of(docs[s], s) for s in got] == expect_ids, (q, sort)


def scenarios(backend_of):
    # SortingOrder :38-120
    docs = [json.loads(l) for l in open(os.path.join(GOLD, "multi_field_documents.jsonl")) if l.strip()]
    pts = [[d["points"] for d in docs]]
    run(backend_of, docs, ("title",), pts, "the", (num(0, False), TM, NONE), ["17", "13", "10", "4", "0", "1", "8", "6", "16", "11"])
    run(backend_of, docs, ("title",), pts, "the", (num(0, True), TM, NONE), ["11", "16", "6", "8", "1", "0", "10", "4", "13", "17"])
    run(backend_of, docs, ("title",), pts, "of", (TM, num(0, True), NONE), ["11", "12", "5", "4", "17"])
    # Int64AsDefaultSortingField :295-349
    big = [343234324234233234, 343234324234233232, 343234324234233235, 343234324234233231]
    docs = [{"title": "foo", "id": str(i)} for i in range(4)]
    run(backend_of, docs, ("title",), [big], "foo", (num(0, False), TM, NONE), ["3", "1", "0", "2"])
    run(backend_of, docs, ("title",), [big], "foo", (num(0, True), TM, NONE), ["2", "0", "1", "3"])
    # SortOnFloatFields :351-419
    docs = [json.loads(l) for l in open(os.path.join(GOLD, "float_documents.jsonl")) if l.strip()]
    cols = [[f2i(d["score"]) for d in docs], [f2i(d["average"]) for d in docs]]
    run(backend_of, docs, ("title",), cols, "Jeremy", (num(0, True), num(1, True), TM), ["2", "0", "3", "1", "5", "4", "6"])
    run(backend_of, docs, ("title",), cols, "Jeremy", (num(0, False), num(1, False), TM), ["6", "4", "5", "1", "3", "0", "2"])
    run(backend_of, docs, ("title",), cols, "Jeremy", (num(0, False), num(1, True), TM), ["5", "4", "6", "1", "3", "0", "2"])


def wildcard_scenarios(wildcard_backend_of):
    # WildcardQuery, test/collection_test.cpp:551-603: q=* over documents.jsonl (25 records with the fixture's dummy one)
    docs = [{"points": 10, "title": "z"}] + [json.loads(l) for l in open(os.path.join(GOLD, "documents.jsonl")) if l.strip()]
    coll = Coll(docs, ("title",), [[d["points"] for d in docs]])
    backend, close = wildcard_backend_of(coll)
    for sort, expect in (((num(0, True), (S.SORT_SEQ_ID, -1, 1, 0), NONE), None),
                         ((num(0, False), (S.SORT_SEQ_ID, -1, 1, 0), NONE), ["21", "24", "17"])):
        kv, cnt, found = backend(S.KwBatch([S.Query([], topk=250, sort=sort)], [0]), 250)
        assert int(found[0]) == 25 and int(cnt[0]) == 25
        if expect:
            assert [str(docs[int(kv["key"][0, i])].get("id", int(kv["key"][0, i]))) for i in range(3)] == expect
    close()
    # WildcardSearchSequenceIdSort, test/collection_sorting_test.cpp:1962-1986: 30 identical docs, _seq_id DESC
    docs = [{"category": "Shoes"} for _ in range(30)]
    coll = Coll(docs, ("category",), [[0] * 30])
    backend, close = wildcard_backend_of(coll)
    kv, cnt, found = backend(S.KwBatch([S.Query([], topk=250, sort=((S.SORT_SEQ_ID, -1, 1, 0), NONE, NONE))], [0]), 250)
    close()
    assert int(found[0]) == 30 and [int(kv["key"][0, i]) for i in range(10)] == list(range(29, 19, -1))


def test_sorting_scenarios_oracle():
    def mk(coll):
        oi = ol.OracleIndex(coll.n_docs, coll.flats, coll.cols)
        return (lambda b, k: oi.keyword_search(b, k)), (lambda: None)
    scenarios(mk)

    def mkw(coll):
        oi = ol.OracleIndex(coll.n_docs, coll.flats, coll.cols)
        return (lambda b, k: oi.wildcard_search(b, k)), (lambda: None)
    wildcard_scenarios(mkw)


@pytest.mark.gpu
def test_sorting_scenarios_gpu():
    from typesense_b200 import capi

    def mk(coll):
        gi = capi.GpuIndex(coll.n_docs, 0)
        for f in coll.flats:
            gi.load_field(f)
        for c in coll.cols:
            gi.load_sort_column(c)
        return (lambda b, k: gi.keyword_search(b, k)), gi.close
    scenarios(mk)

    def mkw(coll):
        gi = capi.GpuIndex(coll.n_docs, 0)
        for f in coll.flats:
            gi.load_field(f)
        for c in coll.cols:
            gi.load_sort_column(c)
        return (lambda b, k: gi.wildcard_search(b, k)), gi.close
    wildcard_scenarios(mkw)
