"""The reference's typo and prefix scenarios (test/collection_test.cpp: QueryWithTypo :374, TypoTokenRankedByScoreAndFrequency
:413, PrefixSearching :605, TypoTokensThreshold :686) through tests/typoflow.py — the host-side cost-combination /
candidate-product / drop-tokens control flow — on the CPU oracle and on the host-compiled device functions."""
import os

import pytest

import oracle_lib as ol
import refflow
import typoflow as tf
from typesense_b200 import structs as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SORT_DESC = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))

# (query, searcher options, expected leading ids, per_page, expected found or None)
CASES = [
    ("kind biologcal", dict(num_typos=2, prefix=False, drop_tokens_threshold=10, typo_tokens_threshold=10), ["19", "3", "20"], 3, None),
    ("lauxnch rcket", dict(num_typos=1, prefix=False, drop_tokens_threshold=10, typo_tokens_threshold=10), ["8", "1", "17"], 3, None),
    ("loox", dict(num_typos=1, token_order=tf.MAX_SCORE, prefix=False), ["22", "3"], 2, 5),
    ("loox", dict(num_typos=1, token_order=tf.FREQUENCY, prefix=False), ["22", "3", "12", "23", "24"], 10, 5),
    ("loox", dict(num_typos=1, token_order=tf.MAX_SCORE, prefix=False), ["22", "3", "12", "23", "24"], 10, 5),
    ("ex", dict(num_typos=0, token_order=tf.FREQUENCY, prefix=True), ["6", "12"], 10, 2),
    ("ex", dict(num_typos=0, token_order=tf.MAX_SCORE, prefix=True), ["6", "12"], 10, 2),
    ("what ex", dict(num_typos=0, token_order=tf.MAX_SCORE, prefix=True, drop_tokens_threshold=10, typo_tokens_threshold=10),
     ["6", "12", "19", "22", "13", "8", "15", "24", "21"], 10, 9),
    ("t", dict(num_typos=0, token_order=tf.MAX_SCORE, prefix=True, drop_tokens_threshold=10, typo_tokens_threshold=10), ["19", "22"], 2, None),
    ("t", dict(num_typos=0, token_order=tf.FREQUENCY, prefix=True, drop_tokens_threshold=10, typo_tokens_threshold=10), ["1", "2"], 2, None),
    ("math fx", dict(num_typos=0, prefix=True, drop_tokens_threshold=0), [], 1, 0),
    ("x", dict(num_typos=2, prefix=True), [], 2, 0),
    ("late propx", dict(num_typos=2, prefix=True), ["16"], 1, None),
    # TextContainingAnActualTypo :473-508
    ("ISSX what", dict(num_typos=1, prefix=False, drop_tokens_threshold=20, typo_tokens_threshold=20), ["19", "6", "21", "22"], 4, 11),
    ("ISSX", dict(num_typos=1, prefix=False, drop_tokens_threshold=10, typo_tokens_threshold=10), ["20", "19", "6", "3", "21"], 10, 5),
    # TypoTokensThreshold: typo correction only until typo_tokens_threshold results exist
    ("redundant", dict(num_typos=2, prefix=True, drop_tokens_threshold=10, typo_tokens_threshold=0), None, 10, 1),
    ("redundant", dict(num_typos=2, prefix=True, drop_tokens_threshold=10, typo_tokens_threshold=10), None, 10, 2),
]


def run_cases(backend, coll):
    for q, opts, expect, per_page, found_expect in CASES:
        got, found = tf.TypoSearcher(backend, coll, SORT_DESC, **opts).search(q)
        ids = [str(coll.docs[s].get("id", s)) for s in got][:per_page]
        if expect is not None:
            assert ids == expect, (q, opts, ids)
        if found_expect is not None:
            assert found == found_expect and (expect is not None or len(ids) == found_expect), (q, opts, found)


def small_collection_cases(make_backend):
    # PrefixRankedAfterExactMatch :3922-3960
    recs = ["Rotini Puttanesca", "Poulet Roti Tout Simple", "Chapatis (Roti)", "School Days Rotini Pasta Salad"]
    coll = refflow.Collection([{"title": t, "points": i} for i, t in enumerate(recs)], ("title",))
    got, found = tf.TypoSearcher(make_backend(coll), coll, SORT_DESC, num_typos=0, prefix=True, drop_tokens_threshold=5).search("roti")
    assert got[:3] == [2, 1, 3] and found == 4
    # MultiOccurrenceString :703-727
    coll = refflow.Collection([{"title": "The brown fox was the tallest of the lot and the quickest of the trot.", "points": 100}], ("title",))
    got, found = tf.TypoSearcher(make_backend(coll), coll, SORT_DESC, num_typos=0, prefix=False, drop_tokens_threshold=0).search("the")
    assert got == [0] and found == 1


def test_typo_and_prefix_scenarios_oracle():
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    oi = ol.OracleIndex(coll.n_docs, [coll.flat], [coll.points])
    run_cases(lambda b, k: oi.keyword_search(b, k), coll)

    def mk(c):
        o = ol.OracleIndex(c.n_docs, [c.flat], [c.points])
        return lambda b, k: o.keyword_search(b, k)
    small_collection_cases(mk)


def test_typo_and_prefix_scenarios_device_functions():
    import test_hostsim as th
    hs = th.hs.__wrapped__() if hasattr(th.hs, "__wrapped__") else None
    if hs is None:
        pytest.skip("hostsim fixture not callable directly")
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    run_cases(th.hostsim_backend(hs, coll), coll)
    small_collection_cases(lambda c: th.hostsim_backend(hs, c))


@pytest.mark.gpu
def test_typo_and_prefix_scenarios_gpu():
    from typesense_b200 import capi
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    gi = capi.GpuIndex(coll.n_docs, 0)
    gi.load_field(coll.flat)
    gi.load_sort_column(coll.points)
    run_cases(lambda b, k: gi.keyword_search(b, k), coll)
    gi.close()
    opened = []

    def mk(c):
        g = capi.GpuIndex(c.n_docs, 0)
        g.load_field(c.flat)
        g.load_sort_column(c.points)
        opened.append(g)
        return lambda b, k: g.keyword_search(b, k)
    small_collection_cases(mk)
    for g in opened:
        g.close()
