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


def test_typo_and_prefix_scenarios_oracle():
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    oi = ol.OracleIndex(coll.n_docs, [coll.flat], [coll.points])
    run_cases(lambda b, k: oi.keyword_search(b, k), coll)


def test_typo_and_prefix_scenarios_device_functions():
    import test_hostsim as th
    hs = th.hs.__wrapped__() if hasattr(th.hs, "__wrapped__") else None
    if hs is None:
        pytest.skip("hostsim fixture not callable directly")
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    run_cases(th.hostsim_backend(hs, coll), coll)


@pytest.mark.gpu
def test_typo_and_prefix_scenarios_gpu():
    from typesense_b200 import capi
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    gi = capi.GpuIndex(coll.n_docs, 0)
    gi.load_field(coll.flat)
    gi.load_sort_column(coll.points)
    run_cases(lambda b, k: gi.keyword_search(b, k), coll)
    gi.close()
