"""String `filter_by` expectations of the reference's own tests, replayed through the id-set primitives of this path
(intersect, phrase / exact / prefix matches, or / exclude) on three implementations: the CPU oracle, the device
functions compiled for the host, and the C-ABI (GPU; on CPU through the test double in test_gpu_tests_dryrun.py).

Sources: test/collection_filtering_test.cpp (FilterOnTextFields :38, FilterByExactPhraseMatch* :218-302,
FacetFieldStringFiltering :467, FacetFieldStringArrayFiltering :535, ExactFilteringSingleQueryTerm :2301,
ExactFilteringRepeatingTokens* :2347-2450, PrefixFilterOnTextFields :2860) and test/collection_specific_more_test.cpp
(ExactFilteringOnArray2 :729). The numeric_array_documents rows are the string columns of the reference's
test/numeric_array_documents.jsonl."""
import json
import os

import numpy as np
import pytest

import filterflow as ff
import oracle_lib as ol
import refflow
from test_hostsim import hs  # noqa: F401  (fixture: the device functions compiled for the host)
from typesense_b200 import structs as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")

NUMERIC_ARRAY_DOCS = [
    {"name": "Jeremy Howard", "points": 24, "tags": ["gold", "silver"]},
    {"name": "Jeremy Howard", "points": 44, "tags": ["FINE PLATINUM"]},
    {"name": "Jeremy Howard", "points": 21, "tags": ["bronze", "gold"]},
    {"name": "Jeremy Howard", "points": 63, "tags": ["silver"]},
    {"name": "Jeremy Howard", "points": 32, "tags": ["silver", "gold", "bronze"]},
]


def _multi_field_docs():
    return [json.loads(l) for l in open(os.path.join(GOLD, "multi_field_documents.jsonl")) if l.strip()]


def _p(docs):
    return [dict(d, points=d.get("points", 0)) for d in docs]


# (source, docs, fields, [(field, raw filter value, expected seq_ids)])
CASES = [
    ("FilterOnTextFields :38", NUMERIC_ARRAY_DOCS, ("name", "tags"), [
        ("tags", "gold", [0, 2, 4]),
        ("tags", "fine PLATINUM", [1]),
        ("tags", "foobarbaz", []),
        ("tags", "PLATINUM", [1]),
        ("tags", "WHITE", []),
        ("tags", "WHITE PLATINUM", []),
        ("tags", "= PLATINUM", []),
        ("tags", "bronze", [2, 4]),
        ("tags", "[bronze,   silver]", [0, 2, 3, 4]),
    ]),
    ("FilterOnTextFields :143 (title / titles)", _p([
        {"title": "foo bar baz", "titles": []},
        {"title": "foo bar baz", "titles": ["foo bar baz"]},
        {"title": "foo bar baz", "titles": ["bar foo baz", "foo bar baz"]},
        {"title": "bar foo baz", "titles": ["bar foo baz"]}]), ("title", "titles"), [
        ("title", "= foo bar baz", [0, 1, 2]),
        ("titles", "= foo bar baz", [1, 2]),
    ]),
    ("FilterByExactPhraseMatch :218", _p([
        {"text": "z"},
        {"text": "Lewis Hamilton has won multiple Formula One World Championships."},
        {"text": "The scientist created a new formula, and this was just one of many groundbreaking discoveries in the lab."},
        {"text": "Formula One is a popular sport."}]), ("text",), [
        ("text", '"Formula One"', [1, 3]),
    ]),
    ("FilterByNegatedExactPhraseMatch :236", _p([
        {"text": "z"}, {"text": "this is a test"}, {"text": "this is not a test"}, {"text": "another test case"}]), ("text",), [
        ("text", '!="this is a test"', [0, 2, 3]),          # seq 0 is this harness's padding document
    ]),
    ("FilterByExactPhraseMatchInArray :259 / Negated :282", _p([
        {"tags": ["zz"]}, {"tags": ["new york", "travel"]}, {"tags": ["new", "york", "travel"]}, {"tags": ["paris", "travel"]},
        {"tags": ["new york", "paris"]}]), ("tags",), [
        ("tags", '["new york", paris]', [1, 3, 4]),
        ("tags", '!=["new york", paris]', [0, 2]),
    ]),
    ("FacetFieldStringFiltering :467", None, ("title", "starring", "cast"), [
        ("starring", "= samuel", []),
        ("starring", "= ssamuel l. Jackson", []),
        ("starring", "= samuel l. Jackson", 2),
        ("starring", "= `samuel l. Jackson`", 2),
        ("starring", "jackson", 2),
        ("starring", "samuel", 2),
        ("starring", "samuel johnson", []),
    ]),
    ("FacetFieldStringArrayFiltering :535", NUMERIC_ARRAY_DOCS, ("name", "tags"), [
        ("tags", "= PLATINUM", []),
        ("tags", "= FINE", []),
        ("tags", "= FFINE PLATINUM", []),
        ("tags", "PLATINUM", [1]),
        ("tags", "FINE", [1]),
        ("tags", "= FINE PLATINUM", [1]),
        ("name", "= Jeremy Howard", [0, 1, 2, 3, 4]),
        ("tags", "= [Gold, bronze]", [0, 2, 4]),
        ("tags", "= [Gold, bronze, fine PLATINUM]", [0, 1, 2, 4]),
        ("tags", "= [fine PLATINUM]", [1]),
    ]),
    ("ExactFilteringSingleQueryTerm :2301", _p([
        {"name": "AT&T GoPhone", "tags": ["AT&T GoPhone"]}, {"name": "AT&T", "tags": ["AT&T"]},
        {"name": "Phone", "tags": ["Samsung Phone", "Phone"]}]), ("name", "tags"), [
        ("name", "=AT&T", [1]),
        ("tags", "=AT&T", [1]),
        ("tags", "=Phone", [2]),
    ]),
    ("ExactFilteringRepeatingTokensSingularField :2347", _p([
        {"name": "Cardiology - Interventional Cardiology"}, {"name": "Cardiology - Interventional"},
        {"name": "Cardiology - Interventional Cardiology Department"},
        {"name": "Interventional Cardiology - Interventional Cardiology"}]), ("name",), [
        ("name", "=Cardiology - Interventional Cardiology", [0]),
        ("name", "=Cardiology - Interventional", [1]),
        ("name", "=Interventional Cardiology", []),
        ("name", "=Cardiology", []),
    ]),
    ("ExactFilteringRepeatingTokensArrayField :2394", _p([
        {"name": ["Cardiology - Interventional Cardiology"]}, {"name": ["Cardiology - Interventional"]},
        {"name": ["Cardiology - Interventional Cardiology Department"]},
        {"name": ["Interventional Cardiology - Interventional Cardiology"]}]), ("name",), [
        ("name", "=Cardiology - Interventional Cardiology", [0]),
        ("name", "=Cardiology - Interventional", [1]),
        ("name", "=Interventional Cardiology", []),
        ("name", "=Cardiology", []),
    ]),
    ("ExactFilteringOnArray2 (collection_specific_more_test.cpp:729)", _p([
        {"capability": ["Encoding capabilities for network communications", "Obfuscation capabilities"]}]), ("capability",), [
        ("capability", "=Encoding capabilities", []),
    ]),
    # hits are ordered by points DESC there; ids of multi_field_documents.jsonl are line numbers
    ("PrefixFilterOnTextFields :2860", None, ("title", "starring", "cast"), [
        ("cast", "Chris", [1, 6, 7, 8]),
        ("cast", "Ch*", [1, 6, 7, 8]),
        ("cast", "M*", [2, 3, 16]),
        ("cast", "Chris P*", [1, 7]),
        ("cast", "[Martin, Chris P*]", [1, 2, 7]),
        ("cast", "[M*, Chris P*]", [1, 2, 3, 7, 16]),
    ]),
    ("PrefixFilterOnTextFields :2950 (Names)", _p([{"name": "Steve Jobs"}, {"name": "Adam Stator"}]), ("name",), [
        ("name", "= S*", [0]),
        ("name", "S*", [0, 1]),
    ]),
    ("PrefixFilterOnTextFields :3020-3100 (Names, 5 docs)", _p([
        {"name": "Steve Jobs"}, {"name": "Adam Stator"}, {"name": "Steve Reiley"}, {"name": "Storm"}, {"name": "Steve Rogers"}]), ("name",), [
        ("name", "= St*", [0, 2, 3, 4]),
        ("name", "St*", [0, 1, 2, 3, 4]),
        ("name", "= Steve R*", [2, 4]),
        ("name", "Steve R*", [2, 4]),
    ]),
    ("PrefixFilterOnTextFields :3100-3230 (names[])", _p([
        {"names": []}, {"names": ["Steve Jobs"]}, {"names": ["Adam Stator"]}, {"names": ["Steve Reiley"]}, {"names": ["Storm"]},
        {"names": ["Adam", "Steve Rogers"]}]), ("names",), [
        ("names", "= St*", [1, 3, 4, 5]),
        ("names", "St*", [1, 2, 3, 4, 5]),
        ("names", "= Steve*", [1, 3, 5]),
    ]),
]


def run_cases(make_ops):
    n = 0
    for src, docs, fields, rows in CASES:
        coll = refflow.Collection(docs if docs is not None else _multi_field_docs(), fields)
        ops, close = make_ops(coll)
        try:
            for field, raw, expect in rows:
                got = ff.string_filter_ids(ops, coll, field, raw)
                if isinstance(expect, int):
                    assert len(got) == expect, (src, field, raw, got)
                else:
                    assert got == expect, (src, field, raw, got)
                n += 1
        finally:
            close()
    return n


def test_parse_string_filter():
    """filter.cpp:674-733"""
    assert ff.parse_string_filter("= PLATINUM") == (["PLATINUM"], [ff.EQUALS], False)
    assert ff.parse_string_filter("gold") == (["gold"], [ff.CONTAINS], False)
    assert ff.parse_string_filter("!= [a, b c]") == (["a", "b c"], [ff.NOT_EQUALS] * 2, True)
    assert ff.parse_string_filter("! a") == (["a"], [ff.CONTAINS], True)
    assert ff.parse_string_filter('"Formula One"') == (["Formula One"], [ff.CONTAINS_PHRASE], False)
    assert ff.parse_string_filter('["new york", paris]') == (["new york", "paris"], [ff.CONTAINS_PHRASE, ff.EQUALS], False)
    assert ff.parse_string_filter("[`a, b`, c]") == (["a, b", "c"], [ff.CONTAINS] * 2, False)


def test_filter_scenarios_oracle():
    assert run_cases(lambda coll: (ff.OracleOps(coll), lambda: None)) == sum(len(c[3]) for c in CASES)


def test_filter_scenarios_device_functions(hs):
    run_cases(lambda coll: (ff.HostsimOps(coll, hs), lambda: None))


def filtered_search_order(run, coll, ids, token="jeremy"):
    """`q: Jeremy, filter_by: tags: ...` sorted by age DESC (FilterOnTextFields): the filter ids go in as the query's filter."""
    fids = list(range(len(coll.fields)))
    rows = [[v.get(token, S.NO_LIST) for v in coll.vocabs]]
    q = S.Query([S.Combo(rows, 1)], topk=10, filter=0, sort=((S.SORT_NUMERIC, 0, 1, 0), (S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NONE, -1, 1, 0)),
                field_weight=[15 - f for f in fids])
    kv, cnt, found = run(S.KwBatch([q], fids, filters=[np.asarray(ids, np.uint32)]), 10)
    return [int(kv["key"][0, i]) for i in range(int(cnt[0]))], int(found[0])


def check_filtered_search(make):
    coll = refflow.Collection(NUMERIC_ARRAY_DOCS, ("name", "tags"))
    ops, run, close = make(coll)
    try:
        for raw, expect in (("gold", [4, 0, 2]), ("bronze", [4, 2]), ("[bronze,   silver]", [3, 4, 0, 2]), ("= [Gold, bronze, fine PLATINUM]", [1, 4, 0, 2])):
            ids = ff.string_filter_ids(ops, coll, "tags", raw)
            got, found = filtered_search_order(run, coll, ids)
            assert got == expect and found == len(expect), (raw, got, found)
    finally:
        close()


def test_filtered_search_oracle():
    def make(coll):
        oi = ol.OracleIndex(coll.n_docs, coll.flats, [coll.points])
        return ff.OracleOps(coll), (lambda b, k: oi.keyword_search(b, k)), (lambda: None)
    check_filtered_search(make)


@pytest.mark.gpu
def test_filter_scenarios_gpu():
    from typesense_b200 import capi

    def mk(coll):
        gi = capi.GpuIndex(coll.n_docs, 0)
        fids = [gi.load_field(f) for f in coll.flats]
        return ff.CapiOps(coll, gi, fids), gi.close
    run_cases(mk)

    def make(coll):
        gi = capi.GpuIndex(coll.n_docs, 0)
        fids = [gi.load_field(f) for f in coll.flats]
        gi.load_sort_column(coll.points)
        return ff.CapiOps(coll, gi, fids), (lambda b, k: gi.keyword_search(b, k)), gi.close
    check_filtered_search(make)
