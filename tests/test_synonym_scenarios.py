"""Synonym scenarios of the reference's test/collection_synonyms_test.cpp through tests/typoflow.py: the synonym variants of
a query run as queries of their own (Index::do_synonym_search, src/index.cpp:6088-6142) with syn_orig_num_tokens /
orig_num_tokens / is_synonym_query driving the rescaling branch of score_results2 (src/index.cpp:7038-7061) and
demote_synonym_match. The SynonymIndex lookup itself is host work: the resolved variants are given as input. Oracle, host-
compiled device functions and (with -m gpu) libtsgpu."""
import pytest

import oracle_lib as ol
import refflow
import typoflow as tf
from typesense_b200 import structs as S

SORT = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
TD, T = ("title", "description"), ("title",)
LOL = [{"title": t, "description": x, "points": p} for t, x, p in
       [("Laughing out Loud", "Description 1", 100), ("Stop Laughing", "Description 2", 120), ("LOL sure", "Laughing out loud sure", 200),
        ("Really ROFL now", "Description 3", 250)]]
LOL2 = [{"title": t, "description": x, "points": p} for t, x, p in
        [("LOL really", "Description 1", 50), ("Never stop", "Description 2", 120), ("Yes and no", "Laughing out loud sure", 100),
         ("And so on", "Description 3", 250)]]
LEMON = [{"title": "Smashed Lemon", "points": 100}, {"title": "Lulu Lemon", "points": 100}, {"title": "Lululemon", "points": 200}]

# (reference test, fields, docs, query, synonym variants, options, expected ids, relation between the first two text_match values)
CASES = [
    ("SynonymsTextMatchSameAsRootQuery :497", ("name", "title"),
     [{"name": "Dan Fisher", "title": "Chief Executive Officer", "points": 10}, {"name": "Jack Sparrow", "title": "CEO", "points": 20}],
     "ceo", ["chief executive officer"], dict(num_typos=0, prefix=True, drop_tokens_threshold=0), [1, 0], "eq"),
    ("ExactMatchRankedSameAsSynonymMatch :589", T, LOL, "laughing", ["lol", "rofl"], dict(num_typos=0, prefix=False, drop_tokens_threshold=0), [3, 2, 1, 0], None),
    ("ExactMatchVsSynonymMatchCrossFields :644", TD,
     [{"title": "Head of Marketing", "description": "The Chief Marketing Officer", "points": 100},
      {"title": "VP of Sales", "description": "Preparing marketing and sales materials.", "points": 120}],
     "cmo", ["Chief Marketing Officer", "VP of Marketing"], dict(num_typos=0, prefix=False, drop_tokens_threshold=0), [0, 1], None),
    ("SynonymFieldOrdering :696", TD, LOL2, "laughing", ["lol", "rofl"], dict(num_typos=0, prefix=False, drop_tokens_threshold=0), [0, 2], None),
    ("SynonymSingleTokenExactMatch :852", T,
     [{"title": "Smashed Lemon", "points": 100}, {"title": "Lulu Guinness", "points": 100}, {"title": "Lululemon", "points": 100}],
     "lulu lemon", ["lululemon"], dict(num_typos=2, prefix=True, drop_tokens_threshold=0), [2], None),
    ("SynonymExpansionAndCompressionRanking :894 (expansion)", T, LEMON, "lululemon", ["lulu lemon"], dict(num_typos=2, prefix=True, drop_tokens_threshold=0), [2, 1], "eq"),
    ("SynonymExpansionAndCompressionRanking :894 (compression)", T, LEMON, "lulu lemon", ["lululemon"], dict(num_typos=2, prefix=True, drop_tokens_threshold=0), [2, 1], "eq"),
    ("SynonymMatchShouldNotOutrankCloserDirectMatch :1808", T,
     [{"title": "Horween Brown Chromexcel Horsehide brwn", "points": 100}, {"title": "The Chromexcel For Brown", "points": 100}],
     "brown chromexcel", ["brwn chromexcel"], dict(num_typos=2, prefix=True, drop_tokens_threshold=0), [0, 1], "ne"),
    ("SynonymDirectMatchOutrankDirectMatch :1851", T,
     [{"title": "Marketing Officer", "points": 100}, {"title": "chief Marketing really very extremely amazingly far Officer", "points": 100}],
     "marketing officer", ["chief marketing officer"], dict(num_typos=0, prefix=True, drop_tokens_threshold=0), [0, 1], "ne"),
    ("DemoteSynonymMatch :1882", T, [{"title": "cmo", "points": 100}, {"title": "chief Marketing Officer", "points": 100}],
     "cmo", ["chief marketing officer"], dict(num_typos=0, prefix=True, drop_tokens_threshold=0, typo_tokens_threshold=40, demote=True), [0, 1], "gt"),
]


def run_cases(make_backend):
    for name, fields, docs, q, syns, opts, expect, rel in CASES:
        coll = refflow.Collection(docs, fields)
        backend, close = make_backend(coll)
        kw = dict(opts)
        demote = kw.pop("demote", False)
        s = tf.TypoSearcher(backend, coll, SORT, **kw)
        got, found = s.search(q, synonyms=syns, demote_synonym_match=demote)
        close()
        assert got == expect and found == len(expect), (name, got)
        if rel:
            a, b = s.best[got[0]][0], s.best[got[1]][0]
            assert {"eq": a == b, "ne": a != b, "gt": a > b}[rel], (name, a, b)


def test_synonym_scenarios_oracle():
    def mk(coll):
        oi = ol.OracleIndex(coll.n_docs, coll.flats, [coll.points])
        return (lambda b, k: oi.keyword_search(b, k)), (lambda: None)
    run_cases(mk)


def test_synonym_scenarios_device_functions():
    import test_hostsim as th
    hs = th.hs.__wrapped__()
    run_cases(lambda coll: (th.hostsim_backend(hs, coll), (lambda: None)))


@pytest.mark.gpu
def test_synonym_scenarios_gpu():
    from typesense_b200 import capi

    def mk(coll):
        gi = capi.GpuIndex(coll.n_docs, 0)
        for f in coll.flats:
            gi.load_field(f)
        gi.load_sort_column(coll.points)
        return (lambda b, k: gi.keyword_search(b, k)), gi.close
    run_cases(mk)
