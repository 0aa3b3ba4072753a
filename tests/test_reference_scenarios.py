"""End-to-end ranking scenarios of the reference's own test-suite (test/collection_test.cpp over test/documents.jsonl),
replayed through the oracle on the CPU and — with -m gpu — through libtsgpu. These pin intersection + Match + score
packing + sort keys + Topster order + the drop-tokens flow together."""
import os

import pytest

import oracle_lib as ol
import refflow
from typesense_b200 import structs as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SORT_DESC = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
SORT_POINTS_ASC_2ND = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, -1, 0), (S.SORT_NONE, -1, 1, 0))
# sort_by points:asc alone: _text_match is appended because the list has < 3 entries (src/collection.cpp:1736-1812)
SORT_POINTS_ASC = ((S.SORT_NUMERIC, 0, -1, 0), (S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NONE, -1, 1, 0))


def ids_of(coll, seq_ids):
    return [str(coll.docs[s].get("id", s)) for s in seq_ids]


def scenarios(backend, coll):
    # MultiTokenSearch, test/collection_test.cpp:162-236 (drop_tokens_threshold = 10)
    got, found = refflow.search(backend, coll, "rocket launch", SORT_DESC, drop_tokens_threshold=10)
    assert ids_of(coll, got) == ["8", "1", "17", "16", "13"] and found == 5
    got, found = refflow.search(backend, coll, "rocket launch", SORT_POINTS_ASC_2ND, drop_tokens_threshold=10)
    assert ids_of(coll, got) == ["8", "17", "1", "16", "13"] and found == 5
    # ExactSearchShouldBeStable, test/collection_test.cpp:117-160
    got, found = refflow.search(backend, coll, "the", SORT_DESC)
    assert ids_of(coll, got) == ["1", "6", "foo", "13", "10", "8", "16"] and found == 7
    got, found = refflow.search(backend, coll, "the", SORT_POINTS_ASC)
    assert ids_of(coll, got) == ["16", "13", "10", "8", "6", "foo", "1"] and found == 7
    got, found = refflow.search(backend, coll, "zxsadqewsad", SORT_POINTS_ASC)
    assert got == [] and found == 0


def more_scenarios(backend, coll, wildcard_backend=None):
    """num_typos = 0 cases of PartialMultiTokenSearch, SkipUnindexedTokensDuringMultiTokenSearch and
    SearchWithExcludedTokens (test/collection_test.cpp:238-372): drop-tokens flow in both directions, tokens missing
    from the index, exclusion tokens."""
    # PartialMultiTokenSearch :358-372
    got, found = refflow.search(backend, coll, "rocket research", SORT_DESC, drop_tokens_threshold=10)
    assert ids_of(coll, got) == ["19", "1", "10", "8", "16", "17"]
    # SkipUnindexedTokensDuringMultiTokenSearch :269-356
    got, found = refflow.search(backend, coll, "DoesNotExist from", SORT_DESC)
    assert ids_of(coll, got) == ["2", "17"]
    got, found = refflow.search(backend, coll, "the a", SORT_DESC, drop_tokens_threshold=10)
    assert len(got) == 9
    got, found = refflow.search(backend, coll, "the a", SORT_DESC, drop_tokens_threshold=0)
    assert ids_of(coll, got) == ["8", "16", "10"]
    got, found = refflow.search(backend, coll, "the a insurance", SORT_DESC, drop_tokens_threshold=0)
    assert got == []
    got, found = refflow.search(backend, coll, "DoesNotExist1 DoesNotExist2", SORT_DESC)
    assert got == []
    # SearchWithExcludedTokens :238-267
    got, found = refflow.search(backend, coll, "how -propellants -are", SORT_DESC, drop_tokens_threshold=10)
    assert ids_of(coll, got) == ["9", "17"] and found == 2
    if wildcard_backend is not None:
        # 23 documents + the fixture's dummy record 0 (test/collection_test.cpp:53-55), which matches `*` too
        got, found = refflow.search(backend, coll, "-rocket", SORT_DESC, wildcard_backend=wildcard_backend)
        assert found == 21 and len(got) == 21
        got, found = refflow.search(backend, coll, "-rocket -cryovolcanism", SORT_DESC, wildcard_backend=wildcard_backend)
        assert found == 20


def multi_field_scenarios(make_backend):
    # MultiFieldRelevance, test/collection_test.cpp:3173-3258: title + artist, default weights 15/14, drop tokens <= 10
    q = "Dustin Kensrue Down There by the Train"
    for records, expect in (([("Down There by the Train", "Dustin Kensrue"), ("Down There by the Train", "Gord Downie"),
                              ("State Trooper", "Dustin Kensrue")], [0, 1, 2]),
                            ([("State Trooper", "Dustin Kensrue"), ("Down There by the Train", "Gord Downie"),
                              ("Down There by the Train", "Dustin Kensrue")], [2, 1, 0])):
        coll = refflow.Collection([{"title": t, "artist": a, "points": i} for i, (t, a) in enumerate(records)], ("title", "artist"))
        backend, close = make_backend(coll)
        got, found = refflow.search(backend, coll, q, SORT_DESC, drop_tokens_threshold=10)
        close()
        assert got == expect and found == 3


def exact_match_scenario(make_backend):
    # ExactMatch, test/collection_test.cpp:3638-3688 (query_by title only; typos/prefix find no other candidate in this
    # vocabulary, so the flow is the exact tokens and then the drop-tokens round)
    records = [("Alpha", "DJ"), ("Alpha Beta", "DJ"), ("Alpha Beta Gamma", "DJ")]
    coll = refflow.Collection([{"title": t, "artist": a, "points": i} for i, (t, a) in enumerate(records)], ("title",))
    backend, close = make_backend(coll)
    got, found = refflow.search(backend, coll, "alpha beta", SORT_DESC, drop_tokens_threshold=10)
    assert got == [1, 2, 0] and found == 3
    got, found = refflow.search(backend, coll, "alpha", SORT_DESC, drop_tokens_threshold=10)
    assert got == [0, 2, 1] and found == 3
    close()


def ranked_weights(given):
    """Collection::process_search_field_weights (src/collection.cpp:4210-4275): weights already in descending order and
    <= 15 are used as they are; otherwise they are re-ranked into 15, 14, ... preserving ties."""
    if all(given[i] <= given[i - 1] for i in range(1, len(given))) and all(w <= 15 for w in given):
        return list(given)
    order = sorted(set(given), reverse=True)
    return [15 - order.index(w) for w in given]


def match_ranking_scenarios(make_backend):
    # MultiFieldMatchRanking, test/collection_test.cpp:3788-3835: query_by artist,title; drop_tokens_threshold 5
    titles = ["Style", "Blank Space", "Balance Overkill", "Cardigan", "Invisible String", "The Last Great American Dynasty",
              "Mirrorball", "Peace", "Betty", "Mad Woman"]
    coll = refflow.Collection([{"title": t, "artist": "Taylor Swift", "points": i} for i, t in enumerate(titles)], ("artist", "title"))
    backend, close = make_backend(coll)
    got, found = refflow.search(backend, coll, "taylor swift style", SORT_DESC, drop_tokens_threshold=5)
    close()
    assert got[:3] == [0, 9, 8] and found == 10
    # MultiFieldMatchRankingOnArray :3837-3877: two string[] fields, drop_tokens_threshold 1
    recs = [(["Golang", "Vue", "React"], ["Docker", "Goa", "Elixir"]), (["Golang", "Phoenix", "React"], ["Docker", "Vue", "Kubernetes"])]
    coll = refflow.Collection([{"strong_skills": a, "skills": b, "points": i} for i, (a, b) in enumerate(recs)], ("strong_skills", "skills"))
    backend, close = make_backend(coll)
    got, found = refflow.search(backend, coll, "golang vue", SORT_DESC, drop_tokens_threshold=1)
    close()
    assert got == [0, 1] and found == 2
    # MultiFieldMatchRankingOnFieldOrder :3879-3920: query_by title,artist with query_by_weights {1, 6}
    recs = [("Toxic", "Britney Spears"), ("Bad", "Michael Jackson")]
    coll = refflow.Collection([{"title": t, "artist": a, "points": i} for i, (t, a) in enumerate(recs)], ("title", "artist"))
    backend, close = make_backend(coll)
    got, found = refflow.search(backend, coll, "michael jackson toxic", SORT_DESC, drop_tokens_threshold=5, field_weights=ranked_weights([1, 6]))
    close()
    assert got == [1, 0] and found == 2


def relevance2_scenarios(make_backend):
    # MultiFieldRelevance2, test/collection_test.cpp:3276-3355: query_by title,artist; drop_tokens_threshold 10
    recs = [("A Daikon Freestyle", "Ghosts on a Trampoline"), ("Leaving on a Jetplane", "Coby Grant")]
    coll = refflow.Collection([{"title": t, "artist": a, "points": i} for i, (t, a) in enumerate(recs)], ("title", "artist"))
    backend, close = make_backend(coll)
    for weights in (None, ranked_weights([1, 4]), ranked_weights([1, 1])):
        got, found = refflow.search(backend, coll, "on a jetplane", SORT_DESC, drop_tokens_threshold=10, field_weights=weights)
        assert got == [1, 0] and found == 2, weights
    got, found = refflow.search(backend, coll, "on a helicopter", SORT_DESC, drop_tokens_threshold=10, field_weights=ranked_weights([1, 4]))
    close()
    assert got == [0, 1] and found == 2


def relevance36_scenarios(make_backend):
    same = ranked_weights([1, 1])
    # MultiFieldRelevance3, test/collection_test.cpp:3403-3460
    recs = [("Taylor Swift Karaoke: reputation", "Taylor Swift"), ("Style", "Taylor Swift")]
    coll = refflow.Collection([{"title": t, "artist": a, "points": i} for i, (t, a) in enumerate(recs)], ("title", "artist"))
    backend, close = make_backend(coll)
    got, found = refflow.search(backend, coll, "style taylor swift", SORT_DESC, drop_tokens_threshold=10, field_weights=same)
    assert got == [1, 0] and found == 2
    got, found = refflow.search(backend, coll, "swift", SORT_DESC, drop_tokens_threshold=10, field_weights=same)
    close()
    assert got == [0, 1] and found == 2
    # MultiFieldRelevance6 :3581-3636: the number of fields with an exact match is not a ranking signal
    recs = [("Taylor Swift", "Taylor Swift"), ("Taylor Swift Song", "Taylor Swift")]
    coll = refflow.Collection([{"title": t, "artist": a, "points": i} for i, (t, a) in enumerate(recs)], ("title", "artist"))
    backend, close = make_backend(coll)
    for flags in (S.FLAG_PRIORITIZE_EXACT_MATCH | S.FLAG_PRIORITIZE_NUM_MATCHING_FIELDS, S.FLAG_PRIORITIZE_NUM_MATCHING_FIELDS):
        got, found = refflow.search(backend, coll, "taylor swift", SORT_DESC, drop_tokens_threshold=10, field_weights=same, flags=flags)
        assert got == [1, 0] and found == 2, flags
    close()


def repeating_token_scenario(make_backend):
    # RepeatingTokenRanking, test/collection_sorting_test.cpp:1800-1855: a repeated query token, with the reference's
    # literal text_match values
    recs = [("Mong Mong", 100), ("Mong Spencer", 200), ("Mong Mong Spencer", 300), ("Spencer Mong Mong", 400)]
    coll = refflow.Collection([{"title": t, "points": p} for t, p in recs], ("title",))
    backend, close = make_backend(coll)
    got, found = refflow.search(backend, coll, "mong mong", SORT_DESC, drop_tokens_threshold=10, field_weights=ranked_weights([3]))
    close()
    assert got == [0, 3, 2, 1]
    tm = {k: v[0] for k, v in refflow.search.last_scores.items()}
    assert tm[0] == 1157451471583709209 and tm[3] == tm[2] == tm[1] == 1157451471575320601


def text_match_literals_scenario(make_backend):
    # test/collection_vector_search_test.cpp:5403-5497: text_match of "nike running shoes" per document. The reference
    # reports doc 0 from the keyword pass and docs 1/2 through compute_aux_scores; the same values fall out of the
    # drop-tokens rounds ("nike running", "shoes") here. 578730123365189753 is also test/union_test.cpp:810.
    names = ["Nike running shoes for men", "Nike running sneakers", "adidas shoes", "puma"]
    coll = refflow.Collection([{"name": n, "points": 0} for n in names], ("name",))
    backend, close = make_backend(coll)
    got, found = refflow.search(backend, coll, "nike running shoes", SORT_DESC, drop_tokens_threshold=10)
    close()
    tm = {k: v[0] for k, v in refflow.search.last_scores.items()}
    assert got == [0, 1, 2] and tm == {0: 1736172819517016185, 1: 1157451471441102969, 2: 578730123365189753}


def test_multi_field_scenarios_oracle():
    def mk(coll):
        oi = ol.OracleIndex(coll.n_docs, coll.flats, [coll.points])
        return (lambda b, k: oi.keyword_search(b, k)), (lambda: None)
    multi_field_scenarios(mk)
    exact_match_scenario(mk)
    match_ranking_scenarios(mk)
    relevance2_scenarios(mk)
    relevance36_scenarios(mk)
    repeating_token_scenario(mk)
    text_match_literals_scenario(mk)


@pytest.mark.gpu
def test_multi_field_scenarios_gpu():
    from typesense_b200 import capi

    def mk(coll):
        gi = capi.GpuIndex(coll.n_docs, 0)
        for f in coll.flats:
            gi.load_field(f)
        gi.load_sort_column(coll.points)
        return (lambda b, k: gi.keyword_search(b, k)), gi.close
    multi_field_scenarios(mk)
    exact_match_scenario(mk)
    match_ranking_scenarios(mk)
    relevance2_scenarios(mk)
    relevance36_scenarios(mk)
    repeating_token_scenario(mk)
    text_match_literals_scenario(mk)


def test_reference_scenarios_oracle():
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    oi = ol.OracleIndex(coll.n_docs, [coll.flat], [coll.points])
    scenarios(lambda b, k: oi.keyword_search(b, k), coll)
    more_scenarios(lambda b, k: oi.keyword_search(b, k), coll, lambda b, k: oi.wildcard_search(b, k))


@pytest.mark.gpu
def test_reference_scenarios_gpu():
    from typesense_b200 import capi
    coll = refflow.Collection.from_jsonl(os.path.join(GOLD, "documents.jsonl"))
    gi = capi.GpuIndex(coll.n_docs, 0)
    gi.load_field(coll.flat)
    gi.load_sort_column(coll.points)
    scenarios(lambda b, k: gi.keyword_search(b, k), coll)
    more_scenarios(lambda b, k: gi.keyword_search(b, k), coll, lambda b, k: gi.wildcard_search(b, k))
    gi.close()
