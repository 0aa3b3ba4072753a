"""Writes tests/golden/reference_kat.json: known-answer vectors lifted verbatim from the reference's OWN unit tests
(literal tables in /root/reference/test/*.cpp; each entry cites file:line). Nothing is computed here — the script only
records the literals so the fixture has a committed generator. Byte copies of the reference's test DATA files (not code) that
its tests read: topster_record_values.txt = test/resources/record_values.txt (test/topster_test.cpp:60-136);
documents.jsonl, multi_field_documents.jsonl, float_documents.jsonl = test/*.jsonl (collection_test.cpp,
collection_sorting_test.cpp fixtures); art_skus.txt, art_ill.txt = test/skus.txt, test/ill.txt (test/art_test.cpp:914, :964).
`cp /root/reference/test/<name> tests/golden/` regenerates them."""
import json, os

kat = {
  "posting_intersect": [  # test/posting_list_test.cpp:603-700 (IntersectionBasics), block size 2
    {"src": "test/posting_list_test.cpp:603-652", "lists": [[0, 2, 3, 20], [1, 3, 5, 10, 20], [2, 3, 5, 7, 20]], "expect": [3, 20]},
    {"src": "test/posting_list_test.cpp:654-662", "lists": [[0, 2, 3, 20]], "expect": [0, 2, 3, 20]},
    {"src": "test/posting_list_test.cpp:1295-1328 (BlockIntersectionOnMixedLists: compact list x full list)", "lists": [[5, 6, 7, 8], [0, 5, 8, 20]], "expect": [5, 8]},
    {"src": "test/posting_list_test.cpp:774-823 (IntersectionSkipBlocks)", "lists": [[9, 11], [1, 2, 3, 4, 5, 6, 7, 8, 9, 11], [2, 3, 8, 9, 11, 20]], "expect": [9, 11]},
  ],
  "posting_merge": [      # test/posting_list_test.cpp:559-601 (MergeBasics)
    {"src": "test/posting_list_test.cpp:559-601", "lists": [[0, 2, 3, 20], [1, 3, 5, 10, 20], [2, 3, 5, 7, 20]], "expect": [0, 1, 2, 3, 5, 7, 10, 20]},
  ],
  "or_iterator": [
    {"src": "test/or_iterator_test.cpp:8-82", "offsets": [0, 1, 3],
     "tokens": [[[0, 2, 3, 20], [1, 3, 5, 10, 20], [2, 3, 6, 7, 20]], [[0, 1, 5, 20], [1, 2, 7, 11, 15], [3, 5, 10, 11, 12]]],
     "filter": None, "expect": [0, 1, 2, 3, 5, 7, 10, 20]},
    {"src": "test/or_iterator_test.cpp:162-217 (IntersectAndFilterThreeIts)", "offsets": [0, 1, 3],
     "tokens": [[[4207, 29159, 47182, 47250, 47337, 48518, 99820]],
                [[62, 330, 367, 4124, 4207, 4242, 4418, 28740, 29099, 29159, 29284, 40795, 43556, 46779, 47182, 47250, 47322, 48494, 48518, 48633, 98813, 98821, 99069, 99368, 99533, 99670, 99820, 99888, 99973]],
                [[723, 1504, 29038, 29164, 29390, 30890, 34743, 35067, 36466, 40268, 40965, 42161, 43425, 45188, 47326, 47443, 49319, 53043, 58436, 58774, 61123, 70973, 71393, 81575, 82323, 88301, 88502, 88594, 88690, 88951, 90662, 91016, 91915, 92069, 92844, 99820]]],
     "filter": [44424, 44425, 44447, 99820, 99834, 99854, 99859, 99963], "expect": [99820]},
    {"src": "test/or_iterator_test.cpp:219-264 (IntersectAndFilterTwoIts)", "offsets": [0, 1, 3],
     "tokens": [[[4207, 29159, 47182, 47250, 47337, 48518, 99820]],
                [[62, 330, 367, 4124, 4207, 4242, 4418, 28740, 29099, 29159, 29284, 40795, 43556, 46779, 47182, 47250, 47322, 48494, 48518, 48633, 98813, 98821, 99069, 99368, 99533, 99670, 99820, 99888, 99973]]],
     "filter": [44424, 44425, 44447, 99820, 99834, 99854, 99859, 99963], "expect": [99820]},
  ],
  "match": [  # test/match_score_test.cpp; positions per token, last_token flag, check_exact -> words_present, distance, exact
    {"src": "test/match_score_test.cpp:9-28", "tokens": [[1]] * 12, "last": [0] * 12, "check_exact": 0, "words_present": 10},
    {"src": "test/match_score_test.cpp:30-50", "tokens": [[25], [26], [11, 18, 24, 60], [14, 27, 63]], "last": [0, 0, 0, 0], "check_exact": 0, "words_present": 4, "distance": 3, "phrase": False},
    {"src": "test/match_score_test.cpp:52-61", "tokens": [[38, 50, 170, 187, 195, 222], [39, 140, 171, 189, 223], [169, 180]], "last": [0, 1, 0], "check_exact": 1, "words_present": 3, "distance": 2, "exact": 0, "phrase": False},
    {"src": "test/match_score_test.cpp:68-78", "tokens": [[38, 50, 187, 195, 201], [120, 167, 171, 223], [240, 250]], "last": [0, 0, 1], "check_exact": 0, "words_present": 1, "distance": 0, "exact": 0, "phrase": False},
    {"src": "test/match_score_test.cpp:92-102", "tokens": [[0], [2], [1]], "last": [0, 1, 0], "check_exact": 1, "words_present": 3, "distance": 2, "exact": 1, "phrase": False},
    {"src": "test/match_score_test.cpp:104-108", "tokens": [[0], [2], [1]], "last": [0, 1, 0], "check_exact": 0, "words_present": 3, "distance": 2, "exact": 0},
    {"src": "test/match_score_test.cpp:110-117", "tokens": [[1], [2], [3]], "last": [0, 0, 1], "check_exact": 1, "exact": 0, "phrase": True},
    {"src": "test/match_score_test.cpp:119-126", "tokens": [[0], [1], [2]], "last": [0, 0, 0], "check_exact": 1, "exact": 0, "phrase": True},
    {"src": "test/match_score_test.cpp:141-148", "tokens": [[38, 50, 187, 195, 201], [120, 167, 171, 196], [197, 250]], "last": [0, 0, 1], "check_exact": 0, "phrase": True},
    {"src": "test/match_score_test.cpp:150-156", "tokens": [[120, 167, 171, 196], [38, 50, 187, 195, 201], [197, 250]], "last": [0, 0, 1], "check_exact": 0, "phrase": False},
  ],
  "topster_max_int": {  # test/topster_test.cpp:7-58
    "src": "test/topster_test.cpp:7-58", "capacity": 5,
    "rows": [[0, 1, 11, 20, 30], [0, 1, 12, 20, 32], [0, 2, 4, 20, 30], [2, 3, 7, 20, 30], [0, 4, 14, 20, 30], [1, 5, 9, 20, 30],
             [1, 5, 10, 20, 32], [1, 5, 9, 20, 30], [0, 6, 6, 20, 30], [2, 7, 6, 22, 30], [2, 7, 6, 22, 30], [1, 8, 9, 20, 30],
             [0, 9, 8, 20, 30], [3, 10, 5, 20, 30]],
    "expect_keys": [4, 1, 5, 8, 9], "expect_score_of": {"1": 12, "5": 10}},
  "vector_cosine": {  # test/collection_vector_search_test.cpp:75-122 (BasicVectorQuerying): cosine, d=4, wildcard query
    "src": "test/collection_vector_search_test.cpp:75-122", "metric": "cosine",
    "docs": [[0.851758, 0.909671, 0.823431, 0.372063], [0.97826, 0.933157, 0.39557, 0.306488], [0.230606, 0.634397, 0.514009, 0.399594]],
    "query": [0.96826, 0.94, 0.39557, 0.306488],
    "expect_ids": [1, 0, 2],
    "expect_distances": [3.409385681152344e-05, 0.04329806566238403, 0.15141665935516357],
    "filtered": {"filter_ids": [0, 1], "expect_ids": [1, 0]}},
  "array_utils": [  # test/array_utils_test.cpp:5-176 (AndScalar, OrScalar*, FilterArray); op 0 and, 1 or, 2 exclude
    {"src": "test/array_utils_test.cpp:5-37", "op": 0, "a": list(range(9)), "b": [3, 6, 9], "expect": [3, 6]},
    {"src": "test/array_utils_test.cpp:39-71", "op": 1, "a": list(range(9)), "b": [3, 6, 9], "expect": list(range(10))},
    {"src": "test/array_utils_test.cpp:73-98", "op": 1, "a": list(range(9)), "b": [0, 4, 5], "expect": list(range(9))},
    {"src": "test/array_utils_test.cpp:100-118", "op": 1, "a": list(range(9)), "b": [], "expect": list(range(9))},
    {"src": "test/array_utils_test.cpp:100-118", "op": 1, "a": [], "b": list(range(9)), "expect": list(range(9))},
    {"src": "test/array_utils_test.cpp:120-144", "op": 2, "a": list(range(9)), "b": [0, 1, 5, 7, 8], "expect": [2, 3, 4, 6]},
    {"src": "test/array_utils_test.cpp:146-158", "op": 2, "a": list(range(9)), "b": list(range(9)), "expect": []},
    {"src": "test/array_utils_test.cpp:163-172", "op": 2, "a": [58, 118, 185, 260, 322, 334, 353],
     "b": [58, 103, 116, 117, 137, 154, 191, 210, 211, 284, 299, 302, 306, 309, 332, 334, 360], "expect": [118, 185, 260, 322, 353]},
  ],
  "text_match_layout": {  # test/union_test.cpp:810: single token, one field, weight 15, cost 0
    "src": "test/union_test.cpp:810", "value": 578730123365189753},
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json"), "w") as f:
    json.dump(kat, f, indent=1)
