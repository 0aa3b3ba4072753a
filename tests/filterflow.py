"""TEST INFRASTRUCTURE: the string `filter_by` flow of the reference restated on top of pluggable id-set primitives
(CPU oracle, the host-compiled device functions, or the tsgpu C-ABI), so the reference's own filter expectations
(test/collection_filtering_test.cpp, test/collection_specific_more_test.cpp) can be replayed against each.

  value parsing       filter::parse_filter_string_value   src/filter.cpp:674-733
  per value tokens    filter_result_iterator_t::init       src/filter_result_iterator.cpp:1739-1905
  id computation      ::compute_iterators                  src/filter_result_iterator.cpp:2964-3100
  `!=` complement     apply_not_equals                     src/filter_result_iterator.cpp:936-953

The primitives are exactly the calls of that code: posting_list_t::intersect, get_phrase_matches / get_exact_matches /
get_prefix_matches and ArrayUtils::or_scalar / exclude_scalar."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

import oracle_lib as ol
import refflow
import typoflow
from typesense_b200 import structs as S

CONTAINS, EQUALS, NOT_EQUALS, CONTAINS_PHRASE = 0, 1, 2, 3
MODE_PHRASE, MODE_EXACT, MODE_PREFIX = 1, 2, 3
SET_AND, SET_OR, SET_EXCLUDE = 0, 1, 2
MAX_FILTER_BY_CANDIDATES = 4                                   # include/filter.h:15


def split_to_values(s: str) -> List[str]:
    """StringUtils::split_to_values: comma separated, back-tick quoting keeps commas, values trimmed."""
    out, cur, tick = [], [], False
    for ch in s:
        if ch == "`":
            tick = not tick
        elif ch == "," and not tick:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return [v for v in out if v]


def parse_string_filter(raw: str):
    """-> (values, comparators, apply_not_equals) as filter.cpp:674-733 fills filter_exp."""
    i, comp, neg = 0, CONTAINS, False
    if raw[0] == "=":
        comp = EQUALS
        i = 1
    elif len(raw) >= 2 and raw[0] == "!":
        i = 1
        if raw[1] == "=":
            comp = NOT_EQUALS
            i = 2
        neg = True
    while i < len(raw) and raw[i] == " ":
        i += 1
    part = raw[i:]
    quoted = lambda v: len(v) > 1 and v[0] == '"' and v[-1] == '"'
    if quoted(part):
        return [part[1:-1]], [CONTAINS_PHRASE], neg
    if part[0] == "[" and part[-1] == "]":
        vals = split_to_values(part[1:-1])
        default = EQUALS if any(quoted(v) for v in vals) else comp
        return [v[1:-1] if quoted(v) else v for v in vals], [CONTAINS_PHRASE if quoted(v) else default for v in vals], neg
    return [part], [comp], neg


class OracleOps:
    def __init__(self, coll: refflow.Collection):
        self.L = ol.oracle()
        self.coll = coll
        self.ix = [ol.OracleIndex(coll.n_docs, [fl], []) for fl in coll.flats]

    def intersect(self, f: int, lists: Sequence[int]) -> np.ndarray:
        fl = self.coll.flats[f]
        arrs = [np.ascontiguousarray(fl.ids[int(fl.list_off[l]):int(fl.list_off[l + 1])], np.uint32) for l in lists]
        ptrs = (C.POINTER(C.c_uint32) * len(arrs))(*[ol.p32(a) for a in arrs])
        lens = (C.c_size_t * len(arrs))(*[len(a) for a in arrs])
        out = np.zeros(max(1, min(len(a) for a in arrs)), np.uint32)
        n = self.L.tso_intersect(len(arrs), ptrs, lens, ol.p32(out), C.c_size_t(len(out)))
        return out[:n]

    def matches(self, f: int, lists: Sequence[int], ids: np.ndarray, mode: int) -> np.ndarray:
        fn = {MODE_PHRASE: self.L.tso_phrase_matches, MODE_EXACT: self.L.tso_exact_matches, MODE_PREFIX: self.L.tso_prefix_matches}[mode]
        out = np.zeros(max(1, len(ids)), np.uint32)
        n = fn(self.ix[f].h, 0, ol.p32(np.asarray(lists, np.uint32)), len(lists), ol.p32(np.ascontiguousarray(ids, np.uint32)), len(ids), ol.p32(out))
        return out[:n]

    def setop(self, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        fn = (self.L.tso_and_scalar, self.L.tso_or_scalar, self.L.tso_exclude_scalar)[op]
        a, b = np.ascontiguousarray(a, np.uint32), np.ascontiguousarray(b, np.uint32)
        out = np.zeros(len(a) + len(b) + 1, np.uint32)
        n = fn(ol.p32(a if len(a) else out), len(a), ol.p32(b if len(b) else out), len(b), ol.p32(out))
        return out[:n].copy()


class HostsimOps(OracleOps):
    """phrase / exact / prefix through the DEVICE functions compiled for the host (tests/hostsim)."""

    def __init__(self, coll, hs):
        super().__init__(coll)
        self.hs = hs

    def matches(self, f, lists, ids, mode):
        fs = self.coll.flats[f].struct()
        out = np.zeros(max(1, len(ids)), np.uint32)
        n = self.hs.hs_idset_matches(C.byref(fs), ol.p32(np.asarray(lists, np.uint32)), len(lists), ol.p32(np.ascontiguousarray(ids, np.uint32)),
                                     len(ids), mode, ol.p32(out))
        return out[:n]


class CapiOps:
    """the tsgpu C-ABI (GpuIndex of typesense_b200.capi); field ids as returned by load_field."""

    def __init__(self, coll, gi, field_ids):
        self.coll, self.gi, self.fids = coll, gi, field_ids

    def intersect(self, f, lists):
        return self.gi.intersect(self.fids[f], list(lists), self.coll.n_docs)

    def matches(self, f, lists, ids, mode):
        fn = {MODE_PHRASE: self.gi.phrase_matches, MODE_EXACT: self.gi.exact_matches, MODE_PREFIX: self.gi.prefix_matches}[mode]
        return fn(self.fids[f], list(lists), np.ascontiguousarray(ids, np.uint32))

    def setop(self, op, a, b):
        return self.gi.ids_setop(op, np.ascontiguousarray(a, np.uint32), np.ascontiguousarray(b, np.uint32))


def prefix_value_token_sets(coll: refflow.Collection, f: int, toks: List[str]) -> List[List[str]]:
    """`Chris P*`: the value runs through Index::fuzzy_search_fields with num_typos 0, the last token prefix-searched,
    token order MAX_SCORE, typo_tokens_threshold 0 and max_candidates = max_filter_by_candidates
    (filter_result_iterator.cpp:1788-1830); every suggestion with a non-empty intersection becomes one OR-ed value."""
    one = refflow.Collection.__new__(refflow.Collection)
    one.__dict__.update(coll.__dict__)
    one.fields, one.vocabs, one.flats = [coll.fields[f]], [coll.vocabs[f]], [coll.flats[f]]
    one.vocab, one.flat = one.vocabs[0], one.flats[0]
    ts = typoflow.TypoSearcher(None, one, None, num_typos=0, token_order=typoflow.MAX_SCORE, prefix=True, max_candidates=MAX_FILTER_BY_CANDIDATES)
    uniq: set = set()
    cands: List[List[str]] = []
    for ti, t in enumerate(toks):
        last = ti == len(toks) - 1
        prev = cands[-1][0] if (last and len(toks) > 1) else None
        c = ts.candidates(t, 0, last, uniq, prev)
        if not c:
            return []
        cands.append(c)
    out, N = [], int(np.prod([len(c) for c in cands]))
    for n in range(min(N, MAX_FILTER_BY_CANDIDATES)):             # combination_limit: one field, prefix (src/index.cpp:1841)
        qn, sugg = n, []
        for c in cands:
            qn, rem = divmod(qn, len(c))
            sugg.append(c[rem])
        out.append(sugg)
    return out


def string_filter_ids(ops, coll: refflow.Collection, field: str, raw: str) -> List[int]:
    f = coll.fields.index(field)
    vocab = coll.vocabs[f]
    values, comps, neg = parse_string_filter(raw)
    comp0 = comps[0]                                               # compute_iterators reads comparators[0] for every value
    plists, prefix_index = [], set()
    for v in values:
        is_prefix = len(v) > 1 and v[-1] == "*"
        toks = refflow.tokenize(v[:-1] if is_prefix else v)
        assert toks, "Filter value cannot be empty."
        if is_prefix:
            for sugg in prefix_value_token_sets(coll, f, toks):
                lists = [vocab[t] for t in sugg]
                if len(ops.intersect(f, lists)):
                    prefix_index.add(len(plists))
                    plists.append(lists)
            continue
        if any(t not in vocab for t in toks):
            continue
        plists.append([vocab[t] for t in toks])
    or_ids = np.zeros(0, np.uint32)
    for i, lists in enumerate(plists):
        ids = ops.intersect(f, lists)
        if not len(ids):
            continue
        if i in prefix_index and comp0 in (EQUALS, NOT_EQUALS):
            ids = ops.matches(f, lists, ids, MODE_PREFIX)
        elif comp0 == CONTAINS_PHRASE:
            ids = ops.matches(f, lists, ids, MODE_PHRASE)
        elif comp0 in (EQUALS, NOT_EQUALS):
            ids = ops.matches(f, lists, ids, MODE_EXACT)
        if len(ids):
            or_ids = ops.setop(SET_OR, or_ids, ids)
    if neg:
        or_ids = ops.setop(SET_EXCLUDE, np.arange(coll.n_docs, dtype=np.uint32), or_ids)
    return [int(x) for x in or_ids]
