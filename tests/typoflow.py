"""TEST INFRASTRUCTURE: the typo / prefix control flow that stays on the host in the reference, restated on top of
refflow's backends so the reference's typo and prefix scenarios (test/collection_test.cpp) can be replayed:
  Index::fuzzy_search_fields     src/index.cpp:4784-5109   cost combinations per token, candidate cache, early exits
  Index::search_all_candidates   src/index.cpp:1794-1894   product of candidates, combination limit, qhash de-dup
  Index::next_suggestion2        src/index.cpp:7204-7248   total_cost = sum(2*typo + prefix-found)
  Index::get_bounded_typo_cost   src/index.cpp:6923-6951
  drop-tokens loop               src/index.cpp:3920-4017
Candidate generation is the reference's ART walk (src/art.cpp: art_fuzzy_search_i): field_candidates() calls
typesense_b200/host/art_mirror.hpp (pinned on the reference's compiled art.cpp, tests/test_art_mirror.py; the same code the
C++ host layer uses) through tests/cpp/libartmirror.so. field_candidates_scan() is the earlier stand-in, kept for
comparison — a brute-force scan of the vocabulary that agrees with the walk on every replayed scenario but accepts a
superset in corner cases the walk's pruning rules skip:
optimal-string-alignment distance exactly equal to the
cost, prefix rule of fuzzy_search_state, leaves ordered by frequency / max_score (ties: token order), the exact leaf
first, at most max_candidates; fields are scanned in query_by order with one shared set of already-produced tokens. For the
last token of a multi-token query the reference first looks only at the fields that hold the previous token (most
documents first) and only at leaves sharing a document with it (validate_and_add_leaf, src/art.cpp:1004-1046), then
falls back to all fields; that is restated too."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

import refflow
from typesense_b200 import structs as S

FREQUENCY, MAX_SCORE = 0, 1

_ART = None


def art_lib():
    """tests/cpp/libartmirror.so: C entry points around typesense_b200/host/art_mirror.hpp (built on demand with g++)."""
    global _ART
    if _ART is None:
        import ctypes as C
        import os
        import subprocess
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        so = os.path.join(root, "tests", "cpp", "libartmirror.so")
        deps = [os.path.join(root, "tests", "cpp", "art_mirror_capi.cpp"), os.path.join(root, "typesense_b200", "host", "art_mirror.hpp"),
                os.path.join(root, "typesense_b200", "csrc", "art_device.cuh")]
        if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", deps[0], "-o", so])
        L = C.CDLL(so)
        L.am_build.restype = C.c_void_p
        L.am_build.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.c_uint32]
        L.am_bind.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.am_free.argtypes = [C.c_void_p]
        L.am_fuzzy_ex.restype = C.c_size_t
        L.am_fuzzy_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t,
                                  C.c_char_p, C.c_size_t]
        _ART = L
    return _ART


def osa_rows(term: str, key: str):
    """rows of the incremental distance matrix: row i = costs after key[:i] against every prefix of term"""
    prev2, prev = None, list(range(len(term) + 1))
    rows = [prev]
    for i, c in enumerate(key):
        cur = [prev[0] + 1]
        for col in range(1, len(term) + 1):
            cost = 0 if c == term[col - 1] else 1
            v = min(cur[col - 1] + 1, prev[col] + 1, prev[col - 1] + cost)
            if i > 1 and col > 1 and c == term[col - 2] and key[i - 1] == term[col - 1]:      # src/art.cpp:1429 (depth > 1)
                v = min(v, prev2[col - 2] + 1)
            cur.append(v)
        prev2, prev = prev, cur
        rows.append(cur)
    return rows


def matches(term: str, key: str, cost: int, prefix: bool) -> bool:
    rows = osa_rows(term, key)
    q = len(term)
    if prefix:
        # fuzzy_search_state case b): once the key is at least as long as the query, a cost within bounds accepts
        for klen in range(q, len(key) + 1):
            if rows[klen][q] == cost:
                return True
    # (the "q=strawberries on key=strawberry" special case of src/art.cpp:1511-1516 needs min_cost <= max_cost - 1 and can
    # never fire here: fuzzy_search_fields always searches with min_cost == max_cost)
    return rows[len(key)][q] == cost


class TypoSearcher:
    def __init__(self, backend, coll: refflow.Collection, sort, num_typos: int = 2, token_order: int = FREQUENCY, prefix: bool = True,
                 drop_tokens_threshold: int = 1, typo_tokens_threshold: int = 1, max_candidates: int = 4, min_len_1typo: int = 4,
                 min_len_2typo: int = 7, topster: int = 250, field_weights=None,
                 flags: int = S.FLAG_PRIORITIZE_EXACT_MATCH | S.FLAG_PRIORITIZE_NUM_MATCHING_FIELDS, match_type: int = S.MATCH_MAX_SCORE,
                 drop_tokens_mode: str = "right_to_left"):
        self.backend, self.coll, self.sort = backend, coll, sort
        self.num_typos, self.token_order, self.prefix = num_typos, token_order, prefix
        self.drop_thr, self.typo_thr, self.max_cand = drop_tokens_threshold, typo_tokens_threshold, max_candidates
        self.min1, self.min2 = min_len_1typo, min_len_2typo
        self.K = max(1, min(max(topster, 250), coll.n_docs))
        self.F = len(coll.fields)
        self.weights = list(field_weights) if field_weights else [max(0, 15 - f) for f in range(self.F)]
        self.flags = flags
        self.match_type = match_type
        self.drop_mode = drop_tokens_mode          # right_to_left | left_to_right | both_sides:N (src/index.cpp:3920-4017)
        self.freq, self.max_score = [], []
        for vocab, fl in zip(coll.vocabs, coll.flats):
            df = np.diff(fl.list_off.astype(np.int64))
            self.freq.append({t: int(df[l]) for t, l in vocab.items()})
            self.max_score.append({t: int(max(coll.points[int(i)] for i in fl.ids[int(fl.list_off[l]):int(fl.list_off[l + 1])]))
                                   for t, l in vocab.items()})

    def bounded_cost(self, token: str) -> int:
        if any(not ch.isalnum() for ch in token) or token.isdigit():
            return 0
        if len(token) < self.min1:
            return 0
        return min(self.num_typos, 1) if len(token) < self.min2 else min(self.num_typos, 2)

    def ids_of(self, f: int, t: str):
        fl, l = self.coll.flats[f], self.coll.vocabs[f][t]
        return fl.ids[int(fl.list_off[l]):int(fl.list_off[l + 1])]

    def __del__(self):
        for h, *_ in getattr(self, "_arts", {}).values():
            try:
                art_lib().am_free(h)
            except Exception:
                pass

    def art_of(self, f: int):
        """the field's ART mirror (built from the vocabulary: max_score = best points of the token's documents)"""
        import ctypes as C
        if not hasattr(self, "_arts"):
            self._arts = {}
        if f not in self._arts:
            L = art_lib()
            vocab, fl = self.coll.vocabs[f], self.coll.flats[f]
            toks = sorted(vocab, key=vocab.get)
            df = np.ascontiguousarray(np.diff(fl.list_off.astype(np.int64)), np.uint32)
            ms = np.asarray([self.max_score[f][t] for t in toks], np.int64)
            blob = "\n".join(toks).encode()
            h = L.am_build(blob, ms.ctypes.data_as(C.POINTER(C.c_int64)), df.ctypes.data_as(C.POINTER(C.c_uint32)), len(toks))
            lo = np.ascontiguousarray(fl.list_off, np.uint64)
            ids = np.ascontiguousarray(fl.ids, np.uint32)
            L.am_bind(h, blob, lo.ctypes.data_as(C.POINTER(C.c_uint64)), ids.ctypes.data_as(C.POINTER(C.c_uint32)))
            self._arts[f] = (h, lo, ids, blob)
        return self._arts[f][0]

    def field_candidates(self, f: int, token: str, cost: int, prefix_search: bool, unique_tokens: set, prev_token: str = "") -> List[str]:
        """art_fuzzy_search_i on the field's ART mirror (typesense_b200/host/art_mirror.hpp, pinned on the reference's compiled
        art.cpp): the candidates, and unique_tokens grown by every leaf the search collected."""
        import ctypes as C
        if not self.coll.vocabs[f] or any(ord(ch) > 127 or ch == "\n" for ch in token):
            return self.field_candidates_scan(f, token, cost, prefix_search, unique_tokens, prev_token)
        L = art_lib()
        out, excl = C.create_string_buffer(1 << 16), C.create_string_buffer(1 << 18)
        L.am_fuzzy_ex(self.art_of(f), token.encode(), cost, cost, self.max_cand, 1 if self.token_order == MAX_SCORE else 0, 1 if prefix_search else 0,
                      (prev_token or "").encode(), "\n".join(sorted(unique_tokens)).encode(), out, len(out), excl, len(excl))
        unique_tokens.clear()
        unique_tokens.update(x for x in excl.value.decode().split("\n") if x)
        return [x for x in out.value.decode().split("\n") if x]

    def field_candidates_scan(self, f: int, token: str, cost: int, prefix_search: bool, unique_tokens: set, prev_token: str = "") -> List[str]:
        """The earlier stand-in, kept for comparison: a brute-force scan of the vocabulary with a plain distance test."""
        vocab = self.coll.vocabs[f]
        rank = self.freq[f] if self.token_order == FREQUENCY else self.max_score[f]
        exact = token if token in vocab else None
        found = [t for t in vocab if matches(token, t, cost, prefix_search) and t not in unique_tokens and t != exact]
        if prev_token and prev_token in vocab:             # leaves that share a document with the previous token in this field
            prev_ids = self.ids_of(f, prev_token)
            found = [t for t in found if len(np.intersect1d(prev_ids, self.ids_of(f, t), assume_unique=True))]
        found.sort(key=lambda t: (-rank[t], t))
        for t in found:
            unique_tokens.add(t)
        if exact is not None and cost == 0 and exact not in unique_tokens:
            found.insert(0, exact)
            unique_tokens.add(exact)
        return found[:self.max_cand]

    def candidates(self, token: str, cost: int, prefix_search: bool, unique_tokens: set, prev_token: str = None):
        """prev_token None: not the last token. Returns None when the last token has no field to look in (the reference
        abandons this cost combination)."""
        out: List[str] = []
        if prev_token is not None:
            popular = sorted([f for f in range(self.F) if prev_token in self.coll.vocabs[f]], key=lambda f: -self.freq[f][prev_token])
            if not popular:
                return None
            for f in popular:
                out += self.field_candidates(f, token, cost, prefix_search, unique_tokens, prev_token)
                if len(out) >= self.max_cand:
                    return out
            if self.F > 1 and len(out) < self.max_cand:
                for f in range(self.F):
                    out += self.field_candidates(f, token, cost, prefix_search, unique_tokens)
                    if len(out) >= self.max_cand:
                        break
            return out
        for f in range(self.F):
            out += self.field_candidates(f, token, cost, prefix_search, unique_tokens)
            if len(out) >= self.max_cand:
                break
        return out

    def search(self, q: str, synonyms=(), demote_synonym_match: bool = False):
        """synonyms: token lists the SynonymIndex resolved for the query (host work; given as input here)."""
        tokens = refflow.tokenize(q)
        self.best: Dict[int, tuple] = {}
        self.all_ids = set()
        self.query_hashes = set()
        syn_lists = [[t for w in syn for t in refflow.tokenize(w)] if isinstance(syn, (list, tuple)) else refflow.tokenize(syn) for syn in synonyms]
        # syn_orig_num_tokens (src/index.cpp:3780-3828): -1 without synonyms, else the longest of the query and its synonyms
        self.syn_orig = max([len(tokens)] + [len(x) for x in syn_lists]) if syn_lists else -1
        self.orig_num = len(tokens)
        self.demote = demote_synonym_match
        self.is_syn = False
        all_queries = [tokens] + syn_lists
        is_prefix = [self.prefix and i == len(tokens) - 1 for i in range(len(tokens))]
        self.fuzzy(list(zip(tokens, is_prefix)), [])
        # do_synonym_search (src/index.cpp:6088-6142): every synonym as its own query, no typos, typo_tokens_threshold 0
        saved = (self.num_typos, self.typo_thr)
        for syn in syn_lists:
            self.query_hashes = set()
            self.is_syn = True
            self.num_typos, self.typo_thr = 0, 0
            self.fuzzy([(t, self.prefix and i == len(syn) - 1) for i, t in enumerate(syn)], [])
        self.num_typos, self.typo_thr = saved
        self.is_syn = False
        for qi, toks_q in enumerate(all_queries):
            self.drop_tokens_rounds(toks_q, qi > 0)
        order = sorted(self.best.items(), key=lambda kv_: (kv_[1], kv_[0]), reverse=True)
        return [k for k, _ in order], len(self.all_ids)

    def drop_tokens_rounds(self, tokens, is_synonym_variant):
        is_prefix = [self.prefix and i == len(tokens) - 1 for i in range(len(tokens))]
        # the drop-token rounds of a synonym variant run like an ordinary query (syn_orig_num_tokens -1, src/index.cpp:4010)
        saved_syn = self.syn_orig
        self.syn_orig = -1
        self.orig_num = len(tokens)
        n = min(len(tokens), 20)
        if len(self.all_ids) < self.drop_thr:
            n_dropped, dirs_done = 0, 0
            direction, both = self.drop_mode, False
            if direction.startswith("both_sides"):
                if n <= int(direction.split(":")[1]):
                    both = True                      # every truncation of both directions runs, whatever the threshold
                    direction = "both_sides"
                else:
                    direction = "right_to_left"
            while len(self.all_ids) < self.drop_thr or both:
                if n_dropped >= n - 1:
                    direction = "left_to_right" if direction == "right_to_left" else "right_to_left"
                    n_dropped = 0
                    dirs_done += 1
                rtl = direction == "right_to_left"
                if n > 1 and dirs_done < 2:
                    toks = list(zip(tokens, is_prefix))
                    if rtl:
                        tl = n - n_dropped - 1
                        trunc, dropped = toks[:tl], toks[tl:n]
                    else:
                        st = n_dropped + 1
                        trunc, dropped = toks[st:n], toks[:st]
                    n_dropped += 1
                    self.orig_num = len(trunc)
                    self.fuzzy(trunc, [t for t, _ in dropped])
                else:
                    break
        self.syn_orig = saved_syn

    # ---- Index::fuzzy_search_fields
    def fuzzy(self, qtokens, dropped: List[str]):
        if not qtokens:
            return
        token_to_costs = [list(range(0, self.bounded_cost(t) + 1)) for t, _ in qtokens]
        cache: Dict[str, List[str]] = {}
        n = 0
        N = 1
        for c in token_to_costs:
            N *= len(c)
        while n < N and n < 10:                                   # COMBINATION_MIN_LIMIT (not exhaustive)
            costs, qn = [0] * len(qtokens), n
            for i in range(len(qtokens) - 1, -1, -1):
                qn, rem = divmod(qn, len(token_to_costs[i]))
                costs[i] = token_to_costs[i][rem]
            unique_tokens: set = set()
            cands = []
            restart = False
            abandon = False
            for ti, (token, pref) in enumerate(qtokens):
                key = token + str(costs[ti])
                if key in cache:
                    leaf_tokens = cache[key]
                else:
                    leaf_tokens = []
                    if costs[ti] <= self.num_typos:
                        last_token = len(qtokens) > 1 and not dropped and ti == len(qtokens) - 1
                        leaf_tokens = self.candidates(token, costs[ti], pref, unique_tokens, cands[-1][3][0] if last_token else None)
                        if leaf_tokens is None:           # popular_field_ids.empty(): `break` out of the token loop
                            abandon = True
                            break
                        if leaf_tokens:
                            cache[key] = leaf_tokens
                if leaf_tokens:
                    cands.append((token, costs[ti], pref, leaf_tokens))
                else:
                    if costs[ti] in token_to_costs[ti]:
                        token_to_costs[ti].remove(costs[ti])
                        if not token_to_costs[ti]:
                            return
                    n = -1
                    N = 1
                    for c in token_to_costs:
                        N *= len(c)
                    restart = True
                    break
            if not restart and not abandon and len(cands) == len(qtokens):
                self.search_all_candidates(cands, dropped)
            if len(self.all_ids) >= self.typo_thr:
                return
            n += 1

    # ---- Index::search_all_candidates + next_suggestion2
    def search_all_candidates(self, cands, dropped: List[str]):
        N = 1
        for c in cands:
            N *= len(c[3])
        limit = self.max_cand if (self.F == 1 and self.prefix) else max(10, self.max_cand)
        vocabs = self.coll.vocabs
        combos = []
        for n in range(min(N, limit)):
            qn, total_cost, sugg = n, 0, []
            for token, cost, pref, leaves in cands:
                qn, rem = divmod(qn, len(leaves))
                cand = leaves[rem]
                is_prefix_searched = pref and len(cand) > len(token) + cost
                total_cost += 2 * cost + (1 if is_prefix_searched else 0)
                sugg.append(cand)
            h = tuple(sugg)
            if h in self.query_hashes:
                continue
            self.query_hashes.add(h)
            rows = [[v.get(t, S.NO_LIST) for v in vocabs] for t in sugg] + [[v.get(t, S.NO_LIST) for v in vocabs] for t in dropped]
            combos.append(S.Combo(rows, len(sugg), total_cost=total_cost, syn_orig_num_tokens=self.syn_orig, orig_num_tokens=self.orig_num,
                                  flags=(S.CFLAG_SYNONYM if self.is_syn else 0) | (S.CFLAG_DEMOTE_SYNONYM if self.demote else 0)))
        if not combos:
            return
        query = S.Query(combos, topk=self.K, sort=self.sort, num_query_tokens=len(cands), field_weight=self.weights, flags=self.flags, match_type=self.match_type)
        kv, cnt, found = self.backend(S.KwBatch([query], list(range(self.F))), self.K)
        for i in range(int(cnt[0])):
            key = int(kv["key"][0, i])
            tup = tuple(int(x) for x in kv["scores"][0, i])
            self.all_ids.add(key)
            if key not in self.best or tup >= self.best[key]:
                self.best[key] = tup
