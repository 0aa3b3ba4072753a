"""TEST INFRASTRUCTURE: the host-side control flow that stays on the host in the reference — ASCII tokenising
(src/tokenizer.cpp:232-290), token lookup (art_search) and the drop-tokens loop of Index::search
(src/index.cpp:3920-4017) — restated in Python on top of a pluggable backend (CPU oracle or the tsgpu C-ABI), so the
reference's END-TO-END ranking expectations (test/collection_test.cpp) can be replayed against both."""
from __future__ import annotations

import json
from typing import Dict, List, Sequence

import numpy as np

from typesense_b200 import structs as S


def tokenize(text: str) -> List[str]:
    """alnum kept and lower-cased; space / newline split; every other ASCII char is dropped, not a boundary."""
    out, cur = [], []
    for ch in text:
        if ch.isalnum() and ord(ch) < 128:
            cur.append(ch.lower())
        elif ch in " \n":
            if cur:
                out.append("".join(cur)[:100])
            cur = []
    if cur:
        out.append("".join(cur)[:100])
    return out


class Collection:
    """String fields (default: `title`) + numeric `points`; seq_id = insertion order (for documents.jsonl a dummy doc
    first, as the reference's fixture does, so seq_id == line number)."""

    def __init__(self, docs: Sequence[dict], fields: Sequence[str] = ("title",)):
        self.docs = list(docs)
        self.fields = list(fields)
        self.vocabs: List[Dict[str, int]] = []
        self.flats: List[S.FlatField] = []
        for fname in self.fields:
            vocab: Dict[str, int] = {}
            per_tok: List[list] = []
            is_array = any(isinstance(d.get(fname), (list, tuple)) for d in self.docs)
            for sid, d in enumerate(self.docs):
                t2o: Dict[str, list] = {}
                if is_array:
                    # string[]: per element and token positions..., last position repeated, array index; the element's last
                    # token also gets 0 (src/index.cpp:1357-1393)
                    for ai, elem in enumerate(d.get(fname) or []):
                        toks = tokenize(elem)
                        seen: List[str] = []
                        for i, t in enumerate(toks):
                            t2o.setdefault(t, []).append(i + 1)
                            if t not in seen:
                                seen.append(t)
                        for t in seen:
                            t2o[t].append(t2o[t][-1])
                            t2o[t].append(ai)
                        if toks:
                            t2o[toks[-1]].append(0)
                else:
                    toks = tokenize(d.get(fname, ""))
                    for i, t in enumerate(toks):
                        t2o.setdefault(t, []).append(i + 1)
                    if toks:
                        t2o[toks[-1]].append(0)           # src/index.cpp:1341-1348
                for t, offs in t2o.items():
                    if t not in vocab:
                        vocab[t] = len(per_tok)
                        per_tok.append([])
                    per_tok[vocab[t]].append((sid, offs))
            self.vocabs.append(vocab)
            self.flats.append(S.FlatField.from_postings(per_tok, is_array))
        self.vocab, self.flat = self.vocabs[0], self.flats[0]
        self.points = np.asarray([d["points"] for d in self.docs], np.int64)
        self.n_docs = len(self.docs)

    @staticmethod
    def from_jsonl(path: str) -> "Collection":
        docs = [{"points": 10, "title": "z"}]
        docs += [json.loads(l) for l in open(path) if l.strip()]
        return Collection(docs)


def split_query(q: str):
    """`-word` is an exclusion token (Collection::parse_search_query, src/collection.cpp): returns (include, exclude)."""
    inc, exc = [], []
    for w in q.split(" "):
        if w.startswith("-") and len(w) > 1:
            exc += tokenize(w[1:])
        else:
            inc += tokenize(w)
    return inc, exc


def excluded_ids(coll: Collection, exc: List[str]) -> List[int]:
    """ids holding any exclusion token in any searched field (the host builds `excluded_result_ids` from the tokens' posting
    lists before run_search)."""
    out = set()
    for t in exc:
        for v, fl in zip(coll.vocabs, coll.flats):
            l = v.get(t)
            if l is not None:
                out.update(int(x) for x in fl.ids[int(fl.list_off[l]):int(fl.list_off[l + 1])])
    return sorted(out)


def search(backend, coll: Collection, q: str, sort, drop_tokens_threshold: int = 1, topster: int = 250, wildcard_backend=None,
           field_weights: Sequence[int] = None,
           flags: int = S.FLAG_PRIORITIZE_EXACT_MATCH | S.FLAG_PRIORITIZE_NUM_MATCHING_FIELDS):   # Collection::search defaults
    """backend(batch, stride) -> (kv, cnt, found). Returns (ordered seq_ids, found)."""
    tokens, exc = split_query(q)
    excl = excluded_ids(coll, exc)
    if not tokens and exc and wildcard_backend is not None:
        # only exclusion tokens: the query becomes `*` minus the excluded ids (Index::search_wildcard path)
        Kw = max(1, min(max(topster, 250), coll.n_docs))
        query = S.Query([], topk=Kw, sort=sort, excl=excl)
        kv, cnt, found = wildcard_backend(S.KwBatch([query], list(range(len(coll.fields)))), Kw)
        return [int(kv["key"][0, i]) for i in range(int(cnt[0]))], int(found[0])
    K = max(1, min(max(topster, 250), coll.n_docs))          # src/index.cpp:3506-3512
    best: Dict[int, tuple] = {}
    all_ids = set()

    F = len(coll.fields)

    def row_of(t):
        return [v.get(t, S.NO_LIST) for v in coll.vocabs]

    def run_round(trunc: List[str], dropped: List[str]):
        if not trunc or any(all(x == S.NO_LIST for x in row_of(t)) for t in trunc):
            return                                           # no candidate at cost 0: fuzzy_search_fields returns
        rows = [row_of(t) for t in trunc] + [row_of(t) for t in dropped]
        query = S.Query([S.Combo(rows, len(trunc))], topk=K, sort=sort, num_query_tokens=len(trunc), excl=excl, flags=flags,
                        field_weight=list(field_weights) if field_weights else [max(0, 15 - f) for f in range(F)])   # src/collection.cpp:4219-4225
        kv, cnt, found = backend(S.KwBatch([query], list(range(F))), K)
        for i in range(int(cnt[0])):
            key = int(kv["key"][0, i])
            tup = tuple(int(x) for x in kv["scores"][0, i])
            all_ids.add(key)
            if key not in best or tup >= best[key]:          # Topster::add keeps the greater KV per key
                best[key] = tup

    run_round(tokens, [])
    n = min(len(tokens), 20)
    if len(all_ids) < drop_tokens_threshold:
        n_dropped, dirs_done, rtl = 0, 0, True
        while len(all_ids) < drop_tokens_threshold:
            if n_dropped >= n - 1:
                rtl = not rtl
                n_dropped = 0
                dirs_done += 1
            if n > 1 and dirs_done < 2:
                if rtl:
                    tl = n - n_dropped - 1
                    trunc, dropped = tokens[:tl], tokens[tl:n]
                else:
                    st = n_dropped + 1
                    trunc, dropped = tokens[st:n], tokens[:st]
                n_dropped += 1
                run_round(trunc, dropped)
            else:
                break
    order = sorted(best.items(), key=lambda kv_: (kv_[1], kv_[0]), reverse=True)
    search.last_scores = {k: v for k, v in order}            # seq_id -> scores[0..2] of the last call (for text_match KATs)
    return [k for k, _ in order], len(all_ids)
