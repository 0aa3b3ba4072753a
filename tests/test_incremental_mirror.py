"""SURVEY §8 f-4, posting half: tsgpu_index_append_lists / tsgpu_index_set_sort_values — the device side of posting_t::upsert / erase
(src/posting.cpp:247-333). A mirror that was loaded from an early state of a collection and then patched must answer exactly like a
mirror (and the oracle) loaded from the final state: same ids, scores, found — for inserted documents, updated documents (new offsets
under an old seq_id) and removed ones."""
import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import capi, structs as S, synth

pytestmark = pytest.mark.gpu


def _lists_of(flat: S.FlatField):
    """per list: [(seq_id, [offsets])]"""
    out = []
    for l in range(flat.n_lists):
        a, b = int(flat.list_off[l]), int(flat.list_off[l + 1])
        out.append([(int(flat.ids[i]), flat.positions[int(flat.pos_off[i]):int(flat.pos_off[i + 1])].tolist()) for i in range(a, b)])
    return out


def _subset(lists, keep_doc):
    return [[p for p in pl if keep_doc(p[0])] for pl in lists]


@pytest.mark.parametrize("is_array", [False, True])
def test_patched_mirror_equals_fresh_mirror(is_array):
    n_docs, vocab = 6000, 400
    fd = synth.make_array_field(n_docs, vocab, 1, 3, 2, 5, seed=21) if is_array else synth.make_string_field(n_docs, vocab, 3, 9, seed=21)
    final_lists = _lists_of(fd.flat)
    L = len(final_lists)
    pts_final = synth.make_points(n_docs, 5)
    rng = np.random.default_rng(3)
    # state 0: documents < 4000 only, and 300 of them with other offsets (they will be "updated"), 200 extra documents that will be removed
    changed = set(rng.choice(4000, 300, replace=False).tolist())
    removed = set(rng.choice(np.arange(4000, 4500), 200, replace=False).tolist())

    def early(pl):
        out = []
        for sid, offs in pl:
            if sid < 4000 or sid in removed:
                o = list(offs)
                if sid in changed and not is_array and len(o) and o[0] > 1:
                    o[0] -= 1                                  # an older version of the document had the token one position earlier
                out.append((sid, o))
        return out
    lists0 = [early(pl) for pl in final_lists]
    # the final state has no `removed` documents
    final_lists = [[p for p in pl if p[0] not in removed] for pl in final_lists]
    flat0 = S.FlatField.from_postings(lists0, is_array)
    flat_final = S.FlatField.from_postings(final_lists, is_array)
    pts0 = pts_final.copy()
    pts0[4000:] = np.iinfo(np.int64).min
    gi = capi.GpuIndex(n_docs, 0)
    f = gi.load_field(flat0)
    col = gi.load_sort_column(pts0)
    # patch in two batches: every list whose content differs between state 0 and the final state is handed over in full
    touched = [l for l in range(L) if lists0[l] != final_lists[l]]
    assert len(touched) > L // 2
    remap = np.arange(L, dtype=np.int64)
    half = len(touched) // 2
    for part in (touched[:half], touched[half:]):
        delta = S.FlatField.from_postings([final_lists[l] for l in part], is_array)
        first = gi.append_lists(f, delta)
        for k, l in enumerate(part):
            remap[l] = first + k
    ids_new = np.arange(4000, n_docs, dtype=np.uint32)
    vals_new = pts_final[4000:].copy()
    vals_new[[i - 4000 for i in removed]] = np.iinfo(np.int64).min
    gi.set_sort_values(col, ids_new, vals_new)
    pts_expect = pts_final.copy()
    pts_expect[list(removed)] = np.iinfo(np.int64).min
    oi = ol.OracleIndex(n_docs, [flat_final], [pts_expect])
    gf = capi.GpuIndex(n_docs, 0)
    gf.load_field(flat_final); gf.load_sort_column(pts_expect)
    sort = ((S.SORT_TEXT_MATCH, -1, 1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
    df = np.diff(flat_final.list_off.astype(np.int64))
    live = np.nonzero(df > 0)[0]
    for seed in range(3):
        r = np.random.default_rng(100 + seed)
        qs_final, qs_patched = [], []
        for i in range(80):
            nt = int(r.integers(1, 4))
            toks = [int(t) for t in r.choice(live, nt, replace=False)]
            if i % 4 == 0:                                   # tokens of one document: guaranteed hits
                d = int(r.integers(0, n_docs))
                cand = [l for l in live[:200] if any(p[0] == d for p in final_lists[l])]
                if len(cand) >= nt:
                    toks = [int(t) for t in cand[:nt]]
            qs_final.append(S.Query([S.Combo([[t] for t in toks], nt)], topk=int(r.choice([10, 250])), sort=sort, num_query_tokens=nt))
            qs_patched.append(S.Query([S.Combo([[int(remap[t])] for t in toks], nt)], topk=qs_final[-1].topk, sort=sort, num_query_tokens=nt))
        bf, bp = S.KwBatch(qs_final, [0]), S.KwBatch(qs_patched, [0])
        okv, ocnt, ofound = oi.keyword_search(bf, 256)
        fkv, fcnt, ffound = gf.keyword_search(bf, 256)
        pkv, pcnt, pfound = gi.keyword_search(bp, 256)
        assert int(ocnt.sum()) > 50
        for kv, cnt, found in ((fkv, fcnt, ffound), (pkv, pcnt, pfound)):
            assert cnt.tolist() == ocnt.tolist() and found.tolist() == ofound.tolist()
            for q in range(len(cnt)):
                n = int(cnt[q])
                assert kv["key"][q, :n].tolist() == okv["key"][q, :n].tolist(), f"query {q}: ids"
                assert kv["scores"][q, :n].tolist() == okv["scores"][q, :n].tolist(), f"query {q}: scores"
                assert kv["text_match_score"][q, :n].tolist() == okv["text_match_score"][q, :n].tolist()
    # the id-set primitives follow the new lists too
    a, b = int(live[0]), int(live[1])
    assert gi.intersect(f, [int(remap[a]), int(remap[b])], n_docs).tolist() == gf.intersect(0, [a, b], n_docs).tolist()
    gi.close(); gf.close()


def test_append_lists_rejects_malformed_input():
    n_docs = 1000
    fd = synth.make_string_field(n_docs, 50, 3, 6, seed=2)
    gi = capi.GpuIndex(n_docs, 0)
    f = gi.load_field(fd.flat)
    bad = S.FlatField.from_postings([[(5, [1]), (3, [2])]])                      # not ascending
    with pytest.raises(capi.TsgpuError):
        gi.append_lists(f, bad)
    with pytest.raises(capi.TsgpuError):
        gi.append_lists(f, S.FlatField.from_postings([[(n_docs + 3, [1])]]))     # beyond the index's capacity
    with pytest.raises(capi.TsgpuError):
        gi.append_lists(f + 7, S.FlatField.from_postings([[(1, [1])]]))
    # the field is unchanged
    assert gi.intersect(f, [0], n_docs).tolist() == fd.flat.ids[int(fd.flat.list_off[0]):int(fd.flat.list_off[1])].tolist()
    gi.close()
