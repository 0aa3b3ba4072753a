// TEST INFRASTRUCTURE: compiles the product's __host__ __device__ headers (typesense_b200/csrc/score_device.cuh,
// postings_device.cuh, postings_pack.h) with g++ and runs, per candidate document, exactly the sequence of calls a
// kw_search_kernel thread makes — so the device scoring / block-probe logic can be checked against the oracle in a
// container without a GPU. Never part of libtsgpu.so; the product has no CPU path.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/tsgpu.h"
#include "../../typesense_b200/csrc/postings_pack.h"

using namespace tsdev;

namespace {
struct HostField {
    tspack::PackedField pk;
    DevField dev;
};
int g_reg_score = 0;          // hs_set_reg_score: score like kw_search_kernel<REGSCORE = true>
long g_reg_score_hits = 0;    // documents x fields scored by that branch
}

extern "C" {

// fields: the index's fields (tsgpu_field), batch b, query q, combination c
size_t hs_keyword_combo(const tsgpu_field* fields, uint32_t n_index_fields, const tsgpu_kw_batch* b, uint32_t q, uint32_t c,
                        uint32_t* out_ids, uint64_t* out_scores, size_t cap) {
    std::vector<HostField> hf(n_index_fields);
    for(uint32_t i = 0; i < n_index_fields; i++) {
        tspack::pack_field(fields[i].n_lists, fields[i].list_off, fields[i].ids, hf[i].pk);
        DevField& d = hf[i].dev;
        d.n_lists = fields[i].n_lists; d.list_off = fields[i].list_off;
        d.is_array = fields[i].is_array ? kFieldIsArray : (tspack::plain_wellformed(fields[i].pos_off, fields[i].positions, fields[i].list_off[fields[i].n_lists]) ? kFieldPlainOk : 0);
        if((d.is_array & kFieldPlainOk) && tspack::positions_fit_u16(fields[i].positions, fields[i].pos_off[fields[i].list_off[fields[i].n_lists]])) d.is_array |= kFieldPos16;
        d.list_blk_off = hf[i].pk.list_blk_off.data(); d.blk_first = hf[i].pk.blk_first.data();
        d.blk_info = hf[i].pk.blk_info.data(); d.packed = hf[i].pk.packed.data();
        d.pos_off = fields[i].pos_off; d.positions = fields[i].positions;
        {   // dense lists: low threshold here so the bitmap/rank path is exercised on small test collections
            uint32_t max_id = 0;
            for(uint64_t k = 0; k < fields[i].list_off[fields[i].n_lists]; k++) if(fields[i].ids[k] > max_id) max_id = fields[i].ids[k];
            tspack::pack_dense(fields[i].n_lists, fields[i].list_off, fields[i].ids, max_id + 1, 64, hf[i].pk);
            d.list_dense = hf[i].pk.list_dense.data(); d.dense_bits = hf[i].pk.dense_bits.data(); d.dense_rank = hf[i].pk.dense_rank.data();
            d.dense_words = hf[i].pk.dense_words; d.dense_groups = hf[i].pk.dense_groups;
        }
    }
    const uint32_t F = b->n_fields;
    const uint32_t row0 = b->c_tok_off[c], n_rows = b->c_tok_off[c + 1] - row0, n_req = b->c_n_required[c];
    auto list_of = [&](uint32_t r, uint32_t f) { return b->t_list[(size_t) (row0 + r) * F + f]; };
    auto fld = [&](uint32_t f) -> const DevField& { return hf[b->field_ids[f]].dev; };
    auto df_of = [&](uint32_t r, uint32_t f) -> uint64_t {
        uint32_t l = list_of(r, f);
        return l == TSGPU_NO_LIST ? 0 : fld(f).list_off[l + 1] - fld(f).list_off[l];
    };
    // required rows that exist in at least one field (a token found in no field is skipped, src/index.cpp:5648)
    std::vector<uint32_t> req;
    for(uint32_t r = 0; r < n_req; r++) { uint64_t s = 0; for(uint32_t f = 0; f < F; f++) s += df_of(r, f); if(s) req.push_back(r); }
    if(req.empty()) return 0;
    uint32_t drv = req[0]; uint64_t best = ~0ull;
    for(uint32_t r: req) { uint64_t s = 0; for(uint32_t f = 0; f < F; f++) s += df_of(r, f); if(s < best) { best = s; drv = r; } }

    ScoreParams P;
    P.total_cost = b->c_total_cost[c];
    P.num_query_tokens = b->q_num_query_tokens[q];
    P.syn_orig_num_tokens = b->c_syn_orig_num_tokens ? b->c_syn_orig_num_tokens[c] : -1;
    P.orig_num_tokens = b->c_orig_num_tokens ? b->c_orig_num_tokens[c] : -1;
    uint8_t cf = b->c_flags ? b->c_flags[c] : 0;
    P.is_synonym_query = cf & 1; P.demote_synonym_match = (cf & 2) ? 1 : 0;
    uint8_t qf = b->q_flags[q];
    P.prioritize_exact_match = qf & 1; P.prioritize_token_position = (qf & 2) ? 1 : 0;
    P.prioritize_num_matching_fields = (qf & 4) ? 1 : 0;
    P.match_type = b->q_match_type[q];

    const uint32_t* excl = b->excl_ids + b->q_excl_off[q];
    size_t n_excl = b->q_excl_off[q + 1] - b->q_excl_off[q];
    const uint32_t* filt = nullptr; size_t n_filt = 0; bool has_filter = b->q_filter[q] >= 0;
    if(has_filter) { filt = b->filter_ids + b->filter_off[b->q_filter[q]]; n_filt = b->filter_off[b->q_filter[q] + 1] - b->filter_off[b->q_filter[q]]; }
    if(has_filter && n_filt == 0) return 0;

    size_t n = 0;
    std::vector<uint32_t> hit(n_rows * F);
    for(uint32_t fd = 0; fd < F; fd++) {
        uint32_t dl = list_of(drv, fd);
        if(dl == TSGPU_NO_LIST) continue;
        const DevField& DF = fld(fd);
        uint64_t ddf = DF.list_off[dl + 1] - DF.list_off[dl];
        uint32_t lb0 = DF.list_blk_off[dl];
        for(uint64_t pi = 0; pi < ddf; pi++) {
            // decode through the packed block, as the kernel does
            uint32_t bblk = lb0 + (uint32_t) (pi / kBlock), idx = (uint32_t) (pi % kBlock);
            uint64_t info = DF.blk_info[bblk];
            uint32_t id = DF.blk_first[bblk] + unpack_at(DF.packed + (info & 0xFFFFFFFFFFull), (uint32_t) (info >> 40) & 0xFF, idx);
            if(n_excl && std::binary_search(excl, excl + n_excl, id)) continue;
            if(has_filter && !std::binary_search(filt, filt + n_filt, id)) continue;
            bool alive = true;
            for(uint32_t r = 0; r < n_rows && alive; r++) {
                bool any = false;
                for(uint32_t f = 0; f < F; f++) {
                    uint32_t l = list_of(r, f);
                    uint32_t h = kNone;
                    if(r == drv && f == fd) h = (uint32_t) pi;
                    else if(l != TSGPU_NO_LIST) {
                        const DevField& G = fld(f);
                        if(G.list_blk_off[l + 1] > G.list_blk_off[l])
                            h = probe_list(G, l, G.list_blk_off[l], G.list_blk_off[l + 1] - 1, id);
                    }
                    hit[r * F + f] = h;
                    if(h != kNone) any = true;
                    if(r == drv && f < fd && h != kNone) alive = false;   // already produced from an earlier field's tile
                }
                bool required = std::find(req.begin(), req.end(), r) != req.end();
                if(required && !any) alive = false;
            }
            if(!alive) continue;
            // score
            FieldAgg agg; field_agg_init(agg);
            uint32_t query_len = 0;
            for(uint32_t r = 0; r < n_rows; r++) { bool any = false; for(uint32_t f = 0; f < F; f++) if(hit[r * F + f] != kNone) any = true; if(any) query_len++; }
            for(uint32_t f = 0; f < F; f++) {
                RawTok toks[kMaxTokens]; int nt = 0;
                const DevField& G = fld(f);
                if(g_reg_score && n_rows <= (uint32_t) kSmallTokens && (G.is_array & kFieldPos16)) {        // the REGSCORE branch of the kernel
                    const uint32_t* tp[kSmallTokens];
                    uint32_t tn[kSmallTokens], present = 0;
                    for(int r = 0; r < kSmallTokens; r++) {
                        tp[r] = G.positions; tn[r] = 0;
                        if((uint32_t) r < n_rows) {
                            uint32_t h = hit[r * F + f];
                            if(h != kNone) {
                                uint64_t p = G.list_off[list_of(r, f)] + h;
                                tp[r] = G.positions + G.pos_off[p]; tn[r] = (uint32_t) (G.pos_off[p + 1] - G.pos_off[p]); present |= 1u << r;
                            }
                        }
                    }
                    if(!present) continue;
                    g_reg_score_hits++;
                    int64_t fs = score_field_plain_small<kSmallTokens>(P, P.total_cost == 0 && P.num_query_tokens == 1, tp, tn, present);
                    field_agg_add(agg, P.match_type, fs, b->q_field_weight[(size_t) q * F + f]);
                    continue;
                }
                for(uint32_t r = 0; r < n_rows; r++) {
                    uint32_t h = hit[r * F + f];
                    if(h == kNone) continue;
                    uint64_t p = G.list_off[list_of(r, f)] + h;
                    toks[nt].p = G.positions + G.pos_off[p];
                    toks[nt].n = (uint32_t) (G.pos_off[p + 1] - G.pos_off[p]);
                    nt++;
                }
                if(nt == 0) continue;
                bool single_exact = (P.total_cost == 0 && P.num_query_tokens == 1);
                int64_t fs = (G.is_array & kFieldPlainOk) ? score_field_plain(P, single_exact, toks, nt) : score_field(P, (G.is_array & kFieldIsArray) != 0, single_exact, toks, nt);
                field_agg_add(agg, P.match_type, fs, b->q_field_weight[(size_t) q * F + f]);
            }
            uint64_t s = field_agg_finish(agg, P, query_len);
            if(n < cap) { out_ids[n] = id; out_scores[n] = s; }
            n++;
        }
    }
    // the kernel emits per driver field; the oracle emits ascending ids — sort for comparison
    if(n <= cap) {
        std::vector<std::pair<uint32_t, uint64_t>> v(n);
        for(size_t i = 0; i < n; i++) v[i] = {out_ids[i], out_scores[i]};
        std::sort(v.begin(), v.end());
        for(size_t i = 0; i < n; i++) { out_ids[i] = v[i].first; out_scores[i] = v[i].second; }
    }
    return n;
}

// probe every query id in a single packed list; out[i] = list-local index or 0xFFFFFFFF
void hs_probe_ids(const uint32_t* ids, uint64_t n, const uint32_t* queries, uint64_t nq, uint32_t* out) {
    uint64_t list_off[2] = {0, n};
    tspack::PackedField pk;
    tspack::pack_field(1, list_off, ids, pk);
    DevField d{};
    d.n_lists = 1; d.list_off = list_off; d.list_blk_off = pk.list_blk_off.data(); d.blk_first = pk.blk_first.data();
    d.blk_info = pk.blk_info.data(); d.packed = pk.packed.data();
    for(uint64_t i = 0; i < nq; i++) out[i] = n ? probe_list(d, 0, 0, pk.list_blk_off[1] - 1, queries[i]) : kNone;
}

// isect_tiles_kernel's id-set modes on one flat field: mode 1 phrase, 2 exact, 3 prefix. Per id: locate it in each list
// (ids absent from a list are dropped) and run the same per-document routine the kernel thread runs.
size_t hs_idset_matches(const tsgpu_field* f, const uint32_t* lists, uint32_t k, const uint32_t* ids, size_t n, int mode,
                        uint32_t* out) {
    size_t n_out = 0;
    for(size_t i = 0; i < n; i++) {
        RawTok toks[kMaxTokens];
        bool alive = true;
        for(uint32_t j = 0; j < k && alive; j++) {
            const uint32_t* b = f->ids + f->list_off[lists[j]];
            const uint32_t* e = f->ids + f->list_off[lists[j] + 1];
            const uint32_t* it = std::lower_bound(b, e, ids[i]);
            if(it == e || *it != ids[i]) { alive = false; break; }
            const uint64_t p = (uint64_t) (it - f->ids);
            toks[j].p = f->positions + f->pos_off[p];
            toks[j].n = (uint32_t) (f->pos_off[p + 1] - f->pos_off[p]);
        }
        if(!alive) continue;
        const bool ok = mode == 1 ? phrase_match_doc(toks, (int) k) : positional_match_doc(toks, (int) k, f->is_array != 0, mode == 2);
        if(ok) out[n_out++] = ids[i];
    }
    return n_out;
}

// Data of test/collection_vector_search_test.cpp:5094-5125 (TestDistanceThresholdWithIP): std::mt19937 seeded with 47,
// five documents of five uniform_real_distribution<>(-1,1) draws (stored as float) each followed by one
// uniform_int_distribution<>(0,100) draw for rank_score. libstdc++'s distributions, like the reference's build.
void hs_ip_kat_data(float* vec_out, int* rank_out) {
    std::mt19937 rng;
    rng.seed(47);
    std::uniform_real_distribution<> distrib(-1, 1);
    std::uniform_int_distribution<> distrib2(0, 100);
    for(int i = 0; i < 5; i++) {
        for(int j = 0; j < 5; j++) vec_out[i * 5 + j] = (float) distrib(rng);
        rank_out[i] = distrib2(rng);
    }
}

int hs_phrase_match_doc(uint32_t k, const uint32_t* tok_off, const uint32_t* raw) {
    RawTok toks[kMaxTokens];
    for(uint32_t t = 0; t < k; t++) { toks[t].p = raw + tok_off[t]; toks[t].n = tok_off[t + 1] - tok_off[t]; }
    return phrase_match_doc(toks, (int) k) ? 1 : 0;
}

void hs_set_reg_score(int on) { g_reg_score = on; }
long hs_reg_score_hits() { return g_reg_score_hits; }

// One plain field, rows 0..n_rows-1 (bit r of `present`: row r matched; its raw offsets are raw[tok_off[r]..tok_off[r+1])):
// out[0] = score_field_plain() on the matched rows, out[1] = score_field_plain_small() — the two must agree.
// params = total_cost, num_query_tokens, syn_orig_num_tokens, orig_num_tokens, is_synonym_query, demote_synonym_match,
//          prioritize_exact_match, prioritize_token_position
void hs_score_plain_both(uint32_t n_rows, uint32_t present, const uint32_t* tok_off, const uint32_t* raw, const int32_t* params,
                         int64_t* out) {
    ScoreParams P;
    P.total_cost = (uint32_t) params[0]; P.num_query_tokens = (uint32_t) params[1]; P.syn_orig_num_tokens = params[2];
    P.orig_num_tokens = params[3]; P.is_synonym_query = (uint8_t) params[4]; P.demote_synonym_match = (uint8_t) params[5];
    P.prioritize_exact_match = (uint8_t) params[6]; P.prioritize_token_position = (uint8_t) params[7];
    P.prioritize_num_matching_fields = 1; P.match_type = 0;
    const bool single_exact = P.total_cost == 0 && P.num_query_tokens == 1;
    RawTok toks[kMaxTokens]; int nt = 0;
    const uint32_t* tp[kSmallTokens]; uint32_t tn[kSmallTokens];
    for(int r = 0; r < kSmallTokens; r++) {
        tp[r] = raw; tn[r] = 0;
        if((uint32_t) r < n_rows && ((present >> r) & 1u)) {
            tp[r] = raw + tok_off[r]; tn[r] = tok_off[r + 1] - tok_off[r];
            toks[nt].p = tp[r]; toks[nt].n = tn[r]; nt++;
        }
    }
    out[0] = score_field_plain(P, single_exact, toks, nt);
    out[1] = score_field_plain_small<kSmallTokens>(P, single_exact, tp, tn, present);
}

void hs_sort_scores(const uint8_t* type, const int8_t* order, const uint8_t* missing_first, const int64_t* const* cols,
                    uint32_t seq_id, int64_t text, float vdist, int64_t* scores, int* msi) {
    SortSpec S;
    for(int i = 0; i < 3; i++) { S.type[i] = type[i]; S.order[i] = order[i]; S.missing_first[i] = missing_first[i]; S.col[i] = cols[i]; }
    *msi = compute_sort_scores(S, seq_id, text, vdist, scores);
}

}  // extern "C"
