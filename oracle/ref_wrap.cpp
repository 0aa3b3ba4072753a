// TEST INFRASTRUCTURE — oracle/_ref: thin C wrapper around the REFERENCE'S OWN sources, compiled in place from
// /root/reference/src/{posting_list,posting,or_iterator,sorted_array,array,array_base,array_utils,
// thread_local_vars}.cpp + include/match_score.h behind the shim headers in oracle/ref_shims/ (recipe:
// oracle/Makefile, outputs only into oracle/_ref/). Nothing here is linked into the product library; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
//
// What is reference code and what is glue:
//   * posting_list_t / iterator_t / or_iterator_t::intersect / get_offsets / Match / ArrayUtils / sorted_array / array
//     are the reference's code, unmodified.
//   * score_results2 (src/index.cpp:6966-7098) and compute_aggregated_score (src/index.cpp:5227-5383) live in
//     index.cpp, which cannot be compiled offline (rocksdb/ICU/s2/hnswlib); the two functions refglue_* below restate
//     them line by line ON TOP OF the reference's iterator_t/Match types so that the standalone oracle
//     (oracle/ts_oracle.cpp, flat arrays, own Match) can be validated against reference-typed execution.
#include <cstdint>
#include <cstring>
#include <vector>
#include <map>
#include <algorithm>

#include "posting_list.h"
#include "posting.h"
#include "or_iterator.h"
#include "array_utils.h"
#include "sorted_array.h"
#include "array.h"
#include "match_score.h"
#include "thread_local_vars.h"
#include "art.h"          // through oracle/_ref/stage (see Makefile): the reference's own adaptive radix tree
#include <set>
#include <string>

extern "C" {

// ---------------------------------------------------------------- containers (pins the libfor port)
void* ref_sorted_array_new() { return new sorted_array(); }
void ref_sorted_array_free(void* h) { delete (sorted_array*) h; }
uint32_t ref_sorted_array_append(void* h, uint32_t v) { return (uint32_t) ((sorted_array*) h)->append(v); }
void ref_sorted_array_load(void* h, const uint32_t* a, uint32_t n) { ((sorted_array*) h)->load(a, n); }
uint32_t ref_sorted_array_at(void* h, uint32_t i) { return ((sorted_array*) h)->at(i); }
uint32_t ref_sorted_array_length(void* h) { return ((sorted_array*) h)->getLength(); }
int ref_sorted_array_contains(void* h, uint32_t v) { return ((sorted_array*) h)->contains(v); }
uint32_t ref_sorted_array_index_of(void* h, uint32_t v) { return ((sorted_array*) h)->indexOf(v); }
void ref_sorted_array_bulk_index_of(void* h, const uint32_t* vals, uint32_t n, uint32_t* out) {
    ((sorted_array*) h)->indexOf(vals, n, out);
}
uint32_t ref_sorted_array_num_found_of(void* h, const uint32_t* vals, uint32_t n) {
    return (uint32_t) ((sorted_array*) h)->numFoundOf(vals, n);
}
void ref_sorted_array_remove_value(void* h, uint32_t v) { ((sorted_array*) h)->remove_value(v); }
void ref_sorted_array_uncompress(void* h, uint32_t* out) {
    auto* a = (sorted_array*) h;
    uint32_t* u = a->uncompress();
    memcpy(out, u, sizeof(uint32_t) * a->getLength());
    delete[] u;
}
void* ref_array_new() { return new array(); }
void ref_array_free(void* h) { delete (array*) h; }
void ref_array_append(void* h, uint32_t v) { ((array*) h)->append(v); }
uint32_t ref_array_at(void* h, uint32_t i) { return ((array*) h)->at(i); }
uint32_t ref_array_length(void* h) { return ((array*) h)->getLength(); }
uint32_t ref_array_index_of(void* h, uint32_t v) { return ((array*) h)->indexOf(v); }
void ref_array_remove_index(void* h, uint32_t s, uint32_t e) { ((array*) h)->remove_index(s, e); }

// ---------------------------------------------------------------- ArrayUtils (include/array_utils.h:13-23)
// caller passes an output buffer of sufficient capacity; returns the result length
size_t ref_and_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out) {
    uint32_t* res = nullptr;
    size_t n = ArrayUtils::and_scalar(a, na, b, nb, &res);
    if(n) memcpy(out, res, n * sizeof(uint32_t));
    delete[] res;
    return n;
}
size_t ref_or_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out) {
    uint32_t* res = nullptr;
    size_t n = ArrayUtils::or_scalar(a, na, b, nb, &res);
    if(n) memcpy(out, res, n * sizeof(uint32_t));
    delete[] res;
    return n;
}
size_t ref_exclude_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out) {
    uint32_t* res = nullptr;
    size_t n = ArrayUtils::exclude_scalar(a, na, b, nb, &res);
    if(n) memcpy(out, res, n * sizeof(uint32_t));
    delete[] res;
    return n;
}

// ---------------------------------------------------------------- posting_list_t
void* ref_plist_new(uint32_t block_max) { return new posting_list_t((uint16_t) block_max); }
void ref_plist_free(void* h) { delete (posting_list_t*) h; }
void ref_plist_upsert(void* h, uint32_t id, const uint32_t* offsets, uint32_t n) {
    ((posting_list_t*) h)->upsert(id, std::vector<uint32_t>(offsets, offsets + n));
}
void ref_plist_erase(void* h, uint32_t id) { ((posting_list_t*) h)->erase(id); }
uint32_t ref_plist_num_ids(void* h) { return (uint32_t) ((posting_list_t*) h)->num_ids(); }
uint32_t ref_plist_num_blocks(void* h) { return (uint32_t) ((posting_list_t*) h)->num_blocks(); }
// bulk build: ids ascending, offs[off_index[i]..off_index[i+1]) are id i's offsets
void ref_plist_bulk(void* h, const uint32_t* ids, const uint32_t* off_index, const uint32_t* offs, uint32_t n) {
    auto* pl = (posting_list_t*) h;
    for(uint32_t i = 0; i < n; i++) {
        pl->upsert(ids[i], std::vector<uint32_t>(offs + off_index[i], offs + off_index[i + 1]));
    }
}
// dump (ids, offset_index, offsets) through the reference's own iterator (pins block decode, a1)
size_t ref_plist_dump(void* h, uint32_t* ids, uint32_t* off_index, uint32_t* offs, size_t cap_ids, size_t cap_offs) {
    auto* pl = (posting_list_t*) h;
    auto it = pl->new_iterator();
    size_t n = 0, no = 0;
    while(it.valid()) {
        std::vector<uint32_t> p;
        posting_list_t::get_offsets(it, p);
        if(n >= cap_ids || no + p.size() > cap_offs) return (size_t) -1;
        ids[n] = it.id();
        off_index[n] = (uint32_t) no;
        for(auto v: p) offs[no++] = v;
        n++;
        it.next();
    }
    off_index[n] = (uint32_t) no;
    return n;
}

size_t ref_plist_intersect(void** hs, uint32_t k, uint32_t* out, size_t cap) {
    std::vector<posting_list_t*> pls;
    for(uint32_t i = 0; i < k; i++) pls.push_back((posting_list_t*) hs[i]);
    std::vector<uint32_t> res;
    posting_list_t::intersect(pls, res);
    if(res.size() > cap) return (size_t) -1;
    if(!res.empty()) memcpy(out, res.data(), res.size() * sizeof(uint32_t));
    return res.size();
}
size_t ref_plist_merge(void** hs, uint32_t k, uint32_t* out, size_t cap) {
    std::vector<posting_list_t*> pls;
    for(uint32_t i = 0; i < k; i++) pls.push_back((posting_list_t*) hs[i]);
    std::vector<uint32_t> res;
    posting_list_t::merge(pls, res);
    if(res.size() > cap) return (size_t) -1;
    if(!res.empty()) memcpy(out, res.data(), res.size() * sizeof(uint32_t));
    return res.size();
}
int ref_plist_contains_atleast_one(void* h, const uint32_t* ids, size_t n) {
    return ((posting_list_t*) h)->contains_atleast_one(ids, n);
}

// posting_t::block_intersector_t-style AND with filter / exclusion (src/posting_list.cpp:794, include/posting_list.h:241)
size_t ref_plist_block_intersect(void** hs, uint32_t k, const uint32_t* excl, size_t n_excl,
                                 const uint32_t* filt, size_t n_filt, uint32_t* out, size_t cap) {
    std::vector<posting_list_t::iterator_t> its;
    its.reserve(k);
    for(uint32_t i = 0; i < k; i++) its.push_back(((posting_list_t*) hs[i])->new_iterator());
    result_iter_state_t istate(excl, n_excl, filt, n_filt);
    std::vector<uint32_t> res;
    posting_list_t::block_intersect(its, istate, [&](uint32_t id, std::vector<posting_list_t::iterator_t>&) {
        res.push_back(id);
    });
    if(res.size() > cap) return (size_t) -1;
    if(!res.empty()) memcpy(out, res.data(), res.size() * sizeof(uint32_t));
    return res.size();
}

// phrase matches over an id set (src/posting_list.cpp:1791)
size_t ref_plist_phrase_matches(void** hs, uint32_t k, int field_is_array, const uint32_t* ids, uint32_t n,
                                uint32_t* out) {
    std::vector<posting_list_t::iterator_t> its;
    its.reserve(k);
    for(uint32_t i = 0; i < k; i++) its.push_back(((posting_list_t*) hs[i])->new_iterator());
    size_t n_out = 0;
    uint32_t* o = out;
    posting_list_t::get_phrase_matches(its, field_is_array, ids, n, o, n_out);
    return n_out;
}
size_t ref_plist_exact_matches(void** hs, uint32_t k, int field_is_array, const uint32_t* ids, uint32_t n,
                               uint32_t* out) {
    std::vector<posting_list_t::iterator_t> its;
    its.reserve(k);
    for(uint32_t i = 0; i < k; i++) its.push_back(((posting_list_t*) hs[i])->new_iterator());
    size_t n_out = 0;
    uint32_t* o = out;
    posting_list_t::get_exact_matches(its, field_is_array, ids, n, o, n_out);
    return n_out;
}
size_t ref_plist_prefix_matches(void** hs, uint32_t k, int field_is_array, const uint32_t* ids, uint32_t n,
                                uint32_t* out) {
    std::vector<posting_list_t::iterator_t> its;
    its.reserve(k);
    for(uint32_t i = 0; i < k; i++) its.push_back(((posting_list_t*) hs[i])->new_iterator());
    size_t n_out = 0;
    uint32_t* o = out;
    posting_list_t::get_prefix_matches(its, field_is_array, ids, n, o, n_out);
    return n_out;
}

// ---------------------------------------------------------------- Match (include/match_score.h:129-275)
// tok_off[t]..tok_off[t+1] index `positions`; last_token[t] flags. out[4] = words_present, distance, max_offset, exact
void ref_match(uint32_t n_tokens, const uint32_t* tok_off, const uint16_t* positions, const uint8_t* last_token,
               int check_exact, uint8_t* out) {
    std::vector<token_positions_t> tp(n_tokens);
    for(uint32_t t = 0; t < n_tokens; t++) {
        tp[t].last_token = last_token[t];
        tp[t].positions.assign(positions + tok_off[t], positions + tok_off[t + 1]);
    }
    Match m(0, tp, false, check_exact);
    out[0] = m.words_present; out[1] = m.distance; out[2] = m.max_offset; out[3] = m.exact_match;
}
uint64_t ref_match_score(uint8_t words_present, uint8_t distance, uint8_t max_offset, uint8_t exact,
                         uint32_t total_cost, uint32_t unique_words, uint8_t syn) {
    Match m(words_present, distance, max_offset, exact);
    return m.get_match_score(total_cost, unique_words, syn);
}
int ref_has_phrase_match(uint32_t n_tokens, const uint32_t* tok_off, const uint16_t* positions) {
    std::vector<token_positions_t> tp(n_tokens);
    for(uint32_t t = 0; t < n_tokens; t++) tp[t].positions.assign(positions + tok_off[t], positions + tok_off[t + 1]);
    return posting_list_t::has_phrase_match(tp);
}

// ---------------------------------------------------------------- reference-typed keyword combination
struct refglue_params {
    uint32_t n_tokens;            // required tokens (AND)
    uint32_t n_dropped;           // optional tokens appended after the required ones in `lists`
    uint32_t n_fields;
    uint32_t total_cost;
    uint32_t num_query_tokens;    // query_tokens.size()
    int32_t  syn_orig_num_tokens; // -1 default
    int32_t  orig_num_tokens;
    uint8_t  is_synonym_query, demote_synonym_match;
    uint8_t  prioritize_exact_match, prioritize_token_position, prioritize_num_matching_fields;
    uint8_t  match_type;          // 0 max_score, 1 max_weight, 2 sum_score (include/index.h text_match_type_t)
    uint8_t  pad[2];
    int64_t  field_weight[32];
    uint8_t  field_is_array[32];
};

// restated src/index.cpp:6966-7098 over reference iterator_t / Match
static void refglue_score_results2(const refglue_params& P, bool field_is_array, bool single_exact_query_token,
                                   int64_t& match_score, uint32_t seq_id,
                                   const std::vector<posting_list_t::iterator_t>& posting_lists) {
    const uint32_t total_cost = P.total_cost;
    const size_t num_query_tokens = P.num_query_tokens;
    const int syn_orig_num_tokens = P.syn_orig_num_tokens, orig_num_tokens = P.orig_num_tokens;
    const bool is_synonym_query = P.is_synonym_query, demote_synonym_match = P.demote_synonym_match;
    if(posting_lists.size() <= 1) {
        const uint8_t is_verbatim_match = uint8_t(P.prioritize_exact_match && single_exact_query_token &&
                posting_list_t::is_single_token_verbatim_match(posting_lists[0], field_is_array));
        size_t words_present = (num_query_tokens == 1 && is_synonym_query) ? syn_orig_num_tokens : 1;
        size_t distance = (num_query_tokens == 1 && is_synonym_query) ? syn_orig_num_tokens - 1 : 0;
        size_t max_offset = P.prioritize_token_position ? posting_list_t::get_last_offset(posting_lists[0], field_is_array) : 255;
        uint8_t synonym_score = (is_synonym_query && demote_synonym_match) ? 0 : 1;
        Match single_token_match = Match(words_present, distance, max_offset, is_verbatim_match);
        match_score = single_token_match.get_match_score(total_cost, words_present, synonym_score);
        return;
    }
    std::map<size_t, std::vector<token_positions_t>> array_token_positions;
    posting_list_t::get_offsets(posting_lists, array_token_positions);
    for(const auto& kv: array_token_positions) {
        const std::vector<token_positions_t>& token_positions = kv.second;
        if(token_positions.empty()) continue;
        const Match& match = Match(seq_id, token_positions, false, P.prioritize_exact_match);
        uint8_t synonym_score = (is_synonym_query && demote_synonym_match) ? 0 : 1;
        uint64_t this_match_score = match.get_match_score(total_cost, posting_lists.size(), synonym_score);
        auto this_words_present = ((this_match_score >> 40) & 0xFF);
        auto unique_words = field_is_array ? this_words_present : ((this_match_score >> 32) & 0xFF);
        auto typo_score = ((this_match_score >> 24) & 0xFF);
        auto proximity = ((this_match_score >> 16) & 0xFF);
        auto verbatim = ((this_match_score >> 12) & 0xF);
        auto offset_score = P.prioritize_token_position ? ((this_match_score >> 4) & 0xFF) : 0;
        synonym_score = ((this_match_score >> 0) & 0xF);
        if(is_synonym_query && num_query_tokens == posting_lists.size()) {
            unique_words = syn_orig_num_tokens;
            this_words_present = syn_orig_num_tokens;
        }
        if(is_synonym_query && syn_orig_num_tokens > 0 && orig_num_tokens > 0) {
            double rel_factor = double(orig_num_tokens) / double(syn_orig_num_tokens);
            auto scale_component = [&](uint64_t v) -> uint64_t {
                double scaled = double(v) * rel_factor;
                if(scaled > 255.0) scaled = 255.0;
                return (uint64_t) scaled;
            };
            this_words_present = scale_component(this_words_present);
            unique_words = scale_component(unique_words);
            auto reversed_typo_score = 255 - typo_score;
            reversed_typo_score = scale_component(reversed_typo_score);
            typo_score = 255 - reversed_typo_score;
            auto reversed_proximity = 100 - proximity;
            reversed_proximity = scale_component(reversed_proximity);
            proximity = 100 - reversed_proximity;
            auto reversed_offset_score = 255 - offset_score;
            reversed_offset_score = scale_component(reversed_offset_score);
            offset_score = P.prioritize_token_position ? 255 - reversed_offset_score : 0;
        }
        uint64_t mod_match_score = ((int64_t(this_words_present) << 40) | (int64_t(unique_words) << 32) |
                                    (int64_t(typo_score) << 24) | (int64_t(proximity) << 16) |
                                    (int64_t(verbatim) << 12) | (int64_t(offset_score) << 4) |
                                    (int64_t(synonym_score) << 0));
        if(mod_match_score > (uint64_t) match_score) match_score = mod_match_score;
    }
}

// restated src/index.cpp:5227-5383 over reference or_iterator_t
static uint64_t refglue_compute_aggregated_score(const refglue_params& P, const std::vector<or_iterator_t>& its,
                                                 std::vector<or_iterator_t>& dropped_token_its, uint32_t seq_id) {
    const size_t num_search_fields = P.n_fields;
    std::vector<std::vector<posting_list_t::iterator_t>> field_to_tokens(num_search_fields);
    size_t query_len = 0;
    for(size_t ti = 0; ti < its.size(); ti++) {
        const auto& field_iters = its[ti].get_its();
        bool found_token = false;
        for(size_t fi = 0; fi < field_iters.size(); fi++) {
            const auto& field_iter = field_iters[fi];
            if(field_iter.valid() && field_iter.id() == seq_id && field_iter.get_field_id() < num_search_fields) {
                field_to_tokens[field_iter.get_field_id()].push_back(field_iter.clone());
                found_token = true;
            }
        }
        if(found_token) query_len++;
    }
    for(size_t ti = 0; ti < dropped_token_its.size(); ti++) {
        or_iterator_t& token_fields_iters = dropped_token_its[ti];
        if(token_fields_iters.skip_to(seq_id) && token_fields_iters.id() == seq_id) {
            const auto& field_iters = token_fields_iters.get_its();
            bool found_token = false;
            for(size_t fi = 0; fi < field_iters.size(); fi++) {
                const auto& field_iter = field_iters[fi];
                if(field_iter.id() == seq_id && field_iter.get_field_id() < num_search_fields) {
                    field_to_tokens[field_iter.get_field_id()].push_back(field_iter.clone());
                    found_token = true;
                }
            }
            if(found_token) query_len++;
        }
    }
    if(P.syn_orig_num_tokens != -1) query_len = P.syn_orig_num_tokens;

    int64_t best_field_match_score = 0, best_field_weight = 0;
    int64_t sum_field_weighted_score = 0;
    uint32_t num_matching_fields = 0;
    for(size_t fi = 0; fi < field_to_tokens.size(); fi++) {
        const auto& token_postings = field_to_tokens[fi];
        if(token_postings.empty()) continue;
        const int64_t field_weight = P.field_weight[fi];
        const bool field_is_array = P.field_is_array[fi];
        int64_t field_match_score = 0;
        bool single_exact_query_token = (P.total_cost == 0 && P.num_query_tokens == 1);
        refglue_score_results2(P, field_is_array, single_exact_query_token, field_match_score, seq_id, token_postings);
        if(P.match_type == 0 && field_match_score > best_field_match_score) {
            best_field_match_score = field_match_score;
            best_field_weight = field_weight;
        }
        if(P.match_type == 1 && field_weight > best_field_weight) {
            best_field_weight = field_weight;
            best_field_match_score = field_match_score;
        }
        if(P.match_type == 2) sum_field_weighted_score += (field_weight * field_match_score);
        num_matching_fields++;
    }
    query_len = (best_field_match_score == 0) ? 0 : std::min<size_t>(15, query_len);
    auto max_field_weight = std::min<size_t>(15 /*FIELD_MAX_WEIGHT*/, best_field_weight);
    num_matching_fields = std::min<size_t>(7, num_matching_fields);
    if(!P.prioritize_num_matching_fields) num_matching_fields = 0;
    uint64_t aggregated_score = 0;
    if(P.match_type == 0) {
        aggregated_score = ((int64_t(query_len) << 59) | (int64_t(best_field_match_score) << 11) |
                            (int64_t(max_field_weight) << 3) | (int64_t(num_matching_fields) << 0));
    } else if(P.match_type == 1) {
        aggregated_score = ((int64_t(query_len) << 59) | (int64_t(max_field_weight) << 51) |
                            (int64_t(best_field_match_score) << 3) | (int64_t(num_matching_fields) << 0));
    } else {
        aggregated_score = ((int64_t(query_len) << 59) | (int64_t(sum_field_weighted_score) << 3) |
                            (int64_t(num_matching_fields) << 0));
    }
    return aggregated_score;
}

// One token combination through the reference's or_iterator_t::intersect (include/or_iterator.h:61-181), the λ of
// src/index.cpp:5479-5551 reduced to (seq_id, aggregated_score).
// lists: (n_tokens + n_dropped) x n_fields posting_list_t* (nullptr where the token is absent from the field).
// filter: if use_fit != 0 the filter goes through the filter_result_iterator_t (lazy) path, else the raw-array path.
size_t ref_keyword_combo(const refglue_params* Pp, void** lists,
                         const uint32_t* excl, size_t n_excl, const uint32_t* filt, size_t n_filt, int use_fit,
                         uint32_t* out_ids, uint64_t* out_scores, size_t cap, uint64_t* out_num_keyword_matches) {
    const refglue_params& P = *Pp;
    search_cutoff = false;
    search_begin_us = 0;
    search_stop_us = UINT64_MAX;
    std::vector<or_iterator_t> token_its, dropped_its;
    for(uint32_t t = 0; t < P.n_tokens + P.n_dropped; t++) {
        std::vector<posting_list_t::iterator_t> its;
        for(uint32_t f = 0; f < P.n_fields; f++) {
            auto* pl = (posting_list_t*) lists[t * P.n_fields + f];
            if(!pl) continue;
            its.push_back(pl->new_iterator(nullptr, nullptr, f));
        }
        if(t < P.n_tokens) {
            if(its.empty()) continue;   // src/index.cpp:5648-5652: a token found in no field is skipped, not fatal
            token_its.push_back(or_iterator_t(its));
        } else {
            dropped_its.push_back(or_iterator_t(its));
        }
    }
    filter_result_iterator_t fit_none;
    filter_result_iterator_t fit_ids(filt, n_filt);
    size_t n = 0;
    bool overflow = false;
    auto fn = [&](single_filter_result_t& fr, const std::vector<or_iterator_t>& its) {
        uint32_t seq_id = fr.seq_id;
        uint64_t s = refglue_compute_aggregated_score(P, its, dropped_its, seq_id);
        if(n < cap) { out_ids[n] = seq_id; out_scores[n] = s; } else overflow = true;
        n++;
    };
    uint64_t nkm = 0;
    if(use_fit) {
        result_iter_state_t istate(excl, n_excl, (n_filt || use_fit == 2) ? &fit_ids : &fit_none);
        or_iterator_t::intersect(token_its, istate, fn);
        nkm = istate.num_keyword_matches;
    } else {
        result_iter_state_t istate(excl, n_excl, filt, n_filt);
        or_iterator_t::intersect(token_its, istate, fn);
        nkm = istate.num_keyword_matches;
    }
    if(out_num_keyword_matches) *out_num_keyword_matches = nkm;
    return overflow ? (size_t) -1 : n;
}

uint32_t ref_sizeof_params() { return (uint32_t) sizeof(refglue_params); }

// ---- the reference's ART (src/art.cpp, compiled in place): token index + fuzzy / prefix candidate search -------------
void* ref_art_new() { art_tree* t = new art_tree; art_tree_init(t); return t; }
void ref_art_free(void* t) { art_tree_destroy((art_tree*) t); delete (art_tree*) t; }
// one (token, document): what Index::index_field_in_memory does per token of a document (src/index.cpp:877, art_inserts)
void ref_art_insert(void* t, const char* token, uint32_t seq_id, int64_t score, const uint32_t* offsets, uint32_t n_offsets) {
    art_document doc(seq_id, score, std::vector<uint32_t>(offsets, offsets + n_offsets));
    art_insert((art_tree*) t, (const unsigned char*) token, (int) strlen(token) + 1, &doc);
}
// art_fuzzy_search_i as Index::fuzzy_search_fields calls it (src/index.cpp:4928-4952): term_len excludes the NUL for a
// prefix search; `exclude` (newline separated) seeds unique_tokens and receives the new tokens; the matching leaves' tokens
// are written to out, newline separated, in the order the reference returns them. Returns their count.
size_t ref_art_fuzzy(void* t, const char* term, int min_cost, int max_cost, size_t max_words, int token_order, int prefix, int last_token,
                     const char* prev_token, const uint32_t* filter_ids, size_t n_filter, int has_filter, const char* exclude,
                     char* out, size_t out_cap) {
    std::set<std::string> excl;
    for(const char* p = exclude; p && *p;) { const char* e = strchr(p, '\n'); std::string tok = e ? std::string(p, e) : std::string(p); if(!tok.empty()) excl.insert(tok); if(!e) break; p = e + 1; }
    filter_result_iterator_t none;
    filter_result_iterator_t some(filter_ids, n_filter);
    std::vector<art_leaf*> leaves;
    const int term_len = prefix ? (int) strlen(term) : (int) strlen(term) + 1;
    art_fuzzy_search_i((art_tree*) t, (const unsigned char*) term, term_len, min_cost, max_cost, max_words,
                       token_order == 1 ? MAX_SCORE : FREQUENCY, prefix != 0, last_token != 0, std::string(prev_token ? prev_token : ""),
                       has_filter ? &some : &none, leaves, excl);
    size_t w = 0;
    for(auto* l: leaves) {
        const size_t n = l->key_len - 1;
        if(w + n + 1 >= out_cap) break;
        memcpy(out + w, l->key, n); w += n; out[w++] = '\n';
    }
    if(out_cap) out[w < out_cap ? w : out_cap - 1] = 0;
    return leaves.size();
}


// Serialises the tree as it stands in memory (structure, compressed-path bytes and the per-node max_score exactly as the
// reference's inserts left them) so that an ART mirror can be loaded from it:
//   inner node: 'N', partial_len u8, partial[8], max_score i64, n_children u16, then per child (ascending byte): byte u8 + record
//   leaf:       'L', key_len u32, key bytes (with the trailing NUL), max_score i64, num_ids u32
//   empty tree: 'E'
// Returns the number of bytes needed; writes only while they fit in cap.
static void art_export_rec(const art_node* n, std::vector<unsigned char>& o) {
    auto put = [&](const void* p, size_t k) { const unsigned char* b = (const unsigned char*) p; o.insert(o.end(), b, b + k); };
    if(((uintptr_t) n) & 1) {
        const art_leaf* l = (const art_leaf*) (((uintptr_t) n) & ~(uintptr_t) 1);
        o.push_back('L');
        uint32_t kl = l->key_len; put(&kl, 4); put(l->key, kl);
        int64_t ms = l->max_score; put(&ms, 8);
        uint32_t df = posting_t::num_ids(l->values); put(&df, 4);
        return;
    }
    o.push_back('N');
    o.push_back(n->partial_len);
    put(n->partial, MAX_PREFIX_LEN);
    int64_t ms = n->max_score; put(&ms, 8);
    std::vector<std::pair<unsigned char, const art_node*>> kids;
    switch(n->type) {
        case NODE4:  for(int i = 0; i < n->num_children; i++) kids.push_back({((const art_node4*) n)->keys[i], ((const art_node4*) n)->children[i]}); break;
        case NODE16: for(int i = 0; i < n->num_children; i++) kids.push_back({((const art_node16*) n)->keys[i], ((const art_node16*) n)->children[i]}); break;
        case NODE48: for(int b = 0; b < 256; b++) { int ix = ((const art_node48*) n)->keys[b]; if(ix) kids.push_back({(unsigned char) b, ((const art_node48*) n)->children[ix - 1]}); } break;
        default:     for(int b = 0; b < 256; b++) if(((const art_node256*) n)->children[b]) kids.push_back({(unsigned char) b, ((const art_node256*) n)->children[b]}); break;
    }
    uint16_t nk = (uint16_t) kids.size(); put(&nk, 2);
    for(auto& k: kids) { o.push_back(k.first); art_export_rec(k.second, o); }
}
size_t ref_art_export(void* t, unsigned char* buf, size_t cap) {
    std::vector<unsigned char> o;
    const art_tree* tr = (const art_tree*) t;
    if(!tr->root) o.push_back('E'); else art_export_rec(tr->root, o);
    if(o.size() <= cap) memcpy(buf, o.data(), o.size());
    return o.size();
}

}  // extern "C"

// the one member the shim header leaves to this file (src/filter_result_iterator.cpp:2273, materialised-filter case)
bool filter_result_iterator_t::contains_atleast_one(const void* obj) {
    if(validity != valid) return false;
    return posting_t::contains_atleast_one(obj, remaining_ids(), remaining_count());
}

