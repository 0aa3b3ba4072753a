// TEST INFRASTRUCTURE — parity oracle (see ts_oracle.h). Standalone CPU restatement of the reference hot path on
// flat arrays. Never linked into the product; the product must fail loudly without its CUDA library.
//
// Sections:
//   1. sorted-set ops          ArrayUtils (include/array_utils.h:13-23), posting_list_t::intersect/merge
//   2. Match                   include/match_score.h:56-68, 129-275
//   3. offsets decode          posting_list_t::get_offsets (src/posting_list.cpp:832-916) & friends
//   4. scoring                 Index::score_results2 (src/index.cpp:6966-7098),
//                              Index::compute_aggregated_score (src/index.cpp:5227-5383)
//   5. AND/OR iteration        or_iterator_t (src/or_iterator.cpp:95-171), or_iterator_t::intersect
//                              (include/or_iterator.h:61-181), take_id (src/or_iterator.cpp:218-272)
//   6. sort keys + Topster     Index::compute_sort_scores (src/index.cpp:5722-5726, 5835-5836, 5864-5903),
//                              KV / Topster (include/topster.h:20-168, 321-473)
//   7. HNSW                    hnswlib (github.com/typesense/hnswlib, pinned 21de18ff… cmake/hnsw.cmake:3;
//                              NOT vendored — published algorithm restated), InnerProductSpace = 1 - dot
//   8. vector / hybrid fusion  src/index.cpp:3345-3445, 3645-3732, 4036-4221
#include "ts_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <bitset>
#include <queue>
#include <random>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ============================================================================ 1. sorted-set ops
// posting_list_t::intersect (src/posting_list.cpp:708-756): "advance all non-largest to the largest" zig-zag.
inline size_t gallop_to(const uint32_t* a, size_t lo, size_t n, uint32_t id) {
    // first index >= lo with a[idx] >= id (exponential then binary), the flat-array analogue of iterator_t::skip_to
    if(lo >= n || a[lo] >= id) return lo;
    size_t step = 1, prev = lo;
    size_t cur = lo + 1;
    while(cur < n && a[cur] < id) { prev = cur; step <<= 1; cur = lo + step; }
    size_t hi = std::min(cur, n);
    return std::lower_bound(a + prev + 1, a + hi, id) - a;
}

// ============================================================================ 2. Match
constexpr size_t WINDOW_SIZE = 10;                       // include/match_score.h:11
constexpr uint16_t MAX_DISPLACEMENT = 0xFFFF;            // include/match_score.h:12

struct TokPos {                  // token_positions_t (include/match_score.h:14-17) as a slice
    const uint16_t* pos;
    uint32_t n;
    bool last_token;
};

struct MatchOut { uint8_t words_present = 0, distance = 0, max_offset = 0, exact_match = 0; };

// include/match_score.h:56-68
inline uint64_t pack_match_score(uint8_t words_present, uint8_t distance, uint8_t max_offset, uint8_t exact_match,
                                 uint32_t total_cost, uint32_t unique_words, uint8_t synonym_score) {
    return (uint64_t) ((int64_t(words_present) << 40) | (int64_t(unique_words) << 32) |
                       (int64_t(255 - total_cost) << 24) | (int64_t(100 - distance) << 16) |
                       (int64_t(exact_match) << 12) | (int64_t(255 - max_offset) << 4) |
                       (int64_t(synonym_score) << 0));
}

// include/match_score.h:129-275 with populate_window=false. The window is kept as parallel small arrays; the
// reference re-sorts it descending by offset every iteration (sort2/sort3/std::sort — all yield the same multiset
// order on the offset key; ties between equal offsets cannot change any quantity computed below except which of
// two equal-offset cursors is popped, see note at the pop).
MatchOut match_window(const TokPos* toks, size_t n_all, bool check_exact_match) {
    MatchOut out;
    const size_t tokens_size = std::min(n_all, WINDOW_SIZE);
    struct W { uint8_t token_id; uint16_t offset; uint32_t offset_index; };
    W window[WINDOW_SIZE];
    size_t wsize = tokens_size;
    for(size_t t = 0; t < tokens_size; t++) window[t] = W{(uint8_t) t, toks[t].pos[0], 0};

    size_t best_num_match = 1;
    size_t best_displacement = MAX_DISPLACEMENT;
    int prev_min_offset = -1;

    while(wsize > 1) {
        // descending by offset. The reference uses sort2 / sort3 / std::sort(greater) depending on size; to follow
        // its tie behaviour exactly we restate the two small sorts and use an insertion-free std::sort otherwise.
        if(wsize == 2) {
            if(window[0].offset < window[1].offset) std::swap(window[0], window[1]);
        } else if(wsize == 3) {
            W* a = window;
            if(a[0].offset > a[1].offset) {
                if(a[1].offset > a[2].offset) {
                } else if(a[0].offset > a[2].offset) {
                    std::swap(a[1], a[2]);
                } else {
                    W tmp = a[0]; a[0] = a[2]; a[2] = a[1]; a[1] = tmp;
                }
            } else {
                if(a[0].offset > a[2].offset) {
                    std::swap(a[0], a[1]);
                } else if(a[2].offset > a[1].offset) {
                    std::swap(a[0], a[2]);
                } else {
                    W tmp = a[0]; a[0] = a[1]; a[1] = a[2]; a[2] = tmp;
                }
            }
        } else {
            std::sort(window, window + wsize, [](const W& x, const W& y) { return x.offset > y.offset; });
        }

        size_t min_offset = window[wsize - 1].offset;
        if(int(min_offset) < prev_min_offset) break;     // wrap-around guard (:164-167)
        prev_min_offset = (int) min_offset;

        size_t this_displacement = 0, this_num_match = 0;
        for(size_t i = 0; i < wsize; i++) {
            if((size_t) (window[i].offset - min_offset) <= WINDOW_SIZE) {
                uint16_t next_offset = (i == wsize - 1) ? window[i].offset : window[i + 1].offset;
                this_displacement += window[i].offset - next_offset;
                this_num_match++;
            }
        }

        if((this_num_match > best_num_match) ||
           (this_num_match == best_num_match && this_displacement < best_displacement)) {
            best_displacement = this_displacement;
            best_num_match = this_num_match;
            out.max_offset = (uint8_t) std::min<uint16_t>(255, window[0].offset);
        }

        if(best_num_match == tokens_size && best_displacement == (wsize - 1)) break;

        const W smallest = window[wsize - 1];
        wsize--;
        const TokPos& tp = toks[smallest.token_id];
        if(smallest.offset == tp.pos[tp.n - 1]) continue;          // no more offsets for this token
        uint32_t next_index = smallest.offset_index + 1;
        window[wsize++] = W{smallest.token_id, tp.pos[next_index], next_index};
    }

    if(best_displacement == MAX_DISPLACEMENT) best_displacement = 0;
    out.words_present = (uint8_t) best_num_match;
    out.distance = uint8_t(best_displacement);
    out.exact_match = 0;

    if(check_exact_match) {
        if(out.distance > n_all - 1) return out;
        int last_token_index = -1;
        size_t total_offsets = 0;
        for(size_t t = 0; t < n_all; t++) {
            if(toks[t].last_token && toks[t].n != 0) last_token_index = toks[t].pos[toks[t].n - 1];
            total_offsets += toks[t].n;
            if(total_offsets > n_all && out.distance == n_all - 1) return out;
        }
        if(last_token_index == int(n_all) - 1) {
            if(total_offsets == n_all && out.distance == n_all - 1) out.exact_match = 1;
            else if(out.distance < n_all - 1) out.exact_match = 1;
        }
    }
    return out;
}

// posting_list_t::found_token_sequence / has_phrase_match (src/posting_list.cpp:1719-1789)
bool found_token_sequence(const TokPos* toks, size_t n, size_t token_index, uint16_t target_pos) {
    if(token_index == n) return true;
    bool found_pos = false;
    int prev_pos = -1;
    for(uint32_t i = 0; i < toks[token_index].n; i++) {
        uint16_t tok_pos = toks[token_index].pos[i];
        if(tok_pos < prev_pos) { found_pos = false; break; }
        if(tok_pos == target_pos) { found_pos = true; break; }
        prev_pos = tok_pos;
    }
    if(!found_pos) return false;
    return found_token_sequence(toks, n, token_index + 1, (uint16_t) (target_pos + 1));
}

bool has_phrase_match(const TokPos* toks, size_t n) {
    int prev_pos = -1;
    for(uint32_t i = 0; i < toks[0].n; i++) {
        uint16_t pos = toks[0].pos[i];
        if(pos < prev_pos) return false;
        if(found_token_sequence(toks, n, 1, (uint16_t) (pos + 1))) return true;
        prev_pos = pos;
    }
    return false;
}

// ============================================================================ 3. offsets decode
struct OffSlice { const uint32_t* p; uint32_t n; };      // raw reference-encoded offsets of one (token, doc)

// posting_list_t::get_offsets (src/posting_list.cpp:832-916): decode every token's raw offsets into
// array_index -> [token_positions_t], tokens appended in input order. `store` owns the uint16 positions.
struct DecodedField {
    std::vector<uint16_t> store;
    struct Ent { size_t array_index; uint32_t start, n; bool last_token; };
    std::vector<Ent> ents;                               // in (token order, then array element order)
};

void decode_offsets(const OffSlice* toks, size_t n_toks, DecodedField& out) {
    out.store.clear();
    out.ents.clear();
    for(size_t j = 0; j < n_toks; j++) {
        const uint32_t* offsets = toks[j].p;
        uint32_t start_offset = 0, end_offset = toks[j].n;
        uint32_t cur_start = (uint32_t) out.store.size();
        int prev_pos = -1;
        bool is_last_token = false;
        while(start_offset < end_offset) {
            int pos = (int) offsets[start_offset];
            start_offset++;
            if(pos == 0) {                                // token is the last token of the doc
                is_last_token = true;
                start_offset++;
                continue;
            }
            if(pos == prev_pos) {                         // end of an array element
                if(out.store.size() != cur_start) {
                    size_t array_index = (size_t) offsets[start_offset];
                    is_last_token = false;
                    if(start_offset + 1 < end_offset) {
                        size_t next_offset = (size_t) offsets[start_offset + 1];
                        if(next_offset == 0) { is_last_token = true; start_offset++; }
                    }
                    out.ents.push_back({array_index, cur_start, (uint32_t) (out.store.size() - cur_start), is_last_token});
                    cur_start = (uint32_t) out.store.size();
                }
                start_offset++;
                prev_pos = -1;
                continue;
            }
            prev_pos = pos;
            out.store.push_back((uint16_t) ((uint16_t) pos - 1));
        }
        if(out.store.size() != cur_start) {               // plain string fields
            out.ents.push_back({0, cur_start, (uint32_t) (out.store.size() - cur_start), is_last_token});
        }
    }
}

// posting_list_t::is_single_token_verbatim_match (src/posting_list.cpp:918-959)
bool is_single_token_verbatim_match(OffSlice s, bool field_is_array) {
    const uint32_t* offsets = s.p;
    uint32_t start_offset = 0, end_offset = s.n;
    if(s.n == 0) return false;
    if(!field_is_array && offsets[start_offset] != 1) return false;
    if(field_is_array) {
        int prev_pos = -1;
        while(start_offset < end_offset) {
            int pos = (int) offsets[start_offset];
            start_offset++;
            if(pos == prev_pos && pos == 1 && start_offset + 1 < end_offset && offsets[start_offset + 1] == 0) return true;
            prev_pos = pos;
        }
        return false;
    } else if((end_offset - start_offset) == 2 && offsets[end_offset - 1] == 0) {
        return true;
    }
    return false;
}

// posting_list_t::get_last_offset (src/posting_list.cpp:1899-1950)
size_t get_last_offset(OffSlice s, bool field_is_array) {
    const uint32_t* offsets = s.p;
    uint32_t end_offset = s.n;
    if(s.n == 0) return 0;
    if(field_is_array) {
        uint32_t start_offset = 0;
        int prev_pos = -1;
        size_t max_offset = 0;
        while(start_offset < end_offset) {
            int pos = (int) offsets[start_offset];
            start_offset++;
            if((size_t) pos > max_offset) max_offset = (size_t) pos;
            if(pos == prev_pos) {
                if(start_offset + 1 < end_offset) {
                    size_t next_offset = (size_t) offsets[start_offset + 1];
                    if(next_offset == 0) start_offset++;
                }
                start_offset++;
                prev_pos = -1;
                continue;
            }
            prev_pos = pos;
        }
        return max_offset;
    }
    return offsets[end_offset - 1] == 0 ? (end_offset >= 2 ? offsets[end_offset - 2] : 0) : offsets[end_offset - 1];
}

// ============================================================================ 4. scoring
struct ScoreParams {
    uint32_t total_cost;
    uint32_t num_query_tokens;
    int syn_orig_num_tokens, orig_num_tokens;
    bool is_synonym_query, demote_synonym_match;
    bool prioritize_exact_match, prioritize_token_position, prioritize_num_matching_fields;
    uint8_t match_type;
};

// Index::score_results2 (src/index.cpp:6966-7098). `toks` = this field's matched tokens for the doc, query order.
int64_t score_field(const ScoreParams& P, bool field_is_array, bool single_exact_query_token,
                    const OffSlice* toks, size_t n_toks, DecodedField& scratch) {
    int64_t match_score = 0;
    if(n_toks <= 1) {
        const uint8_t is_verbatim_match = uint8_t(P.prioritize_exact_match && single_exact_query_token &&
                                                  is_single_token_verbatim_match(toks[0], field_is_array));
        size_t words_present = (P.num_query_tokens == 1 && P.is_synonym_query) ? P.syn_orig_num_tokens : 1;
        size_t distance = (P.num_query_tokens == 1 && P.is_synonym_query) ? P.syn_orig_num_tokens - 1 : 0;
        size_t max_offset = P.prioritize_token_position ? get_last_offset(toks[0], field_is_array) : 255;
        uint8_t synonym_score = (P.is_synonym_query && P.demote_synonym_match) ? 0 : 1;
        return (int64_t) pack_match_score((uint8_t) words_present, (uint8_t) distance, (uint8_t) max_offset,
                                          is_verbatim_match, P.total_cost, (uint32_t) words_present, synonym_score);
    }
    decode_offsets(toks, n_toks, scratch);
    // group by array index ascending (std::map iteration order), tokens in input order within a group
    std::vector<size_t> keys;
    for(auto& e: scratch.ents) keys.push_back(e.array_index);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    TokPos tp[64];
    for(size_t key: keys) {
        size_t n = 0;
        for(auto& e: scratch.ents) {
            if(e.array_index == key && n < 64) tp[n++] = TokPos{scratch.store.data() + e.start, e.n, e.last_token};
        }
        if(n == 0) continue;
        MatchOut m = match_window(tp, n, P.prioritize_exact_match);
        uint8_t synonym_score = (P.is_synonym_query && P.demote_synonym_match) ? 0 : 1;
        uint64_t this_match_score = pack_match_score(m.words_present, m.distance, m.max_offset, m.exact_match,
                                                     P.total_cost, (uint32_t) n_toks, synonym_score);
        uint64_t this_words_present = ((this_match_score >> 40) & 0xFF);
        uint64_t unique_words = field_is_array ? this_words_present : ((this_match_score >> 32) & 0xFF);
        uint64_t typo_score = ((this_match_score >> 24) & 0xFF);
        uint64_t proximity = ((this_match_score >> 16) & 0xFF);
        uint64_t verbatim = ((this_match_score >> 12) & 0xF);
        uint64_t offset_score = P.prioritize_token_position ? ((this_match_score >> 4) & 0xFF) : 0;
        uint64_t syn = ((this_match_score >> 0) & 0xF);
        if(P.is_synonym_query && P.num_query_tokens == n_toks) {
            unique_words = (uint64_t) P.syn_orig_num_tokens;
            this_words_present = (uint64_t) P.syn_orig_num_tokens;
        }
        if(P.is_synonym_query && P.syn_orig_num_tokens > 0 && P.orig_num_tokens > 0) {
            double rel_factor = double(P.orig_num_tokens) / double(P.syn_orig_num_tokens);
            auto scale_component = [&](uint64_t v) -> uint64_t {
                double scaled = double(v) * rel_factor;
                if(scaled > 255.0) scaled = 255.0;
                return (uint64_t) scaled;
            };
            this_words_present = scale_component(this_words_present);
            unique_words = scale_component(unique_words);
            uint64_t reversed_typo_score = 255 - typo_score;
            reversed_typo_score = scale_component(reversed_typo_score);
            typo_score = 255 - reversed_typo_score;
            uint64_t reversed_proximity = 100 - proximity;
            reversed_proximity = scale_component(reversed_proximity);
            proximity = 100 - reversed_proximity;
            uint64_t reversed_offset_score = 255 - offset_score;
            reversed_offset_score = scale_component(reversed_offset_score);
            offset_score = P.prioritize_token_position ? 255 - reversed_offset_score : 0;
        }
        uint64_t mod_match_score = (uint64_t) ((int64_t(this_words_present) << 40) | (int64_t(unique_words) << 32) |
                                               (int64_t(typo_score) << 24) | (int64_t(proximity) << 16) |
                                               (int64_t(verbatim) << 12) | (int64_t(offset_score) << 4) |
                                               (int64_t(syn) << 0));
        if(mod_match_score > (uint64_t) match_score) match_score = (int64_t) mod_match_score;
    }
    return match_score;
}

// Index::compute_aggregated_score (src/index.cpp:5227-5383), reduction part: per-field scores -> one u64.
// field_tok[f] = slices of the tokens (required then dropped, query order) present in field f for this doc.
uint64_t aggregate_score(const ScoreParams& P, size_t n_fields, const uint8_t* field_weight,
                         const uint8_t* field_is_array, size_t query_len_in,
                         const std::vector<OffSlice>* field_tok, DecodedField& scratch) {
    size_t query_len = query_len_in;
    if(P.syn_orig_num_tokens != -1) query_len = (size_t) P.syn_orig_num_tokens;
    int64_t best_field_match_score = 0, best_field_weight = 0, sum_field_weighted_score = 0;
    uint32_t num_matching_fields = 0;
    for(size_t fi = 0; fi < n_fields; fi++) {
        if(field_tok[fi].empty()) continue;
        const int64_t fw = field_weight[fi];
        bool single_exact_query_token = (P.total_cost == 0 && P.num_query_tokens == 1);
        int64_t field_match_score = score_field(P, field_is_array[fi], single_exact_query_token,
                                                field_tok[fi].data(), field_tok[fi].size(), scratch);
        if(P.match_type == TSO_MATCH_MAX_SCORE && field_match_score > best_field_match_score) {
            best_field_match_score = field_match_score;
            best_field_weight = fw;
        }
        if(P.match_type == TSO_MATCH_MAX_WEIGHT && fw > best_field_weight) {
            best_field_weight = fw;
            best_field_match_score = field_match_score;
        }
        if(P.match_type == TSO_MATCH_SUM_SCORE) sum_field_weighted_score += (fw * field_match_score);
        num_matching_fields++;
    }
    query_len = (best_field_match_score == 0) ? 0 : std::min<size_t>(15, query_len);
    size_t max_field_weight = std::min<size_t>(15, (size_t) best_field_weight);     // FIELD_MAX_WEIGHT include/index.h:669
    num_matching_fields = (uint32_t) std::min<size_t>(7, num_matching_fields);
    if(!P.prioritize_num_matching_fields) num_matching_fields = 0;
    if(P.match_type == TSO_MATCH_MAX_SCORE) {
        return (uint64_t) ((int64_t(query_len) << 59) | (int64_t(best_field_match_score) << 11) |
                           (int64_t(max_field_weight) << 3) | (int64_t(num_matching_fields) << 0));
    } else if(P.match_type == TSO_MATCH_MAX_WEIGHT) {
        return (uint64_t) ((int64_t(query_len) << 59) | (int64_t(max_field_weight) << 51) |
                           (int64_t(best_field_match_score) << 3) | (int64_t(num_matching_fields) << 0));
    }
    return (uint64_t) ((int64_t(query_len) << 59) | (int64_t(sum_field_weighted_score) << 3) |
                       (int64_t(num_matching_fields) << 0));
}

// ============================================================================ index object
struct Index {
    uint32_t n_docs = 0;
    std::vector<tso_field> fields;
    std::vector<const int64_t*> sort_cols;
    tso_hnsw hnsw{};
    bool has_hnsw = false;
};

// ============================================================================ 5. AND / OR iteration
struct ListCur {                 // posting_list_t::iterator_t on a flat list
    const uint32_t* ids = nullptr;
    size_t n = 0, i = 0;
    uint64_t base = 0;           // global posting index of ids[0] (-> pos_off)
    uint32_t field = 0;          // field slot
    bool valid() const { return i < n; }
    uint32_t id() const { return ids[i]; }
    void next() { i++; }
    void skip_to(uint32_t v) { i = gallop_to(ids, i, n, v); }
};

struct OrIt {                    // or_iterator_t (src/or_iterator.cpp:95-171); exhausted lists are erased
    std::vector<ListCur> its;
    int curr_index = 0;
    void init() {
        for(size_t i = 0; i < its.size();) { if(!its[i].valid()) its.erase(its.begin() + i); else i++; }
        curr_index = 0;
        for(size_t i = 1; i < its.size(); i++) if(its[i].id() < its[curr_index].id()) curr_index = (int) i;
    }
    bool valid() const { return !its.empty(); }
    uint32_t id() const { return its[curr_index].id(); }
    bool next() {                                                   // advance_smallest
        if(its.empty()) return false;
        uint32_t smallest_value = its[curr_index].id();
        curr_index = 0;
        for(int i = 0; i < int(its.size()); i++) {
            if(its[i].id() == smallest_value) its[i].next();
            if(!its[i].valid()) { its.erase(its.begin() + i); i--; }
        }
        uint32_t new_smallest = UINT32_MAX;
        for(int i = 0; i < int(its.size()); i++) {
            if(its[i].id() < new_smallest) { curr_index = i; new_smallest = its[i].id(); }
        }
        return !its.empty();
    }
    bool skip_to(uint32_t v) {
        uint32_t current_value = UINT32_MAX;
        curr_index = 0;
        for(size_t i = 0; i < its.size(); i++) {
            its[i].skip_to(v);
            if(!its[i].valid()) { its.erase(its.begin() + i); i--; }
            else if(its[i].id() < current_value) { curr_index = (int) i; current_value = its[i].id(); }
        }
        return !its.empty();
    }
};

struct IState {                  // result_iter_state_t raw-array form (include/posting_list.h:13-45)
    const uint32_t* excl = nullptr; size_t n_excl = 0;
    const uint32_t* filt = nullptr; size_t n_filt = 0;
    size_t filt_index = 0;
    uint64_t num_keyword_matches = 0;
    bool is_filter_provided() const { return n_filt > 0; }
    bool is_filter_valid() const { return n_filt > 0 && filt_index < n_filt; }
    uint32_t get_filter_id() const { return (n_filt > 0 && filt_index < n_filt) ? filt[filt_index] : 0; }
};

// or_iterator_t::take_id (src/or_iterator.cpp:218-272), raw filter-array branch
bool take_id(IState& st, uint32_t id, bool& is_excluded) {
    is_excluded = false;
    if(st.n_excl != 0 && std::binary_search(st.excl, st.excl + st.n_excl, id)) { is_excluded = true; return false; }
    if(st.n_filt != 0) {
        if(st.filt_index >= st.n_filt) return false;
        size_t found_index = std::lower_bound(st.filt + st.filt_index, st.filt + st.n_filt, id) - st.filt;
        if(found_index == st.n_filt) { st.filt_index = found_index + 1; return false; }
        if(st.filt[found_index] == id) { st.filt_index = found_index + 1; return true; }
        st.filt_index = found_index;
        return false;
    }
    return true;
}

// or_iterator_t::intersect (include/or_iterator.h:61-181): the three size-specialised loops are the same algorithm;
// restated once. Any exhausted or_iterator ends the loop (its.size() == it_size guard).
template <class F>
void or_intersect(std::vector<OrIt>& its, IState& st, F func) {
    if(its.empty()) return;
    auto at_end = [&]() { for(auto& it: its) if(!it.valid()) return true; return false; };
    auto skip_all_to_filter = [&]() -> bool {
        uint32_t fid = st.get_filter_id();
        bool ok = true;
        for(auto& it: its) { it.skip_to(fid); }
        for(auto& it: its) if(!it.valid()) ok = false;
        return ok;
    };
    if(st.is_filter_provided() && st.is_filter_valid()) skip_all_to_filter();
    bool is_excluded;
    while(!at_end()) {
        bool equal = true;
        for(size_t i = 0; i + 1 < its.size(); i++) if(its[i].id() != its[i + 1].id()) { equal = false; break; }
        if(equal) {
            uint32_t id = its[0].id();
            st.num_keyword_matches++;
            if(take_id(st, id, is_excluded)) func(id);
            if(st.is_filter_provided() && !is_excluded) {
                if(st.is_filter_valid()) skip_all_to_filter();
                else break;
            } else {
                for(auto& it: its) it.next();                     // advance_all
            }
        } else {
            uint32_t greatest = 0;                                // advance_non_largest
            for(auto& it: its) if(it.id() > greatest) greatest = it.id();
            for(auto& it: its) if(it.id() != greatest) it.skip_to(greatest);
        }
    }
}

// ============================================================================ 6. sort keys + Topster
int64_t float_to_int64(float f) {                                 // src/index.cpp:266-274
    int32_t i;
    memcpy(&i, &f, sizeof i);
    if(i < 0) i ^= INT32_MAX;
    return i;
}
float int64_to_float(int64_t n) {                                 // src/index.cpp:276-286
    int32_t i = (int32_t) n;
    if(i < 0) i ^= INT32_MAX;
    float f;
    memcpy(&f, &i, sizeof f);
    return f;
}

struct SortSpec { uint8_t type[3]; int32_t col[3]; int8_t order[3]; uint8_t missing_first[3]; };

// Index::compute_sort_scores (src/index.cpp:5662-5907): _text_match / _seq_id / numeric / _vector_distance
void compute_sort_scores(const Index& ix, const SortSpec& S, uint32_t seq_id, int64_t max_field_match_score,
                         int64_t* scores, int64_t& match_score_index, float vector_distance) {
    for(int i = 0; i < 3; i++) {
        if(S.type[i] == TSO_SORT_NONE) continue;
        if(S.type[i] == TSO_SORT_TEXT_MATCH) { scores[i] = max_field_match_score; match_score_index = i; }
        else if(S.type[i] == TSO_SORT_SEQ_ID) scores[i] = seq_id;
        else if(S.type[i] == TSO_SORT_VECTOR_DISTANCE) scores[i] = float_to_int64(vector_distance);
        else {
            int64_t v = ix.sort_cols[S.col[i]][seq_id];           // INT64_MIN = missing (default_score)
            scores[i] = v;
            if(scores[i] == INT64_MIN && S.missing_first[i]) {
                bool is_asc = (S.order[i] == -1);
                scores[i] = is_asc ? (INT64_MIN + 1) : INT64_MAX;
            }
        }
        if(S.order[i] == -1) scores[i] = (int64_t) (0 - (uint64_t) scores[i]);
    }
}

inline bool kv_greater(const tso_kv& a, const tso_kv& b) {         // KV::is_greater include/topster.h:146-149
    if(a.scores[0] != b.scores[0]) return a.scores[0] > b.scores[0];
    if(a.scores[1] != b.scores[1]) return a.scores[1] > b.scores[1];
    if(a.scores[2] != b.scores[2]) return a.scores[2] > b.scores[2];
    return a.key > b.key;
}
inline bool kv_smaller(const tso_kv& a, const tso_kv& b) {         // KV::is_smaller :151-154
    if(a.scores[0] != b.scores[0]) return a.scores[0] < b.scores[0];
    if(a.scores[1] != b.scores[1]) return a.scores[1] < b.scores[1];
    if(a.scores[2] != b.scores[2]) return a.scores[2] < b.scores[2];
    return a.key < b.key;
}

// Topster<KV> non-grouped path (include/topster.h:242-473): min-heap of slots + key->slot map.
struct Topster {
    uint32_t MAX_SIZE, size = 0;
    std::vector<tso_kv> data;          // storage
    std::vector<uint32_t> kvs;         // heap position -> data slot
    std::vector<uint32_t> array_index; // data slot -> heap position
    std::unordered_map<uint64_t, uint32_t> map;   // key -> data slot
    explicit Topster(uint32_t cap): MAX_SIZE(cap), data(cap), kvs(cap), array_index(cap) {
        for(uint32_t i = 0; i < cap; i++) { kvs[i] = i; array_index[i] = i; }
    }
    void swap_me(uint32_t a, uint32_t b) {
        std::swap(kvs[a], kvs[b]);
        array_index[kvs[a]] = a;
        array_index[kvs[b]] = b;
    }
    const tso_kv& at(uint32_t heap_pos) const { return data[kvs[heap_pos]]; }
    tso_kv& at(uint32_t heap_pos) { return data[kvs[heap_pos]]; }
    int add(const tso_kv& kv) {
        bool less_than_min_heap = (size >= MAX_SIZE) && kv_smaller(kv, at(0));
        if(less_than_min_heap) return 0;
        size_t heap_op_index = 0;
        bool SIFT_DOWN = true;
        auto found_it = map.find(kv.key);
        if(found_it != map.end()) {
            const tso_kv& existing = data[found_it->second];
            if(kv_smaller(kv, existing)) return 0;
            SIFT_DOWN = true;
            heap_op_index = array_index[found_it->second];
            map.erase(at((uint32_t) heap_op_index).key);
        } else {
            if(size < MAX_SIZE) { SIFT_DOWN = false; heap_op_index = size; size++; }
            else { SIFT_DOWN = true; heap_op_index = 0; map.erase(at(0).key); }
        }
        map.emplace(kv.key, kvs[heap_op_index]);
        at((uint32_t) heap_op_index) = kv;
        if(SIFT_DOWN) {
            while((2 * heap_op_index + 1) < size) {
                uint32_t next = (uint32_t) (2 * heap_op_index + 1);
                if(next + 1 < size && kv_greater(at(next), at(next + 1))) next++;
                if(kv_greater(at((uint32_t) heap_op_index), at(next))) swap_me((uint32_t) heap_op_index, next);
                else break;
                heap_op_index = next;
            }
        } else {
            while(heap_op_index > 0) {
                uint32_t parent = (uint32_t) ((heap_op_index - 1) / 2);
                if(kv_greater(at(parent), at((uint32_t) heap_op_index))) {
                    swap_me((uint32_t) heap_op_index, parent);
                    heap_op_index = parent;
                } else break;
            }
        }
        return 1;
    }
    void sort() {                                                   // stable_sort(kvs, kvs+size, is_greater)
        std::stable_sort(kvs.begin(), kvs.begin() + size, [&](uint32_t a, uint32_t b) { return kv_greater(data[a], data[b]); });
        for(uint32_t i = 0; i < size; i++) array_index[kvs[i]] = i;
    }
};

// ---------------------------------------------------------------------------- one token combination
struct ComboCtx {
    const Index* ix;
    const tso_kw_batch* b;
    uint32_t q, c;
    ScoreParams P;
    uint32_t F;
    uint8_t field_is_array[TSO_MAX_FIELDS];
    const uint8_t* field_weight;
};

template <class Emit>
uint64_t run_combo(const Index& ix, const tso_kw_batch& b, uint32_t q, uint32_t c, Emit emit) {
    const uint32_t F = b.n_fields;
    const uint32_t row0 = b.c_tok_off[c], row1 = b.c_tok_off[c + 1];
    const uint32_t n_rows = row1 - row0;
    const uint32_t n_req = b.c_n_required[c];
    ScoreParams P;
    P.total_cost = b.c_total_cost[c];
    P.num_query_tokens = b.q_num_query_tokens[q];
    P.syn_orig_num_tokens = b.c_syn_orig_num_tokens ? b.c_syn_orig_num_tokens[c] : -1;
    P.orig_num_tokens = b.c_orig_num_tokens ? b.c_orig_num_tokens[c] : -1;
    uint8_t cf = b.c_flags ? b.c_flags[c] : 0;
    P.is_synonym_query = cf & TSO_CFLAG_SYNONYM;
    P.demote_synonym_match = cf & TSO_CFLAG_DEMOTE_SYNONYM;
    uint8_t qf = b.q_flags[q];
    P.prioritize_exact_match = qf & TSO_FLAG_PRIORITIZE_EXACT_MATCH;
    P.prioritize_token_position = qf & TSO_FLAG_PRIORITIZE_TOKEN_POSITION;
    P.prioritize_num_matching_fields = qf & TSO_FLAG_PRIORITIZE_NUM_MATCHING_FIELDS;
    P.match_type = b.q_match_type[q];
    uint8_t field_is_array[TSO_MAX_FIELDS];
    for(uint32_t f = 0; f < F; f++) field_is_array[f] = (uint8_t) ix.fields[b.field_ids[f]].is_array;

    // get_field_token_its (src/index.cpp:5598-5660): one or_iterator per token over the fields that have it;
    // a token found in no field is skipped.
    auto make_or = [&](uint32_t row) {
        OrIt o;
        for(uint32_t f = 0; f < F; f++) {
            uint32_t li = b.t_list[(size_t) row * F + f];
            if(li == TSO_NO_LIST) continue;
            const tso_field& fld = ix.fields[b.field_ids[f]];
            ListCur lc;
            lc.ids = fld.ids + fld.list_off[li];
            lc.n = (size_t) (fld.list_off[li + 1] - fld.list_off[li]);
            lc.base = fld.list_off[li];
            lc.field = f;
            o.its.push_back(lc);
        }
        o.init();
        return o;
    };
    std::vector<OrIt> token_its, dropped_its;
    for(uint32_t r = 0; r < n_rows; r++) {
        OrIt o = make_or(row0 + r);
        if(r < n_req) { if(o.its.empty()) continue; token_its.push_back(std::move(o)); }
        else dropped_its.push_back(std::move(o));
    }
    IState st;
    st.excl = b.excl_ids + b.q_excl_off[q];
    st.n_excl = b.q_excl_off[q + 1] - b.q_excl_off[q];
    int32_t fs = b.q_filter[q];
    bool filter_given = fs >= 0;
    if(filter_given) {
        st.filt = b.filter_ids + b.filter_off[fs];
        st.n_filt = (size_t) (b.filter_off[fs + 1] - b.filter_off[fs]);
        if(st.n_filt == 0) return 0;          // fuzzy_search_fields early return (src/index.cpp:4823-4826)
    }

    std::vector<OffSlice> field_tok[TSO_MAX_FIELDS];
    DecodedField scratch;
    auto slice_of = [&](const ListCur& lc) {
        const tso_field& fld = ix.fields[b.field_ids[lc.field]];
        uint64_t p = lc.base + lc.i;
        return OffSlice{fld.positions + fld.pos_off[p], (uint32_t) (fld.pos_off[p + 1] - fld.pos_off[p])};
    };
    or_intersect(token_its, st, [&](uint32_t seq_id) {
        for(uint32_t f = 0; f < F; f++) field_tok[f].clear();
        size_t query_len = 0;
        for(auto& tok: token_its) {
            bool found = false;
            for(auto& lc: tok.its) {
                if(lc.valid() && lc.id() == seq_id) { field_tok[lc.field].push_back(slice_of(lc)); found = true; }
            }
            if(found) query_len++;
        }
        for(auto& tok: dropped_its) {
            if(tok.skip_to(seq_id) && tok.id() == seq_id) {
                bool found = false;
                for(auto& lc: tok.its) {
                    if(lc.id() == seq_id) { field_tok[lc.field].push_back(slice_of(lc)); found = true; }
                }
                if(found) query_len++;
            }
        }
        uint64_t agg = aggregate_score(P, F, b.q_field_weight + (size_t) q * F, field_is_array, query_len,
                                       field_tok, scratch);
        emit(seq_id, agg);
    });
    return st.num_keyword_matches;
}

SortSpec sort_spec_of(const tso_kw_batch& b, uint32_t q) {
    SortSpec S;
    for(int i = 0; i < 3; i++) {
        S.type[i] = b.q_sort_type[q * 3 + i];
        S.col[i] = b.q_sort_col[q * 3 + i];
        S.order[i] = b.q_sort_order[q * 3 + i];
        S.missing_first[i] = b.q_sort_missing_first ? b.q_sort_missing_first[q * 3 + i] : 0;
    }
    return S;
}

// search_all_candidates (src/index.cpp:1794-1894) + the λ of search_across_fields (:5479-5551) for one query:
// every combination into one Topster; result ids OR-ed into all_result_ids.
uint16_t keyword_query(const Index& ix, const tso_kw_batch& b, uint32_t q, Topster& topster,
                       std::vector<uint32_t>& all_result_ids) {
    SortSpec S = sort_spec_of(b, q);
    uint16_t searched_queries = 0;
    std::vector<uint32_t> id_buff;
    for(uint32_t c = b.q_combo_off[q]; c < b.q_combo_off[q + 1]; c++) {
        size_t before = id_buff.size();
        run_combo(ix, b, q, c, [&](uint32_t seq_id, uint64_t agg) {
            tso_kv kv{};
            int64_t match_score_index = -1;
            compute_sort_scores(ix, S, seq_id, (int64_t) agg, kv.scores, match_score_index, 0);
            kv.key = seq_id;
            kv.distinct_key = seq_id;
            kv.query_index = searched_queries;
            kv.match_score_index = (int8_t) match_score_index;
            kv.vector_distance = -1.0f;
            kv.text_match_score = 0;
            if(match_score_index >= 0) kv.text_match_score = kv.scores[match_score_index];   // KV ctor :40-42
            if(match_score_index != -1) { kv.scores[match_score_index] = (int64_t) agg; kv.text_match_score = (int64_t) agg; }
            topster.add(kv);
            id_buff.push_back(seq_id);
        });
        if(id_buff.size() != before) searched_queries++;
    }
    std::sort(id_buff.begin(), id_buff.end());
    id_buff.erase(std::unique(id_buff.begin(), id_buff.end()), id_buff.end());
    all_result_ids = std::move(id_buff);
    return searched_queries;
}

void write_topster(Topster& t, tso_kv* out, uint32_t stride, uint32_t* out_count) {
    t.sort();
    uint32_t n = std::min<uint32_t>(t.size, stride);
    for(uint32_t i = 0; i < n; i++) out[i] = t.at(i);
    *out_count = n;
}

template <class Fn>
void parallel_for(uint32_t n, uint32_t n_threads, Fn fn) {
    if(n_threads <= 1 || n <= 1) { for(uint32_t i = 0; i < n; i++) fn(i); return; }
    std::atomic<uint32_t> next{0};
    std::vector<std::thread> th;
    n_threads = std::min(n_threads, n);
    for(uint32_t t = 0; t < n_threads; t++) {
        th.emplace_back([&]() { for(;;) { uint32_t i = next.fetch_add(1); if(i >= n) break; fn(i); } });
    }
    for(auto& x: th) x.join();
}

}  // namespace

#include "ts_oracle_vec.inc"

// ============================================================================ C API
extern "C" {

size_t tso_facet_counts(uint32_t n_docs, uint32_t n_values, const uint64_t* doc_off, const uint32_t* value_ids,
                        const uint32_t* result_ids, size_t n, uint32_t sample_mod, tso_facet_count* out, size_t cap, uint32_t* out_distinct) {
    // Index::do_facets, hash-index branch (src/index.cpp:1674-1780) over ONE batch of ascending result ids, plain facet (no
    // range / facet query / group-by): per doc the distinct facet ids it holds; count += 1, doc_id = this doc, array_pos =
    // position of the id in the doc's list; `estimate_facets` keeps every sample_mod-th result (i % mod == 0). Then
    // Collection::search's order: (count, id) descending (include/collection.h:552-554).
    std::vector<uint32_t> cnt(n_values, 0), doc(n_values, 0), pos(n_values, 0);
    for(size_t i = 0; i < n; i++) {
        if(sample_mod > 1 && (i % sample_mod) != 0) continue;
        const uint32_t d = result_ids[i];
        if(d >= n_docs) continue;
        const uint64_t o0 = doc_off[d], o1 = doc_off[d + 1];
        for(uint64_t j = o0; j < o1; j++) {
            const uint32_t v = value_ids[j];
            if(v >= n_values) continue;
            bool dup = false;
            for(uint64_t k = o0; k < j && !dup; k++) dup = value_ids[k] == v;
            if(dup) continue;
            cnt[v]++; doc[v] = d; pos[v] = (uint32_t) (j - o0);
        }
    }
    std::vector<tso_facet_count> all;
    for(uint32_t v = 0; v < n_values; v++) if(cnt[v]) all.push_back({v, cnt[v], doc[v], pos[v]});
    std::sort(all.begin(), all.end(), [](const tso_facet_count& a, const tso_facet_count& b) {
        return a.count != b.count ? a.count > b.count : a.value_id > b.value_id; });
    if(out_distinct) *out_distinct = (uint32_t) all.size();
    const size_t m = std::min(cap, all.size());
    for(size_t i = 0; i < m; i++) out[i] = all[i];
    return m;
}

int tso_contains_atleast_one(const uint32_t* list, size_t n_list, const uint32_t* target_ids, size_t n_targets) {
    // posting_list_t::contains_atleast_one (src/posting_list.cpp:1090-1112): two ascending sequences, the smaller head advances
    // (the list side by skip_to, here a gallop)
    size_t li = 0, ti = 0;
    while(ti < n_targets && li < n_list) {
        const uint32_t id = list[li];
        if(id == target_ids[ti]) return 1;
        if(id > target_ids[ti]) { while(ti < n_targets && target_ids[ti] < id) ti++; }
        else li = gallop_to(list, li, n_list, target_ids[ti]);
    }
    return 0;
}

size_t tso_intersect(uint32_t k, const uint32_t* const* lists, const size_t* lens, uint32_t* out, size_t cap) {
    // posting_list_t::intersect (src/posting_list.cpp:708-756)
    if(k == 0) return 0;
    size_t n = 0;
    if(k == 1) { for(size_t i = 0; i < lens[0]; i++) { if(n >= cap) return (size_t) -1; out[n++] = lists[0][i]; } return n; }
    std::vector<size_t> pos(k, 0);
    auto at_end = [&]() { for(uint32_t i = 0; i < k; i++) if(pos[i] >= lens[i]) return true; return false; };
    while(!at_end()) {
        bool eq = true;
        for(uint32_t i = 0; i + 1 < k; i++) if(lists[i][pos[i]] != lists[i + 1][pos[i + 1]]) { eq = false; break; }
        if(eq) {
            if(n >= cap) return (size_t) -1;
            out[n++] = lists[0][pos[0]];
            for(uint32_t i = 0; i < k; i++) pos[i]++;
        } else {
            uint32_t greatest = 0;
            for(uint32_t i = 0; i < k; i++) greatest = std::max(greatest, lists[i][pos[i]]);
            for(uint32_t i = 0; i < k; i++) if(lists[i][pos[i]] != greatest) pos[i] = gallop_to(lists[i], pos[i], lens[i], greatest);
        }
    }
    return n;
}

size_t tso_merge(uint32_t k, const uint32_t* const* lists, const size_t* lens, uint32_t* out, size_t cap) {
    // posting_list_t::merge (src/posting_list.cpp:638-705): unique ascending union
    std::vector<size_t> pos(k, 0);
    size_t n = 0;
    for(;;) {
        uint32_t smallest = UINT32_MAX; bool any = false;
        for(uint32_t i = 0; i < k; i++) if(pos[i] < lens[i]) { any = true; smallest = std::min(smallest, lists[i][pos[i]]); }
        if(!any) break;
        if(n >= cap) return (size_t) -1;
        out[n++] = smallest;
        for(uint32_t i = 0; i < k; i++) if(pos[i] < lens[i] && lists[i][pos[i]] == smallest) pos[i]++;
    }
    return n;
}

size_t tso_and_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out) {
    size_t i = 0, j = 0, n = 0;
    while(i < na && j < nb) { if(a[i] < b[j]) i++; else if(b[j] < a[i]) j++; else { out[n++] = a[i]; i++; j++; } }
    return n;
}
size_t tso_or_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out) {
    size_t i = 0, j = 0, n = 0;
    while(i < na && j < nb) {
        uint32_t v;
        if(a[i] < b[j]) v = a[i++]; else if(b[j] < a[i]) v = b[j++]; else { v = a[i]; i++; j++; }
        if(n == 0 || out[n - 1] != v) out[n++] = v;
    }
    while(i < na) { if(n == 0 || out[n - 1] != a[i]) out[n++] = a[i]; i++; }
    while(j < nb) { if(n == 0 || out[n - 1] != b[j]) out[n++] = b[j]; j++; }
    return n;
}
size_t tso_exclude_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out) {
    size_t i = 0, j = 0, n = 0;
    while(i < na) {
        while(j < nb && b[j] < a[i]) j++;
        if(j < nb && b[j] == a[i]) { i++; continue; }
        out[n++] = a[i++];
    }
    return n;
}

void tso_match(uint32_t n_tokens, const uint32_t* tok_off, const uint16_t* positions, const uint8_t* last_token,
               int check_exact, uint8_t out[4]) {
    std::vector<TokPos> tp(n_tokens);
    for(uint32_t t = 0; t < n_tokens; t++) tp[t] = TokPos{positions + tok_off[t], tok_off[t + 1] - tok_off[t], last_token[t] != 0};
    MatchOut m = match_window(tp.data(), n_tokens, check_exact != 0);
    out[0] = m.words_present; out[1] = m.distance; out[2] = m.max_offset; out[3] = m.exact_match;
}
uint64_t tso_match_score(uint8_t words_present, uint8_t distance, uint8_t max_offset, uint8_t exact,
                         uint32_t total_cost, uint32_t unique_words, uint8_t syn) {
    return pack_match_score(words_present, distance, max_offset, exact, total_cost, unique_words, syn);
}
int tso_has_phrase_match(uint32_t n_tokens, const uint32_t* tok_off, const uint16_t* positions) {
    std::vector<TokPos> tp(n_tokens);
    for(uint32_t t = 0; t < n_tokens; t++) tp[t] = TokPos{positions + tok_off[t], tok_off[t + 1] - tok_off[t], false};
    return has_phrase_match(tp.data(), n_tokens);
}

void* tso_index_new(uint32_t n_docs) { auto* ix = new Index(); ix->n_docs = n_docs; return ix; }
void tso_index_free(void* idx) { delete (Index*) idx; }
int tso_index_add_field(void* idx, const tso_field* f) { auto* ix = (Index*) idx; ix->fields.push_back(*f); return (int) ix->fields.size() - 1; }
void tso_index_set_field(void* idx, int field, const tso_field* f) { auto* ix = (Index*) idx; ix->fields[(size_t) field] = *f; }
int tso_index_add_sort_column(void* idx, const int64_t* vals) { auto* ix = (Index*) idx; ix->sort_cols.push_back(vals); return (int) ix->sort_cols.size() - 1; }
void tso_index_set_hnsw(void* idx, const tso_hnsw* g) { auto* ix = (Index*) idx; ix->hnsw = *g; ix->has_hnsw = true; }

size_t tso_keyword_combo(void* idx, const tso_kw_batch* b, uint32_t q, uint32_t c, uint32_t* out_ids,
                         uint64_t* out_scores, size_t cap, uint64_t* out_num_keyword_matches) {
    size_t n = 0; bool overflow = false;
    uint64_t nkm = run_combo(*(Index*) idx, *b, q, c, [&](uint32_t id, uint64_t s) {
        if(n < cap) { out_ids[n] = id; out_scores[n] = s; } else overflow = true;
        n++;
    });
    if(out_num_keyword_matches) *out_num_keyword_matches = nkm;
    return overflow ? (size_t) -1 : n;
}

int tso_keyword_search_batch(void* idx, const tso_kw_batch* b, tso_kv* out_kv, uint32_t kv_stride,
                             uint32_t* out_count, uint32_t* out_found, uint32_t n_threads) {
    const Index& ix = *(Index*) idx;
    parallel_for(b->n_queries, n_threads, [&](uint32_t q) {
        Topster t(std::max<uint32_t>(1, b->q_topk[q]));
        std::vector<uint32_t> all_ids;
        keyword_query(ix, *b, q, t, all_ids);
        write_topster(t, out_kv + (size_t) q * kv_stride, kv_stride, &out_count[q]);
        out_found[q] = (uint32_t) all_ids.size();
    });
    return 0;
}

static int wildcard_common(void* idx, const tso_kw_batch* b, const int64_t* id_scores, tso_kv* out_kv, uint32_t kv_stride,
                           uint32_t* out_count, uint32_t* out_found, uint32_t n_threads);
int tso_wildcard_search_batch(void* idx, const tso_kw_batch* b, tso_kv* out_kv, uint32_t kv_stride,
                              uint32_t* out_count, uint32_t* out_found, uint32_t n_threads) {
    return wildcard_common(idx, b, nullptr, out_kv, kv_stride, out_count, out_found, n_threads);
}
// Index::do_phrase_search's Topster loop for a phrase-only query (src/index.cpp:6039-6076): ids with their phrase match scores
int tso_scored_ids_search_batch(void* idx, const tso_kw_batch* b, const int64_t* id_scores, tso_kv* out_kv, uint32_t kv_stride,
                                uint32_t* out_count, uint32_t* out_found, uint32_t n_threads) {
    return wildcard_common(idx, b, id_scores, out_kv, kv_stride, out_count, out_found, n_threads);
}
static int wildcard_common(void* idx, const tso_kw_batch* b, const int64_t* id_scores, tso_kv* out_kv, uint32_t kv_stride,
                           uint32_t* out_count, uint32_t* out_found, uint32_t n_threads) {
    const Index& ix = *(Index*) idx;
    parallel_for(b->n_queries, n_threads, [&](uint32_t q) {
        Topster t(std::max<uint32_t>(1, b->q_topk[q]));
        SortSpec S = sort_spec_of(*b, q);
        const uint32_t* excl = b->excl_ids + b->q_excl_off[q];
        const size_t n_excl = b->q_excl_off[q + 1] - b->q_excl_off[q];
        const int32_t fs = b->q_filter[q];
        const uint32_t* filt = nullptr; size_t n = ix.n_docs;
        if(fs >= 0) { filt = b->filter_ids + b->filter_off[fs]; n = (size_t) (b->filter_off[fs + 1] - b->filter_off[fs]); }
        uint32_t found = 0;
        for(size_t i = 0; i < n; i++) {
            const uint32_t seq_id = filt ? filt[i] : (uint32_t) i;
            if(n_excl && std::binary_search(excl, excl + n_excl, seq_id)) continue;      // get_n_ids skips excluded ids
            tso_kv kv{};
            int64_t msi = -1;
            const int64_t ms = (id_scores && filt) ? id_scores[(filt - b->filter_ids) + i] : 100;
            compute_sort_scores(ix, S, seq_id, ms, kv.scores, msi, 0);                   // src/index.cpp:6727-6729 / 6045-6053
            kv.key = seq_id; kv.distinct_key = seq_id; kv.query_index = 0;
            kv.match_score_index = (int8_t) msi;
            kv.vector_distance = -1.0f;
            kv.text_match_score = msi >= 0 ? kv.scores[msi] : 0;
            t.add(kv);
            found++;
        }
        write_topster(t, out_kv + (size_t) q * kv_stride, kv_stride, &out_count[q]);
        out_found[q] = found;
    });
    return 0;
}

uint32_t tso_topster_run(uint32_t capacity, const tso_kv* in, uint32_t n, tso_kv* out) {
    Topster t(capacity);
    for(uint32_t i = 0; i < n; i++) t.add(in[i]);
    t.sort();
    for(uint32_t i = 0; i < t.size; i++) out[i] = t.at(i);
    return t.size;
}

size_t tso_phrase_matches(void* idx, uint32_t field, const uint32_t* lists, uint32_t k,
                          const uint32_t* ids, size_t n, uint32_t* out) {
    // posting_list_t::get_phrase_matches (src/posting_list.cpp:1791-1825)
    const Index& ix = *(Index*) idx;
    const tso_field& fld = ix.fields[field];
    size_t n_out = 0;
    if(k == 1) { for(size_t i = 0; i < n; i++) out[n_out++] = ids[i]; return n_out; }
    std::vector<ListCur> its(k);
    for(uint32_t j = 0; j < k; j++) {
        its[j].ids = fld.ids + fld.list_off[lists[j]];
        its[j].n = (size_t) (fld.list_off[lists[j] + 1] - fld.list_off[lists[j]]);
        its[j].base = fld.list_off[lists[j]];
    }
    DecodedField scratch;
    std::vector<OffSlice> sl;
    for(size_t i = 0; i < n; i++) {
        uint32_t id = ids[i];
        sl.clear();
        for(uint32_t j = 0; j < k; j++) {
            its[j].skip_to(id);
            // reference decodes whatever the iterator points at, even if it is not `id`; ids passed here always
            // come from the intersection of the same lists (src/index.cpp:5956-5960), so the iterator is on `id`.
            if(its[j].valid()) {
                uint64_t p = its[j].base + its[j].i;
                sl.push_back(OffSlice{fld.positions + fld.pos_off[p], (uint32_t) (fld.pos_off[p + 1] - fld.pos_off[p])});
            }
        }
        decode_offsets(sl.data(), sl.size(), scratch);
        std::vector<size_t> keys;
        for(auto& e: scratch.ents) keys.push_back(e.array_index);
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        for(size_t key: keys) {
            std::vector<TokPos> tp;
            for(auto& e: scratch.ents) if(e.array_index == key) tp.push_back(TokPos{scratch.store.data() + e.start, e.n, e.last_token});
            if(tp.size() == k && has_phrase_match(tp.data(), tp.size())) { out[n_out++] = id; break; }
        }
    }
    return n_out;
}

// posting_list_t::get_exact_matches (src/posting_list.cpp:1281-1452) when `exact`, get_prefix_matches (:1129-1279)
// otherwise. Restated per id; the iterator is expected to sit on `id` (callers pass ids of the lists' intersection),
// ids missing from a list are skipped.
static bool positional_match_one(const std::vector<OffSlice>& sl, bool field_is_array, bool exact) {
    const size_t k = sl.size();
    if(k == 1) {
        if(exact) return is_single_token_verbatim_match(sl[0], field_is_array);
        return sl[0].n != 0 && sl[0].p[0] == 1;                    // is_single_token_prefix_match (:1114-1127)
    }
    if(!field_is_array) {
        bool is_match = true;
        for(int j = (int) k - 1; j >= 0; j--) {
            const uint32_t* offsets = sl[j].p;
            size_t start = 0, end = sl[j].n;
            if(end == 0) return false;
            if(exact && j == (int) k - 1) {
                if(offsets[end - 1] != 0 || (end >= 2 && offsets[end - 2] != k)) { is_match = false; break; }
            }
            while(start < end) {
                uint32_t offset = offsets[start++];
                if(offset == (uint32_t) (j + 1)) { is_match = true; break; }
                if(offset > (uint32_t) (j + 1)) { is_match = false; break; }
            }
            if(!is_match) break;
        }
        return is_match;
    }
    struct Meta { std::bitset<128> token_index; bool has_last_token = false; };
    std::map<size_t, Meta> by_elem;
    for(int j = (int) k - 1; j >= 0; j--) {
        const uint32_t* offsets = sl[j].p;
        size_t start = 0, end = sl[j].n;
        int prev_pos = -1;
        bool any_last = false, found = false;
        size_t n_matching = 0;
        while(start < end) {
            int pos = (int) offsets[start++];
            if(pos == prev_pos) {
                if(start >= end) break;
                size_t array_index = offsets[start];
                if(exact && start + 1 < end && offsets[start + 1] == 0 && (size_t) pos == k) {
                    by_elem[array_index].has_last_token = true;
                    any_last = true;
                    start++;
                }
                if(found && j + 1 < 128) by_elem[array_index].token_index.set(j + 1);
                start++;
                prev_pos = -1;
                found = false;
                continue;
            }
            if(pos == j + 1) { found = true; n_matching++; }
            prev_pos = pos;
        }
        if(exact && j == (int) k - 1 && !any_last) return false;
        if(n_matching == 0) return false;
    }
    for(auto& kv: by_elem) if(kv.second.token_index.count() == k && (!exact || kv.second.has_last_token)) return true;
    return false;
}

static size_t positional_matches(void* idx, uint32_t field, const uint32_t* lists, uint32_t k,
                                 const uint32_t* ids, size_t n, uint32_t* out, bool exact) {
    const Index& ix = *(Index*) idx;
    const tso_field& fld = ix.fields[field];
    std::vector<ListCur> its(k);
    for(uint32_t j = 0; j < k; j++) {
        its[j].ids = fld.ids + fld.list_off[lists[j]];
        its[j].n = (size_t) (fld.list_off[lists[j] + 1] - fld.list_off[lists[j]]);
        its[j].base = fld.list_off[lists[j]];
    }
    size_t n_out = 0;
    std::vector<OffSlice> sl;
    for(size_t i = 0; i < n; i++) {
        sl.clear();
        bool on_id = true;
        for(uint32_t j = 0; j < k; j++) {
            its[j].skip_to(ids[i]);
            if(!its[j].valid() || its[j].id() != ids[i]) { on_id = false; break; }
            uint64_t p = its[j].base + its[j].i;
            sl.push_back(OffSlice{fld.positions + fld.pos_off[p], (uint32_t) (fld.pos_off[p + 1] - fld.pos_off[p])});
        }
        if(on_id && positional_match_one(sl, fld.is_array != 0, exact)) out[n_out++] = ids[i];
    }
    return n_out;
}

size_t tso_exact_matches(void* idx, uint32_t field, const uint32_t* lists, uint32_t k,
                         const uint32_t* ids, size_t n, uint32_t* out) {
    return positional_matches(idx, field, lists, k, ids, n, out, true);
}
size_t tso_prefix_matches(void* idx, uint32_t field, const uint32_t* lists, uint32_t k,
                          const uint32_t* ids, size_t n, uint32_t* out) {
    return positional_matches(idx, field, lists, k, ids, n, out, false);
}

int64_t tso_float_to_int64(float f) { return float_to_int64(f); }
float tso_int64_to_float(int64_t v) { return int64_to_float(v); }

}  // extern "C"
