// Per-document scoring for the keyword hot path, written once as __host__ __device__ code.
//
// Device use: called per thread from kw_search_kernel (kw_kernels.cu) on the offsets of one matched document.
// Host use:   tests/hostsim compiles the same header with g++ to check this logic against the oracle on the CPU
//             (this container has no GPU). The host build is test-only: libtsgpu.so never runs it.
//
// Replaces (reference file:line):
//   posting_list_t::get_offsets                 src/posting_list.cpp:832-916   -> seg_next()
//   posting_list_t::is_single_token_verbatim_match   src/posting_list.cpp:918-959
//   posting_list_t::get_last_offset             src/posting_list.cpp:1899-1950
//   Match::Match / get_match_score              include/match_score.h:129-275, 56-68  -> match_window()
//   Index::score_results2                       src/index.cpp:6966-7098        -> score_field()
//   Index::compute_aggregated_score (reduce)    src/index.cpp:5291-5383        -> aggregate_fields()
//   Index::compute_sort_scores + float_to_int64_t   src/index.cpp:5662-5907, 266-274
//
// Design: nothing is materialised. The reference copies every token's offsets into vector<uint16_t> keyed by a
// std::map<array_index,...>; here a token's positions for one array element are a contiguous slice of the raw
// offsets already in HBM, addressed through a small cursor, and array elements are visited by a k-way merge over the
// tokens' segment cursors (segments are stored in ascending array-index order, src/index.cpp:1357-1393).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TS_HD __host__ __device__ __forceinline__
#define TS_HD_NOINLINE __host__ __device__
#else
#define TS_HD inline
#define TS_HD_NOINLINE inline
#endif

namespace tsdev {

constexpr int kWindowSize = 10;            // WINDOW_SIZE include/match_score.h:11
constexpr int kMaxTokens = 16;             // TSGPU_MAX_TOKENS
constexpr uint32_t kMaxDisplacement = 0xFFFF;

struct ScoreParams {
    uint32_t total_cost;
    uint32_t num_query_tokens;
    int32_t  syn_orig_num_tokens;
    int32_t  orig_num_tokens;
    uint8_t  is_synonym_query;
    uint8_t  demote_synonym_match;
    uint8_t  prioritize_exact_match;
    uint8_t  prioritize_token_position;
    uint8_t  prioritize_num_matching_fields;
    uint8_t  match_type;
};

// raw reference-encoded offsets of one (token, doc)
struct RawTok {
    const uint32_t* p;
    uint32_t n;
};

// one array element's worth of a token's positions: positions are (uint16)(p[start+i]) - 1 for i < count
struct Seg {
    uint32_t array_index;
    uint32_t start;
    uint32_t count;
    uint8_t  last_token;
    uint8_t  valid;
};

// Streaming form of posting_list_t::get_offsets for ONE token: *cursor walks the raw stream and returns the next
// emitted (array_index, positions, last_token) group, or valid=0 at the end. State carried between calls mirrors the
// reference's locals: `carry_last` is is_last_token as left by earlier groups of the same token.
struct SegCursor {
    uint32_t pos;          // next raw index to read
    uint8_t  is_last;      // is_last_token carried across groups (it is only reset when a group is emitted)
};

TS_HD Seg seg_next(const RawTok& t, SegCursor& c) {
    Seg s;
    s.valid = 0; s.array_index = 0; s.start = c.pos; s.count = 0; s.last_token = 0;
    uint32_t start_offset = c.pos;
    const uint32_t end_offset = t.n;
    int64_t prev_pos = -1;
    uint32_t npos = 0;              // positions.size()
    uint32_t first = start_offset;  // raw index of positions[0]
    bool is_last_token = c.is_last;
    while(start_offset < end_offset) {
        uint32_t pos = t.p[start_offset];
        start_offset++;
        if(pos == 0) {
            is_last_token = true;
            start_offset++;
            continue;
        }
        if((int64_t) pos == prev_pos) {
            if(npos != 0) {
                uint32_t array_index = (start_offset < end_offset) ? t.p[start_offset] : 0;
                is_last_token = false;
                if(start_offset + 1 < end_offset) {
                    if(t.p[start_offset + 1] == 0) { is_last_token = true; start_offset++; }
                }
                start_offset++;
                s.valid = 1; s.array_index = array_index; s.start = first; s.count = npos;
                s.last_token = is_last_token;
                c.pos = start_offset; c.is_last = is_last_token;
                return s;
            }
            start_offset++;
            prev_pos = -1;
            continue;
        }
        prev_pos = pos;
        if(npos == 0) first = start_offset - 1;
        npos++;
    }
    c.pos = start_offset;
    c.is_last = is_last_token;
    if(npos != 0) {      // plain string field
        s.valid = 1; s.array_index = 0; s.start = first; s.count = npos; s.last_token = is_last_token;
    }
    return s;
}

TS_HD uint16_t seg_pos(const RawTok& t, const Seg& s, uint32_t i) {
    return (uint16_t) ((uint16_t) t.p[s.start + i] - 1);
}

TS_HD bool is_single_token_verbatim_match(const RawTok& t, bool field_is_array) {
    if(t.n == 0) return false;
    uint32_t start_offset = 0, end_offset = t.n;
    if(!field_is_array && t.p[0] != 1) return false;
    if(field_is_array) {
        int64_t prev_pos = -1;
        while(start_offset < end_offset) {
            uint32_t pos = t.p[start_offset];
            start_offset++;
            if((int64_t) pos == prev_pos && pos == 1 && start_offset + 1 < end_offset && t.p[start_offset + 1] == 0) return true;
            prev_pos = pos;
        }
        return false;
    }
    return (end_offset - start_offset) == 2 && t.p[end_offset - 1] == 0;
}

TS_HD uint32_t get_last_offset(const RawTok& t, bool field_is_array) {
    if(t.n == 0) return 0;
    const uint32_t end_offset = t.n;
    if(field_is_array) {
        uint32_t start_offset = 0;
        int64_t prev_pos = -1;
        uint32_t max_offset = 0;
        while(start_offset < end_offset) {
            uint32_t pos = t.p[start_offset];
            start_offset++;
            if(pos > max_offset) max_offset = pos;
            if((int64_t) pos == prev_pos) {
                if(start_offset + 1 < end_offset && t.p[start_offset + 1] == 0) start_offset++;
                start_offset++;
                prev_pos = -1;
                continue;
            }
            prev_pos = pos;
        }
        return max_offset;
    }
    if(t.p[end_offset - 1] == 0) return end_offset >= 2 ? t.p[end_offset - 2] : 0;
    return t.p[end_offset - 1];
}

TS_HD uint64_t pack_match_score(uint32_t words_present, uint32_t distance, uint32_t max_offset, uint32_t exact_match,
                                uint32_t total_cost, uint32_t unique_words, uint32_t synonym_score) {
    // all inputs are uint8 in the reference (Match fields) except total_cost / unique_words (uint32)
    return (uint64_t) (((int64_t) (words_present & 0xFF) << 40) | ((int64_t) unique_words << 32) |
                       ((int64_t) (uint32_t) (255u - total_cost) << 24) | ((int64_t) (100 - (int32_t) (distance & 0xFF)) << 16) |
                       ((int64_t) (exact_match & 0xFF) << 12) | ((int64_t) (255 - (int32_t) (max_offset & 0xFF)) << 4) |
                       ((int64_t) (synonym_score & 0xFF)));
}

struct MatchOut { uint32_t words_present, distance, max_offset, exact_match; };

// Match::Match(doc, token_offsets, populate_window=false, check_exact_match) on n_all tokens given as
// (RawTok, Seg) pairs.
TS_HD MatchOut match_window(const RawTok* toks, const Seg* segs, int n_all, bool check_exact_match) {
    MatchOut out;
    out.words_present = 0; out.distance = 0; out.max_offset = 0; out.exact_match = 0;
    const int tokens_size = n_all < kWindowSize ? n_all : kWindowSize;
    uint8_t  w_tok[kWindowSize];
    uint16_t w_off[kWindowSize];
    uint32_t w_idx[kWindowSize];
    int wsize = tokens_size;
    for(int t = 0; t < tokens_size; t++) { w_tok[t] = (uint8_t) t; w_off[t] = seg_pos(toks[t], segs[t], 0); w_idx[t] = 0; }

    uint32_t best_num_match = 1;
    uint32_t best_displacement = kMaxDisplacement;
    int32_t prev_min_offset = -1;

    while(wsize > 1) {
        // sort descending by offset (insertion sort; ties can only involve duplicate query tokens whose position
        // lists are identical, so tie order cannot change the result — see DESIGN.md)
        for(int i = 1; i < wsize; i++) {
            uint8_t tk = w_tok[i]; uint16_t of = w_off[i]; uint32_t ix = w_idx[i];
            int j = i - 1;
            while(j >= 0 && w_off[j] < of) { w_tok[j + 1] = w_tok[j]; w_off[j + 1] = w_off[j]; w_idx[j + 1] = w_idx[j]; j--; }
            w_tok[j + 1] = tk; w_off[j + 1] = of; w_idx[j + 1] = ix;
        }
        const uint32_t min_offset = w_off[wsize - 1];
        if((int32_t) min_offset < prev_min_offset) break;
        prev_min_offset = (int32_t) min_offset;

        uint32_t this_displacement = 0, this_num_match = 0;
        for(int i = 0; i < wsize; i++) {
            if((uint32_t) (w_off[i] - min_offset) <= (uint32_t) kWindowSize) {
                uint16_t next_offset = (i == wsize - 1) ? w_off[i] : w_off[i + 1];
                this_displacement += (uint32_t) (w_off[i] - next_offset);
                this_num_match++;
            }
        }
        if((this_num_match > best_num_match) ||
           (this_num_match == best_num_match && this_displacement < best_displacement)) {
            best_displacement = this_displacement;
            best_num_match = this_num_match;
            out.max_offset = w_off[0] < 255 ? w_off[0] : 255;
        }
        if(best_num_match == (uint32_t) tokens_size && best_displacement == (uint32_t) (wsize - 1)) break;

        const uint8_t tk = w_tok[wsize - 1];
        const uint16_t sm_off = w_off[wsize - 1];
        const uint32_t sm_idx = w_idx[wsize - 1];
        wsize--;
        const Seg& sg = segs[tk];
        if(sm_off == seg_pos(toks[tk], sg, sg.count - 1)) continue;     // no more offsets for this token
        const uint32_t next_index = sm_idx + 1;
        w_tok[wsize] = tk; w_off[wsize] = seg_pos(toks[tk], sg, next_index); w_idx[wsize] = next_index;
        wsize++;
    }
    if(best_displacement == kMaxDisplacement) best_displacement = 0;
    out.words_present = best_num_match & 0xFF;
    out.distance = best_displacement & 0xFF;
    out.exact_match = 0;
    if(check_exact_match) {
        const uint32_t nm1 = (uint32_t) (n_all - 1);
        if(out.distance > nm1) return out;
        int32_t last_token_index = -1;
        uint32_t total_offsets = 0;
        for(int t = 0; t < n_all; t++) {
            if(segs[t].last_token && segs[t].count != 0) last_token_index = seg_pos(toks[t], segs[t], segs[t].count - 1);
            total_offsets += segs[t].count;
            if(total_offsets > (uint32_t) n_all && out.distance == nm1) return out;
        }
        if(last_token_index == (int32_t) n_all - 1) {
            if(total_offsets == (uint32_t) n_all && out.distance == nm1) out.exact_match = 1;
            else if(out.distance < nm1) out.exact_match = 1;
        }
    }
    return out;
}

TS_HD uint64_t scale_component(uint64_t v, double rel_factor) {
    double scaled = (double) v * rel_factor;
    if(scaled > 255.0) scaled = 255.0;
    return (uint64_t) scaled;
}

// Index::score_results2 for one field. toks[0..n_toks) = the field's matched tokens for this doc in query order.
TS_HD_NOINLINE int64_t score_field(const ScoreParams& P, bool field_is_array, bool single_exact_query_token,
                                   const RawTok* toks, int n_toks) {
    if(n_toks <= 1) {
        const uint32_t is_verbatim_match = (P.prioritize_exact_match && single_exact_query_token &&
                                            is_single_token_verbatim_match(toks[0], field_is_array)) ? 1 : 0;
        const bool syn1 = (P.num_query_tokens == 1 && P.is_synonym_query);
        const uint32_t words_present = syn1 ? (uint32_t) P.syn_orig_num_tokens : 1;     // size_t -> uint8 at Match ctor
        const uint32_t distance = syn1 ? (uint32_t) (P.syn_orig_num_tokens - 1) : 0;
        const uint32_t max_offset = P.prioritize_token_position ? get_last_offset(toks[0], field_is_array) : 255;
        const uint32_t synonym_score = (P.is_synonym_query && P.demote_synonym_match) ? 0 : 1;
        return (int64_t) pack_match_score(words_present & 0xFF, distance & 0xFF, max_offset & 0xFF, is_verbatim_match,
                                          P.total_cost, words_present, synonym_score);
    }
    int64_t match_score = 0;
    SegCursor cur[kMaxTokens];
    Seg nxt[kMaxTokens];
    for(int t = 0; t < n_toks; t++) { cur[t].pos = 0; cur[t].is_last = 0; nxt[t] = seg_next(toks[t], cur[t]); }
    // visit array indices in ascending order (std::map iteration in the reference)
    for(;;) {
        uint32_t key = 0xFFFFFFFFu; bool any = false;
        for(int t = 0; t < n_toks; t++) if(nxt[t].valid) { if(!any || nxt[t].array_index < key) key = nxt[t].array_index; any = true; }
        if(!any) break;
        RawTok gt[kMaxTokens]; Seg gs[kMaxTokens];
        int n = 0;
        // a token can emit several groups with the same array index only for malformed streams; the reference
        // would append them all in token order. Follow that: drain every group of this key, token by token.
        for(int t = 0; t < n_toks; t++) {
            while(nxt[t].valid && nxt[t].array_index == key) {
                if(n < kMaxTokens) { gt[n] = toks[t]; gs[n] = nxt[t]; n++; }
                nxt[t] = seg_next(toks[t], cur[t]);
                // groups of one token arrive in ascending array index; if the next one is smaller it was already
                // passed by the merge and would have been placed under its own (smaller) key in the reference's map
                // — that only happens for malformed input and is not reproduced.
            }
        }
        if(n == 0) continue;
        MatchOut m = match_window(gt, gs, n, P.prioritize_exact_match != 0);
        uint32_t synonym_score0 = (P.is_synonym_query && P.demote_synonym_match) ? 0 : 1;
        uint64_t this_match_score = pack_match_score(m.words_present, m.distance, m.max_offset, m.exact_match,
                                                     P.total_cost, (uint32_t) n_toks, synonym_score0);
        uint64_t this_words_present = ((this_match_score >> 40) & 0xFF);
        uint64_t unique_words = field_is_array ? this_words_present : ((this_match_score >> 32) & 0xFF);
        uint64_t typo_score = ((this_match_score >> 24) & 0xFF);
        uint64_t proximity = ((this_match_score >> 16) & 0xFF);
        uint64_t verbatim = ((this_match_score >> 12) & 0xF);
        uint64_t offset_score = P.prioritize_token_position ? ((this_match_score >> 4) & 0xFF) : 0;
        uint64_t syn = ((this_match_score >> 0) & 0xF);
        if(P.is_synonym_query && P.num_query_tokens == (uint32_t) n_toks) {
            unique_words = (uint64_t) (int64_t) P.syn_orig_num_tokens;
            this_words_present = (uint64_t) (int64_t) P.syn_orig_num_tokens;
        }
        if(P.is_synonym_query && P.syn_orig_num_tokens > 0 && P.orig_num_tokens > 0) {
            double rel_factor = (double) P.orig_num_tokens / (double) P.syn_orig_num_tokens;
            this_words_present = scale_component(this_words_present, rel_factor);
            unique_words = scale_component(unique_words, rel_factor);
            uint64_t reversed_typo_score = 255 - typo_score;
            reversed_typo_score = scale_component(reversed_typo_score, rel_factor);
            typo_score = 255 - reversed_typo_score;
            uint64_t reversed_proximity = 100 - proximity;
            reversed_proximity = scale_component(reversed_proximity, rel_factor);
            proximity = 100 - reversed_proximity;
            uint64_t reversed_offset_score = 255 - offset_score;
            reversed_offset_score = scale_component(reversed_offset_score, rel_factor);
            offset_score = P.prioritize_token_position ? 255 - reversed_offset_score : 0;
        }
        uint64_t mod_match_score = (uint64_t) (((int64_t) this_words_present << 40) | ((int64_t) unique_words << 32) |
                                               ((int64_t) typo_score << 24) | ((int64_t) proximity << 16) |
                                               ((int64_t) verbatim << 12) | ((int64_t) offset_score << 4) |
                                               ((int64_t) syn << 0));
        if(mod_match_score > (uint64_t) match_score) match_score = (int64_t) mod_match_score;
    }
    return match_score;
}

// Fast path for plain string fields whose raw offsets were validated at mirror-load time to be well formed
// (strictly increasing positions, optionally one trailing 0): every token is exactly one group with array index 0, so
// the segment cursors and the k-way merge of score_field() collapse to direct slices. Same result as score_field().
TS_HD_NOINLINE int64_t score_field_plain(const ScoreParams& P, bool single_exact_query_token, const RawTok* toks, int n_toks) {
    if(n_toks <= 1) return score_field(P, false, single_exact_query_token, toks, n_toks);
    Seg gs[kMaxTokens];
    for(int t = 0; t < n_toks; t++) {
        const uint32_t n = toks[t].n;
        const bool last = n && toks[t].p[n - 1] == 0;
        gs[t].array_index = 0; gs[t].start = 0; gs[t].count = n - (last ? 1u : 0u); gs[t].last_token = last; gs[t].valid = 1;
    }
    MatchOut m = match_window(toks, gs, n_toks, P.prioritize_exact_match != 0);
    const uint32_t synonym_score0 = (P.is_synonym_query && P.demote_synonym_match) ? 0 : 1;
    uint64_t this_match_score = pack_match_score(m.words_present, m.distance, m.max_offset, m.exact_match,
                                                 P.total_cost, (uint32_t) n_toks, synonym_score0);
    uint64_t this_words_present = ((this_match_score >> 40) & 0xFF);
    uint64_t unique_words = ((this_match_score >> 32) & 0xFF);
    uint64_t typo_score = ((this_match_score >> 24) & 0xFF);
    uint64_t proximity = ((this_match_score >> 16) & 0xFF);
    uint64_t verbatim = ((this_match_score >> 12) & 0xF);
    uint64_t offset_score = P.prioritize_token_position ? ((this_match_score >> 4) & 0xFF) : 0;
    uint64_t syn = ((this_match_score >> 0) & 0xF);
    if(P.is_synonym_query && P.num_query_tokens == (uint32_t) n_toks) {
        unique_words = (uint64_t) (int64_t) P.syn_orig_num_tokens;
        this_words_present = (uint64_t) (int64_t) P.syn_orig_num_tokens;
    }
    if(P.is_synonym_query && P.syn_orig_num_tokens > 0 && P.orig_num_tokens > 0) {
        double rel_factor = (double) P.orig_num_tokens / (double) P.syn_orig_num_tokens;
        this_words_present = scale_component(this_words_present, rel_factor);
        unique_words = scale_component(unique_words, rel_factor);
        uint64_t r1 = scale_component(255 - typo_score, rel_factor); typo_score = 255 - r1;
        uint64_t r2 = scale_component(100 - proximity, rel_factor); proximity = 100 - r2;
        uint64_t r3 = scale_component(255 - offset_score, rel_factor);
        offset_score = P.prioritize_token_position ? 255 - r3 : 0;
    }
    return (int64_t) (uint64_t) (((int64_t) this_words_present << 40) | ((int64_t) unique_words << 32) | ((int64_t) typo_score << 24) |
                                 ((int64_t) proximity << 16) | ((int64_t) verbatim << 12) | ((int64_t) offset_score << 4) | ((int64_t) syn));
}

// ---- register-resident scoring of a plain field for short queries -------------------------------------------------
// score_field_plain() keeps its tokens, segments and the Match window in arrays indexed at run time, which nvcc can only
// place in local memory (the kernel's stack traffic in profiles/). For queries of at most kSmallTokens rows the same
// computation fits in registers: every per-token array below is indexed by unrolled compile-time constants only.
//
// Match::Match (include/match_score.h:129-275) without the sort. Each round the reference sorts its window by offset,
// descending, and reads from it: the smallest offset (min), the largest (w_off[0]), how many entries lie within
// WINDOW_SIZE of min, and the sum of the gaps between those entries. The entries within the window are a suffix of the
// sorted order, so the gaps telescope to (largest in-window offset - min); all four are functions of the MULTISET of
// offsets. The entry that leaves is the smallest one. Precondition (kFieldPos16, validated when the field is loaded):
// every position fits uint16, so a token's uint16 positions are strictly increasing and two different tokens never
// share an offset. Equal offsets then only occur between duplicate query tokens, whose states are interchangeable —
// whichever of them leaves first, the multiset of offsets of every later round is the same — so the reference's
// tie order (stable insertion sort) need not be tracked; and "this was the token's last position" can be tested by
// value exactly as the reference does.
#ifndef TSGPU_SMALL_TOKENS
#define TSGPU_SMALL_TOKENS 4            // -DTSGPU_SMALL_TOKENS=3: leaner loops, 4-row combinations fall back to score_field_plain()
#endif
constexpr int kSmallTokens = TSGPU_SMALL_TOKENS;

template <int MAXT>
TS_HD MatchOut match_window_small(const uint32_t* const (&tp)[MAXT], const uint32_t (&tn)[MAXT], uint32_t present, int n_all,
                                  bool check_exact_match) {
    MatchOut out;
    out.words_present = 0; out.distance = 0; out.max_offset = 0; out.exact_match = 0;
    const uint32_t* cur[MAXT];             // address of the token's current position
    uint32_t ol[MAXT];                     // current offset | last offset << 16 (both uint16, as Match stores them)
    uint32_t active = present;
    int32_t last_token_index = -1;         // inputs of the exact-match test: they do not depend on the walk
    uint32_t total_offsets = 0;
#pragma unroll
    for(int t = 0; t < MAXT; t++) {
        cur[t] = tp[t]; ol[t] = 0;
        if((present >> t) & 1u) {
            const uint32_t n = tn[t];
            const bool last = tp[t][n - 1] == 0;
            const uint32_t cnt = n - (last ? 1u : 0u);
            const uint32_t lp = (uint16_t) ((uint16_t) tp[t][cnt - 1] - 1);
            ol[t] = (uint32_t) (uint16_t) ((uint16_t) tp[t][0] - 1) | (lp << 16);
            if(last) last_token_index = (int32_t) lp;
            total_offsets += cnt;
        }
    }
    const uint32_t tokens_size = (uint32_t) n_all;           // n_all <= MAXT <= WINDOW_SIZE
    uint32_t wsize = tokens_size;
    uint32_t best_num_match = 1, best_displacement = kMaxDisplacement;
    int32_t prev_min_offset = -1;
    while(wsize > 1) {
        uint32_t min_offset = 0xFFFFFFFFu, max_off = 0;
        int tk = 0;
#pragma unroll
        for(int t = 0; t < MAXT; t++) {
            if((active >> t) & 1u) {
                const uint32_t o = ol[t] & 0xFFFFu;
                if(o > max_off) max_off = o;
                if(o < min_offset) { min_offset = o; tk = t; }
            }
        }
        if((int32_t) min_offset < prev_min_offset) break;
        prev_min_offset = (int32_t) min_offset;
        uint32_t this_num_match = 0, max_in = min_offset;
#pragma unroll
        for(int t = 0; t < MAXT; t++) {
            const uint32_t o = ol[t] & 0xFFFFu;
            if(((active >> t) & 1u) && o - min_offset <= (uint32_t) kWindowSize) { this_num_match++; if(o > max_in) max_in = o; }
        }
        const uint32_t this_displacement = max_in - min_offset;
        if((this_num_match > best_num_match) || (this_num_match == best_num_match && this_displacement < best_displacement)) {
            best_displacement = this_displacement;
            best_num_match = this_num_match;
            out.max_offset = max_off < 255 ? max_off : 255;
        }
        if(best_num_match == tokens_size && best_displacement == wsize - 1) break;
        // the smallest entry leaves; its token re-enters with its next position unless that was its last one
#pragma unroll
        for(int t = 0; t < MAXT; t++) {
            if(t == tk) {
                if((ol[t] & 0xFFFFu) == (ol[t] >> 16)) { active &= ~(1u << t); wsize--; }
                else { cur[t]++; ol[t] = (ol[t] & 0xFFFF0000u) | (uint32_t) (uint16_t) ((uint16_t) cur[t][0] - 1); }
            }
        }
    }
    if(best_displacement == kMaxDisplacement) best_displacement = 0;
    out.words_present = best_num_match & 0xFF;
    out.distance = best_displacement & 0xFF;
    if(check_exact_match) {
        // (the reference returns early once total_offsets > n_all with distance == n_all - 1; the outcome is the same 0)
        const uint32_t nm1 = (uint32_t) (n_all - 1);
        if(out.distance <= nm1 && last_token_index == (int32_t) n_all - 1) {
            if(total_offsets == (uint32_t) n_all && out.distance == nm1) out.exact_match = 1;
            else if(out.distance < nm1) out.exact_match = 1;
        }
    }
    return out;
}

// score_field_plain() for rows given as (tp[r], tn[r]) with bit r of `present` set when row r matched in this field.
template <int MAXT>
TS_HD int64_t score_field_plain_small(const ScoreParams& P, bool single_exact_query_token, const uint32_t* const (&tp)[MAXT],
                                      const uint32_t (&tn)[MAXT], uint32_t present) {
    int n_toks = 0;
#pragma unroll
    for(int t = 0; t < MAXT; t++) n_toks += (int) ((present >> t) & 1u);
    const uint32_t synonym_score = (P.is_synonym_query && P.demote_synonym_match) ? 0 : 1;
    if(n_toks <= 1) {                      // score_field(), n_toks <= 1, on a plain field
        const uint32_t* p = tp[0]; uint32_t n = tn[0];
#pragma unroll
        for(int t = 1; t < MAXT; t++) if((present >> t) & 1u) { p = tp[t]; n = tn[t]; }
        const uint32_t tail = p[n - 1];
        const uint32_t is_verbatim_match = (P.prioritize_exact_match && single_exact_query_token && p[0] == 1 && n == 2 && tail == 0) ? 1 : 0;
        const bool syn1 = (P.num_query_tokens == 1 && P.is_synonym_query);
        const uint32_t words_present = syn1 ? (uint32_t) P.syn_orig_num_tokens : 1;
        const uint32_t distance = syn1 ? (uint32_t) (P.syn_orig_num_tokens - 1) : 0;
        uint32_t max_offset = 255;
        if(P.prioritize_token_position) max_offset = tail != 0 ? tail : (n >= 2 ? p[n - 2] : 0);     // get_last_offset()
        return (int64_t) pack_match_score(words_present & 0xFF, distance & 0xFF, max_offset & 0xFF, is_verbatim_match, P.total_cost,
                                          words_present, synonym_score);
    }
    const MatchOut m = match_window_small<MAXT>(tp, tn, present, n_toks, P.prioritize_exact_match != 0);
    const uint64_t this_match_score = pack_match_score(m.words_present, m.distance, m.max_offset, m.exact_match, P.total_cost,
                                                       (uint32_t) n_toks, synonym_score);
    uint64_t this_words_present = ((this_match_score >> 40) & 0xFF);
    uint64_t unique_words = ((this_match_score >> 32) & 0xFF);
    uint64_t typo_score = ((this_match_score >> 24) & 0xFF);
    uint64_t proximity = ((this_match_score >> 16) & 0xFF);
    const uint64_t verbatim = ((this_match_score >> 12) & 0xF);
    uint64_t offset_score = P.prioritize_token_position ? ((this_match_score >> 4) & 0xFF) : 0;
    const uint64_t syn = ((this_match_score >> 0) & 0xF);
    if(P.is_synonym_query && P.num_query_tokens == (uint32_t) n_toks) {
        unique_words = (uint64_t) (int64_t) P.syn_orig_num_tokens;
        this_words_present = (uint64_t) (int64_t) P.syn_orig_num_tokens;
    }
    if(P.is_synonym_query && P.syn_orig_num_tokens > 0 && P.orig_num_tokens > 0) {
        const double rel_factor = (double) P.orig_num_tokens / (double) P.syn_orig_num_tokens;
        this_words_present = scale_component(this_words_present, rel_factor);
        unique_words = scale_component(unique_words, rel_factor);
        const uint64_t r1 = scale_component(255 - typo_score, rel_factor); typo_score = 255 - r1;
        const uint64_t r2 = scale_component(100 - proximity, rel_factor); proximity = 100 - r2;
        const uint64_t r3 = scale_component(255 - offset_score, rel_factor);
        offset_score = P.prioritize_token_position ? 255 - r3 : 0;
    }
    return (int64_t) (uint64_t) (((int64_t) this_words_present << 40) | ((int64_t) unique_words << 32) | ((int64_t) typo_score << 24) |
                                 ((int64_t) proximity << 16) | ((int64_t) verbatim << 12) | ((int64_t) offset_score << 4) | ((int64_t) syn));
}

// running reduction of compute_aggregated_score over fields
struct FieldAgg {
    int64_t best_field_match_score, best_field_weight, sum_field_weighted_score;
    uint32_t num_matching_fields;
};
TS_HD void field_agg_init(FieldAgg& a) { a.best_field_match_score = 0; a.best_field_weight = 0; a.sum_field_weighted_score = 0; a.num_matching_fields = 0; }
TS_HD void field_agg_add(FieldAgg& a, uint8_t match_type, int64_t field_match_score, int64_t field_weight) {
    if(match_type == 0 && field_match_score > a.best_field_match_score) { a.best_field_match_score = field_match_score; a.best_field_weight = field_weight; }
    if(match_type == 1 && field_weight > a.best_field_weight) { a.best_field_weight = field_weight; a.best_field_match_score = field_match_score; }
    if(match_type == 2) a.sum_field_weighted_score += field_weight * field_match_score;
    a.num_matching_fields++;
}
TS_HD uint64_t field_agg_finish(const FieldAgg& a, const ScoreParams& P, uint32_t query_len_in) {
    uint64_t query_len = query_len_in;
    if(P.syn_orig_num_tokens != -1) query_len = (uint64_t) (int64_t) P.syn_orig_num_tokens;
    query_len = (a.best_field_match_score == 0) ? 0 : (query_len < 15 ? query_len : 15);
    uint64_t max_field_weight = (uint64_t) a.best_field_weight < 15 ? (uint64_t) a.best_field_weight : 15;
    uint32_t nmf = a.num_matching_fields < 7 ? a.num_matching_fields : 7;
    if(!P.prioritize_num_matching_fields) nmf = 0;
    if(P.match_type == 0) {
        return (uint64_t) (((int64_t) query_len << 59) | ((int64_t) a.best_field_match_score << 11) |
                           ((int64_t) max_field_weight << 3) | ((int64_t) nmf));
    } else if(P.match_type == 1) {
        return (uint64_t) (((int64_t) query_len << 59) | ((int64_t) max_field_weight << 51) |
                           ((int64_t) a.best_field_match_score << 3) | ((int64_t) nmf));
    }
    return (uint64_t) (((int64_t) query_len << 59) | ((int64_t) a.sum_field_weighted_score << 3) | ((int64_t) nmf));
}

TS_HD int64_t float_to_int64(float f) {            // src/index.cpp:266-274
    int32_t i;
#if defined(__CUDA_ARCH__)
    i = __float_as_int(f);
#else
    union { float f; int32_t i; } u; u.f = f; i = u.i;
#endif
    if(i < 0) i ^= 0x7FFFFFFF;
    return (int64_t) i;
}
TS_HD float int64_to_float(int64_t n) {            // src/index.cpp:276-286
    int32_t i = (int32_t) n;
    if(i < 0) i ^= 0x7FFFFFFF;
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    union { float f; int32_t i; } u; u.i = i; return u.f;
#endif
}

struct SortSpec {
    uint8_t type[3];
    int8_t  order[3];
    uint8_t missing_first[3];
    uint8_t pad[3];
    const int64_t* col[3];     // dense sort column (device pointer) for numeric clauses
};

// Index::compute_sort_scores (text_match / seq_id / numeric / vector_distance clauses). Returns match_score_index.
TS_HD int compute_sort_scores(const SortSpec& S, uint32_t seq_id, int64_t max_field_match_score, float vector_distance,
                              int64_t* scores) {
    int msi = -1;
    for(int i = 0; i < 3; i++) {
        scores[i] = 0;
        const uint8_t ty = S.type[i];
        if(ty == 0) continue;
        if(ty == 1) { scores[i] = max_field_match_score; msi = i; }
        else if(ty == 2) scores[i] = (int64_t) seq_id;
        else if(ty == 4) scores[i] = float_to_int64(vector_distance);
        else {
            int64_t v = S.col[i][seq_id];
            if(v == INT64_MIN && S.missing_first[i]) v = (S.order[i] == -1) ? (INT64_MIN + 1) : INT64_MAX;
            scores[i] = v;
        }
        if(S.order[i] == -1) scores[i] = (int64_t) (0 - (uint64_t) scores[i]);
    }
    return msi;
}

// KV order (include/topster.h:146-149): (scores[0], scores[1], scores[2], key) descending.
TS_HD bool kv_greater(int64_t a0, int64_t a1, int64_t a2, uint32_t ak, int64_t b0, int64_t b1, int64_t b2, uint32_t bk) {
    if(a0 != b0) return a0 > b0;
    if(a1 != b1) return a1 > b1;
    if(a2 != b2) return a2 > b2;
    return ak > bk;
}

// posting_list_t::has_phrase_match(token_positions) + found_token_sequence (src/posting_list.cpp:1719-1789), iterative:
// some position p of token 0 such that token i holds p+i for every i. Position lists that wrap around (a later
// position smaller than an earlier one) stop the scan exactly where the reference's loops stop.
TS_HD bool has_phrase_match(const RawTok* toks, const Seg* segs, int n) {
    int32_t prev_pos = -1;
    for(uint32_t a = 0; a < segs[0].count; a++) {
        const uint16_t pos = seg_pos(toks[0], segs[0], a);
        if((int32_t) pos < prev_pos) return false;
        bool all = true;
        for(int i = 1; i < n && all; i++) {
            const uint16_t target = (uint16_t) (pos + i);
            bool found = false;
            int32_t pp = -1;
            for(uint32_t k = 0; k < segs[i].count; k++) {
                const uint16_t tp = seg_pos(toks[i], segs[i], k);
                if((int32_t) tp < pp) { found = false; break; }
                if(tp == target) { found = true; break; }
                pp = tp;
            }
            all = found;
        }
        if(all) return true;
        prev_pos = pos;
    }
    return false;
}

// posting_list_t::get_phrase_matches body for one id (src/posting_list.cpp:1806-1820): any array element that holds
// all k tokens as a consecutive sequence.
TS_HD_NOINLINE bool phrase_match_doc(const RawTok* toks, int k) {
    if(k == 1) return true;
    SegCursor cur[kMaxTokens];
    Seg nxt[kMaxTokens];
    for(int t = 0; t < k; t++) { cur[t].pos = 0; cur[t].is_last = 0; nxt[t] = seg_next(toks[t], cur[t]); }
    for(;;) {
        uint32_t key = 0xFFFFFFFFu; bool any = false;
        for(int t = 0; t < k; t++) if(nxt[t].valid) { if(!any || nxt[t].array_index < key) key = nxt[t].array_index; any = true; }
        if(!any) break;
        RawTok gt[kMaxTokens]; Seg gs[kMaxTokens];
        int n = 0;
        for(int t = 0; t < k; t++) {
            while(nxt[t].valid && nxt[t].array_index == key) {
                if(n < kMaxTokens) { gt[n] = toks[t]; gs[n] = nxt[t]; n++; }
                nxt[t] = seg_next(toks[t], cur[t]);
            }
        }
        if(n == k && has_phrase_match(gt, gs, n)) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------
// posting_list_t::get_exact_matches / get_prefix_matches body for one id (src/posting_list.cpp:1129-1279, 1281-1452):
// token j of the filter value must sit at position j+1 of the field value (of some array element), and for an exact
// match that element must also end with the last token. One walker yields the per-element facts both need.
struct ElemFact { uint32_t array_index; bool matched; bool last; };
struct ElemWalk { uint32_t i; uint32_t n_matching; };

// next array element of token `t` (0-based query index jm1+1 == wanted position). `k` = number of query tokens;
// `want_last` selects the exact-match variant, which also consumes the "last token of the element" flag.
TS_HD bool elem_next(const RawTok& t, ElemWalk& w, uint32_t want_pos, uint32_t k, bool want_last, ElemFact& out) {
    int64_t prev_pos = -1;
    bool found = false;
    while(w.i < t.n) {
        const uint32_t pos = t.p[w.i];
        w.i++;
        if((int64_t) pos == prev_pos) {                 // end of an array element: next word is its index
            if(w.i >= t.n) return false;                // malformed tail
            out.array_index = t.p[w.i];
            out.last = false;
            if(want_last && w.i + 1 < t.n && t.p[w.i + 1] == 0 && pos == k) { out.last = true; w.i++; }
            out.matched = found;
            w.i++;
            return true;
        }
        if(pos == want_pos) { found = true; w.n_matching++; }
        prev_pos = pos;
    }
    return false;
}

// plain string field, k > 1
TS_HD bool positional_match_plain(const RawTok* toks, int k, bool exact) {
    for(int j = k - 1; j >= 0; j--) {
        const RawTok& t = toks[j];
        if(t.n == 0) return false;
        if(exact && j == k - 1) {
            // the last query token has to be the last token of the value: [.., k, 0]
            if(t.p[t.n - 1] != 0 || (t.n >= 2 && t.p[t.n - 2] != (uint32_t) k)) return false;
        }
        for(uint32_t i = 0; i < t.n; i++) {
            const uint32_t off = t.p[i];
            if(off == (uint32_t) (j + 1)) break;
            if(off > (uint32_t) (j + 1)) return false;
        }
    }
    return true;
}

TS_HD_NOINLINE bool positional_match_doc(const RawTok* toks, int k, bool field_is_array, bool exact) {
    if(k == 1) {
        if(exact) return is_single_token_verbatim_match(toks[0], field_is_array);
        return toks[0].n != 0 && toks[0].p[0] == 1;      // is_single_token_prefix_match (src/posting_list.cpp:1114-1127)
    }
    if(!field_is_array) return positional_match_plain(toks, k, exact);
    // array field: some element must hold every token at its position (and, for exact, carry the last-token flag,
    // which any token's walk may set). The reference's early exits are all implied failures of that test except
    // "last query token is never a last token", kept explicitly.
    for(int j = 0; j < k; j++) {
        ElemWalk w{0, 0}; ElemFact f;
        bool any_last = false;
        while(elem_next(toks[j], w, (uint32_t) j + 1, (uint32_t) k, exact, f)) any_last |= f.last;
        if(w.n_matching == 0) return false;
        if(exact && j == k - 1 && !any_last) return false;
    }
    ElemWalk w0{0, 0}; ElemFact f0;
    while(elem_next(toks[0], w0, 1, (uint32_t) k, exact, f0)) {
        if(!f0.matched) continue;
        bool all = true, last = f0.last;
        for(int j = 0; j < k; j++) {
            ElemWalk w{0, 0}; ElemFact f;
            bool m = false;
            while(elem_next(toks[j], w, (uint32_t) j + 1, (uint32_t) k, exact, f)) {
                if(f.array_index != f0.array_index) continue;
                m |= f.matched; last |= f.last;
            }
            if(!m) all = false;                           // keep walking: later tokens may still set `last`
        }
        if(all && (!exact || last)) return true;
    }
    return false;
}

}  // namespace tsdev
