// The walk half of the ART fuzzy / prefix candidate search as __host__ __device__ code: which subtrees
// of a field's token index hold keys within [min_cost, max_cost] edits of a query token.
//
// Replaces (reference file:line):
//   art_fuzzy_recurse / art_fuzzy_children      src/art.cpp:1596-1738, 1435-1485   -> art_walk()
//   fuzzy_search_state                          src/art.cpp:1487-1594              -> art_state()
//   levenshtein_dist                            src/art.cpp:1412-1433              -> art_next_row()
// The other half (art_topk_iter, validate_and_add_leaf, the final sort: a few dozen leaves per search) stays on the host,
// typesense_b200/host/art_mirror.hpp::finish — whose own walk() is the readable, recursive statement of this file.
//
// Device use: art_walk_kernel (art_kernels.cu) runs one search per thread — a batch is (queries x tokens x typo costs), tens of
// thousands of independent walks; each is a pointer chase through the flat node arrays, so threads rather than warps are
// the unit. Host use: tests/hostsim compiles this header with g++ and checks the hit lists against art_mirror_t::walk_hits
// (itself pinned on the reference's compiled art.cpp). Not measured on a GPU yet (written after round 1's GPU budget).
//
// The recursion of the reference becomes an explicit stack: a frame is an inner node whose own bytes have been consumed
// (the two DP rows children start from, the byte that led here, the depth) plus the next child to visit, largest byte
// first. A walk cannot run deep: past query length + 4 the cost exceeds every tolerance of art_state().
#pragma once
#include <stdint.h>

#ifndef TS_HD                      // as in score_device.cuh (not included: this header stands alone in art_kernels.cu)
#if defined(__CUDACC__)
#define TS_HD __host__ __device__ __forceinline__
#else
#define TS_HD inline
#endif
#endif

namespace tsdev {

constexpr int kArtMaxQuery = 31;       // query bytes incl. the terminator of a whole-word search (one DP column each, + column 0)
constexpr int kArtMaxStack = 48;
constexpr int kArtPartialBytes = 8;    // MAX_PREFIX_LEN include/art.h:23

struct ArtNodeDev {
    uint32_t first_child;
    uint16_t n_children;
    uint8_t  partial_len;
    uint8_t  partial[kArtPartialBytes];
    uint8_t  pad;
};

struct ArtDev {
    const ArtNodeDev* nodes;
    const uint8_t*  child_byte;
    const int32_t*  child_ref;         // >= 0 inner node, < 0 leaf ~ref
    const uint64_t* leaf_key_off;      // keys without terminator, concatenated
    const uint8_t*  leaf_keys;
    int32_t root;
    uint32_t empty;
};

struct ArtQuery {
    uint8_t q[kArtMaxQuery + 1];
    int qlen, min_cost, max_cost;
    bool prefix;
};

typedef uint8_t ArtRow[kArtMaxQuery + 1];

TS_HD void art_next_row(int depth, uint8_t p, uint8_t c, const ArtQuery& Q, const uint8_t* prev2, const uint8_t* prev, uint8_t* out) {
    out[0] = (uint8_t) (prev[0] + 1);
    for(int col = 1; col <= Q.qlen; col++) {
        int v = prev[col - 1] + (c == Q.q[col - 1] ? 0 : 1);
        const int ins = out[col - 1] + 1, del = prev[col] + 1;
        if(ins < v) v = ins;
        if(del < v) v = del;
        if(depth > 1 && col > 1 && c == Q.q[col - 2] && p == Q.q[col - 1]) { const int tr = prev2[col - 2] + 1; if(tr < v) v = tr; }
        out[col] = (uint8_t) v;
    }
}

// +1 accept every key below, 0 read on, -1 give up
TS_HD int art_state(const ArtQuery& Q, int key_index, uint8_t p, uint8_t c, const uint8_t* row) {
    const bool key_ends = c == 0;
    const int key_len = key_ends ? key_index : key_index + 1;
    const int qlen = Q.qlen;
    if(key_ends) {
        if(row[qlen] >= Q.min_cost && row[qlen] <= Q.max_cost) return 1;
        if(key_len > 5 && qlen > key_len && qlen - key_len <= Q.max_cost && row[key_len] >= Q.min_cost && row[key_len] <= Q.max_cost - 1) return 1;
        return -1;
    }
    const int cost = row[key_len < qlen ? key_len : qlen];
    if(Q.prefix && key_len >= qlen && cost >= Q.min_cost && cost <= Q.max_cost) return 1;
    if(cost <= Q.max_cost) return 0;
    // query bytes behind the query's end read as 0 (the reference reads them unguarded, src/art.cpp:1557, 1587): never a key byte
    if(cost == 2 || cost == 3) {
        if((key_index + 1 < qlen && Q.q[key_index + 1] == c) || (key_index > 0 && key_index - 1 < qlen && Q.q[key_index - 1] == c)) return 0;
    }
    if(cost == 3 || cost == 4) {
        if(key_index + 2 < qlen && Q.q[key_index + 1] == p && Q.q[key_index + 2] == c) return 0;
        if(key_index > 1 && key_index - 2 < qlen && Q.q[key_index - 2] == c) return 0;
    }
    return -1;
}

// One visit: `ref` is about to be entered through key byte c (previous byte p) at `depth` with the two DP rows of the
// path so far (depth -1: the root, no byte leads to it).
struct ArtItem {
    ArtRow prev2, prev;
    int32_t ref;
    int16_t depth;
    uint8_t p, c;
};

// Consumes the byte leading to the item's node and the node's own bytes (its compressed path, or a leaf's remaining key).
// Returns true when the node's children have to be visited: the item then holds the state they start from (rows, depth, and
// c = the last byte read, their `p`). *hit: every key below item.ref is a candidate (the walk does not descend further).
TS_HD bool art_enter(const ArtDev& A, const ArtQuery& Q, ArtItem& it, bool* hit) {
    ArtRow rows[3];
    for(int i = 0; i <= Q.qlen; i++) { rows[0][i] = it.prev2[i]; rows[1][i] = it.prev[i]; }
    int i2 = 0, i1 = 1, i0 = 2;
    int depth = it.depth;
    uint8_t p = it.p, c = it.c;
    const int32_t ref = it.ref;
    bool decided = false;
    *hit = false;
#define ART_STEP(BYTE, ADVANCE)                                                                          \
    {                                                                                                    \
        const uint8_t byte_ = (BYTE);                                                                    \
        if(ADVANCE) { art_next_row(depth, p, byte_, Q, rows[i2], rows[i1], rows[i0]); const int t_ = i2; i2 = i1; i1 = i0; i0 = t_; } \
        const int a_ = art_state(Q, depth, p, byte_, rows[i1]);                                          \
        if(a_ == 1) *hit = true;                                                                         \
        if(a_ != 0) decided = true; else { p = byte_; depth++; }                                         \
    }
    if(depth == -1) depth = 0;
    else ART_STEP(c, !(Q.prefix && c == 0))
    if(decided) return false;
    if(ref < 0) {
        const uint64_t o = A.leaf_key_off[~ref];
        const int klen = (int) (A.leaf_key_off[~ref + 1] - o) + 1;            // with the terminator
        const int iter_len = klen < Q.qlen + Q.max_cost ? klen : Q.qlen + Q.max_cost;
        if(depth >= iter_len) { *hit = art_state(Q, depth, 0, 0, rows[i1]) == 1; return false; }
        while(depth < iter_len && !decided) {
            c = depth < klen - 1 ? A.leaf_keys[o + depth] : 0;
            ART_STEP(c, !(Q.prefix && c == 0))
        }
        return false;
    }
    const ArtNodeDev& n = A.nodes[ref];
    int seen = n.partial_len < kArtPartialBytes ? n.partial_len : kArtPartialBytes;
    for(int i = 0; i < seen && !decided; i++) { c = n.partial[i]; ART_STEP(c, true) }
    // only the first kArtPartialBytes of a compressed path are stored: the rest is assumed to agree with the query
    while(!decided && seen < (int) n.partial_len && depth < Q.qlen) { c = Q.q[depth]; ART_STEP(c, true) seen++; }
#undef ART_STEP
    if(decided) return false;
    for(int i = 0; i <= Q.qlen; i++) { it.prev2[i] = rows[i2][i]; it.prev[i] = rows[i1][i]; }
    it.depth = (int16_t) depth; it.c = c;
    return true;
}

// art_enter() for query tokens of at most QL bytes with every DP row in REGISTERS: arrays of compile-time size, every loop
// unrolled to QL with the column test as a predicate, the three rows rotated by copies (moves the compiler renames away),
// dynamic row reads as select chains, a leaf's remaining key bytes loaded together before the first of them is used
// (independent loads instead of one dependent byte load per step). Same arithmetic, same rules, same results as art_enter():
// the frontier kernel's items spend their time here, and the general form keeps rows[3][32] in local memory behind run-time
// indices (profiles/r02s: 31 % of the end-to-end bench's kernel time was art_frontier_kernel).
template <int QL>
TS_HD uint8_t art_pick(const uint8_t (&row)[QL + 1], int idx) {
    uint8_t v = row[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 1; i <= QL; i++) v = (i == idx) ? row[i] : v;
    return v;
}
template <int QL>
TS_HD int art_state_t(const ArtQuery& Q, const uint8_t (&qq)[QL + 3], int key_index, uint8_t p, uint8_t c, const uint8_t (&row)[QL + 1]) {
    // qq[i + 1] = Q.q[i] for 0 <= i < qlen, 0 elsewhere (so that index -1 .. QL + 1 can be read)
    const bool key_ends = c == 0;
    const int key_len = key_ends ? key_index : key_index + 1;
    const int qlen = Q.qlen;
    if(key_ends) {
        const uint8_t rq = art_pick<QL>(row, qlen);
        if(rq >= Q.min_cost && rq <= Q.max_cost) return 1;
        if(key_len > 5 && qlen > key_len && qlen - key_len <= Q.max_cost) {
            const uint8_t rk = art_pick<QL>(row, key_len);
            if(rk >= Q.min_cost && rk <= Q.max_cost - 1) return 1;
        }
        return -1;
    }
    const int cost = art_pick<QL>(row, key_len < qlen ? key_len : qlen);
    if(Q.prefix && key_len >= qlen && cost >= Q.min_cost && cost <= Q.max_cost) return 1;
    if(cost <= Q.max_cost) return 0;
    auto qat = [&](int i) -> uint8_t {          // Q.q[i] for i in [-1, QL + 1], 0 outside the query
        uint8_t v = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int j = 0; j < QL + 3; j++) v = (j == i + 1) ? qq[j] : v;
        return v;
    };
    if(cost == 2 || cost == 3) {
        if((key_index + 1 < qlen && qat(key_index + 1) == c) || (key_index > 0 && key_index - 1 < qlen && qat(key_index - 1) == c)) return 0;
    }
    if(cost == 3 || cost == 4) {
        if(key_index + 2 < qlen && qat(key_index + 1) == p && qat(key_index + 2) == c) return 0;
        if(key_index > 1 && key_index - 2 < qlen && qat(key_index - 2) == c) return 0;
    }
    return -1;
}

template <int QL>
TS_HD bool art_enter_t(const ArtDev& A, const ArtQuery& Q, ArtItem& it, bool* hit) {
    const int qlen = Q.qlen;                    // <= QL (caller's dispatch)
    uint8_t r2[QL + 1], r1[QL + 1], r0[QL + 1], qq[QL + 3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i <= QL; i++) { r2[i] = it.prev2[i]; r1[i] = it.prev[i]; r0[i] = 0; }
    qq[0] = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i < QL + 2; i++) qq[i + 1] = (i < qlen) ? Q.q[i] : (uint8_t) 0;
    int depth = it.depth;
    uint8_t p = it.p, c = it.c;
    const int32_t ref = it.ref;
    bool decided = false;
    *hit = false;
    auto step = [&](uint8_t byte_, bool advance) {
        if(advance) {
            // art_next_row(depth, p, byte_, Q, r2, r1, r0)
            r0[0] = (uint8_t) (r1[0] + 1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for(int col = 1; col <= QL; col++) {
                if(col <= qlen) {
                    int v = r1[col - 1] + (byte_ == qq[col] ? 0 : 1);
                    const int ins = r0[col - 1] + 1, del = r1[col] + 1;
                    if(ins < v) v = ins;
                    if(del < v) v = del;
                    if(col > 1) { if(depth > 1 && byte_ == qq[col - 1] && p == qq[col]) { const int tr = r2[col - 2] + 1; if(tr < v) v = tr; } }
                    r0[col] = (uint8_t) v;
                }
            }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for(int i = 0; i <= QL; i++) { r2[i] = r1[i]; r1[i] = r0[i]; }
        }
        const int a_ = art_state_t<QL>(Q, qq, depth, p, byte_, r1);
        if(a_ == 1) *hit = true;
        if(a_ != 0) decided = true; else { p = byte_; depth++; }
    };
    if(depth == -1) depth = 0;
    else step(c, !(Q.prefix && c == 0));
    if(decided) return false;
    if(ref < 0) {
        const uint64_t o = A.leaf_key_off[~ref];
        const int klen = (int) (A.leaf_key_off[~ref + 1] - o) + 1;            // with the terminator
        const int iter_len = klen < qlen + Q.max_cost ? klen : qlen + Q.max_cost;
        if(depth >= iter_len) { *hit = art_state_t<QL>(Q, qq, depth, 0, 0, r1) == 1; return false; }
        // the bytes this loop can read: key positions depth .. iter_len - 1, at most QL + 3 of them from `depth`
        constexpr int KB = QL + 4;
        uint8_t kb[KB];
        const int d0 = depth;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int j = 0; j < KB; j++) { const int pos = d0 + j; kb[j] = (pos < iter_len && pos < klen - 1) ? A.leaf_keys[o + pos] : (uint8_t) 0; }
        if(iter_len - d0 > KB) {                 // cannot happen for max_cost <= 3 (iter_len <= qlen + max_cost); keep the general loop for safety
            while(depth < iter_len && !decided) { c = depth < klen - 1 ? A.leaf_keys[o + depth] : 0; step(c, !(Q.prefix && c == 0)); }
            return false;
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int j = 0; j < KB; j++) {
            if(d0 + j < iter_len && !decided && depth == d0 + j) { c = kb[j]; step(c, !(Q.prefix && c == 0)); }
        }
        return false;
    }
    const ArtNodeDev& n = A.nodes[ref];
    const int plen = n.partial_len;
    int seen = plen < kArtPartialBytes ? plen : kArtPartialBytes;
    uint8_t pb[kArtPartialBytes];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i < kArtPartialBytes; i++) pb[i] = n.partial[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i < kArtPartialBytes; i++) { if(i < seen && !decided) { c = pb[i]; step(c, true); } }
    // only the first kArtPartialBytes of a compressed path are stored: the rest is assumed to agree with the query
    while(!decided && seen < plen && depth < qlen) { c = Q.q[depth]; step(c, true); seen++; }
    if(decided) return false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i <= QL; i++) { it.prev2[i] = r2[i]; it.prev[i] = r1[i]; }
    it.depth = (int16_t) depth; it.c = c;
    return true;
}

// dispatch on the query's length
TS_HD bool art_enter_fast(const ArtDev& A, const ArtQuery& Q, ArtItem& it, bool* hit) {
    if(Q.qlen <= 10) return art_enter_t<10>(A, Q, it, hit);
    if(Q.qlen <= 16) return art_enter_t<16>(A, Q, it, hit);
    return art_enter(A, Q, it, hit);
}

TS_HD void art_root_item(const ArtDev& A, const ArtQuery& Q, ArtItem& it) {
    it.ref = A.root; it.p = 0; it.c = 0; it.depth = -1;
    if(A.root < 0) {              // a one-key index: the root is that leaf and its first byte is read like any other
        const uint64_t o = A.leaf_key_off[~A.root];
        it.c = A.leaf_key_off[~A.root + 1] > o ? A.leaf_keys[o] : 0;
        it.depth = 0;
    }
    for(int i = 0; i <= Q.qlen; i++) { it.prev2[i] = (uint8_t) i; it.prev[i] = (uint8_t) i; }
}
// the item of child k of an entered inner node
TS_HD void art_child_item(const ArtDev& A, const ArtQuery& Q, const ArtItem& parent, uint32_t k, ArtItem& ch) {
    const ArtNodeDev& n = A.nodes[parent.ref];
    ch.ref = A.child_ref[n.first_child + k]; ch.c = A.child_byte[n.first_child + k]; ch.p = parent.c; ch.depth = parent.depth;
    for(int i = 0; i <= Q.qlen; i++) { ch.prev2[i] = parent.prev2[i]; ch.prev[i] = parent.prev[i]; }
}

// Depth-first driver (one thread does a whole search): a frame is an entered inner node and the next child to visit.
struct ArtFrame {
    ArtItem at;
    uint16_t next_child;       // children [0, next_child) are still to visit, from the top
};

// Writes the matching refs (walk order) to hits[0..cap) and returns how many there are (may exceed cap: caller's overflow).
// *stack_overflow is set when the walk needed more than kArtMaxStack frames (the caller falls back to the host walk).
TS_HD uint32_t art_walk(const ArtDev& A, const ArtQuery& Q, int32_t* hits, uint32_t cap, ArtFrame* stack, bool* stack_overflow) {
    uint32_t n_hits = 0;
    *stack_overflow = false;
    if(A.empty) return 0;
    int sp = 0;
    ArtItem it;
    art_root_item(A, Q, it);
    bool have = true;
    while(true) {
        if(!have) {
            if(sp == 0) break;
            ArtFrame& f = stack[sp - 1];
            if(f.next_child == 0) { sp--; continue; }
            art_child_item(A, Q, f.at, --f.next_child, it);
        }
        have = false;
        bool hit = false;
        const bool descend = art_enter(A, Q, it, &hit);
        if(hit) { if(n_hits < cap) hits[n_hits] = it.ref; n_hits++; }
        if(!descend) continue;
        if(sp == kArtMaxStack) { *stack_overflow = true; return n_hits; }
        stack[sp].at = it;
        stack[sp].next_child = A.nodes[it.ref].n_children;
        sp++;
    }
    return n_hits;
}

// Breadth-first driver, one level: every item of `in` is entered; hits go to (hit_search, hit_ref), the children of the nodes
// that have to be descended into become the next level's items. The device form gives one thread (later: one warp) per item and
// appends with atomic counters; this serial form is what tests/ run on the host. Hits come out in no particular order: the
// reference's order is the pre-order of the tree with children descending, a static rank per node (art_mirror_t::preorder_ranks).
struct ArtWorkItem { ArtItem at; uint32_t search; };

}  // namespace tsdev
