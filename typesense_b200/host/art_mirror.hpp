// art_mirror_t — a flat, read-only mirror of one string field's token index (the reference keeps it in an adaptive radix
// tree, `art_tree`, one per field: include/art.h) and the fuzzy / prefix candidate search over it. SURVEY §8 row f-1.
//
//   reference                                                     here
//   art_fuzzy_search_i          src/art.cpp:1825-1894             art_mirror_t::fuzzy_search
//   art_fuzzy_recurse           src/art.cpp:1596-1738             art_mirror_t::walk          (which subtrees match)
//   fuzzy_search_state          src/art.cpp:1487-1594             art_mirror_t::search_state  (accept / continue / prune)
//   levenshtein_dist            src/art.cpp:1412-1433             art_mirror_t::next_row      (one OSA row per key byte)
//   art_topk_iter               src/art.cpp:1143-1240             art_mirror_t::collect       (best leaves of a subtree)
//   validate_and_add_leaf       src/art.cpp:1003-1044             art_mirror_t::admit
//   art_search                  src/art.cpp:321-356               art_mirror_t::find
//
// Layout: inner nodes, leaves and child links are three flat arrays (children of a node contiguous and ascending by key
// byte — every node type of the reference iterates its children in byte order, so the 4/16/48/256 distinction carries no
// information) — the form a device kernel would read. A mirror is either LOADED from an export of the live tree
// (load_export: structure, compressed-path bytes and per-node max_score exactly as the reference's inserts left them; the
// search then returns the same leaves in the same order as the reference, ties included — tests/test_art_mirror.py pins
// that against the reference's own art.cpp compiled in oracle/_ref), or BUILT from a vocabulary (build: the canonical
// radix tree of the key set with max_score = the subtree's maximum) where no live tree exists (this repository's harness).
//
// The order of equal-score candidates is part of the contract because max_candidates truncates the list: it is decided by
// std::priority_queue and std::sort driven with the reference's (non-strict) comparators in the reference's push order,
// which this file reproduces call for call.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <queue>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace tsgpu {

class art_mirror_t {
public:
    static constexpr int kPartialBytes = 8;          // MAX_PREFIX_LEN include/art.h:23
    enum token_ordering { FREQUENCY = 1, MAX_SCORE = 2 };

    struct leaf_t { std::string key; int64_t max_score; uint32_t num_ids; uint32_t list; };     // key without the NUL
    struct node_t { uint8_t partial_len; uint8_t partial[kPartialBytes]; int64_t max_score; uint32_t first_child, n_children; };

    std::vector<node_t> nodes;
    std::vector<leaf_t> leaves;
    std::vector<uint8_t> child_byte;
    std::vector<int32_t> child_ref;                  // >= 0: inner node, < 0: leaf ~ref
    int32_t root = 0;
    bool empty = true;

    // ---- construction ------------------------------------------------------------------------------------------------
    // from ref_art_export's byte stream (oracle/ref_wrap.cpp; a server-side exporter would walk art_tree the same way)
    bool load_export(const unsigned char* buf, size_t n) {
        clear();
        size_t at = 0;
        if(n == 0) return false;
        if(buf[0] == 'E') return n == 1;
        bool ok = true;
        root = parse(buf, n, at, ok);
        empty = false;
        index_leaves();
        return ok && at == n;
    }
    struct vocab_entry { std::string token; int64_t max_score; uint32_t num_ids; uint32_t list; };
    void build(std::vector<vocab_entry> v) {
        clear();
        if(v.empty()) return;
        std::sort(v.begin(), v.end(), [](const vocab_entry& a, const vocab_entry& b) { return a.token < b.token; });
        std::vector<std::string> keys;
        for(auto& e: v) { keys.push_back(e.token); keys.back().push_back('\0'); }
        root = build_range(v, keys, 0, v.size(), 0);
        empty = false;
        index_leaves();
    }
    // list id of every leaf, for mirrors loaded from an export
    void bind_lists(const std::function<uint32_t(const std::string&)>& list_of) { for(auto& l: leaves) l.list = list_of(l.key); }

    static uint64_t& last_walk_visited() { static thread_local uint64_t v = 0; return v; }      // size of the last walk_hits() (tuning)
    // art_search: the leaf whose key equals `token`, or -1
    int32_t find(const std::string& token) const {
        auto it = by_key.find(token);
        return it == by_key.end() ? -1 : (int32_t) it->second;
    }

    // the arrays of tsgpu_art (include/tsgpu.h) that are not already members: per-node columns and the key arena
    struct flat_t {
        std::vector<uint32_t> node_first_child;
        std::vector<uint16_t> node_n_children;
        std::vector<uint8_t> node_partial_len, node_partial;          // node_partial: kPartialBytes per node
        std::vector<uint64_t> leaf_key_off;
        std::vector<uint8_t> leaf_keys;
    };
    flat_t flatten() const {
        flat_t f;
        for(auto& n: nodes) {
            f.node_first_child.push_back(n.first_child);
            f.node_n_children.push_back((uint16_t) n.n_children);
            f.node_partial_len.push_back(n.partial_len);
            f.node_partial.insert(f.node_partial.end(), n.partial, n.partial + kPartialBytes);
        }
        f.leaf_key_off.push_back(0);
        for(auto& l: leaves) { f.leaf_keys.insert(f.leaf_keys.end(), l.key.begin(), l.key.end()); f.leaf_key_off.push_back(f.leaf_keys.size()); }
        return f;
    }

    // Position of every node / leaf in the order the reference's recursion visits the tree (pre-order, children from the largest
    // byte down). The hits of any search, in any order, sorted by this rank are the hits in the reference's order — what a
    // breadth-first (frontier-parallel) device walk needs to hand back the same list as the depth-first one.
    struct ranks_t { std::vector<uint32_t> node, leaf; };
    ranks_t preorder_ranks() const {
        ranks_t r;
        r.node.assign(nodes.size(), 0);
        r.leaf.assign(leaves.size(), 0);
        if(empty) return r;
        uint32_t next = 0;
        std::vector<int32_t> stack{root};
        while(!stack.empty()) {
            const int32_t ref = stack.back();
            stack.pop_back();
            if(ref < 0) { r.leaf[~ref] = next++; continue; }
            r.node[ref] = next++;
            const node_t& n = nodes[ref];
            for(uint32_t k = 0; k < n.n_children; k++) stack.push_back(child_ref[n.first_child + k]);      // popped largest byte first
        }
        return r;
    }

    // ---- search ------------------------------------------------------------------------------------------------------
    // has_filter_doc(list): the posting list holds a document of the active filter (only consulted when filter_active);
    // share_doc(a, b): lists a and b hold a common document (that also passes the filter when one is active).
    struct doc_tests {
        bool filter_active = false;
        std::function<bool(uint32_t)> has_filter_doc;
        std::function<bool(uint32_t, uint32_t)> share_doc;
    };
    // Leaves (indices into `leaves`) at `cost` edits of `term` — or, for a prefix search, whose key has such a
    // prefix — best first, at most max_words; tokens already in exclude_leaves are skipped and the returned ones added.
    std::vector<uint32_t> fuzzy_search(const std::string& term, int cost, size_t max_words, token_ordering order, bool prefix,
                                       const std::string& prev_token, const doc_tests& docs, std::set<std::string>& exclude_leaves) const {
        return fuzzy_search(term, cost, cost, max_words, order, prefix, prev_token, docs, exclude_leaves);
    }
    // the general form (the search path always asks for one exact cost; test/art_test.cpp uses ranges)
    std::vector<uint32_t> fuzzy_search(const std::string& term, int min_cost, int max_cost, size_t max_words, token_ordering order, bool prefix,
                                       const std::string& prev_token, const doc_tests& docs, std::set<std::string>& exclude_leaves) const {
        return finish(term, min_cost, max_words, order, prev_token, docs, exclude_leaves, walk_hits(term, min_cost, max_cost, prefix));
    }

    // The two halves of fuzzy_search. walk_hits: the subtrees / leaves whose keys match, in the order the walk meets them
    // (the part that touches the whole tree — tsgpu_art_walk_batch computes the same list on the device for a batch of
    // searches). finish: best leaves of those subtrees, sorted, exact token first, truncated.
    std::vector<int32_t> walk_hits(const std::string& term, int min_cost, int max_cost, bool prefix) const {
        search_t s;
        if(empty) return s.hits;
        s.q.assign(term.begin(), term.end());
        if(!prefix) s.q.push_back('\0');                           // the key's terminator takes part in a whole-word match
        s.min_cost = min_cost; s.max_cost = max_cost;
        s.prefix = prefix;
        if(s.q.size() + 1 > (size_t) kMaxCols) return s.hits;      // cannot happen for indexed tokens (cut at 100 bytes)
        int row0[kMaxCols];
        for(size_t i = 0; i <= s.q.size(); i++) row0[i] = (int) i;
        if(root < 0) walk(s, 0, (uint8_t) leaves[~root].key.c_str()[0], root, 0, row0, row0);
        else walk(s, 0, 0, root, -1, row0, row0);
        last_walk_visited() = s.visited;
        return s.hits;
    }
    std::vector<uint32_t> finish(const std::string& term, int min_cost, size_t max_words, token_ordering order, const std::string& prev_token,
                                 const doc_tests& docs, std::set<std::string>& exclude_leaves, const std::vector<int32_t>& hits) const {
        std::vector<uint32_t> results;
        if(empty) return results;
        const int32_t exact_leaf = find(term);
        const int32_t prev_leaf = find(prev_token);
        for(int32_t n: hits) collect(n, order, max_words, exact_leaf, prev_token, prev_leaf, docs, exclude_leaves, results);
        if(order == FREQUENCY) std::sort(results.begin(), results.end(), [&](uint32_t a, uint32_t b) { return leaves[a].num_ids > leaves[b].num_ids; });
        else std::sort(results.begin(), results.end(), [&](uint32_t a, uint32_t b) { return leaves[a].max_score > leaves[b].max_score; });
        if(exact_leaf >= 0 && min_cost == 0 && !exclude_leaves.count(leaves[exact_leaf].key)) {
            results.insert(results.begin(), (uint32_t) exact_leaf);
            exclude_leaves.insert(leaves[exact_leaf].key);
        }
        if(results.size() > max_words) results.resize(max_words);
        return results;
    }

private:
    std::unordered_map<std::string, uint32_t> by_key;

    void clear() { nodes.clear(); leaves.clear(); child_byte.clear(); child_ref.clear(); by_key.clear(); root = 0; empty = true; }
    void index_leaves() { for(uint32_t i = 0; i < leaves.size(); i++) by_key[leaves[i].key] = i; }

    int32_t parse(const unsigned char* b, size_t n, size_t& at, bool& ok) {
        if(at >= n) { ok = false; return 0; }
        const unsigned char tag = b[at++];
        if(tag == 'L') {
            if(at + 4 > n) { ok = false; return 0; }
            uint32_t kl; std::memcpy(&kl, b + at, 4); at += 4;
            if(kl == 0 || at + kl + 12 > n) { ok = false; return 0; }
            leaf_t l;
            l.key.assign((const char*) b + at, kl - 1); at += kl;
            std::memcpy(&l.max_score, b + at, 8); at += 8;
            std::memcpy(&l.num_ids, b + at, 4); at += 4;
            l.list = 0xFFFFFFFFu;
            leaves.push_back(l);
            return ~(int32_t) (leaves.size() - 1);
        }
        if(tag != 'N' || at + 1 + kPartialBytes + 8 + 2 > n) { ok = false; return 0; }
        node_t nd;
        nd.partial_len = b[at++];
        std::memcpy(nd.partial, b + at, kPartialBytes); at += kPartialBytes;
        std::memcpy(&nd.max_score, b + at, 8); at += 8;
        uint16_t nk; std::memcpy(&nk, b + at, 2); at += 2;
        const int32_t me = (int32_t) nodes.size();
        nodes.push_back(nd);
        std::vector<std::pair<uint8_t, int32_t>> kids;
        for(uint16_t i = 0; i < nk && ok; i++) {
            if(at >= n) { ok = false; break; }
            const uint8_t byte = b[at++];
            kids.push_back({byte, parse(b, n, at, ok)});
        }
        nodes[me].first_child = (uint32_t) child_byte.size();
        nodes[me].n_children = (uint32_t) kids.size();
        for(auto& k: kids) { child_byte.push_back(k.first); child_ref.push_back(k.second); }
        return me;
    }

    // keys[lo, hi) are sorted, distinct (NUL-terminated) and agree on their first `depth` bytes
    int32_t build_range(const std::vector<vocab_entry>& v, const std::vector<std::string>& keys, size_t lo, size_t hi, size_t depth) {
        if(hi - lo == 1) {
            leaves.push_back({v[lo].token, v[lo].max_score, v[lo].num_ids, v[lo].list});
            return ~(int32_t) (leaves.size() - 1);
        }
        size_t lcp = depth;                       // first and last key of a sorted range bound the common prefix
        while(lcp < keys[lo].size() && lcp < keys[hi - 1].size() && keys[lo][lcp] == keys[hi - 1][lcp]) lcp++;
        node_t nd;
        std::memset(&nd, 0, sizeof nd);
        nd.partial_len = (uint8_t) std::min<size_t>(lcp - depth, 255);
        for(size_t i = 0; i < std::min<size_t>(lcp - depth, kPartialBytes); i++) nd.partial[i] = (uint8_t) keys[lo][depth + i];
        nd.max_score = std::numeric_limits<int64_t>::min();
        const int32_t me = (int32_t) nodes.size();
        nodes.push_back(nd);
        std::vector<std::pair<uint8_t, int32_t>> kids;
        for(size_t a = lo; a < hi;) {
            size_t e = a;
            while(e < hi && keys[e][lcp] == keys[a][lcp]) e++;
            kids.push_back({(uint8_t) keys[a][lcp], build_range(v, keys, a, e, lcp + 1)});
            a = e;
        }
        int64_t best = std::numeric_limits<int64_t>::min();
        for(size_t i = lo; i < hi; i++) best = std::max(best, v[i].max_score);
        nodes[me].max_score = best;
        nodes[me].first_child = (uint32_t) child_byte.size();
        nodes[me].n_children = (uint32_t) kids.size();
        for(auto& k: kids) { child_byte.push_back(k.first); child_ref.push_back(k.second); }
        return me;
    }

    struct search_t {
        std::vector<uint8_t> q;
        int min_cost, max_cost;
        bool prefix;
        std::vector<int32_t> hits;                // subtrees (or single leaves) every key of which is a candidate
        uint64_t visited = 0;                     // nodes and leaves the walk entered
    };

    // optimal-string-alignment distance, one more key byte `c` (previous byte `p`): prev2 / prev are the rows of the two
    // shorter key prefixes
    static void next_row(int depth, uint8_t p, uint8_t c, const std::vector<uint8_t>& q, const int* prev2, const int* prev, int* out) {
        out[0] = prev[0] + 1;
        for(size_t col = 1; col <= q.size(); col++) {
            const int subst = prev[col - 1] + (c == q[col - 1] ? 0 : 1);
            out[col] = std::min(std::min(out[col - 1] + 1, prev[col] + 1), subst);
            if(depth > 1 && col > 1 && c == q[col - 2] && p == q[col - 1]) out[col] = std::min(out[col], prev2[col - 2] + 1);
        }
    }

    // +1: every key below is a candidate; 0: read on; -1: give this branch up. Not a plain distance test: a cost that is
    // momentarily too high is tolerated when the next / previous query bytes explain it, and a prefix search accepts as soon
    // as the whole query has been consumed within bounds.
    static int search_state(const search_t& s, int key_index, uint8_t p, uint8_t c, const int* row) {
        const int qlen = (int) s.q.size();
        const bool key_ends = c == 0;
        const int key_len = key_ends ? key_index : key_index + 1;
        auto within = [&](int v, int hi) { return v >= s.min_cost && v <= hi; };
        if(key_ends) {
            if(within(row[qlen], s.max_cost)) return 1;
            // a long key that the query only extends ("strawberry" for q=strawberries)
            if(key_len > 5 && qlen > key_len && qlen - key_len <= s.max_cost && within(row[key_len], s.max_cost - 1)) return 1;
            return -1;
        }
        const int cost = row[std::min(key_len, qlen)];
        if(s.prefix && key_len >= qlen && within(cost, s.max_cost)) return 1;
        if(cost <= s.max_cost) return 0;
        // The reference reads query[key_index - 1] / [key_index - 2] without a bound (src/art.cpp:1557, 1587); once the key is
        // longer than the query that is the terminator or memory behind it. Defined here as 0, which never equals a key byte.
        auto qat = [&](int i) -> uint8_t { return i >= 0 && i < qlen ? s.q[(size_t) i] : 0; };
        if(cost == 2 || cost == 3) {
            if((key_index + 1 < qlen && qat(key_index + 1) == c) || (key_index > 0 && qat(key_index - 1) == c)) return 0;
        }
        if(cost == 3 || cost == 4) {
            if(key_index + 2 < qlen && qat(key_index + 1) == p && qat(key_index + 2) == c) return 0;
            if(key_index > 1 && qat(key_index - 2) == c) return 0;
        }
        return -1;
    }

    // depth -1: `ref` is the root and no byte has led to it. Children are visited from the largest byte down.
    static constexpr int kMaxCols = 104;               // query bytes (tokens are cut at 100) + terminator + column 0
    void walk(search_t& s, uint8_t p, uint8_t c, int32_t ref, int depth, const int* in_prev2, const int* in_prev) const {
        s.visited++;
        int rows[3][kMaxCols];                          // on the stack: a walk visits thousands of nodes
        const size_t cols = s.q.size() + 1;
        std::memcpy(rows[0], in_prev2, cols * sizeof(int));
        std::memcpy(rows[1], in_prev, cols * sizeof(int));
        int i2 = 0, i1 = 1, i0 = 2;                     // rows[i1] is the row of the bytes read so far
        auto feed = [&](uint8_t byte, bool advance) -> int {
            if(advance) { next_row(depth, p, byte, s.q, rows[i2], rows[i1], rows[i0]); const int t = i2; i2 = i1; i1 = i0; i0 = t; }
            return search_state(s, depth, p, byte, rows[i1]);
        };
        auto step = [&](uint8_t byte, bool advance) -> bool {     // false: this branch is decided
            const int a = feed(byte, advance);
            if(a == 1) s.hits.push_back(ref);
            if(a != 0) return false;
            p = byte; depth++;
            return true;
        };
        if(depth == -1) depth = 0;
        else if(!step(c, !(s.prefix && c == 0))) return;

        if(ref < 0) {
            const leaf_t& l = leaves[~ref];
            const int key_len = (int) l.key.size() + 1;
            const int iter_len = std::min(key_len, (int) s.q.size() + s.max_cost);      // look a little past the query for trailing typos
            if(depth >= iter_len) {                   // the path so far already spells the whole (relevant part of the) key
                if(search_state(s, depth, 0, 0, rows[i1]) == 1) s.hits.push_back(ref);
                return;
            }
            while(depth < iter_len) {
                c = (uint8_t) l.key.c_str()[depth];
                if(!step(c, !(s.prefix && c == 0))) return;
            }
            return;
        }
        const node_t& n = nodes[ref];
        int seen = std::min<int>(kPartialBytes, n.partial_len);
        for(int i = 0; i < seen; i++) { c = n.partial[i]; if(!step(c, true)) return; }
        // only the first kPartialBytes of a compressed path are stored: the rest is assumed to agree with the query
        while(seen < (int) n.partial_len && depth < (int) s.q.size()) { c = s.q[(size_t) depth]; if(!step(c, true)) return; seen++; }
        for(uint32_t k = n.n_children; k-- > 0;)
            walk(s, c, child_byte[n.first_child + k], child_ref[n.first_child + k], depth, rows[i2], rows[i1]);
    }

    bool admit(uint32_t leaf, int32_t exact_leaf, const std::string& prev_token, int32_t prev_leaf, const doc_tests& docs,
               std::set<std::string>& exclude_leaves, std::vector<uint32_t>& results) const {
        if((int32_t) leaf == exact_leaf) return false;
        const leaf_t& l = leaves[leaf];
        if(exclude_leaves.count(l.key)) return false;
        if(prev_token.empty() || prev_leaf < 0) {
            if(docs.filter_active && !docs.has_filter_doc(l.list)) return false;
        } else if(!docs.share_doc(leaves[prev_leaf].list, l.list)) return false;
        exclude_leaves.insert(l.key);
        results.push_back(leaf);
        return true;
    }

    // Best-first expansion of one matching subtree: a max-heap on the node's max_score (score order) or on a leaf's document
    // count with inner nodes counting 0 (frequency order); stops once 4 x max_words leaves are in hand.
    void collect(int32_t top, token_ordering order, size_t max_words, int32_t exact_leaf, const std::string& prev_token, int32_t prev_leaf,
                 const doc_tests& docs, std::set<std::string>& exclude_leaves, std::vector<uint32_t>& results) const {
        auto score_of = [&](int32_t r) { return r < 0 ? leaves[~r].max_score : nodes[r].max_score; };
        auto freq_of = [&](int32_t r) { return r < 0 ? leaves[~r].num_ids : 0u; };
        std::function<bool(int32_t, int32_t)> below;             // "a is not better than b" — deliberately not strict
        if(order == FREQUENCY) below = [&](int32_t a, int32_t b) { return !(freq_of(a) > freq_of(b)); };
        else below = [&](int32_t a, int32_t b) { return !(score_of(a) > score_of(b)); };
        std::priority_queue<int32_t, std::vector<int32_t>, std::function<bool(int32_t, int32_t)>> pq(below);
        pq.push(top);
        while(!pq.empty() && results.size() < max_words * 4) {
            const int32_t r = pq.top();
            pq.pop();
            if(r < 0) { admit((uint32_t) ~r, exact_leaf, prev_token, prev_leaf, docs, exclude_leaves, results); continue; }
            const node_t& n = nodes[r];
            for(uint32_t k = 0; k < n.n_children; k++) pq.push(child_ref[n.first_child + k]);
        }
    }
};

}  // namespace tsgpu
