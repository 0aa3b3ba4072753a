// TEST INFRASTRUCTURE: C entry points around typesense_b200/host/art_mirror.hpp so tests/test_art_mirror.py can drive the
// mirror with the same calls it makes to the reference's art.cpp (oracle/_ref).
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../typesense_b200/host/art_mirror.hpp"
#include "../../typesense_b200/csrc/art_device.cuh"      // the device walk, compiled for the host

namespace {
struct handle_t {
    tsgpu::art_mirror_t m;
    std::vector<std::vector<uint32_t>> lists;       // posting ids per list id (for the document tests)
};
std::vector<std::string> split_nl(const char* s) {
    std::vector<std::string> out;
    for(const char* p = s; p && *p;) {
        const char* e = strchr(p, '\n');
        out.emplace_back(e ? std::string(p, e) : std::string(p));
        if(!e) break;
        p = e + 1;
    }
    return out;
}
bool intersects(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b, const std::vector<uint32_t>* c) {
    size_t i = 0, j = 0;
    while(i < a.size() && j < b.size()) {
        if(a[i] == b[j]) { if(!c || std::binary_search(c->begin(), c->end(), a[i])) return true; i++; j++; }
        else if(a[i] < b[j]) i++; else j++;
    }
    return false;
}
}

extern "C" {

void* am_load(const unsigned char* buf, size_t n) {
    auto* h = new handle_t;
    if(!h->m.load_export(buf, n)) { delete h; return nullptr; }
    return h;
}
void* am_build(const char* tokens_nl, const int64_t* scores, const uint32_t* dfs, uint32_t n) {
    auto* h = new handle_t;
    auto toks = split_nl(tokens_nl);
    std::vector<tsgpu::art_mirror_t::vocab_entry> v;
    for(uint32_t i = 0; i < n; i++) v.push_back({toks[i], scores[i], dfs[i], i});
    h->m.build(v);
    return h;
}
void am_free(void* h) { delete (handle_t*) h; }
// vocabulary in list-id order + its postings: binds every leaf to its list
void am_bind(void* hv, const char* tokens_nl, const uint64_t* list_off, const uint32_t* ids) {
    auto* h = (handle_t*) hv;
    auto toks = split_nl(tokens_nl);
    std::unordered_map<std::string, uint32_t> id_of;
    h->lists.assign(toks.size(), {});
    for(uint32_t l = 0; l < toks.size(); l++) { id_of[toks[l]] = l; h->lists[l].assign(ids + list_off[l], ids + list_off[l + 1]); }
    h->m.bind_lists([&](const std::string& k) { auto it = id_of.find(k); return it == id_of.end() ? 0xFFFFFFFFu : it->second; });
}
size_t am_num_nodes(void* hv) { return ((handle_t*) hv)->m.nodes.size(); }
size_t am_num_leaves(void* hv) { return ((handle_t*) hv)->m.leaves.size(); }

size_t am_fuzzy(void* hv, const char* term, int min_cost, int max_cost, size_t max_words, int order, int prefix, const char* prev_token,
                const uint32_t* filter_ids, size_t n_filter, int has_filter, const char* exclude, char* out, size_t out_cap) {
    auto* h = (handle_t*) hv;
    std::set<std::string> excl;
    for(auto& t: split_nl(exclude)) if(!t.empty()) excl.insert(t);
    std::vector<uint32_t> filt(filter_ids, filter_ids + n_filter);
    tsgpu::art_mirror_t::doc_tests docs;
    docs.filter_active = has_filter && n_filter > 0;            // an empty filter leaves the iterator invalid: no test
    docs.has_filter_doc = [&](uint32_t l) { return intersects(h->lists[l], filt, nullptr); };
    docs.share_doc = [&](uint32_t a, uint32_t b) { return intersects(h->lists[a], h->lists[b], docs.filter_active ? &filt : nullptr); };
    auto res = h->m.fuzzy_search(term, min_cost, max_cost, max_words, order == 1 ? tsgpu::art_mirror_t::MAX_SCORE : tsgpu::art_mirror_t::FREQUENCY,
                                 prefix != 0, prev_token ? prev_token : "", docs, excl);
    size_t w = 0;
    for(uint32_t li: res) {
        const std::string& k = h->m.leaves[li].key;
        if(w + k.size() + 1 >= out_cap) break;
        memcpy(out + w, k.data(), k.size()); w += k.size(); out[w++] = '\n';
    }
    if(out_cap) out[w < out_cap ? w : out_cap - 1] = 0;
    return res.size();
}


// am_fuzzy that also hands back the exclude set as the search left it (it grows by every leaf the search collected, also those
// the final truncation dropped — exactly what the reference's unique_tokens sees)
size_t am_fuzzy_ex(void* hv, const char* term, int min_cost, int max_cost, size_t max_words, int order, int prefix, const char* prev_token,
                   const char* exclude, char* out, size_t out_cap, char* exclude_out, size_t exclude_cap) {
    auto* h = (handle_t*) hv;
    std::set<std::string> excl;
    for(auto& t: split_nl(exclude)) if(!t.empty()) excl.insert(t);
    tsgpu::art_mirror_t::doc_tests docs;
    docs.share_doc = [&](uint32_t a, uint32_t b) { return intersects(h->lists[a], h->lists[b], nullptr); };
    auto res = h->m.fuzzy_search(term, min_cost, max_cost, max_words, order == 1 ? tsgpu::art_mirror_t::MAX_SCORE : tsgpu::art_mirror_t::FREQUENCY,
                                 prefix != 0, prev_token ? prev_token : "", docs, excl);
    auto put = [](const std::vector<std::string>& v, char* dst, size_t cap) {
        size_t w = 0;
        for(auto& k: v) { if(w + k.size() + 1 >= cap) break; memcpy(dst + w, k.data(), k.size()); w += k.size(); dst[w++] = '\n'; }
        if(cap) dst[w < cap ? w : cap - 1] = 0;
    };
    std::vector<std::string> keys;
    for(uint32_t li: res) keys.push_back(h->m.leaves[li].key);
    put(keys, out, out_cap);
    put(std::vector<std::string>(excl.begin(), excl.end()), exclude_out, exclude_cap);
    return res.size();
}

// walk_hits on the host mirror (mode 0) or through the device function art_walk() on the flattened arrays (mode 1); refs as int32
size_t am_walk(void* hv, int mode, const char* term, int min_cost, int max_cost, int prefix, int32_t* out, size_t cap, int* stack_overflow) {
    auto* h = (handle_t*) hv;
    *stack_overflow = 0;
    if(mode == 0) {
        auto hits = h->m.walk_hits(term, min_cost, max_cost, prefix != 0);
        for(size_t i = 0; i < hits.size() && i < cap; i++) out[i] = hits[i];
        return hits.size();
    }
    auto f = h->m.flatten();
    std::vector<tsdev::ArtNodeDev> nodes(h->m.nodes.size());
    for(size_t i = 0; i < nodes.size(); i++) {
        nodes[i].first_child = f.node_first_child[i]; nodes[i].n_children = f.node_n_children[i]; nodes[i].partial_len = f.node_partial_len[i];
        memcpy(nodes[i].partial, &f.node_partial[i * 8], 8); nodes[i].pad = 0;
    }
    tsdev::ArtDev A{nodes.data(), h->m.child_byte.data(), h->m.child_ref.data(), f.leaf_key_off.data(), f.leaf_keys.data(), h->m.root, h->m.empty ? 1u : 0u};
    tsdev::ArtQuery Q;
    const size_t tl = strlen(term);
    if(tl + (prefix ? 0 : 1) > (size_t) tsdev::kArtMaxQuery) { *stack_overflow = 2; return 0; }
    memcpy(Q.q, term, tl);
    Q.qlen = (int) tl;
    if(!prefix) Q.q[Q.qlen++] = 0;
    Q.min_cost = min_cost; Q.max_cost = max_cost; Q.prefix = prefix != 0;
    std::vector<tsdev::ArtFrame> stack(tsdev::kArtMaxStack);
    bool so = false;
    const uint32_t n = tsdev::art_walk(A, Q, out, (uint32_t) cap, stack.data(), &so);
    *stack_overflow = so ? 1 : 0;
    return n;
}


// the arrays of tsgpu_art, for handing a mirror to tsgpu_index_load_art from Python: sizes first (out == NULL), then the copy
// which: 0 node_first_child(u32) 1 node_n_children(u16) 2 node_partial_len(u8) 3 node_partial(u8 x8) 4 child_byte(u8) 5 child_ref(i32)
//        6 leaf_key_off(u64) 7 leaf_keys(u8); returns the element count
size_t am_flat(void* hv, int which, void* out) {
    auto* h = (handle_t*) hv;
    auto f = h->m.flatten();
    auto give = [&](const void* p, size_t n, size_t elem) { if(out && n) memcpy(out, p, n * elem); return n; };
    switch(which) {
        case 0: return give(f.node_first_child.data(), f.node_first_child.size(), 4);
        case 1: return give(f.node_n_children.data(), f.node_n_children.size(), 2);
        case 2: return give(f.node_partial_len.data(), f.node_partial_len.size(), 1);
        case 3: return give(f.node_partial.data(), f.node_partial.size(), 1);
        case 4: return give(h->m.child_byte.data(), h->m.child_byte.size(), 1);
        case 5: return give(h->m.child_ref.data(), h->m.child_ref.size(), 4);
        case 6: return give(f.leaf_key_off.data(), f.leaf_key_off.size(), 8);
        default: return give(f.leaf_keys.data(), f.leaf_keys.size(), 1);
    }
}
int32_t am_root(void* hv) { return ((handle_t*) hv)->m.root; }

// mode 2 of the walk: breadth-first over the flattened arrays with art_enter(), level by level as the frontier kernel will do it,
// hits sorted by pre-order rank at the end; returns the hit count, *levels and *peak the number of levels and the largest frontier
size_t am_walk_frontier(void* hv, const char* term, int min_cost, int max_cost, int prefix, int32_t* out, size_t cap, int* levels, size_t* peak) {
    auto* h = (handle_t*) hv;
    auto f = h->m.flatten();
    std::vector<tsdev::ArtNodeDev> nodes(h->m.nodes.size());
    for(size_t i = 0; i < nodes.size(); i++) {
        nodes[i].first_child = f.node_first_child[i]; nodes[i].n_children = f.node_n_children[i]; nodes[i].partial_len = f.node_partial_len[i];
        memcpy(nodes[i].partial, &f.node_partial[i * 8], 8); nodes[i].pad = 0;
    }
    tsdev::ArtDev A{nodes.data(), h->m.child_byte.data(), h->m.child_ref.data(), f.leaf_key_off.data(), f.leaf_keys.data(), h->m.root, h->m.empty ? 1u : 0u};
    tsdev::ArtQuery Q;
    const size_t tl = strlen(term);
    *levels = 0; *peak = 0;
    if(tl + (prefix ? 0 : 1) > (size_t) tsdev::kArtMaxQuery || h->m.empty) return 0;
    memcpy(Q.q, term, tl);
    Q.qlen = (int) tl;
    if(!prefix) Q.q[Q.qlen++] = 0;
    Q.min_cost = min_cost; Q.max_cost = max_cost; Q.prefix = prefix != 0;
    const auto ranks = h->m.preorder_ranks();
    std::vector<tsdev::ArtItem> cur(1), next;
    tsdev::art_root_item(A, Q, cur[0]);
    std::vector<int32_t> hits;
    while(!cur.empty()) {
        (*levels)++;
        *peak = std::max(*peak, cur.size());
        next.clear();
        for(auto& it: cur) {                               // on the device: one thread / warp per item
            bool hit = false;
            const bool descend = tsdev::art_enter_fast(A, Q, it, &hit);      // the form the frontier kernel calls (register rows for short queries)
            if(hit) hits.push_back(it.ref);
            if(descend) for(uint32_t k = 0; k < nodes[it.ref].n_children; k++) { next.emplace_back(); tsdev::art_child_item(A, Q, it, k, next.back()); }
        }
        cur.swap(next);
    }
    auto rank = [&](int32_t r) { return r < 0 ? ranks.leaf[~r] : ranks.node[r]; };
    std::sort(hits.begin(), hits.end(), [&](int32_t a, int32_t b) { return rank(a) < rank(b); });
    for(size_t i = 0; i < hits.size() && i < cap; i++) out[i] = hits[i];
    return hits.size();
}
unsigned long long am_last_visited() { return tsgpu::art_mirror_t::last_walk_visited(); }

}
