// TEST INFRASTRUCTURE: C entry points around typesense_b200/host/art_mirror.hpp so tests/test_art_mirror.py can drive the
// mirror with the same calls it makes to the reference's art.cpp (oracle/_ref).
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../typesense_b200/host/art_mirror.hpp"

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

}
