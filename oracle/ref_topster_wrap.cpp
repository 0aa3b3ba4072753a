// TEST INFRASTRUCTURE (part of oracle/_ref/liboracle_ref.so): the reference's own Topster<KV> (include/topster.h, compiled where it lies)
// behind a C entry point, so that the host layer's group Topster (typesense_b200/host/tsgpu_host.hpp: host_group_topster_t) and the
// oracle's Topster restatement are pinned on the reference's compiled code, not only on its test tables.
#include <cstddef>
#include <cstdint>
#include <vector>

#include "topster.h"

extern "C" {

// Topster<KV>(capacity).add(...) for every KV, sort(): keys in result order. scores: [n * 3].
size_t ref_topster(const uint64_t* keys, const int64_t* scores, size_t n, uint32_t capacity, uint64_t* out_keys) {
    Topster<KV> t(capacity);
    for(size_t i = 0; i < n; i++) { KV kv(0, keys[i], keys[i], 0, scores + 3 * i); t.add(&kv); }
    t.sort();
    for(uint32_t i = 0; i < t.size; i++) out_keys[i] = t.getKeyAt(i);
    return t.size;
}

// group_by: Topster<KV>(capacity, distinct = group_limit, first pass = false) fed with every KV (include/topster.h:357-376), then the
// distinct branch of Index::populate_result_kvs (src/index.cpp:8968-9013; a member of Index, which cannot be compiled here — its dozen
// lines are restated: every group's Topster sorted, the heads through a Topster of `capacity`, the groups in that Topster's order).
// Output: the groups one after the other (keys, distinct keys), out_group_sizes[g] hits each; returns the number of groups.
size_t ref_group_topster(const uint64_t* keys, const uint64_t* distinct, const int64_t* scores, size_t n, uint32_t capacity, uint32_t group_limit,
                         uint64_t* out_keys, uint64_t* out_distinct, uint32_t* out_group_sizes) {
    Topster<KV> t(capacity, group_limit, false);
    for(size_t i = 0; i < n; i++) { KV kv(0, keys[i], distinct[i], 0, scores + 3 * i); t.add(&kv); }
    Topster<KV> heads(t.MAX_SIZE);
    for(auto& g: t.group_kv_map) {
        g.second->sort();
        if(g.second->size != 0) heads.add(g.second->getKV(0));
    }
    heads.sort();
    size_t w = 0;
    for(uint32_t i = 0; i < heads.size; i++) {
        auto* gt = t.group_kv_map[heads.getKV(i)->distinct_key];
        for(uint32_t j = 0; j < gt->size; j++, w++) { out_keys[w] = gt->getKV(j)->key; out_distinct[w] = gt->getKV(j)->distinct_key; }
        out_group_sizes[i] = gt->size;
    }
    return heads.size;
}

}  // extern "C"
