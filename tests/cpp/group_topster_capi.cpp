// C entry points over the host layer's Topsters (typesense_b200/host/tsgpu_host.hpp: host_topster_t, host_group_topster_t) for
// tests/test_group_topster_ref.py, which compares them with the reference's own compiled Topster<KV> (oracle/_ref). Same signatures as
// oracle/ref_topster_wrap.cpp.
#include <cstddef>
#include <cstdint>

#include "../../typesense_b200/host/tsgpu_host.hpp"

extern "C" {

size_t host_topster(const uint64_t* keys, const int64_t* scores, size_t n, uint32_t capacity, uint64_t* out_keys) {
    tsgpu::host_topster_t t(capacity);
    for(size_t i = 0; i < n; i++) {
        tsgpu::KV kv{};
        kv.key = keys[i]; kv.distinct_key = keys[i];
        for(int s = 0; s < 3; s++) kv.scores[s] = scores[3 * i + s];
        t.add(kv);
    }
    const auto v = t.sort();
    for(size_t i = 0; i < v.size(); i++) out_keys[i] = v[i].key;
    return v.size();
}

size_t host_group_topster(const uint64_t* keys, const uint64_t* distinct, const int64_t* scores, size_t n, uint32_t capacity, uint32_t group_limit,
                          uint64_t* out_keys, uint64_t* out_distinct, uint32_t* out_group_sizes) {
    tsgpu::host_group_topster_t t(capacity, group_limit);
    for(size_t i = 0; i < n; i++) {
        tsgpu::KV kv{};
        kv.key = keys[i]; kv.distinct_key = distinct[i];
        for(int s = 0; s < 3; s++) kv.scores[s] = scores[3 * i + s];
        t.add(kv);
    }
    const auto groups = t.result();
    size_t w = 0;
    for(size_t g = 0; g < groups.size(); g++) {
        for(const auto& kv: groups[g]) { out_keys[w] = kv.key; out_distinct[w] = kv.distinct_key; w++; }
        out_group_sizes[g] = (uint32_t) groups[g].size();
    }
    return groups.size();
}

}  // extern "C"
