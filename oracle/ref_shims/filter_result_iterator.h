// TEST INFRASTRUCTURE (oracle/_ref build only).
// Minimal stand-in for the reference's include/filter_result_iterator.h (the real header pulls
// num_tree.h -> filter.h -> store.h -> <rocksdb/db.h>, absent in this image). It provides exactly the consumer
// interface the posting-list / or_iterator hot path touches: a materialised sorted-id filter
// with the reference's is_valid()/next()/reset() contract (src/filter_result_iterator.cpp:2131, :825, :2341 for the
// "is_filter_result_initialized" case, i.e. after compute_iterators()).
#pragma once
#include <cstdint>
#include <cstddef>
#include <map>
#include <string>
#include <vector>
#include <algorithm>

struct reference_filter_result_t {
    uint32_t count = 0;
    uint32_t* docs = nullptr;
};

struct single_filter_result_t {
    uint32_t seq_id = 0;
    std::map<std::string, reference_filter_result_t> reference_filter_results = {};
    bool is_reference_array_field = true;
    single_filter_result_t() = default;
    single_filter_result_t(uint32_t seq_id, std::map<std::string, reference_filter_result_t>&& refs,
                           bool is_reference_array_field = true)
        : seq_id(seq_id), reference_filter_results(std::move(refs)), is_reference_array_field(is_reference_array_field) {}
};

class filter_result_iterator_t {
    std::vector<uint32_t> ids;
    size_t idx = 0;
    bool provided = false;
public:
    enum validity_t : int { timed_out = -1, invalid = 0, valid = 1 };
    uint32_t seq_id = 0;
    std::map<std::string, reference_filter_result_t> reference;
    validity_t validity = invalid;
    uint32_t approx_filter_ids_length = 0;

    filter_result_iterator_t() = default;
    filter_result_iterator_t(const uint32_t* filter_ids, size_t n) : ids(filter_ids, filter_ids + n), provided(true) {
        approx_filter_ids_length = (uint32_t) n;
        reset();
    }
    bool is_filter_provided() const { return provided; }
    void reset(bool = false) {
        idx = 0;
        if(provided && !ids.empty()) { validity = valid; seq_id = ids[0]; } else { validity = invalid; }
    }
    void next() {
        if(validity != valid) return;
        if(++idx >= ids.size()) { validity = invalid; return; }
        seq_id = ids[idx];
    }
    void skip_to(uint32_t id) {
        if(validity != valid) return;
        idx = std::lower_bound(ids.begin() + idx, ids.end(), id) - ids.begin();
        if(idx >= ids.size()) { validity = invalid; return; }
        seq_id = ids[idx];
    }
    int is_valid(uint32_t id, const bool& = false) {
        if(validity == invalid) return -1;
        skip_to(id);
        return validity ? (seq_id == id ? 1 : 0) : -1;
    }
    void compute_iterators() {}
    // src/filter_result_iterator.cpp:2273: does the posting list `obj` hold one of the remaining filter ids? Used by
    // validate_and_add_leaf (src/art.cpp:1019-1023). Defined in oracle/ref_wrap.cpp on top of the reference's own
    // posting_t::contains_atleast_one.
    bool contains_atleast_one(const void* obj);
    const uint32_t* remaining_ids() const { return ids.data() + idx; }
    size_t remaining_count() const { return ids.size() - idx; }
};
