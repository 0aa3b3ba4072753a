// TEST INFRASTRUCTURE (part of oracle/_ref/liboracle_ref.so): the reference's own numeric index (src/num_tree.cpp, compiled where it
// lies behind the filter.h stand-in) behind a C entry point — pins the comparator semantics that tests/test_filters_device.py expects
// of tsgpu_filter_numeric.
#include <cstddef>
#include <cstdint>
#include <climits>
#include <vector>

#include "num_tree.h"

extern "C" {

// values[d] = the document's value (INT64_MIN: none — the document is simply not in the tree, as for an optional field).
// op: include/tsgpu.h TSGPU_CMP_* (0 =, 1 !=, 2 <, 3 <=, 4 >, 5 >=, 6 range [v1, v2]). Returns the number of ids written (ascending).
size_t ref_num_tree_search(const int64_t* values, uint32_t n_docs, int op, int64_t v1, int64_t v2, uint32_t* out_ids) {
    num_tree_t tree;
    for(uint32_t d = 0; d < n_docs; d++) if(values[d] != INT64_MIN) tree.insert(values[d], d);
    uint32_t* ids = nullptr;
    size_t n = 0;
    switch(op) {
        case 0: case 1: tree.search(EQUALS, v1, &ids, n); break;
        case 2: tree.search(LESS_THAN, v1, &ids, n); break;
        case 3: tree.search(LESS_THAN_EQUALS, v1, &ids, n); break;
        case 4: tree.search(GREATER_THAN, v1, &ids, n); break;
        case 5: tree.search(GREATER_THAN_EQUALS, v1, &ids, n); break;
        case 6: tree.range_inclusive_search(v1, v2, &ids, n); break;
        default: return 0;
    }
    size_t w = 0;
    if(op == 1) {
        // `!=`: the equal ids are taken out of ALL the collection's ids (filter_result_iterator_t::apply_not_equals over index->seq_ids,
        // src/filter_result_iterator.cpp) — the complement is restated here, the equal set above is the reference's
        size_t j = 0;
        for(uint32_t d = 0; d < n_docs; d++) {
            while(j < n && ids[j] < d) j++;
            if(j < n && ids[j] == d) continue;
            out_ids[w++] = d;
        }
    } else {
        for(size_t i = 0; i < n; i++) out_ids[w++] = ids[i];
    }
    delete[] ids;
    return w;
}

}  // extern "C"
