// Facet counting over result ids on sm_100a.
//
// Replaces (reference file:line):
//   Index::do_facets, hash-index branch            src/index.cpp:1674-1780   (per result id: the doc's facet ids -> count,
//                                                                              doc_id, array_pos)
//   the per-thread batches + aggregate_facet        src/index.cpp:4355-4400   (here: one histogram, no merge step)
//   Collection::search's top max_facet_values        src/collection.cpp:3252-3255 with facet_count_compare
//                                                    (include/collection.h:552-554): (count, id) descending
//
// Mirror of a facet field (facet_index_t's seq_id -> facet ids posting list, facet ids as the dense counter the reference
// assigns): CSR doc_off[n_docs+1] / value_ids[]. Counting is an HBM-bound histogram: one thread per result id (or per
// word of a query's all_result_ids bitmap), one atomicAdd per (doc, distinct facet id); `doc_id` / `array_pos` of a value
// are those of the largest result id that holds it — what one sequential do_facets pass over the ascending ids leaves
// (the reference's own multi-threaded merge order is not deterministic). Selection is a CTA-local bitonic top-N.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsfc {

struct FacetDev {
    uint32_t n_values;
    const unsigned long long* doc_off;     // [n_docs + 1]
    const uint32_t* value_ids;
};

struct FacetCountOut { uint32_t value_id, count, doc_id, array_pos; };     // tsgpu_facet_count

__device__ __forceinline__ void facet_count_doc(const FacetDev& f, uint32_t doc, uint32_t* __restrict__ cnt, unsigned long long* __restrict__ last) {
    const unsigned long long o0 = f.doc_off[doc], o1 = f.doc_off[doc + 1];
    for(unsigned long long j = o0; j < o1; j++) {
        const uint32_t v = __ldg(f.value_ids + j);
        if(v >= f.n_values) continue;
        bool dup = false;                                  // std::set<uint32_t> unique_facet_hashes of the reference
        for(unsigned long long k = o0; k < j && !dup; k++) dup = __ldg(f.value_ids + k) == v;
        if(dup) continue;
        atomicAdd(cnt + v, 1u);
        atomicMax(last + v, ((unsigned long long) doc << 32) | (uint32_t) (j - o0));
    }
}

// explicit ascending result ids (one query)
__global__ void __launch_bounds__(256)
facet_count_ids_kernel(const __grid_constant__ FacetDev f, const uint32_t* __restrict__ ids, size_t n, uint32_t sample_mod, uint32_t n_docs,
                       uint32_t* __restrict__ cnt, unsigned long long* __restrict__ last) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(sample_mod > 1 && (i % sample_mod) != 0) return;    // estimate_facets: every sample_mod-th result
    const uint32_t doc = ids[i];
    if(doc < n_docs) facet_count_doc(f, doc, cnt, last);
}

// all_result_ids bitmaps of a batch: grid = (word tiles, queries); histograms are [query][n_values]
__global__ void __launch_bounds__(256)
facet_count_bitmaps_kernel(const __grid_constant__ FacetDev f, const uint32_t* const* __restrict__ q_bitmap, uint32_t n_words, uint32_t n_docs,
                           uint32_t* __restrict__ cnt, unsigned long long* __restrict__ last) {
    const uint32_t q = blockIdx.y;
    const uint32_t* bm = q_bitmap[q];
    if(!bm) return;
    uint32_t* c = cnt + (size_t) q * f.n_values;
    unsigned long long* l = last + (size_t) q * f.n_values;
    for(uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += gridDim.x * blockDim.x) {
        uint32_t w = __ldg(bm + wi);
        while(w) {
            const uint32_t bit = __ffs(w) - 1;
            w &= w - 1;
            const uint32_t doc = (wi << 5) | bit;
            if(doc < n_docs) facet_count_doc(f, doc, c, l);
        }
    }
}

// One CTA per query: the top_n values by (count, id) descending, and the number of values with a count.
constexpr int kTopThreads = 256;
__device__ __forceinline__ void key_sort_desc(unsigned long long* k, uint32_t N) {       // bitonic, N power of two
    for(uint32_t sz = 2; sz <= N; sz <<= 1)
        for(uint32_t j = sz >> 1; j > 0; j >>= 1) {
            for(uint32_t t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i + j;
                const bool up = ((i & sz) == 0);
                const unsigned long long a = k[i], b = k[p];
                if(up ? (a < b) : (a > b)) { k[i] = b; k[p] = a; }
            }
            __syncthreads();
        }
}
__global__ void __launch_bounds__(kTopThreads)
facet_topn_kernel(uint32_t n_values, const uint32_t* __restrict__ cnt, const unsigned long long* __restrict__ last, uint32_t top_n, uint32_t NP,
                  FacetCountOut* __restrict__ out, uint32_t* __restrict__ out_n, uint32_t* __restrict__ out_distinct) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);      // [2 * NP], key = count << 32 | value id
    __shared__ uint32_t s_n, s_distinct;
    __shared__ unsigned long long s_thr;
    const uint32_t q = blockIdx.x;
    const uint32_t* c = cnt + (size_t) q * n_values;
    const unsigned long long* l = last + (size_t) q * n_values;
    const uint32_t N2 = 2 * NP;
    if(threadIdx.x == 0) { s_n = 0; s_distinct = 0; s_thr = 0; }
    __syncthreads();
    for(uint32_t base = 0; base < n_values; base += kTopThreads) {
        const uint32_t v = base + threadIdx.x;
        const uint32_t cv = v < n_values ? c[v] : 0;
        const unsigned long long key = ((unsigned long long) cv << 32) | v;
        const bool nz = cv != 0;
        const bool keep = nz && key > s_thr;
        if(nz) atomicAdd(&s_distinct, 1u);
        if(keep) { const uint32_t slot = atomicAdd(&s_n, 1u); keys[slot] = key; }      // s_n + 256 <= N2 is kept below
        __syncthreads();
        if(s_n + kTopThreads > N2) {
            const uint32_t n = s_n;
            for(uint32_t i = n + threadIdx.x; i < N2; i += kTopThreads) keys[i] = 0;
            __syncthreads();
            key_sort_desc(keys, N2);
            if(threadIdx.x == 0) { s_n = n < top_n ? n : top_n; if(n >= top_n) s_thr = keys[top_n - 1]; }
            __syncthreads();
        }
    }
    const uint32_t n = s_n;
    for(uint32_t i = n + threadIdx.x; i < N2; i += kTopThreads) keys[i] = 0;
    __syncthreads();
    key_sort_desc(keys, N2);
    const uint32_t m = n < top_n ? n : top_n;
    for(uint32_t i = threadIdx.x; i < m; i += kTopThreads) {
        const unsigned long long k = keys[i];
        const uint32_t v = (uint32_t) k;
        const unsigned long long lv = l[v];
        out[(size_t) q * top_n + i] = FacetCountOut{v, (uint32_t) (k >> 32), (uint32_t) (lv >> 32), (uint32_t) lv};
    }
    if(threadIdx.x == 0) { out_n[q] = m; out_distinct[q] = s_distinct; }
}

// a query's all_result_ids bitmap -> ascending ids (count first, then write)
__global__ void __launch_bounds__(256)
bitmap_count_kernel(const uint32_t* __restrict__ bm, uint32_t n_words, uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t s[8];
    const uint32_t wi = blockIdx.x * 256 + threadIdx.x;
    uint32_t c = wi < n_words ? __popc(bm[wi]) : 0;
    for(int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if(threadIdx.x == 0) { uint32_t t = 0; for(int i = 0; i < 8; i++) t += s[i]; tile_cnt[blockIdx.x] = t; }
}

}  // namespace tsfc
