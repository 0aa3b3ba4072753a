// tsgpu_index_load_art / tsgpu_art_walk_batch (include/tsgpu.h): the device side of the typo / prefix
// candidate search — one thread per (token, cost) search walks the flat ART mirror with art_walk() (art_device.cuh).
// A translation unit of its own, like kw_regscore.cu: the kernels of tsgpu.cu that were measured this round keep their
// code byte for byte. Written after round 1's GPU budget was spent: checked on the CPU only (tests/test_art_mirror.py runs
// art_walk() compiled for the host against the host walk and the reference's art.cpp); tests/test_zz_art_gpu.py is its
// first run on a GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tsgpu.h"
#include "art_device.cuh"

// tsgpu.cu
extern "C" int tsgpu_index_device_(const tsgpu_index* idx);
extern "C" tsgpu_status tsgpu_fail_(tsgpu_status s, const char* msg);

namespace {

using namespace tsdev;

struct ArtState {
    std::vector<void*> alloc;
    ArtDev dev{};
    std::shared_ptr<std::vector<uint32_t>> node_rank, leaf_rank;          // pre-order ranks (host side): restore the recursion's hit order; shared so that a walk call can sort after it let go of the lock
};
// per tsgpu_index: the fields' mirrors, a stream of its own and a grow-only staging buffer; walk calls on one index serialise
struct ArtIndexState {
    std::unordered_map<uint32_t, ArtState> fields;
    cudaStream_t stream = nullptr;
    unsigned char* scratch = nullptr;
    size_t scratch_cap = 0;
    std::mutex call_mu;
};
std::mutex g_mu;                                                       // guards the table itself
std::unordered_map<const tsgpu_index*, ArtIndexState*> g_art;

void release(ArtState& s) { for(void* p: s.alloc) cudaFree(p); s.alloc.clear(); }
ArtIndexState* state_of(const tsgpu_index* idx, bool create) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_art.find(idx);
    if(it != g_art.end()) return it->second;
    if(!create) return nullptr;
    return g_art[idx] = new ArtIndexState;
}

#define CUA(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e__ = (call);                                                                                  \
        if(e__ != cudaSuccess) return tsgpu_fail_(TSGPU_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e__)).c_str()); \
    } while(0)

__global__ void __launch_bounds__(64)
art_walk_kernel(const ArtDev A, uint32_t n, const uint32_t* __restrict__ term_off, const uint8_t* __restrict__ terms,
                const uint8_t* __restrict__ min_cost, const uint8_t* __restrict__ max_cost, const uint8_t* __restrict__ prefix,
                int32_t* __restrict__ out_hits, uint32_t cap, uint32_t* __restrict__ out_counts, uint8_t* __restrict__ out_flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint32_t o = term_off[i], len = term_off[i + 1] - o;
    const bool pre = prefix[i] != 0;
    if(len + (pre ? 0u : 1u) > (uint32_t) kArtMaxQuery) { out_counts[i] = 0; out_flags[i] = 2; return; }
    ArtQuery Q;
    for(uint32_t k = 0; k < len; k++) Q.q[k] = terms[o + k];
    Q.qlen = (int) len;
    if(!pre) Q.q[Q.qlen++] = 0;                 // the key's terminator takes part in a whole-word match
    Q.min_cost = min_cost[i]; Q.max_cost = max_cost[i]; Q.prefix = pre;
    ArtFrame stack[kArtMaxStack];
    bool deep = false;
    const uint32_t cnt = art_walk(A, Q, out_hits + (size_t) i * cap, cap, stack, &deep);
    out_counts[i] = cnt;
    out_flags[i] = deep ? 1 : (cnt > cap ? 4 : 0);
}

// ---- frontier form: the node visit is the unit of parallelism -----------------------------------------------------------
// Level by level: every item of the current frontier (search, node about to be entered, its two DP rows) is entered by one
// thread; accepted subtrees go to the hit list, the children of nodes to descend into are appended to the next frontier.
struct Hit { uint32_t search; int32_t ref; };

__global__ void art_frontier_init_kernel(const ArtDev A, const ArtQuery* __restrict__ queries, const uint32_t* __restrict__ search_ids, uint32_t n,
                                         ArtWorkItem* __restrict__ items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    ArtWorkItem w;
    w.search = search_ids[i];
    art_root_item(A, queries[w.search], w.at);
    items[i] = w;
}

__global__ void __launch_bounds__(128)
art_frontier_kernel(const ArtDev A, const ArtQuery* __restrict__ queries, const ArtWorkItem* __restrict__ in, uint32_t n_in,
                    ArtWorkItem* __restrict__ out, uint32_t out_cap, uint32_t* __restrict__ counters /* 0 next items, 1 hits, 2 overflow */,
                    Hit* __restrict__ hits, uint32_t hit_cap) {
    // One slot allocation per WARP and level (prefix sum of the lanes' child counts, one atomicAdd by the last lane): the per-item
    // atomicAdd on one counter was the kernel's bound — millions of same-address atomics per level.
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool active = i < n_in;
    ArtWorkItem w;
    bool hit = false, descend = false;
    if(active) {
        w = in[i];
        descend = art_enter_fast(A, queries[w.search], w.at, &hit);
    }
    const uint32_t hit_mask = __ballot_sync(0xffffffffu, hit);
    if(hit_mask) {
        uint32_t hbase = 0;
        if(lane == 0) hbase = atomicAdd(counters + 1, (uint32_t) __popc(hit_mask));
        hbase = __shfl_sync(0xffffffffu, hbase, 0);
        if(hit) {
            const uint32_t pos = hbase + __popc(hit_mask & ((1u << lane) - 1u));
            if(pos < hit_cap) hits[pos] = Hit{w.search, w.at.ref}; else atomicOr(counters + 2, 1u);
        }
    }
    const uint32_t nch = descend ? A.nodes[w.at.ref].n_children : 0u;
    uint32_t incl = nch;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if((int) lane >= o) incl += t; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if(total == 0) return;
    uint32_t wbase = 0;
    if(lane == 31) wbase = atomicAdd(counters + 0, total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    if(wbase + total > out_cap) { if(lane == 0) atomicOr(counters + 2, 2u); return; }
    if(descend) {
        const uint32_t base = wbase + incl - nch;
        const uint32_t first = A.nodes[w.at.ref].first_child;
        w.at.p = w.at.c;                          // art_child_item(): a child is its parent's state + the byte and link that lead to it
        for(uint32_t k = 0; k < nch; k++) {
            w.at.ref = A.child_ref[first + k];
            w.at.c = A.child_byte[first + k];
            out[base + k] = w;
        }
    }
}

}  // namespace

// called by tsgpu_index_destroy
extern "C" __attribute__((visibility("hidden"))) void tsgpu_art_release_(const tsgpu_index* idx) {
    ArtIndexState* st = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_art.find(idx);
        if(it == g_art.end()) return;
        st = it->second;
        g_art.erase(it);
    }
    for(auto& f: st->fields) release(f.second);
    if(st->scratch) cudaFree(st->scratch);
    if(st->stream) cudaStreamDestroy(st->stream);
    delete st;
}

extern "C" tsgpu_status tsgpu_index_load_art(tsgpu_index* idx, uint32_t field, const tsgpu_art* a) {
    if(!idx || !a) return tsgpu_fail_(TSGPU_ERR_INVALID, "null argument");
    if(a->n_leaves && (!a->leaf_key_off || !a->leaf_keys)) return tsgpu_fail_(TSGPU_ERR_INVALID, "tsgpu_art: leaf arrays missing");
    if(a->n_nodes && (!a->node_first_child || !a->node_n_children || !a->node_partial_len || !a->node_partial || !a->child_byte || !a->child_ref))
        return tsgpu_fail_(TSGPU_ERR_INVALID, "tsgpu_art: node arrays missing");
    CUA(cudaSetDevice(tsgpu_index_device_(idx)));
    // validate the links once, so the kernel can follow them blindly
    std::vector<uint32_t> first(a->n_nodes);
    std::vector<uint16_t> nch(a->n_nodes);
    std::vector<uint8_t> plen(a->n_nodes), part((size_t) a->n_nodes * kArtPartialBytes), cbyte(a->n_children);
    std::vector<int32_t> cref(a->n_children);
    std::vector<uint64_t> koff((size_t) a->n_leaves + 1);
    if(a->n_nodes) {
        CUA(cudaMemcpy(first.data(), a->node_first_child, first.size() * 4, cudaMemcpyDefault));
        CUA(cudaMemcpy(nch.data(), a->node_n_children, nch.size() * 2, cudaMemcpyDefault));
        CUA(cudaMemcpy(plen.data(), a->node_partial_len, plen.size(), cudaMemcpyDefault));
        CUA(cudaMemcpy(part.data(), a->node_partial, part.size(), cudaMemcpyDefault));
    }
    if(a->n_children) {
        CUA(cudaMemcpy(cbyte.data(), a->child_byte, cbyte.size(), cudaMemcpyDefault));
        CUA(cudaMemcpy(cref.data(), a->child_ref, cref.size() * 4, cudaMemcpyDefault));
    }
    if(a->n_leaves) CUA(cudaMemcpy(koff.data(), a->leaf_key_off, koff.size() * 8, cudaMemcpyDefault));
    auto ref_ok = [&](int32_t r) { return r >= 0 ? (uint32_t) r < a->n_nodes : (uint32_t) ~r < a->n_leaves; };
    if(a->n_leaves && !ref_ok(a->root)) return tsgpu_fail_(TSGPU_ERR_INVALID, "tsgpu_art: root out of range");
    for(uint32_t i = 0; i < a->n_nodes; i++)
        if((uint64_t) first[i] + nch[i] > a->n_children) return tsgpu_fail_(TSGPU_ERR_INVALID, "tsgpu_art: child range out of bounds");
    for(uint32_t i = 0; i < a->n_children; i++) if(!ref_ok(cref[i])) return tsgpu_fail_(TSGPU_ERR_INVALID, "tsgpu_art: child ref out of range");
    for(uint32_t i = 0; i < a->n_leaves; i++) if(koff[i + 1] < koff[i]) return tsgpu_fail_(TSGPU_ERR_INVALID, "tsgpu_art: key offsets not ascending");
    std::vector<uint8_t> keys(a->n_leaves ? koff[a->n_leaves] : 0);
    if(!keys.empty()) CUA(cudaMemcpy(keys.data(), a->leaf_keys, keys.size(), cudaMemcpyDefault));
    std::vector<ArtNodeDev> nodes(a->n_nodes);
    for(uint32_t i = 0; i < a->n_nodes; i++) {
        nodes[i].first_child = first[i]; nodes[i].n_children = nch[i]; nodes[i].partial_len = plen[i]; nodes[i].pad = 0;
        for(int k = 0; k < kArtPartialBytes; k++) nodes[i].partial[k] = part[(size_t) i * kArtPartialBytes + k];
    }
    ArtState st;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> cudaError_t {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, bytes ? bytes : 16);
        if(e != cudaSuccess) return e;
        st.alloc.push_back(d);
        *dst = d;
        return bytes ? cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    cudaError_t e = up(nodes.data(), nodes.size() * sizeof(ArtNodeDev), (const void**) &st.dev.nodes);
    if(e == cudaSuccess) e = up(cbyte.data(), cbyte.size(), (const void**) &st.dev.child_byte);
    if(e == cudaSuccess) e = up(cref.data(), cref.size() * 4, (const void**) &st.dev.child_ref);
    if(e == cudaSuccess) e = up(koff.data(), koff.size() * 8, (const void**) &st.dev.leaf_key_off);
    if(e == cudaSuccess) e = up(keys.data(), keys.size(), (const void**) &st.dev.leaf_keys);
    if(e != cudaSuccess) { release(st); return tsgpu_fail_(TSGPU_ERR_CUDA, cudaGetErrorString(e)); }
    st.dev.root = a->root;
    st.dev.empty = a->n_leaves == 0 ? 1u : 0u;
    st.node_rank = std::make_shared<std::vector<uint32_t>>(a->n_nodes, 0u);
    st.leaf_rank = std::make_shared<std::vector<uint32_t>>(a->n_leaves, 0u);
    if(a->node_rank && a->leaf_rank) {
        if(a->n_nodes) cudaMemcpy(st.node_rank->data(), a->node_rank, (size_t) a->n_nodes * 4, cudaMemcpyDefault);
        if(a->n_leaves) cudaMemcpy(st.leaf_rank->data(), a->leaf_rank, (size_t) a->n_leaves * 4, cudaMemcpyDefault);
    } else if(a->n_leaves) {                  // pre-order with children from the largest byte down
        uint32_t next = 0;
        std::vector<int32_t> stack{a->root};
        while(!stack.empty()) {
            const int32_t r = stack.back();
            stack.pop_back();
            if(r < 0) { (*st.leaf_rank)[~r] = next++; continue; }
            (*st.node_rank)[r] = next++;
            if(next > a->n_nodes + a->n_leaves) { release(st); return tsgpu_fail_(TSGPU_ERR_INVALID, "tsgpu_art: the links do not form a tree"); }
            for(uint32_t k = 0; k < nch[r]; k++) stack.push_back(cref[first[r] + k]);
        }
    }
    CUA(cudaDeviceSynchronize());      // the uploads used the default stream; the walk stream is non-blocking and does not order with it
    ArtIndexState* is = state_of(idx, true);
    std::lock_guard<std::mutex> lk(is->call_mu);
    auto& slot = is->fields[field];
    release(slot);
    slot = st;
    return TSGPU_OK;
}

extern "C" tsgpu_status tsgpu_art_walk_batch(tsgpu_index* idx, uint32_t field, uint32_t n, const uint32_t* term_off, const uint8_t* terms,
                                             const uint8_t* min_cost, const uint8_t* max_cost, const uint8_t* prefix,
                                             int32_t* out_hits, uint32_t cap, uint32_t* out_counts, uint8_t* out_flags) {
    if(!idx) return tsgpu_fail_(TSGPU_ERR_INVALID, "null index");
    if(n == 0) return TSGPU_OK;
    if(!term_off || !terms || !min_cost || !max_cost || !prefix || !out_hits || !out_counts || !out_flags || cap == 0)
        return tsgpu_fail_(TSGPU_ERR_INVALID, "null argument");
    CUA(cudaSetDevice(tsgpu_index_device_(idx)));
    ArtIndexState* is = state_of(idx, false);
    if(!is) return tsgpu_fail_(TSGPU_ERR_INVALID, "no ART mirror loaded for this index");
    std::unique_lock<std::mutex> lk(is->call_mu);
    auto fit = is->fields.find(field);
    if(fit == is->fields.end()) return tsgpu_fail_(TSGPU_ERR_INVALID, "no ART mirror loaded for this field");
    const ArtDev A = fit->second.dev;
    // default: the frontier form (one thread per node visit). Measured (profiles/r02a_art_gpu_*.json, 200 K tokens, 4096 searches): one thread per
    // search ("dfs") needs 74 us per 2-typo search and, being one serial chain of dependent loads per search, ~0.7 ms per search when a
    // multi_search only has a few hundred walks to do (profiles/r02j); the frontier form spreads each search over the machine.
    static const bool frontier_mode = !(getenv("TSGPU_ART_MODE") && std::string(getenv("TSGPU_ART_MODE")) == "dfs");
    if(frontier_mode) {
        // ---- breadth-first: chunks of searches, one launch per tree level, hits sorted back into the recursion's order on the host
        const ArtState& AS = fit->second;
        std::vector<uint32_t> h_off((size_t) n + 1);
        CUA(cudaMemcpy(h_off.data(), term_off, h_off.size() * 4, cudaMemcpyDefault));
        for(uint32_t i = 0; i < n; i++) if(h_off[i + 1] < h_off[i]) return tsgpu_fail_(TSGPU_ERR_INVALID, "term offsets not ascending");
        std::vector<uint8_t> h_terms(h_off[n] + 1), h_min(n), h_max(n), h_pre(n);
        if(h_off[n]) CUA(cudaMemcpy(h_terms.data(), terms, h_off[n], cudaMemcpyDefault));
        CUA(cudaMemcpy(h_min.data(), min_cost, n, cudaMemcpyDefault));
        CUA(cudaMemcpy(h_max.data(), max_cost, n, cudaMemcpyDefault));
        CUA(cudaMemcpy(h_pre.data(), prefix, n, cudaMemcpyDefault));
        std::vector<ArtQuery> hq(n);
        std::vector<uint8_t> flags(n, 0);
        std::vector<uint32_t> counts(n, 0);
        static thread_local std::vector<int32_t> out;        // [n * cap], kept per calling thread (a fresh vector per call was 24 MB of page faults)
        if(out.size() < (size_t) n * cap) out.resize((size_t) n * cap);          // entries beyond counts[i] are unspecified
        for(uint32_t i = 0; i < n; i++) {
            const uint32_t len = h_off[i + 1] - h_off[i];
            ArtQuery& Q = hq[i];
            memset(&Q, 0, sizeof Q);
            if(len + (h_pre[i] ? 0u : 1u) > (uint32_t) kArtMaxQuery) { flags[i] = 2; continue; }
            memcpy(Q.q, h_terms.data() + h_off[i], len);
            Q.qlen = (int) len;
            if(!h_pre[i]) Q.q[Q.qlen++] = 0;
            Q.min_cost = h_min[i]; Q.max_cost = h_max[i]; Q.prefix = h_pre[i] != 0;
        }
        // searches per chunk: a 2-typo search over a large vocabulary has frontiers of 10^5 items, a 1-typo search of 10^3-10^4; a chunk
        // that overflows the item buffers is split in two and run again (a single search that does not fit goes back to the host walk)
        const uint32_t chunk = (uint32_t) std::max(1, getenv("TSGPU_ART_CHUNK") ? atoi(getenv("TSGPU_ART_CHUNK")) : 1024);
        const uint32_t chunk2 = (uint32_t) std::max(1, getenv("TSGPU_ART_CHUNK2") ? atoi(getenv("TSGPU_ART_CHUNK2")) : 64);
        const uint32_t item_cap = (uint32_t) std::max(1024, getenv("TSGPU_ART_ITEMS") ? atoi(getenv("TSGPU_ART_ITEMS")) : (16 << 20));
        const uint32_t hit_cap = std::max<uint32_t>(item_cap, std::max(chunk, chunk2) * cap);
        auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
        const size_t o_q = 0, o_ids = al((size_t) n * sizeof(ArtQuery)), o_cnt = al(o_ids + (size_t) std::max(chunk, chunk2) * 4), o_a = al(o_cnt + 16);
        const size_t o_b = al(o_a + (size_t) item_cap * sizeof(ArtWorkItem)), o_hits = al(o_b + (size_t) item_cap * sizeof(ArtWorkItem));
        const size_t total = o_hits + (size_t) hit_cap * sizeof(Hit);
        if(!is->stream) CUA(cudaStreamCreateWithFlags(&is->stream, cudaStreamNonBlocking));
        if(total > is->scratch_cap) {
            if(is->scratch) cudaFree(is->scratch);
            is->scratch = nullptr; is->scratch_cap = 0;
            CUA(cudaMalloc(&is->scratch, total));
            is->scratch_cap = total;
        }
        unsigned char* d = is->scratch;
        cudaStream_t st = is->stream;
        CUA(cudaMemcpyAsync(d + o_q, hq.data(), (size_t) n * sizeof(ArtQuery), cudaMemcpyHostToDevice, st));
        ArtWorkItem* bufs[2] = {(ArtWorkItem*) (d + o_a), (ArtWorkItem*) (d + o_b)};
        uint32_t* d_cnt = (uint32_t*) (d + o_cnt);
        // the chunks' hits are collected under the lock and put in order after it is released (the sort was a third of the call's time
        // and kept other requests' walks waiting)
        std::vector<Hit> all_hits;
        std::vector<size_t> chunk_end;
        std::vector<std::vector<uint32_t>> chunk_ids;
        const std::shared_ptr<std::vector<uint32_t>> node_rank = AS.node_rank, leaf_rank = AS.leaf_rank;
        auto rank_of = [&](int32_t r) { return r < 0 ? (*leaf_rank)[~r] : (*node_rank)[r]; };
        std::vector<std::vector<uint32_t>> work;               // chunks still to run (a stack)
        {
            std::vector<uint32_t> light, heavy;
            for(uint32_t i = 0; i < n; i++) if(flags[i] == 0) (h_max[i] >= 2 ? heavy : light).push_back(i);
            for(size_t b0 = 0; b0 < heavy.size(); b0 += chunk2) work.emplace_back(heavy.begin() + b0, heavy.begin() + std::min(heavy.size(), b0 + chunk2));
            for(size_t b0 = 0; b0 < light.size(); b0 += chunk) work.emplace_back(light.begin() + b0, light.begin() + std::min(light.size(), b0 + chunk));
        }
        static const bool art_timing = getenv("TSGPU_ART_TIMING") != nullptr;
        const auto t_begin = std::chrono::steady_clock::now();
        double ms_levels = 0, ms_sort = 0;
        size_t n_levels = 0, n_chunks = 0, n_hits_total = 0;
        while(!work.empty()) {
            std::vector<uint32_t> ids = std::move(work.back());
            work.pop_back();
            if(ids.empty()) continue;
            CUA(cudaMemcpyAsync(d + o_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, st));
            CUA(cudaMemsetAsync(d_cnt, 0, 16, st));
            art_frontier_init_kernel<<<((uint32_t) ids.size() + 127) / 128, 128, 0, st>>>(A, (const ArtQuery*) (d + o_q), (const uint32_t*) (d + o_ids),
                                                                                           (uint32_t) ids.size(), bufs[0]);
            CUA(cudaGetLastError());
            uint32_t n_cur = (uint32_t) ids.size(), h_cnt[4] = {0, 0, 0, 0};
            int cur = 0;
            bool overflow = false;
            const auto t_lv = std::chrono::steady_clock::now();
            n_chunks++;
            for(int level = 0; n_cur && level < 256; level++) {
                n_levels++;
                CUA(cudaMemsetAsync(d_cnt, 0, 4, st));                    // next-frontier counter only: hits accumulate over the levels
                art_frontier_kernel<<<(n_cur + 127) / 128, 128, 0, st>>>(A, (const ArtQuery*) (d + o_q), bufs[cur], n_cur, bufs[cur ^ 1], item_cap, d_cnt,
                                                                         (Hit*) (d + o_hits), hit_cap);
                CUA(cudaGetLastError());
                CUA(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
                CUA(cudaStreamSynchronize(st));
                if(h_cnt[2]) { overflow = true; break; }
                n_cur = h_cnt[0];
                cur ^= 1;
            }
            ms_levels += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_lv).count();
            if(overflow || n_cur) {
                if(ids.size() > 1) {                                  // too many items for the buffers: two half chunks
                    const size_t half = ids.size() / 2;
                    work.emplace_back(ids.begin(), ids.begin() + half);
                    work.emplace_back(ids.begin() + half, ids.end());
                } else flags[ids[0]] = 4;                             // the host walks this search
                continue;
            }
            n_hits_total += h_cnt[1];
            const size_t h0 = all_hits.size();
            all_hits.resize(h0 + h_cnt[1]);
            if(h_cnt[1]) CUA(cudaMemcpyAsync(all_hits.data() + h0, d + o_hits, (size_t) h_cnt[1] * sizeof(Hit), cudaMemcpyDeviceToHost, st));
            CUA(cudaStreamSynchronize(st));
            chunk_end.push_back(all_hits.size());
            chunk_ids.push_back(std::move(ids));
        }
        lk.unlock();                 // the device part is over: other requests' walks may start
        {
            const auto t_so = std::chrono::steady_clock::now();
            size_t h0 = 0;
            for(size_t c = 0; c < chunk_end.size(); c++) {
                std::sort(all_hits.begin() + h0, all_hits.begin() + chunk_end[c],
                          [&](const Hit& x, const Hit& y) { return x.search != y.search ? x.search < y.search : rank_of(x.ref) < rank_of(y.ref); });
                for(size_t i = h0; i < chunk_end[c]; i++) {
                    const Hit& hh = all_hits[i];
                    uint32_t& cnt_ = counts[hh.search];
                    if(cnt_ < cap) out[(size_t) hh.search * cap + cnt_] = hh.ref;
                    cnt_++;
                }
                for(uint32_t i: chunk_ids[c]) if(counts[i] > cap) flags[i] = 4;
                h0 = chunk_end[c];
            }
            ms_sort += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_so).count();
        }
        if(art_timing)
            fprintf(stderr, "[tsgpu art] %u searches: %.2f ms in %zu chunks / %zu levels (launch + sync per level), %.2f ms hit copy + sort (%zu hits), %.2f ms before the copies out\n",
                    n, ms_levels, n_chunks, n_levels, ms_sort, n_hits_total, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
        CUA(cudaMemcpy(out_counts, counts.data(), (size_t) n * 4, cudaMemcpyDefault));
        CUA(cudaMemcpy(out_flags, flags.data(), n, cudaMemcpyDefault));
        {   // a host destination gets the found entries only; a device destination the whole block
            cudaPointerAttributes pa{};
            const bool dev_dst = cudaPointerGetAttributes(&pa, out_hits) == cudaSuccess && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
            (void) cudaGetLastError();
            if(dev_dst) CUA(cudaMemcpy(out_hits, out.data(), (size_t) n * cap * 4, cudaMemcpyDefault));
            else for(uint32_t i = 0; i < n; i++) { const uint32_t c = std::min(counts[i], cap); if(c) memcpy(out_hits + (size_t) i * cap, out.data() + (size_t) i * cap, (size_t) c * 4); }
        }
        return TSGPU_OK;
    }
    std::vector<uint32_t> h_off((size_t) n + 1);
    CUA(cudaMemcpy(h_off.data(), term_off, h_off.size() * 4, cudaMemcpyDefault));
    for(uint32_t i = 0; i < n; i++) if(h_off[i + 1] < h_off[i]) return tsgpu_fail_(TSGPU_ERR_INVALID, "term offsets not ascending");
    const size_t n_bytes = h_off[n];
    auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t o_off = 0, o_terms = al((size_t) (n + 1) * 4), o_min = al(o_terms + n_bytes + 1), o_max = al(o_min + n), o_pre = al(o_max + n);
    const size_t o_cnt = al(o_pre + n), o_flag = al(o_cnt + (size_t) n * 4), o_hits = al(o_flag + n), total = o_hits + (size_t) n * cap * 4;
    if(!is->stream) CUA(cudaStreamCreateWithFlags(&is->stream, cudaStreamNonBlocking));
    if(total > is->scratch_cap) {
        if(is->scratch) cudaFree(is->scratch);
        is->scratch = nullptr; is->scratch_cap = 0;
        CUA(cudaMalloc(&is->scratch, total + total / 4));
        is->scratch_cap = total + total / 4;
    }
    unsigned char* d = is->scratch;
    cudaStream_t st = is->stream;
    CUA(cudaMemcpyAsync(d + o_off, h_off.data(), h_off.size() * 4, cudaMemcpyHostToDevice, st));
    if(n_bytes) CUA(cudaMemcpyAsync(d + o_terms, terms, n_bytes, cudaMemcpyDefault, st));
    CUA(cudaMemcpyAsync(d + o_min, min_cost, n, cudaMemcpyDefault, st));
    CUA(cudaMemcpyAsync(d + o_max, max_cost, n, cudaMemcpyDefault, st));
    CUA(cudaMemcpyAsync(d + o_pre, prefix, n, cudaMemcpyDefault, st));
    art_walk_kernel<<<(n + 63) / 64, 64, 0, st>>>(A, n, (const uint32_t*) (d + o_off), d + o_terms, d + o_min, d + o_max, d + o_pre,
                                                  (int32_t*) (d + o_hits), cap, (uint32_t*) (d + o_cnt), d + o_flag);
    CUA(cudaGetLastError());
    CUA(cudaMemcpyAsync(out_counts, d + o_cnt, (size_t) n * 4, cudaMemcpyDefault, st));
    CUA(cudaMemcpyAsync(out_flags, d + o_flag, n, cudaMemcpyDefault, st));
    CUA(cudaMemcpyAsync(out_hits, d + o_hits, (size_t) n * cap * 4, cudaMemcpyDefault, st));
    CUA(cudaStreamSynchronize(st));              // h_off and the caller's buffers are done with
    return TSGPU_OK;
}
