// Vector hot path on sm_100a: batched HNSW search (K6), flat filtered scan (K7).
//
// Replaces (reference file:line):
//   hnswlib::HierarchicalNSW<float>::searchKnnCloserFirst(q, k, ef, filter)   call site src/index.cpp:3384-3386
//       (hnswlib itself is a build-time download, cmake/hnsw.cmake:3 — its published searchKnn / searchBaseLayerST
//        algorithm is what is implemented)
//   VectorFilterFunctor::operator()                                            include/index.h:338-353
//   InnerProductSpace distance 1 - <a,b>, process_results_bruteforce           src/index.cpp:3345-3374
//
// Execution model: the graph (level-0 link table, upper-level records) and the fp32 vectors stay
// resident in HBM. ONE WARP PER QUERY, persistent warps pulling queries from an atomic counter:
//   - the query vector lives in registers (d/128 float4 per lane); a neighbour's vector is read with coalesced
//     512-byte warp loads, two neighbours (2*d/128 LDG.128 per lane) in flight before the first FMA;
//   - dot products use the fixed "W128" order (element e -> accumulator e mod 128 by FMA, 4->1 per lane, xor
//     butterfly), the same order the oracle uses, so CPU and GPU take identical branches on the same graph;
//   - visited set = a two-tier hash set: shared memory first, a per-slot HBM table for long walks (see vis_insert);
//   - result heap (size ef) in shared memory, candidate heap in a per-warp global arena (L2-resident);
//     both are binary heaps of u64 keys (order-preserving float bits << 32 | id) reproducing std::priority_queue<
//     pair<float,id>> order exactly;
//   - the 32 neighbours of an expansion are tested/marked in parallel (ballot), distances are computed two at a
//     time, and admission to the heaps is replayed in neighbour order by lane 0 — the same sequence of
//     lowerBound updates as hnswlib's loop.
// This kernel is HBM-latency/bandwidth bound (one 4*d-byte vector per 2*d flop); tensor cores are not used here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace tsv {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kKnnThreads = 128;            // 4 warps = 4 queries per CTA
constexpr uint32_t kCandSmem = 256;         // candidate-heap entries kept in shared memory per warp (rest spills to HBM)
constexpr uint32_t kVisSmem = 2048;         // visited set, tier 1: open-addressing slots in shared memory per warp (power of two)
constexpr uint32_t kVisSmemLimit = 1536;    // ... filled to 75 %, then tier 2 (a per-slot table in HBM) takes the new keys

struct HnswDev {
    uint32_t n_nodes, dim, M, max_level, entry_point, metric;
    const float* vectors;
    const uint32_t* labels;        // nullptr = identity
    const uint8_t* levels;
    const uint32_t* links0;        // [n*(2M+1)]
    const unsigned long long* upper_off;
    const uint32_t* links_up;
    const uint32_t* deleted;       // one bit per label: hnswlib's markDelete (never a result, still traversed); nullptr = none
};

struct KnnParams {
    const float* queries;          // [nq*dim]
    uint32_t nq, k, ef;
    const uint32_t* const* q_filter_bitmap;   // [nq] device pointers (nullptr = no filter) or nullptr array
    const uint32_t* const* q_excl;            // [nq] or nullptr
    const uint32_t* q_n_excl;
    const uint8_t* q_skip;                    // [nq] 1 = query not run (e.g. goes to the flat path) or nullptr
    const uint32_t* q_order;                  // [n_order] queries in hand-out order (longest walks first) or nullptr = 0..nq-1
    uint32_t n_order;
    const uint32_t* n_order_dev;              // when set, the number of tickets is read from device memory (retry launch)
    float* out_dist;               // [nq*k]
    uint32_t* out_labels;          // [nq*k]
    uint32_t* out_n;               // [nq]
    // per-warp-slot scratch
    uint32_t* vis2;                // [n_slots * vis2_slots] tier 2 of the visited set, all-zero between queries
    uint32_t vis2_slots;           // power of two
    unsigned long long* cand;      // [n_slots * cand_cap]
    uint32_t cand_cap;
    uint32_t* counter;             // query ticket
    unsigned long long* stats;     // [0] n_dist, [1] n_expanded, [2] speculation hits, [3] walks that used tier 2
    uint32_t* retry_n;             // walks that outgrew the per-slot scratch: re-run by the host with a full-size slot
    uint32_t* retry_list;          // [nq]
    uint32_t vis_cache;            // register-queue kernel: entries of the shared-memory visited cache per walk (power of two)
    uint32_t walk_prefetch;        // register-queue kernel: L2 hints for the likely next expansion (bit 0 table line, 1 rows, 2 filter word)
    uint32_t* q_work;              // [2*nq] expansions, distance evaluations of each walk (instrumentation; may be nullptr)
};

__device__ __forceinline__ uint32_t ord_f32(float f) {          // order-preserving float -> u32
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(u);
}

// W128 dot product. NCH = dim/128 when dim is a multiple of 128 (query in registers), 0 = generic (scalar, guarded).
template <int NCH>
struct QReg { float4 v[NCH > 0 ? NCH : 1]; };

__device__ __forceinline__ float warp_tree(float a0, float a1, float a2, float a3) {
    float t = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
#pragma unroll
    for(int off = 16; off >= 1; off >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, off));
    return t;
}

template <int NCH>
__device__ __forceinline__ float dot_one(const QReg<NCH>& q, const float* __restrict__ qs, const float* __restrict__ v,
                                         uint32_t dim, uint32_t lane) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if(NCH > 0) {
        float4 x[NCH > 0 ? NCH : 1];
#pragma unroll
        for(int s = 0; s < NCH; s++) x[s] = __ldg(reinterpret_cast<const float4*>(v + s * 128) + lane);
#pragma unroll
        for(int s = 0; s < NCH; s++) {
            a0 = __fmaf_rn(q.v[s].x, x[s].x, a0); a1 = __fmaf_rn(q.v[s].y, x[s].y, a1);
            a2 = __fmaf_rn(q.v[s].z, x[s].z, a2); a3 = __fmaf_rn(q.v[s].w, x[s].w, a3);
        }
    } else {
        for(uint32_t s = 0; s * 128 < dim; s++) {
            const uint32_t e = s * 128 + 4 * lane;
            if(e + 0 < dim) a0 = __fmaf_rn(qs[e + 0], __ldg(v + e + 0), a0);
            if(e + 1 < dim) a1 = __fmaf_rn(qs[e + 1], __ldg(v + e + 1), a1);
            if(e + 2 < dim) a2 = __fmaf_rn(qs[e + 2], __ldg(v + e + 2), a2);
            if(e + 3 < dim) a3 = __fmaf_rn(qs[e + 3], __ldg(v + e + 3), a3);
        }
    }
    return warp_tree(a0, a1, a2, a3);
}

// two vectors with all loads issued before the first FMA (memory-level parallelism)
template <int NCH>
__device__ __forceinline__ void dot_two(const QReg<NCH>& q, const float* __restrict__ qs, const float* __restrict__ va,
                                        const float* __restrict__ vb, uint32_t dim, uint32_t lane, float& da, float& db) {
    if(NCH > 0) {
        float4 x[NCH > 0 ? NCH : 1], y[NCH > 0 ? NCH : 1];
#pragma unroll
        for(int s = 0; s < NCH; s++) {
            x[s] = __ldg(reinterpret_cast<const float4*>(va + s * 128) + lane);
            y[s] = __ldg(reinterpret_cast<const float4*>(vb + s * 128) + lane);
        }
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
        for(int s = 0; s < NCH; s++) {
            a0 = __fmaf_rn(q.v[s].x, x[s].x, a0); a1 = __fmaf_rn(q.v[s].y, x[s].y, a1);
            a2 = __fmaf_rn(q.v[s].z, x[s].z, a2); a3 = __fmaf_rn(q.v[s].w, x[s].w, a3);
            b0 = __fmaf_rn(q.v[s].x, y[s].x, b0); b1 = __fmaf_rn(q.v[s].y, y[s].y, b1);
            b2 = __fmaf_rn(q.v[s].z, y[s].z, b2); b3 = __fmaf_rn(q.v[s].w, y[s].w, b3);
        }
        da = warp_tree(a0, a1, a2, a3);
        db = warp_tree(b0, b1, b2, b3);
    } else {
        da = dot_one<NCH>(q, qs, va, dim, lane);
        db = dot_one<NCH>(q, qs, vb, dim, lane);
    }
}

// ---- 4-ary heaps of u64 keys, single-lane operations. Keys are unique ((distance, id) pairs), so the pop order — the
// only thing the search observes — is the key order whatever the heap's shape; four children per node halve the depth
// of the dependent load chain of a sift and their loads are independent.
__device__ __forceinline__ void heap_push_max(unsigned long long* h, uint32_t& n, unsigned long long key) {
    uint32_t i = n++;
    while(i > 0) { const uint32_t p = (i - 1) >> 2; const unsigned long long hp = h[p]; if(hp >= key) break; h[i] = hp; i = p; }
    h[i] = key;
}
__device__ __forceinline__ void heap_sift_max(unsigned long long* h, uint32_t n, unsigned long long key) {
    uint32_t i = 0;
    for(;;) {
        const uint32_t c = 4 * i + 1;
        if(c >= n) break;
        unsigned long long best = h[c]; uint32_t bi = c;
        const unsigned long long h1 = c + 1 < n ? h[c + 1] : 0ull, h2 = c + 2 < n ? h[c + 2] : 0ull, h3 = c + 3 < n ? h[c + 3] : 0ull;
        if(c + 1 < n && h1 > best) { best = h1; bi = c + 1; }
        if(c + 2 < n && h2 > best) { best = h2; bi = c + 2; }
        if(c + 3 < n && h3 > best) { best = h3; bi = c + 3; }
        if(best <= key) break;
        h[i] = best; i = bi;
    }
    h[i] = key;
}
__device__ __forceinline__ void heap_pop_max(unsigned long long* h, uint32_t& n) {
    const unsigned long long last = h[--n];
    if(n) heap_sift_max(h, n, last);
}
// pop the maximum and push `key` (key below the current maximum) in one sift-down
__device__ __forceinline__ void heap_replace_max(unsigned long long* h, uint32_t n, unsigned long long key) { heap_sift_max(h, n, key); }

__device__ __forceinline__ void heap_push_min(unsigned long long* h, uint32_t& n, unsigned long long key) {
    uint32_t i = n++;
    while(i > 0) { const uint32_t p = (i - 1) >> 2; const unsigned long long hp = h[p]; if(hp <= key) break; h[i] = hp; i = p; }
    h[i] = key;
}
__device__ __forceinline__ void heap_pop_min(unsigned long long* h, uint32_t& n) {
    const unsigned long long last = h[--n];
    if(n == 0) return;
    uint32_t i = 0;
    for(;;) {
        const uint32_t c = 4 * i + 1;
        if(c >= n) break;
        unsigned long long best = h[c]; uint32_t bi = c;
        const unsigned long long h1 = c + 1 < n ? h[c + 1] : ~0ull, h2 = c + 2 < n ? h[c + 2] : ~0ull, h3 = c + 3 < n ? h[c + 3] : ~0ull;
        if(h1 < best) { best = h1; bi = c + 1; }
        if(h2 < best) { best = h2; bi = c + 2; }
        if(h3 < best) { best = h3; bi = c + 3; }
        if(best >= last) break;
        h[i] = best; i = bi;
    }
    h[i] = last;
}

// result key: max-heap on (dist, id)            -> ord(dist) << 32 | id
// candidate key: min-heap on (dist, larger id first) (priority_queue<pair<-dist,id>> top) -> ord(dist) << 32 | ~id
__device__ __forceinline__ unsigned long long res_key(float d, uint32_t id) { return ((unsigned long long) ord_f32(d) << 32) | id; }
__device__ __forceinline__ unsigned long long cand_key(float d, uint32_t id) { return ((unsigned long long) ord_f32(d) << 32) | (uint32_t) ~id; }

__device__ __forceinline__ bool is_excluded(const uint32_t* excl, uint32_t n_excl, uint32_t label) {
    uint32_t lo = 0, hi = n_excl;
    while(lo < hi) { const uint32_t mid = (lo + hi) >> 1; if(excl[mid] < label) lo = mid + 1; else hi = mid; }
    return lo < n_excl && excl[lo] == label;
}
__device__ __forceinline__ bool allowed(const HnswDev& g, const uint32_t* fbm, const uint32_t* excl, uint32_t n_excl, uint32_t node) {
    const uint32_t label = g.labels ? __ldg(g.labels + node) : node;
    if(g.deleted && ((__ldg(g.deleted + (label >> 5)) >> (label & 31)) & 1u)) return false;       // isMarkedDeleted
    if(n_excl && is_excluded(excl, n_excl, label)) return false;
    if(!fbm) return true;
    return (__ldg(fbm + (label >> 5)) >> (label & 31)) & 1;
}

// ---- visited set (hnswlib's VisitedList) as a two-tier hash set. Round 1 kept one bit per node in a per-warp bitmap in
// HBM: at 10 M nodes that is 1.25 MB per walk, 3.7 GB over the resident walks, so every test-and-set was a DRAM sector
// read + write-back (6.7 GB of the 32.7 GB the kernel moved per launch, profiles/r01c) and a DRAM round trip on the
// critical path of every expansion. A walk touches ~1.3 K nodes, so the set fits in shared memory: tier 1 is an
// open-addressing table of kVisSmem keys per warp (shared-memory CAS, no HBM traffic at all); walks that outgrow it
// (selective filters) continue in a per-slot table in HBM. Membership is exact, so the walk is unchanged.
__device__ __forceinline__ uint32_t vis_hash(uint32_t x) { x *= 0x9E3779B1u; return x ^ (x >> 15); }

// test-and-set of `node`; lanes of a warp call it together with distinct nodes. use2: tier 1 is frozen (read-only),
// new keys go to tier 2. Returns true when the node was not in the set.
__device__ __forceinline__ bool vis_insert(uint32_t* t1, uint32_t* t2, uint32_t mask2, uint32_t node, bool use2) {
    const uint32_t key = node + 1;
    const uint32_t h = vis_hash(node);
    uint32_t i = h & (kVisSmem - 1);
    if(!use2) {
        for(;;) {
            const uint32_t old = atomicCAS(t1 + i, 0u, key);
            if(old == 0) return true;
            if(old == key) return false;
            i = (i + 1) & (kVisSmem - 1);
        }
    }
    for(;;) { const uint32_t v = t1[i]; if(v == key) return false; if(v == 0) break; i = (i + 1) & (kVisSmem - 1); }
    uint32_t j = (h >> 7) & mask2;
    for(;;) {
        const uint32_t old = atomicCAS(t2 + j, 0u, key);
        if(old == 0) return true;
        if(old == key) return false;
        j = (j + 1) & mask2;
    }
}
// read-only membership (speculation hints)
__device__ __forceinline__ bool vis_contains(const uint32_t* t1, const uint32_t* t2, uint32_t mask2, uint32_t node, bool t2_used) {
    const uint32_t key = node + 1;
    const uint32_t h = vis_hash(node);
    uint32_t i = h & (kVisSmem - 1);
    for(;;) { const uint32_t v = t1[i]; if(v == key) return true; if(v == 0) break; i = (i + 1) & (kVisSmem - 1); }
    if(!t2_used) return false;
    uint32_t j = (h >> 7) & mask2;
    for(;;) { const uint32_t v = __ldcg(t2 + j); if(v == key) return true; if(v == 0) return false; j = (j + 1) & mask2; }
}

// One query per warp. Measured on the 10 M x 768 workload: 96 registers / 5 CTAs per SM walks 4096 queries in 10.7 ms,
// the 72-register build (7 CTAs per SM, a warp for every query) needs 12.4 ms — the extra residency does not pay for the
// serialised loads the tighter register budget forces, because the batch time is set by the slowest walks.
#ifndef TSGPU_KNN_MIN_CTAS
#define TSGPU_KNN_MIN_CTAS 5
#endif
template <int NCH>
__global__ void __launch_bounds__(kKnnThreads, TSGPU_KNN_MIN_CTAS)
hnsw_search_kernel(const __grid_constant__ HnswDev g, const __grid_constant__ KnnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * (kKnnThreads / 32) + warp;
    const uint32_t ef = P.ef > P.k ? P.ef : P.k;           // Typesense fork: effective ef = max(ef, k)
    const uint32_t dim = g.dim;
    const uint32_t dim_pad = (dim + 3) & ~3u;
    // shared per warp: [ef+1] u64 result heap + [kCandSmem] u64 first tier of the candidate heap + [kVisSmem] u32 visited
    // tier 1; then (generic path) the query vectors
    const size_t per_warp = (size_t) (ef + 1) + kCandSmem + kVisSmem / 2;          // in u64 units
    unsigned long long* res = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t) warp * per_warp;
    unsigned long long* cand_s = res + (ef + 1);
    uint32_t* vis1 = reinterpret_cast<uint32_t*>(cand_s + kCandSmem);
    float* qs = reinterpret_cast<float*>(reinterpret_cast<unsigned long long*>(smem_raw) + (size_t) (kKnnThreads / 32) * per_warp) + (size_t) warp * dim_pad;
    uint32_t* vis2 = P.vis2 + (size_t) slot * P.vis2_slots;
    const uint32_t mask2 = P.vis2_slots - 1;
    const uint32_t limit2 = P.vis2_slots - (P.vis2_slots >> 2);
    unsigned long long* cand = P.cand + (size_t) slot * P.cand_cap;
    const uint32_t L0 = 2 * g.M + 1, LU = g.M + 1;
    const uint32_t n_tickets = P.n_order_dev ? __ldcg(P.n_order_dev) : (P.q_order ? P.n_order : P.nq);
    unsigned long long n_dist_acc = 0, n_exp_acc = 0, n_hit_acc = 0, n_t2_acc = 0;

    for(;;) {
        uint32_t qi = 0;
        if(lane == 0) qi = atomicAdd(P.counter, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if(qi >= n_tickets) break;
        if(P.q_order) qi = __ldg(P.q_order + qi);
        if(P.q_skip && P.q_skip[qi]) { if(lane == 0) P.out_n[qi] = 0; continue; }
        const float* qv = P.queries + (size_t) qi * dim;
        QReg<NCH> q;
        if(NCH > 0) {
#pragma unroll
            for(int s = 0; s < NCH; s++) q.v[s] = __ldg(reinterpret_cast<const float4*>(qv + s * 128) + lane);
        } else {
            for(uint32_t e = lane; e < dim; e += 32) qs[e] = qv[e];
            __syncwarp();
        }
        const uint32_t* fbm = P.q_filter_bitmap ? P.q_filter_bitmap[qi] : nullptr;
        const uint32_t* excl = P.q_excl ? P.q_excl[qi] : nullptr;
        const uint32_t n_excl = P.q_n_excl ? P.q_n_excl[qi] : 0;
        if(g.n_nodes == 0 || g.entry_point == kNone) { if(lane == 0) P.out_n[qi] = 0; continue; }
        for(uint32_t i = lane; i < kVisSmem; i += 32) vis1[i] = 0;

        // ---- greedy descent through the upper layers (searchKnn)
        uint32_t cur = g.entry_point;
        float curdist = 1.0f - dot_one<NCH>(q, qs, g.vectors + (size_t) cur * dim, dim, lane);
        n_dist_acc++;
        for(int level = (int) g.max_level; level > 0; level--) {
            bool changed = true;
            while(changed) {
                changed = false;
                const uint32_t* rec = g.links_up + (g.upper_off[cur] + (unsigned long long) (level - 1)) * LU;
                const uint32_t size = __ldg(rec);
                const uint32_t nb = (lane < size) ? __ldg(rec + 1 + lane) : kNone;     // M <= 32
                for(uint32_t i = 0; i < size; i += 2) {
                    const uint32_t c0 = __shfl_sync(0xffffffffu, nb, i);
                    const uint32_t c1 = (i + 1 < size) ? __shfl_sync(0xffffffffu, nb, i + 1) : c0;
                    float d0, d1;
                    dot_two<NCH>(q, qs, g.vectors + (size_t) c0 * dim, g.vectors + (size_t) c1 * dim, dim, lane, d0, d1);
                    d0 = 1.0f - d0; d1 = 1.0f - d1;
                    n_dist_acc += (i + 1 < size) ? 2 : 1;
                    if(d0 < curdist) { curdist = d0; cur = c0; changed = true; }
                    if(i + 1 < size && d1 < curdist) { curdist = d1; cur = c1; changed = true; }
                }
            }
        }

        // ---- best-first search of the base layer (searchBaseLayerST, non-"bare bone" branch)
        uint32_t n_res = 0, n_cand = 0, n_cs = 0, n_cg = 0;     // n_cand = n_cs (shared tier) + n_cg (HBM tier)
        const unsigned long long exp0 = n_exp_acc, dist0 = n_dist_acc;
        uint32_t n_v1 = 0, n_v2 = 0;                             // visited keys per tier (warp-uniform)
        bool overflow = false;                                   // the walk outgrew this slot's scratch (warp-uniform)
        float lowerBound;
        __syncwarp();
        {
            const bool ok = allowed(g, fbm, excl, n_excl, cur);
            if(ok) {
                const float d = curdist;      // same value hnswlib recomputes for the entry point
                n_dist_acc++;
                lowerBound = d;
                if(lane == 0) { heap_push_max(res, n_res, res_key(d, cur)); heap_push_min(cand_s, n_cs, cand_key(d, cur)); }
            } else {
                lowerBound = FLT_MAX;
                if(lane == 0) heap_push_min(cand_s, n_cs, cand_key(FLT_MAX, cur));
            }
            if(lane == 0) vis_insert(vis1, vis2, mask2, cur, false);
            n_v1 = 1;
            n_res = __shfl_sync(0xffffffffu, n_res, 0);
            n_cand = 1;
            __syncwarp();
        }
        uint32_t prev_spec = kNone, prev_nb2 = kNone, prev_size2 = 0;
        for(;;) {
            if(n_cand == 0 || overflow) break;
            // global minimum = the smaller of the two tiers' tops
            unsigned long long top = 0;
            int from_g = 0;
            if(lane == 0) {
                top = n_cs ? cand_s[0] : ~0ull;
                if(n_cg) { const unsigned long long tg = cand[0]; if(tg < top) { top = tg; from_g = 1; } }
            }
            top = __shfl_sync(0xffffffffu, top, 0);
            const float cdist = unord_f32((uint32_t) (top >> 32));
            if(cdist > lowerBound && n_res == ef) break;
            const uint32_t cnode = ~(uint32_t) top;
            if(lane == 0) { if(from_g) heap_pop_min(cand, n_cg); else heap_pop_min(cand_s, n_cs); }
            n_cand--;
            n_exp_acc++;

            const uint32_t* rec = g.links0 + (size_t) cnode * L0;
            // a right guess already holds this node's link row in registers
            const bool reuse = cnode == prev_spec;
            n_hit_acc += reuse ? 1 : 0;
            const uint32_t size = reuse ? prev_size2 : __ldg(rec);
            const uint32_t nb_first = reuse ? prev_nb2 : ((lane + 1 < L0) ? __ldg(rec + 1 + lane) : kNone);
            // Speculation (hints only, no effect on results): the new top of the shared-memory tier is the most likely next
            // expansion. Its link row rides along with this node's, and the vectors of its unvisited neighbours are started
            // towards L2 before this node's distances are computed — so the next expansion finds links and vectors in L2
            // instead of paying two DRAM round trips.
            uint32_t spec = kNone, nb2 = kNone, size2 = 0;
            if(lane == 0 && n_cs) spec = ~(uint32_t) cand_s[0];
            spec = __shfl_sync(0xffffffffu, spec, 0);
            if(spec != kNone) {
                const uint32_t* rec2 = g.links0 + (size_t) spec * L0;
                size2 = __ldg(rec2);
                nb2 = (lane + 1 < L0) ? __ldg(rec2 + 1 + lane) : kNone;
            }
            uint32_t nb = kNone;
            bool fresh = false;
            // 2M <= 32 neighbours handled one per lane; (2M > 32 is processed in chunks of 32)
            for(uint32_t base = 0; base < size; base += 32) {
                const uint32_t j = base + lane;
                nb = (j < size) ? (base == 0 ? nb_first : __ldg(rec + 1 + j)) : kNone;
                const bool use2 = n_v1 + 32 > kVisSmemLimit;          // warp-uniform: tier 1 frozen once it may pass 75 %
                if(use2 && n_v2 + 32 > limit2) { overflow = true; break; }
                fresh = (nb != kNone) && vis_insert(vis1, vis2, mask2, nb, use2);
                uint32_t mask = __ballot_sync(0xffffffffu, fresh);
                if(use2) n_v2 += __popc(mask); else n_v1 += __popc(mask);
                __syncwarp();
                if(base == 0 && lane < size2 && nb2 < g.n_nodes && !vis_contains(vis1, vis2, mask2, nb2, n_v2 != 0)) {
                    const char* vp = reinterpret_cast<const char*>(g.vectors + (size_t) nb2 * dim);
                    for(uint32_t o = 0; o < dim * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(vp + o));
                }
                // every fresh neighbour's vector will be read below, two at a time: start all of them towards L2 now so only
                // the first pair pays the full DRAM (and page-walk) latency
                if(fresh && __popc(mask) > 2) {
                    const char* vp = reinterpret_cast<const char*>(g.vectors + (size_t) nb * dim);
                    for(uint32_t o = 0; o < dim * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(vp + o));
                }
                while(mask) {
                    const int j0 = __ffs(mask) - 1; mask &= mask - 1;
                    int j1 = -1;
                    if(mask) { j1 = __ffs(mask) - 1; mask &= mask - 1; }
                    const uint32_t c0 = __shfl_sync(0xffffffffu, nb, j0);
                    const uint32_t c1 = j1 >= 0 ? __shfl_sync(0xffffffffu, nb, j1) : c0;
                    // filter-functor loads are issued ahead of the vector loads so their latency hides under them
                    const bool ok0 = allowed(g, fbm, excl, n_excl, c0);
                    const bool ok1 = j1 >= 0 ? allowed(g, fbm, excl, n_excl, c1) : false;
                    float d0, d1;
                    dot_two<NCH>(q, qs, g.vectors + (size_t) c0 * dim, g.vectors + (size_t) c1 * dim, dim, lane, d0, d1);
                    d0 = 1.0f - d0; d1 = 1.0f - d1;
                    n_dist_acc += j1 >= 0 ? 2 : 1;
                    // admission replayed in neighbour order (lane 0 owns the heaps)
                    uint32_t ovf = 0;
                    if(lane == 0) {
#pragma unroll
                        for(int t = 0; t < 2; t++) {
                            if(t == 1 && j1 < 0) break;
                            const float d = t ? d1 : d0;
                            const uint32_t c = t ? c1 : c0;
                            const bool ok = t ? ok1 : ok0;
                            if(n_res < ef || lowerBound > d) {
                                {   // the link row of a pushed candidate will be needed when it is expanded: start pulling it into L2
                                    const char* lp = reinterpret_cast<const char*>(g.links0 + (size_t) c * L0);
                                    asm volatile("prefetch.global.L2 [%0];" :: "l"(lp));
                                    asm volatile("prefetch.global.L2 [%0];" :: "l"(lp + 128));
                                }
                                if(n_cs < kCandSmem) heap_push_min(cand_s, n_cs, cand_key(d, c));
                                else if(n_cg < P.cand_cap) heap_push_min(cand, n_cg, cand_key(d, c));
                                else ovf = 1;
                                if(ok) {       // push, then pop while over ef (hnswlib) == replace the maximum when already full
                                    if(n_res < ef) heap_push_max(res, n_res, res_key(d, c));
                                    else heap_replace_max(res, n_res, res_key(d, c));
                                }
                                if(n_res) lowerBound = unord_f32((uint32_t) (res[0] >> 32));
                            }
                        }
                    }
                    lowerBound = __shfl_sync(0xffffffffu, lowerBound, 0);
                    n_res = __shfl_sync(0xffffffffu, n_res, 0);
                    n_cand = __shfl_sync(0xffffffffu, n_cs + n_cg, 0);
                    if(__shfl_sync(0xffffffffu, ovf, 0)) { overflow = true; break; }
                }
                if(overflow) break;
            }
            prev_spec = spec; prev_nb2 = nb2; prev_size2 = size2;
            __syncwarp();
        }

        if(overflow) {
            // hand the query back: the host re-runs it alone with a slot sized for the whole graph (the other queries of the
            // batch are unaffected)
            if(lane == 0) { const uint32_t r = atomicAdd(P.retry_n, 1u); P.retry_list[r] = qi; P.out_n[qi] = 0; }
        } else if(lane == 0) {
            // ---- emit: keep the k closest, closest first (searchKnnCloserFirst)
            while(n_res > P.k) heap_pop_max(res, n_res);
            const uint32_t n = n_res;
            P.out_n[qi] = n;
            for(uint32_t i = n; i-- > 0;) {
                const unsigned long long t = res[0];
                heap_pop_max(res, n_res);
                const uint32_t node = (uint32_t) t;
                P.out_dist[(size_t) qi * P.k + i] = unord_f32((uint32_t) (t >> 32));
                P.out_labels[(size_t) qi * P.k + i] = g.labels ? g.labels[node] : node;
            }
        }
        if(lane == 0 && P.q_work) { P.q_work[2 * qi] = (uint32_t) (n_exp_acc - exp0); P.q_work[2 * qi + 1] = (uint32_t) (n_dist_acc - dist0); }
        __syncwarp();
        // ---- tier 2 goes back to all-zero for the slot's next walk
        if(n_v2) {
            n_t2_acc++;
            uint4* z = reinterpret_cast<uint4*>(vis2);
            for(uint32_t i = lane; i < (P.vis2_slots >> 2); i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
    }
    if(lane == 0) { atomicAdd(P.stats + 0, n_dist_acc); atomicAdd(P.stats + 1, n_exp_acc); atomicAdd(P.stats + 2, n_hit_acc); atomicAdd(P.stats + 3, n_t2_acc); }
}

// ===============================================================================================================
// hnsw_walk_kernel — the same search with its two priority queues held ACROSS THE WARP'S REGISTERS instead of as binary
// heaps that one lane maintains in shared / global memory. Why: profiles/r02b showed the heap kernel executing ~1150 warp
// instructions per expanded node, most of them single-lane heap sifts (dependent shared-memory chains; dependent
// GLOBAL-memory chains once a filtered walk's candidate set has outgrown its shared tier), at ~10 cycles each with 5 warps
// per scheduler — and the batch time is the latency of the longest walks (2048 filtered walks of ~1000 expansions on 2960
// warp slots: no second wave to hide behind), so instructions on one walk's critical path are what the kernel time is made of.
//   result set R   the ef (<= 128) closest allowed nodes as ONE sorted array, element e in lane e / 4, register e % 4
//                  (keys ord(dist) << 32 | id, ascending): admission is a warp-wide compare + shift (~25 instructions, no
//                  memory), lowerBound is a shuffle
//   candidates C   the 128 smallest candidate keys sorted the same way (the "buffer") + two UNSORTED pools in HBM for the
//                  rest: "near" (keys below a pivot pv) and "far" (keys >= pv). Invariant: buffer <= near < pv <= far, so
//                  the buffer's front is the global minimum. A key below the buffer's last goes into the buffer (the
//                  evicted last goes to the near pool, one store); others are appended to the pool their side of pv says.
//                  Only when the buffer runs empty are the pools read: the near pool — split first around a sampled pivot
//                  when it holds more than kNearMax keys; replaced by the far pool when empty — gives up its 128 smallest
//                  through a sorting network (sort each 128-key chunk, merge against the buffer, write the upper half
//                  back): ~30 instructions per popped key. (The first version scanned ONE pool with a serial insert per
//                  qualifying key: 26 % of the kernel's instructions on filtered walks, profiles/r02g.)
// Pop order, admission order and every comparison are those of hnswlib's two std::priority_queues, so results are unchanged
// (same GPU parity tests). Used when max(ef, k) <= 128 and 2M <= 32 (Typesense defaults: ef 10..100s, M 16); the heap kernel
// above stays as the general path.
constexpr unsigned long long kKeyInf = ~0ull;

template <int S> struct WArr { unsigned long long a[S]; };

template <int S>
__device__ __forceinline__ void warr_clear(WArr<S>& w) {
#pragma unroll
    for(int i = 0; i < S; i++) w.a[i] = kKeyInf;
}
// insert k (unique, < +inf) keeping ascending order; returns what fell off the end (+inf while the array is not full)
template <int S>
__device__ __forceinline__ unsigned long long warr_insert(WArr<S>& w, unsigned long long k, uint32_t lane) {
    const unsigned long long last = w.a[S - 1];
    unsigned long long x = __shfl_up_sync(0xffffffffu, last, 1);          // the element just before my block
    if(lane == 0) x = 0;
    const unsigned long long evicted = __shfl_sync(0xffffffffu, last, 31);
    if(k < last) {                                                        // my block changes
        if(k < x) {                                                       // k landed in an earlier lane: shift right, x comes in
#pragma unroll
            for(int i = S - 1; i > 0; i--) w.a[i] = w.a[i - 1];
            w.a[0] = x;
        } else {                                                          // k lands here, my last falls to the next lane
            unsigned long long carry = k;
#pragma unroll
            for(int i = 0; i < S; i++) { const unsigned long long t = w.a[i]; const bool sw = carry < t; w.a[i] = sw ? carry : t; carry = sw ? t : carry; }
        }
    }
    return evicted;
}
template <int S>
__device__ __forceinline__ unsigned long long warr_front(const WArr<S>& w) { return __shfl_sync(0xffffffffu, w.a[0], 0); }
template <int S>
__device__ __forceinline__ void warr_pop_front(WArr<S>& w, uint32_t lane) {
    unsigned long long nx = __shfl_down_sync(0xffffffffu, w.a[0], 1);
    if(lane == 31) nx = kKeyInf;
#pragma unroll
    for(int i = 0; i < S - 1; i++) w.a[i] = w.a[i + 1];
    w.a[S - 1] = nx;
}
template <int S>
__device__ __forceinline__ unsigned long long warr_get(const WArr<S>& w, uint32_t e) {       // e warp-uniform
    const uint32_t r = e % S;
    unsigned long long v = w.a[0];
    if(S > 1 && r == 1) v = w.a[S > 1 ? 1 : 0];
    if(S > 2 && r == 2) v = w.a[S > 2 ? 2 : 0];
    if(S > 3 && r == 3) v = w.a[S > 3 ? 3 : 0];
    return __shfl_sync(0xffffffffu, v, e / S);
}
template <int S>
__device__ __forceinline__ void warr_set_inf(WArr<S>& w, uint32_t e, uint32_t lane) {
    const bool me = lane == e / S;
    const uint32_t r = e % S;
    if(me && r == 0) w.a[0] = kKeyInf;
    if(S > 1 && me && r == 1) w.a[S > 1 ? 1 : 0] = kKeyInf;
    if(S > 2 && me && r == 2) w.a[S > 2 ? 2 : 0] = kKeyInf;
    if(S > 3 && me && r == 3) w.a[S > 3 ? 3 : 0] = kKeyInf;
}
static_assert(true, "WArr accessors are written out for S <= 4");

// ---- 128 keys (4 per lane) as a sorting network: bitonic stages over the element index e = lane * 4 + register.
// Strides below 4 stay inside a lane; larger ones are one xor-shuffle per register. Used by the candidate buffer's refill,
// which takes the 128 smallest of a few hundred pooled keys by sort + merge instead of one serial insert per key.
__device__ __forceinline__ void key_cx(unsigned long long& a, unsigned long long& b, bool asc) {      // afterwards a <= b iff asc
    const bool sw = (a > b) == asc;
    const unsigned long long t = a;
    a = sw ? b : a; b = sw ? t : b;
}
template <uint32_t K, uint32_t J>
__device__ __forceinline__ void warr_stage(WArr<4>& w, uint32_t lane) {
    if constexpr (J >= 4) {
        constexpr uint32_t LJ = J >> 2;
        const bool lower = (lane & LJ) == 0;
        const bool asc = ((lane << 2) & K) == 0;             // K >= 8 here: the direction is a lane bit
        const bool take_min = lower == asc;
#pragma unroll
        for(int r = 0; r < 4; r++) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, w.a[r], LJ);
            const bool less = w.a[r] < o;
            w.a[r] = (less == take_min) ? w.a[r] : o;
        }
    } else if constexpr (J == 2) {
        const bool asc = ((lane << 2) & K) == 0;             // K >= 4
        key_cx(w.a[0], w.a[2], asc); key_cx(w.a[1], w.a[3], asc);
    } else {
        if constexpr (K == 2) { key_cx(w.a[0], w.a[1], true); key_cx(w.a[2], w.a[3], false); }
        else { const bool asc = ((lane << 2) & K) == 0; key_cx(w.a[0], w.a[1], asc); key_cx(w.a[2], w.a[3], asc); }
    }
}
template <uint32_t K, uint32_t J>
__device__ __forceinline__ void warr_merge_from(WArr<4>& w, uint32_t lane) {       // stages (K, J), (K, J/2), ... (K, 1)
    warr_stage<K, J>(w, lane);
    if constexpr (J > 1) warr_merge_from<K, J / 2>(w, lane);
}
template <uint32_t K>
__device__ __forceinline__ void warr_sort_from(WArr<4>& w, uint32_t lane) {        // blocks of K, then 2K, ... 128
    warr_merge_from<K, K / 2>(w, lane);
    if constexpr (K < 128) warr_sort_from<K * 2>(w, lane);
}
// any order -> ascending (element e in lane e / 4, register e % 4: the layout of warr_insert / warr_front)
__device__ __forceinline__ void warr_sort128(WArr<4>& w, uint32_t lane) { warr_sort_from<2>(w, lane); }
// a bitonic sequence -> ascending
__device__ __forceinline__ void warr_bmerge128(WArr<4>& w, uint32_t lane) { warr_merge_from<128, 64>(w, lane); }
// b, c ascending -> b = the 128 smallest of both, ascending; c = the 128 others as a bitonic sequence
__device__ __forceinline__ void warr_keep_low(WArr<4>& b, WArr<4>& c, uint32_t lane) {
    unsigned long long o[4];
#pragma unroll
    for(int r = 0; r < 4; r++) o[r] = __shfl_sync(0xffffffffu, c.a[3 - r], 31 - lane);      // c reversed
#pragma unroll
    for(int r = 0; r < 4; r++) {
        const bool less = b.a[r] < o[r];
        c.a[r] = less ? o[r] : b.a[r];
        b.a[r] = less ? b.a[r] : o[r];
    }
    warr_bmerge128(b, lane);
}

constexpr int kResPerLane = 4;           // R: 128 entries
constexpr int kBufPerLane = 4;           // C buffer: 128 entries
constexpr uint32_t kBufCap = 32 * kBufPerLane;
constexpr uint32_t kNearMax = 1024;       // a near pool above this is split before the buffer refills from it ...
constexpr uint32_t kNearTarget = 512;     // ... with a pivot aimed at leaving this many keys
constexpr int kWalkWarps = 2;            // walks (warps) per CTA
constexpr int kStageRows = 4;            // default: neighbour rows in flight per walk (shared-memory ring filled by cp.async.bulk)
constexpr uint32_t kVisCache = 2048;     // default: direct-mapped cache of recently visited nodes per walk (shared memory)

// ---- mbarrier + bulk async copy (TMA engine, 1-D): rows land in shared memory without passing through registers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Visited set of the register-queue kernel: the authoritative set is an open-addressing table in HBM (one per warp slot,
// all-zero between walks); in front of it a direct-mapped cache of recently visited nodes in shared memory answers most of
// the "already visited" tests (9 of 10 neighbours of an expanded node have been seen, almost always recently) with one
// shared-memory load. Only cache misses — the fresh neighbours plus a few evicted ones — go to the table: one compare-
// and-swap each, in parallel across the lanes. Exact by construction (the table decides), any size of walk.
__device__ __forceinline__ bool vis_test_and_set(uint32_t* cache, uint32_t cmask, uint32_t* table, uint32_t mask2, uint32_t node, bool* went_to_table = nullptr) {
    const uint32_t key = node + 1;
    const uint32_t h = vis_hash(node);
    uint32_t* c = cache + (h & cmask);
    if(*c == key) return false;
    if(went_to_table) *went_to_table = true;
    bool fresh = false;
    uint32_t j = (h >> 11) & mask2;
    for(;;) {
        const uint32_t old = atomicCAS(table + j, 0u, key);
        if(old == 0) { fresh = true; break; }
        if(old == key) break;
        j = (j + 1) & mask2;
    }
    *c = key;
    return fresh;
}

#ifndef TSGPU_WALK_MIN_CTAS
#define TSGPU_WALK_MIN_CTAS 5
#endif
// NCH = dim / 128 (dim a multiple of 128: rows are staged by bulk copies of dim * 4 bytes); other dimensions take the heap kernel.
template <int NCH, int RS>
__global__ void __launch_bounds__(32 * kWalkWarps, RS == 2 ? 2 * TSGPU_WALK_MIN_CTAS : (RS == 8 ? 2 : TSGPU_WALK_MIN_CTAS))
hnsw_walk_kernel(const __grid_constant__ HnswDev g, const __grid_constant__ KnnParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * kWalkWarps + warp;
    const uint32_t ef = P.ef > P.k ? P.ef : P.k;           // <= 128 (host checks)
    constexpr uint32_t dim = NCH * 128, row_bytes = dim * 4;
    // shared per CTA: [kWalkWarps] mbarriers (128 B), then per warp: [RS] rows, [kVisCache] u32
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw);
    float* ring = reinterpret_cast<float*>(smem_raw + 128) + (size_t) warp * RS * dim;
    const uint32_t n_cache = P.vis_cache, cmask = n_cache - 1;          // power of two
    uint32_t* cache = reinterpret_cast<uint32_t*>(smem_raw + 128 + (size_t) kWalkWarps * RS * row_bytes) + (size_t) warp * n_cache;
    const uint32_t bar = smem_u32(bars + warp), ring_s = smem_u32(ring);
    uint32_t parity = 0;
    if(lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t* vis2 = P.vis2 + (size_t) slot * P.vis2_slots;
    const uint32_t mask2 = P.vis2_slots - 1;
    const uint32_t limit2 = P.vis2_slots - (P.vis2_slots >> 2);
    // candidate pools of this slot: two halves, "near" (keys below the pivot) and "far" (the rest); they trade places
    unsigned long long* const pool0 = P.cand + (size_t) slot * P.cand_cap;
    const uint32_t pool_half = P.cand_cap >> 1;
    const uint32_t L0 = 2 * g.M + 1, LU = g.M + 1;
    const uint32_t n_tickets = P.n_order_dev ? __ldcg(P.n_order_dev) : (P.q_order ? P.n_order : P.nq);
    unsigned long long n_dist_acc = 0, n_exp_acc = 0, n_hit_acc = 0, n_t2_acc = 0, n_tab_acc = 0;
    float* const qs = nullptr;

    for(;;) {
        uint32_t qi = 0;
        if(lane == 0) qi = atomicAdd(P.counter, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if(qi >= n_tickets) break;
        if(P.q_order) qi = __ldg(P.q_order + qi);
        if(P.q_skip && P.q_skip[qi]) { if(lane == 0) P.out_n[qi] = 0; continue; }
        const float* qv = P.queries + (size_t) qi * dim;
        QReg<NCH> q;
#pragma unroll
        for(int s = 0; s < NCH; s++) q.v[s] = __ldg(reinterpret_cast<const float4*>(qv + s * 128) + lane);
        const uint32_t* fbm = P.q_filter_bitmap ? P.q_filter_bitmap[qi] : nullptr;
        const uint32_t* excl = P.q_excl ? P.q_excl[qi] : nullptr;
        const uint32_t n_excl = P.q_n_excl ? P.q_n_excl[qi] : 0;
        if(g.n_nodes == 0 || g.entry_point == kNone) { if(lane == 0) P.out_n[qi] = 0; continue; }
        for(uint32_t i = lane; i < n_cache; i += 32) cache[i] = 0;
        const unsigned long long exp0 = n_exp_acc, dist0 = n_dist_acc;

        // ---- greedy descent through the upper layers (searchKnn)
        uint32_t cur = g.entry_point;
        float curdist = 1.0f - dot_one<NCH>(q, qs, g.vectors + (size_t) cur * dim, dim, lane);
        n_dist_acc++;
        for(int level = (int) g.max_level; level > 0; level--) {
            bool changed = true;
            while(changed) {
                changed = false;
                const uint32_t* rec = g.links_up + (g.upper_off[cur] + (unsigned long long) (level - 1)) * LU;
                const uint32_t size = __ldg(rec);
                const uint32_t nb = (lane < size) ? __ldg(rec + 1 + lane) : kNone;     // M <= 32
                for(uint32_t i = 0; i < size; i += 2) {
                    const uint32_t c0 = __shfl_sync(0xffffffffu, nb, i);
                    const uint32_t c1 = (i + 1 < size) ? __shfl_sync(0xffffffffu, nb, i + 1) : c0;
                    float d0, d1;
                    dot_two<NCH>(q, qs, g.vectors + (size_t) c0 * dim, g.vectors + (size_t) c1 * dim, dim, lane, d0, d1);
                    d0 = 1.0f - d0; d1 = 1.0f - d1;
                    n_dist_acc += (i + 1 < size) ? 2 : 1;
                    if(d0 < curdist) { curdist = d0; cur = c0; changed = true; }
                    if(i + 1 < size && d1 < curdist) { curdist = d1; cur = c1; changed = true; }
                }
            }
        }

        // ---- best-first search of the base layer (searchBaseLayerST, non-"bare bone" branch)
        WArr<kResPerLane> R; warr_clear(R);
        WArr<kBufPerLane> B; warr_clear(B);
        uint32_t n_res = 0, n_b = 0;                            // result count, buffered candidates (warp-uniform, like all below)
        unsigned long long* nearp = pool0;                       // pooled candidates: keys < pv ...
        unsigned long long* farp = pool0 + pool_half;            // ... and keys >= pv, both unsorted
        uint32_t n_near = 0, n_far = 0;
        unsigned long long pv = kKeyInf;
        unsigned long long b_last = 0;                           // the buffer's largest key (0 when empty)
        uint32_t n_vis = 0;                                      // keys in the visited table
        bool overflow = false;
        float lowerBound;
        __syncwarp();
        {
            const bool ok = allowed(g, fbm, excl, n_excl, cur);
            if(ok) {
                const float d = curdist;      // same value hnswlib recomputes for the entry point
                n_dist_acc++;
                lowerBound = d;
                warr_insert(R, res_key(d, cur), lane); n_res = 1;
                warr_insert(B, cand_key(d, cur), lane); n_b = 1; b_last = cand_key(d, cur);
            } else {
                lowerBound = FLT_MAX;
                warr_insert(B, cand_key(FLT_MAX, cur), lane); n_b = 1; b_last = cand_key(FLT_MAX, cur);
            }
            if(lane == 0) vis_test_and_set(cache, cmask, vis2, mask2, cur);
            n_vis = 1;
            __syncwarp();
        }
        uint32_t prev_spec = kNone, prev_nb2 = kNone, prev_size2 = 0;
        for(;;) {
            if(overflow) break;
            if(n_b == 0) {
                if(n_near == 0) {
                    if(n_far == 0) break;
                    unsigned long long* t = nearp; nearp = farp; farp = t;           // the far pool becomes the near one
                    n_near = n_far; n_far = 0; pv = kKeyInf;
                }
                __syncwarp();
                if(n_near > kNearMax) {
                    // ---- split: a pivot from a 32-key sample (the rank-q sample, q chosen for ~kNearTarget keys below it);
                    // keys below stay (compacted in place), the others move to the far pool. Any pivot is correct — the
                    // invariant is only near < pv <= far — a lucky or unlucky one just changes how soon the next split comes.
                    const uint32_t n = n_near;
                    const unsigned long long sk = nearp[(unsigned long long) lane * n >> 5];
                    uint32_t rank = 0;
#pragma unroll 8
                    for(int o = 0; o < 32; o++) rank += __shfl_sync(0xffffffffu, sk, o) < sk ? 1u : 0u;
                    uint32_t q = (32u * kNearTarget + (n >> 1)) / n;
                    q = q < 1 ? 1 : (q > 16 ? 16 : q);
                    const unsigned long long pivot = __shfl_sync(0xffffffffu, sk, __ffs(__ballot_sync(0xffffffffu, rank == q)) - 1);
                    uint32_t w = 0;
                    const uint32_t lt = (1u << lane) - 1u;
                    for(uint32_t base = 0; base < n; base += 32) {
                        const bool valid = base + lane < n;
                        const unsigned long long k = valid ? nearp[base + lane] : kKeyInf;
                        const uint32_t mb = __ballot_sync(0xffffffffu, valid && k < pivot);
                        const uint32_t ma = __ballot_sync(0xffffffffu, valid && k >= pivot);
                        if(n_far + __popc(ma) > pool_half) { overflow = true; break; }
                        if(valid) {
                            if(k < pivot) nearp[w + __popc(mb & lt)] = k;
                            else farp[n_far + __popc(ma & lt)] = k;
                        }
                        w += __popc(mb); n_far += __popc(ma);
                    }
                    if(overflow) break;
                    n_near = w; pv = pivot;
                    __syncwarp();
                }
                // ---- refill: the buffer takes the 128 smallest near keys. The pool's tail is sorted into the buffer, then
                // every 128-key chunk of the rest is sorted and merged against it: the lower half stays in registers, the
                // upper half goes back where the chunk came from (so the pool never holds gaps). ~1 K instructions per chunk,
                // where one serial insert per qualifying key cost 30-40 K per refill on a filtered walk's pool.
                {
                    uint32_t n = n_near;
                    const uint32_t seed = n < kBufCap ? n : kBufCap;
                    n -= seed;
#pragma unroll
                    for(int r = 0; r < 4; r++) { const uint32_t e = r * 32 + lane; B.a[r] = e < seed ? nearp[n + e] : kKeyInf; }
                    WArr<4> C; warr_clear(C);
                    if(n) {
#pragma unroll
                        for(int r = 0; r < 4; r++) { const uint32_t e = r * 32 + lane; C.a[r] = e < n ? nearp[e] : kKeyInf; }
                    }
                    warr_sort128(B, lane);
                    n_b = seed;
                    for(uint32_t base = 0; base < n; base += 128) {
                        const uint32_t m = n - base < 128 ? n - base : 128;
                        WArr<4> N;                                        // the next chunk's loads ride under this chunk's sort
                        const uint32_t nb_ = base + 128;
#pragma unroll
                        for(int r = 0; r < 4; r++) { const uint32_t e = nb_ + r * 32 + lane; N.a[r] = e < n ? nearp[e] : kKeyInf; }
                        warr_sort128(C, lane);
                        warr_keep_low(B, C, lane);
                        if(m == 128) {
#pragma unroll
                            for(int r = 0; r < 4; r++) nearp[base + r * 32 + lane] = C.a[r];
                        } else {                                          // a partial chunk: its m real keys first
                            warr_bmerge128(C, lane);
#pragma unroll
                            for(int r = 0; r < 4; r++) { const uint32_t e = lane * 4 + r; if(e < m) nearp[base + e] = C.a[r]; }
                        }
                        C = N;
                    }
                    n_near = n;
                    b_last = warr_get(B, n_b - 1);
                }
                __syncwarp();
            }
            const unsigned long long top = warr_front(B);
            const float cdist = unord_f32((uint32_t) (top >> 32));
            if(cdist > lowerBound && n_res == ef) break;
            const uint32_t cnode = ~(uint32_t) top;
            warr_pop_front(B, lane); n_b--;
            if(n_b == 0) b_last = 0;
            n_exp_acc++;

            const uint32_t* rec = g.links0 + (size_t) cnode * L0;
            // a right guess already holds this node's link row in registers
            const bool reuse = cnode == prev_spec;
            n_hit_acc += reuse ? 1 : 0;
            const uint32_t size = reuse ? prev_size2 : __ldg(rec);
            const uint32_t nb = reuse ? prev_nb2 : ((lane + 1 < L0) ? __ldg(rec + 1 + lane) : kNone);
            // the buffer's new front is the most likely next expansion: its link row rides along (a hint, no effect on results)
            uint32_t spec = kNone, nb2 = kNone, size2 = 0;
            if(n_b) spec = ~(uint32_t) warr_front(B);
            if(spec != kNone) {
                const uint32_t* rec2 = g.links0 + (size_t) spec * L0;
                size2 = __ldg(rec2);
                nb2 = (lane + 1 < L0) ? __ldg(rec2 + 1 + lane) : kNone;
            }
            if(n_vis + 32 > limit2) { overflow = true; break; }
            // the filter functor of every neighbour at once (one bitmap word per lane): the load is issued ahead of the visited
            // test, whose table probe is the longer wait, and its result is only looked at afterwards, for the fresh ones
            uint32_t fw = 0xffffffffu, dw = 0u;                  // only the LOADS go first; the bits are looked at after the table's answer
            if(lane < size && !g.labels) {
                if(fbm) fw = __ldg(fbm + (nb >> 5));
                if(g.deleted) dw = __ldg(g.deleted + (nb >> 5));
            }
            bool to_table = false;
            const bool fresh = (lane < size) && vis_test_and_set(cache, cmask, vis2, mask2, nb, &to_table);
            uint32_t mask = __ballot_sync(0xffffffffu, fresh);
            n_vis += __popc(mask);
            n_tab_acc += __popc(__ballot_sync(0xffffffffu, to_table));
            bool ok_mine = false;
            if(fresh) {
                if(g.labels) ok_mine = allowed(g, fbm, excl, n_excl, nb);
                else ok_mine = ((fw >> (nb & 31)) & 1u) && !((dw >> (nb & 31)) & 1u) && !(n_excl && is_excluded(excl, n_excl, nb));
            }
            const uint32_t ok_mask = __ballot_sync(0xffffffffu, ok_mine);
            n_dist_acc += __popc(mask);
            // Hints for the likely next expansion (no effect on results): the neighbours of the buffer's front that the visited
            // cache does not recognise will most probably be probed in the table and have their rows fetched one iteration
            // from now — start the table line, the row (one bulk L2 prefetch per row) and the filter word towards L2 while this
            // expansion's rows are in flight. P.walk_prefetch: bit 0 table line, bit 1 row, bit 2 filter word.
            bool hinted = P.walk_prefetch == 0;
            auto hint_next = [&]() {
                hinted = true;
                if(spec != kNone && lane < size2) {
                    const uint32_t h2 = vis_hash(nb2);
                    if(cache[h2 & cmask] != nb2 + 1) {
                        if(P.walk_prefetch & 1u) asm volatile("prefetch.global.L2 [%0];" :: "l"(vis2 + ((h2 >> 11) & mask2)));
                        if(P.walk_prefetch & 2u) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(g.vectors + (size_t) nb2 * dim), "r"(row_bytes) : "memory");
                        if((P.walk_prefetch & 4u) && fbm && !g.labels) asm volatile("prefetch.global.L2 [%0];" :: "l"(fbm + (nb2 >> 5)));
                    }
                }
            };
            if(!mask && !hinted) hint_next();
            while(mask) {
                // ---- up to RS fresh neighbours at a time: their rows are copied into the ring by the bulk-copy engine,
                // all in flight together, completion counted in bytes on the warp's mbarrier
                uint32_t grp = 0, cnt = 0;
                while(mask && cnt < (uint32_t) RS) { grp |= mask & (0u - mask); mask &= mask - 1; cnt++; }
                __syncwarp();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the ring's previous readers are done
                if(lane == 0) mbar_expect_tx(bar, cnt * row_bytes);
                __syncwarp();
                if((grp >> lane) & 1u) {
                    const uint32_t rs = __popc(grp & ((1u << lane) - 1u));
                    bulk_g2s(ring_s + rs * row_bytes, g.vectors + (size_t) nb * dim, row_bytes, bar);
                }
                if(!hinted) hint_next();
                {
                    uint32_t spins = 0;
                    while(!mbar_try_wait(bar, parity)) { if(++spins > (1u << 28)) { __trap(); } }
                    parity ^= 1u;
                }
                float dres[RS];
                {
                    float acc[RS][4];
#pragma unroll
                    for(int r = 0; r < RS; r++) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
#pragma unroll
                    for(int sgm = 0; sgm < NCH; sgm++) {
#pragma unroll
                        for(int r = 0; r < RS; r++) {
                            if((uint32_t) r < cnt) {
                                const float4 x = reinterpret_cast<const float4*>(ring + (size_t) r * dim + sgm * 128)[lane];
                                acc[r][0] = __fmaf_rn(q.v[sgm].x, x.x, acc[r][0]); acc[r][1] = __fmaf_rn(q.v[sgm].y, x.y, acc[r][1]);
                                acc[r][2] = __fmaf_rn(q.v[sgm].z, x.z, acc[r][2]); acc[r][3] = __fmaf_rn(q.v[sgm].w, x.w, acc[r][3]);
                            }
                        }
                    }
#pragma unroll
                    for(int r = 0; r < RS; r++) dres[r] = 1.0f - warp_tree(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                }
                // ---- admission in neighbour order (every value below is warp-uniform)
#pragma unroll
                for(int r = 0; r < RS; r++) {
                    if((uint32_t) r < cnt && !overflow) {
                        const int src = __ffs(grp) - 1; grp &= grp - 1;
                        const float d = dres[r];
                        const uint32_t c = __shfl_sync(0xffffffffu, nb, src);
                        const bool ok = (ok_mask >> src) & 1u;
                        if(n_res < ef || lowerBound > d) {
                            if(lane == 0) {   // the link row of a pushed candidate will be needed when it is expanded: start pulling it into L2
                                const char* lp = reinterpret_cast<const char*>(g.links0 + (size_t) c * L0);
                                asm volatile("prefetch.global.L2 [%0];" :: "l"(lp));
                                asm volatile("prefetch.global.L2 [%0];" :: "l"(lp + 128));
                            }
                            const unsigned long long ck = cand_key(d, c);
                            // buffer unless a pooled key might be smaller: below the buffer's last, or nothing pooled and room left
                            if(ck >= pv) { if(n_far < pool_half) { if(lane == 0) farp[n_far] = ck; n_far++; } else overflow = true; }
                            else {
                                if((n_near == 0 && n_b < kBufCap) || ck < b_last) {
                                    const unsigned long long ev = warr_insert(B, ck, lane);
                                    if(n_b < kBufCap) { n_b++; b_last = ck > b_last ? ck : b_last; }
                                    else {
                                        b_last = __shfl_sync(0xffffffffu, B.a[kBufPerLane - 1], 31);
                                        if(n_near < pool_half) { if(lane == 0) nearp[n_near] = ev; n_near++; } else overflow = true;
                                    }
                                } else { if(n_near < pool_half) { if(lane == 0) nearp[n_near] = ck; n_near++; } else overflow = true; }
                            }
                            if(ok) {       // push, then pop while over ef (hnswlib) == the insert drops the last when already full
                                warr_insert(R, res_key(d, c), lane);
                                if(n_res < ef) n_res++; else warr_set_inf(R, ef, lane);
                                lowerBound = unord_f32((uint32_t) (warr_get(R, n_res - 1) >> 32));
                            }
                        }
                    }
                }
                if(overflow) break;
            }
            prev_spec = spec; prev_nb2 = nb2; prev_size2 = size2;
            __syncwarp();
        }

        if(overflow) {
            if(lane == 0) { const uint32_t r = atomicAdd(P.retry_n, 1u); P.retry_list[r] = qi; P.out_n[qi] = 0; }
        } else {
            // ---- emit: the k closest, closest first (searchKnnCloserFirst): R is already in that order
            const uint32_t n = n_res < P.k ? n_res : P.k;
            if(lane == 0) P.out_n[qi] = n;
#pragma unroll
            for(int i = 0; i < kResPerLane; i++) {
                const uint32_t e = lane * kResPerLane + i;
                if(e < n) {
                    const unsigned long long t = R.a[i];
                    const uint32_t node = (uint32_t) t;
                    P.out_dist[(size_t) qi * P.k + e] = unord_f32((uint32_t) (t >> 32));
                    P.out_labels[(size_t) qi * P.k + e] = g.labels ? g.labels[node] : node;
                }
            }
        }
        if(lane == 0 && P.q_work) { P.q_work[2 * qi] = (uint32_t) (n_exp_acc - exp0); P.q_work[2 * qi + 1] = (uint32_t) (n_dist_acc - dist0); }
        __syncwarp();
        // ---- the table goes back to all-zero for the slot's next walk
        {
            if(n_vis > kVisSmemLimit) n_t2_acc++;
            uint4* z = reinterpret_cast<uint4*>(vis2);
            for(uint32_t i = lane; i < (P.vis2_slots >> 2); i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
    }
    if(lane == 0) { atomicAdd(P.stats + 0, n_dist_acc); atomicAdd(P.stats + 1, n_exp_acc); atomicAdd(P.stats + 2, n_hit_acc); atomicAdd(P.stats + 3, n_t2_acc); atomicAdd(P.stats + 5, n_tab_acc); }
}

// process_results_bruteforce: one warp per (query, id)
struct FlatParams {
    const float* queries;          // [nq*dim]
    const uint32_t* ids;           // concatenated candidate ids
    const unsigned long long* q_off; // [nq+1] ranges into ids
    uint32_t nq;
    float* out_dist;               // aligned with ids
    unsigned long long* stats;
    const uint8_t* q_skip;         // [nq] 1 = this query's range is answered by the tensor-core scan (flat_tc.cu); nullptr = none
};

template <int NCH>
__global__ void __launch_bounds__(256)
flat_distance_kernel(const __grid_constant__ HnswDev g, const __grid_constant__ FlatParams P, unsigned long long total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dim = g.dim, dim_pad = (dim + 3) & ~3u;
    float* qs = reinterpret_cast<float*>(smem_raw) + (size_t) warp * dim_pad;
    const unsigned long long wid = (unsigned long long) blockIdx.x * (blockDim.x >> 5) + warp;
    const unsigned long long nw = (unsigned long long) gridDim.x * (blockDim.x >> 5);
    uint32_t cur_q = kNone;
    QReg<NCH> q;
    for(unsigned long long i = wid; i < total; i += nw) {
        // locate the query of element i (binary search over q_off)
        uint32_t lo = 0, hi = P.nq;
        while(lo + 1 < hi) { const uint32_t mid = (lo + hi) >> 1; if(P.q_off[mid] <= i) lo = mid; else hi = mid; }
        if(P.q_skip && P.q_skip[lo]) continue;
        if(lo != cur_q) {
            cur_q = lo;
            const float* qv = P.queries + (size_t) cur_q * dim;
            if(NCH > 0) {
#pragma unroll
                for(int s = 0; s < NCH; s++) q.v[s] = __ldg(reinterpret_cast<const float4*>(qv + s * 128) + lane);
            } else {
                __syncwarp();
                for(uint32_t e = lane; e < dim; e += 32) qs[e] = qv[e];
                __syncwarp();
            }
        }
        const uint32_t id = P.ids[i];
        float d = 0.f;
        if(id < g.n_nodes) d = 1.0f - dot_one<NCH>(q, qs, g.vectors + (size_t) id * dim, dim, lane);
        if(lane == 0) P.out_dist[i] = d;
    }
}


// Load-time check of an exported graph (the search kernels follow links and offsets without bounds tests): first violation
// class wins. 1/2: level-0 count / id, 3: levels vs upper_off, 4/5: upper count / id, 6: level > max_level, 7: entry point level.
__global__ void __launch_bounds__(256)
hnsw_validate_kernel(const HnswDev g, unsigned long long n_up, uint32_t* __restrict__ bad) {
    const uint32_t L0 = 2 * g.M + 1, LU = g.M + 1;
    for(size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < g.n_nodes; i += (size_t) gridDim.x * blockDim.x) {
        uint32_t err = 0;
        const uint32_t* l0 = g.links0 + i * L0;
        const uint32_t c0 = l0[0];
        if(c0 > 2 * g.M) err = 1;
        else for(uint32_t j = 0; j < c0; j++) if(l0[1 + j] >= g.n_nodes) { err = 2; break; }
        const unsigned long long u0 = g.upper_off[i], u1 = g.upper_off[i + 1];
        const uint32_t lv = g.levels[i];
        if(!err && (u1 < u0 || u1 - u0 != lv || u1 > n_up)) err = 3;
        if(!err && lv > g.max_level) err = 6;
        if(!err && i == g.entry_point && lv != g.max_level) err = 7;
        if(!err) for(unsigned long long r = u0; r < u1 && !err; r++) {
            const uint32_t* lu = g.links_up + r * LU;
            const uint32_t c = lu[0];
            if(c > g.M) { err = 4; break; }
            for(uint32_t j = 0; j < c; j++) if(lu[1 + j] >= g.n_nodes) { err = 5; break; }
        }
        if(err) atomicCAS(bad, 0u, err);
    }
}


__global__ void __launch_bounds__(256)
mark_deleted_kernel(const uint32_t* __restrict__ labels, size_t n, uint32_t* __restrict__ bitmap, bool set) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint32_t l = labels[i];
    if(set) atomicOr(bitmap + (l >> 5), 1u << (l & 31)); else atomicAnd(bitmap + (l >> 5), ~(1u << (l & 31)));
}

}  // namespace tsv
