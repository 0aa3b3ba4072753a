// HNSW construction on sm_100a: hnswlib's addPoint as batched rounds (vector half).
//
// Replaces (reference file:line):
//   Index::batch_memory_index -> vecdex->addPoint(data, seq_id, true)         src/index.cpp:1003-1054 (4 threads, :1009)
//   hnswlib HierarchicalNSW::addPoint / searchBaseLayer / getNeighborsByHeuristic2 / mutuallyConnectNewElement
//       (hnswlib is a build-time download of the reference; the published algorithm, restated in
//        oracle/ts_oracle_vec.inc:HnswBuilder, is what is implemented)
//
// Execution model. Levels are drawn on the host exactly as hnswlib draws them (same engine, same seed), so the whole
// insertion schedule is known up front: nodes are inserted in ROUNDS of consecutive ids; a round's nodes search the graph
// as it stood when the round began (entry point and top level are constants of the round; a node that raises the top level
// is always the last of its round) — the same relaxation hnswlib's own multi-threaded build makes (the reference inserts
// with 4 threads), except that here the result is deterministic. A round of ONE node is hnswlib's sequential insertion,
// and with max_batch = 1 the graph is bit-identical to the oracle's single-threaded build (tests/test_hnsw_build.py).
//   insert_search_kernel   one warp per new node: greedy descent, then per level searchBaseLayer (ef_construction) +
//                          getNeighborsByHeuristic2(M) -> the node's own link row; shared-memory heaps and visited set as in
//                          knn_kernels.cuh
//   rev_count / rev_scan / rev_fill   the reverse links (neighbour <- new node) of a round grouped by the row they modify
//   rev_apply_kernel       one warp per modified row, its new links applied in ascending node order: append, or — when the
//                          row is full — re-select with the heuristic among the row + the new node, as hnswlib does
// All distances use the W128 summation order of knn_kernels.cuh, so the oracle takes the same branches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "knn_kernels.cuh"

namespace tsb {

using namespace tsv;

struct BuildDev {
    uint32_t n, dim, M, efc;
    const float* vectors;
    const uint8_t* levels;
    uint32_t* links0;                      // [n*(2M+1)]
    const unsigned long long* upper_off;   // [n+1]
    uint32_t* links_up;                    // [R*(M+1)]
};

struct RoundParams {
    uint32_t first, count;                 // nodes [first, first+count) are inserted this round
    uint32_t entry, max_level;             // graph state when the round begins
    uint32_t* ticket;                      // zeroed before the round
    uint32_t* vis2; uint32_t vis2_slots;   // per warp slot (tier 2 of the visited set), all-zero between walks
    unsigned long long* cand; uint32_t cand_cap;
    int* error;                            // 1: a walk outgrew its slot (cannot happen when the slots are sized for n)
    unsigned long long* stats;             // [0] distance evaluations of the searches, [1] expansions, [2] heuristic dots
};

__device__ __forceinline__ uint32_t* row_ptr(const BuildDev& g, uint32_t node, uint32_t level) {
    return level == 0 ? g.links0 + (size_t) node * (2 * g.M + 1)
                      : g.links_up + (g.upper_off[node] + (unsigned long long) (level - 1)) * (g.M + 1);
}

template <int NCH>
__device__ __forceinline__ void load_q(QReg<NCH>& q, float* qs, const float* v, uint32_t dim, uint32_t lane) {
    if(NCH > 0) {
#pragma unroll
        for(int s = 0; s < NCH; s++) q.v[s] = __ldg(reinterpret_cast<const float4*>(v + s * 128) + lane);
    } else {
        __syncwarp();
        for(uint32_t e = lane; e < dim; e += 32) qs[e] = v[e];
        __syncwarp();
    }
}

// getNeighborsByHeuristic2 over keys[0..m) (ascending (distance, larger id first): the pop order of hnswlib's
// queue_closest), distances to the query encoded in the keys. Leaves the chosen (<= Mlim) keys in sel[0..n_sel) in the
// order a max-heap of (distance, id) pops them — the order hnswlib writes them to the link row. All lanes call it.
// cand_key layout: ord(dist) << 32 | ~id.
template <int NCH>
__device__ uint32_t heuristic_select(const BuildDev& g, QReg<NCH>& q, float* qs, const unsigned long long* keys, uint32_t m,
                                     uint32_t Mlim, unsigned long long* sel, uint32_t lane, unsigned long long& n_dots) {
    uint32_t n_sel = 0;
    const uint32_t dim = g.dim;
    if(m < Mlim) {                                   // hnswlib returns early: every candidate is kept
        for(uint32_t i = lane; i < m; i += 32) sel[i] = keys[i];
        n_sel = m;
        __syncwarp();
    } else {
        for(uint32_t i = 0; i < m && n_sel < Mlim; i++) {
            const unsigned long long k = keys[i];
            const uint32_t c = ~(uint32_t) k;
            const float dq = unord_f32((uint32_t) (k >> 32));
            bool good = true;
            if(n_sel) {
                load_q<NCH>(q, qs, g.vectors + (size_t) c * dim, dim, lane);
                for(uint32_t j = 0; j < n_sel; j += 2) {
                    const uint32_t s0 = ~(uint32_t) sel[j];
                    const uint32_t s1 = (j + 1 < n_sel) ? ~(uint32_t) sel[j + 1] : s0;
                    float d0, d1;
                    dot_two<NCH>(q, qs, g.vectors + (size_t) s0 * dim, g.vectors + (size_t) s1 * dim, dim, lane, d0, d1);
                    d0 = 1.0f - d0; d1 = 1.0f - d1;
                    n_dots += (j + 1 < n_sel) ? 2 : 1;
                    if(d0 < dq) { good = false; break; }                 // hnswlib tests the chosen ones in order and stops at the first
                    if(j + 1 < n_sel && d1 < dq) { good = false; break; }
                }
            }
            if(good) { if(lane == 0) sel[n_sel] = k; n_sel++; __syncwarp(); }
        }
    }
    // row order = pops of a max-heap of (distance, id): distance descending, larger id first among equals. sel is ascending
    // by (distance, ~id), so walking it backwards gives distance descending with SMALLER id first among equals: fix the runs.
    if(lane == 0 && n_sel > 1) {
        for(uint32_t a = 0, b = n_sel - 1; a < b; a++, b--) { const unsigned long long t = sel[a]; sel[a] = sel[b]; sel[b] = t; }
        uint32_t r0 = 0;
        while(r0 < n_sel) {
            uint32_t r1 = r0 + 1;
            while(r1 < n_sel && (uint32_t) (sel[r1] >> 32) == (uint32_t) (sel[r0] >> 32)) r1++;
            for(uint32_t a = r0, b = r1 - 1; a < b; a++, b--) { const unsigned long long t = sel[a]; sel[a] = sel[b]; sel[b] = t; }
            r0 = r1;
        }
    }
    __syncwarp();
    return n_sel;
}

// 4 warps per CTA, one new node per warp at a time. Shared memory per warp: [efc+1] u64 result heap (reused as the sorted
// candidate list), [kCandSmem] u64 candidate tier, [kVisSmem] u32 visited tier 1, [M] u64 chosen neighbours.
template <int NCH>
__global__ void __launch_bounds__(kKnnThreads, TSGPU_KNN_MIN_CTAS)
insert_search_kernel(const __grid_constant__ BuildDev g, const __grid_constant__ RoundParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * (kKnnThreads / 32) + warp;
    const uint32_t ef = g.efc, dim = g.dim, dim_pad = (dim + 3) & ~3u, M = g.M;
    const size_t per_warp = (size_t) (ef + 1) + kCandSmem + kVisSmem / 2 + M;          // u64 units
    unsigned long long* res = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t) warp * per_warp;
    unsigned long long* cand_s = res + (ef + 1);
    uint32_t* vis1 = reinterpret_cast<uint32_t*>(cand_s + kCandSmem);
    unsigned long long* sel = cand_s + kCandSmem + kVisSmem / 2;
    float* qs = reinterpret_cast<float*>(reinterpret_cast<unsigned long long*>(smem_raw) + (size_t) (kKnnThreads / 32) * per_warp) + (size_t) warp * dim_pad;
    uint32_t* vis2 = P.vis2 + (size_t) slot * P.vis2_slots;
    const uint32_t mask2 = P.vis2_slots - 1, limit2 = P.vis2_slots - (P.vis2_slots >> 2);
    unsigned long long* cand = P.cand + (size_t) slot * P.cand_cap;
    unsigned long long n_dist_acc = 0, n_exp_acc = 0, n_dots_acc = 0;

    for(;;) {
        uint32_t t = 0;
        if(lane == 0) t = atomicAdd(P.ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if(t >= P.count) break;
        const uint32_t c = P.first + t;
        const float* qv = g.vectors + (size_t) c * dim;
        QReg<NCH> q;
        load_q<NCH>(q, qs, qv, dim, lane);
        const uint32_t curlevel = g.levels[c];
        uint32_t cur = P.entry;
        float curdist = 0.f;
        if(curlevel < P.max_level) {
            curdist = 1.0f - dot_one<NCH>(q, qs, g.vectors + (size_t) cur * dim, dim, lane);
            n_dist_acc++;
            for(uint32_t level = P.max_level; level > curlevel; level--) {
                bool changed = true;
                while(changed) {
                    changed = false;
                    const uint32_t* rec = row_ptr(g, cur, level);
                    const uint32_t size = rec[0];
                    const uint32_t nb = (lane < size) ? rec[1 + lane] : kNone;
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
        }
        const uint32_t top_level = curlevel < P.max_level ? curlevel : P.max_level;
        for(int level = (int) top_level; level >= 0; level--) {
            // ---- searchBaseLayer(cur, q, level), ef = ef_construction, no filter
            for(uint32_t i = lane; i < kVisSmem; i += 32) vis1[i] = 0;
            uint32_t n_res = 0, n_cand = 0, n_cs = 0, n_cg = 0, n_v1 = 0, n_v2 = 0;
            bool overflow = false;
            float lowerBound;
            __syncwarp();
            {
                const float d = 1.0f - dot_one<NCH>(q, qs, g.vectors + (size_t) cur * dim, dim, lane);
                n_dist_acc++;
                lowerBound = d;
                if(lane == 0) {
                    heap_push_max(res, n_res, res_key(d, cur)); heap_push_min(cand_s, n_cs, cand_key(d, cur));
                    vis_insert(vis1, vis2, mask2, cur, false);
                }
                n_v1 = 1; n_res = 1; n_cand = 1;
                __syncwarp();
            }
            const uint32_t Lrow = level ? M : 2 * M;
            while(n_cand && !overflow) {
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
                const uint32_t* rec = row_ptr(g, cnode, (uint32_t) level);
                const uint32_t size = min(rec[0], Lrow);
                const uint32_t nb = (lane < size) ? rec[1 + lane] : kNone;          // 2M <= 32
                const bool use2 = n_v1 + 32 > kVisSmemLimit;
                if(use2 && n_v2 + 32 > limit2) { overflow = true; break; }
                const bool fresh = (nb != kNone) && vis_insert(vis1, vis2, mask2, nb, use2);
                uint32_t mask = __ballot_sync(0xffffffffu, fresh);
                if(use2) n_v2 += __popc(mask); else n_v1 += __popc(mask);
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
                    float d0, d1;
                    dot_two<NCH>(q, qs, g.vectors + (size_t) c0 * dim, g.vectors + (size_t) c1 * dim, dim, lane, d0, d1);
                    d0 = 1.0f - d0; d1 = 1.0f - d1;
                    n_dist_acc += j1 >= 0 ? 2 : 1;
                    uint32_t ovf = 0;
                    if(lane == 0) {
#pragma unroll
                        for(int tt = 0; tt < 2; tt++) {
                            if(tt == 1 && j1 < 0) break;
                            const float d = tt ? d1 : d0;
                            const uint32_t cc = tt ? c1 : c0;
                            if(n_res < ef || lowerBound > d) {
                                if(n_cs < kCandSmem) heap_push_min(cand_s, n_cs, cand_key(d, cc));
                                else if(n_cg < P.cand_cap) heap_push_min(cand, n_cg, cand_key(d, cc));
                                else ovf = 1;
                                if(n_res < ef) heap_push_max(res, n_res, res_key(d, cc));
                                else heap_replace_max(res, n_res, res_key(d, cc));
                                lowerBound = unord_f32((uint32_t) (res[0] >> 32));
                            }
                        }
                    }
                    lowerBound = __shfl_sync(0xffffffffu, lowerBound, 0);
                    n_res = __shfl_sync(0xffffffffu, n_res, 0);
                    n_cand = __shfl_sync(0xffffffffu, n_cs + n_cg, 0);
                    if(__shfl_sync(0xffffffffu, ovf, 0)) { overflow = true; break; }
                }
            }
            if(n_v2) {
                uint4* z = reinterpret_cast<uint4*>(vis2);
                for(uint32_t i = lane; i < (P.vis2_slots >> 2); i += 32) z[i] = make_uint4(0, 0, 0, 0);
            }
            if(overflow) { if(lane == 0) *P.error = 1; break; }
            // ---- heap -> ascending list in place (heapsort: the maximum moves to the end), then as candidate keys
            // (ord(dist) << 32 | ~id) in queue_closest order: distance ascending, larger id first among equals
            const uint32_t m = n_res;
            if(lane == 0) {
                uint32_t nn = n_res;
                while(nn > 1) { const unsigned long long mx = res[0]; heap_pop_max(res, nn); res[nn] = mx; }
                for(uint32_t i = 0; i < m; i++) res[i] = (res[i] & 0xFFFFFFFF00000000ull) | (uint32_t) ~(uint32_t) res[i];
                uint32_t r0 = 0;
                while(r0 < m) {                                   // res_key order put smaller ids first among equal distances
                    uint32_t r1 = r0 + 1;
                    while(r1 < m && (uint32_t) (res[r1] >> 32) == (uint32_t) (res[r0] >> 32)) r1++;
                    for(uint32_t a = r0, b = r1 - 1; a < b; a++, b--) { const unsigned long long tk = res[a]; res[a] = res[b]; res[b] = tk; }
                    r0 = r1;
                }
            }
            __syncwarp();
            // ---- getNeighborsByHeuristic2(top_candidates, M) and the node's own row (mutuallyConnectNewElement, first half)
            const uint32_t n_sel = heuristic_select<NCH>(g, q, qs, res, m, M, sel, lane, n_dots_acc);
            uint32_t* row = row_ptr(g, c, (uint32_t) level);
            if(lane < n_sel) row[1 + lane] = ~(uint32_t) sel[lane];
            if(lane == 0) row[0] = n_sel;
            cur = ~(uint32_t) sel[n_sel - 1];                     // next_closest_entry_point = the closest chosen neighbour
            __syncwarp();
            load_q<NCH>(q, qs, qv, dim, lane);                    // the heuristic used q's registers for its candidates
        }
    }
    if(lane == 0) { atomicAdd(P.stats + 0, n_dist_acc); atomicAdd(P.stats + 1, n_exp_acc); atomicAdd(P.stats + 2, n_dots_acc); }
}

// ---------------------------------------------------------------------------------------------------------------
// Reverse links of a round. A modified row is identified by rid: level 0 -> node id, level l > 0 -> n + record index.
struct RevParams {
    uint32_t first, count;
    uint32_t max_level;                // top level when the round began (rows above it stay empty, as in hnswlib)
    uint32_t* row_cnt;                 // [n + R] zero between rounds: new links per row
    uint32_t* row_slot;                // [n + R] target slot of a row (valid where row_cnt > 0)
    uint32_t* n_targets;               // zero before the round
    uint32_t* t_rid; uint32_t* t_node; uint32_t* t_level; uint32_t* t_off; uint32_t* t_fill;   // [cap_targets]
    uint32_t* jobs;                    // [cap_jobs] new node ids grouped by target
    uint32_t cap_targets;
    unsigned long long* stats;         // [2] heuristic dots, [3] rows re-selected
};

// one thread per (new node, level, link slot)
__global__ void rev_count_kernel(const __grid_constant__ BuildDev g, const __grid_constant__ RevParams P, int fill) {
    const uint32_t slots = 2 * g.M;                                    // per (node, level); level > 0 uses the first M
    const unsigned long long tid = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t t = (uint32_t) (tid / slots), j = (uint32_t) (tid % slots);
    // thread space: count * (max_level + 1) * slots
    const uint32_t per_node = P.max_level + 1;
    const uint32_t ni = t / per_node, level = t % per_node;
    if(ni >= P.count) return;
    const uint32_t c = P.first + ni;
    if(level > g.levels[c]) return;
    const uint32_t* row = row_ptr(g, c, level);
    if(j >= row[0]) return;
    const uint32_t s = row[1 + j];
    const uint32_t rid = level == 0 ? s : g.n + (uint32_t) (g.upper_off[s] + (level - 1));
    if(!fill) {
        const uint32_t old = atomicAdd(P.row_cnt + rid, 1u);
        if(old == 0) {
            const uint32_t ts = atomicAdd(P.n_targets, 1u);
            if(ts < P.cap_targets) { P.row_slot[rid] = ts; P.t_rid[ts] = rid; P.t_node[ts] = s; P.t_level[ts] = level; }
        }
    } else {
        const uint32_t ts = P.row_slot[rid];
        const uint32_t pos = atomicAdd(P.t_fill + ts, 1u);
        P.jobs[P.t_off[ts] + pos] = c;
    }
}

// exclusive scan of the targets' link counts (single CTA), also clears the fill cursors
__global__ void __launch_bounds__(1024) rev_scan_kernel(const __grid_constant__ RevParams P) {
    __shared__ uint32_t s[1024];
    __shared__ uint32_t carry;
    const uint32_t n = min(*P.n_targets, P.cap_targets);
    if(threadIdx.x == 0) carry = 0;
    __syncthreads();
    for(uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? P.row_cnt[P.t_rid[i]] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for(uint32_t d = 1; d < 1024; d <<= 1) {
            const uint32_t tt = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += tt;
            __syncthreads();
        }
        if(i < n) { P.t_off[i] = carry + s[threadIdx.x] - v; P.t_fill[i] = 0; }
        __syncthreads();
        if(threadIdx.x == 0) carry += s[1023];
        __syncthreads();
    }
}

// One warp per modified row: the round's new links in ascending node order (the order sequential insertion would apply
// them), each by mutuallyConnectNewElement's second half: append while the row has room, else re-select among the row and
// the new node with the heuristic (distances to the row's owner).
template <int NCH>
__global__ void __launch_bounds__(kKnnThreads)
rev_apply_kernel(const __grid_constant__ BuildDev g, const __grid_constant__ RevParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dim = g.dim, dim_pad = (dim + 3) & ~3u, M = g.M;
    // per warp: keys[64] + sel[64] u64, then the generic-path vector
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t) warp * 128;
    unsigned long long* sel = keys + 64;
    float* qs = reinterpret_cast<float*>(reinterpret_cast<unsigned long long*>(smem_raw) + (size_t) (kKnnThreads / 32) * 128) + (size_t) warp * dim_pad;
    const uint32_t n_t = min(*P.n_targets, P.cap_targets);
    unsigned long long n_dots = 0, n_resel = 0;
    for(uint32_t t = blockIdx.x * (kKnnThreads / 32) + warp; t < n_t; t += gridDim.x * (kKnnThreads / 32)) {
        const uint32_t rid = P.t_rid[t], s = P.t_node[t], level = P.t_level[t];
        const uint32_t nj = P.row_cnt[rid];
        uint32_t* jobs = P.jobs + P.t_off[t];
        // ascending node order (a handful of entries: insertion sort by one lane)
        if(lane == 0) {
            for(uint32_t i = 1; i < nj; i++) { const uint32_t v = jobs[i]; uint32_t k = i; while(k > 0 && jobs[k - 1] > v) { jobs[k] = jobs[k - 1]; k--; } jobs[k] = v; }
            P.row_cnt[rid] = 0;                                  // back to all-zero for the next round
        }
        __syncwarp();
        uint32_t* row = row_ptr(g, s, level);
        const uint32_t Mcur = level ? M : 2 * M;
        QReg<NCH> q;
        for(uint32_t ji = 0; ji < nj; ji++) {
            const uint32_t c = jobs[ji];
            const uint32_t sz = row[0];
            if(sz < Mcur) {
                if(lane == 0) { row[1 + sz] = c; row[0] = sz + 1; }
                __syncwarp();
                continue;
            }
            n_resel++;
            // candidates = the row (sz == Mcur <= 32 entries) + c, distances to the row's owner s
            load_q<NCH>(q, qs, g.vectors + (size_t) s * dim, dim, lane);
            const uint32_t mine = lane < sz ? row[1 + lane] : kNone;
            for(uint32_t i = 0; i <= sz; i += 2) {
                const uint32_t c0 = i < sz ? __shfl_sync(0xffffffffu, mine, i) : c;
                const uint32_t c1 = (i + 1 < sz) ? __shfl_sync(0xffffffffu, mine, i + 1) : c;     // i + 1 == sz: the new node
                float d0, d1;
                dot_two<NCH>(q, qs, g.vectors + (size_t) c0 * dim, g.vectors + (size_t) c1 * dim, dim, lane, d0, d1);
                n_dots += (i + 1 <= sz) ? 2 : 1;
                if(lane == 0) {
                    sel[i] = cand_key(1.0f - d0, c0);
                    if(i + 1 <= sz) sel[i + 1] = cand_key(1.0f - d1, c1);
                }
            }
            __syncwarp();
            // rank sort of the sz + 1 <= 33 unique keys, ascending (= queue_closest's pop order)
            const uint32_t m = sz + 1;
            for(uint32_t i = lane; i < m; i += 32) {
                const unsigned long long k = sel[i];
                uint32_t r = 0;
                for(uint32_t j = 0; j < m; j++) r += sel[j] < k;
                keys[r] = k;
            }
            __syncwarp();
            const uint32_t n_sel = heuristic_select<NCH>(g, q, qs, keys, m, Mcur, sel, lane, n_dots);
            if(lane < n_sel) row[1 + lane] = ~(uint32_t) sel[lane];
            if(lane == 0) row[0] = n_sel;
            __syncwarp();
        }
    }
    if(lane == 0 && (n_dots || n_resel)) { atomicAdd(P.stats + 2, n_dots); atomicAdd(P.stats + 3, n_resel); }
}

// the five launches of one round, on one stream
template <int NCH>
cudaError_t launch_build_round(const BuildDev& g, const RoundParams& R, const RevParams& V, unsigned grid_a, size_t smem_a,
                               unsigned grid_rev, size_t smem_rev, cudaStream_t st) {
    cudaFuncSetAttribute(insert_search_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_a, 48 * 1024));
    cudaFuncSetAttribute(rev_apply_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_rev, 48 * 1024));
    insert_search_kernel<NCH><<<grid_a, tsv::kKnnThreads, smem_a, st>>>(g, R);
    const unsigned long long threads = (unsigned long long) V.count * (V.max_level + 1) * 2 * g.M;
    const unsigned blocks = (unsigned) ((threads + 255) / 256);
    rev_count_kernel<<<blocks, 256, 0, st>>>(g, V, 0);
    rev_scan_kernel<<<1, 1024, 0, st>>>(V);
    rev_count_kernel<<<blocks, 256, 0, st>>>(g, V, 1);
    rev_apply_kernel<NCH><<<grid_rev, tsv::kKnnThreads, smem_rev, st>>>(g, V);
    return cudaGetLastError();
}


}  // namespace tsb
