// Vector-result assembly and hybrid rank fusion (K8) on sm_100a.
//
// Replaces (reference file:line):
//   wildcard + vector query result loop            src/index.cpp:3682-3732
//   process_results_hnsw_index post-processing     src/index.cpp:3389-3438
//   hybrid reciprocal-rank fusion                   src/index.cpp:4036-4221
//       including its exact Topster behaviour: topster->sort() turns the heap array into a descending array, the
//       scores are then rewritten in place, and vector-only hits go through Topster::add() on that array
//       (include/topster.h:321-466). Results depend on that sequence, so it is replayed step for step: one CTA per
//       query, the sequential part on warp 0 with lane-parallel key look-ups.
#pragma once
#include "kw_kernels.cuh"

namespace tsf {

using namespace tsk;

struct VecParams {
    uint32_t k;                 // effective k used for the knn call
    float distance_threshold;
    float alpha;
    uint32_t metric;            // 1 = cosine -> abs()
};

// ---------------------------------------------------------------------------------------------------------------
// Wildcard + vector query: (seq_id, dist) stream -> KVs with sort scores -> top-K in KV order.
struct VecAssembleParams {
    const QDesc* qd;
    const float* res_dist;      // [nq*res_stride] (knn) or concatenated (flat)
    const uint32_t* res_ids;
    const uint32_t* res_n;      // [nq] for knn layout
    const float* flat_dist;     // flat (brute-force) results, concatenated; query q owns [res_off[q], res_off[q+1])
    const uint32_t* flat_ids;
    const unsigned long long* res_off;
    const uint8_t* q_is_flat;   // [nq] or nullptr (no flat query in the batch)
    uint32_t res_stride;
    VecParams vp;
    KVOut* out_kv; uint32_t* out_count; uint32_t* out_found;
    uint32_t kv_stride, KP;
};

__global__ void __launch_bounds__(kFinalThreads)
vec_assemble_kernel(const __grid_constant__ VecAssembleParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t KP = P.KP, N2 = 2 * KP;
    TopBuf tb;
    tb.s0 = reinterpret_cast<int64_t*>(smem_raw);
    tb.s1 = tb.s0 + N2; tb.s2 = tb.s1 + N2;
    tb.key = reinterpret_cast<uint32_t*>(tb.s2 + N2);
    tb.vd = reinterpret_cast<float*>(tb.key + N2);
    tb.cmb = nullptr;
    __shared__ uint32_t s_warp[8];
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    const QDesc qd = P.qd[q];
    const uint32_t K = qd.topk;
    SortSpec SS;
    for(int i = 0; i < 3; i++) { SS.type[i] = qd.sort_type[i]; SS.order[i] = qd.sort_order[i]; SS.missing_first[i] = qd.missing_first[i]; SS.col[i] = qd.sort_col[i]; }
    const bool flat = P.q_is_flat && P.q_is_flat[q];
    const float* rd; const uint32_t* ri; unsigned long long n_in;
    if(flat) { rd = P.flat_dist + P.res_off[q]; ri = P.flat_ids + P.res_off[q]; n_in = P.res_off[q + 1] - P.res_off[q]; }
    else { rd = P.res_dist + (size_t) q * P.res_stride; ri = P.res_ids + (size_t) q * P.res_stride; n_in = P.res_n[q]; }

    auto reduce = [&](uint32_t n) -> uint32_t {
        tb_fill_invalid(tb, n, N2);
        __syncthreads();
        tb_sort<false>(tb, N2);
        uint32_t lo = 0, hi = N2;
        while(lo < hi) { const uint32_t mid = (lo + hi) >> 1; if(tb.key[mid] != kNone) lo = mid + 1; else hi = mid; }
        return lo < K ? lo : K;
    };
    uint32_t n = 0, found = 0;
    for(unsigned long long base = 0; base < n_in; base += kFinalThreads) {
        if(n + kFinalThreads > N2) { n = reduce(n); __syncthreads(); }
        bool keep = false;
        int64_t sc[3] = {0, 0, 0};
        uint32_t id = 0; float vds = 0.f;
        if(base + tid < n_in) {
            id = ri[base + tid];
            const float d = rd[base + tid];
            vds = (P.vp.metric == 1) ? fabsf(d) : d;                     // src/index.cpp:3699-3700
            keep = !(vds > P.vp.distance_threshold);
            if(keep) compute_sort_scores(SS, id, 0, vds, sc);
        }
        uint32_t tot;
        const uint32_t rank = cta_rank(keep, s_warp, &tot);
        if(keep) { const uint32_t s = n + rank; tb.s0[s] = sc[0]; tb.s1[s] = sc[1]; tb.s2[s] = sc[2]; tb.key[s] = id; tb.vd[s] = vds; }
        n += tot; found += tot;
        __syncthreads();
    }
    n = reduce(n);
    __syncthreads();
    int msi = -1;
    for(int i = 0; i < 3; i++) if(qd.sort_type[i] == 1) msi = i;
    const uint32_t n_out = n < P.kv_stride ? n : P.kv_stride;
    for(uint32_t i = tid; i < n_out; i += kFinalThreads) {
        KVOut kv;
        kv.key = tb.key[i]; kv.distinct_key = tb.key[i];
        kv.scores[0] = tb.s0[i]; kv.scores[1] = tb.s1[i]; kv.scores[2] = tb.s2[i];
        kv.text_match_score = msi >= 0 ? kv.scores[msi] : 0;
        kv.vector_distance = tb.vd[i];
        kv.match_score_index = (int8_t) msi; kv.pad0 = 0; kv.query_index = 0;
        P.out_kv[(size_t) q * P.kv_stride + i] = kv;
    }
    if(tid == 0) { P.out_count[q] = n_out; P.out_found[q] = found; }
}

// ---------------------------------------------------------------------------------------------------------------
// Hybrid fusion.
struct HybridParams {
    IndexDev ix;
    const QDesc* qd;
    const CDesc* cd;
    uint32_t F;
    uint32_t field_ids[kMaxFieldSlots];
    const KVOut* kw_kv;          // [nq*kw_stride] keyword Topster after sort()
    const uint32_t* kw_count;    // [nq]
    const uint32_t* kw_found;    // [nq]
    const uint32_t* kw_searched; // [nq] searched_queries.size()
    uint32_t kw_stride;
    const float* res_dist; const uint32_t* res_ids; const uint32_t* res_n;
    const float* flat_dist; const uint32_t* flat_ids;
    const unsigned long long* res_off; const uint8_t* q_is_flat;
    uint32_t res_stride;
    VecParams vp;
    KVOut* out_kv; uint32_t* out_count; uint32_t* out_found;
    uint32_t kv_stride;
    uint32_t KMAX;               // max topk in batch (shared memory sizing)
    uint32_t VMAX;               // max knn results per query (<= 1024)
    // rerank_hybrid_matches (Index::compute_aux_scores): the query vectors and the vector index for the keyword-only results' distances
    const float* queries;        // [nq * dim]
    const float* vectors;        // [n_nodes * dim]
    uint32_t dim, n_nodes;
};

__device__ __forceinline__ bool kvo_greater(const KVOut& a, const KVOut& b) {
    return kv_greater(a.scores[0], a.scores[1], a.scores[2], (uint32_t) a.key, b.scores[0], b.scores[1], b.scores[2], (uint32_t) b.key);
}
__device__ __forceinline__ bool kvo_smaller(const KVOut& a, const KVOut& b) { return kvo_greater(b, a); }

// does `id` satisfy some token combination of the query (every required row present in at least one field)?
__device__ bool is_keyword_match(const HybridParams& P, const QDesc& qd, uint32_t id) {
    for(uint32_t c = qd.combo_begin; c < qd.combo_end; c++) {
        const CDesc& cd = P.cd[c];
        bool all = cd.req_mask != 0;
        for(uint32_t r = 0; r < cd.n_rows && all; r++) {
            if(!((cd.req_mask >> r) & 1)) continue;
            bool any = false;
            for(uint32_t f = 0; f < P.F && !any; f++) {
                const uint32_t l = cd.lists[r * P.F + f];
                if(l == kNone) continue;
                const DevField& g = P.ix.fields[P.field_ids[f]];
                any = probe_list(g, l, g.list_blk_off[l], g.list_blk_off[l + 1] - 1, id) != kNone;
            }
            all = any;
        }
        if(all) return true;
    }
    return false;
}

// compute_aggregated_score for ONE doc over the tokens of the query's first combination (every token optional, total_cost 0):
// what compute_aux_scores gives a result that only the vector query found (src/index.cpp:8801-8845)
__device__ int64_t aux_text_score(const HybridParams& P, const QDesc& qd, uint32_t doc) {
    if(qd.combo_end == qd.combo_begin) return 0;
    const CDesc& cd = P.cd[qd.combo_begin];
    ScoreParams SP;
    SP.total_cost = 0; SP.num_query_tokens = cd.n_rows; SP.syn_orig_num_tokens = -1; SP.orig_num_tokens = (int32_t) cd.n_rows;
    SP.is_synonym_query = 0; SP.demote_synonym_match = 0;
    SP.prioritize_exact_match = qd.flags & 1; SP.prioritize_token_position = (qd.flags >> 1) & 1; SP.prioritize_num_matching_fields = (qd.flags >> 2) & 1;
    SP.match_type = qd.match_type;
    FieldAgg agg; field_agg_init(agg);
    uint32_t found_rows = 0, query_len = 0;
    for(uint32_t f = 0; f < P.F; f++) {
        const DevField& g = P.ix.fields[P.field_ids[f]];
        RawTok toks[kMaxTokens];
        int nt = 0;
        for(uint32_t r = 0; r < cd.n_rows; r++) {
            const uint32_t l = cd.lists[r * P.F + f];
            if(l == kNone) continue;
            const uint32_t b0 = g.list_blk_off[l], b1 = g.list_blk_off[l + 1];
            if(b1 == b0) continue;
            const uint32_t h = probe_list(g, l, b0, b1 - 1, doc);
            if(h == kNone) continue;
            const unsigned long long p = g.list_off[l] + h;
            const unsigned long long o0 = g.pos_off[p], o1 = g.pos_off[p + 1];
            toks[nt].p = g.positions + o0; toks[nt].n = (uint32_t) (o1 - o0); nt++;
            found_rows |= 1u << r;
        }
        if(nt == 0) continue;
        const bool single_exact = (SP.total_cost == 0 && SP.num_query_tokens == 1);
        const int64_t fs = score_field(SP, (g.is_array & kFieldIsArray) != 0, single_exact, toks, nt);
        field_agg_add(agg, SP.match_type, fs, (int64_t) qd.field_weight[f]);
    }
    query_len = __popc(found_rows);
    return (int64_t) field_agg_finish(agg, SP, query_len);
}

__global__ void __launch_bounds__(kThreads)
hybrid_fuse_kernel(const __grid_constant__ HybridParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KVOut* data = reinterpret_cast<KVOut*>(smem_raw);                       // [KMAX]
    uint16_t* kvs = reinterpret_cast<uint16_t*>(data + P.KMAX);             // heap position -> data slot, [pow2 >= KMAX]
    uint32_t KP2 = 1; while(KP2 < P.KMAX) KP2 <<= 1;
    uint32_t VP2 = 1; while(VP2 < P.VMAX) VP2 <<= 1;
    uint32_t* v_id = reinterpret_cast<uint32_t*>(kvs + KP2 + (KP2 & 1));    // [VP2] knn results sorted by seq_id
    float* v_dist = reinterpret_cast<float*>(v_id + VP2);
    uint32_t* v_rank = reinterpret_cast<uint32_t*>(v_dist + VP2);
    // per vector result, computed by all threads before the serial replay (graph results only): the three sort scores of
    // the vector-only KV, its match_score_index, and flags (bit0 passes the filter, bit1 counts towards `found`)
    int64_t* v_sc = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(v_rank + VP2) + ((8 - ((size_t) (v_rank + VP2) & 7)) & 7));
    int8_t* v_msi = reinterpret_cast<int8_t*>(v_sc + 3 * VP2);
    uint8_t* v_flag = reinterpret_cast<uint8_t*>(v_msi + VP2);
    __shared__ uint32_t s_size, s_vec_only_new;

    const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const QDesc qd = P.qd[q];
    const uint32_t MAXSZ = qd.topk;
    SortSpec SS;
    for(int i = 0; i < 3; i++) { SS.type[i] = qd.sort_type[i]; SS.order[i] = qd.sort_order[i]; SS.missing_first[i] = qd.missing_first[i]; SS.col[i] = qd.sort_col[i]; }
    const float VECTOR_SEARCH_WEIGHT = P.vp.alpha;
    const float TEXT_MATCH_WEIGHT = (float) (1.0 - (double) VECTOR_SEARCH_WEIGHT);

    // ---- keyword Topster after topster->sort()
    const uint32_t n_kw = P.kw_count[q];
    for(uint32_t i = tid; i < n_kw; i += kThreads) { data[i] = P.kw_kv[(size_t) q * P.kw_stride + i]; kvs[i] = (uint16_t) i; }
    if(tid == 0) { s_size = n_kw; s_vec_only_new = 0; }
    // ---- vector results: rank = position in the distance-ordered list; iteration order = ascending seq_id
    const bool flat = P.q_is_flat && P.q_is_flat[q];
    unsigned long long n_v = 0;
    const float* rd; const uint32_t* ri;
    if(flat) { rd = P.flat_dist + P.res_off[q]; ri = P.flat_ids + P.res_off[q]; n_v = P.res_off[q + 1] - P.res_off[q]; }
    else {
        rd = P.res_dist + (size_t) q * P.res_stride; ri = P.res_ids + (size_t) q * P.res_stride;
        // drop > distance_threshold first (process_results_hnsw_index, :3419-3428); the survivors keep their
        // (distance, seq_id) order, which is what the stable re-sort by distance yields
        const uint32_t n_raw = P.res_n[q];
        if(tid == 0) {
            uint32_t m = 0;
            for(uint32_t i = 0; i < n_raw; i++) {
                const float d = rd[i];
                const float s = (P.vp.metric == 1) ? fabsf(d) : d;
                if(s > P.vp.distance_threshold) continue;
                v_id[m] = ri[i]; v_dist[m] = d; v_rank[m] = m; m++;
            }
            s_vec_only_new = m;        // borrowed as a temporary
        }
        __syncthreads();
        n_v = s_vec_only_new;
        __syncthreads();
        if(tid == 0) s_vec_only_new = 0;
        for(uint32_t i = (uint32_t) n_v + tid; i < VP2; i += kThreads) v_id[i] = kNone;
        __syncthreads();
        // bitonic sort by seq_id ascending
        for(uint32_t k = 2; k <= VP2; k <<= 1) {
            for(uint32_t j = k >> 1; j > 0; j >>= 1) {
                for(uint32_t t = tid; t < (VP2 >> 1); t += kThreads) {
                    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i + j;
                    const bool up = ((i & k) == 0);
                    const bool sw = up ? (v_id[p] < v_id[i]) : (v_id[i] < v_id[p]);
                    if(sw) {
                        uint32_t a = v_id[i]; v_id[i] = v_id[p]; v_id[p] = a;
                        float b = v_dist[i]; v_dist[i] = v_dist[p]; v_dist[p] = b;
                        uint32_t c = v_rank[i]; v_rank[i] = v_rank[p]; v_rank[p] = c;
                    }
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();

    // ---- text ranks (src/index.cpp:4094-4112): rank grows only when the score strictly drops
    if(tid == 0) {
        int64_t text_rank = 0, last = INT64_MAX;
        for(uint32_t i = 0; i < n_kw; i++) {
            KVOut& r = data[kvs[i]];
            if(r.match_score_index < 0 || r.match_score_index > 2) continue;
            r.text_match_score = r.scores[r.match_score_index];
            if(r.text_match_score < last) ++text_rank;
            last = r.text_match_score;
            r.scores[r.match_score_index] = float_to_int64(__double2float_rn((1.0 / (double) text_rank) * (double) TEXT_MATCH_WEIGHT));
        }
    }
    __syncthreads();

    // ---- order-independent part of the loop body, one vector result per thread: filter bit, the vector-only KV's sort
    // scores, and whether the id adds to `found` (a vector result that is not a keyword match; ids found in the Topster
    // are keyword matches, so the count does not depend on the replay). Takes ~100 dependent global loads per result off
    // the single-lane critical path.
    if(!flat) {
        for(uint32_t vi = tid; vi < (uint32_t) n_v; vi += kThreads) {
            const uint32_t seq_id = v_id[vi];
            uint8_t fl = 0;
            bool pass = !qd.filter_empty;
            if(pass && qd.filter_bitmap) pass = (qd.filter_bitmap[seq_id >> 5] >> (seq_id & 31)) & 1;
            if(pass) {
                fl = 1;
                // all_result_ids |= vec_search_ids (src/index.cpp:4197, 4215-4219) when the query keeps its id set
                if(qd.keep_all && qd.found_bitmap) atomicOr(qd.found_bitmap + (seq_id >> 5), 1u << (seq_id & 31));
                const double vec_part = (1.0 / (double) (v_rank[vi] + 1)) * (double) VECTOR_SEARCH_WEIGHT;
                int64_t sc[3];
                v_msi[vi] = (int8_t) compute_sort_scores(SS, seq_id, float_to_int64(__double2float_rn(vec_part)), v_dist[vi], sc);
                v_sc[3 * vi] = sc[0]; v_sc[3 * vi + 1] = sc[1]; v_sc[3 * vi + 2] = sc[2];
                bool in_kw = true;
                if(qd.n_excl && excluded(qd.excl, qd.n_excl, seq_id)) in_kw = false;
                if(in_kw) in_kw = is_keyword_match(P, qd, seq_id);
                if(!in_kw) fl |= 2;
            }
            v_flag[vi] = fl;
        }
    }
    __syncthreads();

    // ---- replay of the fusion loop (warp 0)
    if(tid < 32) {
        uint32_t size = s_size, vec_new = 0;
        for(unsigned long long vi = 0; vi < n_v; vi++) {
            uint32_t seq_id; float dist; uint32_t rank;
            if(flat) { seq_id = ri[vi]; dist = rd[vi]; rank = (uint32_t) vi; }
            else { seq_id = v_id[vi]; dist = v_dist[vi]; rank = v_rank[vi]; }
            if(!flat) { if(!(v_flag[vi] & 1)) continue; }
            else {
                if(qd.filter_bitmap && !((qd.filter_bitmap[seq_id >> 5] >> (seq_id & 31)) & 1)) continue;
                if(qd.filter_empty) continue;
                if(lane == 0 && qd.keep_all && qd.found_bitmap) atomicOr(qd.found_bitmap + (seq_id >> 5), 1u << (seq_id & 31));
            }
            // topster->map.find(seq_id)
            uint32_t found_pos = kNone;
            for(uint32_t b0 = 0; b0 < size; b0 += 32) {
                const uint32_t i = b0 + lane;
                const bool hit = i < size && (uint32_t) data[kvs[i]].key == seq_id;
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if(m) { found_pos = b0 + __ffs(m) - 1; break; }
            }
            if(lane == 0) {
                const double vec_part = (1.0 / (double) (rank + 1)) * (double) VECTOR_SEARCH_WEIGHT;
                if(found_pos != kNone) {
                    KVOut& fk = data[kvs[found_pos]];
                    if(!(fk.match_score_index < 0 || fk.match_score_index > 2)) {
                        fk.vector_distance = dist;
                        const int64_t match_score = float_to_int64(__double2float_rn((double) int64_to_float(fk.scores[fk.match_score_index]) + vec_part));
                        int64_t sc[3];
                        const int msi = compute_sort_scores(SS, seq_id, match_score, dist, sc);
                        fk.scores[0] = sc[0]; fk.scores[1] = sc[1]; fk.scores[2] = sc[2];
                        fk.match_score_index = (int8_t) msi;
                    }
                } else {
                    KVOut kv;
                    int64_t sc[3];
                    int msi;
                    if(!flat) { sc[0] = v_sc[3 * vi]; sc[1] = v_sc[3 * vi + 1]; sc[2] = v_sc[3 * vi + 2]; msi = v_msi[vi]; }
                    else msi = compute_sort_scores(SS, seq_id, float_to_int64(__double2float_rn(vec_part)), dist, sc);
                    kv.key = seq_id; kv.distinct_key = seq_id;
                    kv.scores[0] = sc[0]; kv.scores[1] = sc[1]; kv.scores[2] = sc[2];
                    kv.match_score_index = (int8_t) msi; kv.pad0 = 0;
                    kv.query_index = (uint16_t) P.kw_searched[q];
                    kv.text_match_score = 0;
                    kv.vector_distance = dist;
                    // Topster::add(&kv) on the (no longer heap-ordered) array
                    bool add = true;
                    if(size >= MAXSZ && kvo_smaller(kv, data[kvs[0]])) add = false;
                    if(add) {
                        uint32_t hidx; bool sift_down;
                        if(size < MAXSZ) { sift_down = false; hidx = size; kvs[hidx] = (uint16_t) size; size++; }
                        else { sift_down = true; hidx = 0; }
                        data[kvs[hidx]] = kv;
                        if(sift_down) {
                            while((2 * hidx + 1) < size) {
                                uint32_t next = 2 * hidx + 1;
                                if(next + 1 < size && kvo_greater(data[kvs[next]], data[kvs[next + 1]])) next++;
                                if(kvo_greater(data[kvs[hidx]], data[kvs[next]])) { const uint16_t t = kvs[hidx]; kvs[hidx] = kvs[next]; kvs[next] = t; }
                                else break;
                                hidx = next;
                            }
                        } else {
                            while(hidx > 0) {
                                const uint32_t parent = (hidx - 1) / 2;
                                if(kvo_greater(data[kvs[parent]], data[kvs[hidx]])) { const uint16_t t = kvs[hidx]; kvs[hidx] = kvs[parent]; kvs[parent] = t; hidx = parent; }
                                else break;
                            }
                        }
                    }
                    // vec_search_ids.push_back(seq_id) happens whether or not the heap took it (src/index.cpp:4197)
                    if(!flat) { if(v_flag[vi] & 2) vec_new++; }
                    else {
                        bool in_kw = true;
                        if(qd.n_excl && excluded(qd.excl, qd.n_excl, seq_id)) in_kw = false;
                        if(in_kw) in_kw = is_keyword_match(P, qd, seq_id);
                        if(!in_kw) vec_new++;
                    }
                }
            }
            size = __shfl_sync(0xffffffffu, size, 0);
            __syncwarp();
        }
        if(lane == 0) { s_size = size; s_vec_only_new = vec_new; }
    }
    __syncthreads();

    // ---- rerank_hybrid_matches: Index::compute_aux_scores (src/index.cpp:8793-8923) on the fused Topster
    if((qd.rerank) && s_size) {
        const uint32_t n = s_size;                       // data slots 0 .. n-1 hold the entries (slots are handed out in order)
        for(uint32_t i = tid; i < n; i += kThreads) {    // results only the vector query found: their text match score
            KVOut& kv = data[i];
            if(kv.text_match_score == 0) kv.text_match_score = aux_text_score(P, qd, (uint32_t) kv.key);
        }
        {   // results only the keyword query found: their vector distance (one warp per result, W128 order as everywhere)
            const uint32_t lane2 = tid & 31, warp2 = tid >> 5;
            const float* qv = P.queries + (size_t) q * P.dim;
            for(uint32_t i = warp2; i < n; i += kThreads / 32) {
                const bool need = data[i].vector_distance == -1.0f && data[i].text_match_score != 0 && (uint32_t) data[i].key < P.n_nodes;
                if(!need) continue;                        // warp-uniform
                const float* v = P.vectors + (size_t) (uint32_t) data[i].key * P.dim;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for(uint32_t sgm = 0; sgm * 128 < P.dim; sgm++) {
                    const uint32_t e = sgm * 128 + 4 * lane2;
                    if(e + 0 < P.dim) a0 = __fmaf_rn(qv[e + 0], __ldg(v + e + 0), a0);
                    if(e + 1 < P.dim) a1 = __fmaf_rn(qv[e + 1], __ldg(v + e + 1), a1);
                    if(e + 2 < P.dim) a2 = __fmaf_rn(qv[e + 2], __ldg(v + e + 2), a2);
                    if(e + 3 < P.dim) a3 = __fmaf_rn(qv[e + 3], __ldg(v + e + 3), a3);
                }
                float t = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
                for(int off = 16; off >= 1; off >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, off));
                if(lane2 == 0) data[i].vector_distance = 1.0f - t;
            }
        }
        __syncthreads();
        // ranks: keyword rank by (text_match_score, key) descending; semantic rank by distance ascending, stable on the keyword order
        const double alpha = (double) P.vp.alpha;
        for(uint32_t i = tid; i < n; i += kThreads) {
            const int64_t ti = data[i].text_match_score; const uint32_t ki = (uint32_t) data[i].key; const float di = data[i].vector_distance;
            uint32_t kr = 1;
            for(uint32_t j = 0; j < n; j++) { const int64_t tj = data[j].text_match_score; if(tj > ti || (tj == ti && (uint32_t) data[j].key > ki)) kr++; }
            uint32_t sr = 1;
            for(uint32_t j = 0; j < n; j++) {
                if(j == i) continue;
                const float dj = data[j].vector_distance;
                if(dj < di) sr++;
                else if(dj == di) { const int64_t tj = data[j].text_match_score; if(tj > ti || (tj == ti && (uint32_t) data[j].key > ki)) sr++; }
            }
            data[i].distinct_key = (uint64_t) float_to_int64(__double2float_rn((1.0 / (double) kr) * (1.0 - alpha) + (1.0 / (double) sr) * alpha));   // parked until every rank is known
        }
        __syncthreads();
        for(uint32_t i = tid; i < n; i += kThreads) {
            const int m = data[i].match_score_index;
            if(m >= 0 && m <= 2) data[i].scores[m] = (int64_t) data[i].distinct_key;
            data[i].distinct_key = data[i].key;
        }
        __syncthreads();
    }

    // ---- final topster->sort(): stable_sort by KV order (total order: keys are unique)
    const uint32_t size = s_size;
    for(uint32_t i = size + tid; i < KP2; i += kThreads) kvs[i] = 0xFFFF;
    __syncthreads();
    for(uint32_t k = 2; k <= KP2; k <<= 1) {
        for(uint32_t j = k >> 1; j > 0; j >>= 1) {
            for(uint32_t t = tid; t < (KP2 >> 1); t += kThreads) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i + j;
                const bool up = ((i & k) == 0);
                const uint16_t a = kvs[i], b = kvs[p];
                const bool p_gt = (b != 0xFFFF) && (a == 0xFFFF || kvo_greater(data[b], data[a]));
                const bool i_gt = (a != 0xFFFF) && (b == 0xFFFF || kvo_greater(data[a], data[b]));
                if(up ? p_gt : i_gt) { kvs[i] = b; kvs[p] = a; }
            }
            __syncthreads();
        }
    }
    const uint32_t n_out = size < P.kv_stride ? size : P.kv_stride;
    for(uint32_t i = tid; i < n_out; i += kThreads) P.out_kv[(size_t) q * P.kv_stride + i] = data[kvs[i]];
    if(tid == 0) { P.out_count[q] = n_out; P.out_found[q] = P.kw_found[q] + s_vec_only_new; }
}

}  // namespace tsf
