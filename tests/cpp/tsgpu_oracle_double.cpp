// TEST INFRASTRUCTURE — a test double of the tsgpu C-ABI (include/tsgpu.h) that answers from the CPU oracle, so the C++
// host layer (typesense_b200/host/tsgpu_host.hpp: tokenising, field mirrors, the drop-tokens loop, host_topster_t, the
// call marshalling) can be exercised by tests/cpp/host_scenarios.cpp on a machine WITHOUT a GPU. It lives under tests/,
// is linked only into the CPU build of that one test program, and is never part of libtsgpu.so: the product has no CPU
// path. Layouts of tsgpu_* and tso_* structs are identical by construction (tests/oracle_lib.py passes one ctypes
// struct to both), which the static_asserts below re-check.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tsgpu.h"
#include "../../oracle/ts_oracle.h"
#include "../../typesense_b200/csrc/art_device.cuh"      // art_walk(): the device function, compiled for the host
#include <map>

static_assert(sizeof(tsgpu_kv) == sizeof(tso_kv), "KV layout");
static_assert(sizeof(tsgpu_field) == sizeof(tso_field), "field layout");
static_assert(sizeof(tsgpu_kw_batch) == sizeof(tso_kw_batch), "batch layout");
static_assert(sizeof(tsgpu_hnsw) == sizeof(tso_hnsw), "hnsw layout");
static_assert(sizeof(tsgpu_vec_params) == sizeof(tso_vec_params) + sizeof(uint32_t), "vec params layout: the oracle reads the leading fields (flags only select a device code path)");

namespace {
struct FieldCopy {
    std::vector<uint64_t> list_off, pos_off;
    std::vector<uint32_t> ids, positions;
    tso_field view{};
};
struct Double {
    uint32_t n_docs = 0;
    void* oi = nullptr;
    std::vector<FieldCopy*> fields;
    std::vector<std::vector<int64_t>*> cols;
    std::vector<float> vec; std::vector<uint32_t> labels, links0, links_up; std::vector<uint8_t> levels; std::vector<uint64_t> upper_off;
    tso_hnsw g{};
    bool has_g = false;
    std::vector<std::vector<uint32_t>> filters;     // persistent filters: handle = -(slot + 2)
};
// a batch whose q_filter entries refer to persistent handles (<= -2) is rewritten so they become inline slots appended
// after the batch's own filters — the oracle only knows inline filters
struct Rewritten {
    tso_kw_batch b;
    std::vector<int32_t> q_filter;
    std::vector<uint64_t> filter_off;
    std::vector<uint32_t> filter_ids;
};
static const tso_kw_batch* rewrite(const Double* d, const tsgpu_kw_batch* in, Rewritten& r) {
    const tso_kw_batch* src = reinterpret_cast<const tso_kw_batch*>(in);
    bool any = false;
    for(uint32_t q = 0; q < in->n_queries; q++) if(in->q_filter[q] <= -2) any = true;
    if(!any) return src;
    r.b = *src;
    r.filter_off.assign(1, 0);
    for(uint32_t f = 0; f < in->n_filters; f++) {
        r.filter_ids.insert(r.filter_ids.end(), in->filter_ids + in->filter_off[f], in->filter_ids + in->filter_off[f + 1]);
        r.filter_off.push_back(r.filter_ids.size());
    }
    r.q_filter.assign(in->q_filter, in->q_filter + in->n_queries);
    std::map<int32_t, int32_t> slot_of;                   // one inline slot per distinct handle
    for(uint32_t q = 0; q < in->n_queries; q++) {
        if(in->q_filter[q] > -2) continue;
        auto it = slot_of.find(in->q_filter[q]);
        if(it == slot_of.end()) {
            const auto& ids = d->filters[(size_t) (-(in->q_filter[q] + 2))];
            r.filter_ids.insert(r.filter_ids.end(), ids.begin(), ids.end());
            r.filter_off.push_back(r.filter_ids.size());
            it = slot_of.emplace(in->q_filter[q], (int32_t) r.filter_off.size() - 2).first;
        }
        r.q_filter[q] = it->second;
    }
    if(r.filter_ids.empty()) r.filter_ids.push_back(0);
    r.b.n_filters = (uint32_t) r.filter_off.size() - 1;
    r.b.q_filter = r.q_filter.data(); r.b.filter_off = r.filter_off.data(); r.b.filter_ids = r.filter_ids.data();
    return &r.b;
}
thread_local std::string g_err;
// oracle threads per batch call: 1 for the scenario tests; bench.py's CPU arm sets TSGPU_DOUBLE_THREADS to the host's core count
uint32_t n_thr() { static const uint32_t n = getenv("TSGPU_DOUBLE_THREADS") ? (uint32_t) std::max(1, atoi(getenv("TSGPU_DOUBLE_THREADS"))) : 1u; return n; }
Double* D(tsgpu_index* p) { return reinterpret_cast<Double*>(p); }
}

extern "C" {

const char* tsgpu_last_error(void) { return g_err.c_str(); }
int tsgpu_device_count(void) { return 1; }   /* the double stands in for one device so the scenario program runs */

tsgpu_status tsgpu_index_create(uint32_t n_docs, int, tsgpu_index** out) {
    Double* d = new Double();
    d->n_docs = n_docs;
    d->oi = tso_index_new(n_docs);
    *out = reinterpret_cast<tsgpu_index*>(d);
    return TSGPU_OK;
}
void tsgpu_index_destroy(tsgpu_index* idx) {
    if(!idx) return;
    Double* d = D(idx);
    tso_index_free(d->oi);
    for(auto f: d->fields) delete f;
    for(auto c: d->cols) delete c;
    delete d;
}
tsgpu_status tsgpu_host_alloc(size_t bytes, void** out) { *out = malloc(bytes ? bytes : 16); return *out ? TSGPU_OK : TSGPU_ERR_CUDA; }
tsgpu_status tsgpu_host_free(void* p) { free(p); return TSGPU_OK; }
tsgpu_status tsgpu_index_load_field(tsgpu_index* idx, const tsgpu_field* f, uint32_t* out_field) {
    Double* d = D(idx);
    FieldCopy* c = new FieldCopy();
    const uint64_t n_post = f->list_off[f->n_lists];
    c->list_off.assign(f->list_off, f->list_off + f->n_lists + 1);
    c->ids.assign(f->ids, f->ids + n_post);
    c->pos_off.assign(f->pos_off, f->pos_off + n_post + 1);
    c->positions.assign(f->positions, f->positions + f->pos_off[n_post]);
    if(c->ids.empty()) c->ids.push_back(0);
    if(c->positions.empty()) c->positions.push_back(0);
    c->view.n_lists = f->n_lists; c->view.is_array = f->is_array;
    c->view.list_off = c->list_off.data(); c->view.ids = c->ids.data(); c->view.pos_off = c->pos_off.data(); c->view.positions = c->positions.data();
    d->fields.push_back(c);
    const int id = tso_index_add_field(d->oi, &c->view);
    if(out_field) *out_field = (uint32_t) id;
    return TSGPU_OK;
}
tsgpu_status tsgpu_index_append_lists(tsgpu_index* idx, uint32_t field, const tsgpu_field* f, uint32_t* out_first_list) {
    Double* d = D(idx);
    if(field >= d->fields.size()) { g_err = "no such field"; return TSGPU_ERR_INVALID; }
    FieldCopy* c = d->fields[field];
    const uint32_t L0 = c->view.n_lists;
    if(out_first_list) *out_first_list = L0;
    const uint64_t n_post0 = c->list_off[L0], n_pos0 = c->pos_off[n_post0], np = f->list_off[f->n_lists], npos = f->pos_off[np];
    c->ids.resize(n_post0); c->positions.resize(n_pos0);           // drop the placeholders of an empty field
    for(uint32_t l = 1; l <= f->n_lists; l++) c->list_off.push_back(n_post0 + f->list_off[l]);
    c->ids.insert(c->ids.end(), f->ids, f->ids + np);
    for(uint64_t i = 1; i <= np; i++) c->pos_off.push_back(n_pos0 + f->pos_off[i]);
    c->positions.insert(c->positions.end(), f->positions, f->positions + npos);
    if(c->ids.empty()) c->ids.push_back(0);
    if(c->positions.empty()) c->positions.push_back(0);
    c->view.n_lists = L0 + f->n_lists;
    c->view.list_off = c->list_off.data(); c->view.ids = c->ids.data(); c->view.pos_off = c->pos_off.data(); c->view.positions = c->positions.data();
    tso_index_set_field(d->oi, (int) field, &c->view);
    return TSGPU_OK;
}
tsgpu_status tsgpu_index_set_sort_values(tsgpu_index* idx, uint32_t sort_col, const uint32_t* ids, const int64_t* vals, size_t n) {
    Double* d = D(idx);
    if(sort_col >= d->cols.size()) { g_err = "no such sort column"; return TSGPU_ERR_INVALID; }
    for(size_t i = 0; i < n; i++) (*d->cols[sort_col])[ids[i]] = vals[i];
    return TSGPU_OK;
}
tsgpu_status tsgpu_index_load_sort_column(tsgpu_index* idx, const int64_t* vals, uint32_t* out_col) {
    Double* d = D(idx);
    auto* c = new std::vector<int64_t>(vals, vals + d->n_docs);
    d->cols.push_back(c);
    const int id = tso_index_add_sort_column(d->oi, c->data());
    if(out_col) *out_col = (uint32_t) id;
    return TSGPU_OK;
}
tsgpu_status tsgpu_index_load_hnsw(tsgpu_index* idx, const tsgpu_hnsw* g) {
    Double* d = D(idx);
    const size_t n = g->n_nodes;
    d->vec.assign(g->vectors, g->vectors + n * g->dim);
    if(g->labels) d->labels.assign(g->labels, g->labels + n);
    else { d->labels.resize(n); for(size_t i = 0; i < n; i++) d->labels[i] = (uint32_t) i; }   // tsgpu: NULL labels = identity
    d->levels.assign(g->levels, g->levels + n);
    d->links0.assign(g->links0, g->links0 + n * (2 * g->M + 1));
    d->upper_off.assign(g->upper_off, g->upper_off + n + 1);
    d->links_up.assign(g->links_up, g->links_up + (size_t) g->upper_off[n] * (g->M + 1));
    if(d->links_up.empty()) d->links_up.push_back(0);
    d->g.n_nodes = g->n_nodes; d->g.dim = g->dim; d->g.M = g->M; d->g.max_level = g->max_level; d->g.entry_point = g->entry_point; d->g.metric = g->metric;
    d->g.vectors = d->vec.data(); d->g.labels = d->labels.data(); d->g.levels = d->levels.data();
    d->g.links0 = d->links0.data(); d->g.upper_off = d->upper_off.data(); d->g.links_up = d->links_up.data();
    d->has_g = true;
    tso_index_set_hnsw(d->oi, &d->g);
    return TSGPU_OK;
}

tsgpu_status tsgpu_intersect(tsgpu_index* idx, uint32_t field, const uint32_t* lists, uint32_t k, uint32_t* out_ids, size_t cap, size_t* out_n) {
    Double* d = D(idx);
    const FieldCopy& f = *d->fields[field];
    std::vector<const uint32_t*> ptr(k); std::vector<size_t> len(k);
    for(uint32_t j = 0; j < k; j++) { ptr[j] = f.ids.data() + f.list_off[lists[j]]; len[j] = (size_t) (f.list_off[lists[j] + 1] - f.list_off[lists[j]]); }
    *out_n = tso_intersect(k, ptr.data(), len.data(), out_ids, cap);
    return TSGPU_OK;
}
tsgpu_status tsgpu_contains_atleast_one(tsgpu_index* idx, uint32_t field, uint32_t list, const uint32_t* ids, size_t n, int* out) {
    const FieldCopy& f = *D(idx)->fields[field];
    *out = tso_contains_atleast_one(f.ids.data() + f.list_off[list], (size_t) (f.list_off[list + 1] - f.list_off[list]), ids, n);
    return TSGPU_OK;
}
tsgpu_status tsgpu_phrase_matches(tsgpu_index* idx, uint32_t field, const uint32_t* lists, uint32_t k, const uint32_t* ids, size_t n, uint32_t* out_ids, size_t* out_n) {
    *out_n = tso_phrase_matches(D(idx)->oi, field, lists, k, ids, n, out_ids);
    return TSGPU_OK;
}
tsgpu_status tsgpu_exact_matches(tsgpu_index* idx, uint32_t field, const uint32_t* lists, uint32_t k, const uint32_t* ids, size_t n, uint32_t* out_ids, size_t* out_n) {
    *out_n = tso_exact_matches(D(idx)->oi, field, lists, k, ids, n, out_ids);
    return TSGPU_OK;
}
tsgpu_status tsgpu_prefix_matches(tsgpu_index* idx, uint32_t field, const uint32_t* lists, uint32_t k, const uint32_t* ids, size_t n, uint32_t* out_ids, size_t* out_n) {
    *out_n = tso_prefix_matches(D(idx)->oi, field, lists, k, ids, n, out_ids);
    return TSGPU_OK;
}
// f-1: no oracle exists for the walk other than the reference itself (tests/test_art_mirror.py); the double runs the device
// function on the host so the Python side of the GPU test executes here too
namespace {
struct ArtCopy { std::vector<tsdev::ArtNodeDev> nodes; std::vector<uint8_t> cbyte, keys; std::vector<int32_t> cref; std::vector<uint64_t> koff; int32_t root = 0; bool empty = true; };
std::map<std::pair<const tsgpu_index*, uint32_t>, ArtCopy> g_arts;
}
tsgpu_status tsgpu_index_load_art(tsgpu_index* idx, uint32_t field, const tsgpu_art* a) {
    ArtCopy c;
    c.nodes.resize(a->n_nodes);
    for(uint32_t i = 0; i < a->n_nodes; i++) {
        c.nodes[i].first_child = a->node_first_child[i]; c.nodes[i].n_children = a->node_n_children[i]; c.nodes[i].partial_len = a->node_partial_len[i];
        memcpy(c.nodes[i].partial, a->node_partial + (size_t) i * 8, 8); c.nodes[i].pad = 0;
    }
    c.cbyte.assign(a->child_byte, a->child_byte + a->n_children);
    c.cref.assign(a->child_ref, a->child_ref + a->n_children);
    if(a->n_leaves) { c.koff.assign(a->leaf_key_off, a->leaf_key_off + a->n_leaves + 1); c.keys.assign(a->leaf_keys, a->leaf_keys + c.koff.back()); }
    else c.koff = {0};
    c.root = a->root; c.empty = a->n_leaves == 0;
    g_arts[{idx, field}] = std::move(c);
    return TSGPU_OK;
}
tsgpu_status tsgpu_art_walk_batch(tsgpu_index* idx, uint32_t field, uint32_t n, const uint32_t* term_off, const uint8_t* terms, const uint8_t* min_cost,
                                  const uint8_t* max_cost, const uint8_t* prefix, int32_t* out_hits, uint32_t cap, uint32_t* out_counts, uint8_t* out_flags) {
    auto it = g_arts.find({idx, field});
    if(it == g_arts.end()) { g_err = "no ART mirror loaded for this field"; return TSGPU_ERR_INVALID; }
    const ArtCopy& c = it->second;
    tsdev::ArtDev A{c.nodes.data(), c.cbyte.data(), c.cref.data(), c.koff.data(), c.keys.data(), c.root, c.empty ? 1u : 0u};
    std::vector<tsdev::ArtFrame> stack(tsdev::kArtMaxStack);
    for(uint32_t i = 0; i < n; i++) {
        const uint32_t len = term_off[i + 1] - term_off[i];
        if(len + (prefix[i] ? 0u : 1u) > (uint32_t) tsdev::kArtMaxQuery) { out_counts[i] = 0; out_flags[i] = 2; continue; }
        tsdev::ArtQuery Q;
        memcpy(Q.q, terms + term_off[i], len);
        Q.qlen = (int) len;
        if(!prefix[i]) Q.q[Q.qlen++] = 0;
        Q.min_cost = min_cost[i]; Q.max_cost = max_cost[i]; Q.prefix = prefix[i] != 0;
        bool deep = false;
        out_counts[i] = tsdev::art_walk(A, Q, out_hits + (size_t) i * cap, cap, stack.data(), &deep);
        out_flags[i] = deep ? 1 : (out_counts[i] > cap ? 4 : 0);
    }
    return TSGPU_OK;
}
tsgpu_status tsgpu_ids_setop(tsgpu_index*, int op, const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out_ids, size_t, size_t* out_n) {
    *out_n = op == TSGPU_SET_AND ? tso_and_scalar(a, na, b, nb, out_ids) : op == TSGPU_SET_OR ? tso_or_scalar(a, na, b, nb, out_ids) : tso_exclude_scalar(a, na, b, nb, out_ids);
    return TSGPU_OK;
}
tsgpu_status tsgpu_keyword_search_batch(tsgpu_index* idx, const tsgpu_kw_batch* b, tsgpu_kv* out_kv, uint32_t kv_stride, uint32_t* out_count, uint32_t* out_found) {
    Rewritten r;
    tso_keyword_search_batch(D(idx)->oi, rewrite(D(idx), b, r), reinterpret_cast<tso_kv*>(out_kv), kv_stride, out_count, out_found, n_thr());
    return TSGPU_OK;
}
tsgpu_status tsgpu_wildcard_search_batch(tsgpu_index* idx, const tsgpu_kw_batch* b, tsgpu_kv* out_kv, uint32_t kv_stride, uint32_t* out_count, uint32_t* out_found) {
    Rewritten r;
    tso_wildcard_search_batch(D(idx)->oi, rewrite(D(idx), b, r), reinterpret_cast<tso_kv*>(out_kv), kv_stride, out_count, out_found, n_thr());
    return TSGPU_OK;
}
tsgpu_status tsgpu_vector_search_batch(tsgpu_index* idx, const tsgpu_kw_batch* b, const float* qvecs, const tsgpu_vec_params* vp, tsgpu_kv* out_kv,
                                       uint32_t kv_stride, uint32_t* out_count, uint32_t* out_found) {
    Rewritten r;
    if(tso_vector_search_batch(D(idx)->oi, rewrite(D(idx), b, r), qvecs, reinterpret_cast<const tso_vec_params*>(vp), reinterpret_cast<tso_kv*>(out_kv),
                               kv_stride, out_count, out_found, n_thr()) != 0) { g_err = "no vector index loaded"; return TSGPU_ERR_INVALID; }
    return TSGPU_OK;
}
tsgpu_status tsgpu_hybrid_search_batch(tsgpu_index* idx, const tsgpu_kw_batch* b, const float* qvecs, const tsgpu_vec_params* vp, tsgpu_kv* out_kv,
                                       uint32_t kv_stride, uint32_t* out_count, uint32_t* out_found) {
    Rewritten r;
    if(tso_hybrid_search_batch(D(idx)->oi, rewrite(D(idx), b, r), qvecs, reinterpret_cast<const tso_vec_params*>(vp), reinterpret_cast<tso_kv*>(out_kv),
                               kv_stride, out_count, out_found, n_thr()) != 0) { g_err = "no vector index loaded"; return TSGPU_ERR_INVALID; }
    return TSGPU_OK;
}
tsgpu_status tsgpu_scored_ids_search_batch(tsgpu_index* idx, const tsgpu_kw_batch* b, const int64_t* id_scores, tsgpu_kv* out_kv, uint32_t kv_stride,
                                           uint32_t* out_count, uint32_t* out_found) {
    tso_scored_ids_search_batch(D(idx)->oi, reinterpret_cast<const tso_kw_batch*>(b), id_scores, reinterpret_cast<tso_kv*>(out_kv), kv_stride, out_count, out_found, n_thr());
    return TSGPU_OK;
}
tsgpu_status tsgpu_hybrid_fuse_batch(tsgpu_index* idx, const tsgpu_kw_batch* b, const tsgpu_kv* kw_kv, uint32_t kw_stride, const uint32_t* kw_count,
                                     const uint32_t* kw_found, const uint32_t* kw_searched, const float* qvecs, const tsgpu_vec_params* vp,
                                     tsgpu_kv* out_kv, uint32_t kv_stride, uint32_t* out_count, uint32_t* out_found) {
    Rewritten r;
    if(tso_hybrid_fuse_batch(D(idx)->oi, rewrite(D(idx), b, r), reinterpret_cast<const tso_kv*>(kw_kv), kw_stride, kw_count, kw_found, kw_searched, qvecs,
                             reinterpret_cast<const tso_vec_params*>(vp), reinterpret_cast<tso_kv*>(out_kv), kv_stride, out_count, out_found, n_thr()) != 0) {
        g_err = "no vector index loaded"; return TSGPU_ERR_INVALID;
    }
    return TSGPU_OK;
}
tsgpu_status tsgpu_flat_distances(tsgpu_index* idx, const float* query, const uint32_t* ids, size_t n, float* out_dist) {
    Double* d = D(idx);
    if(!d->has_g) { g_err = "no vector index loaded"; return TSGPU_ERR_INVALID; }
    tso_flat_distances(&d->g, query, ids, n, out_dist);
    return TSGPU_OK;
}
tsgpu_status tsgpu_flat_distances_batch(tsgpu_index* idx, const float* queries, uint32_t nq, const uint32_t* ids, size_t n, float* out_dist) {
    Double* d = D(idx);
    if(!d->has_g) { g_err = "no vector index loaded"; return TSGPU_ERR_INVALID; }
    for(uint32_t q = 0; q < nq; q++) tso_flat_distances(&d->g, queries + (size_t) q * d->g.dim, ids, n, out_dist + (size_t) q * n);
    return TSGPU_OK;
}
tsgpu_status tsgpu_filter_create(tsgpu_index* idx, const uint32_t* ids, size_t n, int32_t* out_handle) {
    Double* d = D(idx);
    d->filters.emplace_back(ids, ids + n);
    *out_handle = -((int32_t) d->filters.size() - 1 + 2);
    return TSGPU_OK;
}
// filter_by leaves / tree answered on the host (the comparators of src/num_tree.cpp over the dense column)
tsgpu_status tsgpu_filter_numeric(tsgpu_index* idx, uint32_t col, int op, int64_t v1, int64_t v2, int32_t* out_handle, size_t* out_n) {
    Double* d = D(idx);
    if(col >= d->cols.size()) { g_err = "column out of range"; return TSGPU_ERR_INVALID; }
    const std::vector<int64_t>& c = *d->cols[col];
    std::vector<uint32_t> ids;
    for(uint32_t i = 0; i < d->n_docs; i++) {
        const int64_t v = c[i];
        const bool has = v != INT64_MIN;
        bool m;
        switch(op) {
            case 0: m = has && v == v1; break;
            case 1: m = !(has && v == v1); break;
            case 2: m = has && v < v1; break;
            case 3: m = has && v <= v1; break;
            case 4: m = has && v > v1; break;
            case 5: m = has && v >= v1; break;
            default: m = has && v >= v1 && v <= v2; break;
        }
        if(m) ids.push_back(i);
    }
    if(out_n) *out_n = ids.size();
    d->filters.push_back(std::move(ids));
    *out_handle = -((int32_t) d->filters.size() - 1 + 2);
    return TSGPU_OK;
}
tsgpu_status tsgpu_filter_combine(tsgpu_index* idx, int op, int32_t a, int32_t b, int32_t* out_handle, size_t* out_n) {
    Double* d = D(idx);
    const size_t ia = (size_t) (-(a + 2)), ib = (size_t) (-(b + 2));
    if(a > -2 || b > -2 || ia >= d->filters.size() || ib >= d->filters.size()) { g_err = "unknown filter handle"; return TSGPU_ERR_INVALID; }
    const auto& A = d->filters[ia]; const auto& B = d->filters[ib];
    std::vector<uint32_t> out(A.size() + B.size() + 1);
    size_t n = op == 0 ? tso_and_scalar(A.data(), A.size(), B.data(), B.size(), out.data())
             : op == 1 ? tso_or_scalar(A.data(), A.size(), B.data(), B.size(), out.data())
                       : tso_exclude_scalar(A.data(), A.size(), B.data(), B.size(), out.data());
    out.resize(n);
    if(out_n) *out_n = n;
    d->filters.push_back(std::move(out));
    *out_handle = -((int32_t) d->filters.size() - 1 + 2);
    return TSGPU_OK;
}
tsgpu_status tsgpu_filter_ids(tsgpu_index* idx, int32_t handle, uint32_t* out_ids, size_t cap, size_t* out_n) {
    Double* d = D(idx);
    const size_t h = (size_t) (-(handle + 2));
    if(handle > -2 || h >= d->filters.size()) { g_err = "bad filter handle"; return TSGPU_ERR_INVALID; }
    *out_n = d->filters[h].size();
    if(*out_n > cap) { g_err = "output buffer too small"; return TSGPU_ERR_CAPACITY; }
    std::copy(d->filters[h].begin(), d->filters[h].end(), out_ids);
    return TSGPU_OK;
}
tsgpu_status tsgpu_filter_destroy(tsgpu_index* idx, int32_t handle) {
    Double* d = D(idx);
    const size_t h = (size_t) (-(handle + 2));
    if(handle > -2 || h >= d->filters.size()) { g_err = "bad filter handle"; return TSGPU_ERR_INVALID; }
    d->filters[h].clear();
    return TSGPU_OK;
}
tsgpu_status tsgpu_get_stats(tsgpu_index*, tsgpu_stats* out) { memset(out, 0, sizeof(*out)); return TSGPU_OK; }
tsgpu_status tsgpu_index_build_hnsw(tsgpu_index*, const float*, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, int) {
    g_err = "the test double has no device build"; return TSGPU_ERR_NO_DEVICE;
}
tsgpu_status tsgpu_index_hnsw_info(tsgpu_index* idx, uint32_t* n_nodes, uint32_t* dim, uint32_t* M, uint32_t* max_level, uint32_t* entry_point,
                                   uint64_t* n_upper_records, uint64_t* build_counters) {
    Double* d = D(idx);
    if(!d->has_g) { g_err = "no vector index loaded"; return TSGPU_ERR_INVALID; }
    if(n_nodes) *n_nodes = d->g.n_nodes;
    if(dim) *dim = d->g.dim;
    if(M) *M = d->g.M;
    if(max_level) *max_level = d->g.max_level;
    if(entry_point) *entry_point = d->g.entry_point;
    if(n_upper_records) *n_upper_records = d->g.n_nodes ? d->upper_off[d->g.n_nodes] : 0;
    if(build_counters) for(int i = 0; i < 5; i++) build_counters[i] = 0;
    return TSGPU_OK;
}
tsgpu_status tsgpu_index_append_hnsw(tsgpu_index*, const float*, uint32_t, uint32_t, uint32_t, uint32_t) { g_err = "the test double has no device build"; return TSGPU_ERR_NO_DEVICE; }
tsgpu_status tsgpu_index_mark_deleted(tsgpu_index*, const uint32_t*, size_t, int) { g_err = "the test double has no deletion marks"; return TSGPU_ERR_NO_DEVICE; }
tsgpu_status tsgpu_index_export_hnsw(tsgpu_index*, uint8_t*, uint32_t*, uint64_t*, uint32_t*) { g_err = "the test double has no device build"; return TSGPU_ERR_NO_DEVICE; }
namespace { struct FacetCopy { uint32_t n_values; std::vector<uint64_t> off; std::vector<uint32_t> vals; }; std::vector<FacetCopy*> g_facets; }
tsgpu_status tsgpu_index_load_facet(tsgpu_index* idx, const tsgpu_facet* f, uint32_t* out_facet) {
    const size_t n = D(idx)->n_docs;
    FacetCopy* c = new FacetCopy;
    c->n_values = f->n_values;
    c->off.assign(f->doc_off, f->doc_off + n + 1);
    c->vals.assign(f->value_ids, f->value_ids + c->off[n]);
    if(c->vals.empty()) c->vals.push_back(0);
    *out_facet = (uint32_t) g_facets.size();
    g_facets.push_back(c);
    return TSGPU_OK;
}
tsgpu_status tsgpu_facet_counts(tsgpu_index* idx, uint32_t facet, const uint32_t* result_ids, size_t n, uint32_t sample_mod, uint32_t top_n,
                                tsgpu_facet_count* out, uint32_t* out_n, uint32_t* out_distinct) {
    static_assert(sizeof(tsgpu_facet_count) == sizeof(tso_facet_count), "facet count layout");
    const FacetCopy& c = *g_facets[facet];
    *out_n = (uint32_t) tso_facet_counts(D(idx)->n_docs, c.n_values, c.off.data(), c.vals.data(), result_ids, n, sample_mod,
                                         reinterpret_cast<tso_facet_count*>(out), top_n, out_distinct);
    return TSGPU_OK;
}
tsgpu_status tsgpu_facet_counts_last(tsgpu_index*, uint32_t, uint32_t, tsgpu_facet_count*, uint32_t*, uint32_t*) { g_err = "the test double keeps no all_result_ids"; return TSGPU_ERR_NO_DEVICE; }
tsgpu_status tsgpu_all_result_ids_last(tsgpu_index*, uint32_t, uint32_t*, size_t, size_t*) { g_err = "the test double keeps no all_result_ids"; return TSGPU_ERR_NO_DEVICE; }
tsgpu_status tsgpu_comm_unique_id(void*) { g_err = "the test double has no communicator"; return TSGPU_ERR_NO_DEVICE; }
tsgpu_status tsgpu_comm_init(tsgpu_index*, int, int, const void*) { g_err = "the test double has no communicator"; return TSGPU_ERR_NO_DEVICE; }
tsgpu_status tsgpu_comm_destroy(tsgpu_index*) { return TSGPU_OK; }
tsgpu_status tsgpu_comm_gather(tsgpu_index*, const void*, size_t, void*, int) { g_err = "the test double has no communicator"; return TSGPU_ERR_NO_DEVICE; }
tsgpu_status tsgpu_comm_last_ms(tsgpu_index*, float* out) { *out = 0; return TSGPU_OK; }
tsgpu_status tsgpu_debug_knn_work(tsgpu_index*, uint32_t*, uint32_t, uint32_t* out_n) { *out_n = 0; return TSGPU_OK; }
tsgpu_status tsgpu_knn_batch(tsgpu_index* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef, const int32_t* q_filter, uint32_t,
                             const uint64_t* filter_off, const uint32_t* filter_ids, float* out_dist, uint32_t* out_labels, uint32_t* out_n) {
    Double* d = D(idx);
    if(!d->has_g) { g_err = "no vector index loaded"; return TSGPU_ERR_INVALID; }
    uint64_t stats[2] = {0, 0};
    tso_hnsw_search_batch(&d->g, queries, nq, k, ef, q_filter, filter_off, filter_ids, out_dist, out_labels, out_n, stats, n_thr());
    return TSGPU_OK;
}

}  // extern "C"
