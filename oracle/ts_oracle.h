/* TEST INFRASTRUCTURE — parity oracle for the Typesense query hot path. NOT part of the product.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library,
 * and only as the checker / the CPU baseline. The product (libtsgpu.so) never links or calls it.
 *
 * This is a standalone CPU restatement (plain C++17, flat arrays) of the reference algorithm for the path
 * SURVEY.md §8 names. Every function cites the reference file:line it follows. Where the reference's own sources
 * compile here (posting_list/or_iterator/Match — oracle/_ref), tests/test_oracle_ref.py pins this restatement
 * against them; the remaining pieces are pinned against the reference's own unit-test vectors (tests/golden/).
 *
 * The structs below mirror the layout of include/tsgpu.h on purpose (same field order and types) so the same
 * numpy buffers can be handed to both libraries in the parity tests; the two headers are maintained separately.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSO_NO_LIST 0xFFFFFFFFu
#define TSO_MAX_FIELDS 8
#define TSO_MAX_TOKENS 16

/* one searchable string field, flattened: token t owns postings [list_off[t], list_off[t+1]) */
typedef struct {
    uint32_t n_lists;
    uint32_t is_array;            /* string[] field (offset encoding src/index.cpp:1384-1393) */
    const uint64_t* list_off;     /* [n_lists+1] */
    const uint32_t* ids;          /* ascending seq_ids per list */
    const uint64_t* pos_off;      /* [n_postings+1] -> positions */
    const uint32_t* positions;    /* reference-encoded offsets (src/index.cpp:1341-1348, 1384-1393) */
} tso_field;

/* result record: scalar fields of the reference's KV (include/topster.h:20-33) */
typedef struct {
    uint64_t key;
    uint64_t distinct_key;
    int64_t  scores[3];
    int64_t  text_match_score;
    float    vector_distance;
    int8_t   match_score_index;
    uint8_t  pad0;
    uint16_t query_index;
} tso_kv;   /* 56 bytes */

enum { TSO_SORT_NONE = 0, TSO_SORT_TEXT_MATCH = 1, TSO_SORT_SEQ_ID = 2, TSO_SORT_NUMERIC = 3, TSO_SORT_VECTOR_DISTANCE = 4 };
enum { TSO_MATCH_MAX_SCORE = 0, TSO_MATCH_MAX_WEIGHT = 1, TSO_MATCH_SUM_SCORE = 2 };
enum { TSO_FLAG_PRIORITIZE_EXACT_MATCH = 1, TSO_FLAG_PRIORITIZE_TOKEN_POSITION = 2, TSO_FLAG_PRIORITIZE_NUM_MATCHING_FIELDS = 4,
       TSO_FLAG_RERANK_HYBRID_MATCHES = 0x40 /* hybrid calls: Index::compute_aux_scores after the fusion */ };
enum { TSO_CFLAG_SYNONYM = 1, TSO_CFLAG_DEMOTE_SYNONYM = 2 };

/* A batch of keyword searches; query q = one Index::search_all_candidates call (src/index.cpp:1794), i.e. a list of
 * token combinations all scored into the same Topster. */
typedef struct {
    uint32_t n_queries;
    uint32_t n_combos;
    uint32_t n_fields;                 /* F: searched fields, slot f -> index field field_ids[f] */
    uint32_t n_filters;
    const uint32_t* field_ids;         /* [F] */
    /* per query */
    const uint32_t* q_combo_off;       /* [nq+1] */
    const int32_t*  q_filter;          /* [nq] filter slot or -1 */
    const uint32_t* q_excl_off;        /* [nq+1] -> excl_ids */
    const uint32_t* excl_ids;          /* sorted per query */
    const uint32_t* q_topk;            /* [nq] Topster capacity (src/index.cpp:3506-3512) */
    const uint8_t*  q_sort_type;       /* [nq*3] */
    const int32_t*  q_sort_col;        /* [nq*3] sort column id for TSO_SORT_NUMERIC */
    const int8_t*   q_sort_order;      /* [nq*3] +1 desc, -1 asc */
    const uint8_t*  q_sort_missing_first; /* [nq*3] */
    const uint8_t*  q_flags;           /* [nq] TSO_FLAG_* */
    const uint8_t*  q_match_type;      /* [nq] */
    const uint8_t*  q_num_query_tokens;/* [nq] query_tokens.size() as seen by compute_aggregated_score */
    const uint8_t*  q_field_weight;    /* [nq*F] */
    /* per combo */
    const uint32_t* c_tok_off;         /* [nc+1] -> token rows */
    const uint32_t* c_total_cost;      /* [nc] */
    const uint8_t*  c_n_required;      /* [nc] first n tokens are ANDed, the rest are dropped (optional) tokens */
    const uint8_t*  c_flags;           /* [nc] TSO_CFLAG_* */
    const int32_t*  c_syn_orig_num_tokens; /* [nc] -1 default */
    const int32_t*  c_orig_num_tokens; /* [nc] */
    /* per token row */
    const uint32_t* t_list;            /* [n_rows*F] posting list id within field slot f, TSO_NO_LIST if absent */
    /* filters: sorted seq_id arrays */
    const uint64_t* filter_off;        /* [n_filters+1] */
    const uint32_t* filter_ids;
} tso_kw_batch;

typedef struct {
    uint32_t n_nodes;        /* internal ids 0..n-1 */
    uint32_t dim;
    uint32_t M;              /* links per node above level 0; level 0 holds 2M */
    uint32_t max_level;
    uint32_t entry_point;    /* internal id, 0xFFFFFFFF if empty */
    uint32_t metric;         /* 0 ip, 1 cosine (vectors already normalised at index time) */
    const float*    vectors; /* [n*dim] by internal id */
    const uint32_t* labels;  /* [n] internal id -> seq_id */
    const uint8_t*  levels;  /* [n] */
    const uint32_t* links0;  /* [n*(2M+1)] count + neighbours */
    const uint64_t* upper_off; /* [n+1] -> links_up, in units of (M+1) u32 records, one record per level 1..levels[i] */
    const uint32_t* links_up;
} tso_hnsw;

/* ---- posting lists on flat arrays */
/* Index::do_facets hash-index branch (src/index.cpp:1674-1780) + Collection::search's (count, id) order; returns entries written */
typedef struct { uint32_t value_id, count, doc_id, array_pos; } tso_facet_count;
size_t tso_facet_counts(uint32_t n_docs, uint32_t n_values, const uint64_t* doc_off, const uint32_t* value_ids,
                        const uint32_t* result_ids, size_t n, uint32_t sample_mod, tso_facet_count* out, size_t cap, uint32_t* out_distinct);
/* posting_list_t::contains_atleast_one (src/posting_list.cpp:1090-1112) on an ascending id list */
int tso_contains_atleast_one(const uint32_t* list, size_t n_list, const uint32_t* target_ids, size_t n_targets);
size_t tso_intersect(uint32_t k, const uint32_t* const* lists, const size_t* lens, uint32_t* out, size_t cap);
size_t tso_merge(uint32_t k, const uint32_t* const* lists, const size_t* lens, uint32_t* out, size_t cap);
size_t tso_and_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out);
size_t tso_or_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out);
size_t tso_exclude_scalar(const uint32_t* a, size_t na, const uint32_t* b, size_t nb, uint32_t* out);

/* ---- Match */
void tso_match(uint32_t n_tokens, const uint32_t* tok_off, const uint16_t* positions, const uint8_t* last_token,
               int check_exact, uint8_t out[4]);
uint64_t tso_match_score(uint8_t words_present, uint8_t distance, uint8_t max_offset, uint8_t exact,
                         uint32_t total_cost, uint32_t unique_words, uint8_t syn);
int tso_has_phrase_match(uint32_t n_tokens, const uint32_t* tok_off, const uint16_t* positions);

/* ---- index object (borrows all arrays) */
void* tso_index_new(uint32_t n_docs);
void  tso_index_free(void* idx);
int   tso_index_add_field(void* idx, const tso_field* f);                    /* returns field id */
void  tso_index_set_field(void* idx, int field, const tso_field* f);          /* the field's arrays were replaced (incremental mirror tests) */
int   tso_index_add_sort_column(void* idx, const int64_t* vals);             /* [n_docs], INT64_MIN = missing */
void  tso_index_set_hnsw(void* idx, const tso_hnsw* g);

/* ---- one token combination, unsorted stream of (id, aggregated score) — mirrors ref_keyword_combo */
size_t tso_keyword_combo(void* idx, const tso_kw_batch* b, uint32_t q, uint32_t c, uint32_t* out_ids,
                         uint64_t* out_scores, size_t cap, uint64_t* out_num_keyword_matches);

/* ---- batched keyword search (search_all_candidates + Topster), n_threads worker threads.
 * out_kv [nq*kv_stride], out_count[nq] valid kvs (sorted, Topster::sort order), out_found[nq] = |all_result_ids| */
int tso_keyword_search_batch(void* idx, const tso_kw_batch* b, tso_kv* out_kv, uint32_t kv_stride,
                             uint32_t* out_count, uint32_t* out_found, uint32_t n_threads);

/* ---- wildcard query (q=*): Index::search_wildcard (src/index.cpp:6616-6800): every filter id (all docs when the query
 * has no filter) minus the exclusion list gets sort scores with text-match value 100 and goes into the Topster.
 * Combinations of the batch are ignored. query_index is left 0 (the reference reads searched_queries.size() from worker
 * threads while the enqueuing thread is still appending to it, i.e. it is not deterministic there). */
int tso_scored_ids_search_batch(void* idx, const tso_kw_batch* b, const int64_t* id_scores, tso_kv* out_kv, uint32_t kv_stride,
                                uint32_t* out_count, uint32_t* out_found, uint32_t n_threads);
int tso_wildcard_search_batch(void* idx, const tso_kw_batch* b, tso_kv* out_kv, uint32_t kv_stride,
                              uint32_t* out_count, uint32_t* out_found, uint32_t n_threads);

/* ---- Topster fed with an explicit stream (pins include/topster.h against test/topster_test.cpp) */
uint32_t tso_topster_run(uint32_t capacity, const tso_kv* in, uint32_t n, tso_kv* out);

/* ---- phrase search over an id set (posting_list_t::get_phrase_matches, src/posting_list.cpp:1791) */
size_t tso_phrase_matches(void* idx, uint32_t field, const uint32_t* lists, uint32_t k,
                          const uint32_t* ids, size_t n, uint32_t* out);

/* ---- `:=` / prefix string-filter checks over an id set (posting_list_t::get_exact_matches, src/posting_list.cpp:1281;
 *      get_prefix_matches, :1129) */
size_t tso_exact_matches(void* idx, uint32_t field, const uint32_t* lists, uint32_t k,
                         const uint32_t* ids, size_t n, uint32_t* out);
size_t tso_prefix_matches(void* idx, uint32_t field, const uint32_t* lists, uint32_t k,
                          const uint32_t* ids, size_t n, uint32_t* out);

/* ---- vectors */
float tso_ip_distance(const float* a, const float* b, uint32_t dim);
void  tso_normalize(const float* src, float* dst, uint32_t dim);
/* hnswlib-equivalent construction (single thread, seed, M, ef_construction); arrays sized by the caller:
 * levels[n], links0[n*(2M+1)], upper links returned through tso_hnsw_build_fetch */
void* tso_hnsw_build(const float* vectors, uint32_t n, uint32_t dim, uint32_t M, uint32_t ef_construction,
                     uint32_t seed);
void  tso_hnsw_build_info(void* bld, uint32_t* max_level, uint32_t* entry_point, uint64_t* n_upper_records);
void  tso_hnsw_build_fetch(void* bld, uint8_t* levels, uint32_t* links0, uint64_t* upper_off, uint32_t* links_up);
void  tso_hnsw_build_free(void* bld);

/* searchKnnCloserFirst(q, k, ef, filter) (src/index.cpp:3384). filter: sorted allowed labels (NULL = none),
 * excl: sorted excluded labels. Outputs closest-first; returns count. stats[0]+=n_dist, stats[1]+=n_expanded */
uint32_t tso_hnsw_search(const tso_hnsw* g, const float* q, uint32_t k, uint32_t ef,
                         const uint32_t* filter, size_t n_filter, const uint32_t* excl, size_t n_excl,
                         float* out_dist, uint32_t* out_labels, uint64_t* stats);
/* batch of nq queries over worker threads; filter slot per query (-1 none) */
void tso_hnsw_search_batch(const tso_hnsw* g, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                           const int32_t* q_filter, const uint64_t* filter_off, const uint32_t* filter_ids,
                           float* out_dist, uint32_t* out_labels, uint32_t* out_n, uint64_t* stats, uint32_t n_threads);
/* process_results_bruteforce (src/index.cpp:3345): distances of all filter ids, in filter order */
void tso_flat_distances(const tso_hnsw* g, const float* q, const uint32_t* ids, size_t n, float* out_dist);

/* ---- vector-only and hybrid result assembly (src/index.cpp:3645-3732, 4036-4221) */
typedef struct {
    uint32_t k;                 /* vector_query.k (0 = default) */
    uint32_t ef;
    uint32_t flat_search_cutoff;
    float    distance_threshold;
    float    alpha;
    uint32_t fetch_size;
} tso_vec_params;

/* hybrid batch: keyword batch + one vector per query; same output contract as tso_keyword_search_batch */
int tso_hybrid_search_batch(void* idx, const tso_kw_batch* b, const float* qvecs, const tso_vec_params* vp,
                            tso_kv* out_kv, uint32_t kv_stride, uint32_t* out_count, uint32_t* out_found,
                            uint32_t n_threads);
/* pure vector search batch (wildcard q + vector_query): sort clauses / filter / exclusion from the kw batch
 * (combos unused) */
int tso_hybrid_fuse_batch(void* idx, const tso_kw_batch* b, const tso_kv* kw_kv, uint32_t kw_stride, const uint32_t* kw_count, const uint32_t* kw_found,
                          const uint32_t* kw_searched, const float* qvecs, const tso_vec_params* vp, tso_kv* out_kv, uint32_t kv_stride,
                          uint32_t* out_count, uint32_t* out_found, uint32_t n_threads);
int tso_vector_search_batch(void* idx, const tso_kw_batch* b, const float* qvecs, const tso_vec_params* vp,
                            tso_kv* out_kv, uint32_t kv_stride, uint32_t* out_count, uint32_t* out_found,
                            uint32_t n_threads);

int64_t tso_float_to_int64(float f);
float   tso_int64_to_float(int64_t v);

#ifdef __cplusplus
}
#endif
