// C++ host-side scenarios over the tsgpu C-ABI, written the way the reference's own gtest cases read
// (test/posting_list_test.cpp, test/or_iterator_test.cpp, test/collection_test.cpp, test/collection_vector_search_test.cpp).
// Built and run by tests/test_cpp_host.py on a machine with a GPU; exit code = number of failed checks.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../typesense_b200/host/tsgpu_host.hpp"
#include "../../oracle/ts_oracle.h"      // test infrastructure: only used to build an HNSW graph for the vector case

static int failures = 0;
#define CHECK(cond) do { if(!(cond)) { failures++; printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } } while(0)

static std::string json_str(const std::string& line, const std::string& key) {
    size_t p = line.find("\"" + key + "\"");
    if(p == std::string::npos) return "";
    p = line.find(':', p);
    p = line.find('"', p);
    size_t e = line.find('"', p + 1);
    return line.substr(p + 1, e - p - 1);
}
static long json_int(const std::string& line, const std::string& key) {
    size_t p = line.find("\"" + key + "\"");
    p = line.find(':', p);
    return std::strtol(line.c_str() + p + 1, nullptr, 10);
}

// TEST(PostingListTest, IntersectionBasics) test/posting_list_test.cpp:603-700
static void posting_list_intersection_basics() {
    tsgpu::Index index(64);
    tsgpu::field_mirror_t f;
    const std::vector<uint32_t> offsets = {0, 1, 3};
    for(uint32_t i: {0u, 2u, 3u, 20u}) f.upsert("p1", i, offsets);
    for(uint32_t i: {1u, 3u, 5u, 10u, 20u}) f.upsert("p2", i, offsets);
    for(uint32_t i: {2u, 3u, 5u, 7u, 20u}) f.upsert("p3", i, offsets);
    CHECK(index.add_field("f", f).ok());
    std::vector<uint32_t> result_ids;
    CHECK(index.intersect("f", {"p1", "p2", "p3"}, result_ids).ok());
    CHECK((result_ids == std::vector<uint32_t>{3, 20}));
    CHECK(index.intersect("f", {"p1"}, result_ids).ok());
    CHECK((result_ids == std::vector<uint32_t>{0, 2, 3, 20}));
    CHECK(index.intersect("f", {"p1", "missing"}, result_ids).ok());
    CHECK(result_ids.empty());
}

// TEST(OrIteratorTest, IntersectAndFilterThreeIts) test/or_iterator_test.cpp:162-217
static void or_iterator_intersect_and_filter() {
    tsgpu::Index index(100000);
    tsgpu::field_mirror_t f;
    const std::vector<uint32_t> offsets = {1, 0};
    for(uint32_t i: {4207u, 29159u, 47182u, 47250u, 47337u, 48518u, 99820u}) f.upsert("a", i, offsets);
    for(uint32_t i: {62u, 330u, 367u, 4124u, 4207u, 4242u, 4418u, 28740u, 29099u, 29159u, 29284u, 40795u, 43556u, 46779u, 47182u, 47250u, 47322u,
                     48494u, 48518u, 48633u, 98813u, 98821u, 99069u, 99368u, 99533u, 99670u, 99820u, 99888u, 99973u}) f.upsert("b", i, offsets);
    for(uint32_t i: {723u, 1504u, 29038u, 29164u, 29390u, 30890u, 34743u, 35067u, 36466u, 40268u, 40965u, 42161u, 43425u, 45188u, 47326u, 47443u,
                     49319u, 53043u, 58436u, 58774u, 61123u, 70973u, 71393u, 81575u, 82323u, 88301u, 88502u, 88594u, 88690u, 88951u, 90662u, 91016u,
                     91915u, 92069u, 92844u, 99820u}) f.upsert("c", i, offsets);
    CHECK(index.add_field("f", f).ok());
    tsgpu::host_topster_t topster(250);
    size_t found = 0;
    const std::vector<uint32_t> filter_ids = {44424, 44425, 44447, 99820, 99834, 99854, 99859, 99963};
    auto op = index.search_across_fields({{"a", "b", "c"}}, 0, {0}, {"f"}, {15}, {{tsgpu::sort_by::text_match, "", true}}, filter_ids, true, {}, 250,
                                         true, topster, found);
    CHECK(op.ok());
    auto kvs = topster.sort();
    CHECK(kvs.size() == 1 && found == 1);
    if(!kvs.empty()) CHECK(kvs[0].key == 99820);
}

// Collection::search(q, fields, "", facets, sort, {num_typos}, per_page, page, token_order, {prefix}, drop_tokens_threshold, ...,
// typo_tokens_threshold): the switches each reference test passes
static tsgpu::search_options opt(uint32_t num_typos, bool prefix, size_t typo_tokens_threshold = 1,
                                 tsgpu::search_options::token_ordering order = tsgpu::search_options::FREQUENCY) {
    tsgpu::search_options o;
    o.num_typos = num_typos; o.prefix = prefix; o.typo_tokens_threshold = typo_tokens_threshold; o.token_order = order;
    o.device_art_walk = getenv("TSGPU_HOST_DEVICE_ART") != nullptr;      // f-1 opt-in: candidate walks through tsgpu_art_walk_batch
    return o;
}

// TEST_F(CollectionTest, MultiTokenSearch / ExactSearchShouldBeStable) test/collection_test.cpp:117-236
static void collection_scenarios(const std::string& jsonl) {
    std::vector<std::pair<std::string, long>> docs = {{"z", 10}};        // dummy record for id 0
    std::vector<std::string> ext_ids = {"0"};
    std::ifstream in(jsonl);
    std::string line;
    while(std::getline(in, line)) {
        if(line.empty()) continue;
        docs.emplace_back(json_str(line, "title"), json_int(line, "points"));
        std::string id = json_str(line, "id");
        ext_ids.push_back(id.empty() ? std::to_string(docs.size() - 1) : id);
    }
    CHECK(docs.size() == 25);        // 24 lines + the dummy record
    tsgpu::Index index((uint32_t) docs.size());
    tsgpu::field_mirror_t title;
    std::unordered_map<uint32_t, int64_t> points;
    for(uint32_t i = 0; i < docs.size(); i++) { title.index_plain_string(i, tsgpu::tokenize_ascii(docs[i].first)); points[i] = docs[i].second; }
    CHECK(index.add_field("title", title).ok());
    CHECK(index.add_sort_field("points", points).ok());
    auto ids_of = [&](const std::vector<tsgpu::KV>& kvs) { std::vector<std::string> r; for(auto& kv: kvs) r.push_back(ext_ids[kv.key]); return r; };
    std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "points", true}};
    std::vector<tsgpu::KV> kvs;
    size_t found = 0;
    CHECK(index.search(tsgpu::tokenize_ascii("rocket launch"), {"title"}, sort_fields, 10, 250, kvs, found, opt(0, false)).ok());
    CHECK((ids_of(kvs) == std::vector<std::string>{"8", "1", "17", "16", "13"}));
    CHECK(found == 5);
    std::vector<tsgpu::sort_by> sort_fields_asc = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "points", false}};
    CHECK(index.search(tsgpu::tokenize_ascii("rocket launch"), {"title"}, sort_fields_asc, 10, 250, kvs, found, opt(0, false)).ok());
    CHECK((ids_of(kvs) == std::vector<std::string>{"8", "17", "1", "16", "13"}));
    CHECK(index.search(tsgpu::tokenize_ascii("the"), {"title"}, sort_fields, 1, 250, kvs, found, opt(0, false)).ok());
    CHECK((ids_of(kvs) == std::vector<std::string>{"1", "6", "foo", "13", "10", "8", "16"}));
    CHECK(found == 7);
    CHECK(index.search(tsgpu::tokenize_ascii("zxsadqewsad"), {"title"}, sort_fields, 1, 250, kvs, found, opt(0, false)).ok());
    CHECK(kvs.empty() && found == 0);
    // PartialMultiTokenSearch :358-372, SkipUnindexedTokensDuringMultiTokenSearch :269-356, SearchWithExcludedTokens :238-267
    CHECK(index.search(tsgpu::tokenize_ascii("rocket research"), {"title"}, sort_fields, 10, 250, kvs, found, opt(0, false)).ok());
    CHECK((ids_of(kvs) == std::vector<std::string>{"19", "1", "10", "8", "16", "17"}));
    CHECK(index.search(tsgpu::tokenize_ascii("DoesNotExist from"), {"title"}, sort_fields, 1, 250, kvs, found, opt(0, true)).ok());
    CHECK((ids_of(kvs) == std::vector<std::string>{"2", "17"}));
    CHECK(index.search(tsgpu::tokenize_ascii("the a"), {"title"}, sort_fields, 10, 250, kvs, found, opt(0, false, 10)).ok());
    CHECK(kvs.size() == 9);
    CHECK(index.search(tsgpu::tokenize_ascii("the a"), {"title"}, sort_fields, 0, 250, kvs, found, opt(0, false)).ok());
    CHECK((ids_of(kvs) == std::vector<std::string>{"8", "16", "10"}));
    CHECK(index.search(tsgpu::tokenize_ascii("the a insurance"), {"title"}, sort_fields, 0, 250, kvs, found, opt(0, false)).ok());
    CHECK(kvs.empty());
    {
        tsgpu::search_options o = opt(0, false, 10);
        o.exclude_tokens = {"propellants", "are"};
        CHECK(index.search(tsgpu::tokenize_ascii("how"), {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((ids_of(kvs) == std::vector<std::string>{"9", "17"}) && found == 2);
    }
    // QueryWithTypo :374, TypoTokenRankedByScoreAndFrequency :413, PrefixSearching :605, TypoTokensThreshold :686
    {
        using O = tsgpu::search_options;
        struct Case { const char* q; tsgpu::search_options o; size_t drop; std::vector<std::string> expect; size_t per_page; long found; };
        const std::vector<Case> cases = {
            {"kind biologcal", opt(2, false, 10), 10, {"19", "3", "20"}, 3, -1},
            {"lauxnch rcket", opt(1, false, 10), 10, {"8", "1", "17"}, 3, -1},
            {"loox", opt(1, false, 1, O::MAX_SCORE), 1, {"22", "3"}, 2, 5},
            {"loox", opt(1, false, 1, O::FREQUENCY), 1, {"22", "3", "12", "23", "24"}, 10, 5},
            {"loox", opt(1, false, 1, O::MAX_SCORE), 1, {"22", "3", "12", "23", "24"}, 10, 5},
            {"ex", opt(0, true, 1, O::FREQUENCY), 1, {"6", "12"}, 10, 2},
            {"ex", opt(0, true, 1, O::MAX_SCORE), 1, {"6", "12"}, 10, 2},
            {"what ex", opt(0, true, 10, O::MAX_SCORE), 10, {"6", "12", "19", "22", "13", "8", "15", "24", "21"}, 10, 9},
            {"t", opt(0, true, 10, O::MAX_SCORE), 10, {"19", "22"}, 2, -1},
            {"t", opt(0, true, 10, O::FREQUENCY), 10, {"1", "2"}, 2, -1},
            {"ISSX what", opt(1, false, 20), 20, {"19", "6", "21", "22"}, 4, 11},        // TextContainingAnActualTypo :473-508
            {"ISSX", opt(1, false, 10), 10, {"20", "19", "6", "3", "21"}, 10, 5},
            {"math fx", opt(0, true), 0, {}, 1, 0},
            {"x", opt(2, true), 1, {}, 2, 0},
            {"late propx", opt(2, true), 1, {"16"}, 1, -1},
        };
        for(auto& c: cases) {
            CHECK(index.search(tsgpu::tokenize_ascii(c.q), {"title"}, sort_fields, c.drop, 250, kvs, found, c.o).ok());
            auto ids = ids_of(kvs);
            if(ids.size() > c.per_page) ids.resize(c.per_page);
            if(!(ids == c.expect)) { printf("typo case '%s':", c.q); for(auto& i: ids) printf(" %s", i.c_str()); printf("\n"); }
            CHECK(ids == c.expect);
            if(c.found >= 0) CHECK((long) found == c.found);
        }
        // the same searches as ONE multi_search (src/core_api.cpp:1080): identical answers; with the device walk switched on all of
        // their candidate walks are fetched ahead in one launch
        {
            std::vector<tsgpu::Index::search_request> reqs;
            for(auto& c: cases) reqs.push_back({tsgpu::tokenize_ascii(c.q), {"title"}, sort_fields, c.drop, 250, c.o});
            index.clear_walk_cache();
            const uint64_t launches0 = tsgpu::Index::art_walk_stats().launches, searches0 = tsgpu::Index::art_walk_stats().searches;
            const uint64_t calls0 = tsgpu::Index::kw_device_calls();
            auto seq = index.multi_search(reqs, /*in_lockstep=*/false);
            const uint64_t calls_seq = tsgpu::Index::kw_device_calls() - calls0;
            auto resps = index.multi_search(reqs);                        // lock-step: pending queries of all searches share a device call
            const uint64_t calls_lock = tsgpu::Index::kw_device_calls() - calls0 - calls_seq;
            {   // a request naming an unknown field fails alone; the others of the list are answered
                auto bad = reqs;
                bad[1].the_fields = {"no_such_field"};
                auto r2 = index.multi_search(bad);
                CHECK(!r2[1].status.ok() && r2[0].status.ok() && r2[2].status.ok() && ids_of(r2[0].raw_result_kvs) == ids_of(resps[0].raw_result_kvs));
            }
            printf("multi_search of %zu searches: %llu keyword device calls one by one, %llu in lock-step\n", reqs.size(),
                   (unsigned long long) calls_seq, (unsigned long long) calls_lock);
            CHECK(calls_lock * 3 <= calls_seq && seq.size() == resps.size());
            for(size_t i = 0; i < seq.size(); i++) {
                CHECK(seq[i].found == resps[i].found && ids_of(seq[i].raw_result_kvs) == ids_of(resps[i].raw_result_kvs));
                for(size_t k = 0; k < seq[i].raw_result_kvs.size() && k < resps[i].raw_result_kvs.size(); k++)
                    CHECK(seq[i].raw_result_kvs[k].scores[0] == resps[i].raw_result_kvs[k].scores[0]);
            }
            if(getenv("TSGPU_HOST_DEVICE_ART")) {      // 15 searches, one field: one launch up front carries (nearly) all walks
                const auto& after = tsgpu::Index::art_walk_stats();
                printf("multi_search: %llu launches for %llu walks\n", (unsigned long long) (after.launches - launches0),
                       (unsigned long long) (after.searches - searches0));
                CHECK(after.searches - searches0 >= 20 && after.launches - launches0 <= 2);
            }
            {   // the same list once more through the replay-batched multi_search (no thread per request; device walks fetched
                // lazily, one batch per pass): identical answers, and as few keyword device calls as the longest search has rounds
                std::vector<tsgpu::Index::batched_request> breqs;
                for(auto& r: reqs) { tsgpu::Index::batched_request b; b.r = r; breqs.push_back(b); }
                index.clear_walk_cache();
                tsgpu::Index::batched_stats bst;
                tsgpu::Index::batched_stats bst_any;
                for(size_t threads: {size_t(1), size_t(4)}) {
                    auto br = index.multi_search_batched(breqs, threads, threads == 1 ? &bst : &bst_any);
                    CHECK(br.size() == seq.size());
                    for(size_t i = 0; i < seq.size() && i < br.size(); i++) {
                        CHECK(br[i].status.ok() && br[i].found == seq[i].found && ids_of(br[i].raw_result_kvs) == ids_of(seq[i].raw_result_kvs));
                        for(size_t k = 0; k < seq[i].raw_result_kvs.size() && k < br[i].raw_result_kvs.size(); k++)
                            CHECK(seq[i].raw_result_kvs[k].scores[0] == br[i].raw_result_kvs[k].scores[0]);
                    }
                }
                printf("multi_search_batched: %zu passes, %zu keyword batches for %zu queries, %zu walk batches for %zu walks (%zu on the host)\n",
                       bst.passes, bst.kw_batches, bst.kw_queries, bst.walk_batches, bst.walks, bst.host_walk_fallbacks);
                CHECK(bst.kw_batches * 2 <= calls_seq);          // lazily fetched walks put the searches out of step by a pass or two
                auto bad = breqs;
                bad[1].r.the_fields = {"no_such_field"};
                auto r3 = index.multi_search_batched(bad, 2);
                CHECK(!r3[1].status.ok() && r3[0].status.ok() && ids_of(r3[0].raw_result_kvs) == ids_of(resps[0].raw_result_kvs));
            }
            for(size_t i = 0; i < cases.size(); i++) {
                CHECK(resps[i].status.ok());
                auto ids = ids_of(resps[i].raw_result_kvs);
                if(ids.size() > cases[i].per_page) ids.resize(cases[i].per_page);
                CHECK(ids == cases[i].expect);
                if(cases[i].found >= 0) CHECK((long) resps[i].found == cases[i].found);
            }
        }
        {   // concurrent searches on one Index, as the reference's request threads do (each holds only the shared lock)
            index.clear_walk_cache();
            std::vector<std::thread> threads;
            std::vector<int> ok(4, 0);
            for(int t = 0; t < 4; t++) threads.emplace_back([&, t] {
                int good = 0;
                for(int rep = 0; rep < 3; rep++) for(size_t i = t; i < cases.size(); i += 2) {
                    std::vector<tsgpu::KV> kv2; size_t f2 = 0;
                    if(!index.search(tsgpu::tokenize_ascii(cases[i].q), {"title"}, sort_fields, cases[i].drop, 250, kv2, f2, cases[i].o).ok()) continue;
                    auto ids = ids_of(kv2);
                    if(ids.size() > cases[i].per_page) ids.resize(cases[i].per_page);
                    good += ids == cases[i].expect;
                }
                ok[t] = good;
            });
            for(auto& th: threads) th.join();
            for(int t = 0; t < 4; t++) CHECK(ok[t] == 3 * (int) ((cases.size() - t + 1) / 2));
        }
        CHECK(index.search(tsgpu::tokenize_ascii("redundant"), {"title"}, sort_fields, 10, 250, kvs, found, opt(2, true, 0)).ok());
        CHECK(kvs.size() == 1 && found == 1);
        CHECK(index.search(tsgpu::tokenize_ascii("redundant"), {"title"}, sort_fields, 10, 250, kvs, found, opt(2, true, 10)).ok());
        CHECK(kvs.size() == 2 && found == 2);
    }
    // phrase: "rocket launch" as a phrase only in doc 8 ("... of a rocket launch these days")
    std::vector<uint32_t> both, phrase;
    CHECK(index.intersect("title", {"rocket", "launch"}, both).ok());
    CHECK(index.get_phrase_matches("title", {"rocket", "launch"}, both, phrase).ok());
    CHECK((phrase == std::vector<uint32_t>{8}));
}

// TEST_F(CollectionVectorTest, BasicVectorQuerying) test/collection_vector_search_test.cpp:75-135 (cosine, d = 4)
static void vector_scenario() {
    std::vector<std::vector<float>> values = {{0.851758f, 0.909671f, 0.823431f, 0.372063f}, {0.97826f, 0.933157f, 0.39557f, 0.306488f},
                                               {0.230606f, 0.634397f, 0.514009f, 0.399594f}};
    std::vector<float> vecs(12), q = {0.96826f, 0.94f, 0.39557f, 0.306488f}, nq(4);
    for(int i = 0; i < 3; i++) tso_normalize(values[i].data(), vecs.data() + 4 * i, 4);        // hnsw_index_t::normalize_vector
    tso_normalize(q.data(), nq.data(), 4);
    void* bld = tso_hnsw_build(vecs.data(), 3, 4, 16, 200, 100);
    uint32_t max_level = 0, entry = 0; uint64_t n_up = 0;
    tso_hnsw_build_info(bld, &max_level, &entry, &n_up);
    std::vector<uint8_t> levels(3); std::vector<uint32_t> links0(3 * 33), links_up((n_up + 1) * 17); std::vector<uint64_t> upper_off(4);
    tso_hnsw_build_fetch(bld, levels.data(), links0.data(), upper_off.data(), links_up.data());
    tso_hnsw_build_free(bld);
    tsgpu_hnsw g{3, 4, 16, max_level, entry, 1, vecs.data(), nullptr, levels.data(), links0.data(), upper_off.data(), links_up.data()};
    tsgpu::Index index(3);
    CHECK(index.add_vector_field(g).ok());
    auto pairs = index.searchKnnCloserFirst(nq.data(), 10, 10);
    CHECK(pairs.size() == 3);
    if(pairs.size() == 3) {
        CHECK(pairs[0].second == 1 && pairs[1].second == 0 && pairs[2].second == 2);
        CHECK(std::fabs(std::fabs(pairs[0].first) - 3.409385681152344e-05f) < 2e-7f);
        CHECK(std::fabs(pairs[1].first - 0.04329806566238403f) < 2e-7f);
        CHECK(std::fabs(pairs[2].first - 0.15141665935516357f) < 2e-7f);
    }
    std::vector<uint32_t> filter = {0, 1};          // filter_by points:[0,1]
    pairs = index.searchKnnCloserFirst(nq.data(), 10, 10, &filter);
    CHECK(pairs.size() == 2);
    if(pairs.size() == 2) CHECK(pairs[0].second == 1 && pairs[1].second == 0);
}

static std::vector<uint32_t> keys_of(const std::vector<tsgpu::KV>& kvs) { std::vector<uint32_t> r; for(auto& kv: kvs) r.push_back((uint32_t) kv.key); return r; }
// Rank fusion through Index::hybrid_search: HybridSearchRankFusionTest (test/collection_test.cpp:4779-4850) and
// TestRankFusionOrdering (test/collection_vector_search_test.cpp:5674-5753) with stand-in vectors of the same distance order
// (the reference embeds text with a model that is not available offline; only the order matters to its assertions);
// DistanceThresholdTest (:1548-1598) through Index::vector_search; process_results_bruteforce through Index::flat_distances.
static std::vector<float> unit_vec(std::vector<float> v) {
    float n = 0; for(float x: v) n += x * x;
    n = std::sqrt(n);
    for(float& x: v) x /= n;
    return v;
}
struct GraphHolder {
    std::vector<float> vecs; std::vector<uint8_t> levels; std::vector<uint32_t> links0, links_up; std::vector<uint64_t> upper_off;
    tsgpu_hnsw g{};
    GraphHolder(const std::vector<std::vector<float>>& rows, uint32_t metric) {
        const uint32_t n = (uint32_t) rows.size(), dim = (uint32_t) rows[0].size();
        for(auto& r: rows) vecs.insert(vecs.end(), r.begin(), r.end());
        void* bld = tso_hnsw_build(vecs.data(), n, dim, 16, 200, 100);
        uint32_t max_level = 0, entry = 0; uint64_t n_up = 0;
        tso_hnsw_build_info(bld, &max_level, &entry, &n_up);
        levels.resize(n); links0.resize((size_t) n * 33); links_up.resize((n_up + 1) * 17); upper_off.resize(n + 1);
        tso_hnsw_build_fetch(bld, levels.data(), links0.data(), upper_off.data(), links_up.data());
        tso_hnsw_build_free(bld);
        g = tsgpu_hnsw{n, dim, 16, max_level, entry, metric, vecs.data(), nullptr, levels.data(), links0.data(), upper_off.data(), links_up.data()};
    }
};
static float fused_score(const tsgpu::KV& kv) {
    const int32_t bits = (int32_t) kv.scores[kv.match_score_index];
    float f; memcpy(&f, &bits, 4);                  // int64_t_to_float for non-negative values (src/index.cpp:276-286)
    return f;
}
static void hybrid_scenarios() {
    const std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::seq_id, "", true}};
    const tsgpu_vec_params vp{0, 10, 0, FLT_MAX, 0.3f, 10};
    const auto q = unit_vec({1, 0.2f, 0, 0});
    {   // "butter" with prefix search: butter (cost 0), butterfly / butterball (prefix-found: cost 1); vector order butter < butterball < butterfly
        tsgpu::Index index(3);
        tsgpu::field_mirror_t title;
        const char* docs[] = {"butter", "butterball", "butterfly"};
        for(uint32_t i = 0; i < 3; i++) title.index_plain_string(i, tsgpu::tokenize_ascii(docs[i]));
        CHECK(index.add_field("title", title).ok());
        GraphHolder gh({q, unit_vec({1, 0.5f, 0, 0}), unit_vec({1, 1.5f, 0, 0})}, 0);
        CHECK(index.add_vector_field(gh.g).ok());
        std::vector<tsgpu::KV> kvs; size_t found = 0;
        CHECK(index.hybrid_search({{"butter"}, {"butterfly"}, {"butterball"}}, {0, 1, 1}, {"title"}, {15}, sort_fields, nullptr, {}, 250, q.data(), vp, kvs, found).ok());
        CHECK(found == 3 && (keys_of(kvs) == std::vector<uint32_t>{0, 1, 2}));
        const float expect[3] = {1.0f / 1 * 0.7f + 1.0f / 1 * 0.3f, 1.0f / 2 * 0.7f + 1.0f / 2 * 0.3f, 1.0f / 2 * 0.7f + 1.0f / 3 * 0.3f};
        for(size_t i = 0; i < kvs.size() && i < 3; i++) CHECK(std::fabs(fused_score(kvs[i]) - expect[i]) <= 3e-7f * expect[i]);
    }
    {   // "apple" matches all three with one shared text rank; vector order green apple < apple pie < red apple
        tsgpu::Index index(3);
        tsgpu::field_mirror_t title;
        const char* docs[] = {"red apple", "green apple", "apple pie"};
        for(uint32_t i = 0; i < 3; i++) title.index_plain_string(i, tsgpu::tokenize_ascii(docs[i]));
        CHECK(index.add_field("title", title).ok());
        GraphHolder gh({unit_vec({1, 2.0f, 0, 0}), q, unit_vec({1, 0.8f, 0, 0})}, 0);
        CHECK(index.add_vector_field(gh.g).ok());
        std::vector<tsgpu::KV> kvs; size_t found = 0;
        CHECK(index.hybrid_search({{"apple"}}, {0}, {"title"}, {15}, sort_fields, nullptr, {}, 250, q.data(), vp, kvs, found).ok());
        CHECK(found == 3 && (keys_of(kvs) == std::vector<uint32_t>{1, 2, 0}));
        const float expect[3] = {0.7f + 0.3f / 1, 0.7f + 0.3f / 2, 0.7f + 0.3f / 3};
        for(size_t i = 0; i < kvs.size() && i < 3; i++) CHECK(std::fabs(fused_score(kvs[i]) - expect[i]) <= 3e-7f * expect[i]);
        // process_results_bruteforce: distances of the query to given ids = 1 - dot
        std::vector<float> dist;
        CHECK(index.flat_distances(q.data(), {0, 1, 2}, dist).ok());
        CHECK(dist.size() == 3 && std::fabs(dist[1]) < 1e-6f && dist[2] < dist[0] && dist[2] > dist[1]);
    }
    {   // DistanceThresholdTest: cosine, wildcard + vector query
        tsgpu::Index index(2);
        GraphHolder gh({unit_vec({0.1f, 0.2f, 0.3f}), unit_vec({0.6f, 0.7f, 0.8f})}, 1);
        CHECK(index.add_vector_field(gh.g).ok());
        const auto vq = unit_vec({0.3f, 0.4f, 0.5f});
        const std::vector<tsgpu::sort_by> vsort = {{tsgpu::sort_by::vector_distance, "", false}, {tsgpu::sort_by::seq_id, "", true}};
        std::vector<tsgpu::KV> kvs; size_t found = 0;
        tsgpu_vec_params v2{0, 10, 0, FLT_MAX, 0.3f, 20};
        CHECK(index.vector_search(vsort, nullptr, {}, 250, vq.data(), v2, kvs, found).ok());
        CHECK(found == 2 && (keys_of(kvs) == std::vector<uint32_t>{1, 0}));
        v2.distance_threshold = 0.01f;
        CHECK(index.vector_search(vsort, nullptr, {}, 250, vq.data(), v2, kvs, found).ok());
        CHECK(found == 1 && (keys_of(kvs) == std::vector<uint32_t>{1}));
    }
}

// posting_t::get_exact_matches / get_prefix_matches (src/posting_list.cpp:1129-1452) on the ExactMatch documents of
// test/collection_test.cpp:3638, and ArrayUtils (test/array_utils_test.cpp:5-172)
static void exact_prefix_and_setops() {
    tsgpu::Index index(3);
    tsgpu::field_mirror_t title;
    const char* docs[] = {"Alpha", "Alpha Beta", "Alpha Beta Gamma"};
    for(uint32_t i = 0; i < 3; i++) title.index_plain_string(i, tsgpu::tokenize_ascii(docs[i]));
    CHECK(index.add_field("title", title).ok());
    std::vector<uint32_t> both, exact, prefix;
    CHECK(index.intersect("title", {"alpha", "beta"}, both).ok());
    CHECK((both == std::vector<uint32_t>{1, 2}));
    CHECK(index.get_exact_matches("title", {"alpha", "beta"}, both, exact).ok());
    CHECK((exact == std::vector<uint32_t>{1}));
    CHECK(index.get_exact_matches("title", {"alpha", "beta"}, both, prefix, true).ok());
    CHECK((prefix == std::vector<uint32_t>{1, 2}));
    std::vector<uint32_t> all;
    CHECK(index.intersect("title", {"alpha"}, all).ok());
    CHECK(index.get_exact_matches("title", {"alpha"}, all, exact).ok());
    CHECK((exact == std::vector<uint32_t>{0}));

    tsgpu::Index ids(400);
    std::vector<uint32_t> a = {0, 1, 2, 3, 4, 5, 6, 7, 8}, out;
    CHECK(ids.ids_setop(TSGPU_SET_AND, a, {3, 6, 9}, out).ok());
    CHECK((out == std::vector<uint32_t>{3, 6}));
    CHECK(ids.ids_setop(TSGPU_SET_OR, a, {3, 6, 9}, out).ok());
    CHECK((out == std::vector<uint32_t>{0, 1, 2, 3, 4, 5, 6, 7, 8, 9}));
    CHECK(ids.ids_setop(TSGPU_SET_EXCLUDE, a, {0, 1, 5, 7, 8}, out).ok());
    CHECK((out == std::vector<uint32_t>{2, 3, 4, 6}));
    CHECK(ids.ids_setop(TSGPU_SET_EXCLUDE, {58, 118, 185, 260, 322, 334, 353},
                        {58, 103, 116, 117, 137, 154, 191, 210, 211, 284, 299, 302, 306, 309, 332, 334, 360}, out).ok());
    CHECK((out == std::vector<uint32_t>{118, 185, 260, 322, 353}));
}

// ---- more of the reference's end-to-end expectations, through Index::search with Collection::search's defaults
struct Rec { std::vector<std::string> values; };     // one string per field


// builds an index over `fields` (plain strings) with points = row number
static void build_plain(tsgpu::Index& index, const std::vector<std::string>& fields, const std::vector<Rec>& recs) {
    for(size_t f = 0; f < fields.size(); f++) {
        tsgpu::field_mirror_t m;
        for(uint32_t i = 0; i < recs.size(); i++) m.index_plain_string(i, tsgpu::tokenize_ascii(recs[i].values[f]));
        CHECK(index.add_field(fields[f], m).ok());
    }
    std::unordered_map<uint32_t, int64_t> points;
    for(uint32_t i = 0; i < recs.size(); i++) points[i] = i;
    CHECK(index.add_sort_field("points", points).ok());
}

static void relevance_scenarios() {
    const std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "points", true}};
    std::vector<tsgpu::KV> kvs;
    size_t found = 0;
    {   // ExactMatch, test/collection_test.cpp:3638-3688
        tsgpu::Index index(3);
        build_plain(index, {"title"}, {{{"Alpha"}}, {{"Alpha Beta"}}, {{"Alpha Beta Gamma"}}});
        CHECK(index.search(tsgpu::tokenize_ascii("alpha beta"), {"title"}, sort_fields, 10, 250, kvs, found, opt(2, true)).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 2, 0}) && found == 3);
        CHECK(index.search(tsgpu::tokenize_ascii("alpha"), {"title"}, sort_fields, 10, 250, kvs, found, opt(2, true)).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 2, 1}) && found == 3);
    }
    {   // PrefixRankedAfterExactMatch :3922-3960
        tsgpu::Index index(4);
        build_plain(index, {"title"}, {{{"Rotini Puttanesca"}}, {{"Poulet Roti Tout Simple"}}, {{"Chapatis (Roti)"}}, {{"School Days Rotini Pasta Salad"}}});
        CHECK(index.search(tsgpu::tokenize_ascii("roti"), {"title"}, sort_fields, 5, 250, kvs, found, opt(0, true)).ok());
        CHECK(found == 4 && kvs.size() == 4);
        if(kvs.size() >= 3) CHECK(kvs[0].key == 2 && kvs[1].key == 1 && kvs[2].key == 3);
    }
    {   // MultiFieldMatchRanking :3788-3835 (query_by artist,title)
        const char* titles[] = {"Style", "Blank Space", "Balance Overkill", "Cardigan", "Invisible String", "The Last Great American Dynasty",
                                "Mirrorball", "Peace", "Betty", "Mad Woman"};
        std::vector<Rec> recs;
        for(auto t: titles) recs.push_back({{"Taylor Swift", t}});
        tsgpu::Index index(10);
        build_plain(index, {"artist", "title"}, recs);
        CHECK(index.search(tsgpu::tokenize_ascii("taylor swift style"), {"artist", "title"}, sort_fields, 5, 250, kvs, found, opt(0, true)).ok());
        CHECK(found == 10 && kvs.size() == 10);
        if(kvs.size() >= 3) CHECK(kvs[0].key == 0 && kvs[1].key == 9 && kvs[2].key == 8);
    }
    {   // MultiFieldMatchRankingOnArray :3837-3877 (two string[] fields)
        tsgpu::Index index(2);
        const std::vector<std::vector<std::vector<std::string>>> strong = {{{"golang"}, {"vue"}, {"react"}}, {{"golang"}, {"phoenix"}, {"react"}}};
        const std::vector<std::vector<std::vector<std::string>>> skills = {{{"docker"}, {"goa"}, {"elixir"}}, {{"docker"}, {"vue"}, {"kubernetes"}}};
        tsgpu::field_mirror_t a(true), b(true);
        for(uint32_t i = 0; i < 2; i++) { a.index_string_array(i, strong[i]); b.index_string_array(i, skills[i]); }
        CHECK(index.add_field("strong_skills", a).ok());
        CHECK(index.add_field("skills", b).ok());
        CHECK(index.add_sort_field("points", {{0, 0}, {1, 1}}).ok());
        CHECK(index.search(tsgpu::tokenize_ascii("golang vue"), {"strong_skills", "skills"}, sort_fields, 1, 250, kvs, found, opt(0, true)).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 1}) && found == 2);
    }
    {   // MultiFieldMatchRankingOnFieldOrder :3879-3920 (query_by_weights {1, 6})
        tsgpu::Index index(2);
        build_plain(index, {"title", "artist"}, {{{"Toxic", "Britney Spears"}}, {{"Bad", "Michael Jackson"}}});
        tsgpu::search_options o = opt(0, true);
        o.query_by_weights = {1, 6};
        CHECK(index.search(tsgpu::tokenize_ascii("michael jackson toxic"), {"title", "artist"}, sort_fields, 5, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 0}) && found == 2);
    }
    {   // MultiFieldRelevance2 :3276-3355
        tsgpu::Index index(2);
        build_plain(index, {"title", "artist"}, {{{"A Daikon Freestyle", "Ghosts on a Trampoline"}}, {{"Leaving on a Jetplane", "Coby Grant"}}});
        for(auto w: std::vector<std::vector<uint32_t>>{{}, {1, 4}, {1, 1}}) {
            tsgpu::search_options o = opt(0, true, 40);
            o.query_by_weights = w;
            CHECK(index.search(tsgpu::tokenize_ascii("on a jetplane"), {"title", "artist"}, sort_fields, 10, 250, kvs, found, o).ok());
            CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 0}) && found == 2);
        }
        tsgpu::search_options o = opt(0, true, 40);
        o.query_by_weights = {1, 4};
        CHECK(index.search(tsgpu::tokenize_ascii("on a helicopter"), {"title", "artist"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 1}) && found == 2);
    }
    {   // MultiFieldRelevance3 :3403-3460 and 6 :3581-3636
        tsgpu::search_options same = opt(0, true, 40);
        same.query_by_weights = {1, 1};
        tsgpu::Index i3(2);
        build_plain(i3, {"title", "artist"}, {{{"Taylor Swift Karaoke: reputation", "Taylor Swift"}}, {{"Style", "Taylor Swift"}}});
        CHECK(i3.search(tsgpu::tokenize_ascii("style taylor swift"), {"title", "artist"}, sort_fields, 10, 250, kvs, found, same).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 0}) && found == 2);
        CHECK(i3.search(tsgpu::tokenize_ascii("swift"), {"title", "artist"}, sort_fields, 10, 250, kvs, found, same).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 1}) && found == 2);
        tsgpu::Index i6(2);
        build_plain(i6, {"title", "artist"}, {{{"Taylor Swift", "Taylor Swift"}}, {{"Taylor Swift Song", "Taylor Swift"}}});
        for(bool exact: {true, false}) {
            tsgpu::search_options o = same;
            o.num_typos = 2;
            o.prioritize_exact_match = exact;
            CHECK(i6.search(tsgpu::tokenize_ascii("taylor swift"), {"title", "artist"}, sort_fields, 10, 250, kvs, found, o).ok());
            CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 0}) && found == 2);
        }
    }
    {   // RepeatingTokenRanking, test/collection_sorting_test.cpp:1800-1855: literal text_match values, weight {3}
        tsgpu::Index index(4);
        tsgpu::field_mirror_t m;
        const char* t[] = {"Mong Mong", "Mong Spencer", "Mong Mong Spencer", "Spencer Mong Mong"};
        for(uint32_t i = 0; i < 4; i++) m.index_plain_string(i, tsgpu::tokenize_ascii(t[i]));
        CHECK(index.add_field("title", m).ok());
        CHECK(index.add_sort_field("points", {{0, 100}, {1, 200}, {2, 300}, {3, 400}}).ok());
        tsgpu::search_options o = opt(2, true, 20);
        o.query_by_weights = {3};
        CHECK(index.search(tsgpu::tokenize_ascii("mong mong"), {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 3, 2, 1}));
        if(kvs.size() == 4) {
            CHECK(kvs[0].scores[0] == 1157451471583709209LL);
            CHECK(kvs[1].scores[0] == 1157451471575320601LL && kvs[2].scores[0] == 1157451471575320601LL && kvs[3].scores[0] == 1157451471575320601LL);
        }
    }
    {   // RelevanceConsiderAllFields, test/collection_specific_more_test.cpp:895-952: literal score, weights {3,2,1} used as given
        tsgpu::Index index(3);
        build_plain(index, {"f1", "f2", "f3"}, {{{"alpha", "alpha", "alpha"}}, {{"alpha", "alpha", "beta"}}, {{"alpha", "beta", "gamma"}}});
        tsgpu::search_options o = opt(2, true, 40);
        o.query_by_weights = {3, 2, 1};
        CHECK(index.search(tsgpu::tokenize_ascii("alpha"), {"f1", "f2", "f3"}, {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::seq_id, "", true}}, 0, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 1, 2}));
        if(kvs.size() == 3) {
            CHECK(kvs[0].scores[0] == 578730123373578267LL);
            for(int i = 0; i < 3; i++) CHECK(((kvs[i].scores[0] >> 11) & ((1LL << 48) - 1)) == 1108091342849LL && ((kvs[i].scores[0] >> 3) & 0xFF) == 3 && (kvs[i].scores[0] & 7) == 3 - i);
        }
    }
    {   // WeightTakingPrecendeceOverMatch :2196-2237: max_weight layout, literal best_field_score / weight / fields_matched
        tsgpu::Index index(2);
        build_plain(index, {"brand", "title"}, {{{"Light Plus", "Healthy Mayo"}}, {{"Vegabond", "Healthy Light Mayo"}}});
        tsgpu::search_options o = opt(2, true, 20);
        o.text_match_type = TSGPU_MATCH_MAX_WEIGHT;
        CHECK(index.search(tsgpu::tokenize_ascii("light mayo"), {"brand", "title"}, {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::seq_id, "", true}}, 5, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 1}));
        if(kvs.size() == 2) {
            CHECK(((kvs[0].scores[0] >> 3) & ((1LL << 48) - 1)) == 1108091338753LL && ((kvs[0].scores[0] >> 51) & 0xFF) == 15 && (kvs[0].scores[0] & 7) == 2);
            CHECK(((kvs[1].scores[0] >> 3) & ((1LL << 48) - 1)) == 2211897868289LL && ((kvs[1].scores[0] >> 51) & 0xFF) == 14 && (kvs[1].scores[0] & 7) == 1);
        }
    }
    {   // text_match literals of test/collection_vector_search_test.cpp:5462-5496 (and test/union_test.cpp:810)
        tsgpu::Index index(4);
        build_plain(index, {"name"}, {{{"Nike running shoes for men"}}, {{"Nike running sneakers"}}, {{"adidas shoes"}}, {{"puma"}}});
        CHECK(index.search(tsgpu::tokenize_ascii("nike running shoes"), {"name"}, {{tsgpu::sort_by::text_match, "", true}}, 10, 250, kvs, found, opt(0, false)).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0, 1, 2}));
        if(kvs.size() == 3) CHECK(kvs[0].scores[0] == 1736172819517016185LL && kvs[1].scores[0] == 1157451471441102969LL && kvs[2].scores[0] == 578730123365189753LL);
    }
}

// PhraseSearch, test/collection_specific_test.cpp:2504-2622 — the cases that combine phrases with tokens or exclusions
// and the phrase-only queries with do_phrase_search's own score
static void phrase_scenarios() {
    tsgpu::Index index(3);
    tsgpu::field_mirror_t m;
    const char* t[] = {"Then and there by the down", "Down There by the Train", "The State Trooper"};
    for(uint32_t i = 0; i < 3; i++) m.index_plain_string(i, tsgpu::tokenize_ascii(t[i]));
    CHECK(index.add_field("title", m).ok());
    const std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::seq_id, "", true}};   // no default_sorting_field
    std::vector<tsgpu::KV> kvs;
    size_t found = 0;
    auto P = [](std::initializer_list<const char*> w) { std::vector<std::string> v; for(auto x: w) v.push_back(x); return v; };
    // without phrase search: "down there by", drop_tokens_threshold 0
    CHECK(index.search(tsgpu::tokenize_ascii("down there by"), {"title"}, sort_fields, 0, 250, kvs, found, opt(0, false)).ok());
    CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 0}));
    {   // "by the" and
        tsgpu::search_options o = opt(0, false);
        o.phrases = {P({"by", "the"})};
        CHECK(index.search({"and"}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0}));
        // "by the" state
        CHECK(index.search({"state"}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK(kvs.empty());
    }
    {   // -"by the down"  and  -"by the"  and  -"by the dinosaur"
        tsgpu::search_options o = opt(0, false);
        o.exclude_phrases = {P({"by", "the", "down"})};
        CHECK(index.search({}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{2, 1}));
        o.exclude_phrases = {P({"by", "the"})};
        CHECK(index.search({}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{2}));
        o.exclude_phrases = {P({"by", "the", "dinosaur"})};
        CHECK(index.search({}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK(kvs.size() == 3);
    }
    {   // phrase-only queries (test/collection_specific_test.cpp:2535-2545): do_phrase_search's own scoring, 100000 + field weight
        tsgpu::search_options o = opt(0, false);
        o.phrases = {P({"down", "there", "by"})};                      // " down there by "
        CHECK(index.search({}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{1}) && found == 1);
        CHECK(kvs.size() == 1 && kvs[0].scores[0] == 100000 + 15 && kvs[0].match_score_index == 0);
        o.phrases = {P({"by", "the"})};                                // "by the" -train
        o.exclude_tokens = {"train"};
        CHECK(index.search({}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{0}) && found == 1);
        o.exclude_tokens.clear();                                      // "by the": both documents, same score, the later id first
        CHECK(index.search({}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 0}) && found == 2);
        o.phrases = {P({"by", "the", "dinosaur"})};                    // a phrase with an unknown token: nothing
        CHECK(index.search({}, {"title"}, sort_fields, 10, 250, kvs, found, o).ok());
        CHECK(kvs.empty() && found == 0);
    }
}

// test/collection_specific_test.cpp scenarios (typos, prefixes, several fields, weights, string[]): the table is
// generated from the Python harness' CASES so both run the same reference expectations
struct SpecificCase {
    const char* name;
    std::vector<std::string> fields;
    std::vector<bool> is_array;
    std::vector<std::vector<std::vector<std::string>>> docs;      // [doc][field][element]
    std::vector<long> points;
    const char* query;
    uint32_t num_typos; bool prefix; size_t drop, typo_thr;
    std::vector<uint32_t> weights;
    std::vector<uint32_t> expect;
    int token_order; size_t max_candidates; long found; bool head;
    int flags, match_type, drop_mode;
    size_t both_sides_limit;
};
static void specific_scenarios() {
    const std::vector<SpecificCase> cases = {
#include "specific_cases.inc"
    };
    const std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "points", true}};
    for(auto& c: cases) {
        tsgpu::Index index((uint32_t) c.docs.size());
        for(size_t f = 0; f < c.fields.size(); f++) {
            tsgpu::field_mirror_t m(c.is_array[f]);
            for(uint32_t d = 0; d < c.docs.size(); d++) {
                if(c.is_array[f]) {
                    std::vector<std::vector<std::string>> elems;
                    for(auto& e: c.docs[d][f]) elems.push_back(tsgpu::tokenize_ascii(e));
                    m.index_string_array(d, elems);
                } else m.index_plain_string(d, tsgpu::tokenize_ascii(c.docs[d][f].empty() ? std::string() : c.docs[d][f][0]));
            }
            CHECK(index.add_field(c.fields[f], m).ok());
        }
        std::unordered_map<uint32_t, int64_t> points;
        for(uint32_t d = 0; d < c.points.size(); d++) points[d] = c.points[d];
        CHECK(index.add_sort_field("points", points).ok());
        tsgpu::search_options o = opt(c.num_typos, c.prefix, c.typo_thr, c.token_order ? tsgpu::search_options::MAX_SCORE : tsgpu::search_options::FREQUENCY);
        o.query_by_weights = c.weights;
        o.max_candidates = c.max_candidates;
        o.prioritize_exact_match = (c.flags & TSGPU_FLAG_PRIORITIZE_EXACT_MATCH) != 0;
        o.prioritize_token_position = (c.flags & TSGPU_FLAG_PRIORITIZE_TOKEN_POSITION) != 0;
        o.prioritize_num_matching_fields = (c.flags & TSGPU_FLAG_PRIORITIZE_NUM_MATCHING_FIELDS) != 0;
        o.text_match_type = c.match_type;
        o.drop_tokens_mode = c.drop_mode == 0 ? tsgpu::search_options::right_to_left : c.drop_mode == 1 ? tsgpu::search_options::left_to_right : tsgpu::search_options::both_sides;
        o.drop_both_sides_token_limit = c.both_sides_limit;
        std::vector<tsgpu::KV> kvs;
        size_t found = 0;
        CHECK(index.search(tsgpu::tokenize_ascii(c.query), c.fields, sort_fields, c.drop, 250, kvs, found, o).ok());
        auto got = keys_of(kvs);
        if(c.found >= 0) CHECK((long) found == c.found && (long) got.size() == c.found);
        if(c.head && got.size() > c.expect.size()) got.resize(c.expect.size());
        if(got != c.expect) { printf("specific case %s: got", c.name); for(auto k: got) printf(" %u", k); printf("\n"); }
        CHECK(got == c.expect);
    }
}

// test/collection_synonyms_test.cpp scenarios (table generated from tests/test_synonym_scenarios.py)
struct SynonymCase {
    const char* name;
    std::vector<std::string> fields;
    std::vector<std::vector<std::string>> docs;       // [doc][field]
    std::vector<long> points;
    const char* query;
    std::vector<std::vector<std::string>> synonyms;
    uint32_t num_typos; bool prefix; size_t drop, typo_thr; bool demote;
    std::vector<uint32_t> expect;
    const char* relation;      // "", "eq", "ne", "gt": first two text_match values
};
static void synonym_scenarios() {
    const std::vector<SynonymCase> cases = {
#include "synonym_cases.inc"
    };
    const std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "points", true}};
    for(auto& c: cases) {
        tsgpu::Index index((uint32_t) c.docs.size());
        for(size_t f = 0; f < c.fields.size(); f++) {
            tsgpu::field_mirror_t m;
            for(uint32_t d = 0; d < c.docs.size(); d++) m.index_plain_string(d, tsgpu::tokenize_ascii(c.docs[d][f]));
            CHECK(index.add_field(c.fields[f], m).ok());
        }
        std::unordered_map<uint32_t, int64_t> points;
        for(uint32_t d = 0; d < c.points.size(); d++) points[d] = c.points[d];
        CHECK(index.add_sort_field("points", points).ok());
        tsgpu::search_options o = opt(c.num_typos, c.prefix, c.typo_thr);
        o.synonyms = c.synonyms;
        o.demote_synonym_match = c.demote;
        std::vector<tsgpu::KV> kvs;
        size_t found = 0;
        CHECK(index.search(tsgpu::tokenize_ascii(c.query), c.fields, sort_fields, c.drop, 250, kvs, found, o).ok());
        const auto got = keys_of(kvs);
        if(got != c.expect) { printf("synonym case %s: got", c.name); for(auto k: got) printf(" %u", k); printf("\n"); }
        CHECK(got == c.expect && found == c.expect.size());
        const std::string rel = c.relation;
        if(!rel.empty() && kvs.size() >= 2) {
            const int64_t a = kvs[0].scores[0], b = kvs[1].scores[0];
            CHECK(rel == "eq" ? a == b : rel == "ne" ? a != b : a > b);
        }
    }
}

// String filter_by scenarios of test/collection_filtering_test.cpp (table generated from tests/test_filter_scenarios.py)
struct FilterCheck { const char* field; const char* raw; long count; std::vector<uint32_t> ids; };
struct FilterCase {
    const char* source;
    std::vector<std::string> fields;
    std::vector<bool> is_array;
    std::vector<std::vector<std::vector<std::string>>> docs;      // [doc][field][element]
    std::vector<long> points;
    std::vector<FilterCheck> checks;
};
static void filter_scenarios() {
    const std::vector<FilterCase> cases = {
#include "filter_cases.inc"
    };
    for(auto& c: cases) {
        tsgpu::Index index((uint32_t) c.docs.size());
        for(size_t f = 0; f < c.fields.size(); f++) {
            tsgpu::field_mirror_t m(c.is_array[f]);
            for(uint32_t d = 0; d < c.docs.size(); d++) {
                if(c.is_array[f]) {
                    std::vector<std::vector<std::string>> elems;
                    for(auto& e: c.docs[d][f]) elems.push_back(tsgpu::tokenize_ascii(e));
                    m.index_string_array(d, elems);
                } else m.index_plain_string(d, tsgpu::tokenize_ascii(c.docs[d][f].empty() ? std::string() : c.docs[d][f][0]));
            }
            CHECK(index.add_field(c.fields[f], m).ok());
        }
        std::unordered_map<uint32_t, int64_t> points;
        for(uint32_t d = 0; d < c.points.size(); d++) points[d] = c.points[d];
        CHECK(index.add_sort_field("points", points).ok());
        for(auto& k: c.checks) {
            std::vector<uint32_t> ids;
            CHECK(index.string_filter_ids(k.field, k.raw, ids).ok());
            const bool ok = k.count >= 0 ? (long) ids.size() == k.count : ids == k.ids;
            if(!ok) { printf("filter case %s `%s:%s`: got", c.source, k.field, k.raw); for(auto i: ids) printf(" %u", i); printf("\n"); }
            CHECK(ok);
        }
    }
    std::vector<uint32_t> ids;
    tsgpu::string_filter_exp exp;
    CHECK(!tsgpu::parse_string_filter("tags", "=", exp).ok());        // "Filter value cannot be empty." (FilterOnTextFields :139)
}

// documents added, updated and removed after the mirror was loaded (Index::update_field -> tsgpu_index_append_lists):
// the patched index answers like one built from the final documents — exact tokens, prefixes, typo candidates (the ART mirror is
// rebuilt from the patched vocabulary) and the default sorting field of the new documents.
static void incremental_scenarios() {
    const std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "points", true}};
    std::vector<std::string> titles = {"The quick brown fox", "Rocket launch delayed by weather", "Brown bears of the north", "A rocket science primer",
                                       "Launch day for the new rover", "Quick guide to foxes"};
    const uint32_t capacity = 16;
    tsgpu::Index index(capacity);
    tsgpu::field_mirror_t m;
    for(uint32_t i = 0; i < titles.size(); i++) m.index_plain_string(i, tsgpu::tokenize_ascii(titles[i]));
    CHECK(index.add_field("title", m).ok());
    m.take_delta();
    std::unordered_map<uint32_t, int64_t> points;
    for(uint32_t i = 0; i < titles.size(); i++) points[i] = 100 - i;
    CHECK(index.add_sort_field("points", points).ok());
    std::vector<tsgpu::KV> kvs;
    size_t found = 0;
    CHECK(index.search(tsgpu::tokenize_ascii("rocket"), {"title"}, sort_fields, 10, 250, kvs, found, opt(2, true)).ok());
    CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 3}) && found == 2);
    // one batch of writes: four new documents, document 2 rewritten, document 4 removed
    titles.push_back("Rocket fuel and brown sugar");          // 6
    titles.push_back("Foxes launch a quick rocket");          // 7
    titles.push_back("Northern lights guide");                // 8
    titles.push_back("Rover lands after long launch");        // 9
    for(uint32_t i = 6; i < 10; i++) m.index_plain_string(i, tsgpu::tokenize_ascii(titles[i]));
    titles[2] = "Rocket bears of the north";
    m.remove(2); m.index_plain_string(2, tsgpu::tokenize_ascii(titles[2]));
    m.remove(4);
    CHECK(index.update_field("title", m.take_delta()).ok());
    CHECK(index.set_sort_values("points", {6, 7, 8, 9, 4}, {94, 93, 92, 91, INT64_MIN}).ok());
    // the same final state, loaded at once
    tsgpu::Index fresh(capacity);
    tsgpu::field_mirror_t mf;
    std::unordered_map<uint32_t, int64_t> pf;
    for(uint32_t i = 0; i < titles.size(); i++) if(i != 4) { mf.index_plain_string(i, tsgpu::tokenize_ascii(titles[i])); pf[i] = 100 - i; }
    CHECK(fresh.add_field("title", mf).ok());
    CHECK(fresh.add_sort_field("points", pf).ok());
    CHECK(index.search(tsgpu::tokenize_ascii("rocket"), {"title"}, sort_fields, 10, 250, kvs, found, opt(2, true)).ok());
    CHECK((keys_of(kvs) == std::vector<uint32_t>{1, 2, 3, 6, 7}) && found == 5);          // equal text scores: points descending
    for(const char* q: {"rocket", "launch", "brown", "quick rocket", "rocket launch", "bears", "rover", "north", "lau", "rokcet", "foxs", "luanch day", "guide", "the"}) {
        std::vector<tsgpu::KV> a, b;
        size_t fa = 0, fb = 0;
        const bool prefix = std::strlen(q) <= 3;
        CHECK(index.search(tsgpu::tokenize_ascii(q), {"title"}, sort_fields, 1, 250, a, fa, opt(2, prefix)).ok());
        CHECK(fresh.search(tsgpu::tokenize_ascii(q), {"title"}, sort_fields, 1, 250, b, fb, opt(2, prefix)).ok());
        CHECK(fa == fb && keys_of(a) == keys_of(b));
        if(!(fa == fb && keys_of(a) == keys_of(b))) printf("  incremental scenario: query '%s' differs (%zu vs %zu hits)\n", q, fa, fb);
        for(size_t i = 0; i < a.size() && i < b.size(); i++) CHECK(a[i].scores[0] == b[i].scores[0] && a[i].scores[1] == b[i].scores[1]);
        for(auto& kv: a) CHECK(kv.key != 4);                                       // the removed document is gone
    }
    // a removed token leaves the vocabulary: "day" only lived in document 4
    CHECK(index.search(tsgpu::tokenize_ascii("day"), {"title"}, sort_fields, 0, 250, kvs, found, opt(0, false)).ok());
    CHECK(found == 0 && kvs.empty());
}

// group_by: Topster<KV> with distinct > 0 (include/topster.h:357-376) + Index::populate_result_kvs (src/index.cpp:8961-9014).
// (1) the table of TEST(TopsterTest, DistinctIntValues), test/topster_test.cpp:181-262, through host_group_topster_t: groups ranked by
// their best KV, each group's KVs best first, the greatest KV per key. (2) Index::search_grouped against the definition: EVERY hit of
// the search (a Topster larger than the result set) fed to the group Topsters — with a first Topster of 4 so that the truncated-list
// paths run (larger Topster, rounds without the placed groups, rounds restricted to one group).
static void group_by_scenarios() {
    {
        struct { uint64_t distinct_key; int64_t match_score, primary_attr, secondary_attr; } data[14] = {
            {1, 11, 20, 30}, {1, 12, 20, 32}, {2, 4, 20, 30}, {3, 7, 20, 30}, {4, 14, 20, 30}, {5, 9, 20, 30}, {5, 10, 20, 32},
            {5, 9, 20, 30}, {6, 6, 20, 30}, {7, 6, 22, 30}, {7, 6, 22, 30}, {8, 9, 20, 30}, {9, 8, 20, 30}, {10, 5, 20, 30}};
        tsgpu::host_group_topster_t gt(5, 2);
        for(int i = 0; i < 14; i++) {
            tsgpu::KV kv{};
            kv.key = (uint64_t) i + 100; kv.distinct_key = data[i].distinct_key;
            kv.scores[0] = data[i].match_score; kv.scores[1] = data[i].primary_attr; kv.scores[2] = data[i].secondary_attr;
            gt.add(kv);
        }
        const auto groups = gt.result();
        std::vector<uint64_t> order;
        for(auto& g: groups) order.push_back(g[0].distinct_key);
        CHECK((order == std::vector<uint64_t>{4, 1, 5, 8, 9}));
        for(auto& g: groups) {
            if(g[0].distinct_key == 1) CHECK(g.size() == 2 && g[0].scores[0] == 12 && g[1].scores[0] == 11);
            if(g[0].distinct_key == 5) CHECK(g.size() == 2 && g[0].scores[0] == 10 && g[0].key == 106 && g[1].scores[0] == 9 && g[1].key == 107);
            if(g[0].distinct_key == 4) CHECK(g.size() == 1 && g[0].scores[0] == 14);
        }
    }
    const std::vector<tsgpu::sort_by> sort_fields = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "points", true}};
    const char* words[] = {"running", "shoe", "trail", "road", "light", "boot", "winter", "sandal", "leather", "kids"};
    const uint32_t n = 120;
    tsgpu::Index index(n);
    tsgpu::field_mirror_t m;
    std::unordered_map<uint32_t, int64_t> points, brand;
    uint32_t rng = 12345;
    auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for(uint32_t i = 0; i < n; i++) {
        std::string title = "shoe";
        const uint32_t extra = 1 + next() % 3;
        for(uint32_t k = 0; k < extra; k++) { title += " "; title += words[next() % 10]; }
        m.index_plain_string(i, tsgpu::tokenize_ascii(title));
        points[i] = (int64_t) (next() % 50);
        if(i % 11 != 0) brand[i] = 1000 + (int64_t) (next() % 9) * (int64_t) (1 + next() % 2);      // ~14 brands of uneven size; every 11th document has none
    }
    CHECK(index.add_field("title", m).ok());
    CHECK(index.add_sort_field("points", points).ok());
    CHECK(index.add_sort_field("brand", brand).ok());
    for(const char* q: {"shoe", "trail shoe", "shoe runing", "boot", "light road shoe", "sandl"}) {
        for(int variant = 0; variant < 3; variant++) {
            const size_t capacity = variant == 0 ? 3 : (variant == 1 ? 6 : 40), L = variant == 1 ? 1 : 2;
            const bool group_missing = variant == 2;
            // the definition: every hit -> group Topsters
            std::vector<tsgpu::KV> all;
            size_t found = 0;
            CHECK(index.search(tsgpu::tokenize_ascii(q), {"title"}, sort_fields, 1, 1024, all, found, opt(2, true)).ok());
            CHECK(all.size() < 1024);
            tsgpu::host_group_topster_t want(capacity, L);
            std::set<uint64_t> want_groups;
            for(auto kv: all) {
                auto it = brand.find((uint32_t) kv.key);
                kv.distinct_key = it == brand.end() ? (group_missing ? 1ull : kv.key) : (uint64_t) it->second;
                want_groups.insert(kv.distinct_key);
                want.add(kv);
            }
            const auto expect = want.result();
            std::vector<std::vector<tsgpu::KV>> got;
            size_t found_groups = 0;
            // first Topster 4, never more than 8 hits per list: every follow-up path runs
            CHECK(index.search_grouped(tsgpu::tokenize_ascii(q), {"title"}, sort_fields, 1, capacity, "brand", L, group_missing, got, found_groups, opt(2, true), 4, 8).ok());
            bool same = got.size() == expect.size();
            for(size_t g = 0; same && g < got.size(); g++) {
                same = got[g].size() == expect[g].size();
                for(size_t i = 0; same && i < got[g].size(); i++)
                    same = got[g][i].key == expect[g][i].key && got[g][i].distinct_key == expect[g][i].distinct_key && got[g][i].scores[0] == expect[g][i].scores[0] &&
                           got[g][i].scores[1] == expect[g][i].scores[1];
            }
            CHECK(same);
            if(!same) {
                printf("  group_by scenario: query '%s' variant %d: %zu groups vs %zu expected\n", q, variant, got.size(), expect.size());
                for(size_t g = 0; g < std::max(got.size(), expect.size()); g++) {
                    printf("    group %zu: got", g);
                    if(g < got.size()) for(auto& kv: got[g]) printf(" (%llu|%llu|%lld,%lld)", (unsigned long long) kv.distinct_key, (unsigned long long) kv.key, (long long) kv.scores[0], (long long) kv.scores[1]);
                    printf("  want");
                    if(g < expect.size()) for(auto& kv: expect[g]) printf(" (%llu|%llu|%lld,%lld)", (unsigned long long) kv.distinct_key, (unsigned long long) kv.key, (long long) kv.scores[0], (long long) kv.scores[1]);
                    printf("\n");
                }
            }
            // and with the default first Topster (nothing truncated): the same groups
            std::vector<std::vector<tsgpu::KV>> got2;
            CHECK(index.search_grouped(tsgpu::tokenize_ascii(q), {"title"}, sort_fields, 1, capacity, "brand", L, group_missing, got2, found_groups, opt(2, true)).ok());
            CHECK(got2.size() == expect.size() && found_groups == want_groups.size());
            for(size_t g = 0; g < got2.size() && g < expect.size(); g++) CHECK(got2[g].size() == expect[g].size() && got2[g][0].key == expect[g][0].key);
        }
    }
}

// TEST_F(CollectionGroupingTest, GroupingBasics), test/collection_grouping_test.cpp:64-198, over test/group_documents.jsonl
// (tests/golden/group_documents.jsonl): group_by an int field and a float field under different sort clauses (wildcard query), the
// per-group `found`, and "typo_tokens_threshold should respect num_groups" (the typo loop counts groups).
static double json_float(const std::string& line, const std::string& key) {
    size_t p = line.find("\"" + key + "\"");
    p = line.find(':', p);
    return std::strtod(line.c_str() + p + 1, nullptr);
}
static int64_t float_to_int64(float f) {                    // Index::float_to_int64_t, src/index.cpp:266-274
    int32_t i;
    std::memcpy(&i, &f, sizeof i);
    if(i < 0) i ^= INT32_MAX;
    return i;
}
static void grouping_basics(const std::string& jsonl) {
    std::ifstream in(jsonl);
    if(!in.good()) { printf("grouping fixture %s not found\n", jsonl.c_str()); CHECK(false); return; }
    std::vector<std::string> lines;
    for(std::string l; std::getline(in, l);) if(!l.empty()) lines.push_back(l);
    const uint32_t n = (uint32_t) lines.size();
    CHECK(n == 12);
    tsgpu::Index index(n);
    tsgpu::field_mirror_t title, brand;
    std::unordered_map<uint32_t, int64_t> rating, size_col, size_key, rating_key, brand_key, size_brand_key;
    std::map<std::string, int64_t> brand_ids;
    for(uint32_t i = 0; i < n; i++) {
        title.index_plain_string(i, tsgpu::tokenize_ascii(json_str(lines[i], "title")));
        const float r = (float) json_float(lines[i], "rating");
        rating[i] = float_to_int64(r);
        size_col[i] = json_int(lines[i], "size");
        size_key[i] = 100 + json_int(lines[i], "size");                 // any injective stand-in for the facet hash get_distinct_id combines
        rating_key[i] = 1000000 + float_to_int64(r);
        size_brand_key[i] = 100000 * json_int(lines[i], "size");          // compound key: the optional brand adds nothing when absent
        if(lines[i].find("\"brand\"") != std::string::npos) {
            const std::string b = json_str(lines[i], "brand");
            brand.index_plain_string(i, tsgpu::tokenize_ascii(b));
            if(!brand_ids.count(b)) brand_ids[b] = 7000 + (int64_t) brand_ids.size();
            brand_key[i] = brand_ids[b];
            size_brand_key[i] += brand_ids[b];
        }
    }
    CHECK(index.add_field("title", title).ok());
    CHECK(index.add_field("brand", brand).ok());
    CHECK(index.add_sort_field("rating", rating).ok());                  // the collection's default sorting field
    CHECK(index.add_sort_field("size", size_col).ok());
    CHECK(index.add_sort_field("size_key", size_key).ok());
    CHECK(index.add_sort_field("rating_key", rating_key).ok());
    CHECK(index.add_sort_field("brand_key", brand_key).ok());
    CHECK(index.add_sort_field("size_brand_key", size_brand_key).ok());
    std::vector<std::vector<tsgpu::KV>> groups;
    std::vector<size_t> gfound;
    size_t n_groups = 0;
    auto ids_of = [](const std::vector<tsgpu::KV>& g) { std::vector<uint32_t> r; for(auto& kv: g) r.push_back((uint32_t) kv.key); return r; };
    {   // `*` grouped by size, group_limit 2, default sort (rating desc)
        const std::vector<tsgpu::sort_by> by_rating = {{tsgpu::sort_by::numeric, "rating", true}};
        CHECK(index.search_grouped({}, {"title"}, by_rating, 1, 250, "size_key", 2, false, groups, n_groups, opt(0, false), 0, 1024, &gfound).ok());
        CHECK(n_groups == 3 && groups.size() == 3);
        if(groups.size() == 3) {
            CHECK((ids_of(groups[0]) == std::vector<uint32_t>{5, 1}) && groups[0][0].distinct_key == 111 && gfound[0] == 2);
            CHECK((ids_of(groups[1]) == std::vector<uint32_t>{4, 3}) && gfound[1] == 7);
            CHECK((ids_of(groups[2]) == std::vector<uint32_t>{2, 8}) && gfound[2] == 3);
        }
    }
    {   // `*` grouped by rating, sort by size desc: 7 unique ratings
        const std::vector<tsgpu::sort_by> by_size = {{tsgpu::sort_by::numeric, "size", true}};
        CHECK(index.search_grouped({}, {"title"}, by_size, 1, 250, "rating_key", 2, false, groups, n_groups, opt(0, false), 0, 1024, &gfound).ok());
        CHECK(n_groups == 7 && groups.size() == 7);
        if(groups.size() == 7) {
            CHECK((ids_of(groups[0]) == std::vector<uint32_t>{8}) && gfound[0] == 1);
            CHECK((ids_of(groups[1]) == std::vector<uint32_t>{6, 1}) && gfound[1] == 4);
            CHECK((ids_of(groups[5]) == std::vector<uint32_t>{9}) && gfound[5] == 1);
            CHECK((ids_of(groups[6]) == std::vector<uint32_t>{0}) && gfound[6] == 1);
        }
        // the same through the truncated-list paths
        std::vector<std::vector<tsgpu::KV>> g2;
        CHECK(index.search_grouped({}, {"title"}, by_size, 1, 250, "rating_key", 2, false, g2, n_groups, opt(0, false), 2, 1024).ok());
        CHECK(g2.size() == groups.size());
        for(size_t g = 0; g < g2.size() && g < groups.size(); g++) CHECK(ids_of(g2[g]) == ids_of(groups[g]));
        // ... and with lists of at most 4 hits: placed groups excluded, short groups searched alone
        CHECK(index.search_grouped({}, {"title"}, by_size, 1, 250, "rating_key", 2, false, g2, n_groups, opt(0, false), 2, 4).ok());
        CHECK(g2.size() == groups.size());
        for(size_t g = 0; g < g2.size() && g < groups.size(); g++) CHECK(ids_of(g2[g]) == ids_of(groups[g]));
    }
    {   // TEST_F(CollectionGroupingTest, GroupingCompoundKey), :200-232: `*` grouped by (size, brand), group_limit 2 — 10 groups; the
        // documents without a brand group by their size alone
        const std::vector<tsgpu::sort_by> by_rating = {{tsgpu::sort_by::numeric, "rating", true}};
        CHECK(index.search_grouped({}, {"title"}, by_rating, 1, 250, "size_brand_key", 2, false, groups, n_groups, opt(0, false), 0, 1024, &gfound).ok());
        CHECK(n_groups == 10 && groups.size() == 10);
        if(groups.size() == 10) {
            CHECK((ids_of(groups[0]) == std::vector<uint32_t>{5}) && gfound[0] == 1);
            CHECK((ids_of(groups[1]) == std::vector<uint32_t>{4}) && gfound[1] == 1);
            CHECK((ids_of(groups[2]) == std::vector<uint32_t>{3, 0}) && gfound[2] == 2);
            CHECK((ids_of(groups[5]) == std::vector<uint32_t>{10, 11}));
        }
        // "pagination with page=2, per_page=2": the third and fourth groups
        if(groups.size() >= 4) CHECK(groups[2][0].key == 3 && groups[2][1].key == 0 && groups[2][0].distinct_key == (uint64_t) (100000 * 10 + brand_ids["Omega"]));
    }
    {   // typo_tokens_threshold counts groups: "beta" in brand, threshold 2 -> Beta and (one typo away) Zeta; threshold 1 -> Beta only
        const std::vector<tsgpu::sort_by> text_then_rating = {{tsgpu::sort_by::text_match, "", true}, {tsgpu::sort_by::numeric, "rating", true}};
        CHECK(index.search_grouped({"beta"}, {"brand"}, text_then_rating, 1, 250, "brand_key", 1, false, groups, n_groups, opt(2, false, 2), 0, 1024, &gfound).ok());
        CHECK(n_groups == 2 && groups.size() == 2);
        if(groups.size() == 2) {
            CHECK(groups[0][0].distinct_key == (uint64_t) brand_ids["Beta"] && groups[1][0].distinct_key == (uint64_t) brand_ids["Zeta"]);
            CHECK(gfound[0] == 3 && gfound[1] == 1);
        }
        CHECK(index.search_grouped({"beta"}, {"brand"}, text_then_rating, 1, 250, "brand_key", 1, false, groups, n_groups, opt(2, false, 1), 0, 1024, &gfound).ok());
        CHECK(n_groups == 1 && groups.size() == 1 && gfound.size() == 1 && gfound[0] == 3);
    }
}

int main(int argc, char** argv) {
    if(tsgpu_device_count() == 0) { printf("no CUDA device: nothing to run (the library has no CPU path)\n"); return 99; }
    posting_list_intersection_basics();
    or_iterator_intersect_and_filter();
    collection_scenarios(argc > 1 ? argv[1] : "tests/golden/documents.jsonl");
    vector_scenario();
    if(getenv("TSGPU_HOST_HYBRID_KAT")) hybrid_scenarios();      // tiny-graph hybrid calls: run on the double and, non-gating, on the GPU
    exact_prefix_and_setops();
    relevance_scenarios();
    specific_scenarios();
    phrase_scenarios();
    synonym_scenarios();
    filter_scenarios();
    incremental_scenarios();
    group_by_scenarios();
    if(getenv("TSGPU_HOST_GROUPING_KAT")) {      // the reference's GroupingBasics: gating on the double; on the GPU reported but not gating until its first GPU run (added after the round's GPU budget was spent)
        std::string g = argc > 1 ? argv[1] : "tests/golden/documents.jsonl";
        const size_t p = g.rfind("documents.jsonl");
        grouping_basics(p == std::string::npos ? "tests/golden/group_documents.jsonl" : g.substr(0, p) + "group_documents.jsonl");
    }
    {
        const auto& ws = tsgpu::Index::art_walk_stats();
        if(getenv("TSGPU_HOST_DEVICE_ART")) {
            printf("device ART walks: %llu launches, %llu searches, %llu served from them, %llu host fallbacks\n", (unsigned long long) ws.launches,
                   (unsigned long long) ws.searches, (unsigned long long) ws.served, (unsigned long long) ws.host_fallbacks);
            CHECK(ws.launches > 50 && ws.served > 200 && ws.searches > ws.launches);     // batched, and really used
        } else CHECK(ws.launches == 0);
    }
    printf("%s (%d failed checks)\n", failures ? "FAILED" : "PASSED", failures);
    return failures;
}
