// Stand-alone check of tsgpu_flat_distances_batch (the tensor-core flat scan, csrc/flat_tc.cu) through the C-ABI only: random unit
// vectors, a shared candidate set, several group sizes; compares with a double-precision dot on the host and with the scalar
// device path (tsgpu_flat_distances), prints the largest absolute / relative deviation and the kernel time. No torch, no Python:
//   g++ -O2 -std=c++17 tests/cpp/flat_tc_check.cpp -Iinclude -Ltypesense_b200 -ltsgpu -Wl,-rpath,'$ORIGIN/../../typesense_b200' -o tests/cpp/flat_tc_check
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "tsgpu.h"

static int run_case(tsgpu_index* idx, const std::vector<float>& vec, uint32_t n, uint32_t dim, uint32_t nq, uint32_t n_ids, uint32_t seed, bool time_it) {
    std::mt19937 rng(seed);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> q((size_t) nq * dim);
    for(uint32_t i = 0; i < nq; i++) {
        // half of the queries are perturbed copies of stored vectors (small distances: the hard case for a relative tolerance)
        double nn = 0;
        const uint32_t src = rng() % n;
        for(uint32_t e = 0; e < dim; e++) { float x = (i & 1) ? vec[(size_t) src * dim + e] + 0.02f * nd(rng) : nd(rng); q[(size_t) i * dim + e] = x; nn += (double) x * x; }
        for(uint32_t e = 0; e < dim; e++) q[(size_t) i * dim + e] = (float) (q[(size_t) i * dim + e] / std::sqrt(nn));
    }
    std::vector<uint32_t> ids;
    {   // ascending sample, plus one id beyond the index (the scalar kernel answers 0 there)
        std::vector<uint32_t> all(n); for(uint32_t i = 0; i < n; i++) all[i] = i;
        std::shuffle(all.begin(), all.end(), rng);
        all.resize(n_ids); std::sort(all.begin(), all.end());
        ids = all;
        if(!time_it) ids.back() = n + 5;
    }
    std::vector<float> out((size_t) nq * n_ids, -7.f);
    if(tsgpu_flat_distances_batch(idx, q.data(), nq, ids.data(), ids.size(), out.data()) != TSGPU_OK) { printf("FAIL call: %s\n", tsgpu_last_error()); return 1; }
    tsgpu_stats st{};
    tsgpu_get_stats(idx, &st);
    double max_abs = 0, max_rel = 0; size_t bad = 0;
    const uint32_t check_q = time_it ? std::min(nq, 8u) : nq;
    for(uint32_t qi = 0; qi < check_q; qi++) {
        for(size_t i = 0; i < ids.size(); i++) {
            double ref = 0;
            if(ids[i] < n) { double d = 0; for(uint32_t e = 0; e < dim; e++) d += (double) q[(size_t) qi * dim + e] * vec[(size_t) ids[i] * dim + e]; ref = 1.0 - d; }
            const double got = out[(size_t) qi * ids.size() + i];
            const double ae = std::fabs(got - ref), re = ae / std::max(std::fabs(ref), 1e-30);
            max_abs = std::max(max_abs, ae);
            if(std::fabs(ref) > 1e-3) max_rel = std::max(max_rel, re);
            if(ae > 1e-4 * std::fabs(ref) + 2e-6) bad++;
        }
    }
    // scalar device path on the first query
    std::vector<float> sc(ids.size());
    tsgpu_flat_distances(idx, q.data(), ids.data(), ids.size(), sc.data());
    double max_vs_scalar = 0;
    for(size_t i = 0; i < ids.size(); i++) max_vs_scalar = std::max(max_vs_scalar, (double) std::fabs(sc[i] - out[i]));
    const double flops = 2.0 * nq * (double) n_ids * dim;
    printf("%s nq=%u ids=%u dim=%u: tc_queries=%llu max_abs=%.3g max_rel(|ref|>1e-3)=%.3g vs_scalar=%.3g bad=%zu kernel_ms=%.3f (%.1f TFLOP/s fp32-equivalent, %.1f GB/s of rows)\n",
           bad ? "FAIL" : "ok  ", nq, n_ids, dim, (unsigned long long) st.flat_tc_queries, max_abs, max_rel, max_vs_scalar, bad, st.ms_knn,
           st.ms_knn > 0 ? flops / (st.ms_knn * 1e-3) / 1e12 : 0.0, st.ms_knn > 0 ? (double) n_ids * dim * 4 / (st.ms_knn * 1e-3) / 1e9 : 0.0);
    return bad ? 1 : 0;
}

static int run_dim(uint32_t n, uint32_t dim, bool big) {
    std::mt19937 rng(1234 + dim);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> vec((size_t) n * dim);
    for(uint32_t i = 0; i < n; i++) {
        double nn = 0;
        for(uint32_t e = 0; e < dim; e++) { vec[(size_t) i * dim + e] = nd(rng); nn += (double) vec[(size_t) i * dim + e] * vec[(size_t) i * dim + e]; }
        for(uint32_t e = 0; e < dim; e++) vec[(size_t) i * dim + e] = (float) (vec[(size_t) i * dim + e] / std::sqrt(nn));
    }
    tsgpu_index* idx = nullptr;
    if(tsgpu_index_create(n, 0, &idx) != TSGPU_OK) { printf("FAIL create: %s\n", tsgpu_last_error()); return 1; }
    // a graph without links: the flat scan only reads the vectors
    std::vector<uint8_t> levels(n, 0);
    std::vector<uint32_t> links0((size_t) n * 33, 0);
    std::vector<uint64_t> upper_off((size_t) n + 1, 0);
    uint32_t dummy_up = 0;
    tsgpu_hnsw g{};
    g.n_nodes = n; g.dim = dim; g.M = 16; g.max_level = 0; g.entry_point = 0; g.metric = 0;
    g.vectors = vec.data(); g.labels = nullptr; g.levels = levels.data(); g.links0 = links0.data(); g.upper_off = upper_off.data(); g.links_up = &dummy_up;
    if(tsgpu_index_load_hnsw(idx, &g) != TSGPU_OK) { printf("FAIL load: %s\n", tsgpu_last_error()); return 1; }
    int rc = 0;
    rc |= run_case(idx, vec, n, dim, 8, 128, 1, false);
    rc |= run_case(idx, vec, n, dim, 37, 1000, 2, false);
    rc |= run_case(idx, vec, n, dim, 205, 5000, 3, false);
    rc |= run_case(idx, vec, n, dim, 300, 3001, 4, false);       // two query tiles, ragged last row tile
    if(big) {
        for(int rep = 0; rep < 2; rep++) rc |= run_case(idx, vec, n, dim, 256, n, 5, true);
        rc |= run_case(idx, vec, n, dim, 64, n, 6, true);
        rc |= run_case(idx, vec, n, dim, 16, n, 7, true);
    }
    tsgpu_index_destroy(idx);
    return rc;
}

int main(int argc, char** argv) {
    const uint32_t n_big = argc > 1 ? (uint32_t) atoi(argv[1]) : 200000;
    int rc = 0;
    rc |= run_dim(20000, 128, false);
    rc |= run_dim(20000, 96, false);
    rc |= run_dim(n_big, 768, true);
    printf(rc ? "FLAT_TC_CHECK FAILED\n" : "FLAT_TC_CHECK PASSED\n");
    return rc;
}
