// TEST INFRASTRUCTURE: driver for tools/sanitize_cpu.sh (ASan/UBSan over art_mirror.hpp and art_device.cuh).
// standalone sanitizer driver: random vocabularies, host walk vs device-function walk, built mirror
#include <cstdio>
#include <algorithm>
#include <cstdint>
#include <random>
#include <string>
#include <vector>
#include <dlfcn.h>
extern "C" {
void* am_build(const char*, const int64_t*, const uint32_t*, uint32_t);
void am_free(void*);
size_t am_walk(void*, int, const char*, int, int, int, int32_t*, size_t, int*);
size_t am_fuzzy(void*, const char*, int, int, size_t, int, int, const char*, const uint32_t*, size_t, int, const char*, char*, size_t);
void am_bind(void*, const char*, const uint64_t*, const uint32_t*);
}
int main() {
    std::mt19937 rng(5);
    long n = 0;
    for(int trial = 0; trial < 60; trial++) {
        const std::string alpha = trial % 2 ? "abcde" : "abcdefghijklmnop";
        std::vector<std::string> words;
        int nw = 5 + rng() % (trial % 5 == 4 ? 2000 : 150);
        for(int i = 0; i < nw; i++) { std::string w; int L = 1 + rng() % (trial % 3 == 0 ? 20 : 8); for(int k = 0; k < L; k++) w.push_back(alpha[rng() % alpha.size()]); words.push_back(w); }
        std::sort(words.begin(), words.end()); words.erase(std::unique(words.begin(), words.end()), words.end());
        std::string nl; std::vector<int64_t> sc; std::vector<uint32_t> df; std::vector<uint64_t> lo(1, 0); std::vector<uint32_t> ids;
        for(auto& w: words) { nl += w + "\n"; sc.push_back(rng() % 50); uint32_t d = 1 + rng() % 5; df.push_back(d); uint32_t id = rng() % 7; for(uint32_t k = 0; k < d; k++) { ids.push_back(id); id += 1 + rng() % 9; } lo.push_back(ids.size()); }
        void* h = am_build(nl.c_str(), sc.data(), df.data(), (uint32_t) words.size());
        am_bind(h, nl.c_str(), lo.data(), ids.data());
        std::vector<int32_t> a(1 << 15), b(1 << 15);
        char out[1 << 16];
        for(int q = 0; q < 200; q++) {
            std::string term = words[rng() % words.size()];
            if(rng() % 2 && term.size() > 1) term[rng() % term.size()] = alpha[rng() % alpha.size()];
            if(rng() % 3 == 0) term = term.substr(0, 1 + rng() % term.size());
            if(rng() % 4 == 0) term.insert(rng() % (term.size() + 1), 1, alpha[rng() % alpha.size()]);
            int cost = rng() % 3, lo_c = rng() % 2 ? cost : rng() % (cost + 1), prefix = rng() % 2, so = 0;
            size_t na = am_walk(h, 0, term.c_str(), lo_c, cost, prefix, a.data(), a.size(), &so);
            size_t nb = am_walk(h, 1, term.c_str(), lo_c, cost, prefix, b.data(), b.size(), &so);
            if(so == 2) continue;
            if(na != nb || so) { printf("MISMATCH count %s\n", term.c_str()); return 1; }
            for(size_t i = 0; i < na; i++) if(a[i] != b[i]) { printf("MISMATCH %s\n", term.c_str()); return 1; }
            std::string prev = rng() % 3 == 0 ? words[rng() % words.size()] : "";
            am_fuzzy(h, term.c_str(), lo_c, cost, 1 + rng() % 10, rng() % 2, prefix, prev.c_str(), nullptr, 0, 0, "", out, sizeof out);
            n++;
        }
        am_free(h);
    }
    printf("ok %ld searches\n", n);
    return 0;
}
