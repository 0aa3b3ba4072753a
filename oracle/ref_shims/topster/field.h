// TEST INFRASTRUCTURE. Stands in for the reference's include/field.h where include/topster.h is compiled on its own (oracle/Makefile,
// target ref): field.h pulls in s2, ICU and the JSON library, none of which the Topster needs. What the Topster does need from that
// include chain: sparsepp (vendored by the reference) and the two hashing helpers of StringUtils (include/string_utils.h:316-326) that
// LogLogBeta and Union_KV call — declared here with the reference's own wyhash underneath.
#pragma once
#include <cstdint>
#include <limits>
#include <string>
#include "sparsepp.h"
#include "wyhash_v5.h"

struct StringUtils {
    static uint64_t hash_wy(const void* key, uint64_t len) {
        const uint64_t h = wyhash(key, len, 0, _wyp);
        return h != std::numeric_limits<uint64_t>::max() ? h : std::numeric_limits<uint64_t>::max() - 1;
    }
    static constexpr uint64_t hash_combine(uint64_t combined, uint64_t hash) {
        combined ^= hash + 0x517cc1b727220a95 + (combined << 6) + (combined >> 2);
        return combined;
    }
};
