// K7 on the 5th-generation tensor cores: process_results_bruteforce (src/index.cpp:3345-3374) for a GROUP of queries that share
// one candidate set (the same filter): dist[q][i] = 1 - <query q, vector ids[i]>  is the GEMM  D[ids x queries] = V[ids, dim] . Q^T.
//
//   * tcgen05.mma kind::tf32, M = 128 candidate rows per tile, N = up to 256 queries, accumulators in TMEM (all 512 columns: a pair
//     per tile — large terms / cross terms — and, for tiles of <= 128 queries, a second pair so that the epilogue of one tile
//     overlaps the MMAs of the next), one elected thread issues.
//   * fp32 fidelity by a 3-term split (x = hi + lo, hi = the 10 leading mantissa bits tf32 keeps):  hi.hi + lo.hi + hi.lo,
//     accumulated in fp32 in TMEM — error ~2^-21 per product, the same class as a re-ordered fp32 sum (the parity gate is
//     1e-4 relative, tests/test_flat_tc_gpu.py; a single tf32 or bf16 product would miss it by two orders of magnitude).
//   * B (the queries): split and laid out ONCE per call by flat_tc_pack_queries_kernel as ready-made 128B-swizzled K-major tile
//     images; each pipeline stage's image is fetched by ONE cp.async.bulk (TMA engine) that completes on the stage's mbarrier.
//   * A (the candidates' rows): gathered — ids are arbitrary, so there is no box a tensor map could describe. Eight producer warps
//     read 128 B of 128 rows per k-block (a quarter-warp per row: full 128-byte lines), split hi / lo in registers and store both
//     halves in the swizzled layout (conflict-free: a quarter-warp's 8 chunks cover the 32 banks once); next k-block's loads are
//     in flight while the current one is split and stored.
//   * epilogue warps 8-11: tcgen05.ld 32 lanes x 16 columns, dist = 1 - acc, one coalesced 128-byte store per query column
//     (lanes = consecutive candidates).
//
// A translation unit of its own (like art_kernels.cu): tsgpu.cu calls tsgpu_flat_tc_run_ below.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/tsgpu.h"

namespace tsft {

constexpr int kRows = 128;            // candidates per tile (UMMA M)
constexpr int kKB = 32;               // fp32 elements per k-block: one 128-byte swizzle row
constexpr int kProdWarps = 8;         // row producers: 4 rows of the tile per thread and k-block, two k-blocks of loads in flight
constexpr int kThreads = 448;         // warps 0-7 row producers, 8-11 epilogue, 12 MMA issuer, 13 query-tile copies + TMEM owner
constexpr int kMaxStages = 4;
constexpr int kPrefetch = 4;           // k-blocks of row loads in flight per producer thread
constexpr uint32_t kABytes = kRows * 128;       // one half (hi or lo) of a stage's A tile
constexpr uint32_t kTmemCols = 512;             // two accumulators of up to 256 fp32 columns

struct Params {
    const float* vectors;             // [n_nodes * dim]
    uint32_t n_nodes, dim;
    const uint32_t* ids;              // [n_ids] the shared candidate set
    uint32_t n_ids;
    uint32_t ng;                      // queries in the group
    uint32_t n_tile;                  // queries per tile: multiple of 16, <= 256
    uint32_t n_qtiles, n_mtiles, kblocks, stages;
    const unsigned char* b_img;       // [n_qtiles][kblocks][hi | lo][n_tile rows x 128 B]
    const unsigned long long* out_off;// [ng] start of each query's range in out_dist
    float* out_dist;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while(!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {      // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), start address in
// 16-byte units; the k-step inside the 128-byte row is an offset on the start address (the hardware swizzles absolute address bits,
// so tiles are 1024-byte aligned).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    return (uint64_t) ((addr & 0x3FFFFu) >> 4)          // start address
         | ((uint64_t) 1 << 16)                         // leading byte offset (unused with swizzled K-major; 1 by convention)
         | ((uint64_t) (1024 >> 4) << 32)               // stride byte offset: 8 rows x 128 B
         | ((uint64_t) 1 << 46)                         // descriptor version (Blackwell)
         | ((uint64_t) 2 << 61);                        // layout: SWIZZLE_128B
}
// byte offset of 16-byte chunk c (0..7) of row r inside a tile whose rows are 128 B: Swizzle<3,4,3>
__device__ __host__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c) { return (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4); }

__device__ __forceinline__ void split_tf32(const float4& v, uint4& hi, uint4& lo) {
    const uint32_t m = 0xFFFFE000u;      // sign, exponent, 10 mantissa bits: what kind::tf32 reads
    hi.x = __float_as_uint(v.x) & m; hi.y = __float_as_uint(v.y) & m; hi.z = __float_as_uint(v.z) & m; hi.w = __float_as_uint(v.w) & m;
    lo.x = __float_as_uint(v.x - __uint_as_float(hi.x)) & m; lo.y = __float_as_uint(v.y - __uint_as_float(hi.y)) & m;
    lo.z = __float_as_uint(v.z - __uint_as_float(hi.z)) & m; lo.w = __float_as_uint(v.w - __uint_as_float(hi.w)) & m;
}

// Query tiles as the MMA reads them. One thread per 16-byte chunk.
__global__ void __launch_bounds__(256)
flat_tc_pack_queries_kernel(const float* __restrict__ queries, const uint32_t* __restrict__ qsel, uint32_t ng, uint32_t dim,
                            uint32_t n_tile, uint32_t n_qtiles, unsigned char* __restrict__ img) {
    const uint32_t kblocks = dim / kKB;
    const size_t total = (size_t) n_qtiles * kblocks * n_tile * 8;
    for(size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t) (i & 7);
        const uint32_t n = (uint32_t) ((i >> 3) % n_tile);
        const size_t tk = (i >> 3) / n_tile;                 // qt * kblocks + kb
        const uint32_t kb = (uint32_t) (tk % kblocks), qt = (uint32_t) (tk / kblocks);
        const uint32_t gq = qt * n_tile + n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if(gq < ng) {
            const uint32_t q = qsel ? qsel[gq] : gq;
            v = *reinterpret_cast<const float4*>(queries + (size_t) q * dim + kb * kKB + c * 4);
        }
        uint4 hi, lo;
        split_tf32(v, hi, lo);
        unsigned char* tile = img + tk * 2 * (size_t) n_tile * 128;
        const uint32_t off = sw128_off(n, c);
        *reinterpret_cast<uint4*>(tile + off) = hi;
        *reinterpret_cast<uint4*>(tile + (size_t) n_tile * 128 + off) = lo;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
flat_tc_kernel(const __grid_constant__ Params P) {
    extern __shared__ unsigned char smem_tc_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * kMaxStages + 4];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned long long col_off[256];

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = P.stages, KB = P.kblocks, NT = P.n_tile;
    const uint32_t stage_bytes = 2 * kABytes + NT * 256;
    const uint32_t nbuf = NT <= 128 ? 2u : 1u;          // accumulator PAIRS in the 512 TMEM columns: two pairs (double-buffered against the epilogue) up to 128 queries per tile
    const uint32_t accw = NT <= 128 ? 128u : 256u;      // columns between a pair's two accumulators
    const uint32_t smem0 = (smem_u32(smem_tc_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8 * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8 * (kMaxStages + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8 * (2 * kMaxStages + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8 * (2 * kMaxStages + 2 + a); };

    if(warp == 12 && lane == 0) {
        for(uint32_t s = 0; s < S; s++) { mbar_init(full_bar(s), kProdWarps * 32 + 1); mbar_init(empty_bar(s), 1); }
        for(uint32_t a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if(warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t n_tiles = P.n_mtiles * P.n_qtiles;

    if(warp < kProdWarps) {
        // ===== row producers: A tile of the stage = 128 gathered rows x 32 floats, hi and lo halves
        const uint32_t c = lane & 7, r0 = warp * 4 + (lane >> 3);
        uint32_t it = 0;
        for(uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const uint32_t mt = t / P.n_qtiles;
            const float* ptr[4];
            bool ok[4];
#pragma unroll
            for(int p = 0; p < 4; p++) {
                const uint32_t gid = mt * kRows + p * 32 + r0;
                const uint32_t id = gid < P.n_ids ? __ldg(P.ids + gid) : 0xFFFFFFFFu;
                ok[p] = id < P.n_nodes;
                ptr[p] = P.vectors + (size_t) (ok[p] ? id : 0u) * P.dim + c * 4;
            }
            // kPrefetch k-blocks of loads in flight per thread (a register ring with compile-time slots): the gather is latency-bound,
            // bytes in flight per SM = kPrefetch x 16 KB
            float4 ring[kPrefetch][4];
#pragma unroll
            for(int d = 0; d < kPrefetch; d++) {
#pragma unroll
                for(int p = 0; p < 4; p++)
                    ring[d][p] = (ok[p] && (uint32_t) d < KB) ? __ldg(reinterpret_cast<const float4*>(ptr[p] + (size_t) d * kKB)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for(uint32_t kb0 = 0; kb0 < KB; kb0 += kPrefetch) {
#pragma unroll
                for(int d = 0; d < kPrefetch; d++) {
                    const uint32_t kb = kb0 + d;
                    if(kb < KB) {
                        const uint32_t s = it % S, ph = (it / S) & 1u;
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        const uint32_t a_hi = smem0 + s * stage_bytes, a_lo = a_hi + kABytes;
#pragma unroll
                        for(int p = 0; p < 4; p++) {
                            uint4 hi, lo;
                            split_tf32(ring[d][p], hi, lo);
                            const uint32_t off = sw128_off(p * 32 + r0, c);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(a_hi + off), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(a_lo + off), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
                        }
                        fence_async_smem();             // generic-proxy stores -> visible to the tensor core's async proxy
                        mbar_arrive(full_bar(s));
                        it++;
                        const uint32_t kn = kb + kPrefetch;      // refill the slot just consumed
                        if(kn < KB) {
#pragma unroll
                            for(int p = 0; p < 4; p++)
                                if(ok[p]) ring[d][p] = __ldg(reinterpret_cast<const float4*>(ptr[p] + (size_t) kn * kKB));
                        }
                    }
                }
            }
        }
    } else if(warp == 13) {
        // ===== query tiles: one bulk copy per stage (hi and lo images are adjacent)
        if(lane == 0) {
            uint32_t it = 0;
            const uint32_t b_bytes = NT * 256;
            for(uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const uint32_t qt = t % P.n_qtiles;
                for(uint32_t kb = 0; kb < KB; kb++, it++) {
                    const uint32_t s = it % S, ph = (it / S) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), b_bytes);
                    bulk_g2s(smem0 + s * stage_bytes + 2 * kABytes, P.b_img + ((size_t) qt * KB + kb) * b_bytes, b_bytes, full_bar(s));
                }
            }
        }
    } else if(warp == 12) {
        // ===== MMA issuer
        const uint32_t idesc = (1u << 4)                 // D: fp32
                             | (2u << 7) | (2u << 10)    // A, B: tf32, both K-major
                             | ((NT >> 3) << 17) | ((uint32_t) (kRows >> 4) << 24);
        uint32_t it = 0, ti = 0;
        for(uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ti++) {
            const uint32_t acc = ti % nbuf, aph = (ti / nbuf) & 1u;
            mbar_wait(tempty_bar(acc), aph ^ 1u);        // the epilogue has drained this accumulator pair
            tc_fence_after();
            // two accumulators per tile: the hi.hi products in one, the two cross terms (2^-11 of it) in the other. The tensor core
            // truncates the accumulator at every step; kept apart, the 2/3 of the steps that add small terms no longer round the
            // large sum (measured: 1.1e-5 -> see tests/cpp/flat_tc_check.cpp for the deviation now)
            const uint32_t d_big = tmem_base + acc * 256u, d_small = d_big + accw;
            for(uint32_t kb = 0; kb < KB; kb++, it++) {
                const uint32_t s = it % S, ph = (it / S) & 1u;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                if(lane == 0) {
                    const uint32_t a_hi = smem0 + s * stage_bytes, a_lo = a_hi + kABytes, b_hi = a_lo + kABytes, b_lo = b_hi + NT * 128;
#pragma unroll
                    for(uint32_t k = 0; k < kKB / 8; k++) {          // UMMA K = 8 tf32 = 32 bytes of the 128-byte row
                        const uint64_t dah = smem_desc_sw128(a_hi + k * 32), dal = smem_desc_sw128(a_lo + k * 32);
                        const uint64_t dbh = smem_desc_sw128(b_hi + k * 32), dbl = smem_desc_sw128(b_lo + k * 32);
                        mma_tf32(d_big, dah, dbh, idesc, (kb | k) != 0u);
                        mma_tf32(d_small, dal, dbh, idesc, (kb | k) != 0u);
                        mma_tf32(d_small, dah, dbl, idesc, 1u);
                    }
                    tc_commit(empty_bar(s));                         // the stage is free once these MMAs have read it
                    if(kb + 1 == KB) tc_commit(tfull_bar(acc));      // the accumulators are complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue (warps 8..11 own TMEM lanes 32*(warp-8) ..): dist = 1 - dot, one column = one query
        const uint32_t ew = warp - 8;
        uint32_t ti = 0, col_qt = 0xFFFFFFFFu;
        for(uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ti++) {
            const uint32_t acc = ti % nbuf, aph = (ti / nbuf) & 1u;
            const uint32_t mt = t / P.n_qtiles, qt = t % P.n_qtiles;
            const uint32_t gid = mt * kRows + ew * 32 + lane;
            const bool in_set = gid < P.n_ids;
            const bool ok = in_set && __ldg(P.ids + gid) < P.n_nodes;
            // this tile's output offsets, staged while the MMAs run (the four epilogue warps share the table; named barrier 1)
            if(qt != col_qt) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for(uint32_t j = threadIdx.x - 8 * 32; j < NT; j += 128) col_off[j] = (qt * NT + j < P.ng) ? P.out_off[qt * NT + j] : ~0ull;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                col_qt = qt;
            }
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((ew * 32u) << 16) + acc * 256u;
            for(uint32_t c0 = 0; c0 < NT; c0 += 16) {
                uint32_t r[16], q[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr + c0) : "memory");
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                               "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
                             : "r"(taddr + accw + c0) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if(in_set) {
#pragma unroll
                    for(uint32_t j = 0; j < 16; j++) {
                        const unsigned long long o = col_off[c0 + j];            // shared memory: this tile's columns, ~0ull = beyond the group
                        if(o != ~0ull) P.out_dist[o + gid] = ok ? 1.0f - (__uint_as_float(r[j]) + __uint_as_float(q[j])) : 0.f;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if(warp == 13) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

}  // namespace tsft

// ---- called by tsgpu.cu (all pointers are device memory; `stream` is the index's vector-stage stream) ---------------------------
// Bytes of scratch the packed query tiles need.
extern "C" __attribute__((visibility("hidden")))
size_t tsgpu_flat_tc_image_bytes_(uint32_t ng, uint32_t dim, uint32_t* n_tile_out, uint32_t* n_qtiles_out) {
    // <= 128 queries: one tile, accumulator pairs double-buffered. More: tiles of up to 256 queries with ONE pair in TMEM (the epilogue is
    // not overlapped) — measured faster than twice as many 128-query tiles (0.51 vs 0.57 ms for 256 queries x 200 K rows x 768), because
    // every query tile streams the rows through the producers again.
    const uint32_t n_qtiles = (ng + 255) / 256;
    uint32_t n_tile = ((ng + n_qtiles - 1) / n_qtiles + 15) & ~15u;       // balanced tiles, multiple of 16
    if(n_tile < 16) n_tile = 16;
    if(n_tile_out) *n_tile_out = n_tile;
    if(n_qtiles_out) *n_qtiles_out = n_qtiles;
    return (size_t) n_qtiles * (dim / tsft::kKB) * 2 * n_tile * 128;
}

extern "C" __attribute__((visibility("hidden")))
int tsgpu_flat_tc_supported_(uint32_t dim) { return dim >= 32 && dim % tsft::kKB == 0; }

// dist of `ng` queries (rows qsel[0..ng) of `queries`, or 0..ng-1 when qsel is null) to every id of the shared candidate set:
// out_dist[out_off[g] + i] for candidate i. `img` = scratch of tsgpu_flat_tc_image_bytes_ bytes, 1024-byte aligned.
extern "C" __attribute__((visibility("hidden")))
cudaError_t tsgpu_flat_tc_run_(const float* vectors, uint32_t n_nodes, uint32_t dim, const float* queries, const uint32_t* qsel, uint32_t ng,
                               const uint32_t* ids, uint32_t n_ids, const unsigned long long* out_off, float* out_dist,
                               unsigned char* img, int n_sms, cudaStream_t stream, int* launches) {
    using namespace tsft;
    if(ng == 0 || n_ids == 0) return cudaSuccess;
    Params P{};
    P.vectors = vectors; P.n_nodes = n_nodes; P.dim = dim;
    P.ids = ids; P.n_ids = n_ids; P.ng = ng;
    tsgpu_flat_tc_image_bytes_(ng, dim, &P.n_tile, &P.n_qtiles);
    P.n_mtiles = (n_ids + kRows - 1) / kRows;
    P.kblocks = dim / kKB;
    const uint32_t stage_bytes = 2 * kABytes + P.n_tile * 256;
    uint32_t stages = (uint32_t) ((220u * 1024u) / stage_bytes);
    if(const char* e = getenv("TSGPU_FLAT_TC_STAGES")) stages = (uint32_t) atoi(e);
    P.stages = stages < 2 ? 2 : (stages > (uint32_t) kMaxStages ? (uint32_t) kMaxStages : stages);
    P.b_img = img; P.out_off = out_off; P.out_dist = out_dist;
    const size_t smem = (size_t) P.stages * stage_bytes + 1024;
    {   // opt-in dynamic shared memory: static (barriers) + dynamic must stay within the 227 KB of a CTA
        cudaError_t e = cudaFuncSetAttribute(flat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if(e != cudaSuccess) return e;
    }
    const size_t chunks = (size_t) P.n_qtiles * P.kblocks * P.n_tile * 8;
    const unsigned pgrid = (unsigned) ((chunks + 255) / 256 < 1184 ? (chunks + 255) / 256 : 1184);
    flat_tc_pack_queries_kernel<<<pgrid, 256, 0, stream>>>(queries, qsel, ng, dim, P.n_tile, P.n_qtiles, img);
    const uint32_t n_tiles = P.n_mtiles * P.n_qtiles;
    const unsigned grid = n_tiles < (uint32_t) n_sms ? n_tiles : (unsigned) n_sms;
    flat_tc_kernel<<<grid, kThreads, smem, stream>>>(P);
    if(launches) *launches += 2;
    return cudaGetLastError();
}
