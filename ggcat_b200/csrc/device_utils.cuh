// device_utils.cuh -- bit/packing helpers and the reference's hash functions as device code.
//
// Semantics follow SURVEY.md Appendix A; each helper cites the reference file:line it mirrors
// (paths relative to /root/reference).  Everything is integer arithmetic and must stay bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace ggb {

constexpr uint64_t NT_MULTIPLIER = 0x397f178c6ae330f9ULL;  // crates/hashes/src/nthash_base.rs:49
constexpr uint32_t READ_FLAG_INCL_BEGIN = 1;                // crates/config/src/lib.rs:93
constexpr uint32_t READ_FLAG_INCL_END = 2;                  // crates/config/src/lib.rs:94

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, unsigned r) {
    r &= 63u;
    return r ? ((x << r) | (x >> (64u - r))) : x;
}

// crates/hashes/src/nthash_base.rs:51-61: h(c) = (2c+1)*MULT, rc(c) = (2(c^2)+1)*MULT on 2-bit codes
__host__ __device__ __forceinline__ uint64_t nt_h(uint32_t code) { return (uint64_t)(2u * code + 1u) * NT_MULTIPLIER; }
__host__ __device__ __forceinline__ uint64_t nt_r(uint32_t code) { return (uint64_t)(2u * (code ^ 2u) + 1u) * NT_MULTIPLIER; }

// 32 bits starting at bit position `bitpos` of a little-endian u32 word stream.
// Reads words [bitpos/32] and [bitpos/32 + 1] (callers pad their buffers by one word).
template <typename P>
__device__ __forceinline__ uint32_t extract32(P words, uint64_t bitpos) {
    const uint64_t w = bitpos >> 5;
    const uint32_t s = (uint32_t)bitpos & 31u;
    const uint32_t lo = words[w], hi = words[w + 1];
    return __funnelshift_r(lo, hi, s);
}

// 64 bits starting at `bitpos` (reads three words).
template <typename P>
__device__ __forceinline__ uint64_t extract64(P words, uint64_t bitpos) {
    const uint64_t w = bitpos >> 5;
    const uint32_t s = (uint32_t)bitpos & 31u;
    const uint32_t a = words[w], b = words[w + 1], c = words[w + 2];
    return ((uint64_t)__funnelshift_r(b, c, s) << 32) | (uint64_t)__funnelshift_r(a, b, s);
}

// 2-bit code of base `i` in a packed stream (crates/io/src/compressed_read.rs:882-885 layout:
// base i at bits 2(i%4) of byte i/4 == bits 2(i%16) of LE word i/16).
template <typename P>
__device__ __forceinline__ uint32_t packed_base(P words, uint64_t i) {
    return (words[i >> 4] >> (2u * ((uint32_t)i & 15u))) & 3u;
}

// Reverse the order of the 16 bases of a word and complement each (code ^ 2):
// crates/io/src/compressed_read.rs:621-633 (compress_from_plain_rc) on a whole word.
__host__ __device__ __forceinline__ uint32_t revcomp32(uint32_t x) {
#ifdef __CUDA_ARCH__
    uint32_t y = __brev(x);
#else
    uint32_t y = x;
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    y = ((y >> 2) & 0x33333333u) | ((y & 0x33333333u) << 2);
    y = ((y >> 4) & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4);
    y = ((y >> 8) & 0x00FF00FFu) | ((y & 0x00FF00FFu) << 8);
    y = (y >> 16) | (y << 16);
#endif
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);  // undo the swap inside each pair
    return y ^ 0xAAAAAAAAu;
}
__host__ __device__ __forceinline__ uint64_t revcomp64(uint64_t x) {
    return ((uint64_t)revcomp32((uint32_t)x) << 32) | (uint64_t)revcomp32((uint32_t)(x >> 32));
}

// ---- TMA (bulk async copy) helpers: global -> shared, completion on an mbarrier --------------------------------
// 1-D cp.async.bulk: source, destination and byte count must be multiples of 16.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// Orders earlier generic-proxy accesses of shared memory (plain ld/st by the CTA's threads, made visible to the issuing
// thread by a barrier) before later async-proxy writes (a bulk copy that reuses the same buffer).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    const uint32_t addr = smem_addr(bar);
    uint32_t done;
    unsigned long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(addr), "r"(phase) : "memory");
        if (done) break;
        // a bulk copy that never lands (bad descriptor, wrong byte count) must not hang the device: trap after ~5 s
        if ((spins & 0xFFFu) == 0xFFFu) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 5000000000ull) {
                printf("ggcat_b200: mbarrier wait timed out (block %u thread %u parity %u)\n", blockIdx.x, threadIdx.x, phase);
                __trap();
            }
        }
    }
}

// ---- shared-memory accesses by 32-bit shared-window address (no generic-pointer conversion in hot loops) ------------
// All are `asm volatile`: the compiler keeps their order relative to each other and to barriers.
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_u64x(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned long long atoms_cas64(uint32_t a, unsigned long long cmp, unsigned long long val) {
    unsigned long long old;
    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val) : "memory");
    return old;
}
__device__ __forceinline__ void atoms_inc32(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
__device__ __forceinline__ void atoms_add32(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void atoms_or32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- block-wide helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// Exclusive scan of one value per thread across the block.  `smem` needs (THREADS/32 + 1) words.
// Returns the exclusive prefix; *total receives the block sum.  Contains two __syncthreads().
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem, uint32_t *total) {
    constexpr int WARPS = THREADS / 32;
    const uint32_t lane = lane_id(), warp = warp_id();
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < WARPS ? smem[lane] : 0;
        uint32_t s = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += y;
        }
        if (lane < WARPS) smem[lane] = s - w;  // exclusive warp offsets
        if (lane == 31) smem[WARPS] = s;       // block total (WARPS <= 32)
    }
    __syncthreads();
    const uint32_t res = smem[warp] + x - v;
    *total = smem[WARPS];
    __syncthreads();  // smem may be reused right after
    return res;
}

}  // namespace ggb
