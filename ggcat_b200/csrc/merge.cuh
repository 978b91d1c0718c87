// merge.cuh -- phase 2 (per-unit k-mer merge) kernels.
//
// Replaces, for one merge unit (bucket, second_bucket), the reference's
//   crates/assembler_kmers_merge/src/unitigs_extender/hashmap.rs:361-409  add_sequence()
//   crates/structs/src/map_entry.rs:33-84                                 MapEntry counter/flags/halving
//   crates/hashes/src/base/cn_seqhash_base.rs:22-70,100-116 (+ fw_seqhash) k-mer identity hash
// by  expand -> LSD radix sort -> run-length reduce -> multiplicity filter  (SURVEY.md A.6):
//   record  = (canonical k-mer << 2) | flag bits, one per k-mer occurrence of every super-k-mer;
//   sort    = stable LSD radix sort, 8-bit digits, per-warp histograms + match_any ranking;
//   reduce  = head flag on key change, run length = MapEntry counter, OR of flag bits = MapEntry
//             flags, multiplicity = counter >> (flags == 3), keep iff multiplicity >= -s.
// One CTA owns one unit.  Units whose records fit in shared memory are sorted there (no DRAM
// round trips at all); larger units run the same code over a global-memory scratch pair.
#pragma once
#include "bucketing.cuh"

namespace ggb {

// A bucket chunk as the merge kernels see it (one per push_reads / imported slice).
struct ChunkView {
    const uint4 *desc;           // descriptors, unit-sorted
    const uint32_t *payload;     // packed bases
    const uint32_t *unit_off;    // [n_units + 1] descriptor offsets, relative to `desc`
    const uint32_t *unit_kmers;  // [n_units]
    uint32_t first_unit;         // units [first_unit, first_unit + n_units) are present
    uint32_t n_units;
    uint32_t word_bias;          // subtract from descriptor payload offsets (imported slices)
    uint32_t pad;
};

struct MergeOut {
    uint64_t *keys;                     // survivors: canonical k-mer (2k bits)
    uint32_t *count_flags;              // multiplicity (30 bits, saturating) | flags << 30
    unsigned long long *cursor;         // [0] entries written, [1] distinct keys, [2] k-mer occurrences
    uint64_t *unit_out_off;             // per unit (relative to first unit of the launch): offset
    uint32_t *unit_out_cnt;             //                                                   count
    uint64_t capacity;                  // entries available in keys/count_flags
    uint32_t *overflow;                 // set to 1 if capacity was exceeded
};

// ------------------------------------------------------------------------------------------------
// Expansion of one super-k-mer into records (k <= 31: 62-bit key + 2 flag bits in one u64).
// Rolling update = crates/hashes/src/base/cn_seqhash_base.rs:52-69 roll_hash; flag bits =
// hashmap.rs:385-399 (begin_ignored << !is_forward) | (end_ignored << is_forward).
template <typename OutP>
__device__ __forceinline__ void expand_superkmer64(const uint32_t *__restrict__ pl, uint32_t len, uint32_t flags,
                                                   uint32_t k, uint32_t forward_only, OutP out) {
    const uint64_t mask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
    uint64_t fw = extract64(pl, 0) & mask;
    uint64_t rc = revcomp64(fw) >> (64 - 2 * k);
    const uint32_t last = len - k;
    uint32_t cw = 0;
    for (uint32_t i = 0;; ++i) {
        const bool isf = forward_only ? true : (fw < rc);
        const uint64_t key = forward_only ? fw : (fw < rc ? fw : rc);
        const uint32_t bi = (!(flags & READ_FLAG_INCL_BEGIN) && i == 0) ? 1u : 0u;
        const uint32_t ei = (!(flags & READ_FLAG_INCL_END) && i == last) ? 1u : 0u;
        const uint32_t fb = (bi << (isf ? 0 : 1)) | (ei << (isf ? 1 : 0));
        out[i] = (key << 2) | fb;
        if (i == last) break;
        const uint32_t nb = i + k;  // next base index
        if ((nb & 15u) == 0 || i == 0) cw = pl[nb >> 4];
        const uint64_t b = (cw >> (2u * (nb & 15u))) & 3u;
        fw = (fw >> 2) | (b << (2 * (k - 1)));
        rc = ((rc << 2) | (b ^ 2ull)) & mask;
    }
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of n u64 records living in A (B = scratch of the same size); generic
// pointers, so A/B may be shared or global memory.  hist: [WARPS][256] u32 in shared memory.
// Returns the buffer that holds the sorted records.
template <int THREADS>
__device__ uint64_t *block_radix_sort64(uint64_t *A, uint64_t *B, uint32_t n, uint32_t first_bit, uint32_t end_bit,
                                        uint32_t *hist, uint32_t *s_scan) {
    constexpr int WARPS = THREADS / 32;
    constexpr int EPT = WARPS * 256 / THREADS;  // scan entries per thread (= 8)
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    uint32_t chunk = (n + WARPS - 1) / WARPS;
    chunk = (chunk + 31u) & ~31u;
    const uint32_t wbeg = min(n, warp * chunk), wend = min(n, wbeg + chunk);
    for (uint32_t shift = first_bit; shift < end_bit; shift += 8) {
        for (uint32_t i = tid; i < WARPS * 256; i += THREADS) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = wbeg + lane; i < wend; i += 32) {
            const uint32_t d = (uint32_t)(A[i] >> shift) & 255u;
            atomicAdd(&hist[warp * 256 + d], 1u);
        }
        __syncthreads();
        // exclusive scan in (digit, warp) order
        {
            uint32_t v[EPT];
            uint32_t sum = 0;
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const uint32_t e = tid * EPT + q;
                v[q] = hist[(e % WARPS) * 256 + (e / WARPS)];
                sum += v[q];
            }
            uint32_t tot;
            uint32_t p = block_exclusive_scan<THREADS>(sum, s_scan, &tot);
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const uint32_t e = tid * EPT + q;
                hist[(e % WARPS) * 256 + (e / WARPS)] = p;
                p += v[q];
            }
        }
        __syncthreads();
        for (uint32_t base = wbeg; base < wend; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < wend;
            const uint64_t rec = valid ? A[i] : 0ull;
            const uint32_t d = valid ? ((uint32_t)(rec >> shift) & 255u) : (256u + lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t pos = 0;
            if (valid) pos = hist[warp * 256 + d] + rank;
            __syncwarp();
            if (valid) {
                B[pos] = rec;
                if (rank + 1 == (uint32_t)__popc(peers)) hist[warp * 256 + d] = pos + 1;
            }
            __syncwarp();
        }
        __syncthreads();
        uint64_t *t = A; A = B; B = t;
    }
    return A;
}

// ------------------------------------------------------------------------------------------------
// Run-length reduce + filter of sorted records S[0..n) (aux: u32 scratch of >= n entries that does
// not alias S).  Survivors go to out.keys/out.count_flags at an atomically reserved range.
template <int THREADS>
__device__ void block_reduce_filter(const uint64_t *S, uint32_t *aux, uint32_t n, uint32_t min_mult, const MergeOut &out,
                                    uint32_t unit_rel, uint32_t *s_scan, unsigned long long *s_base) {
    const uint32_t tid = threadIdx.x;
    uint32_t my_keep = 0, my_heads = 0;
    for (uint32_t i = tid; i < n; i += THREADS) {
        const uint64_t r = S[i];
        const uint64_t key = r >> 2;
        uint32_t cf = 0;
        if (i == 0 || (S[i - 1] >> 2) != key) {
            uint32_t cnt = 1, fl = (uint32_t)r & 3u;
            for (uint32_t j = i + 1; j < n; ++j) {
                const uint64_t q = S[j];
                if ((q >> 2) != key) break;
                ++cnt;
                fl |= (uint32_t)q & 3u;
            }
            // structs/src/map_entry.rs:79-84 get_kmer_multiplicity
            const uint32_t mult = cnt >> ((fl == (READ_FLAG_INCL_BEGIN | READ_FLAG_INCL_END)) ? 1 : 0);
            ++my_heads;
            if (mult >= min_mult) {
                cf = (mult > 0x3FFFFFFFu ? 0x3FFFFFFFu : mult) | (fl << 30);
                ++my_keep;
            }
        }
        aux[i] = cf;
    }
    uint32_t tot;
    block_exclusive_scan<THREADS>(my_keep, s_scan, &tot);  // tot = survivors
    uint32_t tot_heads;
    block_exclusive_scan<THREADS>(my_heads, s_scan, &tot_heads);
    if (tid == 0) {
        unsigned long long b = atomicAdd(&out.cursor[0], (unsigned long long)tot);
        atomicAdd(&out.cursor[1], (unsigned long long)tot_heads);
        atomicAdd(&out.cursor[2], (unsigned long long)n);
        *s_base = b;
        out.unit_out_off[unit_rel] = b;
        out.unit_out_cnt[unit_rel] = tot;
        if (b + tot > out.capacity) *out.overflow = 1u;
    }
    __syncthreads();
    const unsigned long long gbase = *s_base;
    if (gbase + tot > out.capacity) return;  // overflow reported; nothing written
    uint32_t running = 0;
    for (uint32_t base = 0; base < n; base += THREADS) {
        const uint32_t i = base + tid;
        const uint32_t cf = i < n ? aux[i] : 0u;
        uint32_t t2;
        const uint32_t p = block_exclusive_scan<THREADS>(cf ? 1u : 0u, s_scan, &t2);
        if (cf) {
            out.keys[gbase + running + p] = S[i] >> 2;
            out.count_flags[gbase + running + p] = cf;
        }
        running += t2;
    }
}

// ------------------------------------------------------------------------------------------------
// k_merge_units: one CTA per work item (a unit id).  Template CAP = record capacity of the
// shared-memory path; units with more records use scratch (global) if provided, else are skipped
// (they are on the other work list).
template <int THREADS, int CAP, bool GLOBAL_SCRATCH>
__global__ void __launch_bounds__(THREADS)
k_merge_units(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ work,
              uint32_t n_work, uint32_t first_unit, DevParams P, uint32_t min_mult, MergeOut out,
              uint64_t *__restrict__ scratch, const uint64_t *__restrict__ scratch_off) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WARPS = THREADS / 32;
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);                 // WARPS*256 u32
    uint32_t *s_scan = hist + WARPS * 256;                                    // 40 u32
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_scan + 40);
    uint64_t *sA = reinterpret_cast<uint64_t *>(s_base + 2);
    uint64_t *sB = sA + (GLOBAL_SCRATCH ? 0 : CAP);

    const uint32_t tid = threadIdx.x;
    for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const uint32_t unit = work[wi];
        uint32_t n = 0;
        for (uint32_t c = 0; c < n_chunks; c++) {
            const ChunkView &cv = chunks[c];
            if (unit >= cv.first_unit && unit < cv.first_unit + cv.n_units) n += cv.unit_kmers[unit - cv.first_unit];
        }
        if (!GLOBAL_SCRATCH && n > (uint32_t)CAP) {  // host routes such units to the scratch variant
            if (tid == 0) *out.overflow = 2u;
            continue;
        }
        uint64_t *A, *B;
        if (GLOBAL_SCRATCH) {
            A = scratch + scratch_off[wi];
            B = A + n;
        } else {
            A = sA; B = sB;
        }
        // ---- expand
        uint32_t running = 0;
        for (uint32_t c = 0; c < n_chunks; c++) {
            const ChunkView cv = chunks[c];
            if (unit < cv.first_unit || unit >= cv.first_unit + cv.n_units) continue;
            const uint32_t d0 = cv.unit_off[unit - cv.first_unit], d1 = cv.unit_off[unit - cv.first_unit + 1];
            for (uint32_t base = d0; base < d1; base += THREADS) {
                const uint32_t di = base + tid;
                uint4 d = make_uint4(0, 0, 0, 0);
                uint32_t cnt = 0;
                if (di < d1) { d = cv.desc[di]; cnt = d.y - P.k + 1; }
                uint32_t tot;
                const uint32_t p = block_exclusive_scan<THREADS>(cnt, s_scan, &tot);
                if (cnt) expand_superkmer64(cv.payload + (d.x - cv.word_bias), d.y, (d.z >> 16) & 3u, P.k, P.forward_only,
                                            A + running + p);
                running += tot;
            }
        }
        __syncthreads();
        // ---- sort (digits above bit 2k+2 are all zero)
        const uint32_t end_bit = min(64u, (2 * P.k + 2 + 7) & ~7u);
        uint64_t *S = block_radix_sort64<THREADS>(A, B, n, 0, end_bit, hist, s_scan);
        uint64_t *other = (S == A) ? B : A;
        // ---- reduce + filter
        block_reduce_filter<THREADS>(S, reinterpret_cast<uint32_t *>(other), n, min_mult, out, unit - first_unit, s_scan,
                                     s_base);
        __syncthreads();
    }
}

template <int THREADS, int CAP>
constexpr size_t merge_smem_bytes(bool global_scratch) {
    return (size_t)(THREADS / 32) * 256 * 4 + 40 * 4 + 16 + (global_scratch ? 0 : (size_t)CAP * 16);
}

}  // namespace ggb
