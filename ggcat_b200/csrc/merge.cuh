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
    const uint32_t *slot_of_unit;       // output slot of a unit (big units own several slots, one per key partition);
                                        // NULL: slot = unit index relative to the first unit of the launch
    __device__ __forceinline__ uint32_t slot(uint32_t unit_rel) const { return slot_of_unit ? slot_of_unit[unit_rel] : unit_rel; }
};

// Key partitions of big units (units that do not fit a shared-memory table): k_partition_units expands the unit once
// and routes every k-mer record by a hash of its key into one of P partitions in HBM (all occurrences of a k-mer
// land in the same partition, so partitions are counted independently, no fold); k_merge_hash<.., SRC_RECORDS>
// counts one partition per CTA in shared memory; k_finish_units orders the unit's survivors.
struct PartSrc {
    const uint64_t *recs;        // [n_parts_total][pcap] records (key << 2 | flag bits)
    const uint32_t *pcount;      // [n_parts_total] records per partition
    const uint32_t *part_slot;   // work item -> output slot
    const uint32_t *part_big;    // work item -> index of its big unit
    const uint32_t *big_ovf;     // [n_big] 1 = a partition overflowed, the unit is redone by the global-table kernel
    uint32_t pcap;
    uint32_t pad;
};
enum { SRC_SUPERKMERS = 0, SRC_RECORDS = 1 };

__device__ __forceinline__ uint32_t part_hash(uint64_t key) {   // independent of the in-table slot hash
    return (uint32_t)((key * 0xD6E8FEB86659FD93ull) >> 40);
}

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
template <int THREADS, bool HAS_VAL = false>
__device__ uint64_t *block_radix_sort64(uint64_t *A, uint64_t *B, uint32_t n, uint32_t first_bit, uint32_t end_bit,
                                        uint32_t *hist, uint32_t *s_scan, uint32_t *Av = nullptr, uint32_t *Bv = nullptr,
                                        uint32_t **vals_out = nullptr) {
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
                if (HAS_VAL) Bv[pos] = Av[i];
                if (rank + 1 == (uint32_t)__popc(peers)) hist[warp * 256 + d] = pos + 1;
            }
            __syncwarp();
        }
        __syncthreads();
        uint64_t *t = A; A = B; B = t;
        if (HAS_VAL) { uint32_t *tv = Av; Av = Bv; Bv = tv; }
    }
    if (HAS_VAL && vals_out) *vals_out = Av;
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
              uint64_t *__restrict__ scratch, uint64_t per_cta_u64,
              const uint32_t *__restrict__ n_work_dev = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WARPS = THREADS / 32;
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);                 // WARPS*256 u32
    uint32_t *s_scan = hist + WARPS * 256;                                    // 40 u32
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_scan + 40);
    uint64_t *sA = reinterpret_cast<uint64_t *>(s_base + 2);
    uint64_t *sB = sA + (GLOBAL_SCRATCH ? 0 : CAP);

    const uint32_t tid = threadIdx.x;
    if (n_work_dev) n_work = *n_work_dev;  // retry list filled by k_merge_hash
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
            A = scratch + (uint64_t)blockIdx.x * per_cta_u64;  // this CTA's slice: >= 2n records
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
        block_reduce_filter<THREADS>(S, reinterpret_cast<uint32_t *>(other), n, min_mult, out, out.slot(unit - first_unit), s_scan,
                                     s_base);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_merge_hash: the same unit semantics with a shared-memory hash table instead of a full sort --
// the device analogue of the reference's FxHashMap<hash, MapEntry> (hashmap.rs:385-399):
//   slot key  = canonical k-mer (CAS-claimed, linear probing, load <= 0.75),
//   slot word = MapEntry: counter in bits 0..29 (atomicAdd), flags in bits 30..31 (atomicOr).
// After all super-k-mers are inserted the table is scanned once: multiplicity = counter >> (flags==3),
// survivors (multiplicity >= -s) are ordered by key (part of the output contract) with a bin-rank sort, so the
// sort cost scales with the table that leaves the GPU, not with the k-mer occurrences.
constexpr uint64_t HASH_EMPTY = ~0ull;

// Slot of a key in a table of TS slots (any size, not only powers of two): multiply-shift on the high hash bits.
__device__ __forceinline__ uint32_t hash_slot(uint64_t key, uint32_t TS) {
    return __umulhi((uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32), TS);
}

// The counter word holds (occurrences - 1): the thread that claims a slot does not add its own occurrence, so a
// distinct k-mer costs one CAS and every further occurrence one atomicAdd (shared-memory atomics are the kernel's
// bound: ~2 cycles per lane).  Flag bits are OR-ed only when they are not set yet.
__device__ __forceinline__ void hash_insert(uint64_t *K, uint32_t *C, uint32_t TS, uint64_t key, uint32_t fb) {
    uint32_t slot = hash_slot(key, TS);
    bool claimed = false;
    while (true) {
        // most occurrences hit a key that is already there (coverage): look before the CAS
        const uint64_t cur = *reinterpret_cast<volatile uint64_t *>(&K[slot]);
        if (cur == key) break;
        if (cur == HASH_EMPTY) {
            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(&K[slot]), HASH_EMPTY, key);
            if (old == HASH_EMPTY) { claimed = true; break; }
            if (old == key) break;
        }
        slot = slot + 1 == TS ? 0u : slot + 1;
    }
    if (!claimed) atomicAdd(&C[slot], 1u);
    if (fb && ((*reinterpret_cast<volatile uint32_t *>(&C[slot]) >> 30) & fb) != fb) atomicOr(&C[slot], fb << 30);
}
__device__ __forceinline__ uint32_t slot_count(uint32_t cc) { return (cc & 0x3FFFFFFFu) + 1u; }   // occurrences of an occupied slot

// ---- load-balanced expansion of a unit ---------------------------------------------------------------------------
// The k-mer records of a unit are numbered 0..tot over all its super-k-mers (all chunks, THREADS descriptors per
// round: one block scan of the k-mer counts).  Every WARP takes an equal, 32-aligned range of records and walks it 32
// records at a time, ONE RECORD PER LANE:
//   owner super-k-mer of record r  = (super-k-mer that contains the window's first record) + number of super-k-mers
//                                    starting inside the window at or before r  (warp-wide OR of start bits + popc);
//   k-mer at offset i of the owner = 64 bits extracted at bit 2i of its payload (three cached word loads), canonical
//                                    form by one bit-reversal (cn_seqhash_base.rs:27-69 evaluated directly, no rolling).
// Every lane does the same work whatever the super-k-mer lengths are (no divergence), and all warps of the CTA finish
// within one window of each other (the unit is small: ~130 windows for 16 warps on the C2 shape).
constexpr int UNIT_MAXC = 32;   // chunks gathered per round

template <int THREADS>
struct UnitStage {                       // lives in the kernel's scratch area while records are inserted
    const uint32_t *c_pl[UNIT_MAXC];     // payload base of the chunk, word bias already subtracted
    uint32_t c_d0[UNIT_MAXC], c_cnt[UNIT_MAXC];
    uint32_t start[THREADS + 1];         // first record number of the staged super-k-mer (exclusive prefix of k-mer counts)
    uint32_t woff[THREADS];              // payload word offset inside its chunk
    uint32_t lf[THREADS];                // len | flags << 30
    uint8_t ci[THREADS];                 // chunk (relative to the round's first)
};

template <int THREADS, typename Emit>
__device__ __forceinline__ void unit_for_each_kmer64(const ChunkView *__restrict__ chunks, uint32_t n_chunks, uint32_t unit,
                                                     uint32_t k, uint32_t forward_only, UnitStage<THREADS> *S, uint32_t *s_scan,
                                                     Emit emit) {
    constexpr uint32_t WARPS = THREADS / 32;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint64_t mask = (1ull << (2 * k)) - 1ull;
    const uint32_t le_mask = 0xFFFFFFFFu >> (31u - lane);
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += UNIT_MAXC) {
        const uint32_t nc = min((uint32_t)UNIT_MAXC, n_chunks - c0);
        if (tid < nc) {
            const ChunkView &cv = chunks[c0 + tid];
            uint32_t d0 = 0, d1 = 0;
            if (unit >= cv.first_unit && unit < cv.first_unit + cv.n_units) {
                d0 = cv.unit_off[unit - cv.first_unit]; d1 = cv.unit_off[unit - cv.first_unit + 1];
            }
            S->c_d0[tid] = d0; S->c_cnt[tid] = d1 - d0;
            S->c_pl[tid] = cv.payload - cv.word_bias;
        }
        __syncthreads();
        uint32_t dtot = 0;
        for (uint32_t c = 0; c < nc; c++) dtot += S->c_cnt[c];
        for (uint32_t g0 = 0; g0 < dtot; g0 += THREADS) {
            const uint32_t dr = min((uint32_t)THREADS, dtot - g0);
            uint32_t cnt = 0;
            if (tid < dr) {
                uint32_t g = g0 + tid, c = 0;
                while (g >= S->c_cnt[c]) { g -= S->c_cnt[c]; ++c; }
                const uint4 d = chunks[c0 + c].desc[S->c_d0[c] + g];
                S->woff[tid] = d.x;
                S->lf[tid] = d.y | (((d.z >> 16) & 3u) << 30);
                S->ci[tid] = (uint8_t)c;
                cnt = d.y - k + 1;
            }
            uint32_t tot;
            const uint32_t p = block_exclusive_scan<THREADS>(cnt, s_scan, &tot);
            if (tid < dr) S->start[tid] = p;
            if (tid == 0) S->start[dr] = tot;
            __syncthreads();
            const uint32_t per = ((tot + WARPS * 32u - 1u) / (WARPS * 32u)) * 32u;   // records per warp, 32-aligned
            const uint32_t r_beg = warp * per, r_end = min(tot, r_beg + per);
            if (r_beg < r_end) {
                uint32_t lo = 0, hi = dr - 1;                                          // last j with start[j] <= r_beg
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1) >> 1;
                    if (S->start[mid] <= r_beg) lo = mid; else hi = mid - 1;
                }
                uint32_t j = lo;                                                       // invariant: start[j] <= r0 < start[j+1]
                for (uint32_t r0 = r_beg; r0 < r_end; r0 += 32u) {
                    const uint32_t cand = j + 1u + lane;
                    const uint32_t rel = (cand < dr ? S->start[cand] : 0xFFFFFFFFu) - r0;  // > 0 by the invariant
                    const uint32_t smask = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
                    const uint32_t owner = j + (uint32_t)__popc(smask & le_mask);
                    const uint32_t r = r0 + lane;
                    if (r < r_end) {
                        const uint32_t lj = S->lf[owner];
                        const uint32_t i = r - S->start[owner], last = (lj & 0x3FFFFFFFu) - k, flags = lj >> 30;
                        const uint64_t fw = extract64(S->c_pl[S->ci[owner]] + S->woff[owner], 2ull * i) & mask;
                        const uint64_t rc = revcomp64(fw) >> (64 - 2 * k);
                        const bool isf = forward_only ? true : (fw < rc);
                        const uint64_t key = isf ? fw : rc;
                        const uint32_t bi = (!(flags & READ_FLAG_INCL_BEGIN) && i == 0) ? 1u : 0u;
                        const uint32_t ei = (!(flags & READ_FLAG_INCL_END) && i == last) ? 1u : 0u;
                        emit(key, (bi << (isf ? 0 : 1)) | (ei << (isf ? 1 : 0)));  // hashmap.rs:385-399
                    }
                    j += (uint32_t)__popc(smask);
                    if (j + 1u < dr && S->start[j + 1u] == r0 + 32u) ++j;              // next window starts a new super-k-mer
                }
            }
            __syncthreads();  // staging is rewritten by the next round
        }
    }
}

__host__ __device__ __forceinline__ uint32_t hash_table_slots(uint32_t n) {  // multiple of 512, >= 1.25 n + 64
    const uint64_t want = (uint64_t)n + n / 4 + 64;
    const uint64_t t = (want + 511) & ~511ull;
    return (uint32_t)(t < 1024 ? 1024 : t);
}

__host__ __device__ __forceinline__ uint32_t hash_table_slots_pow2(uint32_t n) {  // power of two >= 1.5 n (mask-indexed tables, merge128.cuh)
    uint32_t t = 1024;
    const uint64_t want = (uint64_t)n + n / 2;
    while (t < want) t <<= 1;
    return t;
}

constexpr uint32_t SORT_BINS = 512;  // bin-rank sort: bins on the top 9 key bits

// TS_STATIC > 0: table of up to TS_STATIC slots in shared memory (sized per unit: hash_table_slots(n)); survivors
// ordered by a bin-rank sort.  TS_STATIC == 0: table in this CTA's slice of a global scratch buffer (L2-resident
// for typical units); survivors ordered by the LSD radix sort.  Units whose survivors leave no room for the sort's
// second buffer are appended to `retry` and re-done by the sort-based kernel.
template <int THREADS, int TS_STATIC, int SRC = SRC_SUPERKMERS>
__global__ void __launch_bounds__(THREADS)
k_merge_hash(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ work, uint32_t n_work,
             uint32_t first_unit, DevParams P, uint32_t min_mult, MergeOut out, uint32_t *__restrict__ retry,
             uint32_t *__restrict__ retry_count, uint64_t *__restrict__ scratch, uint64_t per_cta_u64, PartSrc ps,
             const uint32_t *__restrict__ n_work_dev) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WARPS = THREADS / 32;
    constexpr uint32_t SCR_BYTES = WARPS * 256 * 4;  // scratch area: descriptor staging / survivor staging + bins / radix histograms
    static_assert(sizeof(UnitStage<THREADS>) <= SCR_BYTES, "descriptor staging must fit the scratch area");
    uint64_t *K = reinterpret_cast<uint64_t *>(smem_raw);                      // TS keys
    uint32_t *C = reinterpret_cast<uint32_t *>(K + TS_STATIC);                 // TS counters|flags
    uint32_t *hist = C + TS_STATIC;                                            // scratch area (SCR_BYTES)
    uint32_t *s_scan = hist + WARPS * 256;                                     // 40
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_scan + 40);
    UnitStage<THREADS> *stage = reinterpret_cast<UnitStage<THREADS> *>(hist);
    const uint32_t tid = threadIdx.x;
    if (n_work_dev) n_work = min(n_work, *n_work_dev);  // device-side list (big units whose partitions overflowed)
    for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        uint32_t *s_cnt = s_scan + 36;  // [0] survivors, [1] occupied slots, [2] records of the unit
        uint32_t unit = 0, oslot, n;
        if (SRC == SRC_RECORDS) {
            if (ps.big_ovf[ps.part_big[wi]]) continue;   // the whole unit is redone from its super-k-mers
            oslot = ps.part_slot[wi];
            n = min(ps.pcount[wi], ps.pcap);
            if (n == 0) continue;
        } else {
            unit = work[wi];
            oslot = out.slot(unit - first_unit);
            if (tid == 0) s_cnt[2] = 0;
            __syncthreads();
            for (uint32_t c = tid; c < n_chunks; c += THREADS) {  // one round of load latency whatever the chunk count
                const ChunkView &cv = chunks[c];
                if (unit >= cv.first_unit && unit < cv.first_unit + cv.n_units) {
                    const uint32_t v = cv.unit_kmers[unit - cv.first_unit];
                    if (v) atomicAdd(&s_cnt[2], v);
                }
            }
            __syncthreads();
            n = s_cnt[2];
        }
        uint32_t TS = hash_table_slots(n);
        if (TS_STATIC == 0) {
            K = scratch + (uint64_t)blockIdx.x * per_cta_u64;
            C = reinterpret_cast<uint32_t *>(K + TS);
        } else {
            if (TS > (uint32_t)TS_STATIC) {  // host routes such units elsewhere
                if (tid == 0) *out.overflow = 2u;
                continue;
            }
            TS = min(TS, (uint32_t)TS_STATIC);
        }
                for (uint32_t i = tid; i < TS; i += THREADS) { K[i] = HASH_EMPTY; C[i] = 0u; }
        __syncthreads();
        if (SRC == SRC_RECORDS) {
            const uint64_t *recs = ps.recs + (uint64_t)wi * ps.pcap;
            for (uint32_t i = tid; i < n; i += THREADS) {
                const uint64_t r = recs[i];
                hash_insert(K, C, TS, r >> 2, (uint32_t)r & 3u);
            }
        } else {
            unit_for_each_kmer64<THREADS>(chunks, n_chunks, unit, P.k, P.forward_only, stage, s_scan,
                                          [&](uint64_t key, uint32_t fb) { hash_insert(K, C, TS, key, fb); });
        }
        __syncthreads();
        // ---- scan the table once: MapEntry -> multiplicity, filter; survivors are appended to a small
        //      staging area (in the scratch area) in arbitrary order
        constexpr uint32_t STAGE_CAP = TS_STATIC ? (SCR_BYTES - SORT_BINS * 8) / 12 : SCR_BYTES / 12;
        constexpr uint32_t RANK_MAX = 512;                                  // rank-sort threshold (global-table variant)
        uint64_t *stage_k = reinterpret_cast<uint64_t *>(hist);
        uint32_t *stage_c = reinterpret_cast<uint32_t *>(stage_k + STAGE_CAP);
        uint32_t *bin_start = stage_c + STAGE_CAP;                          // [SORT_BINS]   (shared-table variant)
        uint32_t *bin_cur = bin_start + SORT_BINS;                          // [SORT_BINS]
        if (tid < 2) s_cnt[tid] = 0;
        __syncthreads();
        {
            uint32_t my_occ = 0;
            for (uint32_t i = tid; i < TS; i += THREADS) {
                const uint64_t kk = K[i];
                if (kk == HASH_EMPTY) continue;
                ++my_occ;
                const uint32_t cc = C[i];
                const uint32_t cnt = slot_count(cc), fl = cc >> 30;
                const uint32_t mult = cnt >> ((fl == (READ_FLAG_INCL_BEGIN | READ_FLAG_INCL_END)) ? 1 : 0);  // map_entry.rs:79-84
                if (mult >= min_mult) {
                    const uint32_t idx = atomicAdd(&s_cnt[0], 1u);
                    if (idx < STAGE_CAP) { stage_k[idx] = (kk << 2) | fl; stage_c[idx] = mult | (fl << 30); }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_occ += __shfl_xor_sync(0xffffffffu, my_occ, o);
            if (lane_id() == 0 && my_occ) atomicAdd(&s_cnt[1], my_occ);
        }
        __syncthreads();
        const uint32_t S = s_cnt[0], n_occ = s_cnt[1];
        const uint32_t end_bit = min(64u, (2 * P.k + 2 + 7) & ~7u);
        if (SRC != SRC_RECORDS && S > STAGE_CAP && S > TS / 2) {  // no room for the sort's second buffer: hand the unit to the sort-based kernel
            if (tid == 0) retry[atomicAdd(retry_count, 1u)] = unit;
            __syncthreads();
            continue;
        }
        if (tid == 0) {
            const unsigned long long b = atomicAdd(&out.cursor[0], (unsigned long long)S);
            atomicAdd(&out.cursor[1], (unsigned long long)n_occ);
            atomicAdd(&out.cursor[2], (unsigned long long)n);
            *s_base = b;
            out.unit_out_off[oslot] = b;
            out.unit_out_cnt[oslot] = S;
            if (b + S > out.capacity) *out.overflow = 1u;
        }
        __syncthreads();
        const unsigned long long gbase = *s_base;
        const bool room = gbase + S <= out.capacity;
        if (SRC == SRC_RECORDS) {
            // a key partition: survivors leave in any order, k_finish_units sorts the whole unit
            if (room) {
                if (S <= STAGE_CAP) {
                    for (uint32_t i = tid; i < S; i += THREADS) { out.keys[gbase + i] = stage_k[i] >> 2; out.count_flags[gbase + i] = stage_c[i]; }
                } else {
                    if (tid == 0) s_cnt[0] = 0;
                    __syncthreads();
                    for (uint32_t base = 0; base < TS; base += THREADS) {
                        const uint32_t i = base + tid;
                        const uint64_t kk = K[i];
                        uint32_t cf = 0;
                        if (kk != HASH_EMPTY) {
                            const uint32_t cc = C[i];
                            const uint32_t cnt = slot_count(cc), fl = cc >> 30;
                            const uint32_t mult = cnt >> ((fl == 3u) ? 1 : 0);
                            if (mult >= min_mult) cf = mult | (fl << 30);
                        }
                        const uint32_t bal = __ballot_sync(0xffffffffu, cf != 0);
                        uint32_t wb = 0;
                        if (lane_id() == 0 && bal) wb = atomicAdd(&s_cnt[0], (uint32_t)__popc(bal));
                        wb = __shfl_sync(0xffffffffu, wb, 0);
                        if (cf) {
                            const unsigned long long o = gbase + wb + __popc(bal & ((1u << lane_id()) - 1u));
                            out.keys[o] = kk; out.count_flags[o] = cf;
                        }
                    }
                }
            }
            __syncthreads();
            continue;
        }
        // many survivors: in-place block-scan compaction of the table itself -> K[0..S), C[0..S)
        auto compact_table = [&]() {
            uint32_t running = 0;
            for (uint32_t base = 0; base < TS; base += THREADS) {
                const uint32_t i = base + tid;
                const uint64_t kk = K[i];
                const uint32_t cc = C[i];
                uint32_t cf = 0, fl = 0;
                if (kk != HASH_EMPTY) {
                    const uint32_t cnt = slot_count(cc);
                    fl = cc >> 30;
                    const uint32_t mult = cnt >> ((fl == (READ_FLAG_INCL_BEGIN | READ_FLAG_INCL_END)) ? 1 : 0);
                    if (mult >= min_mult) cf = mult | (fl << 30);
                }
                uint32_t tot;
                const uint32_t p = block_exclusive_scan<THREADS>(cf ? 1u : 0u, s_scan, &tot);  // all reads precede writes
                if (cf) { K[running + p] = (kk << 2) | fl; C[running + p] = cf; }
                running += tot;
            }
            __syncthreads();
        };
        if (TS_STATIC != 0) {
            // ---- bin-rank sort: scatter the survivors into SORT_BINS bins on their top key bits (counting sort),
            //      then every survivor's position = bin start + number of smaller keys inside its bin.
            const uint64_t *src_k; const uint32_t *src_c;
            uint64_t *dst_k; uint32_t *dst_c;
            if (S <= STAGE_CAP) { src_k = stage_k; src_c = stage_c; dst_k = K; dst_c = C; }
            else { compact_table(); src_k = K; src_c = C; dst_k = K + TS / 2; dst_c = C + TS / 2; }
            const uint32_t bshift = 2 * P.k + 2 > 9 ? 2 * P.k + 2 - 9 : 0;  // staged keys carry 2 flag bits
            for (uint32_t i = tid; i < SORT_BINS; i += THREADS) bin_cur[i] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < S; i += THREADS) atomicAdd(&bin_cur[(uint32_t)(src_k[i] >> bshift) & (SORT_BINS - 1)], 1u);
            __syncthreads();
            {
                uint32_t v = 0;
                if (tid < SORT_BINS) v = bin_cur[tid];
                uint32_t tot;
                const uint32_t p = block_exclusive_scan<THREADS>(v, s_scan, &tot);
                if (tid < SORT_BINS) { bin_start[tid] = p; bin_cur[tid] = p; }
            }
            __syncthreads();
            for (uint32_t i = tid; i < S; i += THREADS) {
                const uint64_t r = src_k[i];
                const uint32_t pos = atomicAdd(&bin_cur[(uint32_t)(r >> bshift) & (SORT_BINS - 1)], 1u);
                dst_k[pos] = r; dst_c[pos] = src_c[i];
            }
            __syncthreads();
            if (room) {
                for (uint32_t i = tid; i < S; i += THREADS) {
                    const uint64_t r = dst_k[i];
                    const uint32_t b = (uint32_t)(r >> bshift) & (SORT_BINS - 1);
                    const uint32_t lo = bin_start[b], hi = bin_cur[b];   // bin_cur == end of the bin after the scatter
                    uint32_t rank = lo;
                    for (uint32_t j = lo; j < hi; j++) rank += dst_k[j] < r ? 1u : 0u;
                    out.keys[gbase + rank] = r >> 2;
                    out.count_flags[gbase + rank] = dst_c[i];
                }
            }
        } else if (S <= RANK_MAX) {
            // rank sort: keys are distinct, so rank = number of smaller keys is the sorted position
            if (room) {
                for (uint32_t i = tid; i < S; i += THREADS) {
                    const uint64_t r = stage_k[i];
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < S; j++) rank += stage_k[j] < r ? 1u : 0u;
                    out.keys[gbase + rank] = r >> 2;
                    out.count_flags[gbase + rank] = stage_c[i];
                }
            }
        } else {
            if (S <= STAGE_CAP) {  // staging -> table memory (the table is dead now), then key-value radix sort
                for (uint32_t i = tid; i < S; i += THREADS) { K[i] = stage_k[i]; C[i] = stage_c[i]; }
                __syncthreads();
            } else compact_table();
            uint32_t *Vs = nullptr;
            uint64_t *Ss = block_radix_sort64<THREADS, true>(K, K + TS / 2, S, 0, end_bit, hist, s_scan, C, C + TS / 2, &Vs);
            if (room) {
                for (uint32_t i = tid; i < S; i += THREADS) {
                    out.keys[gbase + i] = Ss[i] >> 2;
                    out.count_flags[gbase + i] = Vs[i];
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_partition_units: one CTA per big unit.  Expands the unit (load-balanced, as k_merge_hash) and appends every record
// to partition part_hash(key) & (P-1); the CTA owns all partitions of its unit, so the cursors live in shared memory.
constexpr int PART_MAXP = 4096;
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_partition_units(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ big_unit,
                  const uint32_t *__restrict__ big_logp, const uint32_t *__restrict__ big_pbase, uint32_t n_big, DevParams P,
                  uint64_t *__restrict__ recs, uint32_t *__restrict__ pcount, uint32_t pcap, uint32_t *__restrict__ big_ovf,
                  uint32_t *__restrict__ retry, uint32_t *__restrict__ retry_count) {
    __shared__ uint32_t s_cur[PART_MAXP];
    static_assert(THREADS == 1024, "launched with 1024 threads");
    __shared__ __align__(16) unsigned char s_stage_raw[sizeof(UnitStage<1024>)];
    __shared__ uint32_t s_scan[40];
    __shared__ uint32_t s_ovf;
    UnitStage<1024> *stage = reinterpret_cast<UnitStage<1024> *>(s_stage_raw);
    const uint32_t tid = threadIdx.x;
    for (uint32_t bi = blockIdx.x; bi < n_big; bi += gridDim.x) {
        const uint32_t unit = big_unit[bi], np = 1u << big_logp[bi], pbase = big_pbase[bi];
        for (uint32_t i = tid; i < np; i += THREADS) s_cur[i] = 0;
        if (tid == 0) s_ovf = 0;
        __syncthreads();
        uint64_t *dst = recs + (uint64_t)pbase * pcap;
        unit_for_each_kmer64<1024>(chunks, n_chunks, unit, P.k, P.forward_only, stage, s_scan, [&](uint64_t key, uint32_t fb) {
            const uint32_t p = part_hash(key) & (np - 1);
            const uint32_t pos = atomicAdd(&s_cur[p], 1u);
            if (pos < pcap) dst[(uint64_t)p * pcap + pos] = (key << 2) | fb;
            else s_ovf = 1u;
        });
        __syncthreads();
        for (uint32_t i = tid; i < np; i += THREADS) pcount[pbase + i] = min(s_cur[i], pcap);
        if (tid == 0 && s_ovf) { big_ovf[bi] = 1u; retry[atomicAdd(retry_count, 1u)] = unit; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_finish_units: survivors of every unit -> unit-ordered final table.  A unit with one output slot is already
// sorted (copy); a big unit's slots (one per key partition, unsorted) are concatenated and radix-sorted by key,
// in shared memory when they fit, else between the final buffer and a global scratch of the same layout.
template <int THREADS, int SCAP>
constexpr size_t finish_units_smem_bytes() { return (size_t)(THREADS / 32) * 256 * 4 + 48 * 4 + (size_t)SCAP * 24; }

template <int THREADS, int SCAP>
__global__ void __launch_bounds__(THREADS)
k_finish_units(const uint64_t *__restrict__ src_keys, const uint32_t *__restrict__ src_cf, const uint64_t *__restrict__ slot_off,
               const uint32_t *__restrict__ slot_cnt, const uint32_t *__restrict__ slot_of_unit /* n_units + 1, or NULL */,
               const uint64_t *__restrict__ dst_off, uint64_t *__restrict__ dst_keys, uint32_t *__restrict__ dst_cf,
               uint64_t *__restrict__ tmp_keys, uint32_t *__restrict__ tmp_cf, uint32_t n_units, uint32_t end_bit,
               uint64_t capacity, uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the final table grows with the survivors actually seen: when this part does not fit, nothing is written and the
    // host enlarges the table and launches the gather again (overflow bit 2)
    if (dst_off[n_units] > capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(overflow, 4u);
        return;
    }
    const uint64_t tmp_base = dst_off[0];   // tmp_* hold this part only
    constexpr int WARPS = THREADS / 32;
    uint64_t *sA = reinterpret_cast<uint64_t *>(smem_raw), *sB = sA + SCAP;
    uint32_t *sAv = reinterpret_cast<uint32_t *>(sB + SCAP), *sBv = sAv + SCAP;
    uint32_t *hist = sBv + SCAP;
    uint32_t *s_scan = hist + WARPS * 256;
    const uint32_t tid = threadIdx.x;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint32_t s0 = slot_of_unit ? slot_of_unit[u] : u, s1 = slot_of_unit ? slot_of_unit[u + 1] : u + 1;
        const uint64_t d = dst_off[u];
        const uint32_t n = (uint32_t)(dst_off[u + 1] - d);
        if (n == 0) continue;
        if (s1 - s0 == 1) {
            const uint64_t so = slot_off[s0];
            for (uint32_t i = tid; i < n; i += THREADS) { dst_keys[d + i] = src_keys[so + i]; dst_cf[d + i] = src_cf[so + i]; }
            continue;
        }
        const bool in_smem = n <= (uint32_t)SCAP;
        uint64_t *A = in_smem ? sA : dst_keys + d;
        uint32_t *Av = in_smem ? sAv : dst_cf + d;
        uint32_t run = 0;
        for (uint32_t sl = s0; sl < s1; sl++) {   // concatenate the partitions
            const uint32_t c = slot_cnt[sl];
            const uint64_t so = slot_off[sl];
            for (uint32_t i = tid; i < c; i += THREADS) { A[run + i] = src_keys[so + i]; Av[run + i] = src_cf[so + i]; }
            run += c;
        }
        __syncthreads();
        uint64_t *B = in_smem ? sB : tmp_keys + (d - tmp_base);
        uint32_t *Bv = in_smem ? sBv : tmp_cf + (d - tmp_base);
        uint32_t *Vs = nullptr;
        uint64_t *Ss = block_radix_sort64<THREADS, true>(A, B, n, 0, end_bit, hist, s_scan, Av, Bv, &Vs);
        if (Ss != dst_keys + d)
            for (uint32_t i = tid; i < n; i += THREADS) { dst_keys[d + i] = Ss[i]; dst_cf[d + i] = Vs[i]; }
        __syncthreads();
    }
}

// Exclusive scan of per-unit survivor totals (sum over the unit's output slots) -> u64 offsets, off[n] = total.
__global__ void __launch_bounds__(1024) k_scan_unit_slots(const uint32_t *__restrict__ slot_cnt, const uint32_t *__restrict__ slot_of_unit,
                                                           uint64_t *__restrict__ off, uint32_t n, uint64_t base) {
    __shared__ uint32_t s_scan[1024 / 32 + 2];
    uint64_t running = base;
    for (uint32_t b0 = 0; b0 < n; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        uint32_t v = 0;
        if (i < n) {
            if (slot_of_unit) { for (uint32_t s = slot_of_unit[i]; s < slot_of_unit[i + 1]; s++) v += slot_cnt[s]; }
            else v = slot_cnt[i];
        }
        uint32_t tot;
        const uint32_t p = block_exclusive_scan<1024>(v, s_scan, &tot);
        if (i < n) off[i] = running + p;
        running += tot;
    }
    if (threadIdx.x == 0) off[n] = running;
}

template <int THREADS, int TS_STATIC>
constexpr size_t merge_hash_smem_bytes() {
    return (size_t)TS_STATIC * 12 + (size_t)(THREADS / 32) * 256 * 4 + 40 * 4 + 16;
}

template <int THREADS, int CAP>
constexpr size_t merge_smem_bytes(bool global_scratch) {
    return (size_t)(THREADS / 32) * 256 * 4 + 40 * 4 + 16 + (global_scratch ? 0 : (size_t)CAP * 16);
}

}  // namespace ggb
