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
    const uint32_t *unit_woff;   // [n_units + 1] payload word offsets, relative to `payload` (a unit's words are contiguous)
    const uint32_t *unit_kmers;  // [n_units]
    uint32_t first_unit;         // units [first_unit, first_unit + n_units) are present
    uint32_t n_units;
    uint32_t word_bias;          // subtract from descriptor payload offsets (imported slices)
    uint32_t pad;
};

// Output of the merge kernels, before the unit-ordered final layout.  Every unit owns the region
// [static_off[u], static_off[u] + records(u)) of keys/count_flags (survivors <= distinct <= records), so no kernel
// waits for a global allocation: a CTA that owns a whole unit writes at the region's start, the key partitions of a
// big unit append to the region with unit_out_cnt as the fill counter.  cursor[] are statistics only (fire-and-forget
// atomics).
constexpr uint32_t SLOT_SORTED = 0x80000000u;   // flag in unit_out_cnt: the slot's entries are already ordered by key
struct MergeOut {
    uint64_t *keys;                     // survivors: canonical k-mer (2k bits)
    uint32_t *count_flags;              // multiplicity (30 bits, saturating) | flags << 30
    unsigned long long *cursor;         // [0] entries written, [1] distinct keys, [2] k-mer occurrences
    uint64_t *unit_out_off;             // per output slot: offset
    uint32_t *unit_out_cnt;             //                  count | SLOT_SORTED
    const uint64_t *static_off;         // per unit (relative to the first unit of the launch): start of its region
    uint32_t *overflow;                 // set to 2 if a unit was routed to a kernel that cannot hold it
    const uint32_t *slot_of_unit;       // output slot of a unit (big units own several slots, one per key partition);
                                        // NULL: slot = unit index relative to the first unit of the launch
    __device__ __forceinline__ uint32_t slot(uint32_t unit_rel) const { return slot_of_unit ? slot_of_unit[unit_rel] : unit_rel; }
    __device__ __forceinline__ void stats(uint32_t survivors, uint32_t distinct, uint32_t records) const {
        atomicAdd(&cursor[0], (unsigned long long)survivors);
        atomicAdd(&cursor[1], (unsigned long long)distinct);
        atomicAdd(&cursor[2], (unsigned long long)records);
    }
};

// Key partitions of big units (units that do not fit a shared-memory table): k_partition_units expands the unit once
// and routes every k-mer record by a hash of its key into one of P partitions in HBM (all occurrences of a k-mer
// land in the same partition, so partitions are counted independently, no fold); k_merge_parts
// counts one partition per CTA in shared memory; k_finish_units orders the unit's survivors.
struct PartSrc {
    const uint64_t *recs;        // [n_parts_total][pcap] records (key << 2 | flag bits)
    const uint32_t *pcount;      // [n_parts_total] records per partition
    const uint32_t *part_big;    // work item -> index of its big unit
    const uint32_t *big_unit;    // [n_big] unit id
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
                                    uint32_t unit_rel, uint32_t *s_scan) {
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
    const unsigned long long gbase = out.static_off[unit_rel];
    if (tid == 0) {
        out.stats(tot, tot_heads, n);
        const uint32_t sl = out.slot(unit_rel);
        out.unit_out_off[sl] = gbase;
        out.unit_out_cnt[sl] = tot | SLOT_SORTED;
    }
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
    uint64_t *sA = reinterpret_cast<uint64_t *>(s_scan + 44);
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
        block_reduce_filter<THREADS>(S, reinterpret_cast<uint32_t *>(other), n, min_mult, out, unit - first_unit, s_scan);
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
    if (!claimed) {
        // The counter SATURATES at 2^30 - 1 occurrences (the reference's counter is 61-bit, map_entry.rs:5-10; -s only
        // needs "at least"): far below the limit a plain atomicAdd (at most ~10^5 threads race, the margin is 2^20);
        // close to it a CAS loop that never carries into the flag bits.  Only giant units get here.
        uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&C[slot]);
        if ((cur & 0x3FFFFFFFu) < 0x3FF00000u) atomicAdd(&C[slot], 1u);
        else
            while ((cur & 0x3FFFFFFFu) < 0x3FFFFFFEu) {
                const uint32_t old = atomicCAS(&C[slot], cur, cur + 1u);
                if (old == cur) break;
                cur = old;
            }
    }
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

template <int THREADS>
__host__ __device__ constexpr size_t merge_hash_stage_bytes() { return (sizeof(UnitStage<THREADS>) + 15) & ~(size_t)15; }

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


// TS_STATIC > 0: table of up to TS_STATIC slots in shared memory (sized per unit: hash_table_slots(n)).
// TS_STATIC == 0: table in this CTA's slice of a global scratch buffer (L2-resident for typical units).
// After the inserts the table is scanned once (MapEntry -> multiplicity, -s filter) and the survivors are written
// straight into the unit's output region in table order; ordering by key (part of the output contract) is the job of
// k_finish_small / k_finish_units, which touch only the survivors.  Per unit the CTA passes 7 barriers.
// (Key partitions of big units are counted by k_merge_parts below.)
template <int THREADS, int TS_STATIC>
__global__ void __launch_bounds__(THREADS)
k_merge_hash(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ work, uint32_t n_work,
             uint32_t first_unit, DevParams P, uint32_t min_mult, MergeOut out, const uint32_t *__restrict__ unit_n,
             uint64_t *__restrict__ scratch, uint64_t per_cta_u64, const uint32_t *__restrict__ n_work_dev) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *K = reinterpret_cast<uint64_t *>(smem_raw);                      // TS keys
    uint32_t *C = reinterpret_cast<uint32_t *>(K + TS_STATIC);                 // TS counters|flags
    UnitStage<THREADS> *stage = reinterpret_cast<UnitStage<THREADS> *>(C + TS_STATIC);   // descriptor staging
    uint32_t *s_scan = reinterpret_cast<uint32_t *>(smem_raw + (size_t)TS_STATIC * 12 + merge_hash_stage_bytes<THREADS>());  // 40
    const uint32_t tid = threadIdx.x;
    if (n_work_dev) n_work = min(n_work, *n_work_dev);  // device-side list (big units whose partitions overflowed)
    for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        uint32_t *s_cnt = s_scan + 36;  // [0] survivors written, [1] occupied slots
        const uint32_t unit_rel = work[wi] - first_unit;
        const uint32_t oslot = out.slot(unit_rel);
        const uint32_t n = unit_n[unit_rel];
        uint32_t TS = hash_table_slots(n);
        if (TS_STATIC == 0) {
            K = scratch + (uint64_t)blockIdx.x * per_cta_u64;
            C = reinterpret_cast<uint32_t *>(K + TS);
        } else if (TS > (uint32_t)TS_STATIC) {  // host routes such units elsewhere
            if (tid == 0) *out.overflow = 2u;
            continue;
        }
        for (uint32_t i = tid; i < TS; i += THREADS) { K[i] = HASH_EMPTY; C[i] = 0u; }
        if (tid < 2) s_cnt[tid] = 0;
        __syncthreads();
        unit_for_each_kmer64<THREADS>(chunks, n_chunks, work[wi], P.k, P.forward_only, stage, s_scan,
                                      [&](uint64_t key, uint32_t fb) { hash_insert(K, C, TS, key, fb); });
        __syncthreads();
        // ---- scan the table once: MapEntry -> multiplicity, filter, survivors appended to the unit's region
        const unsigned long long gbase = out.static_off[unit_rel];
        {
            uint32_t my_occ = 0;
            for (uint32_t base = 0; base < TS; base += THREADS) {
                const uint32_t i = base + tid;
                uint64_t kk = HASH_EMPTY;
                uint32_t cf = 0;
                if (i < TS) {
                    kk = K[i];
                    if (kk != HASH_EMPTY) {
                        ++my_occ;
                        const uint32_t cc = C[i];
                        const uint32_t cnt = slot_count(cc), fl = cc >> 30;
                        const uint32_t mult = cnt >> ((fl == (READ_FLAG_INCL_BEGIN | READ_FLAG_INCL_END)) ? 1 : 0);  // map_entry.rs:79-84
                        if (mult >= min_mult) cf = (mult > 0x3FFFFFFFu ? 0x3FFFFFFFu : mult) | (fl << 30);
                    }
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, cf != 0);
                if (bal) {
                    uint32_t wb = 0;
                    if (lane_id() == 0) wb = atomicAdd(&s_cnt[0], (uint32_t)__popc(bal));
                    wb = __shfl_sync(0xffffffffu, wb, 0);
                    if (cf) {
                        const unsigned long long o = gbase + wb + __popc(bal & ((1u << lane_id()) - 1u));
                        out.keys[o] = kk; out.count_flags[o] = cf;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_occ += __shfl_xor_sync(0xffffffffu, my_occ, o);
            if (lane_id() == 0 && my_occ) atomicAdd(&s_cnt[1], my_occ);
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t S = s_cnt[0];
            out.stats(S, s_cnt[1], n);
            out.unit_out_off[oslot] = gbase;
            out.unit_out_cnt[oslot] = S;
        }
        __syncthreads();   // s_cnt / table are rewritten by the next unit
    }
}

// ------------------------------------------------------------------------------------------------
// hash_insert_bounded: hash_insert that gives up after `limit` probes (table sized from an ESTIMATE of the distinct
// keys: the caller re-routes the unit / splits the key space when the estimate was too low).
__device__ __forceinline__ bool hash_insert_bounded(uint64_t *K, uint32_t *C, uint32_t TS, uint64_t key, uint32_t fb, uint32_t limit) {
    uint32_t slot = hash_slot(key, TS);
    bool claimed = false;
    for (uint32_t probes = 0;; ++probes) {
        if (probes >= limit) return false;
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
    return true;
}

// ------------------------------------------------------------------------------------------------
// k_merge_tier: the shared-table merge of every unit that fits one staging round -- the kernel that counts almost all
// k-mers of a build (C2: all units; human-scale inputs: whatever is not a "big" unit).
//
//   * PERSISTENT CTAs pull units from a device-side work counter (dynamic balance, no tail of idle CTAs).
//   * A unit-sorted chunk keeps a unit's descriptors AND its payload words contiguous, so staging a unit is two 1-D bulk
//     copies per chunk slice (cp.async.bulk -> UBLKCP, completion on an mbarrier).  Warp 0 arms the copies of the NEXT
//     unit before the CTA touches the current one (NBUF = 2: second landing buffer; NBUF = 1: issued when the inserts of
//     the current unit are done, so they overlap the table scan), which hides the descriptor / payload latency that the
//     first version of this kernel paid at the start of every unit (ncu: 12.6 % of its stall samples on that load).
//   * All bit extraction reads shared memory: descriptors are rewritten in place to {first record, payload word, len|flags}.
//   * A unit is routed to a tier by the DISTINCT keys it is expected to hold (host: distinct/records ratio of what this
//     context merged before, target load <= 0.5), not by its records: at 30x coverage a unit of 16 k records holds ~5 k
//     keys.  Divergence in the probe loop is what costs issue slots (every extra probe of one lane stalls 31 others), so
//     the table is kept sparse.  Inserts probe a bounded number of slots; a unit whose table fills goes to the retry list
//     (global-table kernel).  Scanning the table resets it for the next unit (no separate clear pass).
//   * Barriers per unit: 1 (prefix scan of the k-mer counts) + 3, against 7 + 3 per staging round before.
// Semantics per record: hashmap.rs:385-399 / map_entry.rs:33-84, as k_merge_hash.
constexpr int TIER_MAXSL = 32;             // chunk slices of one unit (one lane of warp 0 each)
constexpr uint32_t TIER_PROBE_LIMIT = 128;
constexpr uint32_t TIER_DONE = 0xFFFFFFFFu;

template <int THREADS, int TS, int SKCAP, int PWCAP, int NBUF>
struct TierSmem {
    static constexpr size_t k_off = 0;
    static constexpr size_t c_off = k_off + (size_t)TS * 8;
    static constexpr size_t sk_off = c_off + (size_t)TS * 4;                              // NBUF x SKCAP uint4 (TMA landing, then staged records)
    static constexpr size_t pay_off = sk_off + (size_t)NBUF * SKCAP * 16;                 // NBUF x (PWCAP + 8) u32 (TMA landing)
    static constexpr size_t start_off = pay_off + (size_t)NBUF * (PWCAP + 8) * 4;         // NBUF x SKCAP u32: first record of a super-k-mer
    static constexpr size_t bar_off = start_off + (size_t)NBUF * SKCAP * 4;               // NBUF mbarriers
    static constexpr size_t meta_off = bar_off + (size_t)NBUF * 8;                        // NBUF x {unit, n_sk, region start lo, hi}
    static constexpr size_t dstart_off = meta_off + (size_t)NBUF * 32;                    // NBUF x (MAXSL + 1): first staged descriptor of a slice
    static constexpr size_t delta_off = dstart_off + (size_t)NBUF * (TIER_MAXSL + 1) * 4; // NBUF x MAXSL: staged payload word = descriptor word + delta
    static constexpr size_t dtab_off = delta_off + (size_t)NBUF * TIER_MAXSL * 4;         // 2 x SKCAP u32: super-k-mer dedup table
    static constexpr size_t scan_off = dtab_off + (size_t)2 * SKCAP * 4;                  // 34 u32
    static constexpr size_t cnt_off = scan_off + 34 * 4;                                  // 4 u32
    static constexpr size_t bytes = ((cnt_off + 4 * 4 + 15) / 16) * 16;
    static_assert((size_t)TS % 4 == 0 && PWCAP % 4 == 0 && SKCAP % 4 == 0, "bulk-copy destinations must be 16-byte aligned");
};

// Exclusive block scan with ONE barrier: every warp re-scans the WARPS partial sums itself.
template <int THREADS>
__device__ __forceinline__ uint32_t block_scan1(uint32_t v, uint32_t *s_w, uint32_t *total) {
    constexpr int WARPS = THREADS / 32;
    const uint32_t lane = lane_id(), warp = warp_id();
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    const uint32_t w = lane < (uint32_t)WARPS ? s_w[lane] : 0u;
    uint32_t s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (uint32_t)o) s += y; }
    const uint32_t base = __shfl_sync(0xffffffffu, s - w, warp);
    *total = __shfl_sync(0xffffffffu, s, WARPS - 1);
    return base + x - v;
}

// Warp 0: take the next unit from the work counter and arm the bulk copies of its slices into one landing buffer.
// meta = {unit, n_sk, start of the unit's output region (lo, hi)}: loaded here, a whole unit ahead of their use.
__device__ __forceinline__ void tier_prefetch(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ work,
                                              uint32_t n_work, uint32_t *__restrict__ work_counter, uint32_t first_unit,
                                              const uint64_t *__restrict__ static_off, const uint32_t *__restrict__ unit_n, uint4 *sk, uint32_t *pay,
                                              uint64_t *bar, uint32_t *meta, uint32_t *sl_dstart, uint32_t *sl_delta) {
    const uint32_t lane = lane_id();
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(work_counter, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= n_work) { if (lane == 0) meta[0] = TIER_DONE; return; }
    const uint32_t unit = work[wi];
    uint32_t cnt = 0, nw = 0, lead = 0, w0 = 0, d0 = 0, bias = 0;
    const uint4 *desc = nullptr;
    const uint32_t *src = nullptr;
    if (lane < n_chunks) {
        const ChunkView &cv = chunks[lane];
        if (unit >= cv.first_unit && unit < cv.first_unit + cv.n_units) {
            const uint32_t u = unit - cv.first_unit;
            d0 = cv.unit_off[u]; cnt = cv.unit_off[u + 1] - d0;
            w0 = cv.unit_woff[u]; nw = cv.unit_woff[u + 1] - w0;
            src = cv.payload + w0;
            lead = (uint32_t)(((uintptr_t)src >> 2) & 3u);      // the bulk copy starts at the 16-byte boundary below
            bias = cv.word_bias; desc = cv.desc;
        }
    }
    unsigned long long gbase = 0;
    uint32_t n_rec = 0;
    if (lane == 0) { gbase = static_off[unit - first_unit]; n_rec = unit_n[unit - first_unit]; }
    const uint32_t cw = cnt ? ((lead + nw + 3u) & ~3u) : 0u;
    uint32_t dx = cnt, px = cw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, dx, o), b = __shfl_up_sync(0xffffffffu, px, o);
        if (lane >= (uint32_t)o) { dx += a; px += b; }
    }
    const uint32_t dtot = __shfl_sync(0xffffffffu, dx, 31), ptot = __shfl_sync(0xffffffffu, px, 31);
    const uint32_t dpos = dx - cnt, ppos = px - cw;
    sl_dstart[lane] = dpos;
    sl_delta[lane] = ppos + lead - w0 - bias;   // descriptor word (relative to its source chunk) -> staged word
    if (lane == 31) sl_dstart[32] = dtot;
    fence_proxy_async_smem();                   // the landing buffers were read / rewritten by plain loads and stores
    if (lane == 0) {
        meta[0] = unit; meta[1] = dtot; meta[2] = (uint32_t)gbase; meta[3] = (uint32_t)(gbase >> 32); meta[4] = n_rec;
        mbar_expect_tx(bar, dtot * 16u + ptot * 4u);
    }
    __syncwarp();
    if (cnt) {
        tma_load_1d(sk + dpos, desc + d0, cnt * 16u, bar);
        tma_load_1d(pay + ppos, src - lead, cw * 4u, bar);
    }
}

template <int THREADS, int TS, int SKCAP, int PWCAP, int NBUF, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_merge_tier(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ work, uint32_t n_work,
             uint32_t first_unit, DevParams P, uint32_t min_mult, MergeOut out, const uint32_t *__restrict__ unit_n,
             uint32_t *__restrict__ work_counter, uint32_t *__restrict__ retry, uint32_t *__restrict__ retry_count) {
    using L = TierSmem<THREADS, TS, SKCAP, PWCAP, NBUF>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t WARPS = THREADS / 32;
    constexpr int ITEMS = (SKCAP + THREADS - 1) / THREADS;
    uint64_t *K = reinterpret_cast<uint64_t *>(smem_raw + L::k_off);
    uint32_t *C = reinterpret_cast<uint32_t *>(smem_raw + L::c_off);
    uint4 *skb = reinterpret_cast<uint4 *>(smem_raw + L::sk_off);
    uint32_t *payb = reinterpret_cast<uint32_t *>(smem_raw + L::pay_off);
    uint32_t *startb = reinterpret_cast<uint32_t *>(smem_raw + L::start_off);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + L::bar_off);
    uint32_t *metab = reinterpret_cast<uint32_t *>(smem_raw + L::meta_off);
    uint32_t *dstartb = reinterpret_cast<uint32_t *>(smem_raw + L::dstart_off);
    uint32_t *deltab = reinterpret_cast<uint32_t *>(smem_raw + L::delta_off);
    uint32_t *dtab = reinterpret_cast<uint32_t *>(smem_raw + L::dtab_off);
    uint32_t *s_scan = reinterpret_cast<uint32_t *>(smem_raw + L::scan_off);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem_raw + L::cnt_off);   // [0] survivors written, [1] occupied slots, [3] table full

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint32_t k = P.k, forward_only = P.forward_only;
    const uint32_t kmask_lo = (uint32_t)((1ull << (2 * k)) - 1ull), kmask_hi = (uint32_t)(((1ull << (2 * k)) - 1ull) >> 32), rc_shift = 64u - 2u * k;
    const uint32_t le_mask = 0xFFFFFFFFu >> (31u - lane);
    const uint32_t aK = smem_addr(K), aC = smem_addr(C), aFull = smem_addr(s_cnt + 3);
    auto prefetch = [&](uint32_t b) {
        tier_prefetch(chunks, n_chunks, work, n_work, work_counter, first_unit, out.static_off, unit_n, skb + (size_t)b * SKCAP,
                      payb + (size_t)b * (PWCAP + 8), bars + b, metab + 8 * b, dstartb + (TIER_MAXSL + 1) * b, deltab + TIER_MAXSL * b);
    };
    for (uint32_t i = tid; i < (uint32_t)TS; i += THREADS) { K[i] = HASH_EMPTY; C[i] = 0u; }
    for (uint32_t i = tid; i < 2u * SKCAP; i += THREADS) dtab[i] = 0xFFFFFFFFu;
    if (tid == 0) {
        for (int b = 0; b < NBUF; b++) mbar_init(bars + b, 1);
        s_cnt[0] = s_cnt[1] = s_cnt[2] = s_cnt[3] = 0;
    }
    __syncthreads();
    if (warp == 0) prefetch(0);
    __syncthreads();
    for (uint32_t it = 0;; ++it) {
        const uint32_t b = NBUF == 2 ? (it & 1u) : 0u;
        const uint32_t unit = metab[8 * b], nsk = metab[8 * b + 1], n_records = metab[8 * b + 4];
        if (unit == TIER_DONE) break;
        const unsigned long long gbase = ((unsigned long long)metab[8 * b + 3] << 32) | metab[8 * b + 2];
        if (NBUF == 2 && warp == 0) prefetch(b ^ 1u);     // the other landing buffer: its unit finished before the last barrier
        mbar_wait(bars + b, NBUF == 2 ? ((it >> 1) & 1u) : (it & 1u));
        const uint32_t unit_rel = unit - first_unit;
        uint4 *sk = skb + (size_t)b * SKCAP;
        uint32_t *start = startb + (size_t)b * SKCAP;
        const uint32_t *sl_dstart = dstartb + (TIER_MAXSL + 1) * b, *sl_delta = deltab + TIER_MAXSL * b;
        // ---- stage 1: raw descriptor {word, len, meta, colour} -> {-, staged payload word, len | flags << 30, multiplicity 1}
        const uint32_t *payw = payb + (size_t)b * (PWCAP + 8);
        uint32_t pwv[ITEMS], lfv[ITEMS];
#pragma unroll
        for (int t = 0; t < ITEMS; t++) {
            const uint32_t j = tid * ITEMS + t;
            pwv[t] = lfv[t] = 0;
            if (j < nsk) {
                const uint4 d = sk[j];
                uint32_t q = 0;
                while (j >= sl_dstart[q + 1]) ++q;
                pwv[t] = d.x + sl_delta[q];
                lfv[t] = d.y | (((d.z >> 16) & 3u) << 30);
                sk[j] = make_uint4(0u, pwv[t], lfv[t], 1u);
            }
        }
        __syncthreads();
        // ---- stage 2: super-k-mer compaction (the reference's BucketsCompactor, crates/minimizer_bucketing/src/compactor.rs:
        //      128-196: identical (length, flags, packed bases) super-k-mers of a sub-bucket merge, multiplicities add up).
        //      At 30x coverage about half of a unit's super-k-mers repeat one seen before, so half of the k-mer inserts
        //      become "+ multiplicity" of one insert.  Open-addressing table of super-k-mer indices, CAS-claimed; a
        //      duplicate adds 1 to its representative's multiplicity and contributes no records.
        uint32_t tot, nrep;
        {
            constexpr uint32_t DCAP = 2u * SKCAP;
            uint32_t myslot[ITEMS], packed[ITEMS], sum = 0;
#pragma unroll
            for (int t = 0; t < ITEMS; t++) {
                const uint32_t j = tid * ITEMS + t;
                myslot[t] = 0xFFFFFFFFu; packed[t] = 0;
                if (j < nsk) {
                    const uint32_t len = lfv[t] & 0x3FFFFFFFu, nw = (len + 15u) >> 4;
                    uint32_t h = lfv[t] * 0x9E3779B1u;
                    for (uint32_t w = 0; w < nw; w++) h = (h ^ payw[pwv[t] + w]) * 0x85EBCA6Bu + 0x27D4EB2Fu;
                    h ^= h >> 15;
                    uint32_t slot = __umulhi(h * 0x2C1B3C6Du, DCAP);
                    while (true) {
                        const uint32_t old = atomicCAS(&dtab[slot], 0xFFFFFFFFu, j);
                        if (old == 0xFFFFFFFFu) { myslot[t] = slot; packed[t] = (len - k + 1u) | (1u << 20); break; }
                        const uint4 o = sk[old];
                        bool same = o.z == lfv[t];
                        for (uint32_t w = 0; same && w < nw; w++) same = payw[o.y + w] == payw[pwv[t] + w];
                        if (same) { atomicAdd(&sk[old].w, 1u); break; }
                        slot = slot + 1u == DCAP ? 0u : slot + 1u;
                    }
                }
                sum += packed[t];
            }
            uint32_t total;
            uint32_t pre = block_scan1<THREADS>(sum, s_scan, &total);     // records in the low 20 bits, representatives above
            tot = total & 0xFFFFFu; nrep = total >> 20;
            uint32_t mult[ITEMS];
#pragma unroll
            for (int t = 0; t < ITEMS; t++) mult[t] = packed[t] ? sk[tid * ITEMS + t].w : 0u;    // final: every duplicate has added itself
            __syncthreads();
#pragma unroll
            for (int t = 0; t < ITEMS; t++) {
                if (packed[t]) {
                    const uint32_t rank = pre >> 20, rs = pre & 0xFFFFFu;
                    sk[rank] = make_uint4(rs, pwv[t], lfv[t], mult[t]);          // compacted in place (rank <= own index)
                    start[rank] = rs;
                    dtab[myslot[t]] = 0xFFFFFFFFu;
                }
                pre += packed[t];
            }
        }
        __syncthreads();
        // ---- insert: every warp walks an equal, 32-aligned range of the unit's records, one record per lane.
        //      Shared memory is addressed by 32-bit shared-window addresses; the probe loop only FINDS (or claims) the
        //      slot, the lanes reconverge, then one counter update / flag update for the whole warp.
        {
            const uint32_t aSk = smem_addr(sk), aStart = smem_addr(start), aPay = smem_addr(payb + (size_t)b * (PWCAP + 8));
            const uint32_t per = ((tot + WARPS * 32u - 1u) / (WARPS * 32u)) * 32u;
            const uint32_t r_beg = warp * per, r_end = min(tot, r_beg + per);
            if (r_beg < r_end) {
                uint32_t lo = 0, hi = nrep - 1;                                         // last j with start[j] <= r_beg
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1) >> 1;
                    if (lds_u32(aStart + 4u * mid) <= r_beg) lo = mid; else hi = mid - 1;
                }
                uint32_t j = lo;                                                       // invariant: start[j] <= r0 < start[j+1]
                for (uint32_t r0 = r_beg; r0 < r_end; r0 += 32u) {
                    if (lds_u32(aFull)) break;                                         // somebody found the table full
                    const uint32_t cand = j + 1u + lane;
                    const uint32_t rel = (cand < nrep ? lds_u32(aStart + 4u * cand) : 0xFFFFFFFFu) - r0;  // > 0 by the invariant
                    const uint32_t smask = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
                    const uint32_t owner = j + (uint32_t)__popc(smask & le_mask);
                    const uint32_t r = r0 + lane;
                    const bool active = r < r_end;
                    uint32_t slot = 0, fb = 0, add = 0;                                // add: what this lane adds to the slot's counter
                    bool claimed = false;                                              // (inactive lanes add nothing)
                    if (active) {
                        const uint4 s = lds_u128(aSk + 16u * owner);
                        const uint32_t i = r - s.x, last = (s.z & 0x3FFFFFFFu) - k, flags = s.z >> 30;
                        const uint32_t wa = aPay + 4u * (s.y + (i >> 4)), sh = (i & 15u) * 2u;
                        const uint32_t w0 = lds_u32(wa), w1 = lds_u32(wa + 4u), w2 = lds_u32(wa + 8u);
                        const uint32_t flo = __funnelshift_r(w0, w1, sh) & kmask_lo, fhi = __funnelshift_r(w1, w2, sh) & kmask_hi;
                        // reverse complement of the 2k-bit k-mer (cn_seqhash_base.rs:27-69 evaluated directly)
                        const unsigned long long fw = ((unsigned long long)fhi << 32) | flo;
                        const unsigned long long rcx = (((unsigned long long)revcomp32(flo) << 32) | revcomp32(fhi)) >> rc_shift;
                        const bool isf = forward_only ? true : (fw < rcx);
                        const unsigned long long key = isf ? fw : rcx;
                        const uint32_t bi = (!(flags & READ_FLAG_INCL_BEGIN) && i == 0) ? 1u : 0u;
                        const uint32_t ei = (!(flags & READ_FLAG_INCL_END) && i == last) ? 1u : 0u;
                        fb = (bi << (isf ? 0 : 1)) | (ei << (isf ? 1 : 0));           // hashmap.rs:385-399
                        const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
                        slot = hash_slot(key, (uint32_t)TS);
                        claimed = false;
                        add = s.w;                                                     // the super-k-mer's multiplicity
                        uint32_t probes = 0;
                        while (true) {
                            const uint2 cur = lds_u64x(aK + 8u * slot);
                            if (cur.x == klo && cur.y == khi) break;
                            if (cur.y == 0xFFFFFFFFu) {                                // empty: keys are < 2^62
                                const unsigned long long old = atoms_cas64(aK + 8u * slot, HASH_EMPTY, key);
                                if (old == HASH_EMPTY) { claimed = true; break; }
                                if (old == key) break;
                            }
                            slot = slot + 1u == (uint32_t)TS ? 0u : slot + 1u;
                            if (++probes >= TIER_PROBE_LIMIT) { s_cnt[3] = 1u; fb = 0; add = 0; break; }
                        }
                    }
                    __syncwarp();
                    if (claimed) --add;                    // the counter word holds occurrences - 1: the claim itself counts one
                    if (add) atoms_add32(aC + 4u * slot, add);
                    if (fb && ((lds_u32(aC + 4u * slot) >> 30) & fb) != fb) atoms_or32(aC + 4u * slot, fb << 30);
                    j += (uint32_t)__popc(smask);
                    if (j + 1u < nrep && lds_u32(aStart + 4u * (j + 1u)) == r0 + 32u) ++j;  // next window starts a new super-k-mer
                }
            }
        }
        __syncthreads();
        const uint32_t full = s_cnt[3];
        if (NBUF == 1 && warp == 0) prefetch(0);          // the landing buffer is free: the copies overlap the table scan
        if (full) {
            for (uint32_t i = tid; i < (uint32_t)TS; i += THREADS) { K[i] = HASH_EMPTY; C[i] = 0u; }
            if (tid == 0) retry[atomicAdd(retry_count, 1u)] = unit;
        } else {
            // ---- table -> survivors, two passes over the thread's own slots (no block barrier in between):
            //   1  MapEntry -> multiplicity (map_entry.rs:79-84), -s filter; the slot's counter word is overwritten with
            //      the entry's final count|flags (0 = dropped)
            //   2  one warp scan + ONE shared atomic per warp reserve the output range; survivors are written to the
            //      unit's region, every slot is reset for the next unit
            uint32_t my_occ = 0, my_keep = 0;
            for (uint32_t i = tid; i < (uint32_t)TS; i += THREADS) {
                const uint2 kk = lds_u64x(aK + 8u * i);
                uint32_t cf = 0;
                if (kk.y != 0xFFFFFFFFu) {
                    const uint32_t cc = C[i];
                    ++my_occ;
                    const uint32_t cnt = slot_count(cc), fl = cc >> 30;
                    const uint32_t mult = cnt >> ((fl == (READ_FLAG_INCL_BEGIN | READ_FLAG_INCL_END)) ? 1 : 0);
                    if (mult >= min_mult) { cf = mult | (fl << 30); ++my_keep; }
                    C[i] = cf;
                }
            }
            uint32_t incl = my_keep;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_occ += __shfl_xor_sync(0xffffffffu, my_occ, o);
            uint32_t wbase = 0;
            if (lane == 31) {
                if (incl) wbase = atomicAdd(&s_cnt[0], incl);
                if (my_occ) atomicAdd(&s_cnt[1], my_occ);
            }
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            unsigned long long o = gbase + wbase + (incl - my_keep);
            for (uint32_t i = tid; i < (uint32_t)TS; i += THREADS) {
                const uint32_t cf = C[i];
                if (cf) {
                    out.keys[o] = K[i]; out.count_flags[o] = cf;
                    ++o;
                    C[i] = 0u;
                }
                K[i] = HASH_EMPTY;
            }
        }
        __syncthreads();
        if (tid == 0) {
            if (!full) {
                const uint32_t S = s_cnt[0], oslot = out.slot(unit_rel);
                out.stats(S, s_cnt[1], n_records);
                out.unit_out_off[oslot] = gbase;
                out.unit_out_cnt[oslot] = S;
            }
            s_cnt[0] = s_cnt[1] = s_cnt[3] = 0;   // next written after the next unit's staging barrier
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_merge_parts: one CTA per key partition of a big unit (records written by k_partition_units).
// Partitions are sized by the DISTINCT keys they are expected to hold, not by their records (at 30x coverage a
// partition of 60 k records has ~2.5 k distinct k-mers): fewer, larger partitions mean a smaller scatter fan-out in
// k_partition_units and fewer table clears / scans per record here.  The expectation comes from the parts already
// merged, so it can be wrong: inserts probe a bounded number of slots, and when the table turns out to be full the CTA
// splits the partition's KEY SPACE four ways by an independent hash and counts each quarter on its own (work stack in
// shared memory, records re-read from HBM/L2).  Survivors of every successful pass are appended to the unit's output
// region (unit_out_cnt is the fill counter), in table order; k_finish_* orders the unit.
__device__ __forceinline__ uint32_t sub_hash(uint64_t key) {   // independent of part_hash and hash_slot
    return (uint32_t)((key * 0xA24BAED4963EE407ull) >> 33);
}

template <int THREADS, int TS_STATIC>
constexpr size_t merge_parts_smem_bytes() { return (size_t)TS_STATIC * 12; }

template <int THREADS, int TS_STATIC>
__global__ void __launch_bounds__(THREADS)
k_merge_parts(uint32_t n_parts, uint32_t first_unit, uint32_t min_mult, MergeOut out, PartSrc ps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *K = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *C = reinterpret_cast<uint32_t *>(K + TS_STATIC);
    __shared__ uint32_t s_cnt[6];          // [0] survivors, [1] occupied, [2] records of the pass, [3] table full, [4] write cursor, [5] region base (low)
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_sj[64], s_sq[64], s_sp;
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t PROBE_LIMIT = 192;
    for (uint32_t wi = blockIdx.x; wi < n_parts; wi += gridDim.x) {
        if (ps.big_ovf[ps.part_big[wi]]) continue;   // a partition overflowed its record buffer: the unit is redone elsewhere
        const uint32_t n = min(ps.pcount[wi], ps.pcap);
        if (n == 0) continue;
        const uint32_t unit_rel = ps.big_unit[ps.part_big[wi]] - first_unit;
        const uint64_t *__restrict__ recs = ps.recs + (uint64_t)wi * ps.pcap;
        if (tid == 0) { s_sp = 1; s_sj[0] = 0; s_sq[0] = 0; }
        __syncthreads();
        while (true) {
            const uint32_t sp = s_sp;
            if (sp == 0) break;
            const uint32_t j = s_sj[sp - 1], q = s_sq[sp - 1];
            __syncthreads();
            const uint32_t TS = j == 0 ? min((uint32_t)TS_STATIC, hash_table_slots(n)) : (uint32_t)TS_STATIC;
            for (uint32_t i = tid; i < TS; i += THREADS) { K[i] = HASH_EMPTY; C[i] = 0u; }
            if (tid < 5) s_cnt[tid] = 0;
            if (tid == 0) s_sp = sp - 1;
            __syncthreads();
            const uint32_t jmask = (1u << j) - 1u;
            uint32_t my_n = 0;
            bool stop = false;
            for (uint32_t i0 = tid; i0 < n && !stop; i0 += 4 * THREADS) {
                // four independent record loads in flight per thread (the records come from HBM / L2, one pass)
                uint64_t rr[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { const uint32_t i = i0 + u * THREADS; rr[u] = i < n ? __ldcs(recs + i) : 0ull; }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint64_t r = rr[u];
                    if (i0 + u * THREADS >= n || stop) continue;
                    const uint64_t key = r >> 2;
                    if (j && (sub_hash(key) & jmask) != q) continue;
                    ++my_n;
                    if (*reinterpret_cast<volatile uint32_t *>(&s_cnt[3])) { stop = true; continue; }   // somebody found the table full
                    if (!hash_insert_bounded(K, C, TS, key, (uint32_t)r & 3u, PROBE_LIMIT)) { s_cnt[3] = 1u; stop = true; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_n += __shfl_xor_sync(0xffffffffu, my_n, o);
            if (lane_id() == 0 && my_n) atomicAdd(&s_cnt[2], my_n);
            __syncthreads();
            if (s_cnt[3]) {
                // split this key subset four ways (j <= 28 keeps the stack within bounds: 3 net entries per level)
                if (tid == 0) {
                    uint32_t p2 = s_sp;
                    if (j <= 28 && p2 + 4 <= 64) {
                        for (uint32_t c = 0; c < 4; c++) { s_sj[p2] = j + 2; s_sq[p2] = q | (c << j); ++p2; }
                        s_sp = p2;
                    } else *out.overflow = 2u;   // cannot happen with < 2^31 records; reported rather than looping
                }
                __syncthreads();
                continue;
            }
            // ---- count survivors, reserve inside the unit's region, write
            {
                uint32_t my_keep = 0, my_occ = 0;
                for (uint32_t i = tid; i < TS; i += THREADS) {
                    const uint64_t kk = K[i];
                    if (kk == HASH_EMPTY) continue;
                    ++my_occ;
                    const uint32_t cc = C[i];
                    const uint32_t cnt = slot_count(cc), fl = cc >> 30;
                    if ((cnt >> ((fl == 3u) ? 1 : 0)) >= min_mult) ++my_keep;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { my_keep += __shfl_xor_sync(0xffffffffu, my_keep, o); my_occ += __shfl_xor_sync(0xffffffffu, my_occ, o); }
                if (lane_id() == 0) { if (my_keep) atomicAdd(&s_cnt[0], my_keep); if (my_occ) atomicAdd(&s_cnt[1], my_occ); }
            }
            __syncthreads();
            if (tid == 0) {
                const uint32_t S = s_cnt[0];
                s_base = out.static_off[unit_rel] + atomicAdd(&out.unit_out_cnt[unit_rel], S);
                out.unit_out_off[unit_rel] = out.static_off[unit_rel];
                out.stats(S, s_cnt[1], s_cnt[2]);
            }
            __syncthreads();
            const unsigned long long gbase = s_base;
            for (uint32_t base = 0; base < TS; base += THREADS) {
                const uint32_t i = base + tid;
                uint64_t kk = HASH_EMPTY;
                uint32_t cf = 0;
                if (i < TS) {
                    kk = K[i];
                    if (kk != HASH_EMPTY) {
                        const uint32_t cc = C[i];
                        const uint32_t cnt = slot_count(cc), fl = cc >> 30;
                        const uint32_t mult = cnt >> ((fl == 3u) ? 1 : 0);                   // map_entry.rs:79-84
                        if (mult >= min_mult) cf = (mult > 0x3FFFFFFFu ? 0x3FFFFFFFu : mult) | (fl << 30);
                    }
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, cf != 0);
                if (bal) {
                    uint32_t wb = 0;
                    if (lane_id() == 0) wb = atomicAdd(&s_cnt[4], (uint32_t)__popc(bal));
                    wb = __shfl_sync(0xffffffffu, wb, 0);
                    if (cf) {
                        const unsigned long long o = gbase + wb + __popc(bal & ((1u << lane_id()) - 1u));
                        out.keys[o] = kk; out.count_flags[o] = cf;
                    }
                }
            }
            __syncthreads();
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
// Unit-ordered final table.  The merge kernels leave every unit's survivors in one or several output slots, in table
// order; the final table wants them contiguous per unit and ascending by key.  Keys inside a unit are spread over the
// whole 2k-bit range, so a BIN-RANK sort does it in one pass over the data: histogram on the top key bits, exclusive
// scan, scatter into the bins, then rank = bin start + number of smaller keys inside the (tiny) bin.
//   k_finish_small   one WARP per unit with <= FIN_WCAP survivors (no block barrier at all; 128 bins)
//   k_finish_units   one CTA per larger unit: bin-rank in shared memory up to BCAP survivors (2048 bins), LSD radix
//                    sort between the final buffer and a scratch of the part's size beyond that
// Slots flagged SLOT_SORTED (sort-based merge kernels) of single-slot units are copied.
constexpr uint32_t FIN_WCAP = 512, FIN_WBINS = 128, FIN_WARPS = 4;

__device__ __forceinline__ uint32_t key_bin(uint64_t key, uint32_t kbits, uint32_t log_bins) {
    return kbits > log_bins ? (uint32_t)(key >> (kbits - log_bins)) : (uint32_t)key;
}

struct FinishArgs {
    const uint64_t *src_keys; const uint32_t *src_cf;
    const uint64_t *slot_off; const uint32_t *slot_cnt;
    const uint32_t *slot_of_unit;      // n_units + 1, or NULL (slot == unit)
    const uint64_t *dst_off;           // n_units + 1 final offsets
    uint64_t *dst_keys; uint32_t *dst_cf;
    uint64_t *tmp_keys; uint32_t *tmp_cf;   // scratch of the part's size (units sorted in global memory)
    uint32_t n_units, kbits;           // kbits = 2k: keys are < 2^kbits
    uint64_t capacity; uint32_t *overflow;
};

__global__ void __launch_bounds__(FIN_WARPS * 32) k_finish_small(FinishArgs a) {
    __shared__ uint64_t s_k[FIN_WARPS][FIN_WCAP];
    __shared__ uint32_t s_v[FIN_WARPS][FIN_WCAP];
    __shared__ uint32_t s_bin[FIN_WARPS][FIN_WBINS + 1], s_cur[FIN_WARPS][FIN_WBINS];
    // the final table grows with the survivors actually seen: when this part does not fit nothing is written and the
    // host enlarges the table and launches the gather again (overflow bit 2)
    if (a.dst_off[a.n_units] > a.capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.overflow, 4u);
        return;
    }
    const uint32_t lane = lane_id(), w = warp_id();
    uint64_t *sk = s_k[w]; uint32_t *sv = s_v[w], *bin = s_bin[w], *cur = s_cur[w];
    for (uint32_t u = blockIdx.x * FIN_WARPS + w; u < a.n_units; u += gridDim.x * FIN_WARPS) {
        const uint64_t d = a.dst_off[u];
        const uint32_t n = (uint32_t)(a.dst_off[u + 1] - d);
        if (n == 0 || n > FIN_WCAP) continue;
        const uint32_t s0 = a.slot_of_unit ? a.slot_of_unit[u] : u, s1 = a.slot_of_unit ? a.slot_of_unit[u + 1] : u + 1;
        if (s1 - s0 == 1 && (a.slot_cnt[s0] & SLOT_SORTED)) {
            const uint64_t so = a.slot_off[s0];
            for (uint32_t i = lane; i < n; i += 32) { a.dst_keys[d + i] = a.src_keys[so + i]; a.dst_cf[d + i] = a.src_cf[so + i]; }
            continue;
        }
        for (uint32_t i = lane; i < FIN_WBINS; i += 32) bin[i] = 0;
        __syncwarp();
        for (uint32_t sl = s0; sl < s1; sl++) {
            const uint32_t c = a.slot_cnt[sl] & ~SLOT_SORTED;
            const uint64_t so = a.slot_off[sl];
            for (uint32_t i = lane; i < c; i += 32) atomicAdd(&bin[key_bin(a.src_keys[so + i], a.kbits, 7)], 1u);
        }
        __syncwarp();
        {   // exclusive scan of the 128 bin counts: 4 per lane
            uint32_t v[4], sum = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) { v[q] = bin[lane * 4 + q]; sum += v[q]; }
            uint32_t x = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
            uint32_t p = x - sum;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; q++) { bin[lane * 4 + q] = p; cur[lane * 4 + q] = p; p += v[q]; }
            if (lane == 31) bin[FIN_WBINS] = p;
        }
        __syncwarp();
        for (uint32_t sl = s0; sl < s1; sl++) {
            const uint32_t c = a.slot_cnt[sl] & ~SLOT_SORTED;
            const uint64_t so = a.slot_off[sl];
            for (uint32_t i = lane; i < c; i += 32) {
                const uint64_t key = a.src_keys[so + i];
                const uint32_t pos = atomicAdd(&cur[key_bin(key, a.kbits, 7)], 1u);
                sk[pos] = key; sv[pos] = a.src_cf[so + i];
            }
        }
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint64_t key = sk[i];
            const uint32_t b = key_bin(key, a.kbits, 7);
            const uint32_t lo = bin[b], hi = bin[b + 1];
            uint32_t rank = lo;
            for (uint32_t j = lo; j < hi; j++) rank += sk[j] < key ? 1u : 0u;
            a.dst_keys[d + rank] = key; a.dst_cf[d + rank] = sv[i];
        }
        __syncwarp();
    }
}

template <int THREADS, int BCAP>
constexpr size_t finish_units_smem_bytes() { return (size_t)(THREADS / 32) * 256 * 4 + 48 * 4 + (size_t)BCAP * 12 + (2048 * 2 + 1) * 4; }

template <int THREADS, int BCAP>
__global__ void __launch_bounds__(THREADS) k_finish_units(FinishArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WARPS = THREADS / 32;
    constexpr uint32_t NB = 2048;
    uint64_t *sk = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *sv = reinterpret_cast<uint32_t *>(sk + BCAP);
    uint32_t *bin = sv + BCAP, *cur = bin + NB + 1;
    uint32_t *hist = cur + NB;                       // radix histograms (global path)
    uint32_t *s_scan = hist + WARPS * 256;
    if (a.dst_off[a.n_units] > a.capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.overflow, 4u);
        return;
    }
    const uint64_t tmp_base = a.dst_off[0];          // tmp_* hold this part only
    const uint32_t tid = threadIdx.x;
    for (uint32_t u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const uint64_t d = a.dst_off[u];
        const uint32_t n = (uint32_t)(a.dst_off[u + 1] - d);
        if (n <= FIN_WCAP) continue;                 // k_finish_small
        const uint32_t s0 = a.slot_of_unit ? a.slot_of_unit[u] : u, s1 = a.slot_of_unit ? a.slot_of_unit[u + 1] : u + 1;
        if (s1 - s0 == 1 && (a.slot_cnt[s0] & SLOT_SORTED)) {
            const uint64_t so = a.slot_off[s0];
            for (uint32_t i = tid; i < n; i += THREADS) { a.dst_keys[d + i] = a.src_keys[so + i]; a.dst_cf[d + i] = a.src_cf[so + i]; }
            continue;
        }
        if (n <= (uint32_t)BCAP) {
            for (uint32_t i = tid; i < NB; i += THREADS) bin[i] = 0;
            __syncthreads();
            for (uint32_t sl = s0; sl < s1; sl++) {
                const uint32_t c = a.slot_cnt[sl] & ~SLOT_SORTED;
                const uint64_t so = a.slot_off[sl];
                for (uint32_t i = tid; i < c; i += THREADS) atomicAdd(&bin[key_bin(a.src_keys[so + i], a.kbits, 11)], 1u);
            }
            __syncthreads();
            {
                constexpr uint32_t PER = NB / THREADS;
                uint32_t v[PER], sum = 0;
#pragma unroll
                for (uint32_t q = 0; q < PER; q++) { v[q] = bin[tid * PER + q]; sum += v[q]; }
                uint32_t tot;
                uint32_t p = block_exclusive_scan<THREADS>(sum, s_scan, &tot);
#pragma unroll
                for (uint32_t q = 0; q < PER; q++) { bin[tid * PER + q] = p; cur[tid * PER + q] = p; p += v[q]; }
                if (tid == THREADS - 1) bin[NB] = p;
            }
            __syncthreads();
            for (uint32_t sl = s0; sl < s1; sl++) {
                const uint32_t c = a.slot_cnt[sl] & ~SLOT_SORTED;
                const uint64_t so = a.slot_off[sl];
                for (uint32_t i = tid; i < c; i += THREADS) {
                    const uint64_t key = a.src_keys[so + i];
                    const uint32_t pos = atomicAdd(&cur[key_bin(key, a.kbits, 11)], 1u);
                    sk[pos] = key; sv[pos] = a.src_cf[so + i];
                }
            }
            __syncthreads();
            for (uint32_t i = tid; i < n; i += THREADS) {
                const uint64_t key = sk[i];
                const uint32_t b = key_bin(key, a.kbits, 11);
                const uint32_t lo = bin[b], hi = bin[b + 1];
                uint32_t rank = lo;
                for (uint32_t j = lo; j < hi; j++) rank += sk[j] < key ? 1u : 0u;
                a.dst_keys[d + rank] = key; a.dst_cf[d + rank] = sv[i];
            }
            __syncthreads();
            continue;
        }
        // larger than the shared buffers: the same bin-rank sort with the entries in the scratch buffer (the unit's
        // slice of it is L2-resident) and 8192 bins in the shared memory the staging would have used
        constexpr uint32_t NBG = 8192, GSKEW = 1024;
        static_assert((size_t)BCAP * 12 >= (size_t)(2 * NBG + 1) * 4, "global-path bins reuse the staging area");
        uint32_t *gbin = reinterpret_cast<uint32_t *>(smem_raw), *gcur = gbin + NBG + 1;
        uint64_t *B = a.tmp_keys + (d - tmp_base);
        uint32_t *Bv = a.tmp_cf + (d - tmp_base);
        for (uint32_t i = tid; i < NBG; i += THREADS) gbin[i] = 0;
        __syncthreads();
        for (uint32_t sl = s0; sl < s1; sl++) {
            const uint32_t c = a.slot_cnt[sl] & ~SLOT_SORTED;
            const uint64_t so = a.slot_off[sl];
            for (uint32_t i = tid; i < c; i += THREADS) atomicAdd(&gbin[key_bin(a.src_keys[so + i], a.kbits, 13)], 1u);
        }
        __syncthreads();
        uint32_t maxbin = 0;
        {
            constexpr uint32_t PER = NBG / THREADS;
            uint32_t v[PER], sum = 0;
#pragma unroll
            for (uint32_t q = 0; q < PER; q++) { v[q] = gbin[tid * PER + q]; sum += v[q]; maxbin = max(maxbin, v[q]); }
            uint32_t tot;
            uint32_t p = block_exclusive_scan<THREADS>(sum, s_scan, &tot);
#pragma unroll
            for (uint32_t q = 0; q < PER; q++) { gbin[tid * PER + q] = p; gcur[tid * PER + q] = p; p += v[q]; }
            if (tid == THREADS - 1) gbin[NBG] = p;
        }
        maxbin = __syncthreads_or(maxbin > GSKEW ? 1 : 0);
        if (!maxbin) {
            for (uint32_t sl = s0; sl < s1; sl++) {
                const uint32_t c = a.slot_cnt[sl] & ~SLOT_SORTED;
                const uint64_t so = a.slot_off[sl];
                for (uint32_t i = tid; i < c; i += THREADS) {
                    const uint64_t key = a.src_keys[so + i];
                    const uint32_t pos = atomicAdd(&gcur[key_bin(key, a.kbits, 13)], 1u);
                    B[pos] = key; Bv[pos] = a.src_cf[so + i];
                }
            }
            __syncthreads();
            for (uint32_t i = tid; i < n; i += THREADS) {
                const uint64_t key = B[i];
                const uint32_t b = key_bin(key, a.kbits, 13);
                const uint32_t lo = gbin[b], hi = gbin[b + 1];
                uint32_t rank = lo;
                for (uint32_t j = lo; j < hi; j++) rank += B[j] < key ? 1u : 0u;
                a.dst_keys[d + rank] = key; a.dst_cf[d + rank] = Bv[i];
            }
            __syncthreads();
            continue;
        }
        // heavily skewed keys (a bin of > GSKEW entries): concatenate into the final buffer, LSD radix sort against the scratch
        uint64_t *A = a.dst_keys + d;
        uint32_t *Av = a.dst_cf + d;
        uint32_t run = 0;
        for (uint32_t sl = s0; sl < s1; sl++) {
            const uint32_t c = a.slot_cnt[sl] & ~SLOT_SORTED;
            const uint64_t so = a.slot_off[sl];
            for (uint32_t i = tid; i < c; i += THREADS) { A[run + i] = a.src_keys[so + i]; Av[run + i] = a.src_cf[so + i]; }
            run += c;
        }
        __syncthreads();
        uint32_t *Vs = nullptr;
        const uint32_t end_bit = min(64u, (a.kbits + 7u) & ~7u);
        uint64_t *Ss = block_radix_sort64<THREADS, true>(A, B, n, 0, end_bit, hist, s_scan, Av, Bv, &Vs);
        if (Ss != A)
            for (uint32_t i = tid; i < n; i += THREADS) { A[i] = Ss[i]; Av[i] = Vs[i]; }
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
            if (slot_of_unit) { for (uint32_t s = slot_of_unit[i]; s < slot_of_unit[i + 1]; s++) v += slot_cnt[s] & ~SLOT_SORTED; }
            else v = slot_cnt[i] & ~SLOT_SORTED;
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
    return (size_t)TS_STATIC * 12 + merge_hash_stage_bytes<THREADS>() + 40 * 4 + 16;
}

template <int THREADS, int CAP>
constexpr size_t merge_smem_bytes(bool global_scratch) {
    return (size_t)(THREADS / 32) * 256 * 4 + 40 * 4 + 16 + (global_scratch ? 0 : (size_t)CAP * 16);
}

}  // namespace ggb
