// merge128.cuh -- phase 2 for wide keys: 128-bit k-mer identities and coloured builds.
//
// Same unit semantics as merge.cuh (hashmap.rs:361-409 add_sequence + map_entry.rs:33-84), for the key
// families the 64-bit path cannot hold:
//   MODE_SEQ128  canonical / forward seq-hash of 32 <= k <= 64 (crates/hashes/src/base/cn_seqhash_base.rs:22-70,
//                fw_seqhash_base.rs:91-107, u128 instantiation cn_seqhash.rs:15-27)
//   MODE_RK128   canonical / forward Rabin-Karp mod 2^128 (crates/hashes/src/base/cn_rkhash_base.rs:64-110,
//                fw_rkhash_base.rs:57-69, constants cn_rkhash.rs:65-78, M^(k-1) from hashes/src/lib.rs:169-191)
//   MODE_COLOR   seq-hash key of k <= 48 with the colour id of the super-k-mer appended:
//                slot key = (k-mer << 32) | colour, so the table holds one entry per distinct (k-mer, colour)
//                -- the device form of "colours appended per occurrence, sort_unstable + dedup"
//                (crates/colors/src/managers/multiple.rs:198-203,221-223).
// Pipeline per bucket range:
//   k_merge_hash128   one CTA per unit: 128-bit CAS hash table (shared memory, or an L2/HBM scratch slice for big
//                     units), MapEntry word per slot; surviving slots are appended UNSORTED to the range output
//   k_sort_units128   one CTA per unit: LSD radix sort of the unit's entries by key (shared memory when they fit,
//                     global ping-pong otherwise) into the unit-ordered final table
//   k_color_count / k_color_write   (MODE_COLOR) fold the (k-mer, colour) entries of each k-mer: sum counters, OR
//                     flags, halve if flags == 3, apply -s, emit CSR colour lists
#pragma once
#include "merge.cuh"

namespace ggb {

typedef unsigned __int128 u128;

struct __align__(16) K128 {
    unsigned long long lo, hi;
};

enum { MODE_SEQ128 = 0, MODE_RK128 = 1, MODE_COLOR = 2 };
constexpr int RK_MAXK = 128;          // rabin-karp128 k-mer lengths (seq-hash keys hold at most 64 bases)
constexpr int SRC_MAXW = RK_MAXK / 32;  // 64-bit words of source bases per entry

// Rabin-Karp tables, indexed by 2-bit base code (A0 C1 T2 G3): cn_rkhash_base.rs:10-44 + cn_rkhash.rs:69-75.
struct RkTables {
    K128 mult;          // MULTIPLIER
    K128 mult_inv;      // MULT_INV
    K128 fwd[4];        // L[c]
    K128 bkw[4];        // L[c ^ 2]
    K128 fwd_mk[4];     // L[c] * M^k        (leaving base, forward hash)
    K128 bkw_mk1[4];    // L[c ^ 2] * M^(k-1) (entering base, reverse hash)
    const K128 *pos_fwd;  // device [RK_MAXK][4]: L[c] * M^j       -- the first k-mer of a super-k-mer is a sum of table terms
    const K128 *pos_bkw;  // device [RK_MAXK][4]: L[c ^ 2] * M^j      (k 128-bit adds instead of k 128-bit multiplies)
};

__host__ __device__ __forceinline__ u128 to_u128(K128 v) { return ((u128)v.hi << 64) | (u128)v.lo; }
__host__ __device__ __forceinline__ K128 to_k128(u128 v) { return K128{(unsigned long long)v, (unsigned long long)(v >> 64)}; }

struct MergeOut128 {
    uint64_t *keys_lo, *keys_hi;
    uint32_t *count_flags;
    unsigned long long *cursor;   // [0] entries written (allocation cursor of the dynamic region), [1] distinct slot keys,
                                  // [2] k-mer occurrences, [3] entries written into the static regions of big units
    uint64_t *unit_out_off;
    uint32_t *unit_out_cnt;
    uint64_t capacity;            // dynamic region [0, capacity); the static regions of partitioned units follow it
    uint32_t *overflow;
    uint64_t *src;                // MODE_RK128: packed bases of one occurrence of every surviving key, src_words words per entry
                                  // (the hash is not invertible: the reference keeps `saved_reads` for this, hashmap.rs:32-33,96-149)
    uint32_t src_words, pad;      // ceil(k / 32), at least 2
};

// Key partitions of big units (same scheme as merge.cuh, 128-bit keys): k_partition_units128 expands a unit once and
// routes every record by a hash of its key into partitions of <= pcap records in HBM; k_merge_hash128<.., SRC_RECORDS>
// counts one partition per CTA in shared memory and appends its survivors to the unit's static output region.
struct PartSrc128 {
    const uint64_t *rec_lo, *rec_hi;   // [n_parts_total][pcap]
    const uint8_t *rec_fl;             // flag bits of the record
    const uint64_t *rec_src;           // MODE_RK128: source locator of the record (src_locator)
    const uint32_t *pcount;            // [n_parts_total]
    const uint32_t *part_big;          // work item -> index of its big unit
    const uint32_t *big_unit;          // [n_big] unit id
    const uint32_t *big_ovf;           // [n_big] 1 = a partition overflowed: the unit is redone by the global-table kernel
    const uint64_t *big_off;           // [n_big] start of the unit's static output region
    uint32_t *big_fill;                // [n_big] entries handed out inside the region
    uint32_t pcap, pad;
};

constexpr unsigned long long EMPTY64 = ~0ull;

__device__ __forceinline__ uint32_t mix128(unsigned long long lo, unsigned long long hi) {
    unsigned long long x = (lo ^ (hi * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
    return (uint32_t)(x >> 32);
}

// Insert one k-mer occurrence: claim / find the slot with a 128-bit CAS (ATOMS.CAS.128 / ATOMG.CAS.128),
// then MapEntry::incr_by_and_check + update_flags on the slot word.
// The all-ones key doubles as the empty-slot sentinel, but it IS a legal key when the key uses all 128 bits (forward-only
// seq-hash of k = 64: poly-G; colours with k = 48; any rk128 hash): such occurrences are counted in `special`, a
// MapEntry word beside the table that the scan treats as one more slot.
// Returns the slot (SLOT_SPECIAL for the all-ones key); *claimed = this call created the entry (it then records where
// the k-mer's bases can be found when the hash is not invertible).
// Shared-memory tables are sized by the distinct keys a unit is EXPECTED to hold (host: distinct / records of what this
// context merged before), so an insert probes a bounded number of slots; SLOT_FULL = give up, the unit is redone by the
// global-table kernel (probe_limit 0 = unbounded: tables sized by the records themselves).
constexpr uint32_t SLOT_SPECIAL = 0xFFFFFFFFu, SLOT_FULL = 0xFFFFFFFEu;
__device__ __forceinline__ uint32_t hash_insert128(K128 *K, uint32_t *C, uint32_t mask, u128 key, uint32_t fb, uint32_t *special,
                                                   bool *claimed, uint32_t probe_limit = 0) {
    const K128 want = to_k128(key);
    const K128 empty = K128{EMPTY64, EMPTY64};
    if (want.lo == EMPTY64 && want.hi == EMPTY64) {
        *claimed = (atomicAdd(special, 1u) & 0x3FFFFFFFu) == 0u;
        if (fb) atomicOr(special, fb << 30);
        return SLOT_SPECIAL;
    }
    uint32_t slot = mix128(want.lo, want.hi) & mask;
    for (uint32_t probes = 1;; ++probes) {
        const K128 old = atomicCAS(&K[slot], empty, want);
        if (old.lo == EMPTY64 && old.hi == EMPTY64) { *claimed = true; break; }
        if (old.lo == want.lo && old.hi == want.hi) { *claimed = false; break; }
        if (probes == probe_limit) { *claimed = false; return SLOT_FULL; }
        slot = (slot + 1) & mask;
    }
    atomicAdd(&C[slot], 1u);
    if (fb) atomicOr(&C[slot], fb << 30);
    return slot;
}

// Where one k-mer occurrence lives: payload word of its super-k-mer inside chunk `c` (32 bits), k-mer index inside the
// super-k-mer (16 bits), 1 = the key is the hash of the reverse complement, chunk index (15 bits).
__device__ __forceinline__ uint64_t src_locator(uint32_t chunk, uint32_t word, uint32_t i, bool isf) {
    return (uint64_t)word | ((uint64_t)(i & 0xFFFFu) << 32) | ((uint64_t)(isf ? 0u : 1u) << 48) | ((uint64_t)chunk << 49);
}

__device__ __forceinline__ u128 revcomp128(u128 x) {
    return ((u128)revcomp64((uint64_t)x) << 64) | (u128)revcomp64((uint64_t)(x >> 64));
}

__device__ __forceinline__ uint32_t flag_bits(uint32_t flags, uint32_t i, uint32_t last, bool isf) {
    const uint32_t bi = (!(flags & READ_FLAG_INCL_BEGIN) && i == 0) ? 1u : 0u;
    const uint32_t ei = (!(flags & READ_FLAG_INCL_END) && i == last) ? 1u : 0u;
    return (bi << (isf ? 0 : 1)) | (ei << (isf ? 1 : 0));  // hashmap.rs:385-399
}

// All k-mers of one stored super-k-mer -> f(key, flag bits).
template <int MODE, typename F>
__device__ __forceinline__ void for_each_kmer128(const uint32_t *__restrict__ pl, uint32_t len, uint32_t flags, uint32_t k,
                                                 uint32_t forward_only, const RkTables &T, F f) {
    const uint32_t last = len - k;
    if (MODE != MODE_RK128) {
        const u128 mask = (k >= 64) ? ~(u128)0 : ((((u128)1) << (2 * k)) - 1);
        u128 fw = 0;
        const uint32_t nw = (2 * k + 31) >> 5;
        for (uint32_t w = 0; w < nw; w++) fw |= (u128)pl[w] << (32 * w);
        fw &= mask;
        u128 rc = revcomp128(fw) >> (128 - 2 * k);
        uint32_t cw = 0;
        for (uint32_t i = 0;; ++i) {
            const bool isf = forward_only ? true : (fw < rc);
            const u128 key = forward_only ? fw : (fw < rc ? fw : rc);
            f(key, flag_bits(flags, i, last, isf), i, isf);
            if (i == last) break;
            const uint32_t nb = i + k;
            if ((nb & 15u) == 0 || i == 0) cw = pl[nb >> 4];
            const u128 b = (cw >> (2u * (nb & 15u))) & 3u;
            fw = (fw >> 2) | (b << (2 * (k - 1)));
            rc = ((rc << 2) | (b ^ (u128)2)) & mask;
        }
    } else {
        const u128 M = to_u128(T.mult), MI = to_u128(T.mult_inv);
        // fw = sum L[b_i] M^(k-1-i), rc = sum L[b_i ^ 2] M^i  (cn_rkhash_base.rs:66-108), from the per-position tables
        u128 fw = 0, rc = 0;
        uint32_t cw0 = 0;
        for (uint32_t i = 0; i < k; i++) {
            if ((i & 15u) == 0) cw0 = pl[i >> 4];
            const uint32_t b = (cw0 >> (2u * (i & 15u))) & 3u;
            const ulonglong2 f = __ldg(reinterpret_cast<const ulonglong2 *>(T.pos_fwd + (k - 1 - i) * 4 + b));
            const ulonglong2 r = __ldg(reinterpret_cast<const ulonglong2 *>(T.pos_bkw + i * 4 + b));
            fw += ((u128)f.y << 64) | (u128)f.x;
            rc += ((u128)r.y << 64) | (u128)r.x;
        }
        for (uint32_t i = 0;; ++i) {
            const bool isf = forward_only ? true : (fw < rc);
            const u128 key = forward_only ? fw : (fw < rc ? fw : rc);
            f(key, flag_bits(flags, i, last, isf), i, isf);
            if (i == last) break;
            const uint32_t co = packed_base(pl, i), ci = packed_base(pl, i + k);
            fw = fw * M - to_u128(T.fwd_mk[co]) + to_u128(T.fwd[ci]);
            rc = (rc - to_u128(T.bkw[co])) * MI + to_u128(T.bkw_mk1[ci]);
        }
    }
}

// bytes per table slot: 16 key + 4 MapEntry word (+ 8 source locator when the hash is not invertible)
template <int MODE>
constexpr uint32_t slot_bytes128() { return MODE == MODE_RK128 ? 28u : 20u; }

template <int THREADS, int TS_STATIC, int MODE>
constexpr size_t merge_hash128_smem_bytes() {
    return (size_t)TS_STATIC * slot_bytes128<MODE>() + 64;
}

// The k bases of a packed super-k-mer starting at base i, reverse-complemented when rc -- the bases whose FORWARD hash is
// the table key -- as `nwords` (= max(2, ceil(k / 32))) 64-bit words at dst (base j at bits 2(j % 32) of word j / 32).
__device__ __forceinline__ void kmer_bases_store(const uint32_t *__restrict__ pl, uint32_t i, uint32_t k, bool rc, uint64_t *dst,
                                                 uint32_t nwords) {
    const uint32_t w0 = i >> 4, sh = 2u * (i & 15u), n32 = (sh + 2u * k + 31u) >> 5;   // 32-bit words read, all inside the super-k-mer
    uint64_t v[SRC_MAXW];
#pragma unroll
    for (int q = 0; q < SRC_MAXW; q++) {
        // 64 bits starting at bit sh + 64 q of the stream pl[w0 ...]
        const uint32_t a = 2u * q < n32 ? pl[w0 + 2u * q] : 0u, b = 2u * q + 1u < n32 ? pl[w0 + 2u * q + 1u] : 0u,
                       c = 2u * q + 2u < n32 ? pl[w0 + 2u * q + 2u] : 0u;
        v[q] = ((uint64_t)__funnelshift_r(b, c, sh) << 32) | (uint64_t)__funnelshift_r(a, b, sh);
    }
    // clear everything above 2k bits
#pragma unroll
    for (int q = 0; q < SRC_MAXW; q++) {
        const uint32_t lo = 64u * q;
        if (2u * k <= lo) v[q] = 0;
        else if (2u * k < lo + 64u) v[q] &= (1ull << (2u * k - lo)) - 1ull;
    }
    if (rc) {
        // reverse complement of the 64 SRC_MAXW-bit string, then shift right by (64 SRC_MAXW - 2k) bits
        uint64_t r[SRC_MAXW];
#pragma unroll
        for (int q = 0; q < SRC_MAXW; q++) r[q] = revcomp64(v[SRC_MAXW - 1 - q]);
        const uint32_t s = 64u * SRC_MAXW - 2u * k, sw = s >> 6, sb = s & 63u;
#pragma unroll
        for (int q = 0; q < SRC_MAXW; q++) {
            const uint32_t a = q + sw, b = a + 1u;
            uint64_t lo = 0, hi = 0;
#pragma unroll
            for (int t = 0; t < SRC_MAXW; t++) { if ((uint32_t)t == a) lo = r[t]; if ((uint32_t)t == b) hi = r[t]; }
            v[q] = sb ? ((lo >> sb) | (hi << (64u - sb))) : lo;
        }
        // revcomp64 complements the padding too: clear above 2k bits again
#pragma unroll
        for (int q = 0; q < SRC_MAXW; q++) {
            const uint32_t lo = 64u * q;
            if (2u * k <= lo) v[q] = 0;
            else if (2u * k < lo + 64u) v[q] &= (1ull << (2u * k - lo)) - 1ull;
        }
    }
#pragma unroll
    for (int q = 0; q < SRC_MAXW; q++)
        if ((uint32_t)q < nwords) dst[q] = v[q];
}

// TS_STATIC > 0: table in shared memory (units with <= 3/4 TS_STATIC records).  TS_STATIC == 0: table in this CTA's
// slice of `scratch` (hash_table_slots_pow2(n) slots of 20 bytes).
template <int THREADS, int TS_STATIC, int MODE, int SRC = SRC_SUPERKMERS>
__global__ void __launch_bounds__(THREADS)
k_merge_hash128(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ work, uint32_t n_work,
                uint32_t first_unit, DevParams P, RkTables T, uint32_t min_mult, MergeOut128 out,
                uint64_t *__restrict__ scratch, uint64_t per_cta_u64, PartSrc128 ps, const uint32_t *__restrict__ n_work_dev,
                uint32_t *__restrict__ retry /* [0] = count, [1..] = units whose shared table filled up; may be NULL */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool WITH_SRC = MODE == MODE_RK128;
    K128 *K = reinterpret_cast<K128 *>(smem_raw);
    uint64_t *L = reinterpret_cast<uint64_t *>(K + TS_STATIC);                       // WITH_SRC: source locator per slot
    uint32_t *C = reinterpret_cast<uint32_t *>(K + TS_STATIC) + (WITH_SRC ? 2 * TS_STATIC : 0);
    __shared__ uint32_t s_cnt[2];
    __shared__ uint32_t s_special;         // MapEntry word of the all-ones key (see hash_insert128)
    __shared__ uint32_t s_full;            // an insert ran out of probes: the unit goes to the retry list
    // bounded probing only where the table was sized from an expectation (shared tables fed by super-k-mers)
    const uint32_t probe_limit = (TS_STATIC > 0 && SRC == SRC_SUPERKMERS && retry) ? 96u : 0u;
    __shared__ unsigned long long s_special_src;
    __shared__ unsigned long long s_base;
    const uint32_t tid = threadIdx.x;
    if (n_work_dev) n_work = min(n_work, *n_work_dev);   // device-side list (big units whose partitions overflowed)
    for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        uint32_t unit = 0, n = 0;
        uint32_t nx_n = 0;   // SRC_RECORDS: record count of this CTA's NEXT partition (loaded now, used after the inserts to
                             // pull that partition's records into L2 while this one's table is scanned)
        if (SRC == SRC_RECORDS) {
            if (wi + gridDim.x < n_work) nx_n = min(ps.pcount[wi + gridDim.x], ps.pcap);
            if (ps.big_ovf[ps.part_big[wi]]) continue;   // the whole unit is redone from its super-k-mers
            unit = ps.big_unit[ps.part_big[wi]];
            n = min(ps.pcount[wi], ps.pcap);
            if (n == 0) continue;
        } else {
            unit = work[wi];
            for (uint32_t c = 0; c < n_chunks; c++) {
                const ChunkView &cv = chunks[c];
                if (unit >= cv.first_unit && unit < cv.first_unit + cv.n_units) n += cv.unit_kmers[unit - cv.first_unit];
            }
        }
        uint32_t TS = TS_STATIC;
        if (TS_STATIC == 0) {
            TS = hash_table_slots_pow2(n);
            K = reinterpret_cast<K128 *>(scratch + (uint64_t)blockIdx.x * per_cta_u64);
            L = reinterpret_cast<uint64_t *>(K + TS);
            C = reinterpret_cast<uint32_t *>(K + TS) + (WITH_SRC ? 2 * (size_t)TS : 0);
        }
        const uint32_t tmask = TS - 1;
        for (uint32_t i = tid; i < TS; i += THREADS) { K[i] = K128{EMPTY64, EMPTY64}; C[i] = 0u; }
        if (tid < 2) s_cnt[tid] = 0;
        if (tid == 0) { s_special = 0; s_full = 0; }
        __syncthreads();
        if (SRC == SRC_RECORDS) {
            const uint64_t ro = (uint64_t)wi * ps.pcap;
            for (uint32_t i = tid; i < n; i += THREADS) {
                bool claimed;
                const uint32_t slot = hash_insert128(K, C, tmask, ((u128)ps.rec_hi[ro + i] << 64) | (u128)ps.rec_lo[ro + i],
                                                     ps.rec_fl[ro + i], &s_special, &claimed);
                if (WITH_SRC && claimed) { if (slot == SLOT_SPECIAL) s_special_src = ps.rec_src[ro + i]; else L[slot] = ps.rec_src[ro + i]; }
            }
            if (nx_n) {   // 128-byte lines of the next partition's records (keys: 16 per line and array, flags: 128 per line)
                const uint64_t rn = (uint64_t)(wi + gridDim.x) * ps.pcap;
                const uint32_t kl = (nx_n + 15u) >> 4, fl = (nx_n + 127u) >> 7;
                for (uint32_t q = tid; q < 2u * kl + fl; q += THREADS) {
                    const void *pa = q < kl ? (const void *)(ps.rec_lo + rn + 16u * q)
                                   : q < 2u * kl ? (const void *)(ps.rec_hi + rn + 16u * (q - kl)) : (const void *)(ps.rec_fl + rn + 128u * (q - 2u * kl));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
                }
            }
        } else {
            for (uint32_t c = 0; c < n_chunks; c++) {
                const ChunkView cv = chunks[c];
                if (unit < cv.first_unit || unit >= cv.first_unit + cv.n_units) continue;
                const uint32_t d0 = cv.unit_off[unit - cv.first_unit], d1 = cv.unit_off[unit - cv.first_unit + 1];
                for (uint32_t di = d0 + tid; di < d1; di += THREADS) {
                    const uint4 d = cv.desc[di];
                    const uint32_t color = d.w;
                    for_each_kmer128<MODE>(cv.payload + (d.x - cv.word_bias), d.y, (d.z >> 16) & 3u, P.k, P.forward_only, T,
                                           [&](u128 key, uint32_t fb, uint32_t ki, bool isf) {
                                               if (MODE == MODE_COLOR) key = (key << 32) | (u128)color;
                                               bool claimed;
                                               const uint32_t slot = hash_insert128(K, C, tmask, key, fb, &s_special, &claimed, probe_limit);
                                               if (slot == SLOT_FULL) { s_full = 1u; return; }
                                               if (WITH_SRC && claimed) {
                                                   const uint64_t loc = src_locator(c, d.x - cv.word_bias, ki, isf);
                                                   if (slot == SLOT_SPECIAL) s_special_src = loc; else L[slot] = loc;
                                               }
                                           });
                    if (probe_limit && s_full) break;   // somebody found the table full: stop feeding it
                }
            }
        }
        __syncthreads();
        if (probe_limit && s_full) {   // block-uniform: nothing of this unit is written, the global-table kernel redoes it
            if (tid == 0) retry[1u + atomicAdd(&retry[0], 1u)] = unit;
            __syncthreads();           // s_full is reset at the top of the next work item
            continue;
        }
        // ---- pass 1 over the table: count survivors (MODE_COLOR: every occupied slot; the -s filter needs the
        //      per-k-mer fold, done after the sort)
        {
            uint32_t my_keep = 0, my_occ = 0;
            if (tid == 0 && (s_special & 0x3FFFFFFFu)) {   // the all-ones key: one more occupied "slot"
                const uint32_t cc = s_special, cnt = cc & 0x3FFFFFFFu, fl = cc >> 30;
                ++my_occ;
                if (MODE == MODE_COLOR || (cnt >> ((fl == 3u) ? 1 : 0)) >= min_mult) ++my_keep;
            }
            for (uint32_t i = tid; i < TS; i += THREADS) {
                const K128 kk = K[i];
                if (kk.lo == EMPTY64 && kk.hi == EMPTY64) continue;
                ++my_occ;
                if (MODE == MODE_COLOR) { ++my_keep; continue; }
                const uint32_t cc = C[i];
                const uint32_t cnt = cc & 0x3FFFFFFFu, fl = cc >> 30;
                const uint32_t mult = cnt >> ((fl == 3u) ? 1 : 0);  // map_entry.rs:79-84
                if (mult >= min_mult) {
                    ++my_keep;
                    if (WITH_SRC) {   // pass 2 reads the survivor's bases at a random payload address: start the fetch now
                        const uint64_t loc = L[i];
                        const uint32_t ki = (uint32_t)(loc >> 32) & 0xFFFFu;
                        const uint32_t *pp = chunks[loc >> 49].payload + (uint32_t)loc + (ki >> 4);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + ((2u * (ki & 15u) + 2u * P.k - 1u) >> 5)));   // last word read
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                my_keep += __shfl_xor_sync(0xffffffffu, my_keep, o);
                my_occ += __shfl_xor_sync(0xffffffffu, my_occ, o);
            }
            if (lane_id() == 0) { if (my_keep) atomicAdd(&s_cnt[0], my_keep); if (my_occ) atomicAdd(&s_cnt[1], my_occ); }
        }
        __syncthreads();
        const uint32_t S = s_cnt[0];
        if (tid == 0) {
            unsigned long long b;
            if (SRC == SRC_RECORDS) {
                // the partitions of a unit share its static region: survivors of all partitions end up contiguous
                const uint32_t bi = ps.part_big[wi];
                b = ps.big_off[bi] + atomicAdd(&ps.big_fill[bi], S);
                atomicAdd(&out.cursor[3], (unsigned long long)S);
                out.unit_out_off[unit - first_unit] = ps.big_off[bi];
                atomicAdd(&out.unit_out_cnt[unit - first_unit], S);
            } else {
                b = atomicAdd(&out.cursor[0], (unsigned long long)S);
                out.unit_out_off[unit - first_unit] = b;
                out.unit_out_cnt[unit - first_unit] = S;
                if (b + S > out.capacity) *out.overflow = 1u;
            }
            atomicAdd(&out.cursor[1], (unsigned long long)s_cnt[1]);
            atomicAdd(&out.cursor[2], (unsigned long long)n);
            s_base = b;
        }
        __syncthreads();
        const unsigned long long gbase = s_base;
        if (tid == 0) {
            uint32_t first = 0;
            if ((s_special & 0x3FFFFFFFu) && (SRC == SRC_RECORDS || gbase + S <= out.capacity)) {
                const uint32_t cc = s_special, cnt = cc & 0x3FFFFFFFu, fl = cc >> 30;
                const uint32_t mult = cnt >> ((fl == 3u) ? 1 : 0);
                if (MODE == MODE_COLOR || mult >= min_mult) {
                    out.keys_lo[gbase] = EMPTY64; out.keys_hi[gbase] = EMPTY64;
                    out.count_flags[gbase] = MODE == MODE_COLOR ? cc : (mult | (fl << 30));
                    if (WITH_SRC) {
                        const unsigned long long loc = s_special_src;
                        kmer_bases_store(chunks[loc >> 49].payload + (uint32_t)loc, (uint32_t)(loc >> 32) & 0xFFFFu, P.k, (loc >> 48) & 1u,
                                         out.src + (size_t)out.src_words * gbase, out.src_words);
                    }
                    first = 1;
                }
            }
            s_cnt[0] = first;   // the slots' survivors follow
        }
        __syncthreads();
        // ---- pass 2: append survivors (arbitrary order inside the unit's range; k_sort_units128 orders them)
        if (SRC == SRC_RECORDS || gbase + S <= out.capacity) {
            for (uint32_t base = 0; base < TS; base += THREADS) {
                const uint32_t i = base + tid;
                bool keep = false;
                K128 kk = K128{0, 0};
                uint32_t cf = 0;
                if (i < TS) {
                    kk = K[i];
                    if (!(kk.lo == EMPTY64 && kk.hi == EMPTY64)) {
                        const uint32_t cc = C[i];
                        if (MODE == MODE_COLOR) { keep = true; cf = cc; }
                        else {
                            const uint32_t cnt = cc & 0x3FFFFFFFu, fl = cc >> 30;
                            const uint32_t mult = cnt >> ((fl == 3u) ? 1 : 0);
                            if (mult >= min_mult) { keep = true; cf = mult | (fl << 30); }
                        }
                    }
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                uint32_t wbase = 0;
                if (lane_id() == 0 && bal) wbase = atomicAdd(&s_cnt[0], (uint32_t)__popc(bal));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (keep) {
                    const unsigned long long o = gbase + wbase + __popc(bal & ((1u << lane_id()) - 1u));
                    out.keys_lo[o] = kk.lo; out.keys_hi[o] = kk.hi; out.count_flags[o] = cf;
                    if (WITH_SRC) {
                        const uint64_t loc = L[i];
                        kmer_bases_store(chunks[loc >> 49].payload + (uint32_t)loc, (uint32_t)(loc >> 32) & 0xFFFFu, P.k, (loc >> 48) & 1u,
                                         out.src + (size_t)out.src_words * o, out.src_words);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_partition_units128: one CTA per big unit.  Expands the unit once (one super-k-mer per thread, rolling hashes) and
// appends every record to partition part_hash128(key) & (np - 1); the CTA owns all partitions of its unit, so the
// cursors live in shared memory.  An overflowing partition flags the unit for the global-table kernel.
__device__ __forceinline__ uint32_t part_hash128(u128 key) {   // independent of the in-table slot hash (mix128)
    const unsigned long long lo = (unsigned long long)key, hi = (unsigned long long)(key >> 64);
    return (uint32_t)(((lo * 0xD6E8FEB86659FD93ull) ^ (hi * 0x9E3779B97F4A7C15ull) ^ (hi >> 29)) >> 40);
}

template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS)
k_partition_units128(const ChunkView *__restrict__ chunks, uint32_t n_chunks, const uint32_t *__restrict__ big_unit,
                     const uint32_t *__restrict__ big_logp, const uint32_t *__restrict__ big_pbase, uint32_t n_big, DevParams P,
                     RkTables T, uint64_t *__restrict__ rec_lo, uint64_t *__restrict__ rec_hi, uint8_t *__restrict__ rec_fl,
                     uint64_t *__restrict__ rec_src /* MODE_RK128 only */, uint32_t *__restrict__ pcount, uint32_t pcap, uint32_t *__restrict__ big_ovf, uint32_t *__restrict__ retry,
                     uint32_t *__restrict__ retry_count) {
    __shared__ uint32_t s_cur[PART_MAXP];
    __shared__ uint32_t s_ovf;
    const uint32_t tid = threadIdx.x;
    for (uint32_t bi = blockIdx.x; bi < n_big; bi += gridDim.x) {
        const uint32_t unit = big_unit[bi], np = 1u << big_logp[bi], pbase = big_pbase[bi];
        for (uint32_t i = tid; i < np; i += THREADS) s_cur[i] = 0;
        if (tid == 0) s_ovf = 0;
        __syncthreads();
        const uint64_t base = (uint64_t)pbase * pcap;
        for (uint32_t c = 0; c < n_chunks; c++) {
            const ChunkView cv = chunks[c];
            if (unit < cv.first_unit || unit >= cv.first_unit + cv.n_units) continue;
            const uint32_t d0 = cv.unit_off[unit - cv.first_unit], d1 = cv.unit_off[unit - cv.first_unit + 1];
            for (uint32_t di = d0 + tid; di < d1; di += THREADS) {
                const uint4 d = cv.desc[di];
                const uint32_t color = d.w;
                for_each_kmer128<MODE>(cv.payload + (d.x - cv.word_bias), d.y, (d.z >> 16) & 3u, P.k, P.forward_only, T,
                                       [&](u128 key, uint32_t fb, uint32_t ki, bool isf) {
                                           // coloured builds: all colours of a k-mer must meet in one partition
                                           const uint32_t p = part_hash128(key) & (np - 1);
                                           if (MODE == MODE_COLOR) key = (key << 32) | (u128)color;
                                           const uint32_t pos = atomicAdd(&s_cur[p], 1u);
                                           if (pos < pcap) {
                                               const uint64_t o = base + (uint64_t)p * pcap + pos;
                                               rec_lo[o] = (uint64_t)key; rec_hi[o] = (uint64_t)(key >> 64); rec_fl[o] = (uint8_t)fb;
                                               if (MODE == MODE_RK128) rec_src[o] = src_locator(c, d.x - cv.word_bias, ki, isf);
                                           } else s_ovf = 1u;
                                       });
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < np; i += THREADS) pcount[pbase + i] = min(s_cur[i], pcap);
        if (tid == 0 && s_ovf) { big_ovf[bi] = 1u; retry[atomicAdd(retry_count, 1u)] = unit; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of n (lo, hi, val) records, 8-bit digits of the 128-bit key from first_bit to end_bit;
// passes whose digit is constant over the unit are skipped.  Generic pointers (shared or global memory).
// Returns 0 if the result is in (Alo, Ahi, Av), 1 if in (Blo, Bhi, Bv).
template <int THREADS>
__device__ int block_radix_sort128(uint64_t *Alo, uint64_t *Ahi, uint32_t *Av, uint64_t *Blo, uint64_t *Bhi, uint32_t *Bv,
                                   uint32_t n, uint32_t first_bit, uint32_t end_bit, uint32_t *hist, uint32_t *s_scan) {
    constexpr int WARPS = THREADS / 32;
    constexpr int EPT = WARPS * 256 / THREADS;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    uint32_t chunk = (n + WARPS - 1) / WARPS;
    chunk = (chunk + 31u) & ~31u;
    const uint32_t wbeg = min(n, warp * chunk), wend = min(n, wbeg + chunk);
    int where = 0;
    for (uint32_t shift = first_bit; shift < end_bit; shift += 8) {
        const uint64_t *src = shift < 64 ? Alo : Ahi;
        const uint32_t sh = shift & 63u;
        for (uint32_t i = tid; i < WARPS * 256; i += THREADS) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&hist[warp * 256 + ((uint32_t)(src[i] >> sh) & 255u)], 1u);
        __syncthreads();
        uint32_t single;
        {
            uint32_t v[EPT];
            uint32_t sum = 0, digit_tot = 0, one = 0;
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const uint32_t e = tid * EPT + q;
                v[q] = hist[(e % WARPS) * 256 + (e / WARPS)];
                sum += v[q];
            }
            // EPT consecutive entries of one thread belong to one digit when WARPS % EPT == 0 (16/8, 32/8)
            digit_tot = sum;
            for (int o = 1; o < WARPS / EPT; o <<= 1) digit_tot += __shfl_xor_sync(0xffffffffu, digit_tot, o);
            one = (digit_tot == n) ? 1u : 0u;
            uint32_t tot;
            uint32_t p = block_exclusive_scan<THREADS>(sum, s_scan, &tot);
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const uint32_t e = tid * EPT + q;
                hist[(e % WARPS) * 256 + (e / WARPS)] = p;
                p += v[q];
            }
            single = __syncthreads_or((int)one);
        }
        if (single) continue;  // every record has the same digit: the pass is the identity
        for (uint32_t base = wbeg; base < wend; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < wend;
            const uint64_t lo = valid ? Alo[i] : 0ull, hi = valid ? Ahi[i] : 0ull;
            const uint32_t d = valid ? ((uint32_t)((shift < 64 ? lo : hi) >> sh) & 255u) : (256u + lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t pos = 0;
            if (valid) pos = hist[warp * 256 + d] + rank;
            __syncwarp();
            if (valid) {
                Blo[pos] = lo; Bhi[pos] = hi; Bv[pos] = Av[i];
                if (rank + 1 == (uint32_t)__popc(peers)) hist[warp * 256 + d] = pos + 1;
            }
            __syncwarp();
        }
        __syncthreads();
        uint64_t *t = Alo; Alo = Blo; Blo = t;
        t = Ahi; Ahi = Bhi; Bhi = t;
        uint32_t *tv = Av; Av = Bv; Bv = tv;
        where ^= 1;
    }
    return where;
}

template <int THREADS, int CAP>
constexpr size_t sort_units128_smem_bytes() {
    return (size_t)(THREADS / 32) * 256 * 4 + 48 * 4 + (size_t)CAP * 40;
}

// One CTA per unit: entries [unit_out_off, +cnt) of src -> sorted at [unit_final_off, +cnt) of dst.
// The src range is scratch after this kernel (used as the ping-pong partner for big units).
// WITH_SRC (rk128): every entry also owns 2 words of source bases.  The sort then carries the entry's index inside the
// unit instead of its MapEntry word (idx_a / idx_b: scratch, laid out like src), and the MapEntry words and the bases are
// gathered through the sorted indices at the end (src_cf and src_bases are read-only here).
template <int THREADS, int CAP, bool WITH_SRC>
__global__ void __launch_bounds__(THREADS)
k_sort_units128(uint64_t *__restrict__ src_lo, uint64_t *__restrict__ src_hi, uint32_t *__restrict__ src_cf,
                const uint64_t *__restrict__ unit_out_off, const uint32_t *__restrict__ unit_out_cnt,
                const uint64_t *__restrict__ unit_final_off, uint64_t *__restrict__ dst_lo, uint64_t *__restrict__ dst_hi,
                uint32_t *__restrict__ dst_cf, uint32_t n_units, uint32_t first_bit, uint32_t end_bit, uint64_t capacity,
                uint32_t *__restrict__ overflow, const uint64_t *__restrict__ src_bases, uint64_t *__restrict__ dst_bases,
                uint32_t *__restrict__ idx_a, uint32_t *__restrict__ idx_b, uint32_t src_words) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (unit_final_off[n_units] > capacity) {   // the host enlarges the final table and launches the sort again
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(overflow, 4u);
        return;
    }
    constexpr int WARPS = THREADS / 32;
    uint64_t *sAlo = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *sAhi = sAlo + CAP, *sBlo = sAhi + CAP, *sBhi = sBlo + CAP;
    uint32_t *sAv = reinterpret_cast<uint32_t *>(sBhi + CAP), *sBv = sAv + CAP;
    uint32_t *hist = sBv + CAP;
    uint32_t *s_scan = hist + WARPS * 256;
    const uint32_t tid = threadIdx.x;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint32_t n = unit_out_cnt[u];
        if (n == 0) continue;
        const uint64_t so = unit_out_off[u], fo = unit_final_off[u];
        if (n <= (uint32_t)CAP) {
            for (uint32_t i = tid; i < n; i += THREADS) { sAlo[i] = src_lo[so + i]; sAhi[i] = src_hi[so + i]; sAv[i] = WITH_SRC ? i : src_cf[so + i]; }
            __syncthreads();
            const int w = block_radix_sort128<THREADS>(sAlo, sAhi, sAv, sBlo, sBhi, sBv, n, first_bit, end_bit, hist, s_scan);
            const uint64_t *rl = w ? sBlo : sAlo, *rh = w ? sBhi : sAhi;
            const uint32_t *rv = w ? sBv : sAv;
            for (uint32_t i = tid; i < n; i += THREADS) {
                dst_lo[fo + i] = rl[i]; dst_hi[fo + i] = rh[i];
                if (WITH_SRC) {
                    const uint64_t j = so + rv[i];
                    dst_cf[fo + i] = src_cf[j];
                    for (uint32_t q = 0; q < src_words; q++) dst_bases[(size_t)src_words * (fo + i) + q] = src_bases[(size_t)src_words * j + q];
                } else dst_cf[fo + i] = rv[i];
            }
        } else if (WITH_SRC) {
            for (uint32_t i = tid; i < n; i += THREADS) idx_a[so + i] = i;
            __syncthreads();
            const int w = block_radix_sort128<THREADS>(src_lo + so, src_hi + so, idx_a + so, dst_lo + fo, dst_hi + fo, idx_b + so,
                                                       n, first_bit, end_bit, hist, s_scan);
            const uint32_t *rv = (w ? idx_b : idx_a) + so;
            for (uint32_t i = tid; i < n; i += THREADS) {
                if (w == 0) { dst_lo[fo + i] = src_lo[so + i]; dst_hi[fo + i] = src_hi[so + i]; }
                const uint64_t j = so + rv[i];
                dst_cf[fo + i] = src_cf[j];
                for (uint32_t q = 0; q < src_words; q++) dst_bases[(size_t)src_words * (fo + i) + q] = src_bases[(size_t)src_words * j + q];
            }
        } else {
            const int w = block_radix_sort128<THREADS>(src_lo + so, src_hi + so, src_cf + so, dst_lo + fo, dst_hi + fo, dst_cf + fo,
                                                       n, first_bit, end_bit, hist, s_scan);
            if (w == 0)
                for (uint32_t i = tid; i < n; i += THREADS) { dst_lo[fo + i] = src_lo[so + i]; dst_hi[fo + i] = src_hi[so + i]; dst_cf[fo + i] = src_cf[so + i]; }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// MODE_COLOR fold.  Sorted entries of a unit: slot key = (k-mer << 32) | colour, value = counter | flags << 30.
// A k-mer's entries are adjacent; head = first entry of a k-mer.
__device__ __forceinline__ bool same_kmer(uint64_t alo, uint64_t ahi, uint64_t blo, uint64_t bhi) {
    return ahi == bhi && (alo >> 32) == (blo >> 32);
}

// WRITE == false: per-unit counts of kept k-mers / kept colour entries.  WRITE == true: emit them.
// A k-mer of a shared region carries one entry per colour (100 at C3), so the fold of a run is WARP-cooperative: every warp
// takes 32 consecutive entries, finds the run heads among them with one ballot, and all 32 lanes then sum / OR / copy the
// entries of each head's run together (a run may extend past the warp's 32 entries; its head's warp follows it).
template <int THREADS, bool WRITE>
__global__ void __launch_bounds__(THREADS)
k_color_fold(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi, const uint32_t *__restrict__ cf,
             const uint64_t *__restrict__ unit_off /* n_units+1, sorted layout */, uint32_t n_units, uint32_t min_mult,
             uint32_t *__restrict__ unit_keys, uint32_t *__restrict__ unit_cols,              // counts (out when !WRITE)
             const uint64_t *__restrict__ key_off, const uint64_t *__restrict__ col_off,      // scanned (in when WRITE)
             uint64_t *__restrict__ o_lo, uint64_t *__restrict__ o_hi, uint32_t *__restrict__ o_cf,
             uint64_t *__restrict__ o_coloff, uint32_t *__restrict__ o_colors) {
    __shared__ uint32_t s_scan[THREADS / 32 + 2];
    const uint32_t tid = threadIdx.x, lane = lane_id();
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint64_t b = unit_off[u], e = unit_off[u + 1];
        uint32_t run_keys = 0, run_cols = 0;  // running totals inside the unit (block-uniform)
        for (uint64_t base = b; base < e; base += THREADS) {
            const uint64_t i = base + tid;
            uint32_t keep = 0, ncol = 0, cfo = 0;
            uint64_t l = 0, h = 0;
            bool head = false;
            if (i < e) {
                l = lo[i]; h = hi[i];
                head = i == b || !same_kmer(l, h, lo[i - 1], hi[i - 1]);
            }
            // every head of this warp's 32 entries, in turn: all lanes fold its run
            uint32_t heads = __ballot_sync(0xffffffffu, head);
            while (heads) {
                const uint32_t hl = __ffs(heads) - 1u;
                heads &= heads - 1u;
                const uint64_t hi0 = __shfl_sync(0xffffffffu, h, hl), lo0 = __shfl_sync(0xffffffffu, l, hl);
                const uint64_t start = __shfl_sync(0xffffffffu, i, hl);
                uint64_t cnt = 0;
                uint32_t fl = 0, len = 0;
                for (uint64_t j0 = start;; j0 += 32) {
                    const uint64_t j = j0 + lane;
                    const bool in = j < e && same_kmer(lo0, hi0, lo[j], hi[j]);
                    if (in) { const uint32_t c = cf[j]; cnt += c & 0x3FFFFFFFu; fl |= c >> 30; ++len; }
                    if (__ballot_sync(0xffffffffu, in) != 0xffffffffu) break;   // the run ends inside these 32 entries
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                    fl |= __shfl_xor_sync(0xffffffffu, fl, o);
                    len += __shfl_xor_sync(0xffffffffu, len, o);
                }
                if (lane == hl) {
                    const uint64_t mult = cnt >> ((fl == 3u) ? 1 : 0);  // map_entry.rs:79-84 on the folded entry
                    if (mult >= min_mult) { keep = 1; ncol = len; cfo = (uint32_t)(mult > 0x3FFFFFFFull ? 0x3FFFFFFFull : mult) | (fl << 30); }
                }
            }
            uint32_t tk, tc;
            const uint32_t pk = block_exclusive_scan<THREADS>(keep, s_scan, &tk);
            const uint32_t pc = block_exclusive_scan<THREADS>(ncol, s_scan, &tc);
            if (WRITE) {
                const uint64_t ko = key_off[u] + run_keys + pk, co = col_off[u] + run_cols + pc;
                if (keep) {
                    o_lo[ko] = (l >> 32) | (h << 32);
                    o_hi[ko] = h >> 32;
                    o_cf[ko] = cfo;
                    o_coloff[ko] = co;
                }
                // the colour lists of this warp's kept heads, copied by all lanes
                uint32_t kept = __ballot_sync(0xffffffffu, keep != 0);
                while (kept) {
                    const uint32_t hl = __ffs(kept) - 1u;
                    kept &= kept - 1u;
                    const uint64_t src = __shfl_sync(0xffffffffu, i, hl), dst = __shfl_sync(0xffffffffu, co, hl);
                    const uint32_t n = __shfl_sync(0xffffffffu, ncol, hl);
                    for (uint32_t q = lane; q < n; q += 32) o_colors[dst + q] = (uint32_t)lo[src + q];
                }
            }
            run_keys += tk; run_cols += tc;
        }
        if (!WRITE && tid == 0) { unit_keys[u] = run_keys; unit_cols[u] = run_cols; }
    }
}

}  // namespace ggb
