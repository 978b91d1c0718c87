// unitigs.cuh -- partial unitigs of every merge unit, on the device, from the sorted k-mer tables.
//
// Replaces the consumer of the table in the reference's phase 2
//   crates/assembler_kmers_merge/src/unitigs_extender/hashmap.rs:65-86    get_kmers (kept, unused entries in table order)
//   crates/assembler_kmers_merge/src/unitigs_extender/hashmap.rs:162-297  try_extend_function (unitigs, not simplitigs)
//   crates/assembler_kmers_merge/src/unitigs_extender/hashmap.rs:442-601  compute_unitigs
//   crates/assembler_kmers_merge/src/final_executor.rs:103-245            output_sequence (where a partial unitig goes)
// The reference walks: from every unused k-mer it extends forward / backward while the current k-mer has EXACTLY ONE
// neighbour in the table on that side and that neighbour has exactly one neighbour back, stops behind a k-mer whose
// MapEntry flags say "contig break" (1 or 2: its outer (k-1)-mer lives in another unit) and marks k-mers used.  That is a
// traversal of a graph that does not depend on the order of the walk:
//   link(x, side) = y  iff  x has one neighbour y on `side`, y has one neighbour (x) on the side facing x, and `side` is
//                          not the ignored side of a flagged x  (flags 1 = begin / backward side, 2 = end / forward side)
// so every k-mer has <= 1 link per side, the components are paths and cycles, and the partial unitigs are exactly the
// paths (ends: no link = closed, ignored side of a flagged k-mer = open "continues in another unit") and the cycles
// (the reference opens a cycle at its smallest k-mer, the first unused one in table order).
//   k_unitig_links    one thread per table entry: both links by binary search in the unit's sorted key range
//   k_unitig_paths    one thread per path end: walk to the other end; the end with the smaller id emits the path
//                     (2-bit packed bases, 16 per word) and its routing record
//   k_unitig_cycles   what no path visited: every k-mer walks its cycle, the smallest one emits it
// Routing = output_sequence: both ends closed -> final output ("lonely"); else result bucket of the smaller-bucket end,
// seq-hash get_bucket (crates/hashes/src/base/cn_seqhash_base.rs:140-150, constants cn_seqhash.rs:22-26), should_rc,
// HASH_ENDING / OTHER_END flags and last_align exactly as final_executor.rs:170-236 computes them.
// 64-bit key path (k <= 31, odd k: the reference special-cases self-complementary k-mers of even k, final_executor.rs:249-264).
#pragma once
#include "device_utils.cuh"

namespace ggb {

constexpr uint32_t UT_NONE = 0xFFFFFFFFu;
enum { UT_BW = 0, UT_FW = 1 };   // sides of a canonical k-mer: predecessors / successors

struct UnitigRec {               // == ggcat_b200_unitig (include/ggcat_b200.h)
    unsigned long long word_offset;
    uint32_t len, unit;
    uint16_t bucket;
    uint8_t flags, last_align;
    uint32_t n_kmers;
};
enum { UTF_OPEN_BEGIN = 1, UTF_OPEN_END = 2, UTF_CIRCULAR = 4, UTF_SHOULD_RC = 8, UTF_HASH_ENDING = 16, UTF_OTHER_END = 32 };

struct UnitigTable {
    const uint64_t *keys; const uint32_t *cf; const uint64_t *unit_off;   // device table of a bucket range
    uint32_t n_units, first_unit, k;
    uint64_t n_entries;
};

__device__ __forceinline__ uint64_t ut_rc(uint64_t x, uint32_t k) { return revcomp64(x) >> (64 - 2 * k); }
__device__ __forceinline__ uint64_t ut_canon(uint64_t s, uint32_t k) { const uint64_t r = ut_rc(s, k); return s < r ? s : r; }
__device__ __forceinline__ uint64_t ut_succ(uint64_t s, uint32_t b, uint32_t k) { return (s >> 2) | ((uint64_t)b << (2 * (k - 1))); }   // manual_roll_forward
__device__ __forceinline__ uint64_t ut_pred(uint64_t s, uint32_t b, uint32_t k) { return ((s << 2) | b) & ((1ull << (2 * k)) - 1ull); }  // manual_roll_reverse

__device__ __forceinline__ long long ut_find(const uint64_t *__restrict__ keys, long long lo, long long hi, uint64_t key) {   // [lo, hi)
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const uint64_t v = keys[mid];
        if (v == key) return mid;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return -1;
}

// link of entry e on `side`: (target entry << 1 | side of the target we arrive at), or UT_NONE
__device__ __forceinline__ uint32_t ut_link(const UnitigTable &T, long long a, long long b, long long self, uint64_t key, uint32_t flags, int side) {
    if ((flags == 1u && side == UT_BW) || (flags == 2u && side == UT_FW)) return UT_NONE;   // ignored side: continues in another unit
    const uint32_t k = T.k;
    const uint64_t s = side == UT_FW ? key : ut_rc(key, k);      // walk forward along s
    int count = 0;
    long long tj = -1;
    uint64_t cand = 0;
    for (uint32_t q = 0; q < 4; q++) {
        const uint64_t c = ut_succ(s, q, k);
        const long long j = ut_find(T.keys, a, b, ut_canon(c, k));
        if (j >= 0) { count++; tj = j; cand = c; }
    }
    if (count != 1) return UT_NONE;
    int back = 0;
    for (uint32_t q = 0; q < 4; q++)
        if (ut_find(T.keys, a, b, ut_canon(ut_pred(cand, q, k), k)) >= 0) back++;
    if (back != 1) return UT_NONE;
    // a k-mer whose only neighbour is itself (its own reverse complement follows it): the reference finds it `used` and
    // stops without taking it twice (hashmap.rs:262-270)
    if (tj == self) return UT_NONE;
    // arriving along `cand`: on the backward side of the target if cand is its canonical form, else on its forward side
    const int arrive = cand == T.keys[tj] ? UT_BW : UT_FW;
    // symmetry: never link INTO the ignored side of a flagged k-mer (its partner on that side lives in another unit)
    const uint32_t tf = T.cf[tj] >> 30;
    if ((tf == 1u && arrive == UT_BW) || (tf == 2u && arrive == UT_FW)) return UT_NONE;
    return ((uint32_t)tj << 1) | (uint32_t)arrive;
}

__device__ __forceinline__ uint32_t ut_unit_of(const uint64_t *__restrict__ unit_off, uint32_t n_units, uint64_t e) {
    uint32_t lo = 0, hi = n_units;     // last u with unit_off[u] <= e
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (unit_off[mid] <= e) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256) k_unitig_links(UnitigTable T, uint32_t *__restrict__ links /* [n_entries][2] */, uint8_t *__restrict__ visited) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T.n_entries) return;
    const uint32_t u = ut_unit_of(T.unit_off, T.n_units, e);
    const long long a = (long long)T.unit_off[u], b = (long long)T.unit_off[u + 1];
    const uint64_t key = T.keys[e];
    const uint32_t flags = T.cf[e] >> 30;
    links[2 * e + UT_BW] = ut_link(T, a, b, (long long)e, key, flags, UT_BW);
    links[2 * e + UT_FW] = ut_link(T, a, b, (long long)e, key, flags, UT_FW);
    visited[e] = 0;
}

// seq-hash get_bucket(0, bits, canonical k-mer): cn_seqhash_base.rs:140-150, in the integer width the reference selects for
// k (crates/api/src/utils.rs:27-45: k <= 8 u16, <= 16 u32, <= 32 u64) with that width's constants (cn_seqhash.rs:1-27)
__device__ __forceinline__ uint32_t ut_result_bucket(uint64_t canon, uint32_t bits, uint32_t k) {
    uint64_t h;
    if (k <= 8) {
        const uint32_t x = ((uint32_t)canon * 0x0193u + 0x9dc5u) & 0xFFFFu;
        h = ((x >> 3) | (x << 13)) & 0xFFFFu;
    } else if (k <= 16) {
        const uint32_t x = (uint32_t)canon * 0x01000193u + 0x811c9dc5u;
        h = (x >> 3) | (x << 29);
    } else {
        h = canon * 0x00000100000001b3ull + 0xcbf29ce484222325ull;
        h = (h >> 3) | (h << 61);
    }
    return (uint32_t)(h & ((1ull << bits) - 1ull));
}

struct UnitigOut {
    UnitigRec *recs; uint32_t *bases;
    unsigned long long *counters;    // [0] unitigs, [1] words, [2] k-mers in unitigs
    uint64_t rec_cap, word_cap;
    uint32_t *overflow;
    uint32_t result_bits;
};

// Emits one partial unitig: the walk starts at entry e0, entered through side s0, and visits n nodes.
__device__ void ut_emit(const UnitigTable &T, const uint32_t *__restrict__ links, uint8_t *__restrict__ visited, const UnitigOut &O,
                        uint32_t e0, int s0, uint32_t n, bool open_begin, bool open_end, bool is_cycle) {
    const uint32_t k = T.k, len = k + n - 1, nw = (len + 15) >> 4;
    const unsigned long long ri = atomicAdd(&O.counters[0], 1ull), w0 = atomicAdd(&O.counters[1], (unsigned long long)nw);
    atomicAdd(&O.counters[2], (unsigned long long)n);
    if (ri >= O.rec_cap || w0 + nw > O.word_cap) { atomicOr(O.overflow, 1u); return; }
    uint32_t *dst = O.bases + w0;
    uint32_t cur = e0;
    int side = s0;
    uint64_t first_kmer = 0, last_kmer = 0;
    uint32_t word = 0, nb = 0, wi = 0;
    for (uint32_t i = 0; i < n; i++) {
        visited[cur] = 1;
        const uint64_t key = T.keys[cur];
        const uint64_t s = side == UT_BW ? key : ut_rc(key, k);   // entered on the backward side: canonical orientation
        if (i == 0) {
            first_kmer = s;
            for (uint32_t j = 0; j < k; j++) {
                word |= (uint32_t)((s >> (2 * j)) & 3ull) << (2 * nb);
                if (++nb == 16) { dst[wi++] = word; word = 0; nb = 0; }
            }
        } else {
            word |= (uint32_t)((s >> (2 * (k - 1))) & 3ull) << (2 * nb);
            if (++nb == 16) { dst[wi++] = word; word = 0; nb = 0; }
        }
        last_kmer = s;
        if (i + 1 < n) {
            const uint32_t l = links[2 * (uint64_t)cur + (side ^ 1)];
            cur = l >> 1; side = (int)(l & 1u);
        }
    }
    if (nb) dst[wi++] = word;
    // ---- routing (final_executor.rs:103-245)
    uint32_t flags = (open_begin ? UTF_OPEN_BEGIN : 0) | (open_end ? UTF_OPEN_END : 0);
    uint32_t bucket = 0xFFFFu, last_align = 0;
    if (!open_begin && !open_end) {
        // hashmap.rs:556-577: circular iff both ends closed and the first and last (k-1)-mers are the same canonical (k-1)-mer
        const uint64_t m1 = (1ull << (2 * (k - 1))) - 1ull;
        const uint64_t a = first_kmer & m1, b = last_kmer >> 2;
        const uint64_t ar = revcomp64(a) >> (64 - 2 * (k - 1)), br = revcomp64(b) >> (64 - 2 * (k - 1));
        if ((a < ar ? a : ar) == (b < br ? b : br) || is_cycle) flags |= UTF_CIRCULAR;
    } else {
        const uint64_t fr = ut_rc(first_kmer, k), lr = ut_rc(last_kmer, k);
        const bool first_fw = first_kmer < fr, last_fw = last_kmer < lr;
        const uint32_t left_bucket = open_begin ? ut_result_bucket(first_fw ? first_kmer : fr, O.result_bits, k) : 0xFFFFu;
        const uint32_t right_bucket = open_end ? ut_result_bucket(last_fw ? last_kmer : lr, O.result_bits, k) : 0xFFFFu;
        const bool left_rc = open_begin ? !first_fw : false, right_rc = open_end ? !last_fw : true;
        const bool hash_beginning = left_bucket <= right_bucket;
        const bool should_rc = hash_beginning ? left_rc : right_rc;
        bucket = hash_beginning ? left_bucket : right_bucket;
        last_align = (hash_beginning != should_rc) ? 0u : ((len - k) & 3u);
        if (should_rc) flags |= UTF_SHOULD_RC;
        if ((!hash_beginning) != should_rc) flags |= UTF_HASH_ENDING;
        if (open_begin && open_end) flags |= UTF_OTHER_END;
    }
    UnitigRec r;
    r.word_offset = w0; r.len = len; r.unit = T.first_unit + ut_unit_of(T.unit_off, T.n_units, e0);
    r.bucket = (uint16_t)bucket; r.flags = (uint8_t)flags; r.last_align = (uint8_t)last_align; r.n_kmers = n;
    O.recs[ri] = r;
}

// One thread per (entry, side) that is a path end (no link on that side): walk to the other end.  Both ends of a path do
// this; the path is emitted by the end whose walk passes the path's SMALLEST entry in canonical orientation -- the
// orientation the reference gives the unitig (it starts at the first unused k-mer in table order, read forward).
__global__ void __launch_bounds__(256) k_unitig_paths(UnitigTable T, const uint32_t *__restrict__ links, uint8_t *__restrict__ visited, UnitigOut O) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * T.n_entries) return;
    const uint32_t e0 = (uint32_t)(t >> 1);
    const int s0 = (int)(t & 1);
    if (links[t] != UT_NONE) return;                     // not an end
    uint32_t cur = e0, n = 1, min_e = e0;
    int side = s0, min_side = s0;
    while (true) {
        const uint32_t l = links[2 * (uint64_t)cur + (side ^ 1)];
        if (l == UT_NONE) break;
        cur = l >> 1; side = (int)(l & 1u); ++n;
        if (cur < min_e) { min_e = cur; min_side = side; }
        if (n > T.n_entries) { atomicOr(O.overflow, 2u); return; }   // cannot happen: links are symmetric
    }
    if (min_side != UT_BW) return;                       // the walk from the other end has the reference's orientation
    const uint32_t f0 = T.cf[e0] >> 30, f1 = T.cf[cur] >> 30;
    const bool open_begin = (f0 == 1u && s0 == UT_BW) || (f0 == 2u && s0 == UT_FW);
    const bool open_end = (f1 == 1u && (side ^ 1) == UT_BW) || (f1 == 2u && (side ^ 1) == UT_FW);
    ut_emit(T, links, visited, O, e0, s0, n, open_begin, open_end, false);
}

// Entries no path visited lie on cycles: the smallest entry of a cycle emits it, starting with its canonical orientation.
__global__ void __launch_bounds__(256) k_unitig_cycles(UnitigTable T, const uint32_t *__restrict__ links, uint8_t *__restrict__ visited, UnitigOut O) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T.n_entries) return;
    if (visited[e] || links[2 * e] == UT_NONE || links[2 * e + 1] == UT_NONE) return;
    uint32_t cur = (uint32_t)e, n = 1;
    int side = UT_BW;
    while (true) {
        const uint32_t l = links[2 * (uint64_t)cur + (side ^ 1)];
        if (l == UT_NONE) return;                        // not a cycle after all (belongs to a path): unreachable
        cur = l >> 1; side = (int)(l & 1u);
        if (cur == (uint32_t)e) break;
        if (cur < (uint32_t)e) return;                   // a smaller entry owns this cycle
        if (++n > T.n_entries) { atomicOr(O.overflow, 2u); return; }
    }
    ut_emit(T, links, visited, O, (uint32_t)e, UT_BW, n, false, false, true);
}


// ------------------------------------------------------------------------------------------------
// Joining partial unitigs into maximal unitigs.
//
// Replaces crates/assembler_pipeline/src/extend_unitigs.rs:348- ("phase: unitigs joining"): the reference reads the
// "result" buckets one after the other, keeps the partial unitigs of a bucket in a hash map keyed by the hash of their
// open end, glues two partial unitigs that end in the same k-mer (stored once in each of the two units, with complementary
// flags) and re-routes the glued piece by its other open end until nothing is open.  The fixed point of that is a graph
// problem again: an open end has exactly one partner (the other partial unitig that ends in the same canonical k-mer), so
// partial unitigs form chains and cycles, and every chain / cycle is one maximal unitig.
//   k_join_ends    one thread per open end: canonical end k-mer -> CAS hash table; the second arrival pairs the two ends
//   k_join_chains  one thread per chain end (closed end, or an open end nobody answered): walk through the partners; the
//                  smaller end id emits -- first piece whole, every further piece without the k bases it shares
//   k_join_cycles  pieces no chain visited: the smallest piece of a cycle emits it (k + L - 1 bases for L k-mers)
// Orientation of a joined unitig is the walk's (the reference's depends on thread timing: SURVEY App. C); consumers
// compare canonical k-mer sets.
struct JoinTable {
    const UnitigRec *recs; const uint32_t *bases;
    uint64_t n;                       // partial unitigs
    uint32_t k;
    unsigned long long *ht_keys;      // open-addressing table of canonical end k-mers (EMPTY = ~0)
    uint32_t *ht_first;               // first end that arrived at the slot
    uint64_t ht_slots;
    uint32_t *partner;                // [2n]: end id (unitig << 1 | end) of the partner end, UT_NONE = none
    uint8_t *visited;                 // [n]
};

__device__ __forceinline__ uint64_t join_end_kmer(const JoinTable &J, uint64_t u, int end) {
    const UnitigRec r = J.recs[u];
    const uint32_t off = end ? r.len - J.k : 0u;
    return extract64(J.bases + r.word_offset, 2ull * off) & ((1ull << (2 * J.k)) - 1ull);
}

__global__ void __launch_bounds__(256) k_join_init(JoinTable J) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < J.ht_slots) { J.ht_keys[i] = ~0ull; J.ht_first[i] = UT_NONE; }
    if (i < 2 * J.n) J.partner[i] = UT_NONE;
    if (i < J.n) J.visited[i] = 0;
}

__global__ void __launch_bounds__(256) k_join_ends(JoinTable J, uint32_t *__restrict__ overflow) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * J.n) return;
    const uint64_t u = t >> 1;
    const int end = (int)(t & 1);
    if (!(J.recs[u].flags & (end ? UTF_OPEN_END : UTF_OPEN_BEGIN))) return;
    const uint64_t key = ut_canon(join_end_kmer(J, u, end), J.k);
    uint64_t slot = ((key * 0x9E3779B97F4A7C15ull) >> 20) % J.ht_slots;
    while (true) {
        const unsigned long long old = atomicCAS(&J.ht_keys[slot], ~0ull, (unsigned long long)key);
        if (old == ~0ull || old == key) break;
        slot = slot + 1 == J.ht_slots ? 0 : slot + 1;
    }
    const uint32_t first = atomicCAS(&J.ht_first[slot], UT_NONE, (uint32_t)t);
    if (first == UT_NONE) return;                                // the partner has not arrived yet
    if (atomicCAS(&J.partner[first], UT_NONE, (uint32_t)t) != UT_NONE) { atomicOr(overflow, 4u); return; }   // a third end with this k-mer
    J.partner[t] = first;
}

// Emits the chain / cycle that starts at piece u0, entered through end e0, n_pieces long.
__device__ void join_emit(const JoinTable &J, const UnitigOut &O, uint64_t u0, int e0, uint32_t n_pieces, uint64_t total_len, bool cycle) {
    const uint32_t k = J.k;
    const uint64_t len = cycle ? total_len - 1 : total_len;
    const uint64_t nw = (len + 15) >> 4;
    const unsigned long long ri = atomicAdd(&O.counters[0], 1ull), w0 = atomicAdd(&O.counters[1], (unsigned long long)nw);
    atomicAdd(&O.counters[2], (unsigned long long)(len - k + 1));
    if (ri >= O.rec_cap || w0 + nw > O.word_cap || len >= (1ull << 32)) { atomicOr(O.overflow, 1u); return; }
    uint32_t *dst = O.bases + w0;
    uint32_t word = 0, nb = 0;
    uint64_t wi = 0, written = 0, cur = u0;
    int enter = e0;
    for (uint32_t p = 0; p < n_pieces; p++) {
        J.visited[cur] = 1;
        const UnitigRec r = J.recs[cur];
        const uint32_t *src = J.bases + r.word_offset;
        for (uint32_t q = p ? k : 0u; q < r.len && written < len; q++) {
            // entered at the beginning: stored orientation; entered at the end: reverse complement
            const uint32_t pos = enter == 0 ? q : r.len - 1 - q;
            uint32_t c = (src[pos >> 4] >> (2 * (pos & 15u))) & 3u;
            if (enter) c ^= 2u;
            word |= c << (2 * nb);
            if (++nb == 16) { dst[wi++] = word; word = 0; nb = 0; }
            ++written;
        }
        if (p + 1 < n_pieces) {
            const uint32_t nx = J.partner[2 * cur + (uint64_t)(enter ^ 1)];
            cur = nx >> 1; enter = (int)(nx & 1u);
        }
    }
    if (nb) dst[wi++] = word;
    UnitigRec o;
    o.word_offset = w0; o.len = (uint32_t)len; o.unit = n_pieces; o.bucket = 0xFFFFu;
    // a lonely partial unitig that was a cycle inside its unit keeps its circular flag (it already has k + L - 1 bases)
    o.flags = (uint8_t)((cycle || (n_pieces == 1 && (J.recs[u0].flags & UTF_CIRCULAR))) ? UTF_CIRCULAR : 0);
    o.last_align = 0; o.n_kmers = (uint32_t)(len - k + 1);
    O.recs[ri] = o;
}

__global__ void __launch_bounds__(256) k_join_chains(JoinTable J, UnitigOut O) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * J.n) return;
    if (J.partner[t] != UT_NONE) return;                         // not a chain end
    const uint64_t u0 = t >> 1;
    const int e0 = (int)(t & 1);
    uint64_t cur = u0, total = J.recs[u0].len;
    int enter = e0;
    uint32_t n = 1;
    while (true) {
        const uint32_t nx = J.partner[2 * cur + (uint64_t)(enter ^ 1)];
        if (nx == UT_NONE) break;
        cur = nx >> 1; enter = (int)(nx & 1u);
        total += J.recs[cur].len - J.k;
        if (++n > J.n) { atomicOr(O.overflow, 2u); return; }
    }
    const uint64_t other = 2 * cur + (uint64_t)(enter ^ 1);
    if (other < t) return;                                        // the other end of the chain emits
    join_emit(J, O, u0, e0, n, total, false);
}

__global__ void __launch_bounds__(256) k_join_cycles(JoinTable J, UnitigOut O) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= J.n) return;
    if (J.visited[u] || J.partner[2 * u] == UT_NONE || J.partner[2 * u + 1] == UT_NONE) return;
    uint64_t cur = u, total = J.recs[u].len;
    int enter = 0;
    uint32_t n = 1;
    while (true) {
        const uint32_t nx = J.partner[2 * cur + (uint64_t)(enter ^ 1)];
        if (nx == UT_NONE) return;                               // a chain after all: unreachable
        cur = nx >> 1; enter = (int)(nx & 1u);
        if (cur == u) break;
        if (cur < u) return;                                     // a smaller piece owns this cycle
        total += J.recs[cur].len - J.k;
        if (++n > J.n) { atomicOr(O.overflow, 2u); return; }
    }
    // closing the cycle: the last piece's last k-mer is the first piece's first k-mer.  total = sum(len) - k (n - 1) has
    // L + 1 k-mers for the cycle's L; join_emit drops the last base
    join_emit(J, O, u, 0, n, total, true);
}

}  // namespace ggb
