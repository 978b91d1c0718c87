// bucketing.cuh -- phase 1 (minimizer bucketing) kernels.
//
// Replaces the per-sequence hot loop of the reference
//   crates/minimizer_bucketing/src/lib.rs:310-376  (normalise / N-split / process_sequence / push)
//   crates/assembler_minimizer_bucketing/src/lib.rs:171-270  (process_sequence)
//   crates/hashes/src/cn_nthash.rs:21-58, crates/hashes/src/rolling/batch_minqueue.rs:37-188
// by a position-parallel formulation (tests/model.py is the executable spec):
// every (k-1)-mer window j of the concatenated batch gets  M_j = min over its w = k-m m-mer values
// with the duplicate rule folded into the value ( comb(a,b) = a==b ? a&~1 : min(a,b) ), and a
// super-k-mer starts at j iff j is the first window of its segment, or M_j != M_{j-1}, or the
// minimum is unique but a different m-mer instance ( v[j-1] == M_j ).
//
// Data flow per batch (all device-resident):
//   k_pack      ASCII -> 2-bit words (16 bases / u32) + `bad` bitmap (non-ACGT)      [1 B/base read]
//   k_mark      read starts -> `brk` bitmap
//   k_windows   tile of 1024 windows: hashes, window minima, split/segment-end entries (compacted)
//   k_scan      exclusive scan of per-tile super-k-mer counts
//   k_emit      entry -> super-k-mer descriptor (start,len,unit,flags,rc,minimizer_pos), unit histogram
//   k_scan      exclusive scan of per-unit counts / payload words
//   k_scatter   stored-orientation payload + descriptor into the unit-sorted bucket chunk
#pragma once
#include "device_utils.cuh"

namespace ggb {

struct DevParams {
    uint32_t k, m, w;  // w = k - m : m-mers per (k-1)-mer window
    uint32_t b1, b2;
    uint32_t forward_only;
    uint32_t colors;
    uint32_t n_units;  // ((1<<b1)+1) << b2
    uint32_t hash_type;  // ggcat_b200_hash_type (k-mer identity hash of phase 2)
};

#ifndef GGB_WIN_T
#define GGB_WIN_T 1024
#endif
#ifndef GGB_WIN_TPC
#define GGB_WIN_TPC 1
#endif
constexpr int WIN_T = GGB_WIN_T;            // windows per tile (1024 or 2048)
constexpr int WIN_THREADS = WIN_T / 8;      // 8 windows per thread
constexpr int WIN_TPC = GGB_WIN_TPC;        // tiles per CTA: the bulk copies of tile t+1 are in flight while tile t is processed
constexpr int WIN_NB = WIN_TPC > 1 ? 2 : 1; // landing buffers
#ifndef GGB_WIN_MINB
#define GGB_WIN_MINB 9
#endif
constexpr int WIN_MINB = GGB_WIN_MINB;      // resident CTAs per SM the register allocation must allow
constexpr int WIN_GROUP = 64;     // tiles per group of the two-level super-k-mer prefix (k_windows -> group scan -> k_emit)
constexpr int WIN_WMAX = 128;     // max supported k - m (k <= 128); k_windows<64> serves k - m <= 64, k <= 66 with smaller buffers
#define PADX(x) ((x) + ((x) >> 3))

// entry bit layout (u64)
constexpr int ENT_POS_BITS = WIN_T <= 1024 ? 11 : 12;  // position of the window inside its tile (0 .. WIN_T)
constexpr uint64_t ENT_S = 1ull << ENT_POS_BITS, ENT_E = ENT_S << 1, ENT_FIRST = ENT_S << 2, ENT_RC = ENT_S << 3,
                   ENT_DUP = ENT_S << 4;
constexpr int ENT_SECOND_SHIFT = ENT_POS_BITS + 5, ENT_BUCKET_SHIFT = ENT_SECOND_SHIFT + 8, ENT_ARG_SHIFT = ENT_BUCKET_SHIFT + 14,
              ENT_SRANK_SHIFT = ENT_ARG_SHIFT + 8;   // then 12 bits of super-k-mer rank

// descriptor meta word: minimizer_pos(16) | flags(2)<<16 | rc<<18 | second_bucket(8)<<19
__host__ __device__ __forceinline__ uint32_t make_meta(uint32_t mpos, uint32_t flags, uint32_t rc, uint32_t second) {
    return (mpos & 0xFFFFu) | (flags << 16) | (rc << 18) | (second << 19);
}

// ------------------------------------------------------------------------------------------------
// k_pack: crates/io/src/sequences_reader.rs:26-37 (normalisation: anything but ACGTacgt is 'N') +
// crates/utils/src/lib.rs:44-46 (code = (c>>1)&3) + crates/io/src/compressed_read.rs:610-618 layout.
// One thread per 32 bases -> two packed words and one `bad` word.  Positions >= n are bad.
__global__ void __launch_bounds__(256) k_pack(const uint8_t *__restrict__ ascii, uint64_t n, uint32_t *__restrict__ pk,
                                              uint32_t *__restrict__ bad, uint64_t n_groups, int aligned16) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint64_t base = g * 32;
    uint32_t x[8];
    if (base + 32 <= n && aligned16) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(ascii + base));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(ascii + base + 16));
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint64_t i = base + q * 4 + b;
                const uint32_t c = i < n ? ascii[i] : 0u;
                v |= c << (8 * b);
            }
            x[q] = v;
        }
    }
    uint32_t w0 = 0, w1 = 0, bd = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t c = (x[q] >> 1) & 0x03030303u;
        const uint32_t codes = (c & 3u) | ((c >> 6) & 0xCu) | ((c >> 12) & 0x30u) | ((c >> 18) & 0xC0u);
        const uint32_t u = x[q] & 0xDFDFDFDFu;
        const uint32_t valid = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) |
                               __vcmpeq4(u, 0x54545454u);
        const uint32_t inv = ~valid;
        const uint32_t bits = (inv & 1u) | ((inv >> 7) & 2u) | ((inv >> 14) & 4u) | ((inv >> 21) & 8u);
        const uint32_t okm = ((valid & 1u) * 3u) | (((valid >> 8) & 1u) * 0xCu) | (((valid >> 16) & 1u) * 0x30u) |
                             (((valid >> 24) & 1u) * 0xC0u);
        const uint32_t cc = codes & okm;  // invalid bases pack as 0 (never read by a valid window)
        if (q < 4) w0 |= cc << (8 * q); else w1 |= cc << (8 * (q - 4));
        bd |= bits << (4 * q);
    }
    pk[2 * g] = w0;
    pk[2 * g + 1] = w1;
    bad[g] = bd;
}

// k_repack: 2-bit packed input (the reference's CompressedRead layout, crates/io/src/compressed_read.rs:610-618: base i at
// bits 2(i % 4) of byte i / 4) -> this batch's pk / bad.  The batch starts `shift` bases into the first source word; one
// thread per 32 bases.  Packed input has no code for N (the host splits at N before packing, as the reference does), so
// only the positions >= n are bad.
__global__ void __launch_bounds__(256) k_repack(const uint32_t *__restrict__ src, uint32_t shift, uint64_t n, uint32_t *__restrict__ pk,
                                                uint32_t *__restrict__ bad, uint64_t n_groups, uint64_t src_words) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint64_t base = g * 32;
    uint32_t w0 = 0, w1 = 0, bd = 0xFFFFFFFFu;
    if (base < n) {
        const uint64_t bit = 2ull * (base + shift), w = bit >> 5;
        const uint32_t sh = (uint32_t)bit & 31u;
        const uint32_t a = w < src_words ? src[w] : 0u, b = w + 1 < src_words ? src[w + 1] : 0u, c = w + 2 < src_words ? src[w + 2] : 0u;
        w0 = __funnelshift_r(a, b, sh); w1 = __funnelshift_r(b, c, sh);
        const uint64_t left = n - base;          // valid bases in this group
        bd = left >= 32 ? 0u : (0xFFFFFFFFu << (uint32_t)left);
        if (left < 16) { w0 &= (1u << (2 * (uint32_t)left)) - 1u; w1 = 0; }
        else if (left < 32) w1 &= left == 16 ? 0u : ((1u << (2 * ((uint32_t)left - 16))) - 1u);
    }
    pk[2 * g] = w0; pk[2 * g + 1] = w1; bad[g] = bd;
}

// k_mark: a record boundary is a segment boundary (each input record is processed independently,
// crates/minimizer_bucketing/src/lib.rs:310-320).
__global__ void k_mark(const uint64_t *__restrict__ offsets, uint64_t n_reads, uint64_t off0, uint64_t n,
                       uint32_t *__restrict__ brk) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint64_t off = offsets[r] - off0;
    if (off < n) atomicOr(&brk[off >> 5], 1u << (off & 31));
}

// comb(): the window-minimum semigroup with the duplicate flag in bit 0
// (crates/hashes/src/rolling/batch_minqueue.rs:63-70,78-84,97-113: equal values clear the unique bit).
__device__ __forceinline__ uint64_t comb(uint64_t a, uint64_t b) { return a == b ? (a & ~1ull) : (a < b ? a : b); }

// comb() that also carries the position of the (unique) minimum: the argmin of a window comes out of the same reduction
// that computes its minimum, so no thread has to search the window afterwards (ties clear the unique bit; the position
// of a non-unique minimum is never used: such windows go to the duplicates bucket).
__device__ __forceinline__ void combp(uint64_t &av, uint32_t &ap, uint64_t bv, uint32_t bp) {
    if (av == bv) av &= ~1ull;
    else if (bv < av) { av = bv; ap = bp; }
}

// `cnt` (<= 64) bits of a shared-memory bitmap starting at bit `s`, as a u64.
__device__ __forceinline__ uint64_t bits64(const uint32_t *bm, uint32_t s, uint32_t cnt) {
    const uint64_t v = extract64(bm, s);
    return cnt >= 64 ? v : (v & ((1ull << cnt) - 1ull));
}

// ------------------------------------------------------------------------------------------------
// k_windows: one CTA per tile of WIN_T windows.
//   A  stage packed bases + bitmaps in shared memory, build the per-position rotation tables
//   B  m-mer hashes by rolling (cn_nthash.rs:43-57), IPT consecutive items per thread; window validity
//   C  window minima, van Herk / Gil-Werman: per block of w items a suffix scan and a prefix scan with
//      comb() -- exactly the two arrays the reference's BatchMinQueue keeps (batch_minqueue.rs:58-113)
//   D  M_x = comb(suffix[x], prefix[x+w-1])
//   E  split-start / segment-end flags, F in-order compaction (one block scan) + bucket / orientation / minimizer offset
//      of every super-k-mer start from the (minimum, argmin) pair its thread holds in registers, G copy-out
template <int WMAX>
__global__ void __launch_bounds__(WIN_THREADS, WIN_MINB)
k_windows(const uint32_t *__restrict__ pk, const uint32_t *__restrict__ bad, const uint32_t *__restrict__ brk,
          uint32_t n /* bases in batch */, DevParams P, uint64_t *__restrict__ ent, uint32_t *__restrict__ tile_cnt,
          uint32_t *__restrict__ tile_scnt, uint32_t *__restrict__ seg_count, uint32_t n_tiles,
          uint32_t *__restrict__ group_scnt /* super-k-mers per group of WIN_GROUP tiles (zeroed by the host) */) {
    // TMA landing buffers (16-byte aligned; the tile's first word sits at offset (W0 & 3) / (BW0 & 3) because the
    // bulk copy starts at the 16-byte boundary below it)
    constexpr int WIN_NI = WIN_T + WMAX;  // m-mer items per tile (upper bound)
    constexpr int PKW = (((WIN_T + 2 * WMAX + 64) / 16 + 12) + 3) & ~3, BMW = (((WIN_T + 2 * WMAX + 64) / 32 + 12) + 3) & ~3;
    __shared__ __align__(16) uint32_t s_pk_all[WIN_NB][PKW];
    __shared__ __align__(16) uint32_t s_bad_all[WIN_NB][BMW];
    __shared__ __align__(16) uint32_t s_cmb_all[WIN_NB][BMW];   // record starts, then bad | record-start
    __shared__ __align__(8) uint64_t s_bar[WIN_NB];
    // m-mer values and window minima: element x lives at PADX(x) = x + x/8, so that the 8-windows-per-thread phase
    // (lane stride 8 elements) touches every 8-byte bank pair twice per warp instead of sixteen times
    __shared__ uint64_t s_v0[PADX(WIN_NI + 16) + 1];
    __shared__ uint8_t s_fwd[WIN_NI + 1];
    __shared__ uint64_t s_ent[WIN_T];
    __shared__ uint64_t s_TF[32][4], s_TR[32][4];  // rotl(h(c), m-1-i), rotl(r(c), i)
    __shared__ uint64_t s_T1[16], s_T2[16];        // roll tables indexed by (leaving base << 2 | entering base)
    __shared__ uint32_t s_scan[WIN_THREADS / 32 + 2];
    __shared__ uint32_t s_nfirst;                  // segments (N-free stretches of >= k bases) that start in this CTA's tiles

    const uint32_t tid = threadIdx.x;
    const uint32_t tile_first = blockIdx.x * WIN_TPC;
    const uint32_t k = P.k, m = P.m, w = P.w;
    const uint32_t n_items = WIN_T + w;                   // m-mer items x in [0, n_items)

    // ---- A: one elected thread issues three TMA bulk copies per tile (packed bases, bad bitmap, record-start bitmap); the
    //         copies of the first two tiles land while the other threads build the hash tables
    auto issue = [&](uint32_t tl) {
        const uint32_t tile = tile_first + tl;
        if (tile >= n_tiles) return;
        const int64_t bf = (int64_t)tile * WIN_T - 1 < 0 ? 0 : (int64_t)tile * WIN_T - 1;
        const uint32_t W0 = (uint32_t)(bf >> 4), BW0 = (uint32_t)(bf >> 5);
        const uint32_t lastb = (uint32_t)((int64_t)tile * WIN_T + WIN_T + k + 1);
        const uint32_t n_pkw = ((lastb + 15) >> 4) - W0 + 2, n_bmw = ((lastb + 31) >> 5) - BW0 + 4;
        const uint32_t pkw = (n_pkw + (W0 & 3u) + 3u) & ~3u, bmw = (n_bmw + (BW0 & 3u) + 3u) & ~3u;
        const uint32_t b = tl & (WIN_NB - 1);
        mbar_expect_tx(&s_bar[b], (pkw + 2 * bmw) * 4);
        tma_load_1d(s_pk_all[b], pk + (W0 & ~3u), pkw * 4, &s_bar[b]);
        tma_load_1d(s_bad_all[b], bad + (BW0 & ~3u), bmw * 4, &s_bar[b]);
        tma_load_1d(s_cmb_all[b], brk + (BW0 & ~3u), bmw * 4, &s_bar[b]);
    };
    if (tid == 0) {
        s_nfirst = 0;
        for (int b = 0; b < WIN_NB; b++) mbar_init(&s_bar[b], 1);
        issue(0);
        if (WIN_TPC > 1) issue(1);
    }
    if (tid < 16) {
        // cn_nthash.rs:43-57 roll_hash with both table terms folded:
        //   fw' = rotl(fw,1) ^ rotl(h(out), m) ^ h(in)          rc' = rotr(rc ^ r(out), 1) ^ rotl(r(in), m-1)
        const uint32_t o = tid >> 2, i = tid & 3;
        s_T1[tid] = rotl64(nt_h(o), m) ^ nt_h(i);
        s_T2[tid] = rotl64(nt_r(o), 63) ^ rotl64(nt_r(i), m - 1);
    }
    if (tid < 4 * m) {
        const uint32_t i = tid >> 2, c = tid & 3;
        s_TF[i][c] = rotl64(nt_h(c), m - 1 - i);
        s_TR[i][c] = rotl64(nt_r(c), i);
    }
    __syncthreads();          // barriers initialised (and tables written) before anyone polls them
#pragma unroll 1
    for (uint32_t tl = 0; tl < (uint32_t)WIN_TPC; tl++) {
    const uint32_t tile = tile_first + tl;
    if (tile >= n_tiles) break;
    const uint32_t lb_ = tl & (WIN_NB - 1);
    const int64_t j0 = (int64_t)tile * WIN_T;  // first window of the tile (== its base index)
    // x-coordinates: window / m-mer item x  <->  global position j0 - 1 + x
    const int64_t gfirst = j0 - 1;
    const int64_t bfirst = gfirst < 0 ? 0 : gfirst;       // first base the tile may touch
    const uint32_t W0 = (uint32_t)(bfirst >> 4);          // first packed word
    const uint32_t BW0 = (uint32_t)(bfirst >> 5);         // first bitmap word
    const uint32_t last_base = (uint32_t)(j0 + WIN_T + k + 1);  // exclusive upper bound of bases touched
    const uint32_t n_bmw = ((last_base + 31) >> 5) - BW0 + 4;
    uint32_t *s_pk_raw = s_pk_all[lb_], *s_bad_raw = s_bad_all[lb_], *s_cmb_raw = s_cmb_all[lb_];
    uint32_t *s_pk = s_pk_raw + (W0 & 3u), *s_bad = s_bad_raw + (BW0 & 3u), *s_cmb = s_cmb_raw + (BW0 & 3u);
    const uint32_t bm_words = (n_bmw + (BW0 & 3u) + 3u) & ~3u;
    mbar_wait(&s_bar[lb_], (tl >> 1) & 1u);
    for (uint32_t i = tid; i < bm_words; i += WIN_THREADS) s_cmb_raw[i] |= s_bad_raw[i];
    __syncthreads();

    // ---- B: m-mer hashes by rolling, IPT consecutive items per thread, bases kept in two 64-bit shift registers
    //         (leaving / entering side)
    {
        const uint32_t IPT = (n_items + WIN_THREADS - 1) / WIN_THREADS;   // <= 9
        const uint32_t x0 = tid * IPT;
        const uint32_t x1 = min(n_items, x0 + IPT);
        if (x0 < x1) {
            int64_t g = gfirst + x0;  // global m-mer position
            uint32_t xs = x0;
            if (g < 0) { s_v0[PADX(xs)] = ~0ull; s_fwd[xs] = 0; ++g; ++xs; }
            if (xs < x1) {
                const uint32_t lb = (uint32_t)(g - ((int64_t)W0 << 4));  // local base index into s_pk
                uint64_t outw = extract64(s_pk, 2ull * lb), inw = extract64(s_pk, 2ull * (lb + m));
                uint64_t fw = 0, rc = 0;
                {
                    uint64_t t = outw;
                    for (uint32_t i = 0; i < m; i++) {
                        const uint32_t c = (uint32_t)t & 3u;
                        t >>= 2;
                        fw ^= s_TF[i][c];
                        rc ^= s_TR[i][c];
                    }
                }
                for (uint32_t x = xs;; ++x) {
                    const bool isf = fw < rc;
                    const uint64_t mn = isf ? fw : rc;
                    s_v0[PADX(x)] = (mn << 1) | (uint64_t)(fw != rc);  // to_unextendable | !is_rc_symmetric
                    s_fwd[x] = isf;
                    if (x + 1 >= x1) break;
                    const uint32_t idx = (((uint32_t)outw & 3u) << 2) | ((uint32_t)inw & 3u);
                    outw >>= 2; inw >>= 2;
                    fw = ((fw << 1) | (fw >> 63)) ^ s_T1[idx];
                    rc = ((rc >> 1) | (rc << 63)) ^ s_T2[idx];
                }
            }
        }
    }
    __syncthreads();

    // ---- C: 8 consecutive windows per thread (x0 .. x0+7, x0 = 1 + 8 tid), everything in registers.
    //   window x = items [x, x+w-1].  The items [x0+7, x0+w-1] are common to the 8 windows; each window adds a suffix
    //   of the 7 items before and a prefix of the 7 items after (a van Herk block of 8 held in registers).  The window
    //   before the first (x0-1) is recomputed here too, so E needs nothing from other threads.
    //   comb() is the reference's min-with-duplicate-flag semigroup (batch_minqueue.rs:63-113).
    constexpr int WPT = 8;
    static_assert(WPT * WIN_THREADS == WIN_T, "8 windows per thread cover the tile");
    const uint32_t x0 = 1 + tid * WPT;
    uint64_t Mreg[WPT], Mprev;
    uint32_t Preg[WPT];    // item index (tile x coordinate) of the window's minimum
    if (w >= (uint32_t)WPT) {
        uint64_t common = s_v0[PADX(x0 + WPT - 1)], common_short = common;   // [x0+7, x0+w-1] and the same without the last item
        uint32_t common_p = x0 + WPT - 1;
        for (uint32_t q = x0 + WPT; q < x0 + w; ++q) { common_short = common; combp(common, common_p, s_v0[PADX(q)], q); }
        uint64_t L[WPT - 1];
        uint32_t Lp[WPT - 1];
        L[WPT - 2] = s_v0[PADX(x0 + WPT - 2)]; Lp[WPT - 2] = x0 + WPT - 2;
#pragma unroll
        for (int i = WPT - 3; i >= 0; --i) { L[i] = s_v0[PADX(x0 + i)]; Lp[i] = x0 + i; combp(L[i], Lp[i], L[i + 1], Lp[i + 1]); }
        // window x0-1 = item x0-1 + items [x0, x0+6] + items [x0+7, x0+w-2]
        Mprev = comb(s_v0[PADX(x0 - 1)], L[0]);
        if (w > (uint32_t)WPT) Mprev = comb(Mprev, common_short);
        uint64_t R = 0;
        uint32_t Rp = 0;
#pragma unroll
        for (int i = 0; i < WPT; i++) {
            uint64_t mi = common;
            uint32_t mp = common_p;
            if (i < WPT - 1) { mi = L[i]; mp = Lp[i]; combp(mi, mp, common, common_p); }
            if (i > 0) {
                const uint64_t nv = s_v0[PADX(x0 + w + i - 1)];
                if (i == 1) { R = nv; Rp = x0 + w; } else combp(R, Rp, nv, x0 + w + i - 1);
                combp(mi, mp, R, Rp);
            }
            Mreg[i] = mi; Preg[i] = mp;
        }
    } else {
#pragma unroll
        for (int i = 0; i < WPT; i++) {
            uint64_t mi = s_v0[PADX(x0 + i)];
            uint32_t mp = x0 + i;
            for (uint32_t q = 1; q < w; ++q) combp(mi, mp, s_v0[PADX(x0 + i + q)], x0 + i + q);
            Mreg[i] = mi; Preg[i] = mp;
        }
        Mprev = s_v0[PADX(x0 - 1)];
        for (uint32_t q = 1; q < w; ++q) Mprev = comb(Mprev, s_v0[PADX(x0 - 1 + q)]);
    }
    // window validity for x0-1 .. x0+8: inside one N-free segment of one record (sequences_splitter.rs:15-40):
    //   no bad base in [j, j+k-1) and no record start in (j, j+k-1)  <=>  base j is good and
    //   (bad | record-start) has no bit in (j, j+k-1)
    uint32_t ok10 = 0;
    {
        const int64_t jf = gfirst + x0 - 1;                               // position of window x0-1 (>= -1)
        const uint32_t rel = (uint32_t)(jf + 1 - ((int64_t)BW0 << 5));    // bit index of position jf + 1
        const uint64_t clo = extract64(s_cmb, rel), chi = extract64(s_cmb, rel + 64);
        const uint64_t kmask = (k - 2 >= 64) ? ~0ull : ((1ull << (k - 2)) - 1ull);
        const uint32_t badn = extract32(s_bad, rel);                      // bad bits of positions jf+1 ..
        const uint32_t bad0 = jf >= 0 ? ((s_bad[(rel - 1) >> 5] >> ((rel - 1) & 31u)) & 1u) : 1u;
#pragma unroll
        for (int i = 0; i < WPT + 2; i++) {
            const int64_t j = jf + i;
            const uint64_t bits = i == 0 ? clo : ((clo >> i) | (chi << (64 - i)));
            const uint32_t bd = i == 0 ? bad0 : ((badn >> (i - 1)) & 1u);
            const bool ok = j >= 0 && (uint64_t)j + (k - 1) <= (uint64_t)n && bd == 0 && (bits & kmask) == 0;
            ok10 |= (ok ? 1u : 0u) << i;
        }
        if (WMAX > 64 && k > 66) {   // (block-uniform) the k-2 positions behind a window start span a second 64-bit word
            const uint64_t cx = extract64(s_cmb, rel + 128);
            const uint64_t kmask2 = (1ull << (k - 66)) - 1ull;      // k <= 128
#pragma unroll
            for (int i = 0; i < WPT + 2; i++) {
                const uint64_t bits2 = i == 0 ? chi : ((chi >> i) | (cx << (64 - i)));
                if (bits2 & kmask2) ok10 &= ~(1u << i);
            }
        }
    }

    // ---- E: split / segment-end flags of the thread's windows, F: in-order compaction (one block scan)
    uint32_t n_ent, n_s;
    {
        uint32_t fl = 0;  // 3 flag bits per window
#pragma unroll
        for (int i = 0; i < WPT; i++) {
            const uint32_t x = x0 + i;
            const bool okp = (ok10 >> i) & 1u, okj = (ok10 >> (i + 1)) & 1u, okn = (ok10 >> (i + 2)) & 1u;
            const bool valid = okj && (okp || okn);        // segment has >= 2 windows <=> length >= k
            const bool first = okj && !okp;
            if (valid) {
                const uint64_t M = Mreg[i], Mp = i ? Mreg[i ? i - 1 : 0] : Mprev;
                const bool S = first || M != Mp || ((M & 1ull) && s_v0[PADX(x - 1)] == M);
                fl |= ((S ? 1u : 0u) | (!okn ? 2u : 0u) | (first ? 4u : 0u)) << (3 * i);
            }
        }
        if (fl & 0x924924u) atomicAdd(&s_nfirst, (uint32_t)__popc(fl & 0x924924u));   // bit 2 of every 3-bit group: segment starts (rare)
        uint32_t mine = 0;
#pragma unroll
        for (int i = 0; i < WPT; i++) {
            const uint32_t f = (fl >> (3 * i)) & 7u;
            if (f & 3u) mine += 1u + ((f & 1u) << 16);
        }
        uint32_t tot;
        uint32_t pre = block_exclusive_scan<WIN_THREADS>(mine, s_scan, &tot);   // (three block barriers)
#pragma unroll
        for (int i = 0; i < WPT; i++) {
            const uint32_t f = (fl >> (3 * i)) & 7u;
            if (f & 3u) {
                uint64_t e = (uint64_t)(x0 + i - 1) | ((f & 1u) ? ENT_S : 0) | ((f & 2u) ? ENT_E : 0) |
                             ((f & 4u) ? ENT_FIRST : 0) | ((uint64_t)(pre >> 16) << ENT_SRANK_SHIFT);
                if (f & 1u)   // a super-k-mer starts here: remember where its minimizer is (and whether it is unique)
                    e |= ((uint64_t)(Preg[i] - (x0 + i)) << ENT_ARG_SHIFT) | ((Mreg[i] & 1ull) ? 0 : ENT_DUP);
                s_ent[pre & 0xFFFFu] = e;
                pre += 1u + ((f & 1u) << 16);
            }
        }
        n_ent = tot & 0xFFFFu; n_s = tot >> 16;
    }
    __syncthreads();
    if (WIN_TPC > 2 && tid == 0 && tl + 2 < (uint32_t)WIN_TPC) {   // this tile's landing buffer is free: start the copies of tile tl+2
        fence_proxy_async_smem();
        issue(tl + 2);
    }

    // ---- G: per super-k-mer: bucket / orientation from its minimizer (assembler_minimizer_bucketing/src/lib.rs:218-238);
    //         the minimizer's position came out of the window reduction, nothing is searched
    for (uint32_t i = tid; i < n_ent; i += WIN_THREADS) {
        uint64_t e = s_ent[i];
        if (e & ENT_S) {
            const uint32_t x = (uint32_t)(e & ((1u << ENT_POS_BITS) - 1)) + 1;
            const uint32_t arg = (uint32_t)(e >> ENT_ARG_SHIFT) & 0xFFu;
            const uint64_t M = s_v0[PADX(x + arg)];     // the window minimum (bit 0 aside: ENT_DUP says whether it is unique)
            uint32_t bucket, rcf = 0;
            if (e & ENT_DUP) {
                bucket = 1u << P.b1;  // duplicates bucket
                e &= ~(0xFFull << ENT_ARG_SHIFT);
            } else {
                rcf = (!P.forward_only && !s_fwd[x + arg]) ? 1u : 0u;
                bucket = (uint32_t)(M >> 1) & ((1u << P.b1) - 1);   // cn_nthash.rs:135-142 get_bucket(0, b1, M)
            }
            const uint32_t second = (uint32_t)(M >> (P.b1 + 1)) & ((1u << P.b2) - 1);
            e |= (rcf ? ENT_RC : 0) | ((uint64_t)second << ENT_SECOND_SHIFT) | ((uint64_t)bucket << ENT_BUCKET_SHIFT);
        }
        ent[(uint64_t)tile * WIN_T + i] = e;
    }
    if (tid == 0) {
        tile_cnt[tile] = n_ent; tile_scnt[tile] = n_s;
        if (n_s) atomicAdd(&group_scnt[tile / WIN_GROUP], n_s);   // the host scans the groups (n_tiles / 64 values), k_emit finishes the prefix
    }
    if (WIN_TPC > 1) __syncthreads();     // s_v0 / s_fwd / s_ent are rewritten by the next tile
    }   // tiles of this CTA
    if (tid == 0 && s_nfirst) atomicAdd(seg_count, s_nfirst);     // SequencesSplitter::valid_bases bookkeeping (one atomic per CTA)
}

// ------------------------------------------------------------------------------------------------
// Single-CTA exclusive scan of a u32 array (n up to a few million); out may alias in.
// total (u64) written to *total_out.  Used for per-tile and per-unit offsets.
__device__ __forceinline__ void exclusive_scan_u32_cta(const uint32_t *in, uint32_t *out, uint32_t n, unsigned long long *total_out);
__global__ void __launch_bounds__(1024) k_exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint32_t n,
                                                             unsigned long long *total_out) {
    exclusive_scan_u32_cta(in, out, n, total_out);
}
// Several independent scans in ONE launch, one CTA each (the per-unit offsets of the slices an owner received: two arrays per
// source rank and chunk).  The job list travels in the kernel parameters.
constexpr int SCAN_MAX_JOBS = 64;
struct ScanJobs { const uint32_t *in[SCAN_MAX_JOBS]; uint32_t *out[SCAN_MAX_JOBS]; uint32_t n; uint32_t pad; };
__global__ void __launch_bounds__(1024) k_exclusive_scan_u32_jobs(const __grid_constant__ ScanJobs jobs) {
    exclusive_scan_u32_cta(jobs.in[blockIdx.x], jobs.out[blockIdx.x], jobs.n, nullptr);
}
__device__ __forceinline__ void exclusive_scan_u32_cta(const uint32_t *in, uint32_t *out, uint32_t n, unsigned long long *total_out) {
    __shared__ uint32_t s_scan[1024 / 32 + 2];
    constexpr uint32_t IPT = 8;
    uint64_t running = 0;
    for (uint32_t base = 0; base < n; base += 1024 * IPT) {
        uint32_t v[IPT];
        uint32_t sum = 0;
        const uint32_t i0 = base + threadIdx.x * IPT;
#pragma unroll
        for (uint32_t q = 0; q < IPT; q++) { v[q] = (i0 + q < n) ? in[i0 + q] : 0u; sum += v[q]; }
        uint32_t tot;
        uint32_t pre = block_exclusive_scan<1024>(sum, s_scan, &tot);
        uint64_t p = running + pre;
#pragma unroll
        for (uint32_t q = 0; q < IPT; q++) {
            if (i0 + q < n) out[i0 + q] = (uint32_t)p;
            p += v[q];
        }
        running += tot;
    }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = running;
        out[n] = (uint32_t)running;  // arrays carry one extra slot
    }
}

// ------------------------------------------------------------------------------------------------
// k_emit: entry -> super-k-mer descriptor in position order + per-unit histogram.
// tmp descriptor: {start (position in the chunk's packed bases), len, meta, unit}.
__global__ void __launch_bounds__(256)
k_emit(const uint64_t *__restrict__ ent, const uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ tile_scnt,
       const uint32_t *__restrict__ group_sbase, uint32_t n_tiles, DevParams P, uint4 *__restrict__ tmp, uint32_t *__restrict__ tmp_color,
       uint32_t base /* position of the batch's first base inside the chunk's packed bases */,
       const uint64_t *__restrict__ offsets, uint64_t n_reads, uint64_t off0, const uint32_t *__restrict__ colors,
       uint32_t *__restrict__ unit_cnt, uint32_t *__restrict__ unit_words, uint32_t *__restrict__ unit_kmers) {
    const uint32_t tile = blockIdx.x;
    const uint32_t cnt = tile_cnt[tile];
    if (cnt == 0) return;
    // first super-k-mer of the tile = scanned total of the groups before + the tiles of this group before this one
    __shared__ uint32_t s_sbase;
    if (threadIdx.x < 32) {
        const uint32_t g0 = (tile / WIN_GROUP) * WIN_GROUP, l = threadIdx.x;
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < WIN_GROUP / 32; q++) { const uint32_t t = g0 + l + 32u * q; if (t < tile) v += tile_scnt[t]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (l == 0) s_sbase = group_sbase[tile / WIN_GROUP] + v;
    }
    __syncthreads();
    const uint32_t tile_first_sk = s_sbase;
    const uint32_t posmask = (1u << ENT_POS_BITS) - 1;
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint64_t e = ent[(uint64_t)tile * WIN_T + i];
        if (!(e & ENT_S)) continue;
        const uint32_t pos = tile * WIN_T + (uint32_t)(e & posmask);
        const bool first = e & ENT_FIRST;
        uint32_t end;
        bool last;
        if (e & ENT_E) {
            end = pos + P.k - 1; last = true;
        } else {
            uint64_t ne;
            uint32_t t2 = tile;
            if (i + 1 < cnt) ne = ent[(uint64_t)tile * WIN_T + i + 1];
            else {
                do { ++t2; } while (t2 < n_tiles && tile_cnt[t2] == 0);
                ne = t2 < n_tiles ? ent[(uint64_t)t2 * WIN_T] : (ENT_E);  // unreachable fallback
            }
            const uint32_t npos = t2 * WIN_T + (uint32_t)(ne & posmask);
            end = npos + P.k - 1;
            last = !(ne & ENT_S);
        }
        const uint32_t start = first ? pos : pos - 1;
        const uint32_t len = end - start;
        const uint32_t rc = (e & ENT_RC) ? 1u : 0u;
        const uint32_t arg = pos + (uint32_t)((e >> ENT_ARG_SHIFT) & 0xFFu);  // global m-mer position
        uint32_t mpos = 0;
        if (!(e & ENT_DUP)) mpos = rc ? (end - arg - P.m) : (arg - start);
        const uint32_t flags = ((first ? 1u : 0u) << rc) | ((last ? 1u : 0u) << (rc ^ 1u));
        const uint32_t second = (uint32_t)(e >> ENT_SECOND_SHIFT) & 0xFFu;
        const uint32_t bucket = (uint32_t)(e >> ENT_BUCKET_SHIFT) & 0x3FFFu;
        const uint32_t unit = (bucket << P.b2) | second;
        const uint32_t idx = tile_first_sk + (uint32_t)((e >> ENT_SRANK_SHIFT) & 0xFFFFu);
        tmp[idx] = make_uint4(base + start, len, make_meta(mpos, flags, rc, second), unit);
        if (P.colors) {
            // record index = last r with offsets[r] <= start
            uint64_t lo = 0, hi = n_reads;
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) >> 1;
                if (offsets[mid] - off0 <= start) lo = mid; else hi = mid;
            }
            tmp_color[idx] = colors ? colors[lo] : 0u;
        }
        atomicAdd(&unit_cnt[unit], 1u);
        atomicAdd(&unit_words[unit], (len + 15u) >> 4);
        atomicAdd(&unit_kmers[unit], len - P.k + 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// k_scatter: one thread per super-k-mer.  Reserves a descriptor slot AND its payload words in its unit with ONE 64-bit
// atomic (cursor = slot << 32 | word), so descriptors and payload of a unit appear in the same order: any contiguous
// descriptor range of a unit owns a contiguous payload range.  Writes the final descriptor {payload word offset, len,
// meta, colour} and the payload in stored orientation
// (crates/io/src/concurrent/temp_reads/creads_utils.rs:389-406: rc => reverse-complement packing).
__global__ void k_init_cursors(const uint32_t *__restrict__ unit_off, const uint32_t *__restrict__ unit_woff, uint32_t n_units,
                               unsigned long long *__restrict__ cur) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_units) cur[u] = ((unsigned long long)unit_off[u] << 32) | (unsigned long long)unit_woff[u];
}

__global__ void __launch_bounds__(256)
k_scatter(const uint4 *__restrict__ tmp, const uint32_t *__restrict__ tmp_color, uint32_t n_sk,
          const uint32_t *__restrict__ pk, unsigned long long *__restrict__ cur,
          uint4 *__restrict__ desc, uint32_t *__restrict__ payload, uint32_t with_color) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sk) return;
    const uint4 t = tmp[i];
    const uint32_t start = t.x, len = t.y, meta = t.z, unit = t.w;
    const uint32_t nw = (len + 15u) >> 4;
    const unsigned long long cw = atomicAdd(&cur[unit], (1ull << 32) | (unsigned long long)nw);
    const uint32_t slot = (uint32_t)(cw >> 32), woff = (uint32_t)cw;
    desc[slot] = make_uint4(woff, len, meta, with_color ? tmp_color[i] : 0u);
    const bool rc = (meta >> 18) & 1u;
    uint32_t *dst = payload + woff;
    const uint32_t tail = len & 15u;
    if (!rc) {
        for (uint32_t q = 0; q < nw; q++) {
            uint32_t v = extract32(pk, 2ull * (start + 16u * q));
            if (q == nw - 1 && tail) v &= (1u << (2 * tail)) - 1u;
            dst[q] = v;
        }
    } else {
        for (uint32_t q = 0; q < nw; q++) {
            // stored bases [16q, 16q+16) = complement of original bases start+len-1-16q downwards
            const int64_t p = (int64_t)start + (int64_t)len - 16 * ((int64_t)q + 1);
            uint32_t v;
            if (p >= (int64_t)start) v = revcomp32(extract32(pk, 2ull * (uint64_t)p));
            else {
                v = revcomp32(extract32(pk, 2ull * start)) >> (2u * (uint32_t)((int64_t)start - p));
                v &= (1u << (2 * tail)) - 1u;
            }
            dst[q] = v;
        }
    }
}

}  // namespace ggb
