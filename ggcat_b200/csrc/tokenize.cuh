// tokenize.cuh -- FASTA / FASTQ text -> sequence records, on the device.
//
// Replaces the reference's line reader + record state machines
//   crates/io/src/lines_reader.rs:140-175        (lines end at '\n', a '\r' right before it is dropped)
//   crates/io/src/sequences_reader.rs:106-179     process_fasta: '>' line = new record, ';' line = comment, every other
//                                                 line is appended to the sequence; records without bases are not emitted
//   crates/io/src/sequences_reader.rs:181-241     process_fastq: strict 4-line records (ident, bases, '+', qualities)
// by data-parallel passes over the raw bytes (normalisation of the bases is k_pack's job, as in phase 1):
//   a line's KIND follows from its first byte (FASTA) or from its line number modulo 4 (FASTQ); a byte knows its line
//   through a scan over the text (FASTA: max-scan of "position | kind" of line starts; FASTQ: sum-scan of newlines);
//   sequence bytes are compacted with a sum-scan; every record start marks the compacted position it lands on, and a
//   second scan over those marks yields the record offsets (records without bases collapse onto one mark and vanish,
//   exactly like `if intermediate[SEQ_STATE].len() > 0` in the reference).
// Scans are three-phase (tile aggregate, single-CTA scan of the aggregates, tile apply): the text is read three times
// and written once, ~4 B of HBM traffic per input byte.
#pragma once
#include "device_utils.cuh"

namespace ggb {

constexpr int TOK_THREADS = 256, TOK_PER_THREAD = 32, TOK_TILE = TOK_THREADS * TOK_PER_THREAD;   // 8192 bytes per CTA
enum { TOK_FASTA = 0, TOK_FASTQ = 1 };
enum { LINE_SEQ = 1, LINE_HEADER = 2, LINE_COMMENT = 3 };

__device__ __forceinline__ uint32_t tok_fasta_kind(uint8_t first) { return first == '>' ? LINE_HEADER : first == ';' ? LINE_COMMENT : LINE_SEQ; }

// Per-tile aggregate of the line scan.  FASTA: max over line starts of (position << 2 | kind), 0 = no line start in the
// tile.  FASTQ: number of '\n' in the tile.
template <int FORMAT>
__global__ void __launch_bounds__(TOK_THREADS) k_tok_tile_lines(const uint8_t *__restrict__ text, uint64_t n, unsigned long long *__restrict__ tile_agg) {
    __shared__ unsigned long long s_red[TOK_THREADS / 32];
    const uint64_t t0 = (uint64_t)blockIdx.x * TOK_TILE + (uint64_t)threadIdx.x * TOK_PER_THREAD;
    unsigned long long agg = 0;
    for (int q = 0; q < TOK_PER_THREAD; q++) {
        const uint64_t i = t0 + q;
        if (i >= n) break;
        const uint8_t b = text[i];
        if (FORMAT == TOK_FASTA) {
            const bool ls = i == 0 || text[i - 1] == '\n';
            if (ls) agg = ((unsigned long long)i << 2) | tok_fasta_kind(b);     // positions increase: the last one wins
        } else agg += b == '\n' ? 1ull : 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, agg, o);
        agg = FORMAT == TOK_FASTA ? (y > agg ? y : agg) : agg + y;
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = agg;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0;
        for (int w = 0; w < TOK_THREADS / 32; w++) a = FORMAT == TOK_FASTA ? (s_red[w] > a ? s_red[w] : a) : a + s_red[w];
        tile_agg[blockIdx.x] = a;
    }
}

// Single CTA: exclusive scan of the tile aggregates (max for FASTA, sum for FASTQ), in place.
template <int FORMAT>
__global__ void __launch_bounds__(1024) k_tok_scan_tiles(unsigned long long *__restrict__ agg, uint32_t n_tiles) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? agg[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= (uint32_t)o) x = FORMAT == TOK_FASTA ? (y > x ? y : x) : x + y;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
        __syncthreads();
        unsigned long long carry = s_run;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) carry = FORMAT == TOK_FASTA ? (s_w[w] > carry ? s_w[w] : carry) : carry + s_w[w];
        // exclusive: combine with the inclusive value of the previous lane
        unsigned long long prev = __shfl_up_sync(0xffffffffu, x, 1);
        if ((threadIdx.x & 31) == 0) prev = 0;
        const unsigned long long excl = FORMAT == TOK_FASTA ? (prev > carry ? prev : carry) : carry + prev;
        if (i < n_tiles) agg[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = FORMAT == TOK_FASTA ? (x > carry ? x : carry) : carry + x;
        __syncthreads();
    }
}

// The state every thread carries along its 32 bytes: kind of the current line (FASTA) / line number (FASTQ).
template <int FORMAT>
struct TokWalk {
    unsigned long long st;   // FASTA: (line start position << 2 | kind); FASTQ: lines completed before this byte
    __device__ __forceinline__ void at(const uint8_t *text, uint64_t i, uint8_t b) {
        if (FORMAT == TOK_FASTA) { if (i == 0 || text[i - 1] == '\n') st = ((unsigned long long)i << 2) | tok_fasta_kind(b); }
    }
    __device__ __forceinline__ void after(uint8_t b) { if (FORMAT == TOK_FASTQ && b == '\n') ++st; }
    __device__ __forceinline__ bool seq_line() const { return FORMAT == TOK_FASTA ? (st & 3ull) == LINE_SEQ : (st & 3ull) == 1ull; }
};

// Thread-level entry state inside a tile: block scan of the per-thread aggregates + the tile's carry-in.
template <int FORMAT>
__device__ __forceinline__ unsigned long long tok_thread_carry(const uint8_t *__restrict__ text, uint64_t n, uint64_t t0, unsigned long long tile_carry,
                                                               unsigned long long *s_w) {
    unsigned long long agg = 0;
    for (int q = 0; q < TOK_PER_THREAD; q++) {
        const uint64_t i = t0 + q;
        if (i >= n) break;
        const uint8_t b = text[i];
        if (FORMAT == TOK_FASTA) { if (i == 0 || text[i - 1] == '\n') agg = ((unsigned long long)i << 2) | tok_fasta_kind(b); }
        else agg += b == '\n' ? 1ull : 0ull;
    }
    unsigned long long x = agg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) x = FORMAT == TOK_FASTA ? (y > x ? y : x) : x + y;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned long long carry = tile_carry;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) carry = FORMAT == TOK_FASTA ? (s_w[w] > carry ? s_w[w] : carry) : carry + s_w[w];
    unsigned long long prev = __shfl_up_sync(0xffffffffu, x, 1);
    if ((threadIdx.x & 31) == 0) prev = 0;
    __syncthreads();
    return FORMAT == TOK_FASTA ? (prev > carry ? prev : carry) : carry + prev;
}

__device__ __forceinline__ bool tok_keep_byte(const uint8_t *text, uint64_t n, uint64_t i, uint8_t b) {
    if (b == '\n') return false;
    if (b == '\r' && (i + 1 == n || text[i + 1] == '\n')) return false;    // lines_reader.rs:152,165
    return true;
}

// WRITE == false: tile_keep[tile] = sequence bytes of the tile.  WRITE == true: compacted bytes -> seq[tile_base + ...] and a
// mark (bit in `marks`, one per compacted position) wherever a record starts.
template <int FORMAT, bool WRITE>
__global__ void __launch_bounds__(TOK_THREADS)
k_tok_compact(const uint8_t *__restrict__ text, uint64_t n, const unsigned long long *__restrict__ tile_carry, uint32_t *__restrict__ tile_keep,
              const uint32_t *__restrict__ tile_base, uint8_t *__restrict__ seq, uint32_t *__restrict__ marks) {
    __shared__ unsigned long long s_w[TOK_THREADS / 32];
    __shared__ uint32_t s_k[TOK_THREADS / 32 + 1];
    const uint64_t t0 = (uint64_t)blockIdx.x * TOK_TILE + (uint64_t)threadIdx.x * TOK_PER_THREAD;
    TokWalk<FORMAT> wk;
    wk.st = tok_thread_carry<FORMAT>(text, n, t0, tile_carry[blockIdx.x], s_w);
    const TokWalk<FORMAT> wk0 = wk;
    uint32_t cnt = 0;
    for (int q = 0; q < TOK_PER_THREAD; q++) {
        const uint64_t i = t0 + q;
        if (i >= n) break;
        const uint8_t b = text[i];
        wk.at(text, i, b);
        if (wk.seq_line() && tok_keep_byte(text, n, i, b)) ++cnt;
        wk.after(b);
    }
    // block exclusive scan of cnt
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= (uint32_t)o) x += y; }
    if ((threadIdx.x & 31) == 31) s_k[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t pre = x - cnt;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) pre += s_k[w];
    if (!WRITE) {
        if (threadIdx.x == TOK_THREADS - 1) tile_keep[blockIdx.x] = pre + cnt;
        return;
    }
    uint64_t o = (uint64_t)tile_base[blockIdx.x] + pre;
    wk = wk0;
    for (int q = 0; q < TOK_PER_THREAD; q++) {
        const uint64_t i = t0 + q;
        if (i >= n) break;
        const uint8_t b = text[i];
        const bool ls = i == 0 || text[i - 1] == '\n';
        wk.at(text, i, b);
        // a record starts where an ident line starts (FASTA '>' line; FASTQ line 0 of 4): mark the compacted position
        if (ls && (FORMAT == TOK_FASTA ? (wk.st & 3ull) == LINE_HEADER : (wk.st & 3ull) == 0ull)) atomicOr(&marks[o >> 5], 1u << (o & 31));
        if (wk.seq_line() && tok_keep_byte(text, n, i, b)) seq[o++] = b;
        wk.after(b);
    }
}

// Record offsets from the marks: per-tile mark counts, (single-CTA scan by the caller), then offsets[rank] = position.
// Position 0 always starts a record (bases before the first ident line form a record in the reference too).
__global__ void __launch_bounds__(256) k_tok_count_marks(uint32_t *__restrict__ marks, uint64_t n_words, uint64_t total, uint32_t *__restrict__ tile_marks) {
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    const uint64_t w = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t c = 0;
    if (w < n_words) {
        uint32_t m = marks[w];
        if (w == 0) m |= 1u;
        // marks at or beyond `total` (an ident line after the last base) start no record
        const uint64_t lo = w * 32;
        if (lo + 32 > total) m &= total > lo ? (0xFFFFFFFFu >> (32 - (uint32_t)(total - lo))) : 0u;
        marks[w] = m;
        c = __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_marks[blockIdx.x] = s_c;
}

__global__ void __launch_bounds__(256) k_tok_write_offsets(const uint32_t *__restrict__ marks, uint64_t n_words, const uint32_t *__restrict__ tile_rank,
                                                          uint64_t *__restrict__ offsets) {
    __shared__ uint32_t s_k[256 / 32];
    const uint64_t w = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    const uint32_t m = w < n_words ? marks[w] : 0u;
    const uint32_t cnt = __popc(m);
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= (uint32_t)o) x += y; }
    if ((threadIdx.x & 31) == 31) s_k[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t r = tile_rank[blockIdx.x] + x - cnt;
    for (uint32_t q = 0; q < (threadIdx.x >> 5); q++) r += s_k[q];
    uint32_t mm = m;
    while (mm) {
        const uint32_t bit = __ffs(mm) - 1;
        mm &= mm - 1;
        offsets[r++] = w * 32 + bit;
    }
}

}  // namespace ggb
