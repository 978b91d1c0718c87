/*
 * ggcat_oracle.c -- CPU restatement of GGCAT's k-mer counting front end.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * ggcat_b200/csrc.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py may load it.  The product path never does.
 *
 * PARITY PINNING: the reference (algbio/ggcat 2.2.0, pure Rust) cannot be compiled in
 * this environment (no cargo/rustc), and its tests contain no golden bucket contents
 * or k-mer counts.  This oracle is therefore pinned by
 *   (1) the reference's known-answer tests for 2-bit packing / rc packing
 *       (crates/io/src/compressed_read.rs:1002-1042) and varints (crates/io/src/varint.rs:106-135),
 *   (2) the reference's hash *property* tests (crates/hashes/src/lib.rs:265-454):
 *       canonical symmetry, roll consistency, invertibility,
 *   (3) all numeric constants copied from source (cited below), and
 *   (4) an independent naive whole-input k-mer counter (orc_naive_*) that does not share
 *       code with the bucketed path.
 * End-to-end bucket contents are "parity unpinned" by reference goldens (none exist).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * This is a restatement, not a copy: the Rust generic/iterator machinery is replaced by plain C.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef unsigned __int128 u128;

#define ORC_HASH_SEQ 1 /* api/src/utils.rs:4-8  HashType::SeqHash = 1 */
#define ORC_HASH_RK128 4 /* HashType::RabinKarp128 = 4 */

#define READ_FLAG_INCL_BEGIN 1 /* config/src/lib.rs:93 */
#define READ_FLAG_INCL_END 2   /* config/src/lib.rs:94 */

/* ------------------------------------------------------------------------------------------
 * A.1  alphabet, normalisation, N-splitting, 2-bit packing
 * ---------------------------------------------------------------------------------------- */

/* utils/src/lib.rs:44-46  Utils::compress_base : A0 C1 T2 G3 */
static inline uint8_t compress_base(uint8_t c) { return (c >> 1) & 3; }
/* utils/src/lib.rs:49-51  compress_base_complement */
static inline uint8_t compress_base_complement(uint8_t c) { return ((c >> 1) & 3) ^ 2; }

/* io/src/sequences_reader.rs:26-37  SEQ_LETTERS_MAPPING; :50-54 normalize_sequence */
void orc_normalize(uint8_t *seq, size_t n) {
    for (size_t i = 0; i < n; i++) {
        switch (seq[i]) {
        case 'A': case 'a': seq[i] = 'A'; break;
        case 'C': case 'c': seq[i] = 'C'; break;
        case 'G': case 'g': seq[i] = 'G'; break;
        case 'T': case 't': seq[i] = 'T'; break;
        default: seq[i] = 'N';
        }
    }
}

/* minimizer_bucketing/src/sequences_splitter.rs:15-40  SequencesSplitter::process_sequences.
 * Writes (start,end) pairs of segments with end-start >= k; returns how many. */
size_t orc_split_segments(const uint8_t *seq, size_t n, size_t k, uint64_t *starts, uint64_t *ends, size_t cap) {
    size_t start, end = 0, cnt = 0;
    while (end < n) {
        start = end;
        while (start < n && (((seq[start] ^ 'N') & 0x7) == 0)) start++;
        end = start;
        while (end < n && (((seq[end] ^ 'N') & 0x7) != 0)) end++;
        if (end - start >= k) {
            if (cnt < cap) { starts[cnt] = start; ends[cnt] = end; }
            cnt++;
        }
    }
    return cnt;
}

/* io/src/compressed_read.rs:610-618  compress_from_plain: 16-base chunks, base i at bits 2(i%4)
 * of byte i/4 (LE); returns bytes written = ceil(n/4). */
size_t orc_compress_from_plain(const uint8_t *seq, size_t n, uint8_t *out) {
    size_t w = 0;
    for (size_t c0 = 0; c0 < n; c0 += 16) {
        size_t len = n - c0 < 16 ? n - c0 : 16;
        uint32_t value = 0;
        for (size_t j = len; j-- > 0;) value = (value << 2) | compress_base(seq[c0 + j]);
        size_t nb = (len + 3) / 4;
        for (size_t b = 0; b < nb; b++) out[w++] = (uint8_t)(value >> (8 * b));
    }
    return w;
}

/* io/src/compressed_read.rs:621-633  compress_from_plain_rc: walk the sequence from the end in
 * 16-base chunks, complementing each base. */
size_t orc_compress_from_plain_rc(const uint8_t *seq, size_t n, uint8_t *out) {
    size_t w = 0, current = n;
    while (current > 0) {
        size_t amount = current < 16 ? current : 16;
        const uint8_t *chunk = seq + current - amount;
        uint32_t value = 0;
        for (size_t j = 0; j < amount; j++) value = (value << 2) | compress_base_complement(chunk[j]);
        size_t nb = (amount + 3) / 4;
        for (size_t b = 0; b < nb; b++) out[w++] = (uint8_t)(value >> (8 * b));
        current -= amount;
    }
    return w;
}

/* io/src/compressed_read.rs:882-885  get_base_unchecked */
static inline uint8_t packed_base(const uint8_t *data, size_t index) {
    return (data[index / 4] >> ((index % 4) * 2)) & 3;
}

/* utils/src/lib.rs:17  C_INV_LETTERS = A C T G ; CompressedRead::to_string */
void orc_unpack(const uint8_t *packed, size_t start_base, size_t n, uint8_t *ascii_out) {
    static const uint8_t inv[4] = {'A', 'C', 'T', 'G'};
    for (size_t i = 0; i < n; i++) ascii_out[i] = inv[packed_base(packed, start_base + i)];
}

/* ------------------------------------------------------------------------------------------
 * varints  (io/src/varint.rs)
 * ---------------------------------------------------------------------------------------- */

/* io/src/varint.rs:8-24  encode_varint */
size_t orc_encode_varint(uint64_t value, uint8_t *out) {
    size_t index = 0;
    while (index < 9) {
        uint8_t rem = (uint8_t)((value > 127) << 7);
        out[index] = ((uint8_t)value & 0x7f) | rem;
        value >>= 7;
        index++;
        if (value == 0) break;
    }
    return index;
}

/* io/src/varint.rs:26-56  encode_varint_flags<FlagsCount> */
size_t orc_encode_varint_flags(uint64_t value, uint8_t flags, int flags_count, uint8_t *out) {
    int useful_first_bits = 8 - flags_count;
    uint8_t first_byte_max_value = (uint8_t)((1u << (useful_first_bits - 1)) - 1);
    uint8_t fr_rem = (uint8_t)((value > first_byte_max_value) << (useful_first_bits - 1));
    out[0] = (uint8_t)(((uint16_t)flags) << useful_first_bits) | ((uint8_t)value & first_byte_max_value) | fr_rem;
    value >>= (useful_first_bits - 1);
    size_t index = 1;
    while (index < 10) {
        if (value == 0) break;
        uint8_t rem = (uint8_t)((value > 127) << 7);
        out[index] = ((uint8_t)value & 0x7f) | rem;
        value >>= 7;
        index++;
    }
    return index;
}

/* io/src/varint.rs:58-84  decode_varint_flags; returns bytes consumed (0 on truncation) */
size_t orc_decode_varint_flags(const uint8_t *in, size_t n, int flags_count, uint64_t *value, uint8_t *flags) {
    if (n == 0) return 0;
    size_t pos = 0;
    uint8_t first_byte = in[pos++];
    int useful_first_bits = 8 - flags_count;
    uint8_t first_byte_max_value = (uint8_t)((1u << (useful_first_bits - 1)) - 1);
    *flags = (uint8_t)(((uint16_t)first_byte) >> useful_first_bits);
    uint64_t result = first_byte & first_byte_max_value;
    int offset = useful_first_bits - 1;
    int next = (first_byte & (1 << (useful_first_bits - 1))) != 0;
    while (next) {
        if (pos >= n) return 0;
        uint8_t v = in[pos++];
        next = (v & 0x80) != 0;
        result |= ((uint64_t)(v & 0x7f)) << offset;
        offset += 7;
    }
    *value = result;
    return pos;
}

/* io/src/varint.rs:86-98  decode_varint */
size_t orc_decode_varint(const uint8_t *in, size_t n, uint64_t *value) {
    uint64_t result = 0;
    unsigned offset = 0;
    size_t pos = 0;
    for (;;) {
        if (pos >= n) return 0;
        uint8_t v = in[pos++];
        int next = (v & 0x80) != 0;
        result |= ((uint64_t)(v & 0x7f)) << offset;
        if (!next) break;
        offset += 7;
    }
    *value = result;
    return pos;
}

/* ------------------------------------------------------------------------------------------
 * A.2  m-mer hash: canonical ntHash-shaped rolling hash
 * ---------------------------------------------------------------------------------------- */

#define NT_MULTIPLIER 0x397f178c6ae330f9ULL /* hashes/src/nthash_base.rs:49 */

static inline uint64_t rotl64(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static inline uint64_t rotr64(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }

/* hashes/src/nthash_base.rs:51-55  h(c, compressed=false): ((c & 6) + 1) * MULTIPLIER */
static inline uint64_t nt_h(uint8_t c) { return (uint64_t)((c & 0x6) + 1) * NT_MULTIPLIER; }
/* hashes/src/nthash_base.rs:57-61  rc(c, compressed=false): (((c & 6) ^ 4) + 1) * MULTIPLIER */
static inline uint64_t nt_rc(uint8_t c) { return (uint64_t)(((c & 0x6) ^ 4) + 1) * NT_MULTIPLIER; }

/* hashes/src/cn_nthash.rs:21-58  CanonicalNtHashIterator::{new, roll_hash}; writes n-m+1 pairs.
 * seq is ASCII (IS_COMPRESSED = false for &[u8]). */
size_t orc_nthash_iter(const uint8_t *seq, size_t n, size_t m, uint64_t *out_fw, uint64_t *out_rc) {
    if (m > n) return 0;
    size_t k_minus1 = m - 1;
    uint64_t fh = 0, bw = 0;
    for (size_t i = 0; i < m - 1; i++) {
        fh ^= rotl64(nt_h(seq[i]), (unsigned)(m - i - 2));
        bw ^= rotl64(nt_rc(seq[i]), (unsigned)i);
    }
    size_t cnt = n - k_minus1;
    for (size_t i = 0; i < cnt; i++) {
        uint8_t base_i = seq[i], base_k = seq[i + k_minus1];
        uint64_t seqi_h = nt_h(base_i), seqk_h = nt_h(base_k);
        uint64_t seqi_rc = nt_rc(base_i), seqk_rc = nt_rc(base_k);
        uint64_t res = rotl64(fh, 1) ^ seqk_h;
        fh = res ^ rotl64(seqi_h, (unsigned)k_minus1);
        uint64_t res_rc = bw ^ rotl64(seqk_rc, (unsigned)k_minus1);
        bw = rotr64(res_rc ^ seqi_rc, 1);
        out_fw[i] = res;
        out_rc[i] = res_rc;
    }
    return cnt;
}

/* hashes/src/cn_nthash.rs:135-142  get_bucket: ((hash >> (used_bits + 1)) % (1 << requested_bits)) */
static inline uint16_t nt_get_bucket(unsigned used_bits, unsigned requested_bits, uint64_t hash) {
    return (uint16_t)((hash >> (used_bits + 1)) % (1ULL << requested_bits));
}

/* ------------------------------------------------------------------------------------------
 * A.3  BatchMinQueue (literal restatement of the batched two-array algorithm)
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    uint64_t v;          /* (min(fw,rc) << 1) | unique_flag */
    uint32_t index;      /* MinimizerExtraData.index       (assembler_minimizer_bucketing/src/lib.rs:31-35) */
    uint8_t is_forward;  /* MinimizerExtraData.is_forward */
} mq_item;

static inline int mq_extra_eq(const mq_item *a, const mq_item *b) {
    return a->index == b->index && a->is_forward == b->is_forward;
}
/* std::cmp::min_by_key(a, b, |x| x.0): returns a when keys are equal */
static inline mq_item mq_min(mq_item a, mq_item b) { return (b.v < a.v) ? b : a; }

typedef void (*mq_min_cb)(void *ctx, mq_item m, size_t index);
typedef void (*mq_flush_cb)(void *ctx, int is_last);

/* hashes/src/rolling/batch_minqueue.rs:37-124  get_minimizers::<_, ENABLE_DUPLICATE_CHECKING=true>.
 * `items` is the iterator (already advanced past skip_beginning), n_items = iter.len(). */
static void mq_get_minimizers(size_t qsize, const mq_item *items, size_t n_items, size_t skip_ending_count,
                              mq_min_cb minimizers_callback, mq_flush_cb minimizers_flush, void *ctx) {
    const uint64_t hash_mask = ~(uint64_t)1;
    size_t blen = qsize - 1; /* self.backward.len() */
    if (n_items < skip_ending_count) return;
    size_t size = n_items - skip_ending_count;
    if (size < blen) return;
    mq_item *backward = (mq_item *)malloc(sizeof(mq_item) * (blen ? blen : 1));
    size_t it = 0;
    for (size_t i = 0; i < blen; i++) backward[i] = items[it++];
    size_t offset = 0;
    size -= blen;
    while (offset < size) {
        size_t remaining = size - offset < blen ? size - offset : blen;
        if (blen >= 2) {
            for (size_t i = blen - 1; i-- > 0;) {
                mq_item current = backward[i];
                mq_item next = backward[i + 1];
                int is_duplicated = current.v == next.v;
                if (is_duplicated) current.v &= hash_mask;
                backward[i] = mq_min(current, next);
            }
        }
        mq_item new_item = items[it++];
        mq_item first_minimum = mq_min(new_item, backward[0]);
        int is_duplicated = new_item.v == backward[0].v;
        if (is_duplicated) first_minimum.v &= hash_mask;
        minimizers_callback(ctx, first_minimum, offset);
        mq_item last_forward = new_item;
        backward[0] = new_item;
        for (size_t i = 1; i < remaining; i++) {
            mq_item ni = items[it++];
            mq_item *current_backward = &backward[i];
            int new_item_duplicated = ni.v == last_forward.v;
            if (new_item_duplicated) last_forward.v &= hash_mask;
            last_forward = mq_min(last_forward, ni);
            mq_item current_minimum = mq_min(last_forward, *current_backward);
            int dup = last_forward.v == current_backward->v;
            if (dup) current_minimum.v &= hash_mask;
            minimizers_callback(ctx, current_minimum, offset + i);
            *current_backward = ni;
        }
        offset += remaining;
        minimizers_flush(ctx, offset == size);
    }
    free(backward);
}

typedef void (*mq_split_cb)(void *ctx, size_t position, mq_item value, int is_last);

typedef struct {
    mq_item last_value;
    int is_first;
    size_t last_index;
    size_t skip_beginning_count;
    mq_item *splits_val;
    size_t *splits_pos;
    size_t splits_len; /* splits_ptr - splits_start */
    mq_split_cb cb;
    void *cb_ctx;
} mq_splits_state;

/* hashes/src/rolling/batch_minqueue.rs:148-169  the per-window closure */
static void mq_splits_on_min(void *vctx, mq_item m, size_t index) {
    mq_splits_state *s = (mq_splits_state *)vctx;
    int has_different_value =
        !s->is_first && (s->last_value.v != m.v || (!mq_extra_eq(&s->last_value, &m) && (m.v & 1) == 1));
    s->splits_val[s->splits_len] = s->last_value;
    s->splits_pos[s->splits_len] = index + s->skip_beginning_count;
    s->splits_len += (size_t)has_different_value;
    s->is_first = 0;
    s->last_value = m;
}

/* hashes/src/rolling/batch_minqueue.rs:170-186  the flush closure */
static void mq_splits_on_flush(void *vctx, int is_last) {
    mq_splits_state *s = (mq_splits_state *)vctx;
    if (is_last) {
        s->splits_val[s->splits_len] = s->last_value;
        s->splits_pos[s->splits_len] = s->last_index;
        s->splits_len += 1;
    }
    for (size_t c = 0; c < s->splits_len; c++) {
        int last = (c + 1 == s->splits_len);
        s->cb(s->cb_ctx, s->splits_pos[c], s->splits_val[c], last && is_last);
    }
    s->splits_len = 0;
}

/* hashes/src/rolling/batch_minqueue.rs:127-188  get_minimizer_splits */
static void mq_get_minimizer_splits(size_t qsize, const mq_item *items, size_t n_items, size_t skip_beginning_count,
                                    size_t skip_ending_count, mq_split_cb cb, void *cb_ctx) {
    mq_splits_state s;
    memset(&s, 0, sizeof(s));
    s.is_first = 1;
    s.last_index = n_items - qsize; /* iter.len() - self.size, before skipping */
    s.skip_beginning_count = skip_beginning_count;
    s.splits_val = (mq_item *)calloc(qsize + 2, sizeof(mq_item));
    s.splits_pos = (size_t *)calloc(qsize + 2, sizeof(size_t));
    s.cb = cb;
    s.cb_ctx = cb_ctx;
    if (skip_beginning_count > n_items) skip_beginning_count = n_items;
    mq_get_minimizers(qsize, items + skip_beginning_count, n_items - skip_beginning_count, skip_ending_count,
                      mq_splits_on_min, mq_splits_on_flush, &s);
    free(s.splits_val);
    free(s.splits_pos);
}

/* Test hook: window minima exactly as the batched code reports them.
 * vals: n (value|flag) items, idx implicit = position; out_v/out_idx receive n-w+1 entries. */
typedef struct { uint64_t *out_v; uint32_t *out_idx; size_t cnt; } mq_dump_ctx;
static void mq_dump_min(void *c, mq_item m, size_t index) {
    mq_dump_ctx *d = (mq_dump_ctx *)c;
    d->out_v[index] = m.v; d->out_idx[index] = m.index; if (index + 1 > d->cnt) d->cnt = index + 1;
}
static void mq_dump_flush(void *c, int l) { (void)c; (void)l; }
size_t orc_window_minima(const uint64_t *vals, size_t n, size_t w, uint64_t *out_v, uint32_t *out_idx) {
    mq_item *items = (mq_item *)malloc(sizeof(mq_item) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) { items[i].v = vals[i]; items[i].index = (uint32_t)i; items[i].is_forward = 0; }
    mq_dump_ctx d = {out_v, out_idx, 0};
    mq_get_minimizers(w, items, n, 0, mq_dump_min, mq_dump_flush, &d);
    free(items);
    return d.cnt;
}

/* ------------------------------------------------------------------------------------------
 * A.3 (cont.)  super-k-mer emission: AssemblerMinimizerBucketingExecutor::process_sequence
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    uint32_t read_index;   /* index of the input record */
    uint32_t start;        /* first base of the super-k-mer inside the input record */
    uint32_t len;          /* bases */
    uint32_t color;        /* file colour (0 when uncoloured) */
    uint16_t bucket;       /* first-level bucket; duplicates bucket = 1 << b1 */
    uint16_t minimizer_pos;
    uint8_t second_bucket;
    uint8_t flags;         /* READ_FLAG_INCL_BEGIN | READ_FLAG_INCL_END, stored orientation */
    uint8_t rc;            /* stored reverse-complemented */
    uint8_t pad;
} orc_superkmer;

typedef struct {
    orc_superkmer *out;
    size_t cap, cnt;
    uint32_t read_index, seg_start, color;
    size_t k, m;
    unsigned first_bits, second_bits;
    int canonical;          /* !forward_only */
    uint16_t duplicates_bucket;
    int include_first, include_last_pre;
    size_t last_index;
} ps_ctx;

/* assembler_minimizer_bucketing/src/lib.rs:214-268  the splits callback */
static void ps_on_split(void *vctx, size_t index, mq_item min_hash, int is_last) {
    ps_ctx *c = (ps_ctx *)vctx;
    int include_last = c->include_last_pre && is_last;
    uint16_t bucket, minimizer_pos;
    int rc;
    if ((min_hash.v & 1) == 0) {
        bucket = c->duplicates_bucket; rc = 0; minimizer_pos = 0;
    } else {
        rc = c->canonical && !min_hash.is_forward;
        bucket = nt_get_bucket(0, c->first_bits, min_hash.v);
        if (rc) minimizer_pos = (uint16_t)(index + c->k - 1 - (size_t)min_hash.index - c->m);
        else minimizer_pos = (uint16_t)(min_hash.index - ((uint32_t)c->last_index - 1));
    }
    uint8_t second_bucket = (uint8_t)nt_get_bucket(c->first_bits, c->second_bits, min_hash.v);
    size_t s = c->last_index - 1, e = index + c->k - 1;
    if (c->cnt < c->cap) {
        orc_superkmer *o = &c->out[c->cnt];
        o->read_index = c->read_index;
        o->start = c->seg_start + (uint32_t)s;
        o->len = (uint32_t)(e - s);
        o->color = c->color;
        o->bucket = bucket;
        o->minimizer_pos = minimizer_pos;
        o->second_bucket = second_bucket;
        o->flags = (uint8_t)(((c->include_first ? 1 : 0) << (rc ? 1 : 0)) | ((include_last ? 1 : 0) << (rc ? 0 : 1)));
        o->rc = (uint8_t)rc;
        o->pad = 0;
    }
    c->cnt++;
    c->last_index = index;
    c->include_first = 0;
}

/* assembler_minimizer_bucketing/src/lib.rs:171-270  process_sequence::<_, _, SEPARATE_DUPLICATES=true>
 * on one N-free ASCII segment (include_first = include_last = true in phase 1, :149-150). */
static void process_sequence(ps_ctx *c, const uint8_t *seq, size_t len, int include_first, int include_last) {
    size_t n_items = len - c->m + 1;
    uint64_t *fw = (uint64_t *)malloc(sizeof(uint64_t) * n_items);
    uint64_t *rc = (uint64_t *)malloc(sizeof(uint64_t) * n_items);
    mq_item *items = (mq_item *)malloc(sizeof(mq_item) * n_items);
    orc_nthash_iter(seq, len, c->m, fw, rc);
    for (size_t i = 0; i < n_items; i++) {
        uint64_t mn = fw[i] < rc[i] ? fw[i] : rc[i];
        /* cn_nthash.rs:96-99 to_unextendable = min << 1 ; amb:200-203 | !is_rc_symmetric */
        items[i].v = (mn << 1) | (uint64_t)(fw[i] != rc[i]);
        items[i].index = (uint32_t)i;
        items[i].is_forward = fw[i] < rc[i];
    }
    c->last_index = 1;
    c->include_first = include_first;
    c->include_last_pre = include_last;
    size_t skip_before = include_first ? 0 : 1, skip_after = include_last ? 0 : 1;
    mq_get_minimizer_splits(c->k - c->m, items, n_items, skip_before, skip_after, ps_on_split, c);
    free(fw); free(rc); free(items);
}

/* Phase 1 over a batch of input records (minimizer_bucketing/src/lib.rs:310-376 hot loop):
 * records shorter than k are dropped (reader.rs:70-72), bytes normalised (sequences_reader.rs:50-54),
 * split at N, each segment >= k processed.  `reads` = concatenated raw ASCII, offsets[n_reads+1].
 * colors may be NULL.  Returns number of super-k-mers (may exceed cap: call again with more room). */
size_t orc_bucketing(const uint8_t *reads, const uint64_t *offsets, size_t n_reads, const uint32_t *colors, size_t k,
                     size_t m, unsigned b1, unsigned b2, int forward_only, orc_superkmer *out, size_t cap,
                     uint64_t *valid_bases_out) {
    ps_ctx c;
    memset(&c, 0, sizeof(c));
    c.out = out; c.cap = cap; c.k = k; c.m = m; c.first_bits = b1; c.second_bits = b2;
    c.canonical = !forward_only;
    c.duplicates_bucket = (uint16_t)(1u << b1);
    uint64_t valid_bases = 0;
    size_t maxlen = 0;
    for (size_t r = 0; r < n_reads; r++) { size_t l = offsets[r + 1] - offsets[r]; if (l > maxlen) maxlen = l; }
    uint8_t *buf = (uint8_t *)malloc(maxlen + 1);
    for (size_t r = 0; r < n_reads; r++) {
        size_t len = offsets[r + 1] - offsets[r];
        if (len < k) continue;
        memcpy(buf, reads + offsets[r], len);
        orc_normalize(buf, len);
        size_t start, end = 0;
        while (end < len) {
            start = end;
            while (start < len && (((buf[start] ^ 'N') & 0x7) == 0)) start++;
            end = start;
            while (end < len && (((buf[end] ^ 'N') & 0x7) != 0)) end++;
            if (end - start >= k) {
                valid_bases += end - start;
                c.read_index = (uint32_t)r;
                c.seg_start = (uint32_t)start;
                c.color = colors ? colors[r] : 0;
                process_sequence(&c, buf + start, end - start, 1, 1);
            }
        }
    }
    free(buf);
    if (valid_bases_out) *valid_bases_out = valid_bases;
    return c.cnt;
}

/* Stored packed bytes of one super-k-mer (creads_utils.rs:389-406: rc => compress_from_plain_rc). */
size_t orc_superkmer_packed(const uint8_t *reads, const uint64_t *offsets, const orc_superkmer *s, uint8_t *out) {
    size_t n = s->len;
    uint8_t *tmp = (uint8_t *)malloc(n + 1);
    memcpy(tmp, reads + offsets[s->read_index] + s->start, n);
    orc_normalize(tmp, n);
    size_t w = s->rc ? orc_compress_from_plain_rc(tmp, n, out) : orc_compress_from_plain(tmp, n, out);
    free(tmp);
    return w;
}

/* A.4 wire record: creads_utils.rs:374-434 write_to (BucketMode on, MultiplicityMode off,
 * MinimizerMode on, FlagsCount=2) + compressed_read.rs:435-467 encode_length.
 * Colour extra (single colour, colors/src/parsers/separate.rs:98-116) is appended by the caller. */
size_t orc_superkmer_record(const uint8_t *reads, const uint64_t *offsets, const orc_superkmer *s, size_t k, uint8_t *out) {
    size_t w = 0;
    out[w++] = s->second_bucket;
    unsigned min_size_log = 0;
    { size_t p = 1; while (p < k) { p <<= 1; min_size_log++; } } /* k.next_power_of_two().ilog2() */
    uint64_t encoded = (((uint64_t)(s->len - k)) << min_size_log) | (uint64_t)s->minimizer_pos;
    w += orc_encode_varint_flags(encoded, s->flags, 2, out + w);
    w += orc_superkmer_packed(reads, offsets, s, out + w);
    return w;
}

/* ------------------------------------------------------------------------------------------
 * A.5  k-mer identity hashes on a packed (2-bit) stored read
 * ---------------------------------------------------------------------------------------- */

/* hashes/src/cn_rkhash.rs:69-75 (u128 module) */
static const u128 RK_MULTIPLIER = ((u128)0x3eb9402f3e733993ULL << 64) | 0xadd64d3ca00e1b6bULL;
static const u128 RK_MULT_INV = ((u128)0x09cb6ff6f1b1a6d7ULL << 64) | 0x33e0952e899c3943ULL;
static const u128 RK_MULT_A = ((u128)0x4751137d01d863c5ULL << 64) | 0xb8c36de2b7d399dfULL;
static const u128 RK_MULT_C = ((u128)0x37ea3a13226503fbULL << 64) | 0x783f5cb69f4552bdULL;
static const u128 RK_MULT_G = ((u128)0x50796b285343f09aULL << 64) | 0x0c53113ae736572bULL;
static const u128 RK_MULT_T = ((u128)0x1e62d96a5e1f5adeULL << 64) | 0x2d4e68d8f88110b7ULL;

/* hashes/src/base/cn_rkhash_base.rs:10-44: compressed codes 0=A 1=C 2=T 3=G */
static inline u128 rk_fwd_l(uint8_t c) { return c == 0 ? RK_MULT_A : c == 1 ? RK_MULT_C : c == 2 ? RK_MULT_T : RK_MULT_G; }
static inline u128 rk_bkw_l(uint8_t c) { return c == 0 ? RK_MULT_T : c == 1 ? RK_MULT_G : c == 2 ? RK_MULT_A : RK_MULT_C; }

/* hashes/src/lib.rs:169-191  init_rmmult: fastexp(multiplier, k-1) */
static u128 rk_rmmult(size_t k) {
    u128 result = 1, sqv = RK_MULTIPLIER;
    size_t e = k ? k - 1 : 0;
    while (e > 0) { if (e & 1) result *= sqv; e /= 2; sqv *= sqv; }
    return result;
}
void orc_rk_constants(uint64_t *out /* 7 x (lo,hi) */, size_t k) {
    u128 v[7] = {RK_MULTIPLIER, RK_MULT_INV, RK_MULT_A, RK_MULT_C, RK_MULT_G, RK_MULT_T, rk_rmmult(k)};
    for (int i = 0; i < 7; i++) { out[2 * i] = (uint64_t)v[i]; out[2 * i + 1] = (uint64_t)(v[i] >> 64); }
}

typedef struct { u128 key; int is_forward; int symmetric; } kmer_hash;

/* Rolling k-mer hashes of a packed read, literal per hash family.  out has len-k+1 entries.
 *  seq-hash canonical : hashes/src/base/cn_seqhash_base.rs:22-70,100-116
 *  seq-hash forward   : hashes/src/base/fw_seqhash_base.rs (fh only; is_forward = true)
 *  rk128 canonical    : hashes/src/base/cn_rkhash_base.rs:64-110,144-160
 *  rk128 forward      : hashes/src/base/fw_rkhash_base.rs:57-69 */
static void kmer_hashes_packed(const uint8_t *packed, size_t start_base, size_t len, size_t k, int hash_type,
                               int forward_only, kmer_hash *out) {
    size_t cnt = len - k + 1;
    if (hash_type == ORC_HASH_SEQ) {
        /* width: reference picks u16/u32/u64/u128 by k (api/src/utils.rs:30-58); values are identical
         * as integers, so one u128 restatement with mask = 2k low bits covers all widths. */
        u128 mask = (k >= 64) ? ~(u128)0 : ((((u128)1) << (2 * k)) - 1);
        u128 fh = 0, bw = 0;
        for (size_t i = 0; i < k - 1; i++) {
            u128 base = packed_base(packed, start_base + i);
            fh |= base << (i * 2);
            bw = (bw << 2) | (base ^ 2);
        }
        fh <<= 2;
        bw &= mask;
        for (size_t idx = k - 1; idx < len; idx++) {
            u128 base = packed_base(packed, start_base + idx);
            fh = (fh >> 2) | (base << ((k - 1) * 2));
            bw = ((bw << 2) | (base ^ 2)) & mask;
            kmer_hash *o = &out[idx - (k - 1)];
            if (forward_only) { o->key = fh; o->is_forward = 1; o->symmetric = 0; }
            else { o->key = fh < bw ? fh : bw; o->is_forward = fh < bw; o->symmetric = fh == bw; }
        }
    } else {
        u128 rmmult = rk_rmmult(k);
        u128 fh = 0, bw = 0;
        for (size_t i = 0; i < k - 1; i++) fh = fh * RK_MULTIPLIER + rk_fwd_l(packed_base(packed, start_base + i));
        for (size_t i = k - 1; i-- > 0;) bw = bw * RK_MULTIPLIER + rk_bkw_l(packed_base(packed, start_base + i));
        bw *= RK_MULTIPLIER;
        for (size_t idx = k - 1; idx < len; idx++) {
            uint8_t in_base = packed_base(packed, start_base + idx);
            uint8_t out_base = packed_base(packed, start_base + idx - (k - 1));
            u128 current_fh = fh * RK_MULTIPLIER + rk_fwd_l(in_base);
            fh = current_fh - rk_fwd_l(out_base) * rmmult;
            u128 current_bk = bw * RK_MULT_INV + rk_bkw_l(in_base) * rmmult;
            bw = current_bk - rk_bkw_l(out_base);
            kmer_hash *o = &out[idx - (k - 1)];
            if (forward_only) { o->key = current_fh; o->is_forward = 1; o->symmetric = 0; }
            else { o->key = current_fh < current_bk ? current_fh : current_bk; o->is_forward = current_fh < current_bk; o->symmetric = current_fh == current_bk; }
        }
    }
    (void)cnt;
}

/* Test hook: hashes of an ASCII sequence (packed internally).  key_lo/key_hi/is_fw have n-k+1 entries. */
size_t orc_kmer_hashes(const uint8_t *ascii, size_t n, size_t k, int hash_type, int forward_only, uint64_t *key_lo,
                       uint64_t *key_hi, uint8_t *is_fw) {
    if (n < k) return 0;
    uint8_t *packed = (uint8_t *)calloc((n + 3) / 4 + 1, 1);
    orc_compress_from_plain(ascii, n, packed);
    size_t cnt = n - k + 1;
    kmer_hash *h = (kmer_hash *)malloc(sizeof(kmer_hash) * cnt);
    kmer_hashes_packed(packed, 0, n, k, hash_type, forward_only, h);
    for (size_t i = 0; i < cnt; i++) { key_lo[i] = (uint64_t)h[i].key; key_hi[i] = (uint64_t)(h[i].key >> 64); is_fw[i] = (uint8_t)h[i].is_forward; }
    free(h); free(packed);
    return cnt;
}

/* hashes/src/base/cn_seqhash_base.rs:140-150 + cn_seqhash.rs:22-26 (u64): partition function used when
 * routing partial unitigs (consumer side; restated for the "next" row) */
uint16_t orc_seqhash64_get_bucket(unsigned used_bits, unsigned requested_bits, uint64_t hash) {
    uint64_t h = hash * 0x00000100000001b3ULL + 0xcbf29ce484222325ULL;
    h = rotr64(h, 3);
    return (uint16_t)((h >> used_bits) % (1ULL << requested_bits));
}

/* ------------------------------------------------------------------------------------------
 * A.6  k-mer table of one merge unit: HashMapUnitigsExtender::add_sequence + MapEntry
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    u128 key;
    uint64_t counter;     /* MapEntry counter (structs/src/map_entry.rs:33-45) */
    uint8_t flags;        /* MapEntry flags   (structs/src/map_entry.rs:68-77) */
    uint8_t used;
    uint32_t *colors;     /* multiset of colours, one per occurrence */
    uint32_t ncolors, capcolors;
} tbl_entry;

typedef struct {
    tbl_entry *slots;
    size_t cap, len;
} tbl;

static uint64_t tbl_mix(u128 key) {
    uint64_t x = (uint64_t)key ^ ((uint64_t)(key >> 64) * 0x9E3779B97F4A7C15ULL);
    x ^= x >> 31; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 29;
    return x;
}
static void tbl_init(tbl *t, size_t cap) {
    size_t c = 64; while (c < cap * 2) c <<= 1;
    t->slots = (tbl_entry *)calloc(c, sizeof(tbl_entry)); t->cap = c; t->len = 0;
}
static tbl_entry *tbl_get(tbl *t, u128 key);
static void tbl_grow(tbl *t) {
    tbl old = *t;
    t->cap = old.cap * 2; t->slots = (tbl_entry *)calloc(t->cap, sizeof(tbl_entry)); t->len = 0;
    for (size_t i = 0; i < old.cap; i++) if (old.slots[i].used) { tbl_entry *e = tbl_get(t, old.slots[i].key); *e = old.slots[i]; }
    free(old.slots);
}
static tbl_entry *tbl_get(tbl *t, u128 key) {
    if ((t->len + 1) * 2 > t->cap) tbl_grow(t);
    size_t mask = t->cap - 1, i = (size_t)tbl_mix(key) & mask;
    for (;;) {
        tbl_entry *e = &t->slots[i];
        if (!e->used) { memset(e, 0, sizeof(*e)); e->used = 1; e->key = key; t->len++; return e; }
        if (e->key == key) return e;
        i = (i + 1) & mask;
    }
}

/* assembler_kmers_merge/src/unitigs_extender/hashmap.rs:361-409  add_sequence (counting part) */
static void add_sequence(tbl *t, const uint8_t *packed, size_t len, uint8_t flags, uint32_t multiplicity, int with_color,
                         uint32_t color, size_t k, int hash_type, int forward_only, kmer_hash *scratch) {
    kmer_hashes_packed(packed, 0, len, k, hash_type, forward_only, scratch);
    size_t last_hash_pos = len - k;
    for (size_t idx = 0; idx <= last_hash_pos; idx++) {
        int begin_ignored = ((flags & READ_FLAG_INCL_BEGIN) == 0) && idx == 0;
        int end_ignored = ((flags & READ_FLAG_INCL_END) == 0) && idx == last_hash_pos;
        int is_forward = scratch[idx].is_forward;
        tbl_entry *e = tbl_get(t, scratch[idx].key);
        e->flags |= (uint8_t)((begin_ignored << (!is_forward)) | (end_ignored << is_forward)); /* update_flags */
        e->counter += multiplicity;                                                              /* incr_by_and_check */
        if (with_color) {
            if (e->ncolors == e->capcolors) {
                e->capcolors = e->capcolors ? e->capcolors * 2 : 4;
                e->colors = (uint32_t *)realloc(e->colors, sizeof(uint32_t) * e->capcolors);
            }
            e->colors[e->ncolors++] = color;
        }
    }
}

typedef struct {
    uint64_t key_lo, key_hi;
    uint64_t counter;       /* raw counter */
    uint64_t multiplicity;  /* get_kmer_multiplicity(): counter >> (flags == 3)  (map_entry.rs:79-84) */
    uint32_t color_off;     /* offset into colour array (sorted-unique list) */
    uint32_t color_len;
    uint8_t flags;
    uint8_t kept;           /* multiplicity >= min_multiplicity (hashmap.rs:77-80,134-137) */
    uint8_t pad[6];
} orc_table_entry;

static int cmp_entry(const void *a, const void *b) {
    const orc_table_entry *x = (const orc_table_entry *)a, *y = (const orc_table_entry *)b;
    if (x->key_hi != y->key_hi) return x->key_hi < y->key_hi ? -1 : 1;
    if (x->key_lo != y->key_lo) return x->key_lo < y->key_lo ? -1 : 1;
    return 0;
}
static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

/* Build the table of one merge unit from super-k-mers selected by (bucket, second_bucket);
 * second_bucket < 0 selects the whole first-level bucket (fold of its sub-buckets).
 * Entries come back sorted ascending by key (u128 order).  colours_out receives the concatenated
 * sorted-unique colour lists (colors/src/managers/multiple.rs:198-203) when with_color.
 * Returns number of distinct keys (may exceed cap; colours may exceed colors_cap via *n_colors_out). */
size_t orc_merge_unit(const uint8_t *reads, const uint64_t *offsets, const orc_superkmer *sk, size_t n_sk, int bucket,
                      int second_bucket, size_t k, uint64_t min_multiplicity, int hash_type, int forward_only,
                      int with_color, orc_table_entry *out, size_t cap, uint32_t *colors_out, size_t colors_cap,
                      uint64_t *n_colors_out, uint64_t *total_kmers_out) {
    tbl t;
    tbl_init(&t, 1024);
    size_t maxlen = 0;
    for (size_t i = 0; i < n_sk; i++) if (sk[i].len > maxlen) maxlen = sk[i].len;
    uint8_t *packed = (uint8_t *)malloc(maxlen / 4 + 8);
    kmer_hash *scratch = (kmer_hash *)malloc(sizeof(kmer_hash) * (maxlen + 1));
    uint64_t total_kmers = 0;
    for (size_t i = 0; i < n_sk; i++) {
        if ((int)sk[i].bucket != bucket) continue;
        if (second_bucket >= 0 && (int)sk[i].second_bucket != second_bucket) continue;
        orc_superkmer_packed(reads, offsets, &sk[i], packed);
        total_kmers += sk[i].len - k + 1;
        add_sequence(&t, packed, sk[i].len, sk[i].flags, 1, with_color, sk[i].color, k, hash_type, forward_only, scratch);
    }
    size_t n = 0;
    orc_table_entry *all = (orc_table_entry *)malloc(sizeof(orc_table_entry) * (t.len ? t.len : 1));
    tbl_entry **src = (tbl_entry **)malloc(sizeof(tbl_entry *) * (t.len ? t.len : 1));
    for (size_t i = 0; i < t.cap; i++) {
        tbl_entry *e = &t.slots[i];
        if (!e->used) continue;
        orc_table_entry *o = &all[n];
        memset(o, 0, sizeof(*o));
        o->key_lo = (uint64_t)e->key; o->key_hi = (uint64_t)(e->key >> 64);
        o->counter = e->counter;
        o->flags = e->flags;
        o->multiplicity = e->counter >> (e->flags == (READ_FLAG_INCL_BEGIN | READ_FLAG_INCL_END));
        o->kept = o->multiplicity >= min_multiplicity;
        o->color_off = (uint32_t)i; /* temporarily: slot index */
        n++;
    }
    qsort(all, n, sizeof(orc_table_entry), cmp_entry);
    uint64_t ncol = 0;
    for (size_t i = 0; i < n; i++) {
        tbl_entry *e = &t.slots[all[i].color_off];
        all[i].color_off = (uint32_t)ncol; all[i].color_len = 0;
        if (with_color && e->ncolors) {
            qsort(e->colors, e->ncolors, sizeof(uint32_t), cmp_u32);
            uint32_t u = 0;
            for (uint32_t j = 0; j < e->ncolors; j++) if (j == 0 || e->colors[j] != e->colors[j - 1]) {
                if (ncol + u < colors_cap) colors_out[ncol + u] = e->colors[j];
                u++;
            }
            all[i].color_len = u; ncol += u;
        }
        if (i < cap) out[i] = all[i];
    }
    for (size_t i = 0; i < t.cap; i++) if (t.slots[i].used) free(t.slots[i].colors);
    free(t.slots); free(all); free(src); free(packed); free(scratch);
    if (n_colors_out) *n_colors_out = ncol;
    if (total_kmers_out) *total_kmers_out = total_kmers;
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Independent cross-check: count every k-mer of the (normalised, N-split, >= k) input directly
 * from the definitions in SURVEY A.5 -- no minimizers, no buckets, no rolling.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t key_lo, key_hi, count; } orc_naive_entry;

static u128 naive_key(const uint8_t *s, size_t k, int hash_type, int forward_only) {
    u128 fw = 0, rc = 0;
    if (hash_type == ORC_HASH_SEQ) {
        for (size_t i = 0; i < k; i++) {
            u128 b = compress_base(s[i]);
            fw += b << (2 * i);
            rc += (b ^ 2) << (2 * (k - 1 - i));
        }
    } else {
        u128 p = 1; /* M^i */
        u128 *pw = (u128 *)malloc(sizeof(u128) * k);
        for (size_t i = 0; i < k; i++) { pw[i] = p; p *= RK_MULTIPLIER; }
        for (size_t i = 0; i < k; i++) {
            uint8_t b = compress_base(s[i]);
            fw += rk_fwd_l(b) * pw[k - 1 - i];
            rc += rk_bkw_l(b) * pw[i];
        }
        free(pw);
    }
    if (forward_only) return fw;
    return fw < rc ? fw : rc;
}
static int cmp_u128(const void *a, const void *b) {
    u128 x = *(const u128 *)a, y = *(const u128 *)b;
    return x < y ? -1 : x > y;
}
size_t orc_naive_count(const uint8_t *reads, const uint64_t *offsets, size_t n_reads, size_t k, int hash_type,
                       int forward_only, orc_naive_entry *out, size_t cap, uint64_t *total_out) {
    size_t total = 0, maxlen = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t l = offsets[r + 1] - offsets[r];
        if (l > maxlen) maxlen = l;
        if (l >= k) total += l - k + 1;
    }
    u128 *keys = (u128 *)malloc(sizeof(u128) * (total ? total : 1));
    uint8_t *buf = (uint8_t *)malloc(maxlen + 1);
    size_t n = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t len = offsets[r + 1] - offsets[r];
        if (len < k) continue;
        memcpy(buf, reads + offsets[r], len);
        orc_normalize(buf, len);
        for (size_t i = 0; i + k <= len; i++) {
            int ok = 1;
            for (size_t j = 0; j < k; j++) if (buf[i + j] == 'N') { ok = 0; break; }
            if (ok) keys[n++] = naive_key(buf + i, k, hash_type, forward_only);
        }
    }
    qsort(keys, n, sizeof(u128), cmp_u128);
    size_t d = 0;
    for (size_t i = 0; i < n;) {
        size_t j = i;
        while (j < n && keys[j] == keys[i]) j++;
        if (d < cap) { out[d].key_lo = (uint64_t)keys[i]; out[d].key_hi = (uint64_t)(keys[i] >> 64); out[d].count = j - i; }
        d++; i = j;
    }
    free(keys); free(buf);
    if (total_out) *total_out = n;
    return d;
}

/* io/src/lib.rs:67-140 compute_stats_from_input_blocks (bucket-count heuristics); config/src/lib.rs:62-90 */
static uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }
static uint64_t u64min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static uint64_t u64max(uint64_t a, uint64_t b) { return a > b ? a : b; }
static unsigned ilog2u(uint64_t x) { unsigned l = 0; while (x >>= 1) l++; return l; }
void orc_bucket_counts(uint64_t bases_count, unsigned *buckets_log, unsigned *second_log) {
    const uint64_t MAX_BUCKET_SIZE = 1024ULL * 1024 * 1024, MIN_BUCKET_SIZE = 512 * 1024;
    uint64_t buckets = u64max(u64min(1ULL << 10, bases_count / MIN_BUCKET_SIZE), bases_count / MAX_BUCKET_SIZE);
    buckets = next_pow2(buckets);
    buckets = u64max(u64min(buckets, 1ULL << 13), 1ULL << 2);
    uint64_t per = bases_count / buckets;
    const uint64_t MAX_SECOND = 4ULL * 1024 * 1024, MIN_SECOND = 2 * 1024;
    uint64_t second = u64max(u64min(1ULL << 6, per / MIN_SECOND), per / MAX_SECOND);
    second = next_pow2(second);
    second = u64max(u64min(second, 1ULL << 8), 1ULL << 1);
    *buckets_log = ilog2u(buckets);
    *second_log = ilog2u(second);
}

/* utils/src/lib.rs:29-40 compute_best_m */
size_t orc_compute_best_m(size_t k) {
    if (k <= 13) return (k / 2) > (k >= 4 ? k - 4 : 0) ? (k / 2) : (k >= 4 ? k - 4 : 0);
    if (k <= 15) return 9;
    if (k <= 21) return 10;
    if (k <= 30) return 11;
    if (k <= 37) return 12;
    if (k <= 42) return 13;
    if (k <= 64) return 14;
    return (size_t)((double)k / 4.0 + 0.5);
}

size_t orc_sizeof_superkmer(void) { return sizeof(orc_superkmer); }
size_t orc_sizeof_table_entry(void) { return sizeof(orc_table_entry); }

/* ------------------------------------------------------------------------------------------
 * Whole-path driver used as the CPU baseline (bench.py cpu_baseline / --impl reference):
 * phase 1 over all records, then phase 2 over every (bucket, second_bucket) unit, OpenMP over
 * records / units the way the reference parallelises over input blocks and bucket jobs
 * (minimizer_bucketing/src/lib.rs:566-567; kmers_transform/src/lib.rs:332-370).
 * ---------------------------------------------------------------------------------------- */
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    uint64_t n_superkmers, n_kmers, n_unique, n_kept, valid_bases, checksum;
    double t_bucketing, t_merge;
    int threads;
} orc_pipeline_stats;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int orc_pipeline(const uint8_t *reads, const uint64_t *offsets, size_t n_reads, size_t k, size_t m, unsigned b1,
                 unsigned b2, int forward_only, uint64_t min_multiplicity, int hash_type, int n_threads,
                 orc_pipeline_stats *st) {
    memset(st, 0, sizeof(*st));
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    st->threads = omp_get_max_threads();
#else
    st->threads = 1;
#endif
    const size_t BLOCK = 2048;
    size_t n_blocks = (n_reads + BLOCK - 1) / BLOCK;
    orc_superkmer **blk_sk = (orc_superkmer **)calloc(n_blocks ? n_blocks : 1, sizeof(orc_superkmer *));
    size_t *blk_n = (size_t *)calloc(n_blocks ? n_blocks : 1, sizeof(size_t));
    uint64_t *blk_vb = (uint64_t *)calloc(n_blocks ? n_blocks : 1, sizeof(uint64_t));
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 1)
    for (long bi = 0; bi < (long)n_blocks; bi++) {
        size_t r0 = (size_t)bi * BLOCK, r1 = r0 + BLOCK < n_reads ? r0 + BLOCK : n_reads;
        size_t bases = offsets[r1] - offsets[r0];
        size_t cap = bases / 4 + (r1 - r0) + 16;
        orc_superkmer *out = (orc_superkmer *)malloc(sizeof(orc_superkmer) * cap);
        uint64_t vb = 0;
        /* offsets are absolute: orc_bucketing indexes reads[offsets[r]] */
        size_t n = orc_bucketing(reads, offsets + r0, r1 - r0, NULL, k, m, b1, b2, forward_only, out, cap, &vb);
        if (n > cap) {
            out = (orc_superkmer *)realloc(out, sizeof(orc_superkmer) * n);
            n = orc_bucketing(reads, offsets + r0, r1 - r0, NULL, k, m, b1, b2, forward_only, out, n, &vb);
        }
        for (size_t i = 0; i < n; i++) out[i].read_index += (uint32_t)r0;
        blk_sk[bi] = out; blk_n[bi] = n; blk_vb[bi] = vb;
    }
    size_t n_sk = 0;
    for (size_t b = 0; b < n_blocks; b++) { n_sk += blk_n[b]; st->valid_bases += blk_vb[b]; }
    /* "write to bucket": counting sort of the super-k-mers by unit */
    size_t n_units = (((size_t)1 << b1) + 1) << b2;
    size_t *unit_off = (size_t *)calloc(n_units + 1, sizeof(size_t));
    for (size_t b = 0; b < n_blocks; b++)
        for (size_t i = 0; i < blk_n[b]; i++) unit_off[(((size_t)blk_sk[b][i].bucket << b2) | blk_sk[b][i].second_bucket) + 1]++;
    for (size_t u = 0; u < n_units; u++) unit_off[u + 1] += unit_off[u];
    orc_superkmer *sorted = (orc_superkmer *)malloc(sizeof(orc_superkmer) * (n_sk ? n_sk : 1));
    size_t *cursor = (size_t *)malloc(sizeof(size_t) * (n_units + 1));
    memcpy(cursor, unit_off, sizeof(size_t) * (n_units + 1));
    for (size_t b = 0; b < n_blocks; b++) {
        for (size_t i = 0; i < blk_n[b]; i++) {
            size_t u = ((size_t)blk_sk[b][i].bucket << b2) | blk_sk[b][i].second_bucket;
            sorted[cursor[u]++] = blk_sk[b][i];
        }
        free(blk_sk[b]);
    }
    double t1 = now_s();
    uint64_t n_kmers = 0, n_unique = 0, n_kept = 0, checksum = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : n_kmers, n_unique, n_kept, checksum)
    for (long u = 0; u < (long)n_units; u++) {
        size_t a = unit_off[u], e = unit_off[u + 1];
        if (a == e) continue;
        tbl t;
        tbl_init(&t, 1024);
        size_t maxlen = 0;
        for (size_t i = a; i < e; i++) if (sorted[i].len > maxlen) maxlen = sorted[i].len;
        uint8_t *packed = (uint8_t *)malloc(maxlen / 4 + 8);
        kmer_hash *scratch = (kmer_hash *)malloc(sizeof(kmer_hash) * (maxlen + 1));
        for (size_t i = a; i < e; i++) {
            orc_superkmer_packed(reads, offsets, &sorted[i], packed);
            n_kmers += sorted[i].len - k + 1;
            add_sequence(&t, packed, sorted[i].len, sorted[i].flags, 1, 0, 0, k, hash_type, forward_only, scratch);
        }
        for (size_t i = 0; i < t.cap; i++) {
            tbl_entry *en = &t.slots[i];
            if (!en->used) continue;
            n_unique++;
            uint64_t mult = en->counter >> (en->flags == 3);
            if (mult >= min_multiplicity) { n_kept++; checksum += (uint64_t)en->key * 0x9E3779B97F4A7C15ULL + mult + ((uint64_t)en->flags << 40); }
        }
        free(t.slots); free(packed); free(scratch);
    }
    double t2 = now_s();
    st->n_superkmers = n_sk; st->n_kmers = n_kmers; st->n_unique = n_unique; st->n_kept = n_kept; st->checksum = checksum;
    st->t_bucketing = t1 - t0; st->t_merge = t2 - t1;
    free(blk_sk); free(blk_n); free(blk_vb); free(unit_off); free(sorted); free(cursor);
    return 0;
}
