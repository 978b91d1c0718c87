/*
 * ggcat_unitigs.c -- CPU restatement of the CONSUMER of the k-mer tables: partial unitigs per merge unit
 * and their joining into maximal unitigs.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as ggcat_oracle.c): used by tests/ to run the north-star's second
 * bit-exactness check ("final maximal-unitig set compared as canonical k-mer set, plus its unitig count")
 * on tables produced by the CUDA path.  The product never loads it.
 *
 * Restates, for invertible seq-hash keys (k <= 63 here):
 *   assembler_kmers_merge/src/unitigs_extender/hashmap.rs:65-86    get_kmers (iterate kept, unused entries)
 *   assembler_kmers_merge/src/unitigs_extender/hashmap.rs:162-297  try_extend_function (COMPUTE_SIMPLITIGS = false)
 *   assembler_kmers_merge/src/unitigs_extender/hashmap.rs:442-601  compute_unitigs
 * The reference then routes open-ended partial unitigs by the hash of their end k-mer and joins them in
 * phases 3-5 (assembler/src/lib.rs, assembler_pipeline/src/{links_compaction,build_unitigs,extend_unitigs}.rs);
 * those phases are out of the hot path, so the join is restated at its SEMANTIC level only:
 * two partial unitigs that end in the same k-mer X (stored once per unit, with complementary flags) are
 * glued with X shared.  PARITY UNPINNED by reference goldens (the reference has no unitig fixtures); the
 * independent cross-check is the flag-free global build (one unit holding every kept k-mer).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct {
    u128 key;       /* canonical (or forward, -f) k-mer: base i at bits 2i (cn_seqhash_base.rs:27-50) */
    uint8_t flags;  /* MapEntry flags relative to the canonical orientation (hashmap.rs:385-399) */
    uint8_t used;   /* MapEntry USED bit (structs/src/map_entry.rs:47-66) */
} ut_node;

typedef struct {
    const ut_node *t;  /* one unit's kept entries, ascending by key */
    ut_node *tm;
    long n;
    unsigned k;
    int forward_only;
    u128 mask;
} ut_table;

static u128 ut_rc(u128 x, unsigned k) {
    u128 r = 0;
    for (unsigned i = 0; i < k; i++) { r = (r << 2) | ((x & 3) ^ 2); x >>= 2; }
    return r;
}
static inline u128 ut_canon(const ut_table *T, u128 s) {
    if (T->forward_only) return s;
    u128 r = ut_rc(s, T->k);
    return s < r ? s : r;
}
static long ut_find(const ut_table *T, u128 key) {
    long lo = 0, hi = T->n - 1;
    while (lo <= hi) {
        long mid = (lo + hi) >> 1;
        if (T->t[mid].key == key) return mid;
        if (T->t[mid].key < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}
/* cn_seqhash_base.rs manual_roll_forward / manual_roll_reverse on the oriented k-mer */
static inline u128 ut_succ(const ut_table *T, u128 s, unsigned b) { return (s >> 2) | ((u128)b << (2 * (T->k - 1))); }
static inline u128 ut_pred(const ut_table *T, u128 s, unsigned b) { return ((s << 2) | b) & T->mask; }

typedef struct { uint8_t *b; size_t n, cap; } ut_seq;
static void ut_push(ut_seq *s, uint8_t c) {
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 256; s->b = (uint8_t *)realloc(s->b, s->cap); }
    s->b[s->n++] = c;
}

/* hashmap.rs:162-297 try_extend_function, unitigs (non-simplitig) branch.  dir = +1: roll forward,
 * -1: roll reverse.  Appends the extension bases to out.  Returns 1 ("Some": open end, continues in another
 * unit) or 0 ("None"). */
static int ut_try_extend(ut_table *T, u128 start, int dir, ut_seq *out) {
    u128 cur = start;
    for (;;) {
        int count = 0;
        u128 cand = 0;
        unsigned cb = 0;
        for (unsigned b = 0; b < 4; b++) {
            u128 nh = dir > 0 ? ut_succ(T, cur, b) : ut_pred(T, cur, b);
            if (ut_find(T, ut_canon(T, nh)) >= 0) { count++; cand = nh; cb = b; }
        }
        if (count != 1) return 0;
        int ocount = 0;
        for (unsigned b = 0; b < 4; b++) {
            u128 bh = dir > 0 ? ut_pred(T, cand, b) : ut_succ(T, cand, b);
            if (ut_find(T, ut_canon(T, bh)) >= 0) { if (ocount > 0) return 0; ocount++; }
        }
        long e = ut_find(T, ut_canon(T, cand));
        if (T->tm[e].used) return 0;
        T->tm[e].used = 1;
        ut_push(out, (uint8_t)cb);
        if (T->t[e].flags == 1 || T->t[e].flags == 2) return 1; /* contig_break */
        cur = cand;
    }
}

typedef struct {
    uint8_t *bases; /* 2-bit codes, one per byte */
    size_t len;
    int open_fw, open_bw;
} ut_partial;

typedef struct { ut_partial *v; size_t n, cap; } ut_plist;
static void ut_emit(ut_plist *L, const uint8_t *b, size_t len, int open_fw, int open_bw) {
    if (L->n == L->cap) { L->cap = L->cap ? L->cap * 2 : 1024; L->v = (ut_partial *)realloc(L->v, L->cap * sizeof(ut_partial)); }
    ut_partial *p = &L->v[L->n++];
    p->bases = (uint8_t *)malloc(len ? len : 1);
    memcpy(p->bases, b, len);
    p->len = len; p->open_fw = open_fw; p->open_bw = open_bw;
}

/* hashmap.rs:442-601 compute_unitigs over one unit */
static void ut_compute_unit(ut_table *T, ut_plist *L) {
    ut_seq fw = {0}, bw = {0};
    for (long i = 0; i < T->n; i++) {
        if (T->tm[i].used) continue;
        const u128 hash = T->t[i].key; /* MH::new(invert(hash)): the canonical k-mer read forward => is_forward */
        const uint8_t st = T->t[i].flags;
        const int begin_ignored = st == 1, end_ignored = st == 2;
        fw.n = bw.n = 0;
        T->tm[i].used = 1;
        int open_fw = end_ignored ? 1 : ut_try_extend(T, hash, +1, &fw);
        int open_bw = begin_ignored ? 1 : ut_try_extend(T, hash, -1, &bw);
        /* out_seq = reverse(backward extension) + k-mer + forward extension */
        size_t len = bw.n + T->k + fw.n;
        uint8_t *s = (uint8_t *)malloc(len);
        for (size_t j = 0; j < bw.n; j++) s[j] = bw.b[bw.n - 1 - j];
        for (unsigned j = 0; j < T->k; j++) s[bw.n + j] = (uint8_t)((hash >> (2 * j)) & 3);
        memcpy(s + bw.n + T->k, fw.b, fw.n);
        ut_emit(L, s, len, open_fw, open_bw);
        free(s);
    }
    free(fw.b); free(bw.b);
}

/* ---- join of open ends (semantic restatement of phases 3-5, see header) ---- */
typedef struct { u128 key; uint32_t unitig; uint8_t end; /* 0 = begin, 1 = end */ } ut_end;
static int ut_end_cmp(const void *a, const void *b) {
    const ut_end *x = (const ut_end *)a, *y = (const ut_end *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->unitig != y->unitig) return x->unitig < y->unitig ? -1 : 1;
    return (int)x->end - (int)y->end;
}
static u128 ut_kmer_at(const uint8_t *b, size_t pos, unsigned k) {
    u128 v = 0;
    for (unsigned j = 0; j < k; j++) v |= (u128)b[pos + j] << (2 * j);
    return v;
}
static int ut_u64_cmp(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
static int ut_u128_cmp(const void *a, const void *b) {
    u128 x = *(const u128 *)a, y = *(const u128 *)b;
    return x < y ? -1 : x > y;
}

/*
 * Tables of n_units units (entries of unit u = [unit_off[u], unit_off[u+1]), ascending by key inside a unit;
 * count_flags = multiplicity | flags << 30 as in include/ggcat_b200.h) -> maximal unitigs.
 * Outputs: *n_unitigs, *n_partial, *n_kmers (k-mers over all maximal unitigs), lengths[] (sorted, cap entries),
 * kmers_lo/hi[] = sorted canonical k-mers of all unitigs (cap_k entries).  Returns 0, or <0 on inconsistency
 * (-1: an open end without exactly one partner, -2: a k-mer appears twice in the output).
 */
int orc_unitigs_from_tables(const uint64_t *keys_lo, const uint64_t *keys_hi, const uint32_t *count_flags,
                            const uint64_t *unit_off, size_t n_units, unsigned k, int forward_only, uint64_t *n_unitigs,
                            uint64_t *n_partial, uint64_t *n_kmers, uint64_t *lengths, size_t cap_len, uint64_t *kmers_lo,
                            uint64_t *kmers_hi, size_t cap_k) {
    ut_plist L = {0};
    const u128 mask = (k >= 64) ? ~(u128)0 : ((((u128)1) << (2 * k)) - 1);
    for (size_t u = 0; u < n_units; u++) {
        const size_t a = unit_off[u], e = unit_off[u + 1];
        if (a == e) continue;
        ut_node *nodes = (ut_node *)malloc(sizeof(ut_node) * (e - a));
        for (size_t i = a; i < e; i++) {
            nodes[i - a].key = ((u128)(keys_hi ? keys_hi[i] : 0) << 64) | keys_lo[i];
            nodes[i - a].flags = (uint8_t)(count_flags[i] >> 30);
            nodes[i - a].used = 0;
        }
        ut_table T = {nodes, nodes, (long)(e - a), k, forward_only, mask};
        ut_compute_unit(&T, &L);
        free(nodes);
    }
    *n_partial = L.n;
    /* open ends keyed by canonical end k-mer */
    size_t n_ends = 0;
    for (size_t i = 0; i < L.n; i++) n_ends += (size_t)L.v[i].open_fw + (size_t)L.v[i].open_bw;
    ut_end *ends = (ut_end *)malloc(sizeof(ut_end) * (n_ends ? n_ends : 1));
    ut_table TC = {NULL, NULL, 0, k, forward_only, mask};
    size_t ne = 0;
    for (size_t i = 0; i < L.n; i++) {
        if (L.v[i].open_bw) { ends[ne].key = ut_canon(&TC, ut_kmer_at(L.v[i].bases, 0, k)); ends[ne].unitig = (uint32_t)i; ends[ne++].end = 0; }
        if (L.v[i].open_fw) { ends[ne].key = ut_canon(&TC, ut_kmer_at(L.v[i].bases, L.v[i].len - k, k)); ends[ne].unitig = (uint32_t)i; ends[ne++].end = 1; }
    }
    qsort(ends, ne, sizeof(ut_end), ut_end_cmp);
    /* partner[2*i + end] = 2*j + end' or -1 */
    long *partner = (long *)malloc(sizeof(long) * (2 * L.n + 2));
    for (size_t i = 0; i < 2 * L.n; i++) partner[i] = -1;
    int rc = 0;
    for (size_t i = 0; i < ne;) {
        size_t j = i;
        while (j < ne && ends[j].key == ends[i].key) j++;
        if (j - i != 2) { rc = -1; break; }
        long a = 2 * (long)ends[i].unitig + ends[i].end, b = 2 * (long)ends[i + 1].unitig + ends[i + 1].end;
        partner[a] = b; partner[b] = a;
        i = j;
    }
    uint64_t nu = 0, nk = 0;
    size_t kcount = 0;
    u128 *allk = (u128 *)malloc(sizeof(u128) * (cap_k ? cap_k : 1));
    uint8_t *seen = (uint8_t *)calloc(L.n ? L.n : 1, 1);
    for (int pass = 0; pass < 2 && rc == 0; pass++) {
        /* pass 0: chains starting at a closed end; pass 1: what is left are cycles of partial unitigs */
        for (size_t i = 0; i < L.n; i++) {
            if (seen[i]) continue;
            int start_end;
            if (pass == 0) {
                if (L.v[i].open_bw && L.v[i].open_fw) continue;
                start_end = L.v[i].open_bw ? 1 : 0; /* enter through the closed side */
            } else start_end = 0;
            /* walk: enter partial unitig `cur` at end `in_end`, leave through the other end */
            long cur = (long)i;
            int in_end = start_end;
            uint64_t len = 0;
            int first = 1;
            for (;;) {
                seen[cur] = 1;
                const ut_partial *p = &L.v[cur];
                /* k-mers of this piece; the k-mer shared with the previous piece is skipped */
                const size_t nkm = p->len - k + 1;
                for (size_t q = 0; q < nkm; q++) {
                    const size_t pos = in_end == 0 ? q : nkm - 1 - q;
                    if (!first && q == 0) continue;
                    if (kcount < cap_k) allk[kcount] = ut_canon(&TC, ut_kmer_at(p->bases, pos, k));
                    kcount++;
                }
                len += first ? p->len : p->len - k;
                first = 0;
                const int out_end = in_end ^ 1;
                const long nx = partner[2 * cur + out_end];
                if (nx < 0) break;
                if (seen[nx >> 1]) { /* closed a cycle: the first piece's entry k-mer was counted twice */
                    if (pass == 1 && (nx >> 1) == (long)i) { kcount--; len -= 1; }
                    break;
                }
                cur = nx >> 1;
                in_end = (int)(nx & 1);
            }
            if (nu < cap_len) lengths[nu] = len;
            nu++;
            nk += len - k + 1;
        }
    }
    if (rc == 0) {
        qsort(lengths, nu < cap_len ? nu : cap_len, sizeof(uint64_t), ut_u64_cmp);
        const size_t kc = kcount < cap_k ? kcount : cap_k;
        qsort(allk, kc, sizeof(u128), ut_u128_cmp);
        for (size_t i = 0; i < kc; i++) {
            if (i && allk[i] == allk[i - 1]) rc = -2;
            kmers_lo[i] = (uint64_t)allk[i]; kmers_hi[i] = (uint64_t)(allk[i] >> 64);
        }
    }
    *n_unitigs = nu; *n_kmers = kcount;
    (void)nk;
    for (size_t i = 0; i < L.n; i++) free(L.v[i].bases);
    free(L.v); free(ends); free(partner); free(allk); free(seen);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Partial unitigs of every unit + their routing, for the parity tests of the device builder
 * (ggcat_b200/csrc/unitigs.cuh).  compute_unitigs is ut_compute_unit above; the routing restates
 *   assembler_kmers_merge/src/final_executor.rs:103-245  output_sequence
 *   hashes/src/base/cn_seqhash_base.rs:140-150           get_bucket, constants hashes/src/cn_seqhash.rs:1-27 in the
 *                                                        integer width api/src/utils.rs:27-45 selects for k
 *   unitigs_extender/hashmap.rs:556-577                  is_circular (equal canonical (k-1)-mers at both closed ends)
 * Output per partial unitig (in emission order): unit, len, flags (1 open begin = bw_hash Some, 2 open end, 4 circular,
 * 8 should_rc, 16 HASH_ENDING, 32 OTHER_END), bucket (0xFFFF = lonely), last_align, byte offset of its bases (one 2-bit code
 * per byte) in `bases`.  Returns the number of partial unitigs (> cap_u / > cap_b total bases: nothing beyond the caps is
 * written, call again with larger buffers).
 * ---------------------------------------------------------------------------------------------- */
static uint32_t ut_get_bucket(u128 canon, unsigned bits, unsigned k) {
    if (k <= 8) {
        uint16_t x = (uint16_t)((uint16_t)canon * (uint16_t)0x0193 + (uint16_t)0x9dc5);
        x = (uint16_t)((x >> 3) | (x << 13));
        return (uint32_t)(x % (1u << bits));
    }
    if (k <= 16) {
        uint32_t x = (uint32_t)canon * 0x01000193u + 0x811c9dc5u;
        x = (x >> 3) | (x << 29);
        return (uint32_t)(x % (1u << bits));
    }
    if (k <= 32) {
        uint64_t x = (uint64_t)canon * 0x00000100000001b3ULL + 0xcbf29ce484222325ULL;
        x = (x >> 3) | (x << 61);
        return (uint32_t)(x % (1ULL << bits));
    }
    {
        const u128 mul = ((u128)0x0000000001000000ULL << 64) | 0x000000000000013bULL;
        const u128 bas = ((u128)0x6c62272e07bb0142ULL << 64) | 0x62b821756295c58dULL;
        u128 x = canon * mul + bas;
        x = (x >> 3) | (x << 125);
        return (uint32_t)(x % ((u128)1 << bits));
    }
}

size_t orc_partial_unitigs(const uint64_t *keys_lo, const uint64_t *keys_hi, const uint32_t *count_flags, const uint64_t *unit_off,
                           size_t n_units, unsigned first_unit, unsigned k, int forward_only, unsigned result_bits,
                           uint32_t *o_unit, uint32_t *o_len, uint8_t *o_flags, uint16_t *o_bucket, uint8_t *o_align,
                           uint64_t *o_off, size_t cap_u, uint8_t *bases, size_t cap_b, uint64_t *total_bases) {
    const u128 mask = (k >= 64) ? ~(u128)0 : ((((u128)1) << (2 * k)) - 1);
    size_t nu = 0, nb = 0;
    ut_table TC = {NULL, NULL, 0, k, forward_only, mask};
    for (size_t u = 0; u < n_units; u++) {
        const size_t a = unit_off[u], e = unit_off[u + 1];
        if (a == e) continue;
        ut_node *nodes = (ut_node *)malloc(sizeof(ut_node) * (e - a));
        for (size_t i = a; i < e; i++) {
            nodes[i - a].key = ((u128)(keys_hi ? keys_hi[i] : 0) << 64) | keys_lo[i];
            nodes[i - a].flags = (uint8_t)(count_flags[i] >> 30);
            nodes[i - a].used = 0;
        }
        ut_table T = {nodes, nodes, (long)(e - a), k, forward_only, mask};
        ut_plist L = {0};
        ut_compute_unit(&T, &L);
        for (size_t i = 0; i < L.n; i++) {
            const ut_partial *p = &L.v[i];
            uint32_t fl = (p->open_bw ? 1u : 0u) | (p->open_fw ? 2u : 0u);
            uint32_t bucket = 0xFFFFu, align = 0;
            const u128 first = ut_kmer_at(p->bases, 0, k), last = ut_kmer_at(p->bases, p->len - k, k);
            if (!p->open_bw && !p->open_fw) {
                const u128 m1 = (((u128)1) << (2 * (k - 1))) - 1;
                ut_table T1 = {NULL, NULL, 0, k - 1, 0, m1};
                if (ut_canon(&T1, first & m1) == ut_canon(&T1, last >> 2)) fl |= 4u;
            } else {
                const u128 fr = ut_rc(first, k), lr = ut_rc(last, k);
                const int first_fw = first < fr, last_fw = last < lr;   /* ExtendableHashTraitType::is_forward */
                const uint32_t lb = p->open_bw ? ut_get_bucket(first_fw ? first : fr, result_bits, k) : 0xFFFFu;
                const uint32_t rb = p->open_fw ? ut_get_bucket(last_fw ? last : lr, result_bits, k) : 0xFFFFu;
                const int lrc = p->open_bw ? !first_fw : 0, rrc = p->open_fw ? !last_fw : 1;
                const int hash_beginning = lb <= rb;
                const int should_rc = hash_beginning ? lrc : rrc;
                bucket = hash_beginning ? lb : rb;
                align = (hash_beginning ^ should_rc) ? 0u : (uint32_t)((p->len - k) % 4);
                if (should_rc) fl |= 8u;
                if ((!hash_beginning) ^ should_rc) fl |= 16u;
                if (p->open_bw && p->open_fw) fl |= 32u;
            }
            if (nu < cap_u) {
                o_unit[nu] = first_unit + (uint32_t)u; o_len[nu] = (uint32_t)p->len; o_flags[nu] = (uint8_t)fl;
                o_bucket[nu] = (uint16_t)bucket; o_align[nu] = (uint8_t)align; o_off[nu] = nb;
                if (nb + p->len <= cap_b) memcpy(bases + nb, p->bases, p->len);
            }
            nb += p->len;
            nu++;
            free(p->bases);
        }
        free(L.v);
        free(nodes);
    }
    (void)TC;
    *total_bases = nb;
    return nu;
}
