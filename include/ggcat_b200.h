/*
 * ggcat_b200.h -- C ABI of the B200-native k-mer counting front end for GGCAT.
 *
 * Drop-in boundary for the reference's phase 1 + phase 2:
 *   phase 1  minimizer bucketing   replaces assembler_minimizer_bucketing::minimizer_bucketing()
 *            (/root/reference/crates/assembler_minimizer_bucketing/src/lib.rs:279-340) and the
 *            GenericMinimizerBucketing::do_bucketing engine it drives
 *            (crates/minimizer_bucketing/src/lib.rs:543-685);
 *   phase 2  per-bucket k-mer merge replaces assembler_kmers_merge::kmers_merge()
 *            (crates/assembler_kmers_merge/src/lib.rs:159-284) up to the point where the
 *            k-mer table (FxHashMap<hash, MapEntry>, crates/structs/src/map_entry.rs:5-85) is
 *            complete, i.e. the input of HashMapUnitigsExtender::compute_unitigs
 *            (crates/assembler_kmers_merge/src/unitigs_extender/hashmap.rs:442-601).
 *
 * Plain pointers and sizes only; no C++ or torch types.  All functions return 0 on success and a
 * negative ggcat_b200_status on failure; ggcat_b200_last_error() gives the thread-local message.
 * Nothing here falls back to the CPU: without a CUDA device every compute call fails with
 * GGCAT_B200_ERR_CUDA.
 */
#ifndef GGCAT_B200_H
#define GGCAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGCAT_B200_ABI_VERSION 8

typedef enum {
    GGCAT_B200_OK = 0,
    GGCAT_B200_ERR_INVALID = -1,   /* bad argument / unsupported parameter combination */
    GGCAT_B200_ERR_CUDA = -2,      /* CUDA runtime error (no device, OOM, launch failure) */
    GGCAT_B200_ERR_STATE = -3,     /* call out of order (e.g. merge before finish_bucketing) */
    GGCAT_B200_ERR_CAPACITY = -4   /* caller-provided buffer too small */
} ggcat_b200_status;

/* Same numeric values as the reference's HashType (crates/api/src/utils.rs:4-8). */
typedef enum { GGCAT_B200_HASH_AUTO = 0, GGCAT_B200_HASH_SEQ = 1, GGCAT_B200_HASH_RK128 = 4 } ggcat_b200_hash_type;

/* The knobs of `ggcat build` that reach the hot path (crates/cmdline/src/main.rs:148-242,
 * crates/api/src/lib.rs:75-99): -k, --minimizer-length, -s, -b, -f, -w, -c. */
typedef struct {
    uint32_t k;                        /* k-mer length, 4 <= k <= 128 (odd k <= 31: 64-bit keys; else 128-bit keys; k > 64: rabin-karp128 only) */
    uint32_t m;                        /* minimizer length, 0 = compute_best_m(k) (crates/utils/src/lib.rs:29-40) */
    uint32_t min_multiplicity;         /* -s */
    uint32_t buckets_count_log;        /* first-level buckets = 1 << this (+1 duplicates bucket) */
    uint32_t second_buckets_count_log; /* sub-buckets per bucket = 1 << this (<= 8) */
    uint32_t forward_only;             /* -f */
    uint32_t hash_type;                /* ggcat_b200_hash_type; AUTO = seq-hash for k <= 64, rabin-karp128 above (crates/api/src/utils.rs:17-26) */
    uint32_t colors;                   /* -c : carry a colour id per record (seq-hash, k <= 48) */
    int32_t device;                    /* CUDA device ordinal */
    uint32_t reserved[7];
} ggcat_b200_params;

typedef struct ggcat_b200_ctx ggcat_b200_ctx;

/* Counters of the same meaning as the reference's GroupProcessStats / phase-1 counters
 * (crates/kmers_transform/src/lib.rs:77-82; crates/minimizer_bucketing/src/lib.rs:402-434). */
typedef struct {
    uint64_t total_bases;      /* bytes pushed */
    uint64_t valid_bases;      /* bases inside N-free segments of length >= k */
    uint64_t n_superkmers;
    uint64_t n_kmers;          /* k-mer occurrences stored in buckets (boundary copies counted twice) */
    uint64_t payload_words;    /* 32-bit words of packed super-k-mer payload */
    uint32_t n_buckets;        /* (1 << buckets_count_log) + 1 */
    uint32_t n_units;          /* n_buckets << second_buckets_count_log */
} ggcat_b200_bucket_stats;

/* One stored super-k-mer = the reference's PushSequenceInfo / CompressedReadsBucketData
 * (crates/minimizer_bucketing/src/lib.rs:105-114; crates/io/src/concurrent/temp_reads/creads_utils.rs:80-86). */
typedef struct {
    uint64_t payload_offset;   /* byte offset of the packed bases in the payload buffer (4-byte aligned) */
    uint32_t len;              /* bases; packed 4 per byte, stored orientation (rc already applied) */
    uint32_t color;
    uint16_t bucket;
    uint16_t minimizer_pos;
    uint8_t second_bucket;
    uint8_t flags;             /* READ_FLAG_INCL_BEGIN=1 | READ_FLAG_INCL_END=2 (crates/config/src/lib.rs:93-94) */
    uint8_t rc;
    uint8_t pad;
} ggcat_b200_superkmer;

/* k-mer table of a range of buckets: what HashMapUnitigsExtender holds after add_sequence().
 * Entries are grouped by merge unit (bucket, second_bucket) and sorted ascending by key inside a
 * unit.  Only entries with multiplicity >= min_multiplicity are present.
 *   count_flags = multiplicity (low 30 bits, saturating at 2^30 - 1: the reference's counter is 61-bit,
 *   crates/structs/src/map_entry.rs:5-10, but -s only needs "at least") | MapEntry flags << 30.
 * All pointers are library-owned pinned host memory, valid until ggcat_b200_release_table(). */
typedef struct {
    uint64_t n_entries;
    const uint64_t *keys_lo;
    const uint64_t *keys_hi;        /* high 64 bits of the key; NULL on the 64-bit key path (seq-hash, k <= 31, no colours) */
    const uint32_t *count_flags;
    uint32_t first_unit;            /* first_bucket << second_buckets_count_log */
    uint32_t n_units;
    const uint64_t *unit_offsets;   /* n_units + 1 entry offsets */
    const uint64_t *color_offsets;  /* n_entries + 1 offsets into colors, NULL when uncoloured */
    const uint32_t *colors;         /* sorted-unique colour ids per entry */
    uint64_t total_kmers;           /* k-mer occurrences processed */
    uint64_t unique_kmers;          /* distinct keys before the multiplicity filter (0 = not tracked, coloured builds) */
    /* Non-invertible keys (rabin-karp128): the bases of one occurrence of every entry, so that the consumer can rebuild
     * sequence from a hash -- the role of the reference's `saved_reads` + `encoded_saved_reads_indexes`
     * (crates/assembler_kmers_merge/src/unitigs_extender/hashmap.rs:32-33,96-149 get_kmers, :413-431 add_sequence).
     * Entry e owns words [e * src_kmer_words, (e + 1) * src_kmer_words), src_kmer_words = max(2, ceil(k / 32)); base j of
     * the k-mer sits at bits 2(j % 32) of word j / 32 (A0 C1 T2 G3), in the orientation whose FORWARD hash equals the key.
     * NULL / 0 for invertible keys (seq-hash). */
    const uint64_t *src_kmers;
    uint32_t src_kmer_words;
    uint32_t reserved0;
    void *opaque;
} ggcat_b200_table;

const char *ggcat_b200_last_error(void);
uint32_t ggcat_b200_abi_version(void);
/* crates/utils/src/lib.rs:29-40 */
uint32_t ggcat_b200_compute_best_m(uint32_t k);
/* crates/io/src/lib.rs:67-140: bucket-count heuristics from the estimated input size in bytes. */
void ggcat_b200_bucket_counts(uint64_t estimated_bases, uint32_t *buckets_count_log, uint32_t *second_buckets_count_log);

int32_t ggcat_b200_create(const ggcat_b200_params *params, ggcat_b200_ctx **out);
void ggcat_b200_destroy(ggcat_b200_ctx *ctx);

/* Pinned host staging memory for callers that want asynchronous copies. */
void *ggcat_b200_host_alloc(uint64_t bytes);
void ggcat_b200_host_free(void *p);

/* Phase 1.  `data` holds n_reads raw ASCII records back to back (record r = data[offsets[r] ..
 * offsets[r+1])), as produced by the reference's reader before normalisation
 * (crates/io/src/sequences_reader.rs:106-179).  colors: one id per record or NULL.
 * The batch is normalised, N-split, hashed, split into super-k-mers and scattered into
 * device-resident buckets.  May be called repeatedly and from several host threads (calls are serialised on the
 * context); each call appends ONE bucket chunk: the host input travels through a ring of three H2D staging batches of <= 24 MB
 * (GGCAT_B200_HOST_BATCH) whose kernels overlap the next copy, and the batches of a call are scattered together. */
int32_t ggcat_b200_push_reads(ggcat_b200_ctx *ctx, const uint8_t *data, const uint64_t *offsets, uint64_t n_reads,
                              const uint32_t *colors);
/* Same, with data/offsets/colors already resident in device memory of ctx's device. */
int32_t ggcat_b200_push_reads_device(ggcat_b200_ctx *ctx, const uint8_t *d_data, const uint64_t *d_offsets,
                                     uint64_t n_reads, uint64_t n_bytes, const uint32_t *d_colors);
/* 2-bit packed input (the `is_packed` form of SURVEY 8(b)): `packed` is ONE contiguous stream in the reference's
 * CompressedRead layout (base i at bits 2(i % 4) of byte i / 4, A0 C1 T2 G3; crates/io/src/compressed_read.rs:610-618) and
 * record r holds the bases [offsets[r], offsets[r+1]) of it -- offsets count bases, records need not start on a byte.
 * The layout has no code for N: a host that packs reads splits them at N first, exactly as the reference does before it
 * compresses a read (crates/minimizer_bucketing/src/sequences_splitter.rs:15-40).  A quarter of the H2D bytes of the ASCII
 * form and no k_pack pass; everything downstream is identical.  The device variant wants a 4-byte aligned stream whose
 * first base is offsets[0] = 0. */
int32_t ggcat_b200_push_reads_packed(ggcat_b200_ctx *ctx, const uint8_t *packed, const uint64_t *offsets, uint64_t n_reads,
                                     const uint32_t *colors);
int32_t ggcat_b200_push_reads_packed_device(ggcat_b200_ctx *ctx, const uint8_t *d_packed, const uint64_t *d_offsets,
                                            uint64_t n_reads, uint64_t n_bases, const uint32_t *d_colors);
/* Input side on the device: raw FASTA / FASTQ TEXT instead of tokenised records.  Replaces the
 * reference's line reader and record state machines (crates/io/src/lines_reader.rs:140-175,
 * crates/io/src/sequences_reader.rs:106-179 process_fasta, :181-241 process_fastq): '>' lines start a record, ';' lines
 * are comments, every other line is appended to the record's bases ('\n' and a '\r' right before it dropped); FASTQ is
 * strict 4-line records; records without bases are not emitted.  The block must hold whole records (a record may not
 * continue in the next call) and at most 2^30 bytes; decompression (.gz) stays on the host.  Every record gets `color`.
 * Records longer than the reference's 4 MiB split (sequences_reader.rs:119-122) are processed in one piece here -- same
 * k-mers, no k-1 overlap copies needed. */
typedef enum { GGCAT_B200_TEXT_FASTA = 0, GGCAT_B200_TEXT_FASTQ = 1 } ggcat_b200_text_format;
int32_t ggcat_b200_push_text(ggcat_b200_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int32_t format, uint32_t color,
                             uint64_t *n_records);
int32_t ggcat_b200_push_text_device(ggcat_b200_ctx *ctx, const uint8_t *d_text, uint64_t n_bytes, int32_t format,
                                    uint32_t color, uint64_t *n_records);
/* The tokenizer alone: d_seq / d_offsets (n_records + 1 entries) are device buffers owned by the context, valid until
 * the next tokenize / push_text call. */
int32_t ggcat_b200_tokenize_device(ggcat_b200_ctx *ctx, const uint8_t *d_text, uint64_t n_bytes, int32_t format,
                                   const uint8_t **d_seq, const uint64_t **d_offsets, uint64_t *n_records,
                                   uint64_t *n_seq_bytes);

/* Returns as soon as the per-unit counts of every chunk are on the host (they are final before the last scatter
 * kernel ends); later calls are ordered on the context stream.  Callers that read chunk buffers from another stream
 * go through export_chunk_slice (which synchronises) or call ggcat_b200_synchronize(). */
int32_t ggcat_b200_finish_bucketing(ggcat_b200_ctx *ctx, ggcat_b200_bucket_stats *stats);

/* The reference's bucket FILE format, so that a GPU phase 1 can feed an unmodified CPU phase 2 and a
 * CPU phase 1 can feed this phase 2: one PLAIN_INTR_BKT_M bucket file per first-level bucket, chunks grouped by
 * sub-bucket with ReadsCheckpointData checkpoints, MinimizerBucketMode::SingleGrouped records -- what the reference's
 * compactor leaves behind (crates/minimizer_bucketing/src/compactor.rs:370-420) and SplittedBucket::generate /
 * decode_sequences read (split_buckets.rs:46-131, decode_helper.rs:14-63); container:
 * libs-crates/parallel-processor-rs/src/buckets/writers/{mod.rs:15-70,lock_free_binary_writer.rs},
 * readers/binary_reader.rs:120-190; record: crates/io/src/concurrent/temp_reads/creads_utils.rs:374-434.
 * write: after finish_bucketing.  import: before finish_bucketing, instead of (or beside) push_reads; lz4
 * (CPLZ4_INTR_BKT_M) files and Compacted / coloured records are rejected.  Uncoloured builds only. */
int32_t ggcat_b200_write_bucket_file(ggcat_b200_ctx *ctx, uint32_t bucket, const char *path, uint64_t *n_records);
int32_t ggcat_b200_import_bucket_file(ggcat_b200_ctx *ctx, uint32_t bucket, const char *path, uint64_t *n_records);

/* Per-unit sizes after finish_bucketing: arrays of stats.n_units entries (either may be NULL). */
int32_t ggcat_b200_unit_sizes(ggcat_b200_ctx *ctx, uint64_t *n_superkmers, uint64_t *n_kmers);

/* Test hook (phase-1 parity): copies the super-k-mers of one first-level bucket to the host.
 * Two-call protocol: with out == NULL only the counts are returned. */
int32_t ggcat_b200_dump_superkmers(ggcat_b200_ctx *ctx, uint32_t bucket, ggcat_b200_superkmer *out, uint64_t cap,
                                   uint8_t *payload, uint64_t payload_cap, uint64_t *n_out, uint64_t *payload_bytes);

/* Phase 2 for buckets [first_bucket, first_bucket + n_buckets) (the duplicates bucket is index
 * 1 << buckets_count_log).  Fills *out with pinned host copies of the filtered table. */
int32_t ggcat_b200_merge_bucket_range(ggcat_b200_ctx *ctx, uint32_t first_bucket, uint32_t n_buckets,
                                      ggcat_b200_table *out);
int32_t ggcat_b200_release_table(ggcat_b200_ctx *ctx, ggcat_b200_table *table);

/* Phase 2 without the device->host copy: the table stays in HBM (bench `value`, multi-GPU owners).
 * Returns entry count and distinct/total k-mer counters. */
int32_t ggcat_b200_merge_bucket_range_device(ggcat_b200_ctx *ctx, uint32_t first_bucket, uint32_t n_buckets,
                                             uint64_t *n_entries, uint64_t *unique_kmers, uint64_t *total_kmers);

/* The table left in HBM by the last merge_bucket_range_device, in the layout of ggcat_b200_table: every pointer is
 * DEVICE memory owned by the context, valid until the next merge / reset (opaque is NULL; nothing to release).  This is
 * what a device-side consumer (partial-unitig construction) or a verifier reads without a host copy. */
int32_t ggcat_b200_device_table(ggcat_b200_ctx *ctx, ggcat_b200_table *out);

/* Partial unitigs of the table left by the last merge_bucket_range_device, built on the device (SURVEY 8(f)-1 + rows
 * a12 / a13): what HashMapUnitigsExtender::compute_unitigs + try_extend_function
 * (crates/assembler_kmers_merge/src/unitigs_extender/hashmap.rs:162-297,442-601) produce per merge unit, with the routing
 * ParallelKmersMergeFinalExecutor::output_sequence (final_executor.rs:103-245) would apply for
 * 1 << result_buckets_log "result" buckets (crates/assembler_kmers_merge/src/lib.rs:206-212).
 * flags: 1 open at the beginning (backward hash is Some: the unitig continues in another unit), 2 open at the end,
 *        4 circular (hashmap.rs:556-577), 8 should_rc, 16 HASH_ENDING_FLAG, 32 OTHER_END_FLAG
 *        (crates/io/src/partial_unitigs_extra_data.rs:16-18) -- the last three as output_sequence stores them.
 * bucket: result bucket of an open unitig; 0xFFFF for unitigs closed at both ends (written straight to the final file).
 * The bases are stored in walk orientation (base i at bits 2(i%16) of word word_offset + i/16); a consumer that writes the
 * reference's "result" buckets applies should_rc itself.  Host pointers are pinned memory owned by the context, valid until
 * the next call; d_* are the same arrays in HBM for a device-side consumer (extend_unitigs).
 * 64-bit key path only (seq-hash, odd k <= 31, uncoloured). */
typedef struct {
    uint64_t word_offset;
    uint32_t len;            /* bases */
    uint32_t unit;           /* merge unit (bucket << second_buckets_count_log | second_bucket) it was built in */
    uint16_t bucket;
    uint8_t flags;
    uint8_t last_align;      /* output_sequence's minimizer_pos field: (len - k) % 4 or 0 */
    uint32_t n_kmers;        /* len - k + 1 */
} ggcat_b200_unitig;
typedef struct {
    uint64_t n_unitigs, n_words, n_kmers;
    const ggcat_b200_unitig *unitigs;
    const uint32_t *bases;
    const void *d_unitigs;
    const uint32_t *d_bases;
} ggcat_b200_unitigs;
int32_t ggcat_b200_partial_unitigs(ggcat_b200_ctx *ctx, uint32_t result_buckets_log, ggcat_b200_unitigs *out);
/* Joins the partial unitigs of the last ggcat_b200_partial_unitigs call into MAXIMAL unitigs, on the device
 * (SURVEY 8(f)-2): the fixed point of the reference's "phase: unitigs joining"
 * (crates/assembler_pipeline/src/extend_unitigs.rs:348-): two partial unitigs that end in the same canonical k-mer are
 * glued with that k-mer shared, chains and cycles of partial unitigs become one unitig each.  The partial unitigs must
 * cover the whole build (merge_bucket_range_device over all buckets of one GPU); across GPUs the partial unitigs are
 * first brought together by their result bucket (flags / bucket above).  Records: unit = number of partial unitigs glued,
 * bucket = 0xFFFF, flags = 4 for circular unitigs (k + L - 1 bases for L k-mers).  Orientation is the walk's (the
 * reference's depends on thread timing); compare canonical k-mer sets.  Invalidates the pointers of partial_unitigs. */
int32_t ggcat_b200_maximal_unitigs(ggcat_b200_ctx *ctx, ggcat_b200_unitigs *out);

/* Drops all bucket chunks so the context can be reused for another build. */
int32_t ggcat_b200_reset(ggcat_b200_ctx *ctx);

/* ---- multi-GPU plumbing: bucket chunks as plain device buffers ---------------------------------
 * After finish_bucketing a context holds one chunk per push.  The host (torch.distributed / NCCL in
 * this repo, any transport in a Rust host) routes the slice of every chunk that belongs to units
 * [first_unit, first_unit + n_units) to the owner rank and registers it there with import_chunk. */
typedef struct {
    uint64_t n_superkmers;        /* descriptors in the slice */
    uint64_t n_words;             /* payload words in the slice */
    uint64_t word_bias;           /* payload word offset of the slice inside its source chunk: descriptors
                                     hold offsets relative to the source chunk, the receiver subtracts this */
    const void *d_descriptors;    /* 16 bytes each */
    const uint32_t *d_payload;
    const uint32_t *d_unit_counts;   /* n_units super-k-mer counts   (device) */
    const uint32_t *d_unit_words;    /* n_units payload word counts  (device) */
    const uint32_t *d_unit_kmers;    /* n_units k-mer counts         (device) */
    /* Host copies of the three per-unit arrays.  export fills them (library-owned, valid until reset/drop);
     * import uses them instead of reading the device arrays back when all three are non-NULL. */
    const uint32_t *h_unit_counts;
    const uint32_t *h_unit_words;
    const uint32_t *h_unit_kmers;
} ggcat_b200_chunk_slice;

uint32_t ggcat_b200_n_chunks(ggcat_b200_ctx *ctx);
int32_t ggcat_b200_export_chunk_slice(ggcat_b200_ctx *ctx, uint32_t chunk, uint32_t first_unit, uint32_t n_units,
                                      ggcat_b200_chunk_slice *out);
/* Registers a received slice (device pointers stay owned by the caller and must outlive the merge). */
int32_t ggcat_b200_import_chunk_slice(ggcat_b200_ctx *ctx, uint32_t first_unit, uint32_t n_units,
                                      const ggcat_b200_chunk_slice *slice);
/* Forget the chunks produced locally by push_reads (after they were exported) but keep imports. */
int32_t ggcat_b200_drop_local_chunks(ggcat_b200_ctx *ctx);

/* ---- multi-GPU exchange over NVLink peer memory (one process per GPU, one NVSwitch box) --------------
 * The sharded build's only exchange: the reference writes per-bucket temp files in phase 1
 * (crates/minimizer_bucketing/src/lib.rs:340-351) and reads them back in phase 2
 * (crates/kmers_transform/src/lib.rs:294-371); here rank r owns the contiguous bucket range
 * ggcat_b200_owner_range(r) (the duplicates bucket goes to the last rank) and every rank pushes the slices of its
 * bucket chunks straight into the owners' receive arenas with a copy kernel (CUDA IPC mapping, 16-byte stores over
 * NVLink, device-side ready/released flags; ggcat_b200/csrc/peer.cuh).  Protocol, collective over all ranks:
 *   peer_init (allocates the arena, returns its IPC handle)  ->  the host all-gathers the 64-byte handles with
 *   any transport  ->  peer_connect(handles in rank order)  ->  per build: push_reads*, finish_bucketing,
 *   peer_exchange, merge_bucket_range[_device](owner range).
 * arena_bytes is split evenly into one region per source rank; a region must hold that rank's descriptors
 * (16 B per super-k-mer) + payload + per-unit counts for this owner (GGCAT_B200_ERR_CAPACITY otherwise). */
typedef struct { uint8_t bytes[64]; } ggcat_b200_peer_handle;
int32_t ggcat_b200_owner_range(uint32_t buckets_count_log, uint32_t rank, uint32_t world, uint32_t *first_bucket,
                               uint32_t *n_buckets);
int32_t ggcat_b200_peer_init(ggcat_b200_ctx *ctx, uint32_t rank, uint32_t world, uint64_t arena_bytes,
                             ggcat_b200_peer_handle *out_handle);
int32_t ggcat_b200_peer_connect(ggcat_b200_ctx *ctx, const ggcat_b200_peer_handle *handles);
/* After finish_bucketing on every rank.  Local chunks stay registered (the owner's own units are merged in
 * place); the slices received from the other ranks are registered as imported chunks living in the arena. */
int32_t ggcat_b200_peer_exchange(ggcat_b200_ctx *ctx);
/* Bytes this rank pushed over NVLink / received into its arena in the last peer_exchange (descriptors, payload,
 * per-unit counts): the numerator of the "all-to-all bytes against NVLink bandwidth" figure of the bench. */
int32_t ggcat_b200_peer_stats(ggcat_b200_ctx *ctx, uint64_t *bytes_sent, uint64_t *bytes_received);

/* ---- measurement hooks ------------------------------------------------------------------------- */
/* Per-kernel-family CUDA-event timing on/off (off by default: two event records per launch). */
int32_t ggcat_b200_set_timing(ggcat_b200_ctx *ctx, int32_t enabled);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on. */
void *ggcat_b200_stream(ggcat_b200_ctx *ctx);
int32_t ggcat_b200_synchronize(ggcat_b200_ctx *ctx);
/* Device time (ms, CUDA events on the context stream) and launch count of each kernel family since
 * the last call with reset != 0.  names/ms/launches receive up to cap entries; returns the number
 * of families. */
int32_t ggcat_b200_kernel_times(ggcat_b200_ctx *ctx, const char **names, float *ms, uint32_t *launches, uint32_t cap,
                                int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* GGCAT_B200_H */
