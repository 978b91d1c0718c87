// capi.cu -- C ABI (include/ggcat_b200.h) over the phase-1 / phase-2 kernels.
// Host orchestration only: buffers, launches, chunk bookkeeping.  No CPU compute path exists;
// every entry point fails with GGCAT_B200_ERR_CUDA when no device is usable.
#include "../../include/ggcat_b200.h"
#include "merge128.cuh"
#include "peer.cuh"
#include "tokenize.cuh"
#include "wire.hpp"
#include "unitigs.cuh"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace ggb;

static thread_local std::string g_err;
static int32_t set_err(int32_t code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
// GGCAT_B200_TRACE=host: wall-clock marks of the host side of every entry point (us since the first mark), on stderr.
static void trace_host(const char *label) {
    const char *e = getenv("GGCAT_B200_TRACE");
    if (!e || strcmp(e, "host") != 0) return;
    static const auto t0 = std::chrono::steady_clock::now();
    fprintf(stderr, "[ggcat_b200 host] %9.1f us  %s\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(), label);
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return set_err(GGCAT_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                                         \
    } while (0)
#define TRY(call)                   \
    do {                            \
        int32_t r__ = (call);       \
        if (r__ != 0) return r__;   \
    } while (0)

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Chunk {
    bool imported = false;
    DevBuf desc, payload, unit_cnt, unit_off, unit_words, unit_woff, unit_kmers;  // owned (local chunks)
    const uint4 *d_desc = nullptr;          // views used by kernels
    const uint32_t *d_payload = nullptr;
    const uint32_t *d_unit_cnt = nullptr, *d_unit_off = nullptr, *d_unit_words = nullptr, *d_unit_woff = nullptr,
                   *d_unit_kmers = nullptr;
    uint32_t first_unit = 0, n_units = 0, word_bias = 0;
    uint64_t n_sk = 0, n_words = 0, n_kmers = 0, n_bases = 0, n_segments = 0;
    std::vector<uint32_t> h_cnt, h_off, h_words, h_woff, h_kmers;  // host mirrors (after finish / import)
    uint32_t *h_pin = nullptr; size_t h_pin_cap = 0;   // pinned landing area of the per-unit counts (cnt | words | kmers),
    bool mirror_queued = false;                        // filled by copies queued right behind k_emit on the copy stream
    cudaEvent_t ev_emit = nullptr, ev_mirror = nullptr;   // histograms complete / mirror landed
    void release() {
        desc.release(); payload.release(); unit_cnt.release(); unit_off.release(); unit_words.release();
        unit_woff.release(); unit_kmers.release();
        if (h_pin) cudaFreeHost(h_pin);
        h_pin = nullptr; h_pin_cap = 0;
        if (ev_emit) cudaEventDestroy(ev_emit);
        if (ev_mirror) cudaEventDestroy(ev_mirror);
        ev_emit = ev_mirror = nullptr;
    }
};

enum Family { F_PACK = 0, F_WINDOWS, F_EMIT, F_SCATTER, F_SCAN, F_MERGE_HASH, F_MERGE_HASH_GLOBAL, F_MERGE_SMEM, F_MERGE_GLOBAL,
              F_GATHER, F_MERGE_HASH128, F_SORT128, F_COLOR_FOLD, F_PARTITION, F_MERGE_HASH_PART, F_PEER, F_PEER_PUSH, F_TOKENIZE, F_UNITIGS, F_COUNT };
static const char *kFamilyNames[F_COUNT] = {"k_pack+k_mark", "k_windows", "k_emit", "k_scatter", "k_exclusive_scan_u32",
                                            "k_merge_hash<smem>", "k_merge_hash<global>", "k_merge_units<smem>",
                                            "k_merge_units<global>", "k_gather_units", "k_merge_hash128", "k_sort_units128",
                                            "k_color_fold", "k_partition_units", "k_merge_hash<partitions>", "k_peer_sync<exposed>", "k_peer_push<side streams>", "k_tok_*", "k_unitig_*"};

struct TimedLaunch { int fam; cudaEvent_t a, b; };

struct HostTable {
    uint64_t *keys = nullptr, *keys_hi = nullptr; uint32_t *cf = nullptr; uint64_t *unit_offsets = nullptr;
    uint64_t *color_offsets = nullptr; uint32_t *colors = nullptr;
    uint64_t *src = nullptr;   // rk128: 2 words of source bases per entry
    size_t cap_entries = 0, cap_units = 0, cap_colors = 0, cap_coloff = 0;
    bool wide = false, colored = false, with_src = false;
    void release() {
        cudaFreeHost(keys); cudaFreeHost(keys_hi); cudaFreeHost(cf); cudaFreeHost(unit_offsets); cudaFreeHost(color_offsets);
        cudaFreeHost(colors); cudaFreeHost(src);
    }
};

// where the finished table of the last merge lives on the device
struct FinalTable {
    const uint64_t *keys_lo = nullptr, *keys_hi = nullptr; const uint32_t *cf = nullptr;
    const uint64_t *src = nullptr;           // rk128: source bases, 2 words per entry
    const uint64_t *unit_off = nullptr;      // n_units + 1
    const uint64_t *color_off = nullptr;     // n_entries (+1 implied = n_colors)
    const uint32_t *colors = nullptr;
    uint64_t n_entries = 0, n_colors = 0;
    uint32_t first_unit = 0, n_units = 0;    // unit range of the merge that produced it
    uint64_t total_kmers = 0, unique_kmers = 0;
};

// NVLink peer exchange (peer.cuh): this rank's receive arena, the peers' arenas mapped through CUDA IPC, staging.
struct PeerState {
    bool inited = false, connected = false;
    uint32_t rank = 0, world = 1, epoch = 0;
    uint8_t *arena = nullptr;
    uint64_t arena_bytes = 0, region_bytes = 0;
    uint8_t *peer_arena[PEER_MAX_WORLD] = {};
    uint32_t meta_slots = 64;                 // S: bucket chunks one build may route (GGCAT_B200_PEER_SLICES, <= PEER_MAX_SLICES)
    cudaStream_t meta_stream = nullptr, data_stream = nullptr;   // per-unit counts + slice entries / bulk slices
    cudaEvent_t ev_scatter = nullptr, ev_data = nullptr, ev_meta = nullptr;
    bool build_started = false;
    uint32_t n_pushed = 0;                    // local chunks already pushed in this build
    uint64_t cursor[PEER_MAX_WORLD] = {};     // next free byte of this rank's region on destination d
    bool overflow[PEER_MAX_WORLD] = {};
    DevBuf d_stage, d_err;                    // d_stage: ring of per-chunk blocks (copy jobs + the slice entries / headers being sent)
    uint8_t *h_stage = nullptr; size_t h_stage_cap = 0;   // pinned mirror of d_stage
    uint8_t *h_recv = nullptr; size_t h_recv_cap = 0;     // pinned: received headers, tables and per-unit counts
    uint64_t last_sent = 0, last_received = 0;            // payload + descriptor + metadata bytes of the last exchange
};

}  // namespace

struct ggcat_b200_ctx {
    std::mutex mu;   // serialises the entry points that change the context (push_reads may be called from many host threads)
    ggcat_b200_params params;
    DevParams P;
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    uint64_t max_batch = 1ull << 30;
    uint64_t host_batch = 24ull << 20;   // push_reads(host): the H2D of batches i+1, i+2 overlaps the kernels of batch i (24 MB measured best on C2: 3.41 vs 3.54 ms at 48 MB)
    uint64_t part_kmers = 36ull << 20;   // merge_bucket_range(host): D2H of part j overlaps the merge of part j+1
    uint64_t part_kmers_dev = 192ull << 20;  // merge_bucket_range_device: bounds the per-part scratch (12 B / record + key partitions)
    double distinct_ratio = 1.0;         // distinct keys / records of the parts merged so far (sizes the key partitions)
    bool part_fixed = false;             // GGCAT_B200_PART_TARGET=fixed: 4096-record partitions whatever the ratio (tests)
    bool part_dev_fixed = false;         // GGCAT_B200_PART_KMERS_DEV given: no grid-filling enlargement (tests)
    uint64_t fin_cap = 0;                // entries the final table (out_keys2 / out_cf2 / out_hi2) can hold
    uint64_t final_hint = 0;             // survivors of the previous build of this context (sizes the next final table)
    cudaStream_t copy_stream = nullptr;
    static constexpr int N_STAGE = 3;    // H2D staging slots: the copy of batch i+2 never waits for the kernels of batch i
    cudaEvent_t ev_h2d[N_STAGE] = {}, ev_free[N_STAGE] = {}, ev_part = nullptr;
    DevBuf st_ascii[N_STAGE], st_off[N_STAGE], st_col[N_STAGE];
    int part_scheme = 1;                 // GGCAT_B200_PART_SCHEME: 0 = shrinking parts (first = half), 1 = a small first part, then equal parts
    bool finished = false;
    bool timing = false;
    int merge_mode = 1;  // 1 = shared-memory hash table (default), 0 = LSD radix sort (GGCAT_B200_MERGE=sort)
    bool no_partition = false;  // GGCAT_B200_NO_PARTITION=1: big units use the global-scratch table (A/B switch)
    bool no_tiers = false;      // GGCAT_B200_NO_TIERS=1: every unit takes the key-partition / global-table path (tests)
    int tier_a_variant = 0;     // GGCAT_B200_TIER_A=1|2: A/B switch of the small-unit tier (see TIER_A1 / TIER_A256)
    int wide_mode = -1;  // -1: 64-bit key path (merge.cuh); else MODE_SEQ128 / MODE_RK128 / MODE_COLOR (merge128.cuh)
    RkTables rk;
    FinalTable fin;
    ggcat_b200_bucket_stats stats;
    // phase-1 workspace
    DevBuf d_ascii, d_offsets, d_colors, pk, bad, brk, ent, tile_cnt, tile_sbase, tile_group, tmp, tmp_color, cur_cnt, totals;
    std::vector<Chunk *> chunks;
    std::vector<Chunk *> chunk_pool;  // recycled local chunks (device buffers kept)
    Chunk *open_chunk = nullptr;      // chunk the batches of the current push accumulate into (closed by flush_open_chunk)
    uint64_t open_bases = 0, open_sk = 0;   // positions used in pk / descriptors held in tmp
    struct UnitTot { uint64_t n; uint32_t sk, w, sl, pad; };
    std::vector<UnitTot> unit_tot;    // merge: per-unit totals over all chunks (kept allocated between merges)
    // phase-2 workspace
    DevBuf ut_links, ut_visited, ut_recs, ut_bases, ut_counters;   // partial unitigs (unitigs.cuh)
    DevBuf mu_recs, mu_bases, mu_ht, mu_first, mu_partner, mu_visited;   // maximal unitigs (join of the partial ones)
    uint64_t ut_n = 0, ut_words = 0;                                 // partial unitigs / words of the last partial_unitigs call
    uint8_t *h_unitigs = nullptr; size_t h_unitigs_cap = 0;        // pinned: records, then packed bases
    DevBuf tok_agg, tok_keep, tok_base, tok_marks, tok_tmarks, tok_seq, tok_offsets, tok_text, tok_colors;   // FASTA / FASTQ tokenizer (tokenize.cuh)
    DevBuf d_static_off;                   // wide path: static output regions of partitioned units
    DevBuf d_rkpos;                        // rabin-karp per-position tables
    DevBuf d_recfl;                        // flag bits of the wide path's partition records
    uint32_t src_words = 2;                // rk128: 64-bit words of source bases per table entry = max(2, ceil(k / 32))
    DevBuf d_recsrc, out_src, out_src2, sort_idx;   // rk128: source locators of partition records, source bases (unsorted / final), sort index ping-pong
    DevBuf d_mstage;                       // merge uploads (views, work lists, unit_n, static_off) in one copy
    uint8_t *h_mstage = nullptr; size_t h_mstage_cap = 0;
    DevBuf d_views, d_work[3], d_scratch, out_keys, out_cf, out_keys2, out_cf2, cursor, unit_out_off,
        unit_out_cnt, unit_final_off, overflow, d_retry, d_partmeta, d_recs, fin_tmp_keys, fin_tmp_cf, out_hi, out_hi2, unit_keys, unit_cols, col_off, out_coloff, out_colors;
    unsigned long long *h_pinned = nullptr;  // small pinned staging (16 u64)
    std::vector<TimedLaunch> launches;
    std::vector<cudaEvent_t> event_pool;
    float fam_ms[F_COUNT];
    uint32_t fam_launches[F_COUNT];
    std::vector<HostTable *> free_tables;
    uint64_t last_entries = 0;
    PeerState peer;
};

namespace {

struct LaunchTimer {
    ggcat_b200_ctx *c; int fam; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    LaunchTimer(ggcat_b200_ctx *ctx, int f, uint32_t n_kernels = 1, cudaStream_t stream = nullptr) : c(ctx), fam(f), st(stream ? stream : ctx->stream) {
        c->fam_launches[fam] += n_kernels;
        if (!c->timing) return;
        auto get = [&]() { cudaEvent_t e; if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); } else cudaEventCreate(&e); return e; };
        a = get(); b = get();
        cudaEventRecord(a, st);
    }
    ~LaunchTimer() {
        if (!c->timing) return;
        cudaEventRecord(b, st);
        c->launches.push_back({fam, a, b});
    }
};

void collect_timings(ggcat_b200_ctx *c) {
    if (c->launches.empty()) return;
    cudaStreamSynchronize(c->stream);
    if (c->peer.data_stream) { cudaStreamSynchronize(c->peer.data_stream); cudaStreamSynchronize(c->peer.meta_stream); }
    const bool trace = getenv("GGCAT_B200_TRACE") != nullptr;   // timeline of every timed launch, ms since the first one
    for (auto &l : c->launches) {
        float ms = 0;
        cudaEventElapsedTime(&ms, l.a, l.b);
        c->fam_ms[l.fam] += ms;
        if (trace) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, c->launches.front().a, l.a);
            cudaEventElapsedTime(&t1, c->launches.front().a, l.b);
            fprintf(stderr, "[ggcat_b200 trace] %-28s %8.3f -> %8.3f ms\n", kFamilyNames[l.fam], t0, t1);
        }
        c->event_pool.push_back(l.a);
        c->event_pool.push_back(l.b);
    }
    c->launches.clear();
}

uint32_t best_m(uint32_t k) {  // crates/utils/src/lib.rs:29-40 compute_best_m
    if (k <= 13) return std::max(k / 2, k >= 4 ? k - 4 : 0u);
    if (k <= 15) return 9;
    if (k <= 21) return 10;
    if (k <= 30) return 11;
    if (k <= 37) return 12;
    if (k <= 42) return 13;
    if (k <= 64) return 14;
    return (uint32_t)((double)k / 4.0 + 0.5);
}

int32_t check_ctx(ggcat_b200_ctx *c) {
    if (!c) return set_err(GGCAT_B200_ERR_INVALID, "null context");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return set_err(GGCAT_B200_ERR_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
    return 0;
}

int32_t mirror_chunk(ggcat_b200_ctx *c, Chunk *ch);
int32_t peer_begin_build(ggcat_b200_ctx *c);
int32_t peer_push_chunk(ggcat_b200_ctx *c, Chunk *ch);
int32_t pinned_reserve(uint8_t **p, size_t *cap, size_t need);

// Grow a device buffer keeping its first `keep` bytes (the chunk being accumulated lives in pk / tmp / tmp_color).
int32_t reserve_keep(ggcat_b200_ctx *c, DevBuf &b, size_t bytes, size_t keep) {
    if (bytes <= b.cap) return 0;
    DevBuf nb;
    CU(nb.reserve(bytes + bytes / 4));
    if (keep && b.p) CU(cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, c->stream));
    if (b.p) { CU(cudaStreamSynchronize(c->stream)); b.release(); }
    b = nb;
    return 0;
}

int32_t flush_open_chunk(ggcat_b200_ctx *c);
void abort_open_chunk(ggcat_b200_ctx *c) {   // an error in the middle of a push: what was accumulated is dropped
    if (c->open_chunk) c->chunk_pool.push_back(c->open_chunk);
    c->open_chunk = nullptr; c->open_sk = 0; c->open_bases = 0;
}

// One batch of <= max_batch bases, inputs already on the device: pack, window minima, super-k-mer descriptors and the
// per-unit histograms.  Batches ACCUMULATE into the open chunk (packed bases, descriptors and unit counts of every batch of
// one push call); flush_open_chunk() scatters them into ONE unit-sorted bucket chunk, so a push of any size makes one
// chunk -- one slice per unit and source in the merge and in the exchange -- while the H2D copy of the next batch still
// overlaps everything but that final scatter.  `reserve_bases`: bases the whole push will bring (sizes the buffers once).
// packed_shift < 0: d_data holds n ASCII bases.  packed_shift = 0..15: d_data (4-byte aligned) is a 2-bit packed stream whose
// base number packed_shift is the batch's first base (ggcat_b200_push_reads_packed).
int32_t bucket_batch_device(ggcat_b200_ctx *c, const uint8_t *d_data, const uint64_t *d_offsets, uint64_t n_reads,
                            uint64_t off0, uint64_t n, const uint32_t *d_colors, uint64_t reserve_bases, int packed_shift = -1) {
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    if (n >= (1ull << 31)) return set_err(GGCAT_B200_ERR_INVALID, "batch of %llu bases exceeds 2^31", (unsigned long long)n);
    c->stats.total_bases += n;
    // sharded build: the "I am done with what I received in the last build" flag goes out BEFORE this build's kernels are
    // queued (behind the last merge in stream order) -- the peers' pushes into this rank's arena wait for it
    if (c->peer.connected && c->peer.world > 1 && !c->peer.build_started) TRY(peer_begin_build(c));
    if (n < P.k) return 0;
    const uint32_t n_tiles = (uint32_t)((n + WIN_T - 1) / WIN_T);
    const uint64_t padded = ((uint64_t)n_tiles * WIN_T + 4 * WIN_WMAX + 256 + 1023) & ~1023ull;   // positions this batch occupies in pk
    const uint64_t n_groups = padded / 32;  // 32-base groups incl. padding
    // positions inside a chunk are 32-bit: a chunk that would outgrow them is closed first
    if (c->open_chunk && c->open_bases + padded >= (1ull << 32) - (1ull << 20)) TRY(flush_open_chunk(c));
    if (!c->open_chunk) {
        Chunk *ch;
        if (!c->chunk_pool.empty()) { ch = c->chunk_pool.back(); c->chunk_pool.pop_back(); }
        else ch = new Chunk();
        ch->imported = false; ch->word_bias = 0; ch->mirror_queued = false;
        ch->h_cnt.clear(); ch->h_off.clear(); ch->h_words.clear(); ch->h_woff.clear(); ch->h_kmers.clear();
        ch->first_unit = 0; ch->n_units = P.n_units; ch->n_sk = 0; ch->n_bases = 0;
        c->open_chunk = ch; c->open_bases = 0; c->open_sk = 0;
        const size_t ub = ((size_t)P.n_units + 2) * 4;
        cudaError_t e = ch->unit_cnt.reserve(ub);
        if (e == cudaSuccess) e = ch->unit_off.reserve(ub);
        if (e == cudaSuccess) e = ch->unit_words.reserve(ub);
        if (e == cudaSuccess) e = ch->unit_woff.reserve(ub);
        if (e == cudaSuccess) e = ch->unit_kmers.reserve(ub);
        if (e == cudaSuccess) e = cudaMemsetAsync(ch->unit_cnt.p, 0, ub, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(ch->unit_words.p, 0, ub, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(ch->unit_kmers.p, 0, ub, st);
        if (e != cudaSuccess) {
            c->chunk_pool.push_back(ch); c->open_chunk = nullptr;
            return set_err(GGCAT_B200_ERR_CUDA, "bucket chunk allocation failed: %s", cudaGetErrorString(e));
        }
    }
    Chunk *ch = c->open_chunk;
    const uint64_t base = c->open_bases;                      // first position of this batch inside the chunk's packed bases
    const uint64_t want_bases = std::max<uint64_t>(base + padded, std::min<uint64_t>(reserve_bases + (reserve_bases >> 4) + 8 * padded, (1ull << 32)));
    TRY(reserve_keep(c, c->pk, (want_bases / 16 + 16) * 4, (size_t)(base / 16) * 4));
    CU(c->bad.reserve((n_groups + 8) * 4));
    CU(c->brk.reserve((n_groups + 8) * 4));
    uint32_t *pkb = c->pk.as<uint32_t>() + base / 16;         // this batch's packed bases (16-byte aligned: base % 1024 == 0)
    CU(cudaMemsetAsync(c->brk.p, 0, (n_groups + 8) * 4, st));
    CU(cudaMemsetAsync(pkb + 2 * n_groups, 0, 8 * 4, st));
    CU(cudaMemsetAsync(c->bad.as<uint32_t>() + n_groups, 0xFF, 8 * 4, st));
    {
        LaunchTimer t(c, F_PACK, 2);
        const int aligned = ((uintptr_t)d_data & 15) == 0;
        if (packed_shift >= 0)
            k_repack<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(d_data), (uint32_t)packed_shift, n, pkb,
                                                                          c->bad.as<uint32_t>(), n_groups, (2 * (n + packed_shift) + 31) / 32);
        else
            k_pack<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(d_data, n, pkb, c->bad.as<uint32_t>(), n_groups, aligned);
        k_mark<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(d_offsets, n_reads, off0, n, c->brk.as<uint32_t>());
    }
    CU(c->ent.reserve((uint64_t)n_tiles * WIN_T * 8));
    CU(c->tile_cnt.reserve(((uint64_t)n_tiles + 2) * 4));
    CU(c->tile_sbase.reserve(((uint64_t)n_tiles + 2) * 4));
    const uint32_t n_tgroups = (n_tiles + WIN_GROUP - 1) / WIN_GROUP;
    CU(c->tile_group.reserve(((uint64_t)n_tgroups + 2) * 4));
    CU(cudaMemsetAsync(c->tile_group.p, 0, ((uint64_t)n_tgroups + 2) * 4, st));
    {
        LaunchTimer t(c, F_WINDOWS);
        auto kwin = (P.w <= 64 && P.k <= 66) ? k_windows<64> : k_windows<WIN_WMAX>;
        kwin<<<(n_tiles + WIN_TPC - 1) / WIN_TPC, WIN_THREADS, 0, st>>>(pkb, c->bad.as<uint32_t>(), c->brk.as<uint32_t>(),
                                                   (uint32_t)n, P, c->ent.as<uint64_t>(), c->tile_cnt.as<uint32_t>(),
                                                   c->tile_sbase.as<uint32_t>(), ch->unit_cnt.as<uint32_t>() + P.n_units /* spare slot: segments */,
                                                   n_tiles, c->tile_group.as<uint32_t>());
    }
    CU(c->totals.reserve(8 * 8));
    {
        LaunchTimer t(c, F_SCAN);
        // two-level prefix: only the per-group totals are scanned here (n_tiles / 64 values, one pass of one CTA)
        k_exclusive_scan_u32<<<1, 1024, 0, st>>>(c->tile_group.as<uint32_t>(), c->tile_group.as<uint32_t>(), n_tgroups,
                                                 c->totals.as<unsigned long long>());
    }
    CU(cudaMemcpyAsync(c->h_pinned, c->totals.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint64_t n_sk = c->h_pinned[0];
    trace_host("bucket_batch: super-k-mer count on the host");
    if (n_sk == 0) return 0;
    if (c->open_sk + n_sk >= (1ull << 32)) return set_err(GGCAT_B200_ERR_INVALID, "too many super-k-mers in one push");
    TRY(reserve_keep(c, c->tmp, (c->open_sk + n_sk) * 16, (size_t)c->open_sk * 16));
    if (P.colors) TRY(reserve_keep(c, c->tmp_color, (c->open_sk + n_sk) * 4, (size_t)c->open_sk * 4));
    {
        LaunchTimer t(c, F_EMIT);
        k_emit<<<n_tiles, 256, 0, st>>>(c->ent.as<uint64_t>(), c->tile_cnt.as<uint32_t>(), c->tile_sbase.as<uint32_t>(),
                                        c->tile_group.as<uint32_t>(), n_tiles, P, c->tmp.as<uint4>() + c->open_sk, c->tmp_color.as<uint32_t>() + c->open_sk, (uint32_t)base,
                                        d_offsets, n_reads, off0, d_colors, ch->unit_cnt.as<uint32_t>(), ch->unit_words.as<uint32_t>(),
                                        ch->unit_kmers.as<uint32_t>());
    }
    CU(cudaGetLastError());
    c->open_sk += n_sk; c->open_bases += padded; ch->n_bases += n;
    trace_host("bucket_batch: kernels queued");
    return 0;
}

// Close the open chunk: per-unit counts to the host (copy stream), offset scans, one scatter of every descriptor /
// payload of the push into the unit-sorted layout; the chunk is registered (and, in a sharded build, pushed to the
// owners) only after the last operation that can fail.
int32_t flush_open_chunk(ggcat_b200_ctx *c) {
    Chunk *ch = c->open_chunk;
    if (!ch) return 0;
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    struct ChunkGuard {
        ggcat_b200_ctx *c; Chunk *ch;
        ~ChunkGuard() { if (ch) c->chunk_pool.push_back(ch); c->open_chunk = nullptr; c->open_sk = 0; c->open_bases = 0; }
    } guard{c, ch};
    const uint64_t n_sk = c->open_sk;
    if (n_sk == 0) return 0;      // nothing but reads shorter than k: the chunk goes back to the pool
    ch->n_sk = n_sk;
    const size_t ub = ((size_t)P.n_units + 2) * 4;
    // host mirror of the per-unit counts: they are final after the last k_emit, so they travel on the copy stream while
    // k_scatter runs; finish_bucketing waits for this event only, and the host side of the merge (unit classification,
    // uploads) overlaps the tail of phase 1
    if (c->copy_stream) {
        const size_t nu = P.n_units;
        if (ch->h_pin_cap < 3 * nu + 1) {
            if (ch->h_pin) cudaFreeHost(ch->h_pin);
            ch->h_pin = nullptr; ch->h_pin_cap = 0;
            CU(cudaHostAlloc(reinterpret_cast<void **>(&ch->h_pin), (3 * nu + 1) * 4, cudaHostAllocDefault));
            ch->h_pin_cap = 3 * nu + 1;
        }
        if (!ch->ev_emit) { CU(cudaEventCreateWithFlags(&ch->ev_emit, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&ch->ev_mirror, cudaEventDisableTiming)); }
        CU(cudaEventRecord(ch->ev_emit, st));
        CU(cudaStreamWaitEvent(c->copy_stream, ch->ev_emit, 0));
        CU(cudaMemcpyAsync(ch->h_pin, ch->unit_cnt.p, nu * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaMemcpyAsync(ch->h_pin + nu, ch->unit_words.p, nu * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaMemcpyAsync(ch->h_pin + 2 * nu, ch->unit_kmers.p, nu * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaMemcpyAsync(ch->h_pin + 3 * nu, ch->unit_cnt.as<uint32_t>() + nu, 4, cudaMemcpyDeviceToHost, c->copy_stream));   // segments
        CU(cudaEventRecord(ch->ev_mirror, c->copy_stream));
        ch->mirror_queued = true;
    }
    {
        LaunchTimer t(c, F_SCAN, 2);
        k_exclusive_scan_u32<<<1, 1024, 0, st>>>(ch->unit_cnt.as<uint32_t>(), ch->unit_off.as<uint32_t>(), P.n_units,
                                                 c->totals.as<unsigned long long>() + 1);
        k_exclusive_scan_u32<<<1, 1024, 0, st>>>(ch->unit_words.as<uint32_t>(), ch->unit_woff.as<uint32_t>(), P.n_units,
                                                 c->totals.as<unsigned long long>() + 2);
    }
    // payload bound: sum ceil(len/16) <= (bases covered)/16 + n_sk, bases covered <= n + n_sk*(k-1)
    const uint64_t words_bound = (ch->n_bases + n_sk * (uint64_t)(P.k - 1)) / 16 + n_sk + 8;
    if (words_bound >= (1ull << 32)) return set_err(GGCAT_B200_ERR_INVALID, "payload of one push exceeds 2^32 words: push smaller batches");
    CU(ch->desc.reserve(n_sk * 16));
    CU(ch->payload.reserve(words_bound * 4));
    CU(c->cur_cnt.reserve(((size_t)P.n_units + 2) * 8));
    {
        LaunchTimer t(c, F_SCATTER, 2);
        k_init_cursors<<<(P.n_units + 255) / 256, 256, 0, st>>>(ch->unit_off.as<uint32_t>(), ch->unit_woff.as<uint32_t>(), P.n_units,
                                                              c->cur_cnt.as<unsigned long long>());
        k_scatter<<<(unsigned)((n_sk + 255) / 256), 256, 0, st>>>(c->tmp.as<uint4>(), c->tmp_color.as<uint32_t>(), (uint32_t)n_sk,
                                                                  c->pk.as<uint32_t>(), c->cur_cnt.as<unsigned long long>(),
                                                                  ch->desc.as<uint4>(), ch->payload.as<uint32_t>(), P.colors);
    }
    ch->d_desc = ch->desc.as<uint4>(); ch->d_payload = ch->payload.as<uint32_t>();
    ch->d_unit_cnt = ch->unit_cnt.as<uint32_t>(); ch->d_unit_off = ch->unit_off.as<uint32_t>();
    ch->d_unit_words = ch->unit_words.as<uint32_t>(); ch->d_unit_woff = ch->unit_woff.as<uint32_t>();
    ch->d_unit_kmers = ch->unit_kmers.as<uint32_t>();
    CU(cudaGetLastError());
    guard.ch = nullptr;
    c->chunks.push_back(ch);
    trace_host("flush_open_chunk: scatter queued");
    // sharded build over NVLink peer memory: the slices of this chunk leave for their owners now, on side streams, while
    // the next push is bucketed
    if (c->peer.connected && c->peer.world > 1) {
        CU(cudaEventRecord(c->peer.ev_scatter, st));
        TRY(peer_push_chunk(c, ch));
    }
    return 0;
}

int32_t mirror_chunk(ggcat_b200_ctx *c, Chunk *ch) {
    if (!ch->h_cnt.empty()) return 0;
    const size_t nu = ch->n_units;
    ch->h_cnt.resize(nu + 1); ch->h_off.resize(nu + 1); ch->h_words.resize(nu + 1); ch->h_woff.resize(nu + 1);
    ch->h_kmers.resize(nu + 1);
    if (ch->mirror_queued && ch->h_pin && !ch->imported) {
        CU(cudaEventSynchronize(ch->ev_mirror));
        memcpy(ch->h_cnt.data(), ch->h_pin, nu * 4);
        memcpy(ch->h_words.data(), ch->h_pin + nu, nu * 4);
        memcpy(ch->h_kmers.data(), ch->h_pin + 2 * nu, nu * 4);
        ch->n_segments = ch->h_pin[3 * nu];
    } else {
        CU(cudaMemcpyAsync(ch->h_cnt.data(), ch->d_unit_cnt, nu * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(ch->h_words.data(), ch->d_unit_words, nu * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(ch->h_kmers.data(), ch->d_unit_kmers, nu * 4, cudaMemcpyDeviceToHost, c->stream));
        uint32_t nseg = 0;
        if (!ch->imported) CU(cudaMemcpyAsync(&nseg, ch->d_unit_cnt + nu, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        ch->n_segments = nseg;
    }
    uint64_t a = 0, b = 0, km = 0;
    for (size_t u = 0; u < nu; u++) {
        ch->h_off[u] = (uint32_t)a; ch->h_woff[u] = (uint32_t)b;
        a += ch->h_cnt[u]; b += ch->h_words[u]; km += ch->h_kmers[u];
    }
    ch->h_off[nu] = (uint32_t)a; ch->h_woff[nu] = (uint32_t)b;
    ch->n_sk = a; ch->n_words = b; ch->n_kmers = km;
    return 0;
}


// ---- NVLink peer exchange, sender side (peer.cuh) ------------------------------------------------------------------
static uint32_t owner_first_bucket(uint32_t b1, uint32_t r, uint32_t world) {
    const uint64_t nb = 1ull << b1;
    if (r >= world) return (uint32_t)nb + 1;                      // the last rank also owns the duplicates bucket
    return (uint32_t)(((uint64_t)r * nb + world - 1) / world);    // owner(b) = b * world >> b1
}
static inline uint64_t peer_align16(uint64_t x) { return (x + 15) & ~15ull; }
static inline uint64_t peer_s_meta(uint32_t nu) { return peer_align16(3ull * nu * 4); }             // cnt | words | kmers of one slice
static inline uint64_t peer_s_uoff(uint32_t nu) { return 2 * peer_align16(((uint64_t)nu + 2) * 4); }  // receiver's two offset scans
static inline uint64_t peer_data_off(uint32_t S, uint32_t nu) { return (PEER_META_OFF + (uint64_t)S * (peer_s_meta(nu) + peer_s_uoff(nu)) + 255) & ~255ull; }
// staging block of one chunk: [meta jobs: 4 per rank][data jobs: 2 per rank][slice entries / headers: 64 B per rank]
static inline size_t peer_block_data_jobs_off(uint32_t W) { return (size_t)W * 4 * sizeof(PeerJob); }
static inline size_t peer_block_entries_off(uint32_t W) { return (size_t)W * 6 * sizeof(PeerJob); }
static inline size_t peer_block_bytes(uint32_t W) { return (size_t)W * (6 * sizeof(PeerJob) + 64); }

// First chunk of a build: new epoch; tell every peer that this rank is done with what it received in the last build
// (stream-ordered behind the last merge), and make both side streams wait until every destination has said the same.
int32_t peer_begin_build(ggcat_b200_ctx *c) {
    PeerState &ps = c->peer;
    const uint32_t W = ps.world, me = ps.rank;
    ps.epoch++; ps.build_started = true; ps.n_pushed = 0; ps.last_sent = ps.last_received = 0;
    for (uint32_t d = 0; d < W; d++) {
        const uint32_t nu = (owner_first_bucket(c->P.b1, d + 1, W) - owner_first_bucket(c->P.b1, d, W)) << c->P.b2;
        ps.cursor[d] = peer_data_off(ps.meta_slots, nu); ps.overflow[d] = false;
    }
    PeerHdrPtrs hp;
    memset(&hp, 0, sizeof(hp));
    for (uint32_t r = 0; r < W; r++) hp.h[r] = reinterpret_cast<PeerHdr *>(ps.peer_arena[r]);
    const unsigned long long timeout_ns = 20ull * 1000000000ull;
    c->fam_launches[F_PEER] += 3;
    k_peer_sync<<<1, PEER_MAX_WORLD, 0, c->stream>>>(hp, me, W, PEER_FLAG_RELEASED, ps.epoch - 1, 1u, 0u, ps.d_err.as<uint32_t>(), timeout_ns);
    k_peer_sync<<<1, PEER_MAX_WORLD, 0, ps.meta_stream>>>(hp, me, W, PEER_FLAG_RELEASED, ps.epoch - 1, 0u, 1u, ps.d_err.as<uint32_t>(), timeout_ns);
    k_peer_sync<<<1, PEER_MAX_WORLD, 0, ps.data_stream>>>(hp, me, W, PEER_FLAG_RELEASED, ps.epoch - 1, 0u, 1u, ps.d_err.as<uint32_t>(), timeout_ns);
    CU(cudaGetLastError());
    return 0;
}

// Push the slices of one local chunk to their owners.  The caller has recorded ps.ev_scatter on the compute stream behind
// the chunk's k_scatter.  Counts + slice entry go out on the meta stream at once (they are final after k_emit, which the
// host mirror of the counts has already waited for); descriptors + payload follow on the data stream behind ev_scatter.
int32_t peer_push_chunk(ggcat_b200_ctx *c, Chunk *ch) {
    PeerState &ps = c->peer;
    const uint32_t W = ps.world, me = ps.rank, j = ps.n_pushed;
    const DevParams &P = c->P;
    if (j >= ps.meta_slots)
        return set_err(GGCAT_B200_ERR_CAPACITY, "the NVLink exchange routes at most %u bucket chunks per build (GGCAT_B200_PEER_SLICES, max %d): "
                       "push larger batches", ps.meta_slots, PEER_MAX_SLICES);
    TRY(mirror_chunk(c, ch));
    trace_host("peer_push_chunk: counts of the chunk on the host");
    const size_t bb = peer_block_bytes(W);
    uint8_t *hb = ps.h_stage + (size_t)j * bb, *db = ps.d_stage.as<uint8_t>() + (size_t)j * bb;
    PeerJob *mjobs = reinterpret_cast<PeerJob *>(hb), *djobs = reinterpret_cast<PeerJob *>(hb + peer_block_data_jobs_off(W));
    PeerSlice *ent = reinterpret_cast<PeerSlice *>(hb + peer_block_entries_off(W));
    uint32_t nm = 0, nd = 0;
    for (uint32_t d = 0; d < W; d++) {
        if (d == me) continue;
        const uint32_t fu = owner_first_bucket(P.b1, d, W) << P.b2, nu = (owner_first_bucket(P.b1, d + 1, W) << P.b2) - fu;
        const uint32_t slot = (d + W - me - 1) % W;   // 0 .. W-2: the grid of k_peer_push is split over the destinations
        const uint64_t d0 = ch->h_off[fu], d1 = ch->h_off[fu + nu], w0 = ch->h_woff[fu], w1 = ch->h_woff[fu + nu];
        PeerSlice &e = ent[d];
        memset(&e, 0, sizeof(e));
        uint8_t *dst = ps.peer_arena[d] + PEER_HDR_BYTES + (uint64_t)me * ps.region_bytes;
        if (!ps.overflow[d]) {
            const uint64_t desc_off = peer_align16(ps.cursor[d]);
            const uint64_t pay_off = peer_align16(desc_off + (d1 - d0) * 16) + ((w0 * 4) & 15u);
            const uint64_t end = pay_off + (w1 - w0 + 8) * 4;
            if (end > ps.region_bytes) ps.overflow[d] = true;
            else {
                e.n_sk = d1 - d0; e.n_words = w1 - w0; e.word_bias = w0; e.desc_off = desc_off; e.pay_off = pay_off;
                ps.cursor[d] = end;
                if (e.n_sk) {
                    djobs[nd++] = {reinterpret_cast<const uint8_t *>(ch->d_desc + d0), dst + desc_off, e.n_sk * 16, slot, 0u};
                    djobs[nd++] = {reinterpret_cast<const uint8_t *>(ch->d_payload + w0), dst + pay_off, e.n_words * 4, slot, 0u};
                }
            }
        }
        uint8_t *meta = dst + PEER_META_OFF + (uint64_t)j * peer_s_meta(nu);
        mjobs[nm++] = {reinterpret_cast<const uint8_t *>(ch->d_unit_cnt + fu), meta, (uint64_t)nu * 4, slot, 0u};
        mjobs[nm++] = {reinterpret_cast<const uint8_t *>(ch->d_unit_words + fu), meta + (uint64_t)nu * 4, (uint64_t)nu * 4, slot, 0u};
        mjobs[nm++] = {reinterpret_cast<const uint8_t *>(ch->d_unit_kmers + fu), meta + (uint64_t)nu * 8, (uint64_t)nu * 4, slot, 0u};
        mjobs[nm++] = {db + peer_block_entries_off(W) + (size_t)d * 64, dst + PEER_TABLE_OFF + (uint64_t)j * 64, 64ull, slot, 0u};
    }
    for (uint32_t q = 0; q < nm; q++) ps.last_sent += mjobs[q].bytes;
    for (uint32_t q = 0; q < nd; q++) ps.last_sent += djobs[q].bytes;
    // the whole block travels on the meta stream; the data stream waits for it through ev_meta
    CU(cudaMemcpyAsync(db, hb, bb, cudaMemcpyHostToDevice, ps.meta_stream));
    CU(cudaEventRecord(ps.ev_meta, ps.meta_stream));
    k_peer_push<<<(W - 1) * 4, 256, 0, ps.meta_stream>>>(reinterpret_cast<const PeerJob *>(db), nm, W - 1);
    c->fam_launches[F_PEER_PUSH] += 1;
    CU(cudaStreamWaitEvent(ps.data_stream, ps.ev_meta, 0));
    CU(cudaStreamWaitEvent(ps.data_stream, ps.ev_scatter, 0));
    if (nd) {
        LaunchTimer t(c, F_PEER_PUSH, 1, ps.data_stream);   // device time of the bulk push (overlaps the bucketing of the next batch)
        k_peer_push<<<(unsigned)c->sm_count * 2, 256, 0, ps.data_stream>>>(reinterpret_cast<const PeerJob *>(db + peer_block_data_jobs_off(W)), nd, W - 1);
    }
    CU(cudaGetLastError());
    ps.n_pushed++;
    trace_host("peer_push_chunk: queued");
    return 0;
}

// exclusive scan u32 counts -> u64 offsets (n_units small: single CTA, sequential per thread chunks)
__global__ void __launch_bounds__(1024) k_scan_counts_u64(const uint32_t *cnt, uint64_t *off, uint32_t n, uint64_t base = 0) {
    __shared__ uint32_t s_scan[1024 / 32 + 2];
    uint64_t running = base;
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? cnt[i] : 0u;
        uint32_t tot;
        const uint32_t p = block_exclusive_scan<1024>(v, s_scan, &tot);
        if (i < n) off[i] = running + p;
        running += tot;
    }
    if (threadIdx.x == 0) off[n] = running;
}

constexpr int SM_THREADS_S = 512, SM_CAP_S = 6144;     // sort-mode (GGCAT_B200_MERGE=sort) capacities, 2 CTAs / SM
constexpr int SM_THREADS_L = 1024, SM_CAP_L = 12288;   // 1 CTA / SM
constexpr int GL_THREADS = 1024;
// k_merge_tier instantiations (merge.cuh): {threads, table slots, super-k-mers, payload words, landing buffers}
//   A  ~73 KB  3 CTAs / SM   typical units of a bacterial-size build (C2: ~4 k records, ~400 super-k-mers, ~1.2 k keys)
//   B  ~112 KB 2 CTAs / SM
//   C  ~221 KB 1 CTA  / SM   units of up to ~25 k records at 30x coverage (C2-sized slices on 8 GPUs: 16 k records)
constexpr int HASH_TS_S = 8192;                       // k_merge_parts: table slots of one key partition
#define TIER_A 512, 4096, 512, 2048, 1, 3      // one landing buffer + a sparser table measured faster than two buffers + 3072 slots (1.35 vs 1.43 ms on C2)
#define TIER_A1 512, 3072, 512, 2048, 2, 3     // A/B variant: two landing buffers, 3072 slots
#define TIER_A256 256, 3072, 512, 2048, 2, 3   // A/B variant: 8 warps per unit
#define TIER_B 512, 4608, 1408, 4608, 1, 2
#define TIER_C 1024, 10240, 2560, 8192, 1, 1
struct TierCap { uint32_t ts, skcap, pwcap; };
constexpr TierCap kTierCaps[3] = {{4096, 512, 2048}, {4608, 1408, 4608}, {10240, 2560, 8192}};
constexpr double kTierMaxLoad[3] = {0.5, 0.5, 0.72};   // expected distinct keys / table slots a unit may have in its tier (the last tier
                                                       // takes denser tables rather than sending the unit to the key-partition path)

// A bucket range may be merged in several parts that append to one final table (merge_range_parts): `eb` = entries
// already in the table, `ub` = units already in unit_final_off, `cap_total` = final-buffer capacity for the whole range.
struct PartBase { uint64_t eb = 0; uint32_t ub = 0; uint64_t cap_total = 0; };
int32_t pinned_reserve(uint8_t **p, size_t *cap, size_t need);

int32_t pinned_reserve(uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    const size_t want = need + need / 4 + 4096;
    if (cudaMallocHost((void **)p, want) != cudaSuccess) return set_err(GGCAT_B200_ERR_CUDA, "cudaMallocHost(%zu) failed", want);
    *cap = want;
    return 0;
}

// The final table is sized by the survivors actually seen, not by the k-mer occurrences (which are 10-25x more at
// 30x coverage): it starts from an estimate and grows (keeping the first `keep` entries) when a part does not fit.
int32_t final_reserve(ggcat_b200_ctx *c, uint64_t need, uint64_t keep, bool wide) {
    const bool with_src = c->wide_mode == MODE_RK128;
    if (need <= c->fin_cap && c->out_keys2.p && c->out_cf2.p && (!wide || c->out_hi2.p) && (!with_src || c->out_src2.p)) return 0;
    CU(cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CU(cudaStreamSynchronize(c->copy_stream));   // copies of earlier parts read the old buffers
    auto grow = [&](DevBuf &b, size_t elem) -> cudaError_t {
        DevBuf nb;
        cudaError_t e = nb.reserve(need * elem);
        if (e != cudaSuccess) return e;
        if (keep && b.p) { e = cudaMemcpy(nb.p, b.p, keep * elem, cudaMemcpyDeviceToDevice); if (e != cudaSuccess) { nb.release(); return e; } }
        b.release();
        b = nb;
        return cudaSuccess;
    };
    CU(grow(c->out_keys2, 8));
    CU(grow(c->out_cf2, 4));
    if (wide) CU(grow(c->out_hi2, 8));
    if (with_src) CU(grow(c->out_src2, 8 * (size_t)c->src_words));
    c->fin_cap = need;
    return 0;
}
uint64_t final_estimate(const ggcat_b200_ctx *c, uint64_t records) {
    if (const char *e = getenv("GGCAT_B200_FINAL_EST")) return std::max<uint64_t>(strtoull(e, nullptr, 10), 1);  // tests: force the growth path
    const uint64_t est = c->final_hint ? c->final_hint + c->final_hint / 4 + 65536 : std::max<uint64_t>(records / 4, 1ull << 20);
    return std::max<uint64_t>(std::min(est, records), 1);
}

int32_t merge_range_device_wide(ggcat_b200_ctx *c, uint32_t first_bucket, uint32_t n_buckets, uint64_t *n_entries,
                                uint64_t *unique, uint64_t *total, PartBase pb);

constexpr uint32_t PART_TARGET = 4096;    // smallest partition target (records): fits the 8192-slot shared table whatever the data
constexpr int FIN_THREADS = 512, FIN_BCAP = 6144;

int32_t merge_range_device(ggcat_b200_ctx *c, uint32_t first_bucket, uint32_t n_buckets, uint64_t *n_entries,
                           uint64_t *unique, uint64_t *total, PartBase pb = PartBase()) {
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    if (!c->finished) return set_err(GGCAT_B200_ERR_STATE, "merge before finish_bucketing");
    const uint32_t nb_total = (1u << P.b1) + 1;
    if (n_buckets == 0 || first_bucket >= nb_total || first_bucket + n_buckets > nb_total)
        return set_err(GGCAT_B200_ERR_INVALID, "bucket range [%u,+%u) outside 0..%u", first_bucket, n_buckets, nb_total);
    if (c->wide_mode >= 0) return merge_range_device_wide(c, first_bucket, n_buckets, n_entries, unique, total, pb);
    trace_host("merge_range_device: enter");
    const uint32_t u0 = first_bucket << P.b2, nu = n_buckets << P.b2;
    const bool hash_mode = c->merge_mode == 1;
    // ---- classify units
    //   tier[0..2]  units that fit one staging round of k_merge_tier (super-k-mers, payload words, expected distinct keys)
    //   big         <= PART_MAXP * part_target records: key partitions in HBM, one shared-table CTA per partition
    //   work[2]     giant units (and every large unit in sort mode): table / sort buffers in a global scratch slice
    //   sort mode (GGCAT_B200_MERGE=sort): work[0] <= 6144 records, work[1] <= 12288
    std::vector<uint32_t> work[3], tier[3];
    std::vector<std::pair<uint64_t, uint32_t>> large, big;  // (records, unit)
    if (c->unit_tot.size() < nu) c->unit_tot.resize(nu);
    ggcat_b200_ctx::UnitTot *ut = c->unit_tot.data();
    memset(ut, 0, (size_t)nu * sizeof(*ut));
    // key partitions are sized by the distinct keys they should hold (~2048 = a quarter of the 8192-slot table), from the
    // distinct/records ratio of the parts merged so far with a 2x margin; k_merge_parts splits a partition that
    // turns out fuller than that.  Unknown ratio (first part of a context): 4096 records, the table's guaranteed capacity.
    uint32_t part_target = PART_TARGET;
    if (!c->part_fixed) {
        const double r = std::min(1.0, 2.0 * c->distinct_ratio);
        part_target = (uint32_t)std::min<double>(65536.0, std::max<double>(PART_TARGET, 2048.0 / std::max(r, 1e-3)));
    }
    const uint32_t part_cap = part_target + part_target / 2;
    // expected distinct keys per k-mer record (15 % margin): routes a unit to the smallest tier whose table stays sparse
    const double keys_per_rec = std::min(1.0, 1.15 * c->distinct_ratio);
    TierCap caps[3] = {kTierCaps[0], kTierCaps[1], kTierCaps[2]};
    if (c->tier_a_variant != 0) caps[0].ts = 3072;
    for (Chunk *ch : c->chunks) {
        const uint32_t lo = std::max(u0, ch->first_unit), hi = std::min(u0 + nu, ch->first_unit + ch->n_units);
        if (lo >= hi) continue;
        const uint32_t *hc = ch->h_cnt.data() + (lo - ch->first_unit), *hk = ch->h_kmers.data() + (lo - ch->first_unit),
                       *hw = ch->h_words.data() + (lo - ch->first_unit);
        ggcat_b200_ctx::UnitTot *t = ut + (lo - u0);
        for (uint32_t i = 0, e = hi - lo; i < e; i++) {
            const uint32_t sk = hc[i];
            t[i].n += hk[i]; t[i].sk += sk; t[i].w += hw[i]; t[i].sl += sk ? 1u : 0u;   // a slice without super-k-mers adds zeros
        }
    }
    for (int q = 0; q < 3; q++) tier[q].reserve(nu);
    const bool tiers_ok = hash_mode && c->chunks.size() <= (size_t)TIER_MAXSL && !c->no_tiers;
    uint64_t tot_kmers = 0, tier_nmax = 0;
    for (uint32_t u = u0; u < u0 + nu; u++) {
        const uint64_t n = ut[u - u0].n;
        if (n == 0) continue;
        if (n >= (1ull << 31)) return set_err(GGCAT_B200_ERR_INVALID, "unit %u holds %llu k-mers (> 2^31)", u, (unsigned long long)n);
        tot_kmers += n;
        int t = -1;
        if (tiers_ok && n < (1u << 20)) {
            const double need_keys = (double)n * keys_per_rec;
            const uint64_t need_w = (uint64_t)ut[u - u0].w + 6ull * ut[u - u0].sl;
            for (int q = 0; q < 3 && t < 0; q++)
                if (ut[u - u0].sk <= caps[q].skcap && need_w <= caps[q].pwcap && need_keys <= kTierMaxLoad[q] * caps[q].ts) t = q;
        }
        if (t >= 0) { tier[t].push_back(u); tier_nmax = std::max(tier_nmax, n); }
        else if (!hash_mode && n <= SM_CAP_S) work[0].push_back(u);
        else if (!hash_mode && n <= SM_CAP_L) work[1].push_back(u);
        else if (hash_mode && n <= (uint64_t)PART_MAXP * part_target && !c->no_partition) big.push_back({n, u});
        else large.push_back({n, u});
    }
    trace_host("merge_range_device: units classified");
    // ---- key partitions of big units and the output-slot map
    std::sort(big.begin(), big.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
    std::vector<uint32_t> big_unit, big_logp, big_pbase, part_big;
    uint32_t n_parts = 0;
    const uint32_t n_slots = nu;   // one output slot per unit
    if (!big.empty()) {
        // every unit keeps ONE output slot: the partitions of a big unit append into its region (unit_out_cnt is the fill counter)
        for (auto &pr : big) {
            uint32_t lp = 1;
            while (((uint64_t)part_target << lp) < pr.first) lp++;
            big_unit.push_back(pr.second); big_logp.push_back(lp); big_pbase.push_back(n_parts);
            for (uint32_t q = 0; q < (1u << lp); q++) part_big.push_back((uint32_t)big_unit.size() - 1);
            n_parts += 1u << lp;
        }
    }
    // ---- giant / sort-mode large units: biggest first (one CTA each), two tiers so that ordinary large units do not
    //      inherit the per-CTA scratch size of a giant one
    std::sort(large.begin(), large.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
    const uint64_t TIER = 1ull << 20;
    size_t n_giant = 0;
    while (n_giant < large.size() && large[n_giant].first > TIER) n_giant++;
    for (auto &pr : large) work[2].push_back(pr.second);
    struct Tier { const uint32_t *wl; size_t count; const uint32_t *count_dev; uint64_t per_cta; unsigned grid; };
    auto make_tier = [&](const uint32_t *wl, size_t count, const uint32_t *count_dev, uint64_t nmax) {
        const uint64_t per_cta = std::max<uint64_t>((uint64_t)hash_table_slots((uint32_t)nmax) * 3 / 2, 2 * nmax) + 16;  // table keys + counters, or 2n records for the sort variant
        const uint64_t budget = 12ull << 30;
        const uint64_t g = std::min<uint64_t>(std::min<uint64_t>(count, (uint64_t)c->sm_count * 2), std::max<uint64_t>(1, budget / (per_cta * 8)));
        return Tier{wl, count, count_dev, per_cta, (unsigned)g};
    };
    // ---- uploads: chunk views, work lists, per-unit record counts and output regions travel in ONE copy from a pinned
    //      staging buffer (every merge call ends with a stream synchronisation, so the buffer is free again)
    std::vector<ChunkView> views;
    for (Chunk *ch : c->chunks) {
        ChunkView v;
        v.desc = ch->d_desc; v.payload = ch->d_payload; v.unit_off = ch->d_unit_off; v.unit_woff = ch->d_unit_woff; v.unit_kmers = ch->d_unit_kmers;
        v.first_unit = ch->first_unit; v.n_units = ch->n_units; v.word_bias = ch->word_bias; v.pad = 0;
        views.push_back(v);
    }
    auto al16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t off_views = 0, off_t0 = al16(off_views + std::max<size_t>(1, views.size()) * sizeof(ChunkView));
    const size_t off_t1 = al16(off_t0 + tier[0].size() * 4), off_t2 = al16(off_t1 + tier[1].size() * 4);
    const size_t off_w0 = al16(off_t2 + tier[2].size() * 4), off_w1 = al16(off_w0 + work[0].size() * 4);
    const size_t off_w2 = al16(off_w1 + work[1].size() * 4), off_un = al16(off_w2 + work[2].size() * 4);
    const size_t off_so = al16(off_un + (size_t)nu * 4), stage_bytes = al16(off_so + ((size_t)nu + 1) * 8);
    TRY(pinned_reserve(&c->h_mstage, &c->h_mstage_cap, stage_bytes));
    CU(c->d_mstage.reserve(stage_bytes));
    {
        uint8_t *h = c->h_mstage;
        if (!views.empty()) memcpy(h + off_views, views.data(), views.size() * sizeof(ChunkView));
        for (int q = 0; q < 3; q++)
            if (!tier[q].empty()) memcpy(h + (q == 0 ? off_t0 : q == 1 ? off_t1 : off_t2), tier[q].data(), tier[q].size() * 4);
        for (int q = 0; q < 3; q++)
            if (!work[q].empty()) memcpy(h + (q == 0 ? off_w0 : q == 1 ? off_w1 : off_w2), work[q].data(), work[q].size() * 4);
        uint32_t *un = reinterpret_cast<uint32_t *>(h + off_un);
        uint64_t *so = reinterpret_cast<uint64_t *>(h + off_so);
        uint64_t acc = 0;   // every unit owns the region [static_off[u], static_off[u] + records(u)) of the part's output buffers
        for (uint32_t i = 0; i < nu; i++) { un[i] = (uint32_t)ut[i].n; so[i] = acc; acc += ut[i].n; }
        so[nu] = acc;
        CU(cudaMemcpyAsync(c->d_mstage.p, h, stage_bytes, cudaMemcpyHostToDevice, st));
    }
    uint8_t *dm = c->d_mstage.as<uint8_t>();
    const uint32_t *d_tier[3] = {reinterpret_cast<const uint32_t *>(dm + off_t0), reinterpret_cast<const uint32_t *>(dm + off_t1),
                                 reinterpret_cast<const uint32_t *>(dm + off_t2)};
    const uint32_t *d_w[3] = {reinterpret_cast<const uint32_t *>(dm + off_w0), reinterpret_cast<const uint32_t *>(dm + off_w1),
                              reinterpret_cast<const uint32_t *>(dm + off_w2)};
    uint32_t *d_big_unit = nullptr, *d_big_logp = nullptr, *d_big_pbase = nullptr, *d_big_ovf = nullptr, *d_part_big = nullptr,
             *d_pcount = nullptr;
    const uint32_t *d_slot_of_unit = nullptr;   // every unit has one output slot (the finish kernels also accept a slot map)
    if (!big.empty()) {
        const size_t nbig = big.size();
        // one metadata buffer: big_unit | big_logp | big_pbase | big_ovf | part_big | pcount
        const size_t words = 4 * nbig + 2 * (size_t)n_parts;
        CU(c->d_partmeta.reserve(words * 4));
        uint32_t *base = c->d_partmeta.as<uint32_t>();
        d_big_unit = base; d_big_logp = base + nbig; d_big_pbase = base + 2 * nbig; d_big_ovf = base + 3 * nbig;
        d_part_big = base + 4 * nbig; d_pcount = d_part_big + n_parts;
        CU(cudaMemcpyAsync(d_big_unit, big_unit.data(), nbig * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_big_logp, big_logp.data(), nbig * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_big_pbase, big_pbase.data(), nbig * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(d_big_ovf, 0, nbig * 4, st));
        CU(cudaMemcpyAsync(d_part_big, part_big.data(), (size_t)n_parts * 4, cudaMemcpyHostToDevice, st));
        CU(c->d_recs.reserve((size_t)n_parts * part_cap * 8));
    }
    std::vector<Tier> tiers;
    uint64_t scratch_u64 = 1;
    if (n_giant) tiers.push_back(make_tier(d_w[2], n_giant, nullptr, large[0].first));
    if (large.size() > n_giant) tiers.push_back(make_tier(d_w[2] + n_giant, large.size() - n_giant, nullptr, large[n_giant].first));
    const uint64_t cap = std::max<uint64_t>(tot_kmers, 1);
    CU(c->out_keys.reserve(cap * 8)); CU(c->out_cf.reserve(cap * 4));
    CU(c->cursor.reserve(64)); CU(c->overflow.reserve(16));
    CU(c->d_retry.reserve(((size_t)nu + 2) * 4));  // [0] = count, [1..] = unit ids
    uint32_t *retry3_cnt = c->d_retry.as<uint32_t>(), *retry3 = retry3_cnt + 1;   // big units with an overflowed partition
    // units that come back: big units with an overflowed partition, tier units whose table filled up
    const size_t n_tier_units = tier[0].size() + tier[1].size() + tier[2].size();
    if (!big.empty() || n_tier_units)
        tiers.push_back(make_tier(retry3, big.size() + n_tier_units, retry3_cnt, std::max<uint64_t>(big.empty() ? 0 : big[0].first, tier_nmax)));
    for (const Tier &tr : tiers) scratch_u64 = std::max(scratch_u64, tr.per_cta * tr.grid);
    CU(c->d_scratch.reserve(scratch_u64 * 8));
    CU(cudaMemsetAsync(retry3_cnt, 0, 4, st));
    CU(c->unit_out_off.reserve(((size_t)n_slots + 1) * 8)); CU(c->unit_out_cnt.reserve(((size_t)n_slots + 1) * 4));
    if (pb.ub == 0) CU(c->unit_final_off.reserve(((size_t)c->P.n_units + 2) * 8));
    CU(cudaMemsetAsync(c->cursor.p, 0, 64, st));
    CU(cudaMemsetAsync(c->overflow.p, 0, 16, st));
    CU(cudaMemsetAsync(c->unit_out_off.p, 0, ((size_t)n_slots + 1) * 8, st));
    CU(cudaMemsetAsync(c->unit_out_cnt.p, 0, ((size_t)n_slots + 1) * 4, st));
    trace_host("merge_range_device: uploads queued");
    const uint32_t *d_unit_n = reinterpret_cast<const uint32_t *>(dm + off_un);
    MergeOut out;
    out.keys = c->out_keys.as<uint64_t>(); out.count_flags = c->out_cf.as<uint32_t>();
    out.cursor = c->cursor.as<unsigned long long>(); out.unit_out_off = c->unit_out_off.as<uint64_t>();
    out.unit_out_cnt = c->unit_out_cnt.as<uint32_t>(); out.overflow = c->overflow.as<uint32_t>();
    out.static_off = reinterpret_cast<const uint64_t *>(dm + off_so);
    out.slot_of_unit = d_slot_of_unit;
    const ChunkView *dv = reinterpret_cast<const ChunkView *>(dm + off_views);
    const uint32_t nch = (uint32_t)views.size();
    const uint32_t ms = c->params.min_multiplicity;
    if (hash_mode) {
        // work counters of the three tier launches live behind the statistics in `cursor` (zeroed above)
        uint32_t *wc = reinterpret_cast<uint32_t *>(c->cursor.as<unsigned long long>() + 4);
        auto launch_tier = [&](auto kern, size_t smem, int threads, int ctas_per_sm, int q) -> int32_t {
            if (tier[q].empty()) return 0;
            LaunchTimer t(c, F_MERGE_HASH);
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned grid = (unsigned)std::min<size_t>(tier[q].size(), (size_t)c->sm_count * ctas_per_sm);
            kern<<<grid, threads, smem, st>>>(dv, nch, d_tier[q], (uint32_t)tier[q].size(), u0, P, ms, out, d_unit_n, wc + q, retry3, retry3_cnt);
            return 0;
        };
        if (c->tier_a_variant == 2) TRY(launch_tier(k_merge_tier<TIER_A256>, TierSmem<256, 3072, 512, 2048, 2>::bytes, 256, 3, 0));
        else if (c->tier_a_variant == 1) TRY(launch_tier(k_merge_tier<TIER_A1>, TierSmem<512, 3072, 512, 2048, 2>::bytes, 512, 3, 0));
        else TRY(launch_tier(k_merge_tier<TIER_A>, TierSmem<512, 4096, 512, 2048, 1>::bytes, 512, 3, 0));
        TRY(launch_tier(k_merge_tier<TIER_B>, TierSmem<512, 4608, 1408, 4608, 1>::bytes, 512, 2, 1));
        TRY(launch_tier(k_merge_tier<TIER_C>, TierSmem<1024, 10240, 2560, 8192, 1>::bytes, 1024, 1, 2));
        if (!big.empty()) {
            {
                LaunchTimer t(c, F_PARTITION);
                const unsigned grid = (unsigned)std::min<size_t>(big.size(), (size_t)c->sm_count * 2);
                k_partition_units<1024><<<grid, 1024, 0, st>>>(dv, nch, d_big_unit, d_big_logp, d_big_pbase, (uint32_t)big.size(), P,
                                                               c->d_recs.as<uint64_t>(), d_pcount, part_cap, d_big_ovf, retry3, retry3_cnt);
            }
            {
                LaunchTimer t(c, F_MERGE_HASH_PART);
                PartSrc ps;
                ps.recs = c->d_recs.as<uint64_t>(); ps.pcount = d_pcount; ps.part_big = d_part_big;
                ps.big_ovf = d_big_ovf; ps.big_unit = d_big_unit; ps.pcap = part_cap; ps.pad = 0;
                auto kern = k_merge_parts<SM_THREADS_S, HASH_TS_S>;
                const size_t smem = merge_parts_smem_bytes<SM_THREADS_S, HASH_TS_S>();
                CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const unsigned grid = (unsigned)std::min<size_t>(n_parts, (size_t)c->sm_count * 2 * 8);
                kern<<<grid, SM_THREADS_S, smem, st>>>(n_parts, u0, ms, out, ps);
            }
        }
    }
    if (!work[0].empty()) {
        LaunchTimer t(c, F_MERGE_SMEM);
        auto kern = k_merge_units<SM_THREADS_S, SM_CAP_S, false>;
        const size_t smem = merge_smem_bytes<SM_THREADS_S, SM_CAP_S>(false);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)std::min<size_t>(work[0].size(), (size_t)c->sm_count * 2 * 8);
        kern<<<grid, SM_THREADS_S, smem, st>>>(dv, nch, d_w[0], (uint32_t)work[0].size(), u0, P, ms, out, nullptr, 0, nullptr);
    }
    if (!work[1].empty()) {
        LaunchTimer t(c, F_MERGE_SMEM);
        auto kern = k_merge_units<SM_THREADS_L, SM_CAP_L, false>;
        const size_t smem = merge_smem_bytes<SM_THREADS_L, SM_CAP_L>(false);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)std::min<size_t>(work[1].size(), (size_t)c->sm_count * 8);
        kern<<<grid, SM_THREADS_L, smem, st>>>(dv, nch, d_w[1], (uint32_t)work[1].size(), u0, P, ms, out, nullptr, 0, nullptr);
    }
    for (const Tier &tr : tiers) {
        auto sortk = k_merge_units<GL_THREADS, 0, true>;
        const size_t smem_sort = merge_smem_bytes<GL_THREADS, 0>(true);
        if (hash_mode) {
            LaunchTimer t(c, F_MERGE_HASH_GLOBAL);
            auto kern = k_merge_hash<GL_THREADS, 0>;
            kern<<<tr.grid, GL_THREADS, merge_hash_smem_bytes<GL_THREADS, 0>(), st>>>(
                dv, nch, tr.wl, (uint32_t)tr.count, u0, P, ms, out, d_unit_n, c->d_scratch.as<uint64_t>(), tr.per_cta, tr.count_dev);
        } else {
            LaunchTimer t(c, F_MERGE_GLOBAL);
            sortk<<<tr.grid, GL_THREADS, smem_sort, st>>>(dv, nch, tr.wl, (uint32_t)tr.count, u0, P, ms, out, c->d_scratch.as<uint64_t>(),
                                                         tr.per_cta, nullptr);
        }
    }
    CU(cudaGetLastError());
    trace_host("merge_range_device: merge kernels launched");
    // ---- unit-ordered final layout (parts append at entry pb.eb / unit pb.ub)
    const uint64_t rec_total = std::max(cap, pb.cap_total);
    if (pb.ub == 0) TRY(final_reserve(c, final_estimate(c, rec_total), 0, false));
    CU(c->fin_tmp_keys.reserve(cap * 8)); CU(c->fin_tmp_cf.reserve(cap * 4));   // units sorted in global memory (> FIN_BCAP survivors)
    uint32_t ovf = 0;
    for (int attempt = 0;; attempt++) {
        {
            LaunchTimer t(c, F_GATHER, 3);
            uint64_t *foff = c->unit_final_off.as<uint64_t>() + pb.ub;
            k_scan_unit_slots<<<1, 1024, 0, st>>>(c->unit_out_cnt.as<uint32_t>(), d_slot_of_unit, foff, nu, pb.eb);
            FinishArgs fa;
            fa.src_keys = c->out_keys.as<uint64_t>(); fa.src_cf = c->out_cf.as<uint32_t>();
            fa.slot_off = c->unit_out_off.as<uint64_t>(); fa.slot_cnt = c->unit_out_cnt.as<uint32_t>();
            fa.slot_of_unit = d_slot_of_unit; fa.dst_off = foff;
            fa.dst_keys = c->out_keys2.as<uint64_t>(); fa.dst_cf = c->out_cf2.as<uint32_t>();
            fa.tmp_keys = c->fin_tmp_keys.as<uint64_t>(); fa.tmp_cf = c->fin_tmp_cf.as<uint32_t>();
            fa.n_units = nu; fa.kbits = 2 * P.k; fa.capacity = c->fin_cap; fa.overflow = c->overflow.as<uint32_t>();
            k_finish_small<<<(unsigned)std::min<uint32_t>((nu + FIN_WARPS - 1) / FIN_WARPS, (uint32_t)c->sm_count * 16), FIN_WARPS * 32, 0, st>>>(fa);
            auto kern = k_finish_units<FIN_THREADS, FIN_BCAP>;
            const size_t smem = finish_units_smem_bytes<FIN_THREADS, FIN_BCAP>();
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)std::min<uint32_t>(nu, (uint32_t)c->sm_count * 2 * 4), FIN_THREADS, smem, st>>>(fa);
        }
        CU(cudaMemcpyAsync(c->h_pinned, c->cursor.p, 24, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(c->h_pinned + 8, c->overflow.p, 4, cudaMemcpyDeviceToHost, st));
        c->h_pinned[9] = 0;
        if (c->peer.connected && c->peer.world > 1) CU(cudaMemcpyAsync(c->h_pinned + 9, c->peer.d_err.p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        if ((uint32_t)c->h_pinned[9]) {   // a peer never signalled that its bulk push was complete: what was merged is incomplete
            cudaMemset(c->peer.d_err.p, 0, 16);
            return set_err(GGCAT_B200_ERR_STATE, "NVLink exchange: rank %u did not complete its push within 20 s", (uint32_t)c->h_pinned[9] - 1);
        }
        ovf = (uint32_t)c->h_pinned[8];
        if (ovf == 4u && attempt == 0) {   // the part's survivors do not fit the final table: grow it, gather again
            const uint64_t need = pb.eb + c->h_pinned[0];
            TRY(final_reserve(c, std::max(need, std::min(rec_total, need + need / 2)), pb.eb, false));
            CU(cudaMemsetAsync(c->overflow.p, 0, 16, st));
            continue;
        }
        break;
    }
    if (ovf) return set_err(GGCAT_B200_ERR_CAPACITY, "merge output overflow (code %u)", ovf);
    c->last_entries = c->h_pinned[0];
    c->fin = FinalTable();
    c->fin.keys_lo = c->out_keys2.as<uint64_t>(); c->fin.cf = c->out_cf2.as<uint32_t>();
    c->fin.unit_off = c->unit_final_off.as<uint64_t>(); c->fin.n_entries = pb.eb + c->h_pinned[0];
    if (c->h_pinned[2] >= (1ull << 20)) c->distinct_ratio = (double)c->h_pinned[1] / (double)c->h_pinned[2];
    if (n_entries) *n_entries = c->h_pinned[0];
    if (unique) *unique = c->h_pinned[1];
    if (total) *total = c->h_pinned[2];
    return 0;
}

// ---- wide path (merge128.cuh): 128-bit keys and coloured builds --------------------------------------------------
constexpr int W_THREADS_S = 512, W_TS_S = 4096;    // 80 KB table: 2 CTAs / SM, units <= 3072 records
constexpr int W_THREADS_L = 1024, W_TS_L = 8192;   // 160 KB table: 1 CTA / SM, units <= 6144 records
constexpr int W_SORT_THREADS = 512, W_SORT_CAP = 2048;

template <int MODE>
int32_t launch_hash128(ggcat_b200_ctx *c, const ChunkView *dv, uint32_t nch, std::vector<uint32_t> *work, uint32_t u0,
                       const MergeOut128 &out, const std::vector<std::pair<uint64_t, uint32_t>> &large, uint32_t *retry_base) {
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    if (!work[0].empty()) {
        LaunchTimer t(c, F_MERGE_HASH128);
        auto kern = k_merge_hash128<W_THREADS_S, W_TS_S, MODE>;
        const size_t smem = merge_hash128_smem_bytes<W_THREADS_S, W_TS_S, MODE>();
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)std::min<size_t>(work[0].size(), (size_t)c->sm_count * 2 * 8);
        kern<<<grid, W_THREADS_S, smem, st>>>(dv, nch, c->d_work[0].as<uint32_t>(), (uint32_t)work[0].size(), u0, P, c->rk,
                                               c->params.min_multiplicity, out, nullptr, 0, PartSrc128(), nullptr, retry_base);
    }
    if (!work[1].empty()) {
        LaunchTimer t(c, F_MERGE_HASH128);
        auto kern = k_merge_hash128<W_THREADS_L, W_TS_L, MODE>;
        const size_t smem = merge_hash128_smem_bytes<W_THREADS_L, W_TS_L, MODE>();
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)std::min<size_t>(work[1].size(), (size_t)c->sm_count * 8);
        kern<<<grid, W_THREADS_L, smem, st>>>(dv, nch, c->d_work[1].as<uint32_t>(), (uint32_t)work[1].size(), u0, P, c->rk,
                                               c->params.min_multiplicity, out, nullptr, 0, PartSrc128(), nullptr, retry_base);
    }
    // large units: table in a per-CTA slice of global scratch, biggest first, two tiers (see merge_range_device)
    const uint64_t TIER = 1ull << 20;
    size_t n_giant = 0;
    while (n_giant < large.size() && large[n_giant].first > TIER) n_giant++;
    for (int tr = 0; tr < 2; tr++) {
        const size_t first = tr == 0 ? 0 : n_giant, count = tr == 0 ? n_giant : large.size() - n_giant;
        if (!count) continue;
        const uint64_t nmax = large[first].first;
        uint64_t per_cta = ((uint64_t)hash_table_slots_pow2((uint32_t)nmax) * slot_bytes128<MODE>() + 15) / 16 * 2 + 2;  // u64 words, 16-byte multiple
        const uint64_t budget = 12ull << 30;
        const uint64_t g = std::min<uint64_t>(std::min<uint64_t>(count, (uint64_t)c->sm_count * 2), std::max<uint64_t>(1, budget / (per_cta * 8)));
        CU(c->d_scratch.reserve(per_cta * g * 8));
        LaunchTimer t(c, F_MERGE_HASH128);
        auto kern = k_merge_hash128<W_THREADS_L, 0, MODE>;
        kern<<<(unsigned)g, W_THREADS_L, merge_hash128_smem_bytes<W_THREADS_L, 0, MODE>(), st>>>(
            dv, nch, c->d_work[2].as<uint32_t>() + first, (uint32_t)count, u0, P, c->rk, c->params.min_multiplicity, out,
            c->d_scratch.as<uint64_t>(), per_cta, PartSrc128(), nullptr, nullptr);
    }
    CU(cudaGetLastError());
    return 0;
}

// Big units of the wide path: key partitions in HBM (k_partition_units128), one shared-table CTA per partition, and the
// global-table kernel for units whose partitions overflowed.
struct BigPlan128 {
    std::vector<uint32_t> unit, logp, pbase, part_big;
    std::vector<uint64_t> off;     // static output region of every big unit
    uint32_t n_parts = 0;
    uint64_t nmax = 0;
};
constexpr uint32_t W_PART_CAP = 3072, W_PART_TARGET = 2048;   // a partition fits the 4096-slot shared table

template <int MODE>
int32_t launch_partitions128(ggcat_b200_ctx *c, const ChunkView *dv, uint32_t nch, uint32_t u0, const MergeOut128 &out,
                             const BigPlan128 &bp, uint32_t *retry_base) {
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    const size_t nbig = bp.unit.size();
    if (!nbig) return 0;
    // metadata: big_unit | big_logp | big_pbase | big_ovf | big_fill | part_big | pcount   (u32), big_off (u64)
    const size_t words = 5 * nbig + 2 * (size_t)bp.n_parts;
    CU(c->d_partmeta.reserve(words * 4));
    CU(c->d_static_off.reserve(nbig * 8));
    uint32_t *base = c->d_partmeta.as<uint32_t>();
    uint32_t *d_unit = base, *d_logp = base + nbig, *d_pbase = base + 2 * nbig, *d_ovf = base + 3 * nbig, *d_fill = base + 4 * nbig;
    uint32_t *d_part_big = base + 5 * nbig, *d_pcount = d_part_big + bp.n_parts;
    CU(cudaMemcpyAsync(d_unit, bp.unit.data(), nbig * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_logp, bp.logp.data(), nbig * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_pbase, bp.pbase.data(), nbig * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_ovf, 0, 2 * nbig * 4, st));   // big_ovf and big_fill
    CU(cudaMemcpyAsync(d_part_big, bp.part_big.data(), (size_t)bp.n_parts * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_static_off.p, bp.off.data(), nbig * 8, cudaMemcpyHostToDevice, st));
    const size_t nrec = (size_t)bp.n_parts * W_PART_CAP;
    CU(c->d_recs.reserve(nrec * 16));
    CU(c->d_recfl.reserve(nrec));
    if (MODE == MODE_RK128) CU(c->d_recsrc.reserve(nrec * 8));
    uint32_t *retry_cnt = retry_base, *retry = retry_base + 1;   // shared with the shared-table kernels (merge_range_device_wide)
    uint64_t *rec_lo = c->d_recs.as<uint64_t>(), *rec_hi = rec_lo + nrec;
    {
        LaunchTimer t(c, F_PARTITION);
        const unsigned grid = (unsigned)std::min<size_t>(nbig, (size_t)c->sm_count * 2);
        k_partition_units128<1024, MODE><<<grid, 1024, 0, st>>>(dv, nch, d_unit, d_logp, d_pbase, (uint32_t)nbig, P, c->rk, rec_lo, rec_hi,
                                                                c->d_recfl.as<uint8_t>(), MODE == MODE_RK128 ? c->d_recsrc.as<uint64_t>() : nullptr,
                                                                d_pcount, W_PART_CAP, d_ovf, retry, retry_cnt);
    }
    {
        LaunchTimer t(c, F_MERGE_HASH_PART);
        PartSrc128 ps;
        ps.rec_lo = rec_lo; ps.rec_hi = rec_hi; ps.rec_fl = c->d_recfl.as<uint8_t>(); ps.pcount = d_pcount; ps.part_big = d_part_big;
        ps.rec_src = MODE == MODE_RK128 ? c->d_recsrc.as<uint64_t>() : nullptr;
        ps.big_unit = d_unit; ps.big_ovf = d_ovf; ps.big_off = c->d_static_off.as<uint64_t>(); ps.big_fill = d_fill;
        ps.pcap = W_PART_CAP; ps.pad = 0;
        auto kern = k_merge_hash128<W_THREADS_S, W_TS_S, MODE, SRC_RECORDS>;
        const size_t smem = merge_hash128_smem_bytes<W_THREADS_S, W_TS_S, MODE>();
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)std::min<size_t>(bp.n_parts, (size_t)c->sm_count * 2 * 8);
        kern<<<grid, W_THREADS_S, smem, st>>>(dv, nch, nullptr, bp.n_parts, u0, P, c->rk, c->params.min_multiplicity, out, nullptr, 0, ps, nullptr, nullptr);
    }
    CU(cudaGetLastError());
    return 0;
}

// Units that came back (a partition overflowed, or a shared table sized from the expected distinct keys filled up):
// global-table kernel over the device-side retry list.
template <int MODE>
int32_t launch_retry128(ggcat_b200_ctx *c, const ChunkView *dv, uint32_t nch, uint32_t u0, const MergeOut128 &out, uint32_t *retry_base,
                        uint64_t n_candidates, uint64_t nmax) {
    if (!n_candidates) return 0;
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    uint64_t per_cta = ((uint64_t)hash_table_slots_pow2((uint32_t)nmax) * slot_bytes128<MODE>() + 15) / 16 * 2 + 2;
    const uint64_t budget = 12ull << 30;
    const uint64_t g = std::min<uint64_t>(std::min<uint64_t>(n_candidates, (uint64_t)c->sm_count * 2), std::max<uint64_t>(1, budget / (per_cta * 8)));
    CU(c->d_scratch.reserve(per_cta * g * 8));
    LaunchTimer t(c, F_MERGE_HASH128);
    auto kern = k_merge_hash128<W_THREADS_L, 0, MODE>;
    kern<<<(unsigned)g, W_THREADS_L, merge_hash128_smem_bytes<W_THREADS_L, 0, MODE>(), st>>>(
        dv, nch, retry_base + 1, (uint32_t)n_candidates, u0, P, c->rk, c->params.min_multiplicity, out, c->d_scratch.as<uint64_t>(), per_cta,
        PartSrc128(), retry_base, nullptr);
    CU(cudaGetLastError());
    return 0;
}

int32_t merge_range_device_wide(ggcat_b200_ctx *c, uint32_t first_bucket, uint32_t n_buckets, uint64_t *n_entries,
                                uint64_t *unique, uint64_t *total, PartBase pb) {
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    const uint32_t u0 = first_bucket << P.b2, nu = n_buckets << P.b2;
    std::vector<uint32_t> work[3];
    std::vector<std::pair<uint64_t, uint32_t>> large, big;
    uint64_t tot_kmers = 0, retry_nmax = 0;
    // Opt-in (GGCAT_B200_WIDE_BY_KEYS=1): measured on the C5 slice (14 k-record units holding ~800 keys) the direct shared-table
    // path takes 83.8 ms against 33.5 + 29.6 ms for key partitions + partition tables -- one super-k-mer per thread leaves the
    // inserts as unbalanced as the rolling hashes, the partition pass balances them.  Kept for units of few, long super-k-mers.
    const char *wbk = getenv("GGCAT_B200_WIDE_BY_KEYS");
    const bool wide_by_keys = wbk && atoi(wbk) != 0;
    const bool by_keys = wide_by_keys && c->wide_mode != MODE_COLOR && !c->no_tiers && c->distinct_ratio < 1.0;
    const double keys_per_rec = std::min(1.0, 1.15 * c->distinct_ratio);
    for (uint32_t u = u0; u < u0 + nu; u++) {
        uint64_t n = 0;
        for (Chunk *ch : c->chunks)
            if (u >= ch->first_unit && u < ch->first_unit + ch->n_units) n += ch->h_kmers[u - ch->first_unit];
        if (n == 0) continue;
        if (n >= (1ull << 30)) return set_err(GGCAT_B200_ERR_INVALID, "unit %u holds %llu k-mers (> 2^30)", u, (unsigned long long)n);
        tot_kmers += n;
        // shared tables are sized by the distinct keys a unit should hold (distinct / records of what this context merged
        // before, 15 % margin, load <= 1/2); inserts probe a bounded number of slots and a unit whose table fills up comes
        // back through the retry list.  Coloured builds key by (k-mer, colour): every record may be a new key.
        const double need_keys = (double)n * keys_per_rec;
        if (n <= W_TS_S * 3 / 4 || (by_keys && need_keys <= W_TS_S / 2)) { work[0].push_back(u); retry_nmax = std::max(retry_nmax, n); }
        else if (n <= W_TS_L * 3 / 4 || (by_keys && need_keys <= W_TS_L / 2)) { work[1].push_back(u); retry_nmax = std::max(retry_nmax, n); }
        else if (n <= (uint64_t)PART_MAXP * W_PART_TARGET && !c->no_partition) { big.push_back({n, u}); retry_nmax = std::max(retry_nmax, n); }
        else large.push_back({n, u});
    }
    std::sort(large.begin(), large.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
    for (auto &pr : large) work[2].push_back(pr.second);
    std::sort(big.begin(), big.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
    BigPlan128 bp;
    uint64_t big_records = 0;
    for (auto &pr : big) {
        uint32_t lp = 2;
        while (((uint64_t)W_PART_TARGET << lp) < pr.first) lp++;
        bp.unit.push_back(pr.second); bp.logp.push_back(lp); bp.pbase.push_back(bp.n_parts);
        for (uint32_t q = 0; q < (1u << lp); q++) bp.part_big.push_back((uint32_t)bp.unit.size() - 1);
        bp.n_parts += 1u << lp;
        big_records += pr.first;
        bp.nmax = std::max(bp.nmax, pr.first);
    }
    std::vector<ChunkView> views;
    for (Chunk *ch : c->chunks) {
        ChunkView v;
        v.desc = ch->d_desc; v.payload = ch->d_payload; v.unit_off = ch->d_unit_off; v.unit_woff = ch->d_unit_woff; v.unit_kmers = ch->d_unit_kmers;
        v.first_unit = ch->first_unit; v.n_units = ch->n_units; v.word_bias = ch->word_bias; v.pad = 0;
        views.push_back(v);
    }
    CU(c->d_views.reserve(std::max<size_t>(1, views.size()) * sizeof(ChunkView)));
    if (!views.empty())
        CU(cudaMemcpyAsync(c->d_views.p, views.data(), views.size() * sizeof(ChunkView), cudaMemcpyHostToDevice, st));
    for (int q = 0; q < 3; q++) {
        CU(c->d_work[q].reserve(std::max<size_t>(1, work[q].size()) * 4));
        if (!work[q].empty())
            CU(cudaMemcpyAsync(c->d_work[q].p, work[q].data(), work[q].size() * 4, cudaMemcpyHostToDevice, st));
    }
    const uint64_t cap = std::max<uint64_t>(tot_kmers, 1);
    // dynamic region [0, cap) (cursor allocation), then the static regions of the partitioned units
    {
        uint64_t acc = cap;
        for (auto &pr : big) { bp.off.push_back(acc); acc += pr.first; }
    }
    const uint64_t cap_all = cap + big_records;
    CU(c->out_keys.reserve(cap_all * 8)); CU(c->out_hi.reserve(cap_all * 8)); CU(c->out_cf.reserve(cap_all * 4));
    const bool with_src = c->wide_mode == MODE_RK128;
    if (with_src) { CU(c->out_src.reserve(cap_all * 8 * c->src_words)); CU(c->sort_idx.reserve(cap_all * 8)); }
    const uint64_t rec_total = std::max(cap, pb.cap_total);
    if (pb.ub == 0) {
        // coloured builds fold in one piece and need every (k-mer, colour) entry; the others grow with the survivors
        TRY(final_reserve(c, c->wide_mode == MODE_COLOR ? rec_total : final_estimate(c, rec_total), 0, true));
        CU(c->unit_final_off.reserve(((size_t)c->P.n_units + 2) * 8));
    }
    CU(c->cursor.reserve(64)); CU(c->overflow.reserve(16));
    CU(c->unit_out_off.reserve(((size_t)nu + 1) * 8)); CU(c->unit_out_cnt.reserve(((size_t)nu + 1) * 4));
    CU(cudaMemsetAsync(c->cursor.p, 0, 64, st));
    CU(cudaMemsetAsync(c->overflow.p, 0, 16, st));
    CU(cudaMemsetAsync(c->unit_out_off.p, 0, ((size_t)nu + 1) * 8, st));
    CU(cudaMemsetAsync(c->unit_out_cnt.p, 0, ((size_t)nu + 1) * 4, st));
    MergeOut128 out;
    out.keys_lo = c->out_keys.as<uint64_t>(); out.keys_hi = c->out_hi.as<uint64_t>(); out.count_flags = c->out_cf.as<uint32_t>();
    out.cursor = c->cursor.as<unsigned long long>(); out.unit_out_off = c->unit_out_off.as<uint64_t>();
    out.unit_out_cnt = c->unit_out_cnt.as<uint32_t>(); out.capacity = cap; out.overflow = c->overflow.as<uint32_t>();
    out.src = with_src ? c->out_src.as<uint64_t>() : nullptr; out.src_words = c->src_words; out.pad = 0;
    const ChunkView *dv = c->d_views.as<ChunkView>();
    const uint32_t nch = (uint32_t)views.size();
    // retry list shared by the shared-table kernels (table full) and the partition kernel (partition overflow)
    CU(c->d_retry.reserve(((size_t)nu + 2) * 4));
    uint32_t *retry_base = c->d_retry.as<uint32_t>();
    CU(cudaMemsetAsync(retry_base, 0, 4, st));
    const uint64_t n_cand = work[0].size() + work[1].size() + big.size();
    if (c->wide_mode == MODE_SEQ128) {
        TRY(launch_hash128<MODE_SEQ128>(c, dv, nch, work, u0, out, large, retry_base)); TRY(launch_partitions128<MODE_SEQ128>(c, dv, nch, u0, out, bp, retry_base));
        TRY(launch_retry128<MODE_SEQ128>(c, dv, nch, u0, out, retry_base, n_cand, retry_nmax));
    } else if (c->wide_mode == MODE_RK128) {
        TRY(launch_hash128<MODE_RK128>(c, dv, nch, work, u0, out, large, retry_base)); TRY(launch_partitions128<MODE_RK128>(c, dv, nch, u0, out, bp, retry_base));
        TRY(launch_retry128<MODE_RK128>(c, dv, nch, u0, out, retry_base, n_cand, retry_nmax));
    } else {
        TRY(launch_hash128<MODE_COLOR>(c, dv, nch, work, u0, out, large, retry_base)); TRY(launch_partitions128<MODE_COLOR>(c, dv, nch, u0, out, bp, retry_base));
        TRY(launch_retry128<MODE_COLOR>(c, dv, nch, u0, out, retry_base, n_cand, retry_nmax));
    }
    // ---- order every unit's entries by key into the unit-ordered layout
    uint32_t end_bit = 128;
    if (c->wide_mode == MODE_SEQ128) end_bit = std::min(128u, (2 * P.k + 7) & ~7u);
    else if (c->wide_mode == MODE_COLOR) end_bit = std::min(128u, (32 + 2 * P.k + 7) & ~7u);
    for (int attempt = 0;; attempt++) {
        {
            LaunchTimer t(c, F_SORT128, 2);
            k_scan_counts_u64<<<1, 1024, 0, st>>>(c->unit_out_cnt.as<uint32_t>(), c->unit_final_off.as<uint64_t>() + pb.ub, nu, pb.eb);
            auto kern = with_src ? k_sort_units128<W_SORT_THREADS, W_SORT_CAP, true> : k_sort_units128<W_SORT_THREADS, W_SORT_CAP, false>;
            const size_t smem = sort_units128_smem_bytes<W_SORT_THREADS, W_SORT_CAP>();
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)std::min<uint32_t>(nu, (uint32_t)c->sm_count * 2 * 8), W_SORT_THREADS, smem, st>>>(
                c->out_keys.as<uint64_t>(), c->out_hi.as<uint64_t>(), c->out_cf.as<uint32_t>(), c->unit_out_off.as<uint64_t>(),
                c->unit_out_cnt.as<uint32_t>(), c->unit_final_off.as<uint64_t>() + pb.ub, c->out_keys2.as<uint64_t>(),
                c->out_hi2.as<uint64_t>(), c->out_cf2.as<uint32_t>(), nu, 0u, end_bit, c->fin_cap, c->overflow.as<uint32_t>(),
                with_src ? c->out_src.as<uint64_t>() : nullptr, with_src ? c->out_src2.as<uint64_t>() : nullptr,
                with_src ? c->sort_idx.as<uint32_t>() : nullptr, with_src ? c->sort_idx.as<uint32_t>() + cap_all : nullptr, c->src_words);
        }
        CU(cudaGetLastError());
        if (c->wide_mode == MODE_COLOR || attempt > 0) break;   // sized for every record / already regrown
        CU(cudaMemcpyAsync(c->h_pinned, c->cursor.p, 32, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(c->h_pinned + 8, c->overflow.p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if ((uint32_t)c->h_pinned[8] != 4u) break;
        const uint64_t need = pb.eb + c->h_pinned[0] + c->h_pinned[3];
        TRY(final_reserve(c, std::max(need, std::min(rec_total, need + need / 2)), pb.eb, true));
        CU(cudaMemsetAsync(c->overflow.p, 0, 16, st));
    }
    c->fin = FinalTable();
    if (c->wide_mode == MODE_COLOR) {
        // fold the (k-mer, colour) entries of every k-mer; the unsorted buffers are free again and take the result
        const size_t ub = ((size_t)nu + 2);
        CU(c->unit_keys.reserve(ub * 4)); CU(c->unit_cols.reserve(ub * 4)); CU(c->col_off.reserve(ub * 8));
        CU(c->out_coloff.reserve((cap + 1) * 8)); CU(c->out_colors.reserve(cap * 4));
        LaunchTimer t(c, F_COLOR_FOLD, 4);
        const unsigned grid = (unsigned)std::min<uint32_t>(nu, (uint32_t)c->sm_count * 8);
        k_color_fold<256, false><<<grid, 256, 0, st>>>(
            c->out_keys2.as<uint64_t>(), c->out_hi2.as<uint64_t>(), c->out_cf2.as<uint32_t>(), c->unit_final_off.as<uint64_t>(), nu,
            c->params.min_multiplicity, c->unit_keys.as<uint32_t>(), c->unit_cols.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr,
            nullptr, nullptr, nullptr);
        // unit_out_off is free now: it receives the per-unit key offsets of the folded table
        k_scan_counts_u64<<<1, 1024, 0, st>>>(c->unit_keys.as<uint32_t>(), c->unit_out_off.as<uint64_t>(), nu);
        k_scan_counts_u64<<<1, 1024, 0, st>>>(c->unit_cols.as<uint32_t>(), c->col_off.as<uint64_t>(), nu);
        k_color_fold<256, true><<<grid, 256, 0, st>>>(
            c->out_keys2.as<uint64_t>(), c->out_hi2.as<uint64_t>(), c->out_cf2.as<uint32_t>(), c->unit_final_off.as<uint64_t>(), nu,
            c->params.min_multiplicity, nullptr, nullptr, c->unit_out_off.as<uint64_t>(), c->col_off.as<uint64_t>(),
            c->out_keys.as<uint64_t>(), c->out_hi.as<uint64_t>(), c->out_cf.as<uint32_t>(), c->out_coloff.as<uint64_t>(),
            c->out_colors.as<uint32_t>());
        CU(cudaMemcpyAsync(c->h_pinned + 4, c->unit_out_off.as<uint64_t>() + nu, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(c->h_pinned + 5, c->col_off.as<uint64_t>() + nu, 8, cudaMemcpyDeviceToHost, st));
        c->fin.keys_lo = c->out_keys.as<uint64_t>(); c->fin.keys_hi = c->out_hi.as<uint64_t>(); c->fin.cf = c->out_cf.as<uint32_t>();
        c->fin.unit_off = c->unit_out_off.as<uint64_t>(); c->fin.color_off = c->out_coloff.as<uint64_t>();
        c->fin.colors = c->out_colors.as<uint32_t>();
    } else {
        c->fin.keys_lo = c->out_keys2.as<uint64_t>(); c->fin.keys_hi = c->out_hi2.as<uint64_t>(); c->fin.cf = c->out_cf2.as<uint32_t>();
        c->fin.unit_off = c->unit_final_off.as<uint64_t>();
        c->fin.src = with_src ? c->out_src2.as<uint64_t>() : nullptr;
    }
    CU(cudaMemcpyAsync(c->h_pinned, c->cursor.p, 32, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_pinned + 8, c->overflow.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    const uint32_t ovf = (uint32_t)c->h_pinned[8];
    if (ovf) return set_err(GGCAT_B200_ERR_CAPACITY, "merge output overflow (code %u)", ovf);
    c->h_pinned[0] += c->h_pinned[3];   // entries of the dynamic region + entries of the partitioned units' regions
    uint64_t uq = c->h_pinned[1];
    if (c->wide_mode != MODE_COLOR && c->h_pinned[2] >= (1ull << 20))
        c->distinct_ratio = (double)c->h_pinned[1] / (double)c->h_pinned[2];   // sizes the shared tables of the next part / build
    if (c->wide_mode == MODE_COLOR) {
        c->fin.n_entries = c->h_pinned[4]; c->fin.n_colors = c->h_pinned[5];
        uq = 0;  // distinct (k-mer, colour) pairs are not the reference's distinct k-mers; not tracked for coloured builds
    } else c->fin.n_entries = pb.eb + c->h_pinned[0];
    c->last_entries = c->fin.n_entries;
    if (n_entries) *n_entries = c->wide_mode == MODE_COLOR ? c->fin.n_entries : c->h_pinned[0];
    if (unique) *unique = uq;
    if (total) *total = c->h_pinned[2];
    return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

const char *ggcat_b200_last_error(void) { return g_err.c_str(); }
uint32_t ggcat_b200_abi_version(void) { return GGCAT_B200_ABI_VERSION; }
uint32_t ggcat_b200_compute_best_m(uint32_t k) { return best_m(k); }

void ggcat_b200_bucket_counts(uint64_t bases_count, uint32_t *buckets_log, uint32_t *second_log) {
    // crates/io/src/lib.rs:75-89,113-128 with the thresholds of crates/config/src/lib.rs:62-90
    auto next_pow2 = [](uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; };
    auto ilog2 = [](uint64_t x) { uint32_t l = 0; while (x >>= 1) l++; return l; };
    uint64_t buckets = std::max<uint64_t>(std::min<uint64_t>(1ull << 10, bases_count / (512 * 1024)), bases_count / (1024ull * 1024 * 1024));
    buckets = std::max<uint64_t>(std::min<uint64_t>(next_pow2(buckets), 1ull << 13), 1ull << 2);
    const uint64_t per = bases_count / buckets;
    uint64_t second = std::max<uint64_t>(std::min<uint64_t>(1ull << 6, per / (2 * 1024)), per / (4ull * 1024 * 1024));
    second = std::max<uint64_t>(std::min<uint64_t>(next_pow2(second), 1ull << 8), 1ull << 1);
    if (buckets_log) *buckets_log = ilog2(buckets);
    if (second_log) *second_log = ilog2(second);
}

void *ggcat_b200_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { set_err(GGCAT_B200_ERR_CUDA, "cudaMallocHost(%llu) failed", (unsigned long long)bytes); return nullptr; }
    return p;
}
void ggcat_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

int32_t ggcat_b200_create(const ggcat_b200_params *params, ggcat_b200_ctx **out) {
    if (!params || !out) return set_err(GGCAT_B200_ERR_INVALID, "null argument");
    ggcat_b200_params p = *params;
    if (p.m == 0) p.m = best_m(p.k);
    if (p.hash_type == GGCAT_B200_HASH_AUTO) p.hash_type = p.k <= 64 ? GGCAT_B200_HASH_SEQ : GGCAT_B200_HASH_RK128;  // crates/api/src/utils.rs:17-26
    if (p.hash_type != GGCAT_B200_HASH_SEQ && p.hash_type != GGCAT_B200_HASH_RK128)
        return set_err(GGCAT_B200_ERR_INVALID, "hash_type=%u unsupported (1 = seq-hash, 4 = rabin-karp128)", p.hash_type);
    // seq-hash keys hold at most 64 bases; longer k-mers take the non-invertible rabin-karp128 hash, as in the reference
    // (crates/api/src/utils.rs:17-26), up to RK_MAXK bases
    if (p.k < 4 || p.k > (uint32_t)RK_MAXK) return set_err(GGCAT_B200_ERR_INVALID, "k=%u unsupported (4..%d)", p.k, RK_MAXK);
    if (p.k > 64 && p.hash_type != GGCAT_B200_HASH_RK128) return set_err(GGCAT_B200_ERR_INVALID, "k=%u needs hash_type rabin-karp128 (seq-hash keys hold <= 64 bases)", p.k);
    if (p.m < 2 || p.m > 32 || p.m >= p.k) return set_err(GGCAT_B200_ERR_INVALID, "m=%u invalid for k=%u", p.m, p.k);
    if (p.k - p.m < 2 || p.k - p.m > (uint32_t)WIN_WMAX) return set_err(GGCAT_B200_ERR_INVALID, "k-m=%u outside 2..%d", p.k - p.m, WIN_WMAX);
    if (p.buckets_count_log > 13) return set_err(GGCAT_B200_ERR_INVALID, "buckets_count_log > 13");  // config MAX_BUCKETS_COUNT_LOG
    if (p.second_buckets_count_log > 8) return set_err(GGCAT_B200_ERR_INVALID, "second_buckets_count_log > 8");
    if (p.min_multiplicity == 0) p.min_multiplicity = 1;
    if (p.colors && (p.hash_type != GGCAT_B200_HASH_SEQ || p.k > 48))
        return set_err(GGCAT_B200_ERR_INVALID, "colors need seq-hash and k <= 48 (k-mer and colour id share one 128-bit slot key)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(GGCAT_B200_ERR_CUDA, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (p.device < 0 || p.device >= ndev) return set_err(GGCAT_B200_ERR_INVALID, "device %d out of range (0..%d)", p.device, ndev - 1);
    CU(cudaSetDevice(p.device));
    ggcat_b200_ctx *c = new ggcat_b200_ctx();
    c->params = p;
    c->device = p.device;
    c->P.k = p.k; c->P.m = p.m; c->P.w = p.k - p.m; c->P.b1 = p.buckets_count_log; c->P.b2 = p.second_buckets_count_log;
    c->P.forward_only = p.forward_only ? 1 : 0; c->P.colors = p.colors ? 1 : 0;
    c->P.n_units = ((1u << c->P.b1) + 1u) << c->P.b2;
    memset(&c->stats, 0, sizeof(c->stats));
    memset(c->fam_ms, 0, sizeof(c->fam_ms));
    memset(c->fam_launches, 0, sizeof(c->fam_launches));
    c->P.hash_type = p.hash_type;
    if (p.colors) c->wide_mode = MODE_COLOR;
    else if (p.hash_type == GGCAT_B200_HASH_RK128) c->wide_mode = MODE_RK128;
    else if (p.k > 31) c->wide_mode = MODE_SEQ128;
    memset(&c->rk, 0, sizeof(c->rk));
    if (c->wide_mode == MODE_RK128) {
        // crates/hashes/src/cn_rkhash.rs:69-75 (u128 module); M^(k-1) as crates/hashes/src/lib.rs:169-191 (fastexp)
        const u128 M = ((u128)0x3eb9402f3e733993ULL << 64) | 0xadd64d3ca00e1b6bULL;
        const u128 MI = ((u128)0x09cb6ff6f1b1a6d7ULL << 64) | 0x33e0952e899c3943ULL;
        const u128 LA = ((u128)0x4751137d01d863c5ULL << 64) | 0xb8c36de2b7d399dfULL;
        const u128 LC = ((u128)0x37ea3a13226503fbULL << 64) | 0x783f5cb69f4552bdULL;
        const u128 LG = ((u128)0x50796b285343f09aULL << 64) | 0x0c53113ae736572bULL;
        const u128 LT = ((u128)0x1e62d96a5e1f5adeULL << 64) | 0x2d4e68d8f88110b7ULL;
        const u128 L[4] = {LA, LC, LT, LG};  // 2-bit codes A0 C1 T2 G3 (cn_rkhash_base.rs:10-44)
        u128 mk1 = 1, sq = M;
        for (uint32_t e = p.k - 1; e > 0; e >>= 1) { if (e & 1) mk1 *= sq; sq *= sq; }
        c->rk.mult = to_k128(M); c->rk.mult_inv = to_k128(MI);
        for (int b = 0; b < 4; b++) {
            c->rk.fwd[b] = to_k128(L[b]); c->rk.bkw[b] = to_k128(L[b ^ 2]);
            c->rk.fwd_mk[b] = to_k128(L[b] * mk1 * M); c->rk.bkw_mk1[b] = to_k128(L[b ^ 2] * mk1);
        }
        // per-position terms L[b] M^j, j < RK_MAXK
        std::vector<K128> pos(2 * RK_MAXK * 4);
        u128 mj = 1;
        for (int j = 0; j < RK_MAXK; j++) {
            for (int b = 0; b < 4; b++) { pos[j * 4 + b] = to_k128(L[b] * mj); pos[RK_MAXK * 4 + j * 4 + b] = to_k128(L[b ^ 2] * mj); }
            mj *= M;
        }
        c->src_words = std::max<uint32_t>(2, (p.k + 31) / 32);
        if (c->d_rkpos.reserve(pos.size() * sizeof(K128)) != cudaSuccess ||
            cudaMemcpy(c->d_rkpos.p, pos.data(), pos.size() * sizeof(K128), cudaMemcpyHostToDevice) != cudaSuccess) {
            delete c;
            return set_err(GGCAT_B200_ERR_CUDA, "rabin-karp table upload failed");
        }
        c->rk.pos_fwd = c->d_rkpos.as<K128>(); c->rk.pos_bkw = c->d_rkpos.as<K128>() + RK_MAXK * 4;
    }
    if (const char *np = getenv("GGCAT_B200_NO_PARTITION")) c->no_partition = atoi(np) != 0;
    if (const char *nt = getenv("GGCAT_B200_NO_TIERS")) c->no_tiers = atoi(nt) != 0;
    if (const char *ta = getenv("GGCAT_B200_TIER_A")) c->tier_a_variant = std::max(0, std::min(2, atoi(ta)));
    if (const char *pt = getenv("GGCAT_B200_PART_TARGET")) c->part_fixed = strcmp(pt, "fixed") == 0;
    if (const char *dr = getenv("GGCAT_B200_DISTINCT_RATIO")) { const double v = atof(dr); if (v > 0) c->distinct_ratio = v; }  // tests: pretend a ratio
    if (const char *mm = getenv("GGCAT_B200_MERGE")) c->merge_mode = (strcmp(mm, "sort") == 0) ? 0 : 1;
    if (const char *mb = getenv("GGCAT_B200_MAX_BATCH")) { uint64_t v = strtoull(mb, nullptr, 10); if (v >= 1024) c->max_batch = std::min<uint64_t>(v, 1ull << 30); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p.device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (const char *hb = getenv("GGCAT_B200_HOST_BATCH")) { uint64_t v = strtoull(hb, nullptr, 10); if (v >= 1024) c->host_batch = std::min<uint64_t>(v, c->max_batch); }
    if (const char *ps = getenv("GGCAT_B200_PART_SCHEME")) c->part_scheme = atoi(ps);
    if (const char *pk = getenv("GGCAT_B200_PART_KMERS")) { uint64_t v = strtoull(pk, nullptr, 10); if (v >= 1024) c->part_kmers = v; }
    if (const char *pk = getenv("GGCAT_B200_PART_KMERS_DEV")) { uint64_t v = strtoull(pk, nullptr, 10); if (v >= 1024) { c->part_kmers_dev = v; c->part_dev_fixed = true; } }
    c->host_batch = std::min(c->host_batch, c->max_batch);
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_part, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < ggcat_b200_ctx::N_STAGE && ok; i++)
        ok = cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok || cudaMallocHost((void **)&c->h_pinned, 16 * 8) != cudaSuccess) {
        delete c;
        return set_err(GGCAT_B200_ERR_CUDA, "stream / pinned allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = c;
    return 0;
}

int32_t ggcat_b200_reset(ggcat_b200_ctx *c) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    cudaStreamSynchronize(c->stream);
    abort_open_chunk(c);
    for (Chunk *ch : c->chunks) {
        if (!ch->imported) c->chunk_pool.push_back(ch);  // keep the device buffers for the next build
        else { ch->release(); delete ch; }
    }
    c->chunks.clear();
    c->ut_n = c->ut_words = 0;
    c->fin = FinalTable();
    c->finished = false;
    c->peer.build_started = false;
    memset(&c->stats, 0, sizeof(c->stats));
    return 0;
}

void ggcat_b200_destroy(ggcat_b200_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    ggcat_b200_reset(c);
    for (Chunk *ch : c->chunk_pool) { ch->release(); delete ch; }
    c->chunk_pool.clear();
    collect_timings(c);
    for (DevBuf *b : {&c->d_ascii, &c->d_offsets, &c->d_colors, &c->pk, &c->bad, &c->brk, &c->ent, &c->tile_cnt, &c->tile_sbase, &c->tile_group,
                      &c->tmp, &c->tmp_color, &c->cur_cnt, &c->totals, &c->d_views, &c->d_work[0], &c->d_work[1],
                      &c->d_work[2], &c->d_scratch, &c->out_keys, &c->out_cf, &c->out_keys2, &c->out_cf2,
                      &c->d_rkpos, &c->d_recfl, &c->d_mstage, &c->d_static_off, &c->cursor, &c->unit_out_off, &c->unit_out_cnt, &c->unit_final_off, &c->overflow, &c->d_retry,
                      &c->d_partmeta, &c->d_recs, &c->fin_tmp_keys, &c->fin_tmp_cf, &c->tok_agg, &c->tok_keep, &c->tok_base, &c->tok_marks,
                      &c->tok_tmarks, &c->tok_seq, &c->tok_offsets, &c->tok_text, &c->tok_colors, &c->ut_links, &c->ut_visited, &c->ut_recs,
                      &c->ut_bases, &c->ut_counters, &c->mu_recs, &c->mu_bases, &c->mu_ht, &c->mu_first, &c->mu_partner, &c->mu_visited,
                      &c->out_hi, &c->out_hi2, &c->d_recsrc, &c->out_src, &c->out_src2, &c->sort_idx, &c->unit_keys, &c->unit_cols, &c->col_off, &c->out_coloff, &c->out_colors})
        b->release();
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    for (HostTable *t : c->free_tables) { t->release(); delete t; }
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (int i = 0; i < ggcat_b200_ctx::N_STAGE; i++) {
        c->st_ascii[i].release(); c->st_off[i].release(); c->st_col[i].release();
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    if (c->ev_part) cudaEventDestroy(c->ev_part);
    {
        PeerState &ps = c->peer;
        for (uint32_t r = 0; r < ps.world; r++)
            if (ps.connected && r != ps.rank && ps.peer_arena[r]) cudaIpcCloseMemHandle(ps.peer_arena[r]);
        if (ps.arena) cudaFree(ps.arena);
        ps.d_stage.release(); ps.d_err.release();
        if (ps.meta_stream) cudaStreamDestroy(ps.meta_stream);
        if (ps.data_stream) cudaStreamDestroy(ps.data_stream);
        for (cudaEvent_t e : {ps.ev_scatter, ps.ev_data, ps.ev_meta}) if (e) cudaEventDestroy(e);
        if (ps.h_stage) cudaFreeHost(ps.h_stage);
        if (ps.h_recv) cudaFreeHost(ps.h_recv);
    }
    if (c->h_mstage) cudaFreeHost(c->h_mstage);
    if (c->h_unitigs) cudaFreeHost(c->h_unitigs);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int32_t push_reads_host(ggcat_b200_ctx *c, const uint8_t *data, const uint64_t *offsets, uint64_t n_reads,
                               const uint32_t *colors, bool packed);
int32_t ggcat_b200_push_reads(ggcat_b200_ctx *c, const uint8_t *data, const uint64_t *offsets, uint64_t n_reads,
                              const uint32_t *colors) {
    return push_reads_host(c, data, offsets, n_reads, colors, false);
}
int32_t ggcat_b200_push_reads_packed(ggcat_b200_ctx *c, const uint8_t *packed, const uint64_t *offsets, uint64_t n_reads,
                                     const uint32_t *colors) {
    return push_reads_host(c, packed, offsets, n_reads, colors, true);
}
// packed: `data` is one contiguous 2-bit stream and offsets count BASES; a batch of records [r0, r1) travels as the bytes
// [offsets[r0] / 4, ceil(offsets[r1] / 4)) and starts offsets[r0] % 4 bases into its first byte.
static int32_t push_reads_host(ggcat_b200_ctx *c, const uint8_t *data, const uint64_t *offsets, uint64_t n_reads,
                               const uint32_t *colors, bool packed) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (c->finished) return set_err(GGCAT_B200_ERR_STATE, "push_reads after finish_bucketing");
    if (n_reads == 0) return 0;
    if (!data || !offsets) return set_err(GGCAT_B200_ERR_INVALID, "null input");
    if (colors && !c->P.colors) colors = nullptr;
    // greedy batches of whole records, <= host_batch bases (a longer single record gets its own batch, <= max_batch)
    std::vector<std::pair<uint64_t, uint64_t>> batches;
    for (uint64_t r0 = 0; r0 < n_reads;) {
        if (offsets[r0 + 1] - offsets[r0] > c->max_batch)
            return set_err(GGCAT_B200_ERR_INVALID, "record %llu is longer than the batch limit %llu; split it with k-1 overlap "
                           "(crates/io/src/sequences_reader.rs:162-173)", (unsigned long long)r0, (unsigned long long)c->max_batch);
        // the first batch is a third of the others (its copy is the only one no kernel overlaps) and the last ones
        // shrink (the kernels of the final batch are the only ones no copy overlaps)
        const uint64_t left = offsets[n_reads] - offsets[r0];
        uint64_t limit = c->host_batch;
        if (batches.empty()) limit = std::max<uint64_t>(c->host_batch / 3, 1024);
        else if (left <= c->host_batch / 2) limit = left;
        else if (left <= c->host_batch + c->host_batch / 2) limit = left * 2 / 3;
        uint64_t lo = r0 + 1, hi = n_reads;
        while (lo < hi) {  // last record that still fits
            const uint64_t mid = (lo + hi + 1) >> 1;
            if (offsets[mid] - offsets[r0] <= limit) lo = mid; else hi = mid - 1;
        }
        batches.push_back({r0, lo});
        r0 = lo;
    }
    // ring of staging slots: the copy stream fills the slots of batches i+1, i+2 while the compute stream works on batch i
    constexpr int NS = ggcat_b200_ctx::N_STAGE;
    auto issue_copy = [&](size_t bi) -> int32_t {
        const int sl = (int)(bi % NS);
        const uint64_t r0 = batches[bi].first, r1 = batches[bi].second;
        const uint64_t nb = offsets[r1] - offsets[r0], nr = r1 - r0;
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_free[sl], 0));
        CU(c->st_ascii[sl].reserve(nb + 64));
        CU(c->st_off[sl].reserve((nr + 1) * 8));
        if (packed) {
            const uint64_t b0 = offsets[r0] >> 2, b1 = (offsets[r1] + 3) >> 2;
            CU(cudaMemcpyAsync(c->st_ascii[sl].p, data + b0, b1 - b0, cudaMemcpyHostToDevice, c->copy_stream));
        } else
        CU(cudaMemcpyAsync(c->st_ascii[sl].p, data + offsets[r0], nb, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaMemcpyAsync(c->st_off[sl].p, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, c->copy_stream));
        if (colors) {
            CU(c->st_col[sl].reserve(nr * 4));
            CU(cudaMemcpyAsync(c->st_col[sl].p, colors + r0, nr * 4, cudaMemcpyHostToDevice, c->copy_stream));
        }
        CU(cudaEventRecord(c->ev_h2d[sl], c->copy_stream));
        return 0;
    };
    const bool trace = getenv("GGCAT_B200_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;   // trace: [origin, per batch: copy start, copy end, compute start, compute end]
    auto tmark = [&](cudaStream_t s) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); tev.push_back(e); } };
    tmark(c->copy_stream);
    auto issue_copy_t = [&](size_t bi) -> int32_t { tmark(c->copy_stream); TRY(issue_copy(bi)); tmark(c->copy_stream); return 0; };
    TRY(issue_copy_t(0));
    if (batches.size() > 1) TRY(issue_copy_t(1));
    for (size_t bi = 0; bi < batches.size(); bi++) {
        const int sl = (int)(bi % NS);
        if (bi + 2 < batches.size()) TRY(issue_copy_t(bi + 2));
        const uint64_t r0 = batches[bi].first, r1 = batches[bi].second;
        CU(cudaStreamWaitEvent(c->stream, c->ev_h2d[sl], 0));
        tmark(c->stream);
        int32_t rc = bucket_batch_device(c, c->st_ascii[sl].as<uint8_t>(), c->st_off[sl].as<uint64_t>(), r1 - r0, offsets[r0],
                                         offsets[r1] - offsets[r0], colors ? c->st_col[sl].as<uint32_t>() : nullptr,
                                         offsets[n_reads] - offsets[0], packed ? (int)(offsets[r0] & 3) : -1);
        if (rc) { abort_open_chunk(c); return rc; }
        CU(cudaEventRecord(c->ev_free[sl], c->stream));
        tmark(c->stream);
    }
    TRY(flush_open_chunk(c));    // one bucket chunk per push call
    if (trace) {
        cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_stream);
        fprintf(stderr, "[ggcat_b200 trace] push_reads: %zu batches; ms since first copy was queued:", batches.size());
        for (size_t i = 1; i < tev.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[i]); fprintf(stderr, " %.2f", ms); }
        fprintf(stderr, "\n  order: c0s c0e c1s c1e, then per batch i: [c(i+2)s c(i+2)e] k(i)s k(i)e\n");
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return 0;
}

int32_t ggcat_b200_push_reads_packed_device(ggcat_b200_ctx *c, const uint8_t *d_packed, const uint64_t *d_offsets, uint64_t n_reads,
                                            uint64_t n_bases, const uint32_t *d_colors) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (c->finished) return set_err(GGCAT_B200_ERR_STATE, "push_reads after finish_bucketing");
    if (n_reads == 0) return 0;
    if (!d_packed || !d_offsets) return set_err(GGCAT_B200_ERR_INVALID, "null input");
    if ((uintptr_t)d_packed & 3) return set_err(GGCAT_B200_ERR_INVALID, "packed device input must be 4-byte aligned");
    if (n_bases > c->max_batch) return set_err(GGCAT_B200_ERR_INVALID, "device batch of %llu bases exceeds limit %llu", (unsigned long long)n_bases, (unsigned long long)c->max_batch);
    int32_t rc = bucket_batch_device(c, d_packed, d_offsets, n_reads, 0, n_bases, c->P.colors ? d_colors : nullptr, n_bases, 0);
    if (rc) { abort_open_chunk(c); return rc; }
    return flush_open_chunk(c);
}

int32_t ggcat_b200_push_reads_device(ggcat_b200_ctx *c, const uint8_t *d_data, const uint64_t *d_offsets, uint64_t n_reads,
                                     uint64_t n_bytes, const uint32_t *d_colors) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (c->finished) return set_err(GGCAT_B200_ERR_STATE, "push_reads after finish_bucketing");
    if (n_reads == 0) return 0;
    if (!d_data || !d_offsets) return set_err(GGCAT_B200_ERR_INVALID, "null input");
    if (n_bytes > c->max_batch) return set_err(GGCAT_B200_ERR_INVALID, "device batch of %llu bases exceeds limit %llu", (unsigned long long)n_bytes, (unsigned long long)c->max_batch);
    int32_t rc = bucket_batch_device(c, d_data, d_offsets, n_reads, 0, n_bytes, c->P.colors ? d_colors : nullptr, n_bytes);
    if (rc) { abort_open_chunk(c); return rc; }
    return flush_open_chunk(c);
}


// ---- FASTA / FASTQ text on the device (tokenize.cuh) ----------------------------------------------
extern "C++" {
template <int FORMAT>
static int32_t tokenize_device_impl(ggcat_b200_ctx *c, const uint8_t *d_text, uint64_t n, uint64_t *n_records, uint64_t *n_seq) {
    cudaStream_t st = c->stream;
    const uint32_t n_tiles = (uint32_t)((n + TOK_TILE - 1) / TOK_TILE);
    CU(c->tok_agg.reserve(((size_t)n_tiles + 2) * 8)); CU(c->tok_keep.reserve(((size_t)n_tiles + 2) * 4)); CU(c->tok_base.reserve(((size_t)n_tiles + 2) * 4));
    CU(c->tok_marks.reserve((n / 32 + 4) * 4)); CU(c->tok_seq.reserve(n + 64)); CU(c->totals.reserve(8 * 8));
    CU(cudaMemsetAsync(c->tok_marks.p, 0, (n / 32 + 4) * 4, st));
    {
        LaunchTimer t(c, F_TOKENIZE, 5);
        k_tok_tile_lines<FORMAT><<<n_tiles, TOK_THREADS, 0, st>>>(d_text, n, c->tok_agg.as<unsigned long long>());
        k_tok_scan_tiles<FORMAT><<<1, 1024, 0, st>>>(c->tok_agg.as<unsigned long long>(), n_tiles);
        k_tok_compact<FORMAT, false><<<n_tiles, TOK_THREADS, 0, st>>>(d_text, n, c->tok_agg.as<unsigned long long>(), c->tok_keep.as<uint32_t>(), nullptr, nullptr, nullptr);
        k_exclusive_scan_u32<<<1, 1024, 0, st>>>(c->tok_keep.as<uint32_t>(), c->tok_base.as<uint32_t>(), n_tiles, c->totals.as<unsigned long long>() + 5);
        k_tok_compact<FORMAT, true><<<n_tiles, TOK_THREADS, 0, st>>>(d_text, n, c->tok_agg.as<unsigned long long>(), nullptr, c->tok_base.as<uint32_t>(),
                                                                    c->tok_seq.as<uint8_t>(), c->tok_marks.as<uint32_t>());
    }
    CU(cudaMemcpyAsync(c->h_pinned + 10, c->totals.as<unsigned long long>() + 5, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint64_t total = c->h_pinned[10];
    uint64_t nrec = 0;
    CU(c->tok_offsets.reserve(16));
    if (total) {
        const uint64_t n_words = total / 32 + 1;
        const uint32_t mt = (uint32_t)((n_words + 255) / 256);
        CU(c->tok_tmarks.reserve(((size_t)mt + 2) * 4 * 2));
        uint32_t *tm = c->tok_tmarks.as<uint32_t>(), *tr = tm + mt + 2;
        LaunchTimer t(c, F_TOKENIZE, 3);
        k_tok_count_marks<<<mt, 256, 0, st>>>(c->tok_marks.as<uint32_t>(), n_words, total, tm);
        k_exclusive_scan_u32<<<1, 1024, 0, st>>>(tm, tr, mt, c->totals.as<unsigned long long>() + 6);
        CU(cudaMemcpyAsync(c->h_pinned + 11, c->totals.as<unsigned long long>() + 6, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        nrec = c->h_pinned[11];
        CU(c->tok_offsets.reserve((nrec + 2) * 8));
        k_tok_write_offsets<<<mt, 256, 0, st>>>(c->tok_marks.as<uint32_t>(), n_words, tr, c->tok_offsets.as<uint64_t>());
    }
    c->h_pinned[12] = total;
    CU(cudaMemcpyAsync(c->tok_offsets.as<uint64_t>() + nrec, c->h_pinned + 12, 8, cudaMemcpyHostToDevice, st));
    CU(cudaGetLastError());
    *n_records = nrec; *n_seq = total;
    return 0;
}

static int32_t tokenize_device(ggcat_b200_ctx *c, const uint8_t *d_text, uint64_t n, int32_t format, uint64_t *n_records, uint64_t *n_seq) {
    if (format != TOK_FASTA && format != TOK_FASTQ) return set_err(GGCAT_B200_ERR_INVALID, "text format %d unsupported (0 = FASTA, 1 = FASTQ)", format);
    if (n == 0) { *n_records = *n_seq = 0; return 0; }
    if (n > c->max_batch) return set_err(GGCAT_B200_ERR_INVALID, "text block of %llu bytes exceeds the batch limit %llu", (unsigned long long)n, (unsigned long long)c->max_batch);
    return format == TOK_FASTA ? tokenize_device_impl<TOK_FASTA>(c, d_text, n, n_records, n_seq) : tokenize_device_impl<TOK_FASTQ>(c, d_text, n, n_records, n_seq);
}
__global__ void k_fill_u32(uint32_t *p, uint64_t n, uint32_t v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
}  // extern "C++"

int32_t ggcat_b200_tokenize_device(ggcat_b200_ctx *c, const uint8_t *d_text, uint64_t n_bytes, int32_t format, const uint8_t **d_seq,
                                   const uint64_t **d_offsets, uint64_t *n_records, uint64_t *n_seq_bytes) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!d_text && n_bytes) return set_err(GGCAT_B200_ERR_INVALID, "null input");
    uint64_t nr = 0, ns = 0;
    TRY(tokenize_device(c, d_text, n_bytes, format, &nr, &ns));
    CU(cudaStreamSynchronize(c->stream));
    if (d_seq) *d_seq = c->tok_seq.as<uint8_t>();
    if (d_offsets) *d_offsets = c->tok_offsets.as<uint64_t>();
    if (n_records) *n_records = nr;
    if (n_seq_bytes) *n_seq_bytes = ns;
    return 0;
}

static int32_t push_text_device_locked(ggcat_b200_ctx *c, const uint8_t *d_text, uint64_t n_bytes, int32_t format, uint32_t color, uint64_t *n_records) {
    if (c->finished) return set_err(GGCAT_B200_ERR_STATE, "push_text after finish_bucketing");
    if (!d_text && n_bytes) return set_err(GGCAT_B200_ERR_INVALID, "null input");
    uint64_t nr = 0, ns = 0;
    TRY(tokenize_device(c, d_text, n_bytes, format, &nr, &ns));
    if (n_records) *n_records = nr;
    if (nr == 0) return 0;
    const uint32_t *d_colors = nullptr;
    if (c->P.colors) {
        CU(c->tok_colors.reserve(nr * 4));
        k_fill_u32<<<(unsigned)((nr + 255) / 256), 256, 0, c->stream>>>(c->tok_colors.as<uint32_t>(), nr, color);
        d_colors = c->tok_colors.as<uint32_t>();
    }
    int32_t rc = bucket_batch_device(c, c->tok_seq.as<uint8_t>(), c->tok_offsets.as<uint64_t>(), nr, 0, ns, d_colors, ns);
    if (rc) { abort_open_chunk(c); return rc; }
    return flush_open_chunk(c);
}

int32_t ggcat_b200_push_text(ggcat_b200_ctx *c, const uint8_t *text, uint64_t n_bytes, int32_t format, uint32_t color, uint64_t *n_records) {
    TRY(check_ctx(c));
    if (!text && n_bytes) return set_err(GGCAT_B200_ERR_INVALID, "null input");
    if (n_bytes > c->max_batch) return set_err(GGCAT_B200_ERR_INVALID, "text block of %llu bytes exceeds the batch limit %llu: push whole records in smaller blocks",
                                               (unsigned long long)n_bytes, (unsigned long long)c->max_batch);
    std::lock_guard<std::mutex> lock__(c->mu);
    CU(c->tok_text.reserve(n_bytes + 64));
    CU(cudaMemcpyAsync(c->tok_text.p, text, n_bytes, cudaMemcpyHostToDevice, c->stream));
    return push_text_device_locked(c, c->tok_text.as<uint8_t>(), n_bytes, format, color, n_records);
}

int32_t ggcat_b200_push_text_device(ggcat_b200_ctx *c, const uint8_t *d_text, uint64_t n_bytes, int32_t format, uint32_t color, uint64_t *n_records) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    return push_text_device_locked(c, d_text, n_bytes, format, color, n_records);
}


// ---- the reference's bucket file format (wire.hpp) --------------------------------------------------
static int32_t dump_superkmers_impl(ggcat_b200_ctx *c, uint32_t bucket, ggcat_b200_superkmer *out, uint64_t cap, uint8_t *payload,
                                    uint64_t payload_cap, uint64_t *n_out, uint64_t *payload_bytes);
int32_t ggcat_b200_write_bucket_file(ggcat_b200_ctx *c, uint32_t bucket, const char *path, uint64_t *n_records) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!path) return set_err(GGCAT_B200_ERR_INVALID, "null path");
    if (c->P.colors) return set_err(GGCAT_B200_ERR_INVALID, "bucket files of coloured builds are not written (the colour varint depends on the reference writer's buffer state)");
    uint64_t n = 0, pb = 0;
    TRY(dump_superkmers_impl(c, bucket, nullptr, 0, nullptr, 0, &n, &pb));
    std::vector<ggcat_b200_superkmer> sk(n);
    std::vector<uint8_t> payload(pb + 16);
    if (n) TRY(dump_superkmers_impl(c, bucket, sk.data(), n, payload.data(), pb, &n, &pb));
    // group by sub-bucket, keeping the order inside each group
    const uint32_t nsub = 1u << c->P.b2;
    std::vector<uint64_t> cnt(nsub + 1, 0);
    for (auto &r : sk) cnt[r.second_bucket + 1]++;
    for (uint32_t q = 0; q < nsub; q++) cnt[q + 1] += cnt[q];
    std::vector<uint32_t> order(n);
    {
        std::vector<uint64_t> cur(cnt.begin(), cnt.end() - 1);
        for (uint64_t i = 0; i < n; i++) order[cur[sk[i].second_bucket]++] = (uint32_t)i;
    }
    ggb_wire::Writer w(c->P.k);
    for (uint32_t q = 0; q < nsub; q++) {
        if (cnt[q + 1] == cnt[q]) continue;
        w.begin_sub_bucket(q, cnt[q + 1] - cnt[q]);
        for (uint64_t i = cnt[q]; i < cnt[q + 1]; i++) {
            const ggcat_b200_superkmer &r = sk[order[i]];
            w.record(r.len, r.minimizer_pos, r.flags, payload.data() + r.payload_offset);
        }
    }
    w.finish();
    FILE *f = fopen(path, "wb");
    if (!f) return set_err(GGCAT_B200_ERR_INVALID, "cannot create %s", path);
    const bool ok = fwrite(w.out.data(), 1, w.out.size(), f) == w.out.size();
    if (fclose(f) != 0 || !ok) return set_err(GGCAT_B200_ERR_INVALID, "short write to %s", path);
    if (n_records) *n_records = n;
    return 0;
}

int32_t ggcat_b200_import_bucket_file(ggcat_b200_ctx *c, uint32_t bucket, const char *path, uint64_t *n_records) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    const DevParams &P = c->P;
    if (!path) return set_err(GGCAT_B200_ERR_INVALID, "null path");
    if (c->finished) return set_err(GGCAT_B200_ERR_STATE, "import_bucket_file after finish_bucketing");
    if (P.colors) return set_err(GGCAT_B200_ERR_INVALID, "bucket files of coloured builds are not read");
    if (bucket > (1u << P.b1)) return set_err(GGCAT_B200_ERR_INVALID, "bucket %u out of range", bucket);
    std::vector<uint8_t> file;
    {
        FILE *f = fopen(path, "rb");
        if (!f) return set_err(GGCAT_B200_ERR_INVALID, "cannot open %s", path);
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        file.resize(sz > 0 ? (size_t)sz : 0);
        const bool ok = file.empty() || fread(file.data(), 1, file.size(), f) == file.size();
        fclose(f);
        if (!ok) return set_err(GGCAT_B200_ERR_INVALID, "short read from %s", path);
    }
    std::vector<ggb_wire::Record> recs;
    const std::string err = ggb_wire::parse(file, P.k, recs);
    if (!err.empty()) return set_err(GGCAT_B200_ERR_INVALID, "%s: %s", path, err.c_str());
    if (n_records) *n_records = recs.size();
    if (recs.empty()) return 0;
    const uint32_t nsub = 1u << P.b2;
    std::vector<uint32_t> h_cnt(nsub + 1, 0), h_words(nsub + 1, 0), h_kmers(nsub + 1, 0), h_off(nsub + 1, 0), h_woff(nsub + 1, 0);
    for (auto &r : recs) {
        if (r.sub_bucket >= nsub) return set_err(GGCAT_B200_ERR_INVALID, "%s: sub-bucket %u >= %u", path, r.sub_bucket, nsub);
        if (r.len < P.k) return set_err(GGCAT_B200_ERR_INVALID, "%s: record shorter than k", path);
        h_cnt[r.sub_bucket]++; h_words[r.sub_bucket] += (r.len + 15) / 16; h_kmers[r.sub_bucket] += r.len - P.k + 1;
    }
    uint64_t a = 0, b = 0, km = 0;
    for (uint32_t q = 0; q < nsub; q++) { h_off[q] = (uint32_t)a; h_woff[q] = (uint32_t)b; a += h_cnt[q]; b += h_words[q]; km += h_kmers[q]; }
    h_off[nsub] = (uint32_t)a; h_woff[nsub] = (uint32_t)b;
    if (b >= (1ull << 32)) return set_err(GGCAT_B200_ERR_INVALID, "%s: payload exceeds 2^32 words", path);
    std::vector<uint4> desc(a);
    std::vector<uint32_t> payload(b + 8, 0);
    {
        std::vector<uint32_t> cd(h_off.begin(), h_off.end()), cw(h_woff.begin(), h_woff.end());
        for (auto &r : recs) {
            const uint32_t slot = cd[r.sub_bucket]++, woff = cw[r.sub_bucket];
            cw[r.sub_bucket] += (r.len + 15) / 16;
            desc[slot] = make_uint4(woff, r.len, make_meta(r.minimizer_pos, r.flags, 0u, r.sub_bucket), 0u);
            memcpy(reinterpret_cast<uint8_t *>(payload.data() + woff), file.data() + r.byte_off, (r.len + 3) / 4);
        }
    }
    Chunk *ch = new Chunk();
    ch->imported = true; ch->first_unit = bucket << P.b2; ch->n_units = nsub; ch->word_bias = 0;
    cudaError_t e = ch->desc.reserve(a * 16);
    if (e == cudaSuccess) e = ch->payload.reserve((b + 8) * 4);
    for (DevBuf *d : {&ch->unit_cnt, &ch->unit_off, &ch->unit_words, &ch->unit_woff, &ch->unit_kmers})
        if (e == cudaSuccess) e = d->reserve(((size_t)nsub + 2) * 4);
    auto up = [&](DevBuf &d, const void *src, size_t bytes) { if (e == cudaSuccess) e = cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, c->stream); };
    up(ch->desc, desc.data(), a * 16); up(ch->payload, payload.data(), (b + 8) * 4);
    up(ch->unit_cnt, h_cnt.data(), (nsub + 1) * 4); up(ch->unit_off, h_off.data(), (nsub + 1) * 4);
    up(ch->unit_words, h_words.data(), (nsub + 1) * 4); up(ch->unit_woff, h_woff.data(), (nsub + 1) * 4);
    up(ch->unit_kmers, h_kmers.data(), (nsub + 1) * 4);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);      // the host vectors go out of scope
    if (e != cudaSuccess) { ch->release(); delete ch; return set_err(GGCAT_B200_ERR_CUDA, "bucket file upload failed: %s", cudaGetErrorString(e)); }
    ch->d_desc = ch->desc.as<uint4>(); ch->d_payload = ch->payload.as<uint32_t>();
    ch->d_unit_cnt = ch->unit_cnt.as<uint32_t>(); ch->d_unit_off = ch->unit_off.as<uint32_t>();
    ch->d_unit_words = ch->unit_words.as<uint32_t>(); ch->d_unit_woff = ch->unit_woff.as<uint32_t>();
    ch->d_unit_kmers = ch->unit_kmers.as<uint32_t>();
    ch->h_cnt = h_cnt; ch->h_words = h_words; ch->h_kmers = h_kmers; ch->h_off = h_off; ch->h_woff = h_woff;
    ch->n_sk = a; ch->n_words = b; ch->n_kmers = km;
    c->chunks.push_back(ch);
    return 0;
}


// ---- partial unitigs on the device (unitigs.cuh) -------------------------------------------
int32_t ggcat_b200_partial_unitigs(ggcat_b200_ctx *c, uint32_t result_buckets_log, ggcat_b200_unitigs *out) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!out) return set_err(GGCAT_B200_ERR_INVALID, "null output");
    memset(out, 0, sizeof(*out));
    const FinalTable &f = c->fin;
    if (!f.unit_off || f.n_units == 0) return set_err(GGCAT_B200_ERR_STATE, "partial_unitigs before merge_bucket_range_device");
    if (c->wide_mode >= 0) return set_err(GGCAT_B200_ERR_INVALID, "partial unitigs are built on the 64-bit key path (seq-hash, k <= 31, no colours)");
    if (c->P.forward_only) return set_err(GGCAT_B200_ERR_INVALID, "partial unitigs of forward-only builds are not built on the device");
    if ((c->P.k & 1u) == 0) return set_err(GGCAT_B200_ERR_INVALID, "partial unitigs need odd k (self-complementary k-mers of even k, final_executor.rs:249-264, are not handled)");
    if (result_buckets_log > 15) return set_err(GGCAT_B200_ERR_INVALID, "result_buckets_log > 15");
    static_assert(sizeof(UnitigRec) == sizeof(ggcat_b200_unitig), "device record == ABI record");
    cudaStream_t st = c->stream;
    const uint64_t ne = f.n_entries;
    if (ne >= (1ull << 31)) return set_err(GGCAT_B200_ERR_INVALID, "table of %llu entries exceeds 2^31: build partial unitigs per bucket range", (unsigned long long)ne);
    if (ne == 0) return 0;
    const uint64_t word_cap = ne * ((c->P.k + 14) / 16 + 2) + 16;
    CU(c->ut_links.reserve(ne * 8)); CU(c->ut_visited.reserve(ne)); CU(c->ut_recs.reserve(ne * sizeof(UnitigRec)));
    CU(c->ut_bases.reserve(word_cap * 4)); CU(c->ut_counters.reserve(64));
    CU(cudaMemsetAsync(c->ut_counters.p, 0, 64, st));
    UnitigTable T;
    T.keys = f.keys_lo; T.cf = f.cf; T.unit_off = f.unit_off; T.n_units = f.n_units; T.first_unit = f.first_unit; T.k = c->P.k; T.n_entries = ne;
    UnitigOut O;
    O.recs = c->ut_recs.as<UnitigRec>(); O.bases = c->ut_bases.as<uint32_t>(); O.counters = c->ut_counters.as<unsigned long long>();
    O.rec_cap = ne; O.word_cap = word_cap; O.overflow = reinterpret_cast<uint32_t *>(c->ut_counters.as<unsigned long long>() + 4);
    O.result_bits = result_buckets_log;
    {
        LaunchTimer t(c, F_UNITIGS, 3);
        k_unitig_links<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(T, c->ut_links.as<uint32_t>(), c->ut_visited.as<uint8_t>());
        k_unitig_paths<<<(unsigned)((2 * ne + 255) / 256), 256, 0, st>>>(T, c->ut_links.as<uint32_t>(), c->ut_visited.as<uint8_t>(), O);
        k_unitig_cycles<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(T, c->ut_links.as<uint32_t>(), c->ut_visited.as<uint8_t>(), O);
    }
    CU(cudaMemcpyAsync(c->h_pinned, c->ut_counters.p, 40, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if ((uint32_t)c->h_pinned[4]) return set_err(GGCAT_B200_ERR_CAPACITY, "partial unitig output overflow (code %u)", (uint32_t)c->h_pinned[4]);
    const uint64_t nu = c->h_pinned[0], nw = c->h_pinned[1];
    const size_t rec_bytes = (nu * sizeof(UnitigRec) + 15) & ~(size_t)15;
    TRY(pinned_reserve(&c->h_unitigs, &c->h_unitigs_cap, rec_bytes + nw * 4 + 16));
    if (nu) CU(cudaMemcpyAsync(c->h_unitigs, c->ut_recs.p, nu * sizeof(UnitigRec), cudaMemcpyDeviceToHost, st));
    if (nw) CU(cudaMemcpyAsync(c->h_unitigs + rec_bytes, c->ut_bases.p, nw * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c->ut_n = nu; c->ut_words = nw;
    out->n_unitigs = nu; out->n_words = nw; out->n_kmers = c->h_pinned[2];
    out->unitigs = reinterpret_cast<const ggcat_b200_unitig *>(c->h_unitigs);
    out->bases = reinterpret_cast<const uint32_t *>(c->h_unitigs + rec_bytes);
    out->d_unitigs = c->ut_recs.p; out->d_bases = c->ut_bases.as<uint32_t>();
    return 0;
}


int32_t ggcat_b200_maximal_unitigs(ggcat_b200_ctx *c, ggcat_b200_unitigs *out) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!out) return set_err(GGCAT_B200_ERR_INVALID, "null output");
    memset(out, 0, sizeof(*out));
    const uint64_t n = c->ut_n;
    if (n == 0) return c->fin.unit_off ? 0 : set_err(GGCAT_B200_ERR_STATE, "maximal_unitigs before partial_unitigs");
    if (n >= (1ull << 31)) return set_err(GGCAT_B200_ERR_INVALID, "too many partial unitigs for one join");
    cudaStream_t st = c->stream;
    const uint64_t slots = 4 * n + 1024, word_cap = c->ut_words + n + 16;
    CU(c->mu_ht.reserve(slots * 8)); CU(c->mu_first.reserve(slots * 4)); CU(c->mu_partner.reserve(2 * n * 4)); CU(c->mu_visited.reserve(n));
    CU(c->mu_recs.reserve(n * sizeof(UnitigRec))); CU(c->mu_bases.reserve(word_cap * 4));
    CU(cudaMemsetAsync(c->ut_counters.p, 0, 64, st));
    JoinTable J;
    J.recs = c->ut_recs.as<UnitigRec>(); J.bases = c->ut_bases.as<uint32_t>(); J.n = n; J.k = c->P.k;
    J.ht_keys = c->mu_ht.as<unsigned long long>(); J.ht_first = c->mu_first.as<uint32_t>(); J.ht_slots = slots;
    J.partner = c->mu_partner.as<uint32_t>(); J.visited = c->mu_visited.as<uint8_t>();
    UnitigOut O;
    O.recs = c->mu_recs.as<UnitigRec>(); O.bases = c->mu_bases.as<uint32_t>(); O.counters = c->ut_counters.as<unsigned long long>();
    O.rec_cap = n; O.word_cap = word_cap; O.overflow = reinterpret_cast<uint32_t *>(c->ut_counters.as<unsigned long long>() + 4); O.result_bits = 0;
    {
        LaunchTimer t(c, F_UNITIGS, 4);
        k_join_init<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(J);
        k_join_ends<<<(unsigned)((2 * n + 255) / 256), 256, 0, st>>>(J, O.overflow);
        k_join_chains<<<(unsigned)((2 * n + 255) / 256), 256, 0, st>>>(J, O);
        k_join_cycles<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(J, O);
    }
    CU(cudaMemcpyAsync(c->h_pinned, c->ut_counters.p, 40, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if ((uint32_t)c->h_pinned[4] & 4u) return set_err(GGCAT_B200_ERR_INVALID, "maximal_unitigs: more than two open ends share a k-mer (partial unitigs of an incomplete bucket range?)");
    if ((uint32_t)c->h_pinned[4]) return set_err(GGCAT_B200_ERR_CAPACITY, "maximal unitig output overflow (code %u)", (uint32_t)c->h_pinned[4]);
    const uint64_t nu = c->h_pinned[0], nw = c->h_pinned[1];
    const size_t rec_bytes = (nu * sizeof(UnitigRec) + 15) & ~(size_t)15;
    TRY(pinned_reserve(&c->h_unitigs, &c->h_unitigs_cap, rec_bytes + nw * 4 + 16));
    if (nu) CU(cudaMemcpyAsync(c->h_unitigs, c->mu_recs.p, nu * sizeof(UnitigRec), cudaMemcpyDeviceToHost, st));
    if (nw) CU(cudaMemcpyAsync(c->h_unitigs + rec_bytes, c->mu_bases.p, nw * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    out->n_unitigs = nu; out->n_words = nw; out->n_kmers = c->h_pinned[2];
    out->unitigs = reinterpret_cast<const ggcat_b200_unitig *>(c->h_unitigs);
    out->bases = reinterpret_cast<const uint32_t *>(c->h_unitigs + rec_bytes);
    out->d_unitigs = c->mu_recs.p; out->d_bases = c->mu_bases.as<uint32_t>();
    return 0;
}

int32_t ggcat_b200_finish_bucketing(ggcat_b200_ctx *c, ggcat_b200_bucket_stats *stats) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    // no stream synchronisation here: the per-unit counts of every chunk arrive on the copy stream right after its
    // k_emit, so this returns while the last k_scatter is still running; everything that follows is stream-ordered
    // (ggcat_b200_synchronize() is there for callers that touch chunk buffers from another stream)
    TRY(flush_open_chunk(c));
    uint64_t sk = 0, km = 0, words = 0, segs = 0;
    for (Chunk *ch : c->chunks) {
        TRY(mirror_chunk(c, ch));
        sk += ch->n_sk; km += ch->n_kmers; words += ch->n_words; segs += ch->n_segments;
    }
    c->stats.n_superkmers = sk; c->stats.n_kmers = km; c->stats.payload_words = words;
    c->stats.n_buckets = (1u << c->P.b1) + 1; c->stats.n_units = c->P.n_units;
    // SequencesSplitter::valid_bases (crates/minimizer_bucketing/src/sequences_splitter.rs:15-40): bases of the N-free
    // segments of length >= k.  A segment of L bases split into n super-k-mers (consecutive ones overlap by k bases)
    // stores (L - k + 1) + (n - 1) k-mers, so  sum L = k-mers - super-k-mers + segments * k.
    c->stats.valid_bases = km - sk + segs * (uint64_t)c->P.k;
    c->finished = true;
    if (stats) *stats = c->stats;
    trace_host("finish_bucketing: return");
    return 0;
}

int32_t ggcat_b200_unit_sizes(ggcat_b200_ctx *c, uint64_t *n_superkmers, uint64_t *n_kmers) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!c->finished) return set_err(GGCAT_B200_ERR_STATE, "unit_sizes before finish_bucketing");
    for (uint32_t u = 0; u < c->P.n_units; u++) {
        uint64_t a = 0, b = 0;
        for (Chunk *ch : c->chunks)
            if (u >= ch->first_unit && u < ch->first_unit + ch->n_units) { a += ch->h_cnt[u - ch->first_unit]; b += ch->h_kmers[u - ch->first_unit]; }
        if (n_superkmers) n_superkmers[u] = a;
        if (n_kmers) n_kmers[u] = b;
    }
    return 0;
}

int32_t ggcat_b200_dump_superkmers(ggcat_b200_ctx *c, uint32_t bucket, ggcat_b200_superkmer *out, uint64_t cap, uint8_t *payload,
                                   uint64_t payload_cap, uint64_t *n_out, uint64_t *payload_bytes) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    return dump_superkmers_impl(c, bucket, out, cap, payload, payload_cap, n_out, payload_bytes);
}
static int32_t dump_superkmers_impl(ggcat_b200_ctx *c, uint32_t bucket, ggcat_b200_superkmer *out, uint64_t cap, uint8_t *payload,
                                    uint64_t payload_cap, uint64_t *n_out, uint64_t *payload_bytes) {
    if (!c->finished) return set_err(GGCAT_B200_ERR_STATE, "dump before finish_bucketing");
    const DevParams &P = c->P;
    if (bucket > (1u << P.b1)) return set_err(GGCAT_B200_ERR_INVALID, "bucket %u out of range", bucket);
    const uint32_t u0 = bucket << P.b2, u1 = (bucket + 1) << P.b2;
    uint64_t n = 0, words = 0;
    for (Chunk *ch : c->chunks) {
        const uint32_t a = std::max(u0, ch->first_unit), b = std::min(u1, ch->first_unit + ch->n_units);
        if (a >= b) continue;
        n += ch->h_off[b - ch->first_unit] - ch->h_off[a - ch->first_unit];
        words += ch->h_woff[b - ch->first_unit] - ch->h_woff[a - ch->first_unit];
    }
    if (n_out) *n_out = n;
    if (payload_bytes) *payload_bytes = words * 4;
    if (!out) return 0;
    if (cap < n || payload_cap < words * 4) return set_err(GGCAT_B200_ERR_CAPACITY, "dump buffers too small");
    uint64_t oi = 0, pw = 0;
    std::vector<uint4> hd;
    for (Chunk *ch : c->chunks) {
        const uint32_t a = std::max(u0, ch->first_unit), b = std::min(u1, ch->first_unit + ch->n_units);
        if (a >= b) continue;
        const uint32_t d0 = ch->h_off[a - ch->first_unit], d1 = ch->h_off[b - ch->first_unit];
        const uint32_t w0 = ch->h_woff[a - ch->first_unit], w1 = ch->h_woff[b - ch->first_unit];
        if (d1 == d0) continue;
        hd.resize(d1 - d0);
        CU(cudaMemcpyAsync(hd.data(), ch->d_desc + d0, (size_t)(d1 - d0) * 16, cudaMemcpyDeviceToHost, c->stream));
        // local chunks: payload slice of these units starts at word w0 (+ bias 0); imported: payload pointer is the slice
        const uint32_t *psrc = ch->d_payload + w0;
        CU(cudaMemcpyAsync(payload + pw * 4, psrc, (size_t)(w1 - w0) * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        // second bucket per descriptor from its unit
        uint32_t u = a;
        for (uint32_t i = 0; i < d1 - d0; i++) {
            while (ch->h_off[u + 1 - ch->first_unit] <= d0 + i) ++u;
            const uint4 d = hd[i];
            ggcat_b200_superkmer &s = out[oi++];
            const uint32_t rel = d.x - ch->word_bias;  // word index into d_payload
            s.payload_offset = (pw + (rel - w0)) * 4;
            s.len = d.y; s.color = d.w; s.bucket = (uint16_t)bucket; s.minimizer_pos = (uint16_t)(d.z & 0xFFFF);
            s.second_bucket = (uint8_t)((d.z >> 19) & 0xFF); s.flags = (uint8_t)((d.z >> 16) & 3); s.rc = (uint8_t)((d.z >> 18) & 1);
            s.pad = 0;
            (void)u;
        }
        pw += w1 - w0;
    }
    return 0;
}

int32_t ggcat_b200_merge_bucket_range_device(ggcat_b200_ctx *c, uint32_t first_bucket, uint32_t n_buckets, uint64_t *n_entries,
                                             uint64_t *unique_kmers, uint64_t *total_kmers) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!c->finished) return set_err(GGCAT_B200_ERR_STATE, "merge before finish_bucketing");
    const DevParams &P = c->P;
    const uint32_t nb_total = (1u << P.b1) + 1;
    if (n_buckets == 0 || first_bucket >= nb_total || first_bucket + n_buckets > nb_total)
        return set_err(GGCAT_B200_ERR_INVALID, "bucket range [%u,+%u) outside 0..%u", first_bucket, n_buckets, nb_total);
    // parts of ~part_kmers_dev records (whole buckets) bound the per-part scratch; every part appends to one final
    // table that stays in HBM.  Coloured builds fold in one piece.
    std::vector<uint64_t> bk(n_buckets, 0);
    uint64_t tot = 0, unit_max = 0;
    {
        std::vector<uint64_t> un((size_t)n_buckets << P.b2, 0);
        const uint32_t ua = first_bucket << P.b2, ub2 = (first_bucket + n_buckets) << P.b2;
        for (Chunk *ch : c->chunks) {
            const uint32_t lo = std::max(ua, ch->first_unit), hi = std::min(ub2, ch->first_unit + ch->n_units);
            for (uint32_t u = lo; u < hi; u++) un[u - ua] += ch->h_kmers[u - ch->first_unit];
        }
        for (size_t i = 0; i < un.size(); i++) { bk[i >> P.b2] += un[i]; unit_max = std::max(unit_max, un[i]); }
        for (uint32_t b = 0; b < n_buckets; b++) tot += bk[b];
    }
    // a part must hold enough big units to fill the grid of the one-CTA-per-unit kernels (k_partition_units,
    // k_finish_units): at human-scale inputs a unit has > 1 M records and 192 M records are only ~150 units
    const uint64_t part_target = c->part_dev_fixed ? c->part_kmers_dev
        : std::min<uint64_t>(std::max<uint64_t>(c->part_kmers_dev, 2ull * c->sm_count * unit_max), 640ull << 20);
    std::vector<std::pair<uint32_t, uint32_t>> parts;
    if (c->wide_mode == MODE_COLOR || tot <= part_target + part_target / 2) parts.push_back({first_bucket, n_buckets});
    else {
        uint32_t b0 = 0; uint64_t acc = 0;
        for (uint32_t b = 0; b < n_buckets; b++) {
            acc += bk[b];
            if (acc >= part_target || b + 1 == n_buckets) { parts.push_back({first_bucket + b0, b + 1 - b0}); b0 = b + 1; acc = 0; }
        }
    }
    uint64_t eb = 0, uq = 0, tk = 0;
    uint32_t ub = 0;
    for (auto &pr : parts) {
        PartBase pb; pb.eb = eb; pb.ub = ub; pb.cap_total = std::max<uint64_t>(tot, 1);
        uint64_t ne = 0, u1 = 0, t1 = 0;
        TRY(merge_range_device(c, pr.first, pr.second, &ne, &u1, &t1, pb));
        eb += ne; uq += u1; tk += t1;
        ub += pr.second << P.b2;
    }
    c->final_hint = eb;
    c->fin.first_unit = first_bucket << P.b2; c->fin.n_units = n_buckets << P.b2;
    c->fin.total_kmers = tk; c->fin.unique_kmers = uq;
    if (n_entries) *n_entries = eb;
    if (unique_kmers) *unique_kmers = uq;
    if (total_kmers) *total_kmers = tk;
    return 0;
}

int32_t ggcat_b200_device_table(ggcat_b200_ctx *c, ggcat_b200_table *out) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!out) return set_err(GGCAT_B200_ERR_INVALID, "null table");
    memset(out, 0, sizeof(*out));
    const FinalTable &f = c->fin;
    if (!f.unit_off || f.n_units == 0) return set_err(GGCAT_B200_ERR_STATE, "device_table before merge_bucket_range_device");
    out->n_entries = f.n_entries; out->keys_lo = f.keys_lo; out->keys_hi = f.keys_hi; out->count_flags = f.cf;
    out->src_kmers = f.src; out->src_kmer_words = f.src ? c->src_words : 0;
    out->first_unit = f.first_unit; out->n_units = f.n_units; out->unit_offsets = f.unit_off;
    out->color_offsets = f.color_off; out->colors = f.colors;   // color_offsets has n_entries entries on the device (the end is n_colors)
    out->total_kmers = f.total_kmers; out->unique_kmers = f.unique_kmers;
    return 0;
}

// Host-table growth for the part-wise D2H: keeps what was already copied.
static int32_t host_table_reserve(ggcat_b200_ctx *c, HostTable *t, uint64_t need_entries, uint64_t copied, bool wide) {
    if (need_entries <= t->cap_entries) return 0;
    CU(cudaStreamSynchronize(c->copy_stream));
    const uint64_t cap = std::max<uint64_t>(need_entries + need_entries / 2, 1024);
    uint64_t *nk = nullptr, *nh = nullptr, *ns = nullptr; uint32_t *nc = nullptr;
    bool ok = cudaMallocHost((void **)&nk, cap * 8) == cudaSuccess && cudaMallocHost((void **)&nc, cap * 4) == cudaSuccess;
    if (ok && wide) ok = cudaMallocHost((void **)&nh, cap * 8) == cudaSuccess;
    if (ok && t->with_src) ok = cudaMallocHost((void **)&ns, cap * 8 * c->src_words) == cudaSuccess;
    if (!ok) { cudaFreeHost(nk); cudaFreeHost(nc); cudaFreeHost(nh); cudaFreeHost(ns); return set_err(GGCAT_B200_ERR_CUDA, "pinned table allocation failed"); }
    if (copied) {
        memcpy(nk, t->keys, copied * 8); memcpy(nc, t->cf, copied * 4);
        if (wide) memcpy(nh, t->keys_hi, copied * 8);
        if (t->with_src) memcpy(ns, t->src, copied * 8 * c->src_words);
    }
    cudaFreeHost(t->keys); cudaFreeHost(t->cf); cudaFreeHost(t->keys_hi); cudaFreeHost(t->src);
    t->keys = nk; t->cf = nc; t->keys_hi = nh; t->src = ns; t->cap_entries = cap;
    return 0;
}

int32_t ggcat_b200_merge_bucket_range(ggcat_b200_ctx *c, uint32_t first_bucket, uint32_t n_buckets, ggcat_b200_table *out) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!out) return set_err(GGCAT_B200_ERR_INVALID, "null table");
    memset(out, 0, sizeof(*out));
    if (!c->finished) return set_err(GGCAT_B200_ERR_STATE, "merge before finish_bucketing");
    const DevParams &P = c->P;
    const uint32_t nb_total = (1u << P.b1) + 1;
    if (n_buckets == 0 || first_bucket >= nb_total || first_bucket + n_buckets > nb_total)
        return set_err(GGCAT_B200_ERR_INVALID, "bucket range [%u,+%u) outside 0..%u", first_bucket, n_buckets, nb_total);
    const uint32_t nu = n_buckets << P.b2;
    const bool colored = c->wide_mode == MODE_COLOR, wide = c->wide_mode >= 0, with_src = c->wide_mode == MODE_RK128;
    // split the range into parts of ~part_kmers records (whole buckets); coloured builds fold in one piece
    std::vector<uint64_t> bk(n_buckets, 0);
    uint64_t tot = 0;
    for (uint32_t b = 0; b < n_buckets; b++) {
        for (uint32_t u = (first_bucket + b) << P.b2; u < ((first_bucket + b + 1) << P.b2); u++)
            for (Chunk *ch : c->chunks)
                if (u >= ch->first_unit && u < ch->first_unit + ch->n_units) bk[b] += ch->h_kmers[u - ch->first_unit];
        tot += bk[b];
    }
    std::vector<std::pair<uint32_t, uint32_t>> parts;  // (first bucket, count)
    if (colored || tot <= c->part_kmers + c->part_kmers / 2) parts.push_back({first_bucket, n_buckets});
    else if (c->part_scheme == 1) {
        // The D2H copies are the critical path (the merge produces the table faster than PCIe drains it): a small first part
        // starts the copy engine early, equal parts of ~part_kmers keep it busy, bounded by the device part size
        const uint64_t n_eq = std::max<uint64_t>(2, (tot + c->part_kmers - 1) / c->part_kmers);
        const uint64_t eq = std::min<uint64_t>(c->part_kmers_dev, (tot + n_eq - 1) / n_eq);
        uint32_t b0 = 0; uint64_t acc = 0, target = std::max<uint64_t>(eq / 2, 1);
        for (uint32_t b = 0; b < n_buckets; b++) {
            acc += bk[b];
            if (acc >= target || b + 1 == n_buckets) { parts.push_back({first_bucket + b0, b + 1 - b0}); b0 = b + 1; acc = 0; target = eq; }
        }
        if (parts.size() >= 2) {   // a tiny last part only adds a host round trip: fold it into its predecessor
            uint64_t last = 0;
            for (uint32_t b = parts.back().first - first_bucket; b < n_buckets; b++) last += bk[b];
            if (last < eq / 4) { const auto lp = parts.back(); parts.pop_back(); parts.back().second += lp.second; }
        }
    } else {
        // shrinking parts: every part costs a fixed host round trip and only the LAST part's D2H is exposed, so the
        // first part takes half of what is left (bounded by the device part size), the last ones ~part_kmers
        uint32_t b0 = 0; uint64_t acc = 0, left = tot;
        uint64_t target = std::min<uint64_t>(c->part_kmers_dev, std::max<uint64_t>(c->part_kmers, left / 2));
        for (uint32_t b = 0; b < n_buckets; b++) {
            acc += bk[b];
            if (acc >= target || b + 1 == n_buckets) {
                parts.push_back({first_bucket + b0, b + 1 - b0}); b0 = b + 1;
                left -= std::min(left, acc); acc = 0;
                target = std::min<uint64_t>(c->part_kmers_dev, std::max<uint64_t>(c->part_kmers, left / 2));
                if (left < c->part_kmers + c->part_kmers / 2) target = left + 1;   // the rest in one piece
            }
        }
    }
    HostTable *t = nullptr;
    for (size_t i = 0; i < c->free_tables.size(); i++) {
        HostTable *q = c->free_tables[i];
        if (q->cap_units >= nu + 1 && q->wide == wide && q->colored == colored && q->with_src == with_src) { t = q; c->free_tables.erase(c->free_tables.begin() + i); break; }
    }
    if (!t) {
        t = new HostTable();
        t->cap_units = nu + 1; t->wide = wide; t->colored = colored; t->with_src = with_src;
        if (cudaMallocHost((void **)&t->unit_offsets, t->cap_units * 8) != cudaSuccess) { delete t; return set_err(GGCAT_B200_ERR_CUDA, "pinned table allocation failed"); }
    }
    auto fail = [&](int32_t rc) { c->free_tables.push_back(t); return rc; };
    uint64_t eb = 0, uq = 0, tk = 0;
    uint32_t ub = 0;
    for (size_t pi = 0; pi < parts.size(); pi++) {
        PartBase pb; pb.eb = eb; pb.ub = ub; pb.cap_total = std::max<uint64_t>(tot, 1);
        uint64_t ne = 0, u1 = 0, t1 = 0;
        int32_t rc = merge_range_device(c, parts[pi].first, parts[pi].second, &ne, &u1, &t1, pb);  // ends with a stream sync
        if (rc) return fail(rc);
        uq += u1; tk += t1;
        if (!colored) {
            rc = host_table_reserve(c, t, eb + ne + (pi + 1 < parts.size() ? ne : 0), eb, wide);
            if (rc) return fail(rc);
            if (ne) {  // the part is complete on the compute stream (synchronised): copy it out while the next part merges
                const FinalTable &f = c->fin;
                cudaError_t e1 = cudaMemcpyAsync(t->keys + eb, f.keys_lo + eb, ne * 8, cudaMemcpyDeviceToHost, c->copy_stream);
                if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(t->cf + eb, f.cf + eb, ne * 4, cudaMemcpyDeviceToHost, c->copy_stream);
                if (e1 == cudaSuccess && wide) e1 = cudaMemcpyAsync(t->keys_hi + eb, f.keys_hi + eb, ne * 8, cudaMemcpyDeviceToHost, c->copy_stream);
                if (e1 == cudaSuccess && with_src) e1 = cudaMemcpyAsync(t->src + (size_t)c->src_words * eb, f.src + (size_t)c->src_words * eb, ne * 8 * c->src_words, cudaMemcpyDeviceToHost, c->copy_stream);
                if (e1 != cudaSuccess) return fail(set_err(GGCAT_B200_ERR_CUDA, "table copy failed: %s", cudaGetErrorString(e1)));
            }
            eb += ne;
        } else eb = ne;
        ub += parts[pi].second << P.b2;
    }
    const FinalTable &f = c->fin;
    const uint64_t ne = eb, ncol = f.n_colors;
    if (colored) {
        int32_t rc = host_table_reserve(c, t, ne, 0, wide);
        if (rc) return fail(rc);
        if (t->cap_colors < ncol || !t->colors) {
            cudaFreeHost(t->colors);
            t->colors = nullptr;
            t->cap_colors = std::max<uint64_t>(ncol + ncol / 4, 1024);
            if (cudaMallocHost((void **)&t->colors, t->cap_colors * 4) != cudaSuccess) { t->cap_colors = 0; return fail(set_err(GGCAT_B200_ERR_CUDA, "pinned table allocation failed")); }
        }
        if (t->cap_coloff < ne + 1 || !t->color_offsets) {
            cudaFreeHost(t->color_offsets);
            t->color_offsets = nullptr;
            t->cap_coloff = std::max<uint64_t>(ne + ne / 4 + 1, 1024);
            if (cudaMallocHost((void **)&t->color_offsets, t->cap_coloff * 8) != cudaSuccess) { t->cap_coloff = 0; return fail(set_err(GGCAT_B200_ERR_CUDA, "pinned table allocation failed")); }
        }
        if (ne) {
            CU(cudaMemcpyAsync(t->keys, f.keys_lo, ne * 8, cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaMemcpyAsync(t->cf, f.cf, ne * 4, cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaMemcpyAsync(t->keys_hi, f.keys_hi, ne * 8, cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaMemcpyAsync(t->color_offsets, f.color_off, ne * 8, cudaMemcpyDeviceToHost, c->copy_stream));
            if (ncol) CU(cudaMemcpyAsync(t->colors, f.colors, ncol * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        }
    }
    CU(cudaMemcpyAsync(t->unit_offsets, f.unit_off, ((size_t)nu + 1) * 8, cudaMemcpyDeviceToHost, c->copy_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    if (colored) t->color_offsets[ne] = ncol;
    out->n_entries = ne; out->keys_lo = t->keys; out->keys_hi = wide ? t->keys_hi : nullptr; out->count_flags = t->cf;
    out->src_kmers = with_src ? t->src : nullptr; out->src_kmer_words = with_src ? c->src_words : 0;
    out->first_unit = first_bucket << P.b2; out->n_units = nu; out->unit_offsets = t->unit_offsets;
    out->color_offsets = colored ? t->color_offsets : nullptr; out->colors = colored ? t->colors : nullptr;
    out->total_kmers = tk; out->unique_kmers = uq; out->opaque = t;
    if (!colored) c->final_hint = ne;
    return 0;
}

int32_t ggcat_b200_release_table(ggcat_b200_ctx *c, ggcat_b200_table *table) {
    if (!c || !table) return set_err(GGCAT_B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock__(c->mu);
    if (table->opaque) c->free_tables.push_back(reinterpret_cast<HostTable *>(table->opaque));
    memset(table, 0, sizeof(*table));
    return 0;
}

uint32_t ggcat_b200_n_chunks(ggcat_b200_ctx *c) { return c ? (uint32_t)c->chunks.size() : 0; }

int32_t ggcat_b200_export_chunk_slice(ggcat_b200_ctx *c, uint32_t chunk, uint32_t first_unit, uint32_t n_units,
                                      ggcat_b200_chunk_slice *out) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!c->finished) return set_err(GGCAT_B200_ERR_STATE, "export before finish_bucketing");
    if (!out || chunk >= c->chunks.size()) return set_err(GGCAT_B200_ERR_INVALID, "bad chunk index");
    Chunk *ch = c->chunks[chunk];
    if (ch->imported) return set_err(GGCAT_B200_ERR_INVALID, "cannot export an imported chunk");
    if (first_unit + n_units > ch->n_units) return set_err(GGCAT_B200_ERR_INVALID, "unit range outside chunk");
    CU(cudaStreamSynchronize(c->stream));   // the pointers handed out are read from other streams: phase 1 must be complete
    const uint32_t d0 = ch->h_off[first_unit], d1 = ch->h_off[first_unit + n_units];
    const uint32_t w0 = ch->h_woff[first_unit], w1 = ch->h_woff[first_unit + n_units];
    out->n_superkmers = d1 - d0; out->n_words = w1 - w0; out->word_bias = w0;
    out->d_descriptors = ch->d_desc + d0; out->d_payload = ch->d_payload + w0;
    out->d_unit_counts = ch->d_unit_cnt + first_unit; out->d_unit_words = ch->d_unit_words + first_unit;
    out->d_unit_kmers = ch->d_unit_kmers + first_unit;
    out->h_unit_counts = ch->h_cnt.data() + first_unit; out->h_unit_words = ch->h_words.data() + first_unit;
    out->h_unit_kmers = ch->h_kmers.data() + first_unit;
    return 0;
}

int32_t ggcat_b200_import_chunk_slice(ggcat_b200_ctx *c, uint32_t first_unit, uint32_t n_units, const ggcat_b200_chunk_slice *s) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    if (!s) return set_err(GGCAT_B200_ERR_INVALID, "null slice");
    if (first_unit + n_units > c->P.n_units) return set_err(GGCAT_B200_ERR_INVALID, "unit range outside 0..%u", c->P.n_units);
    if (s->n_superkmers == 0) return 0;
    Chunk *ch = new Chunk();
    ch->imported = true; ch->first_unit = first_unit; ch->n_units = n_units; ch->word_bias = (uint32_t)s->word_bias;
    ch->d_desc = reinterpret_cast<const uint4 *>(s->d_descriptors); ch->d_payload = s->d_payload;
    ch->d_unit_cnt = s->d_unit_counts; ch->d_unit_words = s->d_unit_words; ch->d_unit_kmers = s->d_unit_kmers;
    cudaError_t e = ch->unit_off.reserve(((size_t)n_units + 2) * 4);
    if (e == cudaSuccess) e = ch->unit_woff.reserve(((size_t)n_units + 2) * 4);
    if (e == cudaSuccess) e = c->totals.reserve(8 * 8);
    if (e != cudaSuccess) { ch->release(); delete ch; return set_err(GGCAT_B200_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    k_exclusive_scan_u32<<<1, 1024, 0, c->stream>>>(s->d_unit_counts, ch->unit_off.as<uint32_t>(), n_units,
                                                     c->totals.as<unsigned long long>() + 3);
    k_exclusive_scan_u32<<<1, 1024, 0, c->stream>>>(s->d_unit_words, ch->unit_woff.as<uint32_t>(), n_units,
                                                     c->totals.as<unsigned long long>() + 4);
    ch->d_unit_off = ch->unit_off.as<uint32_t>();
    ch->d_unit_woff = ch->unit_woff.as<uint32_t>();   // relative to the slice's payload pointer
    if (s->h_unit_counts && s->h_unit_words && s->h_unit_kmers) {
        // host copies supplied by the transport: no device read-back, no synchronisation
        const size_t nu = n_units;
        ch->h_cnt.assign(s->h_unit_counts, s->h_unit_counts + nu); ch->h_cnt.push_back(0);
        ch->h_words.assign(s->h_unit_words, s->h_unit_words + nu); ch->h_words.push_back(0);
        ch->h_kmers.assign(s->h_unit_kmers, s->h_unit_kmers + nu); ch->h_kmers.push_back(0);
        ch->h_off.resize(nu + 1); ch->h_woff.resize(nu + 1);
        uint64_t a = 0, b = 0, km = 0;
        for (size_t u = 0; u < nu; u++) {
            ch->h_off[u] = (uint32_t)a; ch->h_woff[u] = (uint32_t)b;
            a += ch->h_cnt[u]; b += ch->h_words[u]; km += ch->h_kmers[u];
        }
        ch->h_off[nu] = (uint32_t)a; ch->h_woff[nu] = (uint32_t)b;
        ch->n_sk = a; ch->n_words = b; ch->n_kmers = km;
        if (a != s->n_superkmers) {
            ch->release(); delete ch;
            return set_err(GGCAT_B200_ERR_INVALID, "import: unit counts sum to %llu, slice holds %llu super-k-mers", (unsigned long long)a, (unsigned long long)s->n_superkmers);
        }
        c->chunks.push_back(ch);
        return 0;
    }
    const int32_t rc = mirror_chunk(c, ch);
    if (rc) { ch->release(); delete ch; return rc; }
    c->chunks.push_back(ch);   // registered only once nothing can fail any more
    // imported payload pointer already addresses the slice: word offsets inside it are (woff - word_bias)
    return 0;
}

int32_t ggcat_b200_drop_local_chunks(ggcat_b200_ctx *c) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    cudaStreamSynchronize(c->stream);
    std::vector<Chunk *> keep;
    for (Chunk *ch : c->chunks) {
        if (ch->imported) keep.push_back(ch);
        else c->chunk_pool.push_back(ch);  // recycle the device buffers
    }
    c->chunks.swap(keep);
    return 0;
}


// ---- NVLink peer exchange (peer.cuh) -----------------------------------------------------------------------------

int32_t ggcat_b200_owner_range(uint32_t buckets_count_log, uint32_t rank, uint32_t world, uint32_t *first_bucket, uint32_t *n_buckets) {
    if (world == 0 || rank >= world || buckets_count_log > 13 || world > (1u << buckets_count_log))
        return set_err(GGCAT_B200_ERR_INVALID, "owner_range: rank %u of %u ranks for %u buckets", rank, world, 1u << buckets_count_log);
    const uint32_t a = owner_first_bucket(buckets_count_log, rank, world), b = owner_first_bucket(buckets_count_log, rank + 1, world);
    if (first_bucket) *first_bucket = a;
    if (n_buckets) *n_buckets = b - a;
    return 0;
}

int32_t ggcat_b200_peer_init(ggcat_b200_ctx *c, uint32_t rank, uint32_t world, uint64_t arena_bytes, ggcat_b200_peer_handle *out) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    PeerState &ps = c->peer;
    if (ps.inited) return set_err(GGCAT_B200_ERR_STATE, "peer_init called twice");
    if (world == 0 || world > (uint32_t)PEER_MAX_WORLD || rank >= world || world > (1u << c->P.b1))
        return set_err(GGCAT_B200_ERR_INVALID, "peer_init: rank %u of %u ranks (max %d, <= %u buckets)", rank, world, PEER_MAX_WORLD, 1u << c->P.b1);
    if (!out) return set_err(GGCAT_B200_ERR_INVALID, "null handle");
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(ggcat_b200_peer_handle), "IPC handle must fit the ABI struct");
    memset(out, 0, sizeof(*out));
    ps.rank = rank; ps.world = world;
    if (const char *e = getenv("GGCAT_B200_PEER_SLICES")) ps.meta_slots = (uint32_t)std::max(1, std::min(PEER_MAX_SLICES, atoi(e)));
    if (world > 1) {
        // every region starts with the slice table and S meta + scratch slots sized by the LARGEST owner range
        uint32_t nu_max = 0;
        for (uint32_t r = 0; r < world; r++)
            nu_max = std::max(nu_max, (owner_first_bucket(c->P.b1, r + 1, world) - owner_first_bucket(c->P.b1, r, world)) << c->P.b2);
        const uint64_t min_region = peer_data_off(ps.meta_slots, nu_max) + (1ull << 20);
        ps.region_bytes = std::max<uint64_t>(arena_bytes / world, min_region) & ~255ull;
        ps.arena_bytes = PEER_HDR_BYTES + ps.region_bytes * world;
        CU(cudaMalloc((void **)&ps.arena, ps.arena_bytes));
        CU(cudaMemset(ps.arena, 0, PEER_HDR_BYTES));
        CU(cudaDeviceSynchronize());
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, ps.arena));
        memcpy(out->bytes, &h, sizeof(h));
        CU(ps.d_err.reserve(16));
        CU(cudaMemset(ps.d_err.p, 0, 16));
        {   // the count / header pushes are a few CTAs that must not queue behind the thousands of pending CTAs of k_scatter:
            // highest stream priority (their blocks are dispatched first as soon as any SM frees a slot)
            int lo_prio = 0, hi_prio = 0;
            CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            CU(cudaStreamCreateWithPriority(&ps.meta_stream, cudaStreamNonBlocking, hi_prio));
        }
        CU(cudaStreamCreateWithFlags(&ps.data_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ps.ev_scatter, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ps.ev_data, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ps.ev_meta, cudaEventDisableTiming));
        const size_t stage_bytes = (size_t)(ps.meta_slots + 1) * peer_block_bytes(world);
        CU(ps.d_stage.reserve(stage_bytes));
        TRY(pinned_reserve(&ps.h_stage, &ps.h_stage_cap, stage_bytes));
    }
    ps.inited = true;
    return 0;
}

int32_t ggcat_b200_peer_connect(ggcat_b200_ctx *c, const ggcat_b200_peer_handle *handles) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    PeerState &ps = c->peer;
    if (!ps.inited) return set_err(GGCAT_B200_ERR_STATE, "peer_connect before peer_init");
    if (ps.connected) return set_err(GGCAT_B200_ERR_STATE, "peer_connect called twice");
    if (ps.world > 1 && !handles) return set_err(GGCAT_B200_ERR_INVALID, "null handles");
    for (uint32_t r = 0; r < ps.world && ps.world > 1; r++) {
        if (r == ps.rank) { ps.peer_arena[r] = ps.arena; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles[r].bytes, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return set_err(GGCAT_B200_ERR_CUDA, "cudaIpcOpenMemHandle(rank %u): %s (ranks must be separate processes on one NVLink box)", r,
                           cudaGetErrorString(e));
        ps.peer_arena[r] = reinterpret_cast<uint8_t *>(p);
    }
    ps.connected = true;
    return 0;
}

int32_t ggcat_b200_peer_exchange(ggcat_b200_ctx *c) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    PeerState &ps = c->peer;
    if (!ps.connected) return set_err(GGCAT_B200_ERR_STATE, "peer_exchange before peer_connect");
    if (!c->finished) return set_err(GGCAT_B200_ERR_STATE, "peer_exchange before finish_bucketing");
    const uint32_t W = ps.world, me = ps.rank;
    if (W == 1) return 0;
    const DevParams &P = c->P;
    cudaStream_t st = c->stream;
    std::vector<Chunk *> local;
    for (Chunk *ch : c->chunks) {
        if (ch->imported) return set_err(GGCAT_B200_ERR_STATE, "peer_exchange: chunks were already exchanged");
        local.push_back(ch);
    }
    trace_host("peer_exchange: enter");
    if (!ps.build_started) TRY(peer_begin_build(c));
    // chunks that were not pushed eagerly (peer arena connected in the middle of a build)
    for (size_t j = ps.n_pushed; j < local.size(); j++) {
        CU(cudaEventRecord(ps.ev_scatter, st));
        TRY(peer_push_chunk(c, local[j]));
    }
    const uint32_t epoch = ps.epoch, S = ps.meta_slots;
    auto first_unit_of = [&](uint32_t r) { return owner_first_bucket(P.b1, r, W) << P.b2; };
    PeerHdrPtrs hp;
    memset(&hp, 0, sizeof(hp));
    for (uint32_t r = 0; r < W; r++) hp.h[r] = reinterpret_cast<PeerHdr *>(ps.peer_arena[r]);
    const unsigned long long timeout_ns = 20ull * 1000000000ull;
    // ---- region headers (slice count of this build) -> every owner; then "my counts are complete" / wait for everybody's
    {
        const size_t bb = peer_block_bytes(W);
        uint8_t *hb = ps.h_stage + (size_t)S * bb, *db = ps.d_stage.as<uint8_t>() + (size_t)S * bb;
        PeerJob *jobs = reinterpret_cast<PeerJob *>(hb);
        RegionHdr *hdrs = reinterpret_cast<RegionHdr *>(hb + peer_block_entries_off(W));
        uint32_t nj = 0;
        for (uint32_t d = 0; d < W; d++) {
            if (d == me) continue;
            RegionHdr &rh = hdrs[d];
            memset(&rh, 0, sizeof(rh));
            rh.n_slices = ps.n_pushed; rh.overflow = ps.overflow[d] ? 1u : 0u; rh.n_units = first_unit_of(d + 1) - first_unit_of(d); rh.epoch = epoch;
            uint8_t *dst = ps.peer_arena[d] + PEER_HDR_BYTES + (uint64_t)me * ps.region_bytes;
            jobs[nj++] = {db + peer_block_entries_off(W) + (size_t)d * 64, dst, 64ull, (d + W - me - 1) % W, 0u};
        }
        CU(cudaMemcpyAsync(db, hb, bb, cudaMemcpyHostToDevice, ps.meta_stream));
        c->fam_launches[F_PEER_PUSH] += 2;
        k_peer_push<<<W - 1, 64, 0, ps.meta_stream>>>(reinterpret_cast<const PeerJob *>(db), nj, W - 1);
        k_peer_sync<<<1, PEER_MAX_WORLD, 0, ps.meta_stream>>>(hp, me, W, PEER_FLAG_META, epoch, 1u, 1u, ps.d_err.as<uint32_t>(), timeout_ns);
    }
    // ---- receive: headers, slice tables and per-unit counts of every source (the bulk data may still be in flight)
    const uint32_t my_fu = first_unit_of(me), my_nu = first_unit_of(me + 1) - my_fu;
    const uint64_t s_meta = peer_s_meta(my_nu), s_uoff = peer_s_uoff(my_nu);
    uint32_t guess = std::max<uint32_t>(ps.n_pushed, 1);
    for (int attempt = 0; attempt < 2; attempt++) {
        const size_t stride = (size_t)(PEER_META_OFF + (uint64_t)guess * s_meta);
        TRY(pinned_reserve(&ps.h_recv, &ps.h_recv_cap, (size_t)W * stride));
        for (uint32_t s2 = 0; s2 < W; s2++) {
            if (s2 == me) continue;
            const uint8_t *reg = ps.arena + PEER_HDR_BYTES + (uint64_t)s2 * ps.region_bytes;
            CU(cudaMemcpyAsync(ps.h_recv + (size_t)s2 * stride, reg, std::min<uint64_t>(stride, ps.region_bytes), cudaMemcpyDeviceToHost, ps.meta_stream));
        }
        CU(cudaMemcpyAsync(c->h_pinned + 9, ps.d_err.p, 4, cudaMemcpyDeviceToHost, ps.meta_stream));
        CU(cudaEventRecord(ps.ev_meta, ps.meta_stream));
        CU(cudaEventSynchronize(ps.ev_meta));
        trace_host("peer_exchange: counts of all sources on the host");
        CU(cudaGetLastError());
        if ((uint32_t)c->h_pinned[9]) {
            CU(cudaMemset(ps.d_err.p, 0, 16));
            return set_err(GGCAT_B200_ERR_STATE, "peer_exchange: rank %u did not answer within 20 s", (uint32_t)c->h_pinned[9] - 1);
        }
        uint32_t need = 0;
        for (uint32_t s2 = 0; s2 < W; s2++) {
            if (s2 == me) continue;
            const RegionHdr *rh = reinterpret_cast<const RegionHdr *>(ps.h_recv + (size_t)s2 * stride);
            if (rh->epoch != epoch) return set_err(GGCAT_B200_ERR_STATE, "peer_exchange: rank %u delivered epoch %u, expected %u", s2, rh->epoch, epoch);
            if (rh->overflow) return set_err(GGCAT_B200_ERR_CAPACITY, "peer_exchange: the slices of rank %u do not fit a %llu-byte arena region; "
                                             "give peer_init a larger arena", s2, (unsigned long long)ps.region_bytes);
            if (rh->n_units != my_nu) return set_err(GGCAT_B200_ERR_INVALID, "peer_exchange: rank %u assumes %u units for this owner, expected %u", s2, rh->n_units, my_nu);
            need = std::max(need, rh->n_slices);
        }
        if (need <= guess) break;
        guess = need;
    }
    for (uint32_t d = 0; d < W; d++)
        if (d != me && ps.overflow[d])
            return set_err(GGCAT_B200_ERR_CAPACITY, "peer_exchange: local slices do not fit a %llu-byte arena region; give peer_init a larger arena",
                           (unsigned long long)ps.region_bytes);
    // ---- compute stream: my bulk pushes are complete -> tell the owners, wait for my sources; everything that follows
    //      (offset scans of the received slices, the merge) is stream-ordered behind it -- no host synchronisation here
    CU(cudaEventRecord(ps.ev_data, ps.data_stream));
    CU(cudaStreamWaitEvent(st, ps.ev_data, 0));
    {
        LaunchTimer t(c, F_PEER, 1);
        k_peer_sync<<<1, PEER_MAX_WORLD, 0, st>>>(hp, me, W, PEER_FLAG_READY, epoch, 1u, 1u, ps.d_err.as<uint32_t>(), timeout_ns);
    }
    const size_t stride = (size_t)(PEER_META_OFF + (uint64_t)guess * s_meta);
    CU(c->totals.reserve(8 * 8));
    // per-unit offset scans of the received slices: they need only the counts (already here), so they run on the meta stream
    // in batched launches while the bulk data is still in flight; the compute stream picks them up through ev_meta
    ScanJobs sj;
    uint32_t n_sj = 0;
    sj.n = my_nu; sj.pad = 0;
    auto flush_scans = [&]() -> int32_t {
        if (!n_sj) return 0;
        c->fam_launches[F_SCAN] += 1;
        k_exclusive_scan_u32_jobs<<<n_sj, 1024, 0, ps.meta_stream>>>(sj);
        n_sj = 0;
        return 0;
    };
    for (uint32_t s2 = 0; s2 < W; s2++) {
        if (s2 == me) continue;
        const uint8_t *hr = ps.h_recv + (size_t)s2 * stride;
        const RegionHdr *rh = reinterpret_cast<const RegionHdr *>(hr);
        const PeerSlice *tb = reinterpret_cast<const PeerSlice *>(hr + PEER_TABLE_OFF);
        uint8_t *reg = ps.arena + PEER_HDR_BYTES + (uint64_t)s2 * ps.region_bytes;
        for (uint32_t j = 0; j < rh->n_slices; j++) {
            if (tb[j].n_sk == 0) continue;
            Chunk *ch = new Chunk();
            ch->imported = true; ch->first_unit = my_fu; ch->n_units = my_nu; ch->word_bias = (uint32_t)tb[j].word_bias;
            ch->d_desc = reinterpret_cast<const uint4 *>(reg + tb[j].desc_off);
            ch->d_payload = reinterpret_cast<const uint32_t *>(reg + tb[j].pay_off);
            const uint32_t *dm = reinterpret_cast<const uint32_t *>(reg + PEER_META_OFF + (uint64_t)j * s_meta);
            ch->d_unit_cnt = dm; ch->d_unit_words = dm + my_nu; ch->d_unit_kmers = dm + 2 * (size_t)my_nu;
            uint32_t *uoff = reinterpret_cast<uint32_t *>(reg + PEER_META_OFF + (uint64_t)S * s_meta + (uint64_t)j * s_uoff);
            uint32_t *uwoff = uoff + (s_uoff / 8);
            sj.in[n_sj] = dm; sj.out[n_sj] = uoff; ++n_sj;
            sj.in[n_sj] = dm + my_nu; sj.out[n_sj] = uwoff; ++n_sj;
            if (n_sj + 2 > (uint32_t)SCAN_MAX_JOBS) TRY(flush_scans());
            ch->d_unit_off = uoff; ch->d_unit_woff = uwoff;
            const uint32_t *hm = reinterpret_cast<const uint32_t *>(hr + PEER_META_OFF + (uint64_t)j * s_meta);
            const size_t nu = my_nu;
            ch->h_cnt.assign(hm, hm + nu); ch->h_cnt.push_back(0);
            ch->h_words.assign(hm + nu, hm + 2 * nu); ch->h_words.push_back(0);
            ch->h_kmers.assign(hm + 2 * nu, hm + 3 * nu); ch->h_kmers.push_back(0);
            ch->h_off.resize(nu + 1); ch->h_woff.resize(nu + 1);
            uint64_t a = 0, b = 0, km = 0;
            for (size_t u = 0; u < nu; u++) {
                ch->h_off[u] = (uint32_t)a; ch->h_woff[u] = (uint32_t)b;
                a += ch->h_cnt[u]; b += ch->h_words[u]; km += ch->h_kmers[u];
            }
            ch->h_off[nu] = (uint32_t)a; ch->h_woff[nu] = (uint32_t)b;
            ch->n_sk = a; ch->n_words = b; ch->n_kmers = km;
            ps.last_received += a * 16 + b * 4 + 3ull * nu * 4;
            if (a != tb[j].n_sk || b != tb[j].n_words) {
                delete ch;
                return set_err(GGCAT_B200_ERR_INVALID, "peer_exchange: slice %u of rank %u: unit counts sum to %llu super-k-mers / %llu words, header says %llu / %llu",
                               j, s2, (unsigned long long)a, (unsigned long long)b, (unsigned long long)tb[j].n_sk, (unsigned long long)tb[j].n_words);
            }
            c->chunks.push_back(ch);
        }
    }
    TRY(flush_scans());
    CU(cudaEventRecord(ps.ev_meta, ps.meta_stream));
    CU(cudaStreamWaitEvent(st, ps.ev_meta, 0));
    CU(cudaGetLastError());
    trace_host("peer_exchange: return");
    return 0;
}

int32_t ggcat_b200_peer_stats(ggcat_b200_ctx *c, uint64_t *bytes_sent, uint64_t *bytes_received) {
    TRY(check_ctx(c));
    if (bytes_sent) *bytes_sent = c->peer.last_sent;
    if (bytes_received) *bytes_received = c->peer.last_received;
    return 0;
}

void *ggcat_b200_stream(ggcat_b200_ctx *c) { return c ? (void *)c->stream : nullptr; }
int32_t ggcat_b200_synchronize(ggcat_b200_ctx *c) {
    TRY(check_ctx(c));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
int32_t ggcat_b200_set_timing(ggcat_b200_ctx *c, int32_t enabled) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    collect_timings(c);
    c->timing = enabled != 0;
    return 0;
}
int32_t ggcat_b200_kernel_times(ggcat_b200_ctx *c, const char **names, float *ms, uint32_t *launches, uint32_t cap, int32_t reset) {
    TRY(check_ctx(c));
    std::lock_guard<std::mutex> lock__(c->mu);
    collect_timings(c);
    for (uint32_t i = 0; i < (uint32_t)F_COUNT && i < cap; i++) {
        if (names) names[i] = kFamilyNames[i];
        if (ms) ms[i] = c->fam_ms[i];
        if (launches) launches[i] = c->fam_launches[i];
    }
    if (reset) { memset(c->fam_ms, 0, sizeof(c->fam_ms)); memset(c->fam_launches, 0, sizeof(c->fam_launches)); }
    return F_COUNT;
}

}  // extern "C"
