// wire.hpp -- the reference's bucket FILE format, so that a GPU phase 1 can feed an unmodified CPU phase 2 and vice versa
// (SURVEY 8(f)-3).  Host code: bytes in, bytes out.
//
// What phase 2 of the reference reads (crates/minimizer_bucketing/src/split_buckets.rs:46-131, decode_helper.rs:14-63) is
// the file the phase-1 compactor leaves behind (compactor.rs:370-420): one PLAIN ("lock free") bucket file per first-level
// bucket whose chunks are grouped by sub-bucket:
//   [BucketHeader 56 B: magic "PLAIN_INTR_BKT_M" | index_offset u64 LE | data_format_info[32]]      writers/mod.rs:15-24
//   [chunk of sub-bucket s0][chunk of sub-bucket s1] ...                                               raw records back to back
//   [index: bincode(BucketCheckpoints{index: Vec<CheckpointData{offset: u64, data: Option<Vec<u8>>}>})] writers/mod.rs:26-29,40-52
// data_format_info = bincode(MinimizerBucketMode::SingleGrouped) (one byte, 1; minimizer_bucketing/src/lib.rs:55-60);
// every chunk's checkpoint data = bincode(ReadsCheckpointData{target_subbucket: u16, sequences_count: usize})
// (creads_utils.rs:302-306); the first checkpoint sits right behind the header with no data and is skipped by the reader
// (split_buckets.rs:63-64).  bincode::config::standard() (lib.rs:8): integers as varints (< 251: one byte; 251 + u16,
// 252 + u32, 253 + u64 little endian), Option as a 0 / 1 byte, Vec as varint length + elements.
// A record of the SingleGrouped serializer <NoSecondBucket, NoMultiplicity, AssemblerMinimizerPosition, 2 flag bits>
// (creads_utils.rs:374-434, compressed_read.rs:435-467, varint.rs:26-56):
//   varint_flags(((len - k) << ceil_log2(k)) | minimizer_pos, flags)  |  ceil(len / 4) packed bytes (base i at bits 2(i%4))
// Coloured records carry an extra varint (colors/src/parsers/separate.rs:98-116) whose delta state depends on the writer's
// buffer boundaries: not produced here (uncoloured builds only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace ggb_wire {

constexpr char PLAIN_MAGIC[17] = "PLAIN_INTR_BKT_M";
constexpr size_t HEADER_BYTES = 16 + 8 + 32;
constexpr uint8_t MODE_SINGLE_GROUPED = 1;

inline void bincode_varint(std::vector<uint8_t> &o, uint64_t v) {
    if (v < 251) o.push_back((uint8_t)v);
    else if (v < (1ull << 16)) { o.push_back(251); o.push_back((uint8_t)v); o.push_back((uint8_t)(v >> 8)); }
    else if (v < (1ull << 32)) { o.push_back(252); for (int i = 0; i < 4; i++) o.push_back((uint8_t)(v >> (8 * i))); }
    else { o.push_back(253); for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i))); }
}
inline bool bincode_read_varint(const uint8_t *&p, const uint8_t *end, uint64_t *v) {
    if (p >= end) return false;
    const uint8_t b = *p++;
    int nb = b < 251 ? 0 : b == 251 ? 2 : b == 252 ? 4 : b == 253 ? 8 : -1;
    if (nb < 0 || p + nb > end) return false;
    if (nb == 0) { *v = b; return true; }
    uint64_t x = 0;
    for (int i = 0; i < nb; i++) x |= (uint64_t)p[i] << (8 * i);
    p += nb; *v = x;
    return true;
}
// crates/io/src/varint.rs:26-56 with FlagsCount = 2
inline void varint_flags2(std::vector<uint8_t> &o, uint64_t value, uint8_t flags) {
    const uint8_t first_max = 31;                        // 5 value bits, bit 5 = continuation, bits 6-7 = flags
    o.push_back((uint8_t)((flags << 6) | (value & first_max) | ((value > first_max) ? 32 : 0)));
    value >>= 5;
    while (value) { o.push_back((uint8_t)((value & 127) | ((value > 127) ? 128 : 0))); value >>= 7; }
}
inline bool read_varint_flags2(const uint8_t *&p, const uint8_t *end, uint64_t *value, uint8_t *flags) {
    if (p >= end) return false;
    const uint8_t f = *p++;
    *flags = f >> 6;
    uint64_t r = f & 31;
    unsigned off = 5;
    bool next = f & 32;
    while (next) {
        if (p >= end || off > 63) return false;
        const uint8_t b = *p++;
        next = b & 128;
        r |= (uint64_t)(b & 127) << off;
        off += 7;
    }
    *value = r;
    return true;
}
inline uint32_t ceil_log2(uint32_t k) { uint32_t l = 0; while ((1u << l) < k) l++; return l; }   // k.next_power_of_two().ilog2()

struct Record { uint32_t len, minimizer_pos, flags, sub_bucket; size_t byte_off; };   // packed bases at bytes[byte_off ..]

// One bucket file in memory.
struct Writer {
    std::vector<uint8_t> out;
    std::vector<uint8_t> index;      // encoded CheckpointData entries
    uint64_t n_checkpoints = 0;
    uint32_t k, klog;
    explicit Writer(uint32_t k_) : k(k_), klog(ceil_log2(k_)) {
        out.assign(HEADER_BYTES, 0);
        checkpoint(nullptr, 0);      // LockFreeBinaryWriter::new: first checkpoint behind the header, no data
    }
    void checkpoint(const uint8_t *data, size_t n) {
        bincode_varint(index, out.size());
        if (!data) index.push_back(0);
        else { index.push_back(1); bincode_varint(index, n); index.insert(index.end(), data, data + n); }
        n_checkpoints++;
    }
    void begin_sub_bucket(uint32_t sub, uint64_t n_sequences) {
        std::vector<uint8_t> cd;
        bincode_varint(cd, sub);
        bincode_varint(cd, n_sequences);
        checkpoint(cd.data(), cd.size());
    }
    void record(uint32_t len, uint32_t minimizer_pos, uint32_t flags, const uint8_t *packed) {
        varint_flags2(out, ((uint64_t)(len - k) << klog) | minimizer_pos, (uint8_t)flags);
        out.insert(out.end(), packed, packed + (len + 3) / 4);
    }
    void finish() {
        const uint64_t index_offset = out.size();
        std::vector<uint8_t> idx;
        bincode_varint(idx, n_checkpoints);
        idx.insert(idx.end(), index.begin(), index.end());
        out.insert(out.end(), idx.begin(), idx.end());
        memcpy(out.data(), PLAIN_MAGIC, 16);
        for (int i = 0; i < 8; i++) out[16 + i] = (uint8_t)(index_offset >> (8 * i));
        out[24] = MODE_SINGLE_GROUPED;
    }
};

// Parses a PLAIN SingleGrouped bucket file.  Returns an error string, empty on success.
inline std::string parse(const std::vector<uint8_t> &f, uint32_t k, std::vector<Record> &recs) {
    if (f.size() < HEADER_BYTES || memcmp(f.data(), PLAIN_MAGIC, 16) != 0) return "not a PLAIN_INTR_BKT_M bucket file (lz4 buckets are not read)";
    uint64_t index_offset = 0;
    for (int i = 0; i < 8; i++) index_offset |= (uint64_t)f[16 + i] << (8 * i);
    if (index_offset < HEADER_BYTES || index_offset > f.size()) return "index offset outside the file";
    if (f[24] != MODE_SINGLE_GROUPED) return "data format is not MinimizerBucketMode::SingleGrouped";
    const uint8_t *p = f.data() + index_offset, *end = f.data() + f.size();
    uint64_t n = 0;
    if (!bincode_read_varint(p, end, &n)) return "corrupt checkpoint index";
    struct Cp { uint64_t off; bool has; uint64_t sub, count; };
    std::vector<Cp> cps;
    for (uint64_t i = 0; i < n; i++) {
        Cp c{0, false, 0, 0};
        if (!bincode_read_varint(p, end, &c.off) || p >= end) return "corrupt checkpoint index";
        const uint8_t opt = *p++;
        if (opt == 1) {
            uint64_t len = 0;
            if (!bincode_read_varint(p, end, &len) || p + len > end) return "corrupt checkpoint data";
            const uint8_t *q = p;
            if (!bincode_read_varint(q, p + len, &c.sub) || !bincode_read_varint(q, p + len, &c.count)) return "corrupt ReadsCheckpointData";
            c.has = true;
            p += len;
        } else if (opt != 0) return "corrupt checkpoint option";
        cps.push_back(c);
    }
    const uint32_t klog = ceil_log2(k);
    for (size_t i = 0; i < cps.size(); i++) {
        if (!cps[i].has) continue;       // the empty first chunk
        const uint64_t a = cps[i].off, b = i + 1 < cps.size() ? cps[i + 1].off : index_offset;
        if (a > b || b > index_offset) return "checkpoint offsets out of order";
        const uint8_t *q = f.data() + a, *qe = f.data() + b;
        uint64_t seen = 0;
        while (q < qe) {
            uint64_t v = 0; uint8_t fl = 0;
            if (!read_varint_flags2(q, qe, &v, &fl)) return "corrupt record header";
            Record r;
            r.len = (uint32_t)(v >> klog) + k; r.minimizer_pos = (uint32_t)(v & ((1u << klog) - 1)); r.flags = fl;
            r.sub_bucket = (uint32_t)cps[i].sub; r.byte_off = (size_t)(q - f.data());
            const size_t nb = (r.len + 3) / 4;
            if (q + nb > qe) return "record runs past its chunk";
            q += nb;
            recs.push_back(r);
            seen++;
        }
        if (seen != cps[i].count) return "sequences_count of a chunk does not match its records";
    }
    return "";
}

}  // namespace ggb_wire
