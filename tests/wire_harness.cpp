// Test infrastructure: C entry points over ggcat_b200/csrc/wire.hpp (the product's bucket-file writer / parser is host-only
// C++ inside the CUDA library) so that the "not gpu" suite can exercise it without a device.  Built by tests/test_wire_cpu.py.
#include "../ggcat_b200/csrc/wire.hpp"

extern "C" {

// Records must arrive grouped by sub-bucket (ascending), as the library writes them.  Returns 0, or -1 if `cap` is too small.
int wire_build(uint32_t k, uint32_t n, const uint32_t *len, const uint32_t *mpos, const uint32_t *flags, const uint32_t *sub,
               const uint8_t *packed, const uint64_t *packed_off, uint8_t *out, uint64_t cap, uint64_t *out_size) {
    ggb_wire::Writer w(k);
    for (uint32_t i = 0; i < n;) {
        uint32_t j = i;
        while (j < n && sub[j] == sub[i]) j++;
        w.begin_sub_bucket(sub[i], j - i);
        for (uint32_t r = i; r < j; r++) w.record(len[r], mpos[r], flags[r], packed + packed_off[r]);
        i = j;
    }
    w.finish();
    *out_size = w.out.size();
    if (w.out.size() > cap) return -1;
    memcpy(out, w.out.data(), w.out.size());
    return 0;
}

// Returns 0 and the records, or -2 with the parser's message in err.
int wire_parse(const uint8_t *f, uint64_t n, uint32_t k, uint32_t *len, uint32_t *mpos, uint32_t *flags, uint32_t *sub,
               uint64_t *byte_off, uint64_t cap, uint64_t *n_out, char *err, uint32_t errcap) {
    std::vector<uint8_t> buf(f, f + n);
    std::vector<ggb_wire::Record> recs;
    const std::string e = ggb_wire::parse(buf, k, recs);
    if (!e.empty()) { snprintf(err, errcap, "%s", e.c_str()); return -2; }
    *n_out = recs.size();
    if (recs.size() > cap) return -1;
    for (size_t i = 0; i < recs.size(); i++) {
        len[i] = recs[i].len; mpos[i] = recs[i].minimizer_pos; flags[i] = recs[i].flags; sub[i] = recs[i].sub_bucket;
        byte_off[i] = recs[i].byte_off;
    }
    return 0;
}
}
