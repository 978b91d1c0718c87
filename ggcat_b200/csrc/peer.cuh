// peer.cuh -- the one exchange step of the sharded path, over NVLink peer memory.
//
// Replaces the reference's "write bucket files / read bucket files" shuffle between phase 1 and phase 2
//   crates/minimizer_bucketing/src/lib.rs:340-351   (per-bucket writers, one file per bucket)
//   crates/kmers_transform/src/lib.rs:294-371       (readers that feed the merge workers)
// for a build sharded over the GPUs of one NVSwitch box: rank g owns a contiguous range of
// first-level buckets; after phase 1 every rank PUSHES the slice of its unit-sorted bucket chunks that belongs to
// owner d straight into d's receive arena with 16-byte stores over NVLink (the arena is cudaMalloc memory shared
// through CUDA IPC), then raises a flag in d's arena.  No NCCL, no packing pass (chunks are unit-sorted and owners
// hold contiguous unit ranges, so a slice is a plain sub-array), no host round trip on the sender.
//
// Arena of rank X:   [PeerHdr 4 KiB] [region 0] ... [region world-1]      (region s is written by rank s only)
// Region:            [RegionHdr 64 B] [PeerSlice table, PEER_MAX_SLICES entries] [S meta slots: per-unit cnt | words | kmers]
//                    [S scratch slots: the receiver's descriptor / word offset scans] [descriptors / payload of the slices ...]
//                    (S = slices a build may route, fixed at peer_init; slot sizes follow from the owner's unit count)
// A rank pushes EAGERLY: the slices of a bucket chunk leave for their owners as soon as the chunk's k_scatter is done, on a
// side stream, while the next batch is bucketed; the small per-unit counts (meta) and the slice-table entry of a chunk
// travel on a second side stream right after its k_emit, so that every owner's HOST knows the sizes of what it will merge
// (unit classification, work lists) before the bulk data has arrived.  Flow control, all stream-ordered device code
// (k_peer_sync), epochs count builds:
//   released[d] on rank X = last epoch whose data rank d has finished merging     -> X may overwrite its region on d
//   meta[s]     on rank X = last epoch whose headers / counts from rank s are complete -> X's host may plan its merge
//   ready[s]    on rank X = last epoch whose bulk push from rank s is complete        -> X may merge
#pragma once
#include "device_utils.cuh"

namespace ggb {

constexpr int PEER_MAX_WORLD = 64;
constexpr int PEER_MAX_SLICES = 256;                // local chunks one build can route (64 B of slice table each)
constexpr uint64_t PEER_HDR_BYTES = 4096;
constexpr uint64_t PEER_TABLE_OFF = 64;
constexpr uint64_t PEER_META_OFF = PEER_TABLE_OFF + (uint64_t)PEER_MAX_SLICES * 64;

struct PeerHdr {
    uint32_t ready[PEER_MAX_WORLD];
    uint32_t released[PEER_MAX_WORLD];
    uint32_t meta[PEER_MAX_WORLD];
};
enum { PEER_FLAG_RELEASED = 0, PEER_FLAG_READY = 1, PEER_FLAG_META = 2 };

struct RegionHdr {           // 64 bytes, written by the sender
    uint32_t n_slices;
    uint32_t overflow;       // 1 = the sender's slices did not fit the region: nothing was delivered
    uint32_t n_units;        // units of the destination (sanity check)
    uint32_t epoch;
    uint32_t pad[12];
};

struct PeerSlice {           // 64 bytes: one local chunk's slice for this destination
    uint64_t n_sk, n_words, word_bias;
    uint64_t desc_off, pay_off;      // byte offsets from the region base
    uint64_t pad[3];
};

struct PeerJob { const uint8_t *src; uint8_t *dst; uint64_t bytes; uint32_t slot, pad; };   // 4-byte aligned, bytes % 4 == 0;
                                                                                          // slot = destination index among the peers

struct PeerHdrPtrs { PeerHdr *h[PEER_MAX_WORLD]; };

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// k_peer_push: bulk copies into peer memory.  The grid is split evenly over the destinations (CTA b serves destination
// slot b % n_slots), so every owner receives from all its sources at 1/(W-1) of their rate all the time: no rank is
// the target of everybody at once (the incast that a destination-after-destination order produces: measured 194 GB/s
// per GPU at N=8 against 530 GB/s at N=2).  Jobs whose source and destination agree modulo 16 move as uint4
// (4 independent 16-byte loads in flight per thread, 512 contiguous bytes per warp store); the rest as u32.
__global__ void __launch_bounds__(256) k_peer_push(const PeerJob *__restrict__ jobs, uint32_t n_jobs, uint32_t n_slots) {
    const uint32_t my_slot = blockIdx.x % n_slots, my_idx = blockIdx.x / n_slots;
    const uint32_t ctas = (gridDim.x - my_slot + n_slots - 1) / n_slots;     // CTAs serving this slot
    const uint64_t gtid = (uint64_t)my_idx * blockDim.x + threadIdx.x;
    const uint64_t gsz = (uint64_t)ctas * blockDim.x;
    for (uint32_t j = 0; j < n_jobs; j++) {
        const PeerJob job = jobs[j];
        if (job.slot != my_slot || job.bytes == 0) continue;
        const uintptr_t sa = (uintptr_t)job.src, da = (uintptr_t)job.dst;
        if (((sa ^ da) & 15u) == 0) {
            uint64_t head = (16u - (sa & 15u)) & 15u;
            if (head > job.bytes) head = job.bytes;
            const uint64_t nvec = (job.bytes - head) >> 4;
            const uint64_t tail0 = head + (nvec << 4);
            const uint4 *s4 = reinterpret_cast<const uint4 *>(job.src + head);
            uint4 *d4 = reinterpret_cast<uint4 *>(job.dst + head);
            uint64_t i = gtid;
            for (; i + 3 * gsz < nvec; i += 4 * gsz) {
                const uint4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + gsz), c = __ldcs(s4 + i + 2 * gsz), d = __ldcs(s4 + i + 3 * gsz);
                d4[i] = a; d4[i + gsz] = b; d4[i + 2 * gsz] = c; d4[i + 3 * gsz] = d;
            }
            for (; i < nvec; i += gsz) d4[i] = __ldcs(s4 + i);
            if (my_idx == 0) {
                const uint32_t *s1 = reinterpret_cast<const uint32_t *>(job.src);
                uint32_t *d1 = reinterpret_cast<uint32_t *>(job.dst);
                for (uint64_t q = threadIdx.x; q < (head >> 2); q += blockDim.x) d1[q] = s1[q];
                for (uint64_t q = (tail0 >> 2) + threadIdx.x; q < (job.bytes >> 2); q += blockDim.x) d1[q] = s1[q];
            }
        } else {
            const uint32_t *s1 = reinterpret_cast<const uint32_t *>(job.src);
            uint32_t *d1 = reinterpret_cast<uint32_t *>(job.dst);
            for (uint64_t q = gtid; q < (job.bytes >> 2); q += gsz) d1[q] = s1[q];
        }
    }
    __threadfence_system();
}

// k_peer_sync: thread t talks to rank t.  signal: stores `value` into this rank's slot of rank t's flag array `which`;
// wait: spins until rank t's slot of the LOCAL flag array reaches `value`.  Bounded spin: on timeout *err is set and the
// host reports GGCAT_B200_ERR_STATE instead of hanging.
__global__ void k_peer_sync(PeerHdrPtrs peers, uint32_t me, uint32_t world, uint32_t which, uint32_t value, uint32_t do_signal,
                            uint32_t do_wait, uint32_t *err, unsigned long long timeout_ns) {
    const uint32_t t = threadIdx.x;
    if (t >= world || t == me) return;
    __threadfence_system();
    if (do_signal) {
        PeerHdr *h = peers.h[t];
        uint32_t *theirs = which == PEER_FLAG_READY ? &h->ready[me] : which == PEER_FLAG_RELEASED ? &h->released[me] : &h->meta[me];
        st_release_sys(theirs, value);
    }
    if (do_wait) {
        const PeerHdr *h = peers.h[me];
        const uint32_t *mine = which == PEER_FLAG_READY ? &h->ready[t] : which == PEER_FLAG_RELEASED ? &h->released[t] : &h->meta[t];
        const unsigned long long t0 = global_timer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - value) < 0) {
            __nanosleep(200);
            if (global_timer_ns() - t0 > timeout_ns) { atomicExch(err, 1u + t); return; }
        }
    }
}

}  // namespace ggb
