#!/usr/bin/env python
"""bench.py -- build Gbases/s (minimizer bucketing + k-mer merge) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port, all host cores)

A "step" is one full pass of the hot path (phase 1 + phase 2) over one batch of synthetic reads.
Workload at N=1: BASELINE configs[1] (C2): synthetic 5 Mbp genome, 1 M x 150 bp reads (30x), 1 % errors, k=31 -s 2,
seq-hash, buckets 512(+1) x 64 as the reference would choose (crates/io/src/lib.rs:67-140).
At N>1 every rank holds its own 1 M-read slice of a 5N Mbp genome (weak scaling); buckets are owned by
contiguous ranges, super-k-mers are routed with one all-to-all (NCCL) and each owner merges locally.

Printed JSON line (rank 0): see DESIGN.md "Measurement".
  value    device-timed (CUDA events on the library's stream), inputs resident in HBM
  e2e      same metric through the C ABI with HOST buffers (H2D of reads, D2H of the table inside the timed region)
  roofline dominant kernel family (k_merge_hash<smem> on C2) against the measured HBM peak; `achieved` uses the
           algorithmic-bytes model of SURVEY 8(d), `compulsory_bytes_per_launch` what the kernel really has to move
  exchange N > 1: bytes pushed over NVLink per GPU, time of the push + flag kernels, fraction of 770 GB/s
  cpu_baseline  N = 1: the oracle's OpenMP port, repeated passes over the whole batch for >= 10 s (kind "port")

--workload c4 benches a slice of BASELINE configs[3] (error-free reads generated on the device, 1024 x 64 buckets, big
merge units; --reads-per-gpu 77500000 at --gpus 8 is the full 93 Gbases set).  The bench line stays C2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K, M, S = 31, 12, 2
READ_LEN = 150
READS_PER_GPU = 1_000_000
GENOME_PER_GPU = 5_000_000
ERR = 0.01


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_reads(rank: int, world: int, n_reads: int):
    from ggcat_b200 import synth

    g = synth.genome_codes(0xC2, GENOME_PER_GPU * world)
    r = synth.simulate_reads(g, n_reads, READ_LEN, ERR, 0xC2 + 1, first_read=rank * n_reads)
    return synth.reads_to_ascii_batch(r)


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path: Rust cannot be built here, so this is
    the oracle port (oracle/ggcat_oracle.c, OpenMP over all host cores) on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    import ggcat_b200  # noqa: F401  (only for bucket_counts parity with our arm; no GPU work)
    from ggcat_b200 import synth

    sample_reads = args.sample_reads   # default: the whole per-GPU C2 batch, ~1 s of CPU work per step on 16 cores
    cores = os.cpu_count() or 1
    data, offsets = make_reads(0, max(args.gpus, 1), sample_reads)
    b1, b2 = O.bucket_counts(int(READS_PER_GPU * max(args.gpus, 1) * (READ_LEN + 15)))  # same bucket counts as the full workload
    reads = O.Reads(data, offsets)
    times = []
    st = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        st = O.pipeline(reads, K, M, b1, b2, S, n_threads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    bases = int(data.size)
    ms = 1e3 * float(np.mean(times))
    val = bases / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": "build Gbases/s (bucketing+k-mer merge)", "value": val, "unit": "Gbases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"C2: {sample_reads} x {READ_LEN} bp reads of the 5 Mbp/30x/1% set per step, k={K} m={M} -s {S}, "
                               f"buckets {1 << b1}(+1) x {1 << b2}"},
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": int(st.threads), "kind": "port",
                         "sample": f"{sample_reads} reads ({bases} bases); phase1 {st.t_bucketing:.3f}s phase2 {st.t_merge:.3f}s"},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the ggcat_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        try:   # pin this rank to the CPUs / NUMA node of its GPU before any pinned host memory is allocated (e2e copies
               # of 8 ranks share the host's memory system); N=1 keeps every core for the CPU baseline
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    if world > 1:
        # the exchange is one large point-to-point all-to-all: let NCCL spread it over more channels
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import ggcat_b200 as G
    from ggcat_b200 import dist as gdist

    wl = args.workload
    n_reads = args.reads_per_gpu if args.reads_per_gpu else (READS_PER_GPU if wl == "c2" else 20_000_000)
    dev = torch.device("cuda", local_rank)
    if wl == "c2":
        data, offsets = make_reads(rank, world, n_reads)
        n_bases = int(data.size)
        h_data = torch.from_numpy(data).pin_memory()
        h_off = torch.from_numpy(offsets.view(np.int64)).pin_memory()
        d_data = h_data.cuda(non_blocking=True)
        d_off = h_off.cuda(non_blocking=True)
        genome_len, err, seed_note = GENOME_PER_GPU * world, ERR, "1% errors"
        reads_per_push = n_reads
    else:
        # C4 shape (BASELINE configs[3]): error-free 150 bp reads at 30x of one genome shared by all ranks, generated
        # on the device; rank r holds reads [r*R, (r+1)*R).  Full C4 is 77.5 M reads per GPU at N=8.
        from ggcat_b200 import synth
        genome_len, err, seed_note = 5 * n_reads * world, 0.0, "error-free"
        genome = synth.genome_codes_torch(0xC4, genome_len, dev)
        d_data = synth.simulate_reads_torch(genome, n_reads, READ_LEN, 0.0, 0xC4 + 1, first_read=rank * n_reads)
        del genome
        torch.cuda.empty_cache()
        n_bases = int(d_data.numel())
        d_off = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * READ_LEN
        h_data = h_off = None
        if not args.no_e2e:
            h_data = torch.empty(n_bases, dtype=torch.uint8).pin_memory()
            h_data.copy_(d_data)
            h_off = torch.empty(n_reads + 1, dtype=torch.int64).pin_memory()
            h_off.copy_(d_off)
        reads_per_push = args.reads_per_push          # device pushes of 600 Mbases (one bucket chunk each)
    # bucket counts as the reference derives them from the size of the WHOLE input (all ranks' FASTA bytes,
    # crates/io/src/lib.rs:67-140)
    b1, b2 = G.bucket_counts(int(n_reads * world * (READ_LEN + 15)))
    if args.b1 is not None:
        b1 = args.b1
    nb = (1 << b1) + 1
    ctx = G.GGCATB200(G.Params(k=K, m=M, min_multiplicity=S, buckets_count_log=b1, second_buckets_count_log=b2,
                               device=local_rank))
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    # per-push device views (offsets rebased to the push)
    pushes = []
    for r0 in range(0, n_reads, reads_per_push):
        r1 = min(n_reads, r0 + reads_per_push)
        off = (d_off[r0:r1 + 1] - r0 * READ_LEN).contiguous() if r0 else d_off[:r1 + 1]
        pushes.append((d_data.data_ptr() + r0 * READ_LEN, off, r1 - r0, (r1 - r0) * READ_LEN))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    torch.cuda.synchronize()
    owner = gdist.OwnerMap(b1, b2, world)
    transport = "none"
    if world > 1:
        transport = os.environ.get("GGCAT_B200_EXCHANGE", "peer")
        if transport == "peer":
            # receive arena: descriptors (16 B / super-k-mer) + payload ~ 2.4 B per input base; 2.5x headroom on small
            # inputs, 1.5x on large ones (the arena is HBM that the merge cannot use)
            gdist.peer_setup(ctx, rank, world, arena_bytes=max(int((6 if n_bases < (1 << 31) else 3.6) * n_bases), 64 << 20))

    last_stats = [None]

    def step_device():
        ctx.reset()
        for ptr, off, nr, nbytes in pushes:
            ctx.push_reads_device(ptr, off.data_ptr(), nr, nbytes)
        last_stats[0] = ctx.finish_bucketing()   # this rank's own super-k-mers (before the exchange adds imported chunks)
        if world > 1:
            gdist.exchange_and_import(ctx, owner, rank, world, ext)
        fb, cnt = owner.bucket_range(rank)
        return ctx.merge_bucket_range_device(fb, cnt)

    def step_host():
        ctx.reset()
        ctx.push_reads_ptr(h_data.data_ptr(), h_off.data_ptr(), n_reads)   # the library splits it into double-buffered H2D batches
        ctx.finish_bucketing()
        if world > 1:
            gdist.exchange_and_import(ctx, owner, rank, world, ext)
        fb, cnt = owner.bucket_range(rank)
        return ctx.merge_bucket_range(fb, cnt, copy=False)  # what the C ABI hands a host: pinned table, no extra copy

    def l2_flush():
        with torch.cuda.stream(ext):
            flush.fill_(rank & 0xFF)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
        l2_flush()
    # ---- timed region: K steps, device time on the library stream, L2 flushed between steps (not timed)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ctx.kernel_times(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    res = None
    for i in range(args.steps):
        l2_flush()
        barrier()
        ev[i][0].record(ext)
        res = step_device()
        ev[i][1].record(ext)
    barrier()
    clocks = sampler.stop()
    launches = sum(v[1] for v in ctx.kernel_times(reset=True).values())
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = (n_bases * world) / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel timing pass (events around every launch; separate from the timed region)
    ctx.set_timing(True)
    n_prof = 3
    ctx.kernel_times(reset=True)
    for _ in range(n_prof):
        l2_flush()
        step_device()
    kt = ctx.kernel_times(reset=True)
    ctx.set_timing(False)
    st = last_stats[0]
    n_entries, unique, total_kmers = res
    exchange = None
    if world > 1 and transport == "peer":
        sent, recvd = ctx.peer_stats()
        ex_ms = kt.get("k_peer_push+k_peer_sync", (0.0, 0))[0] / n_prof
        ex = torch.tensor([float(sent), float(recvd), ex_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(ex, op=dist.ReduceOp.MAX)
        sent_max, recv_max, ex_ms = (float(x) for x in ex.tolist())
        # NVLink 5: 900 GB/s per direction nominal, ~770 GB/s achievable; the push also waits for the slowest peer
        exchange = {"bytes_sent_per_gpu": int(sent_max), "bytes_received_per_gpu": int(recv_max), "ms_per_step": ex_ms,
                    "achieved_GBps_per_gpu": sent_max / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else None,
                    "frac_of_770_GBps": sent_max / (ex_ms * 1e-3) / 1e9 / 770.0 if ex_ms > 0 else None}

    # ---- e2e through the C ABI with host buffers
    e2e = None
    if h_data is not None:
        for _ in range(2):
            step_host().release()
        e2e_times = []
        d2h = 0
        for _ in range(args.steps):
            l2_flush()
            barrier()
            t0 = time.perf_counter()
            tab = step_host()
            torch.cuda.synchronize()
            e2e_times.append(time.perf_counter() - t0)
            d2h = int(tab.keys_lo.nbytes + tab.count_flags.nbytes + tab.unit_offsets.nbytes)
            tab.release()
        e2e_ms = float(np.mean(e2e_times)) * 1e3
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": (n_bases * world) / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(h_data.numel() + h_off.numel() * 8), "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms}

    # ---- roofline of the dominant kernel (largest device time in the per-kernel pass)
    peak, peak_kind = measured_peak()
    total_kernel_ms = max(sum(v[0] for v in kt.values()), 1e-9)
    dom = max(kt, key=lambda name: kt[name][0])
    dom_ms, dom_launches = kt[dom]
    per_launch_ms = dom_ms / max(dom_launches, 1)
    launches_per_step = max(dom_launches / n_prof, 1)
    B_s = st.payload_words * 4 + st.n_superkmers * 16       # super-k-mer payload + descriptors
    N_k = st.n_kmers
    if dom.startswith("k_merge"):
        # SURVEY 8(d) merge model with this build's record size W+P = 8 B and R = 8 LSD passes:
        #   B_s + N_k*8*(1 expand write + 2R sort + 1 reduce read) + S*12
        model_bytes = B_s + N_k * 8 * (1 + 2 * 8 + 1) + n_entries * 12
        compulsory = B_s + n_entries * 12                   # what the kernel must move: super-k-mers in, table out
        model = "SURVEY 8(d) DRAM-LSD-equivalent bytes (W+P=8, R=8); the kernel counts in shared memory"
    else:
        # bucketing kernels: packed bases + 2 bitmaps in, entries out (SURVEY 8(d) bucketing model share)
        model_bytes = compulsory = n_bases // 4 + 2 * (n_bases // 8) + st.n_superkmers * 8
        model = "compulsory bytes: packed bases + bad/brk bitmaps in, split entries out"
    achieved = model_bytes / launches_per_step / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_kind": peak_kind, "model": model,
                "algorithmic_bytes_per_launch": model_bytes / launches_per_step,
                "compulsory_bytes_per_launch": compulsory / launches_per_step, "ms_per_launch": per_launch_ms,
                "kernel_share_of_step": dom_ms / total_kernel_ms}
    tr = ROOT / "profiles" / "traffic.json"
    if tr.exists():
        try:
            roofline["traffic"] = json.loads(tr.read_text()).get(dom)
        except Exception:
            pass

    line = {
        "metric": "build Gbases/s (bucketing+k-mer merge)", "value": value, "unit": "Gbases/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{wl.upper()} per GPU: {n_reads} x {READ_LEN} bp reads (30x of {genome_len} bp genome, {seed_note}), "
                               f"k={K} m={M} -s {S} seq-hash, buckets {1 << b1}(+1) x {1 << b2}",
                   "l2": "flushed (256 MB write) between timed steps", "reads_per_gpu": n_reads,
                   "parallelism": f"bucket-owner x{world}" if world > 1 else "single GPU",
                   "exchange": {"peer": "k_peer_push over NVLink peer memory (CUDA IPC)", "nccl": "NCCL all_to_all_single",
                                "none": "none"}[transport]},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels_ms_per_step": {k: round(v[0] / n_prof, 4) for k, v in kt.items()},
        "exchange": exchange,
        "counts": {"bases_per_gpu": n_bases, "superkmers": int(st.n_superkmers), "kmer_records": int(st.n_kmers),
                   "unique": int(unique), "kept": int(n_entries)},
    }
    if rank == 0 and world == 1 and wl == "c2" and not args.no_cpu_baseline:
        from oracle import oracle as O

        sr = args.sample_reads
        cdata, coff = make_reads(0, 1, sr)
        creads = O.Reads(cdata, coff)
        O.pipeline(creads, K, M, b1, b2, S, n_threads=os.cpu_count() or 1)   # warm-up (page faults, thread pool)
        passes, t_cpu, pst = 0, 0.0, None
        while t_cpu < args.cpu_seconds and passes < 64:
            t0 = time.perf_counter()
            pst = O.pipeline(creads, K, M, b1, b2, S, n_threads=os.cpu_count() or 1)
            t_cpu += time.perf_counter() - t0
            passes += 1
        line["cpu_baseline"] = {"value": cdata.size * passes / t_cpu / 1e9, "unit": "Gbases/s", "cores": int(pst.threads), "kind": "port",
                                "sample": f"{passes} passes over {sr} reads ({cdata.size} bases each, {t_cpu:.1f} s of CPU work) of the same "
                                          f"workload, oracle C port with OpenMP; last pass phase1 {pst.t_bucketing:.3f}s phase2 {pst.t_merge:.3f}s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=0, help="default: 1 M (c2) / 20 M (c4)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2 = BASELINE configs[1] (the bench line); c4 = a slice of configs[3] (human-scale shape, big merge units)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pass (large c4 slices)")
    ap.add_argument("--reads-per-push", type=int, default=4_000_000, help="c4: reads per device push (one bucket chunk each)")
    ap.add_argument("--sample-reads", type=int, default=READS_PER_GPU, help="reads per CPU pass (default: the whole C2 batch)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline: repeat passes until this much CPU wall time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--b1", type=int, default=None, help="experiment: override buckets_count_log")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
