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
  roofline dominant kernel family against the measured HBM peak: `achieved` = the bytes the kernel has to move (inputs once,
           outputs once) / its CUDA-event time, `traffic` = dram__bytes_read+write measured by an ncu pass of this run
           (fallback: profiles/r02_traffic.json, labelled), `frac` = achieved / peak; `limiter` says what really bounds it
           (issue slots), `step` / `per_kernel` give the same for the whole step and every family;
           `survey_model_equiv_frac` is the SURVEY 8(d) DRAM-LSD pipeline (195 B/base, W+P = 12) at this step time
  parity   N > 1: sampled owned units of every rank (device path and host path) against the oracle on the union of all reads
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
# parameters of the benchable BASELINE configs; wp = key + payload bytes, R = LSD passes of the SURVEY 8(d) byte model
WORKLOADS = {
    "c2": dict(k=31, m=12, s=2, hash_type=1, colors=False, wp=12, R=8),
    "c4": dict(k=31, m=12, s=2, hash_type=1, colors=False, wp=12, R=8),
    "c5": dict(k=63, m=14, s=2, hash_type=4, colors=False, wp=20, R=16),   # -w rabin-karp128
    "c3": dict(k=31, m=12, s=1, hash_type=1, colors=True, wp=16, R=8),     # -c, colour = genome index
}
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


def workload_string(wl: str, n_reads: int, world: int, b1: int, b2: int) -> str:
    """config.workload, shared by both arms (the driver compares the two strings)."""
    if wl == "c2":
        return (f"C2 per GPU: {n_reads} x {READ_LEN} bp reads (30x of {GENOME_PER_GPU * world} bp genome, 1% errors), "
                f"k={K} m={M} -s {S} seq-hash, buckets {1 << b1}(+1) x {1 << b2}")
    if wl == "c5":
        return (f"C5 slice per GPU: {n_reads} x {READ_LEN} bp reads (30x of {5 * n_reads * world} bp genome, error-free), "
                f"k=63 m=14 -s 2 rabin-karp128, buckets {1 << b1}(+1) x {1 << b2}")
    if wl == "c3":
        return (f"C3: {n_reads} genomes x 5000000 bp (20 shared 200 kbp segments mutated at 0.1%), colour = genome, "
                f"k=31 m=12 -s 1 -c, buckets {1 << b1}(+1) x {1 << b2}")
    return (f"C4 per GPU: {n_reads} x {READ_LEN} bp reads (30x of {5 * n_reads * world} bp genome, error-free), "
            f"k={K} m={M} -s {S} seq-hash, buckets {1 << b1}(+1) x {1 << b2}")


FAMILY_OF = [  # kernel-name regex -> family name of ggcat_b200_kernel_times
    (r"k_pack|k_mark", "k_pack+k_mark"), (r"k_windows", "k_windows"), (r"k_emit", "k_emit"),
    (r"k_scatter|k_init_cursors", "k_scatter"), (r"k_exclusive_scan_u32", "k_exclusive_scan_u32"),
    (r"k_merge_tier", "k_merge_hash<smem>"), (r"k_merge_hash<", "k_merge_hash<global>"),
    (r"k_scan_unit_slots|k_finish_small|k_finish_units", "k_gather_units"), (r"k_partition_units", "k_partition_units"),
    (r"k_merge_parts", "k_merge_hash<partitions>"),
]


def traffic_pass(timeout_s: float = 150.0):
    """DRAM traffic per kernel family of ONE warm C2 step, measured now: ncu (dram__bytes_read/write.sum) over
    profiles/prof_driver.py, which brackets its last step with cudaProfilerStart/Stop.  Returns (dict, source)."""
    import csv
    import re
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    try:
        out = subprocess.run([ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
                              "--profile-from-start", "off", "--csv", sys.executable, str(ROOT / "profiles" / "prof_driver.py"), "3"],
                             capture_output=True, text=True, timeout=timeout_s, cwd=str(ROOT))
        rows = list(csv.reader([l for l in out.stdout.splitlines() if l.startswith('"')]))
        hdr = rows[0]
        ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        fam = {}
        for r in rows[1:]:
            if not r[im].startswith("dram__bytes"):
                continue
            name = next((f for pat, f in FAMILY_OF if re.search(pat, r[ik])), None)
            if name is None:
                continue
            fam[name] = fam.get(name, 0.0) + float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        if fam:
            return {k: int(v) for k, v in fam.items()}, "ncu pass in this run (dram__bytes_read.sum + dram__bytes_write.sum, one warm step)"
    except Exception:
        pass
    tr = ROOT / "profiles" / "r02_traffic.json"
    if tr.exists():
        try:
            d = json.loads(tr.read_text())
            return {k: v for k, v in d.items() if not k.startswith("_")}, "profiles/r02_traffic.json (committed ncu pass of the same command; ncu unavailable in this run)"
        except Exception:
            pass
    return {}, "unavailable"


_GENOME_CACHE = {}


def make_reads(rank: int, world: int, n_reads: int):
    from ggcat_b200 import synth

    glen = GENOME_PER_GPU * world
    if glen not in _GENOME_CACHE:          # the N-rank reference arm and the N > 1 parity check ask for every rank's slice
        _GENOME_CACHE.clear()
        _GENOME_CACHE[glen] = synth.genome_codes(0xC2, glen)
    r = synth.simulate_reads(_GENOME_CACHE[glen], n_reads, READ_LEN, ERR, 0xC2 + 1, first_read=rank * n_reads)
    return synth.reads_to_ascii_batch(r)


def make_reads_all(world: int, n_reads: int):
    """The slices of all ranks (reference arm, N > 1 parity check), generated in threads (numpy releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    first = make_reads(0, world, n_reads)          # fills the genome cache
    if world == 1:
        return [first]
    with ThreadPoolExecutor(min(world - 1, os.cpu_count() or 1)) as ex:
        return [first] + list(ex.map(lambda r: make_reads(r, world, n_reads), range(1, world)))


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

    world = max(args.gpus, 1)
    per_gpu = args.reads_per_gpu if args.reads_per_gpu else READS_PER_GPU
    # the same N-GPU workload as our arm: the slices of all N ranks, one pass over all of them per step (about 1 s of
    # CPU work per GPU slice on 16 cores); --sample-reads bounds it (tests)
    if args.sample_reads:
        parts = [make_reads(0, world, args.sample_reads)]
        sample_note = f"{args.sample_reads} reads of rank 0's slice"
    else:
        parts = make_reads_all(world, per_gpu)
        sample_note = f"all {world} x {per_gpu} reads of the workload"
    cores = os.cpu_count() or 1
    data = np.concatenate([d for d, _ in parts])
    offsets = np.arange(data.size // READ_LEN + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    b1, b2 = O.bucket_counts(int(per_gpu * world * (READ_LEN + 15)))  # same bucket counts as the full workload
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
        "config": {"workload": workload_string("c2", per_gpu, world, b1, b2)},
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": int(st.threads), "kind": "port",
                         "sample": f"{sample_note} ({bases} bases per step); phase1 {st.t_bucketing:.3f}s phase2 {st.t_merge:.3f}s"},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def multi_gpu_parity(args, ctx, owner, rank, world, b1, b2, n_reads, step_device, step_host):
    """Every rank hands `parity_units` sampled units of its owner range (tables of the device path and of the host
    path) to rank 0, which runs the oracle on the UNION of all ranks' reads and compares bit for bit."""
    import torch.distributed as dist

    fb, cnt = owner.bucket_range(rank)
    u_lo, u_hi = fb << b2, (fb + cnt) << b2
    rng = np.random.default_rng(1234 + rank)
    units = np.sort(rng.choice(np.arange(u_lo, u_hi), size=min(args.parity_units, u_hi - u_lo), replace=False))
    payload = {"rank": rank, "units": units.tolist(), "tables": {}}
    step_device()
    tab = ctx.read_device_table([int(u) for u in units])      # only the sampled units leave the GPU
    payload["tables"]["device"] = [(np.array(tab.keys_lo[tab.unit_slice_at(i)]), np.array(tab.count_flags[tab.unit_slice_at(i)]))
                                   for i in range(len(units))]
    if step_host is not None:
        tab = step_host()
        payload["tables"]["host"] = [(np.array(tab.keys_lo[tab.unit_slice(int(u))]), np.array(tab.count_flags[tab.unit_slice(int(u))]))
                                     for u in units]
        tab.release()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    return parity_against_oracle(gathered, world, n_reads, b1, b2)


def parity_against_oracle(gathered, world: int, n_reads: int, b1: int, b2: int):
    """Rank 0, host only: the oracle on the UNION of all ranks' reads against the sampled unit tables every rank sent
    (payload = {"rank", "units", "tables": {path: [(keys, count_flags) per unit]}})."""
    from oracle import oracle as O

    parts = make_reads_all(world, n_reads)
    data = np.concatenate([d for d, _ in parts])
    offsets = np.arange(data.size // READ_LEN + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    reads = O.Reads(data, offsets)
    sk, _ = O.bucketing_parallel(reads, K, M, b1, b2)
    # merge_unit selects a unit's super-k-mers with a mask over the array it is given: hand it only the super-k-mers of the
    # sampled units (one pass over the 10^8 rows instead of one per unit)
    sk = sampled_superkmers(sk, [u for pl in gathered for u in pl["units"]], b2)
    checked, entries, bad = 0, 0, []
    for pl in gathered:
        for i, u in enumerate(pl["units"]):
            ref, _, _ = O.merge_unit(reads, sk, int(u) >> b2, int(u) & ((1 << b2) - 1), K, S)
            ref = ref[ref["kept"] == 1]
            rcf = (ref["multiplicity"].astype(np.uint32) & np.uint32(0x3FFFFFFF)) | (ref["flags"].astype(np.uint32) << np.uint32(30))
            for path, rows in pl["tables"].items():
                keys, cf = rows[i]
                ok = np.array_equal(keys, ref["key_lo"]) and np.array_equal(cf, rcf)
                checked += 1
                entries += len(ref)
                if not ok:
                    bad.append({"rank": pl["rank"], "unit": int(u), "path": path})
    return {"ok": not bad, "units_checked": checked, "entries_checked": entries, "paths": sorted(gathered[0]["tables"].keys()),
            "oracle": "oracle/ggcat_oracle.c on the union of all ranks' reads", "mismatches": bad[:8]}


def sampled_superkmers(sk, units, b2: int):
    """The rows of an oracle super-k-mer array that belong to the given units (bucket << b2 | second_bucket)."""
    uid = (sk["bucket"].astype(np.uint32) << np.uint32(b2)) | sk["second_bucket"].astype(np.uint32)
    return np.ascontiguousarray(sk[np.isin(uid, np.asarray(sorted(set(int(u) for u in units)), np.uint32))])


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
    cfg = WORKLOADS[wl]
    if wl in ("c3", "c5") and world > 1:
        raise SystemExit("bench.py: --workload c3 / c5 are single-GPU lines")
    n_reads = args.reads_per_gpu if args.reads_per_gpu else {"c2": READS_PER_GPU, "c4": 20_000_000, "c5": 10_000_000, "c3": 100}[wl]
    dev = torch.device("cuda", local_rank)
    d_col = h_col = None
    if wl == "c3":
        # BASELINE configs[2]: n_reads genomes of 5 Mbp, one record and one colour each (host generation, ~0.5 GB)
        from ggcat_b200 import synth
        data, offsets, colors = synth.config_c3(n_genomes=n_reads)
        n_bases = int(data.size)
        h_data = torch.from_numpy(data).pin_memory()
        h_off = torch.from_numpy(offsets.view(np.int64)).pin_memory()
        h_col = torch.from_numpy(colors.view(np.int32)).pin_memory()
        d_data, d_off, d_col = h_data.cuda(), h_off.cuda(), h_col.cuda()
        reads_per_push = args.reads_per_push if args.reads_per_push_given else 20
    elif wl == "c2":
        data, offsets = make_reads(rank, world, n_reads)
        n_bases = int(data.size)
        h_data = torch.from_numpy(data).pin_memory()
        h_off = torch.from_numpy(offsets.view(np.int64)).pin_memory()
        d_data = h_data.cuda(non_blocking=True)
        d_off = h_off.cuda(non_blocking=True)
        genome_len, err, seed_note = GENOME_PER_GPU * world, ERR, "1% errors"
        # one push per rank (--reads-per-push splits it: the NVLink push of a bucket chunk then overlaps the bucketing of the
        # next one; at C2's size the per-batch fixed costs outweigh that, measured 6.6 vs 5.x ms per step at N=8)
        reads_per_push = args.reads_per_push if args.reads_per_push_given else n_reads
    else:
        # C4 shape (BASELINE configs[3]): error-free 150 bp reads at 30x of one genome shared by all ranks, generated
        # on the device; rank r holds reads [r*R, (r+1)*R).  Full C4 is 77.5 M reads per GPU at N=8.
        from ggcat_b200 import synth
        genome_len, err, seed_note = 5 * n_reads * world, 0.0, "error-free"
        wseed = 0xC4 if wl == "c4" else 0xC5
        genome = synth.genome_codes_torch(wseed, genome_len, dev)
        d_data = synth.simulate_reads_torch(genome, n_reads, READ_LEN, 0.0, wseed + 1, first_read=rank * n_reads)
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
    b1, b2 = G.bucket_counts(int(n_bases * 1.016) if wl == "c3" else int(n_reads * world * (READ_LEN + 15)))
    if args.b1 is not None:
        b1 = args.b1
    nb = (1 << b1) + 1
    ctx = G.GGCATB200(G.Params(k=cfg["k"], m=cfg["m"], min_multiplicity=cfg["s"], buckets_count_log=b1, second_buckets_count_log=b2,
                               hash_type=cfg["hash_type"], colors=cfg["colors"], device=local_rank))
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    # per-push device views (offsets rebased to the push)
    pushes = []
    rec_len = n_bases // n_reads      # fixed-length records in every workload (150 bp reads / 5 Mbp genomes)
    for r0 in range(0, n_reads, reads_per_push):
        r1 = min(n_reads, r0 + reads_per_push)
        off = (d_off[r0:r1 + 1] - r0 * rec_len).contiguous() if r0 else d_off[:r1 + 1]
        pushes.append((d_data.data_ptr() + r0 * rec_len, off, r1 - r0, (r1 - r0) * rec_len,
                       (d_col.data_ptr() + 4 * r0) if d_col is not None else None))
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
    my_fb, my_cnt = owner.bucket_range(rank)
    # the one exchange step: ggcat_b200_peer_exchange (NVLink peer memory) or the NCCL all-to-all of ggcat_b200.dist
    do_exchange = ctx.peer_exchange if transport == "peer" else (lambda: gdist.exchange_and_import(ctx, owner, rank, world, ext))

    def step_device():
        ctx.reset()
        for ptr, off, nr, nbytes, colp in pushes:
            ctx.push_reads_device(ptr, off.data_ptr(), nr, nbytes, colp)
        last_stats[0] = ctx.finish_bucketing()   # this rank's own super-k-mers (before the exchange adds imported chunks)
        if world > 1:
            do_exchange()
        return ctx.merge_bucket_range_device(my_fb, my_cnt)

    def step_host():
        ctx.reset()
        # the library splits the push into double-buffered H2D batches
        ctx.push_reads_ptr(h_data.data_ptr(), h_off.data_ptr(), n_reads, h_col.data_ptr() if h_col is not None else None)
        ctx.finish_bucketing()
        if world > 1:
            do_exchange()
        return ctx.merge_bucket_range(my_fb, my_cnt, copy=False)  # what the C ABI hands a host: pinned table, no extra copy

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
        exposed_ms = kt.get("k_peer_sync<exposed>", (0.0, 0))[0] / n_prof       # flag kernels on the compute stream (waits for the slowest peer)
        push_ms = kt.get("k_peer_push<side streams>", (0.0, 0))[0] / n_prof     # bulk push kernels on the data stream, overlapped with phase 1
        ex = torch.tensor([float(sent), float(recvd), exposed_ms, push_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(ex, op=dist.ReduceOp.MAX)
        sent_max, recv_max, exposed_ms, push_ms = (float(x) for x in ex.tolist())
        # NVLink 5: 900 GB/s per direction nominal, ~770 GB/s achievable
        exchange = {"bytes_sent_per_gpu": int(sent_max), "bytes_received_per_gpu": int(recv_max),
                    "push_ms_per_step": push_ms, "exposed_ms_per_step": exposed_ms, "pushes_per_step": len(pushes),
                    "achieved_GBps_per_gpu": sent_max / (push_ms * 1e-3) / 1e9 if push_ms > 0 else None,
                    "frac_of_770_GBps": sent_max / (push_ms * 1e-3) / 1e9 / 770.0 if push_ms > 0 else None,
                    "note": "bucket chunks are pushed on a side stream while the next batch is bucketed; exposed = flag waits on the compute stream"}

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
            d2h = int(sum(a.nbytes for a in (tab.keys_lo, tab.keys_hi, tab.count_flags, tab.unit_offsets, tab.src_kmers,
                                             tab.color_offsets, tab.colors) if a is not None))
            tab.release()
        e2e_ms = float(np.mean(e2e_times)) * 1e3
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": (n_bases * world) / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(h_data.numel() + h_off.numel() * 8 + (h_col.numel() * 4 if h_col is not None else 0)),
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms}

    # ---- the same through the 2-bit packed entry point (SURVEY 8(b) `is_packed`): a quarter of the H2D bytes, no k_pack.
    #      Reported beside e2e, not instead of it: the reference-facing call takes ASCII records.
    e2e_packed = None
    if h_data is not None and wl == "c2" and not args.no_e2e:
        from ggcat_b200 import synth
        h_pk = torch.from_numpy(synth.pack_2bit(data)).pin_memory()

        def step_host_packed():
            ctx.reset()
            ctx.push_reads_packed_ptr(h_pk.data_ptr(), h_off.data_ptr(), n_reads)
            ctx.finish_bucketing()
            if world > 1:
                do_exchange()
            return ctx.merge_bucket_range(my_fb, my_cnt, copy=False)

        for _ in range(2):
            step_host_packed().release()
        pts = []
        for _ in range(args.steps):
            l2_flush()
            barrier()
            t0 = time.perf_counter()
            tab = step_host_packed()
            torch.cuda.synchronize()
            pts.append(time.perf_counter() - t0)
            assert int(tab.keys_lo.size) == int(n_entries), "packed input must give the same table"
            tab.release()
        p_ms = float(np.mean(pts)) * 1e3
        t = torch.tensor([p_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p_ms = float(t.item())
        e2e_packed = {"value": (n_bases * world) / (p_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": p_ms,
                      "h2d_bytes_per_step": int(h_pk.numel() + h_off.numel() * 8), "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                      "note": "host input already 2-bit packed (ggcat_b200_push_reads_packed); e2e above is the ASCII call"}

    # ---- roofline: measured DRAM traffic and the bytes each family has to move, against the measured HBM peak
    peak, peak_kind = measured_peak()
    total_kernel_ms = max(sum(v[0] for v in kt.values()), 1e-9)
    dom = max(kt, key=lambda name: kt[name][0])
    dom_ms, dom_launches = kt[dom]
    per_launch_ms = dom_ms / max(dom_launches, 1)
    launches_per_step = max(dom_launches / n_prof, 1)
    B_s = st.payload_words * 4 + st.n_superkmers * 16       # super-k-mer payload + descriptors
    N_k = st.n_kmers
    n_ent_est = st.n_superkmers * 1.1                       # split entries (super-k-mer starts + segment ends)
    # algorithmic bytes per step of every family: inputs read once + outputs written once
    algo = {
        "k_pack+k_mark": n_bases + n_bases // 4 + n_bases // 8 + n_reads * 8,
        "k_windows": n_bases // 4 + 2 * (n_bases // 8) + n_ent_est * 8,
        "k_emit": n_ent_est * 8 + st.n_superkmers * 16,
        "k_scatter": st.n_superkmers * 16 + n_bases // 4 + B_s,
        "k_merge_hash<smem>": B_s + n_entries * 12,
        "k_gather_units": 2 * n_entries * 12,
        "k_merge_hash128": B_s + N_k * 0 + n_entries * (36 if wl == "c5" else 20),
        "k_sort_units128": 2 * n_entries * (36 if wl == "c5" else 20),
        # k_partition_units / k_merge_hash<partitions> see only the records of the units above the shared-table tiers (a few
        # units at C2 sizes): no byte model without that count, they are reported by time only
    }
    traffic, traffic_src = ({}, "skipped")
    if rank == 0 and world == 1 and wl == "c2" and not args.no_traffic_pass:
        ctx.synchronize()
        traffic, traffic_src = traffic_pass()
    per_kernel = {}
    for name, (ms_f, ln) in kt.items():
        if ln == 0:
            continue
        ms1 = ms_f / n_prof
        ent = {"ms": round(ms1, 4)}
        if name in algo and ms1 > 0:
            ent["algorithmic_bytes"] = int(algo[name])
            ent["achieved_GBps"] = round(algo[name] / (ms1 * 1e-3) / 1e9, 1)
            ent["frac"] = round(algo[name] / (ms1 * 1e-3) / 1e9 / peak, 4)
        if name in traffic and ms1 > 0:
            ent["traffic_bytes"] = int(traffic[name])
            ent["traffic_GBps"] = round(traffic[name] / (ms1 * 1e-3) / 1e9, 1)
        per_kernel[name] = ent
    dom_algo = algo.get(dom, 0) / launches_per_step
    achieved = dom_algo / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    step_algo = sum(algo[name] for name in per_kernel if name in algo)
    step_traffic = sum(traffic.values()) if traffic else None
    limiter = None
    lim = ROOT / "profiles" / "r02_limiters.json"
    if lim.exists():
        try:
            limiter = json.loads(lim.read_text()).get(dom)
        except Exception:
            limiter = None
    # the DRAM-LSD pipeline moves ~195 B/base at k=31 (W+P = 12, R = 8); its HBM-bound time is what this
    # figure compares the measured step with (> 1 = faster than that pipeline could run at HBM peak, NOT a bandwidth)
    survey_bytes = 4.1 * n_bases + B_s + N_k * cfg["wp"] * (1 + 2 * cfg["R"] + 1) + n_entries * cfg["wp"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": (traffic.get(dom) / launches_per_step) if dom in traffic else None,
                "traffic_source": traffic_src, "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": dom_algo, "ms_per_launch": per_launch_ms,
                "kernel_share_of_step": dom_ms / total_kernel_ms,
                "limiter": limiter or {"kind": "issue", "note": "see profiles/ (ncu --set full): the path is issue/latency-bound, DRAM < 5 % of peak"},
                "step": {"algorithmic_bytes": int(step_algo), "traffic_bytes": int(step_traffic) if step_traffic else None,
                         "ms": ms_per_step, "achieved": step_algo / (ms_per_step * 1e-3) / 1e9,
                         "frac": step_algo / (ms_per_step * 1e-3) / 1e9 / peak,
                         "traffic_frac": (step_traffic / (ms_per_step * 1e-3) / 1e9 / peak) if step_traffic else None},
                "per_kernel": per_kernel,
                "survey_model_equiv_frac": survey_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                "survey_model": f"SURVEY 8(d) DRAM-LSD pipeline bytes (W+P={cfg['wp']}, R={cfg['R']}) / measured step time / peak; not a bandwidth"}

    # ---- N > 1: parity of sampled owned units (device path and host path) against the oracle on the union of all reads
    parity = None
    if world > 1 and wl == "c2" and args.parity_units > 0:
        parity = multi_gpu_parity(args, ctx, owner, rank, world, b1, b2, n_reads, step_device, step_host if h_data is not None else None)

    line = {
        "metric": "build Gbases/s (bucketing+k-mer merge)", "value": value, "unit": "Gbases/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64" if cfg["wp"] == 12 else "u128", "data": "synthetic",
        "config": {"workload": workload_string(wl, n_reads, world, b1, b2),
                   "l2": "flushed (256 MB write) between timed steps", "reads_per_gpu": n_reads,
                   "parallelism": f"bucket-owner x{world}" if world > 1 else "single GPU",
                   "exchange": {"peer": "k_peer_push over NVLink peer memory (CUDA IPC)", "nccl": "NCCL all_to_all_single",
                                "none": "none"}[transport]},
        "e2e": e2e,
        "e2e_packed": e2e_packed,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels_ms_per_step": {k: round(v[0] / n_prof, 4) for k, v in kt.items()},
        "exchange": exchange,
        "parity": parity,
        "counts": {"bases_per_gpu": n_bases, "superkmers": int(st.n_superkmers), "kmer_records": int(st.n_kmers),
                   "unique": int(unique), "kept": int(n_entries)},
    }
    if rank == 0 and world == 1 and wl == "c2" and not args.no_cpu_baseline:
        from oracle import oracle as O

        sr = args.sample_reads or READS_PER_GPU
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5", "c3"],
                    help="c2 = BASELINE configs[1] (the bench line); c4 = a slice of configs[3] (human-scale shape, big merge units)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pass (large c4 slices)")
    ap.add_argument("--reads-per-push", type=int, default=4_000_000, help="c4: reads per device push (one bucket chunk each)")
    ap.add_argument("--sample-reads", type=int, default=0, help="reads per CPU pass (default: the whole workload / the whole C2 batch)")
    ap.add_argument("--no-traffic-pass", action="store_true", help="skip the ncu DRAM-traffic pass (roofline.traffic from profiles/)")
    ap.add_argument("--parity-units", type=int, default=16, help="N > 1: sampled owned units per rank checked against the oracle")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline: repeat passes until this much CPU wall time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--b1", type=int, default=None, help="experiment: override buckets_count_log")
    args = ap.parse_args()
    args.reads_per_push_given = "--reads-per-push" in sys.argv
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
