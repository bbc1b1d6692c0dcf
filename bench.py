#!/usr/bin/env python
"""bench.py — methphase hot-path throughput (BASELINE.json metric: reads/s and bases/s, HBM GB/s vs peak).

Workload (BASELINE.json config 3 at the size one GPU and a few minutes allow): a synthetic 60x haplotagged sample
over `--contigs` regions of the hg38 primary contigs (`--contig-mb` each, 20 kb ONT-like reads with MM/ML + MD,
whatshap-style phased VCF).  Its phase-block windows are staged `--replicas` times as distinct records so that one
batch has the record count of a few percent of a 60x genome (~0.6 M reads, ~8 GB in HBM, seconds of device work per
timed region).  A "step" is one pass of the hot path (decode -> read sets -> pileup -> methmers -> greedy join ->
collect) over that batch.

  value / ms_per_step   K steps, records resident in HBM, `--in-flight` batches (default 2) from as many host threads
  latency_ms_per_step   one step alone on the device (launches, the pool-size round trip, D2H + Fisher included)
  kernel_ms             CUDA-event time of every stage of one step; roofline = decode_kernel, roofline_pileup
  e2e                   the same step through the C ABI from registered HOST buffers: descriptors, device gather over
                        PCIe, kernels, D2H of decisions and tags; region chunks pipelined
  cli_e2e               wall time of the drop-in CLI (`pomfret methphase`, BGZF inflate and writers included) next to the
                        unmodified reference CLI on the same files — the like-for-like number; cli_report: `report`
  untagged              config 4: the -u read haplotagger over an untagged 30x sample (roofline_haptag); cli_untagged, and
                        cli_write_bam: the same run with --write-bam (output BAM re-tagged on the device, blocks compressed
                        on the host threads, BAI from the block table; .mp.bam and .bai compared byte for byte)
  strong                N > 1: the same batch cut into one contiguous region set per rank (no data-path collective,
                        host gather of decisions, checked against rank 0's own full run)
  cpu_baseline          the compiled reference's per-window call sequence on a bounded sample, 1 thread (N = 1 only)

`--impl reference` times the unmodified reference CLI (oracle/_ref/pomfret methphase -t <cores>) on the same sample.
Every rank works on the same sample (rank 0 generates it); weak scaling = every rank runs the whole batch.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# hg38 primary contigs (the ##contig list of the reference's example VCF), longest first
HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
        ("chr6", 170805979), ("chr7", 159345973), ("chrX", 156040895), ("chr8", 145138636), ("chr9", 138394717),
        ("chr11", 135086622), ("chr10", 133797422), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718),
        ("chr15", 101991189), ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr20", 64444167),
        ("chr19", 58617616), ("chrY", 57227415), ("chr22", 50818468), ("chr21", 46709983)]


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_traffic(kernel, reads_per_launch):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/traffic.json, written from
    the .ncu-rep by profiles/summarize_ncu.py), scaled by the record count when the capture is of a smaller launch of the
    same workload; None if there is no capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p)).get(kernel)
        if t and t.get("workload") == "60x" and t["reads_per_launch"] > 0:
            return int(t["dram_bytes"] * reads_per_launch / t["reads_per_launch"])
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.samples = []
        self.reasons = set()
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                    for n, v in zip(names, f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def synth_args(n_contigs, contig_mb, cov, seed, tagged=True):
    args = ["-c", str(cov), "-s", str(seed)]
    for name, length in HG38[:n_contigs]:
        beg = 1000000
        end = min(length, beg + int(contig_mb * 1e6))
        args += ["-C", "%s:%d:%d-%d" % (name, length, beg, end)]
    if not tagged:
        args.append("--untagged")
    return args


def make_workload(tmp, tag, n_contigs, contig_mb, cov, seed, tagged=True):
    import conftest
    return conftest.run_synth(os.path.join(tmp, tag), synth_args(n_contigs, contig_mb, cov, seed, tagged))


def wall(cmd, env=None, want_startup=False):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env, text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("%s failed: %s" % (" ".join(cmd[:3]), p.stderr[-1500:]))
    if want_startup:  # the front end reports how long the CUDA driver / context start-up took
        import re
        m = re.search(r"engine ready after ([0-9.]+)s", p.stderr)
        return dt, float(m.group(1)) if m else None
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cov", type=int, default=60)
    ap.add_argument("--contigs", type=int, default=env_int("POMFRET_BENCH_CONTIGS", 16))
    ap.add_argument("--contig-mb", type=float, default=float(os.environ.get("POMFRET_BENCH_CONTIG_MB", "5")))
    ap.add_argument("--replicas", type=int, default=env_int("POMFRET_BENCH_REPLICAS", 6),
                    help="times the sample's windows are staged (as distinct records) to form one WGS-scale batch")
    ap.add_argument("--cpu-sample-windows", type=int, default=24)
    ap.add_argument("--e2e-batches", type=int, default=8, help="region chunks per step on the end-to-end path")
    ap.add_argument("--in-flight", type=int, default=2, help="batches in flight when measuring device-resident throughput")
    ap.add_argument("--skip-cli", action="store_true", help="leave out the CLI wall-time comparisons (profiling runs)")
    ap.add_argument("--skip-untagged", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    import pomfret_b200 as pb
    import oracle_bindings as ob
    import parity
    from pomfret_b200 import build, shard, _ffi
    build.build_host()

    import atexit
    import shutil
    cov = args.cov
    cfg = pb.make_config(cov)
    workload = ("synthetic %dx ONT-like reads (20 kb, MM/ML+MD, haplotagged) on %d hg38 contig regions of %.1f Mb, phased VCF; "
                "methphase -c %d" % (cov, args.contigs, args.contig_mb, cov))
    ncpu = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        tmp = tempfile.mkdtemp(prefix="pomfret_bench_")
        atexit.register(shutil.rmtree, tmp, True)
        data = make_workload(tmp, "s60", args.contigs, args.contig_mb, cov, seed=100)
        if not os.path.exists(ob.REF_BIN):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/pomfret not built"}))
            return
        host = pb.load_host()
        hb = host.bam_open(data["bam"])
        wins = parity.load_windows(host, hb, data["gaps"], cfg)
        reads = sum(n for _, n, _, _, _ in wins)
        bases = sum(host.window_bases(w) for w, _, _, _, _ in wins)
        n_gap_contigs = len({c for c, _, _, _ in data["gaps"]})
        times = []
        for it in range(args.warmup + args.steps):
            dt = wall([ob.REF_BIN, "methphase", "-t", str(ncpu), "-c", str(cov), "-o", os.path.join(tmp, "ref_out"), "--vcf",
                       data["vcf"], data["bam"]])
            if it >= args.warmup:
                times.append(dt)
        mean = sum(times) / len(times)
        v = reads / mean
        cores = max(1, min(ncpu, n_gap_contigs))
        line = {"impl": "reference", "metric": "methphase reads/s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer + f32 scores", "data": "synthetic",
                "bases_per_s": bases / mean,
                "config": {"workload": workload, "windows": len(wins), "reads_per_step": reads, "bases_per_step": bases,
                           "replicas": 1},
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference",
                                 "sample": "pomfret methphase -t %d on the un-replicated sample (%d contigs with gaps => %d worker "
                                           "threads busy, kt_for over contigs); BGZF inflate, per-window BAM open and the "
                                           "writers included" % (ncpu, n_gap_contigs, cores)},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- our arm ----------------
    import numpy as np
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    gpu = pb.load_gpu()
    if gpu.device_count() < 1:
        raise RuntimeError("no CUDA device: pomfret_b200 has no CPU fallback")
    os.environ.setdefault("POMFRET_GPU_STAGE_THREADS", str(max(1, min(12, ncpu // max(world, 1)))))
    host = pb.load_host()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one sample for all ranks: rank 0 generates, the others wait for it
    tmp_holder = [None]
    if rank == 0:
        tmp_holder[0] = tempfile.mkdtemp(prefix="pomfret_bench_")
        atexit.register(shutil.rmtree, tmp_holder[0], True)
    if world > 1:
        dist.broadcast_object_list(tmp_holder, src=0)
    tmp = tmp_holder[0]
    data = None
    if rank == 0:
        data = make_workload(tmp, "s60", args.contigs, args.contig_mb, cov, seed=100)
    if world > 1:
        box = [data]
        dist.broadcast_object_list(box, src=0)
        data = box[0]
    hb = host.bam_open(data["bam"])
    wins = parity.load_windows(host, hb, data["gaps"], cfg)
    base_reads = sum(n for _, n, _, _, _ in wins)
    base_bases = sum(host.window_bases(w) for w, _, _, _, _ in wins)
    R = max(1, args.replicas)
    # the batch: every window of the sample, R times (window order: replica-major)
    batch_windows = [(r, i) for r in range(R) for i in range(len(wins))]
    reads = base_reads * R
    bases = base_bases * R
    ctx = gpu.init([local_rank])

    # the records of the sample in ONE host buffer registered with the engine (pinned + mapped), the way a loader keeps
    # the buffers it inflates BGZF blocks into; every replica's descriptors point into it
    sizes = [host.window_arena(w)[1] for w, _, _, _, _ in wins]
    big_bytes = sum((x + 63) & ~63 for x in sizes) + 8192
    # pinned by the allocator (cudaHostAlloc through torch: what a loader that owns its inflate buffers would use; the
    # engine takes such memory as it is) unless POMFRET_BENCH_PINNED=0: then plain memory, pinned by host_register()
    pinned_alloc = os.environ.get("POMFRET_BENCH_PINNED", "1") != "0"
    if pinned_alloc:
        try:
            big = torch.empty(big_bytes, dtype=torch.uint8, pin_memory=True)
            big_ptr = big.data_ptr()
            probe = (big_ptr + 4095) & ~4095
            gpu.host_register(ctx, probe, 4096)      # does the engine take this memory?
            gpu.host_unregister(ctx, probe)
        except Exception as exc:  # fall back to plain memory pinned by host_register()
            sys.stderr.write("[bench] pinned allocation not usable (%s): malloc + cudaHostRegister\n" % exc)
            pinned_alloc = False
    if not pinned_alloc:
        big = np.empty(big_bytes, dtype=np.uint8)
        big_ptr = big.ctypes.data
    big_base = (big_ptr + 4095) & ~4095
    win_descs = []
    o = 0
    dsz = C.sizeof(_ffi.ReadDesc)
    for (w, n, _, _, _), nbytes in zip(wins, sizes):
        ptr, _ = host.window_arena(w)
        C.memmove(big_base + o, ptr, nbytes)
        delta = (big_base + o) - ptr
        arr = (_ffi.ReadDesc * max(n, 1))()
        C.memmove(arr, host.window_descs(w), n * dsz)
        for j in range(n):
            d = arr[j]
            for f in ("cigar", "seq", "mm", "ml", "md"):
                v = getattr(d, f)
                if v:
                    setattr(d, f, v + delta)
        win_descs.append(arr)
        o += (nbytes + 63) & ~63
    gpu.host_register(ctx, big_base, o + 4096)

    def stage(bt, window_list):
        # (descriptors point into the registered buffer: add_reads() lays the batch out, the device gathers the payloads)
        bt.reset()
        for _, i in window_list:
            w, n, chrom, s, e = wins[i]
            first = bt.add_reads(win_descs[i], n)
            bt.add_window(s, e, first, n)

    def device_pass(bt):
        bt.decode(cfg.lo, cfg.hi)
        bt.pileup(cfg)
        bt.join(cfg)

    sampler = ClockSampler(local_rank)
    b = gpu.batch_begin(ctx, 0, local_rank)
    stage(b, batch_windows)
    b.submit()
    # ---- single-step latency and per-stage kernel times (records resident, one batch) ----
    dev_times, tsum, launches = [], {}, 0
    res = None
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            sampler.start()
        if it:
            b.rewind()
        barrier()
        t0 = time.perf_counter()
        device_pass(b)
        res, tags, ids, rc = b.collect()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = b.timing()
        if it >= args.warmup:
            dev_times.append(dt)
            launches = tm.launches
            for k in ("decode_ms", "readset_ms", "pileup_ms", "methmer_ms", "join_ms", "h2d_ms"):
                tsum[k] = tsum.get(k, 0.0) + getattr(tm, k)
            tsum["decode_bytes"] = tm.decode_bytes
            tsum["pileup_bytes"] = tm.pileup_bytes
            tsum["h2d"] = tm.bytes_h2d
    full_decisions = [x.decision for x in res]
    for r in range(1, R):  # replicas are independent copies of the same windows
        if full_decisions[r * len(wins):(r + 1) * len(wins)] != full_decisions[:len(wins)]:
            raise RuntimeError("replica %d of the window set disagrees with replica 0" % r)
    # ---- throughput with the inputs resident in HBM: K steps from `--in-flight` host threads, one batch each ----
    nfl = max(1, min(args.in_flight, args.steps))
    fl_batches = [b]
    for i in range(1, nfl):
        bt = gpu.batch_begin(ctx, 100 + i, local_rank)
        stage(bt, batch_windows)
        bt.submit()
        fl_batches.append(bt)
    fl_dec = [None] * nfl

    def fl_worker(j, n_steps):
        bt = fl_batches[j]
        r = None
        for _ in range(n_steps):
            bt.rewind()
            device_pass(bt)
            r = bt.collect()
        fl_dec[j] = [x.decision for x in r[0]] if r else None

    def fl_run(total_steps):
        share = [total_steps // nfl + (1 if j < total_steps % nfl else 0) for j in range(nfl)]
        ths = [threading.Thread(target=fl_worker, args=(j, share[j])) for j in range(nfl) if share[j] > 0]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    fl_run(max(args.warmup, nfl))
    barrier()
    t0 = time.perf_counter()
    fl_run(args.steps)
    barrier()
    fl_mean = (time.perf_counter() - t0) / args.steps
    for dj in fl_dec:
        if dj is not None and dj != full_decisions:
            raise RuntimeError("a batch in flight disagrees with the single-batch run")
    for bt in fl_batches[1:]:
        bt.end()

    # ---- strong scaling over the same batch: one contiguous region set per rank, host gather of the decisions ----
    strong = None
    if world > 1:
        win_list = [(wins[i][2], wins[i][3], wins[i][4]) for _, i in batch_windows]
        a, z = shard.partition_windows(win_list, world, cov)[rank]
        stage(b, batch_windows[a:z])
        b.submit()
        local = None
        for it in range(args.warmup):
            if it:
                b.rewind()
            device_pass(b)
            local = b.collect()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            b.rewind()
            device_pass(b)
            local = b.collect()
        barrier()
        st_mean = (time.perf_counter() - t0) / args.steps
        merged = shard.gather_results([x.decision for x in local[0]], rank, world, dist)
        if merged != full_decisions:
            raise RuntimeError("region-sharded run over %d ranks disagrees with the single-rank run of the same batch" % world)
        strong = st_mean
        stage(b, batch_windows)  # back to the full batch for what follows
        b.submit()

    # ---- end to end through the C ABI from registered host buffers (descriptors -> device gather over PCIe ->
    #      kernels -> D2H of decisions + tags), `--e2e-batches` region chunks pipelined: producer + one consumer each ----
    nb = max(1, min(args.e2e_batches, len(batch_windows)))
    cuts = [round(i * len(batch_windows) / nb) for i in range(nb + 1)]
    chunks = [batch_windows[cuts[i]:cuts[i + 1]] for i in range(nb)]
    batches = [b] + [gpu.batch_begin(ctx, i, local_rank) for i in range(1, nb)]
    e2e_results = [None] * nb
    chunk_descs = []
    for ws in chunks:
        tot = sum(wins[i][1] for _, i in ws)
        arr = (_ffi.ReadDesc * max(tot, 1))()
        k = 0
        for _, i in ws:
            n = wins[i][1]
            C.memmove(C.byref(arr, k * dsz), win_descs[i], n * dsz)
            k += n
        ns = [wins[i][1] for _, i in ws]
        firsts = np.cumsum([0] + ns[:-1]).astype(np.uint32) if ws else np.zeros(0, np.uint32)
        chunk_descs.append((arr, tot, np.array([wins[i][3] for _, i in ws], dtype=np.uint32),
                            np.array([wins[i][4] for _, i in ws], dtype=np.uint32), firsts, np.array(ns, dtype=np.uint32)))
    e2e_prof = {"stage": 0.0, "device": 0.0}

    def produce(evs):
        t_p = time.perf_counter()
        for i, bt in enumerate(batches):
            bt.reset()
            arr, tot, w_s, w_e, w_first, w_n = chunk_descs[i]
            bt.add_reads(arr, tot)
            bt.add_windows(w_s, w_e, w_first, w_n)
            bt.submit()
            evs[i].set()
        e2e_prof["stage"] += time.perf_counter() - t_p

    def consume(i, ev):
        ev.wait()
        t_c = time.perf_counter()
        bt = batches[i]
        device_pass(bt)
        e2e_results[i] = bt.collect()
        e2e_prof["device"] += (time.perf_counter() - t_c) / nb

    def e2e_step():
        evs = [threading.Event() for _ in range(nb)]
        ths = [threading.Thread(target=produce, args=(evs,))] + [threading.Thread(target=consume, args=(i, evs[i])) for i in range(nb)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    e2e_times = []
    h2d_e2e = d2h_e2e = 0
    n_e2e = max(3, min(args.steps, 10))  # (a step moves the whole batch over PCIe: ten of them are seconds of traffic)
    for it in range(3 + n_e2e):
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if it < 3:
            for k_ in list(e2e_prof):
                e2e_prof[k_] = 0.0
        else:
            e2e_times.append(dt)
            h2d_e2e = sum(bt.timing().bytes_h2d for bt in batches)
            d2h_e2e = sum(bt.timing().bytes_d2h for bt in batches)
    if full_decisions != [r.decision for i in range(nb) for r in e2e_results[i][0]]:
        raise RuntimeError("chunked end-to-end run disagrees with the single-batch run")
    gpu.host_unregister(ctx, big_base)
    sampler.stop_flag = True
    for bt in batches[1:]:
        bt.end()

    def maxr(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    lat_mean = maxr(sum(dev_times) / len(dev_times))
    dev_mean = maxr(fl_mean)
    e2e_mean = maxr(sum(e2e_times) / len(e2e_times))
    strong_mean = maxr(strong) if strong is not None else None
    tot_reads = float(reads) * world   # weak scaling: every rank runs the whole batch
    tot_bases = float(bases) * world

    # ---- config 4: the -u read haplotagger over an untagged 30x sample (rank 0) ----
    untagged = None
    if rank == 0 and not args.skip_untagged:
        udata = make_workload(tmp, "u30", min(args.contigs, 6), min(args.contig_mb, 2.0), 30, seed=130, tagged=False)
        lib = host.lib
        lib.pomfret_host_contig_load.restype = C.c_void_p
        lib.pomfret_host_contig_load.argtypes = [C.c_void_p, C.c_char_p]
        lib.pomfret_host_load_variants.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        uhb = host.bam_open(udata["bam"])
        ub = gpu.batch_begin(ctx, 50, local_rank)
        u_reads = u_bases = 0
        u_ms = u_bytes = 0.0
        rep_u = 8
        for name, _ in HG38[:min(args.contigs, 6)]:
            w = lib.pomfret_host_contig_load(uhb, name.encode())
            n = host.window_n(w)
            cap = 1 << 18
            vars_ = (pb.Variant * cap)()
            vb = np.zeros(cap * 4, np.uint8)
            nbv = C.c_int()
            nk = lib.pomfret_host_load_variants(udata["vcf"].encode(), name.encode(), vars_, cap, vb.ctypes.data, cap * 4, C.byref(nbv))
            if n == 0 or nk <= 0:
                continue
            known = np.frombuffer(vars_, dtype=np.uint8, count=nk * C.sizeof(pb.Variant)).copy()
            var_pos = np.array([vars_[i].pos for i in range(nk)], dtype=np.uint32)
            descs = host.window_descs(w)
            starts = np.array([_ffi.ReadDesc.from_address(descs + i * dsz).pos for i in range(n)], dtype=np.uint32)
            kf = np.searchsorted(var_pos, starts, side="left").astype(np.uint32)  # the i_left cursor of blockjoin.c:1716-1720
            ub.reset()
            for _ in range(rep_u):
                ub.add_reads(descs, n)
            ub.submit()
            kfr = np.tile(kf, rep_u)
            for it in range(3):
                if it:
                    ub.rewind()
                ub.haptag(known, nk, vb[:nbv.value], kfr)
                utags, ustat = ub.collect_haptags()
            t = ub.timing()
            u_ms += t.haptag_ms
            u_bytes += t.haptag_bytes
            u_reads += n * rep_u
            u_bases += host.window_bases(w) * rep_u
            host.window_free(w)
        ub.end()
        host.bam_close(uhb)
        if u_ms > 0:
            untagged = {"reads": u_reads, "bases": u_bases, "haptag_ms": u_ms, "bytes": u_bytes, "data": udata}

    # ---- compressed ingest (what the front end does instead of host inflate): BGZF blocks of the sample's region queries
    #      from host memory -> inflate_kernel -> record walk + slicing, per contig (rank 0) ----
    ingest = None
    if rank == 0:
        ib = gpu.batch_begin(ctx, 60, local_rank)
        flt = _ffi.IngestFilter(cfg.min_mapq, cfg.readlen_threshold, 2, 1, 0.1)
        tot = {"in": 0, "out": 0, "inflate_ms": 0.0, "slice_ms": 0.0, "records": 0, "blocks": 0, "wall": 0.0}
        for chrom in sorted({g[0] for g in data["gaps"]}):
            regions = [(max(max(s_ - 50000, 0) - 1, 0), e_ + 50000) for c_, s_, e_, _ in data["gaps"] if c_ == chrom]
            plan = host.ingest_plan(hb, chrom, regions)
            for it in range(2):
                ib.reset()
                t0 = time.perf_counter()
                rc_, recs_, n_ = ib.ingest_bgzf(plan["comp"], plan["comp_bytes"], plan["blocks"], plan["n_blocks"], plan["streams"],
                                                plan["n_streams"], flt)
                dt_ = time.perf_counter() - t0
            t = ib.timing()
            tot["in"] += t.inflate_in_bytes; tot["out"] += t.inflate_out_bytes; tot["inflate_ms"] += t.inflate_ms
            tot["slice_ms"] += t.slice_ms; tot["records"] += n_; tot["blocks"] += plan["n_blocks"]; tot["wall"] += dt_
            host.ingest_free(plan)
        ib.end()
        if tot["inflate_ms"] > 0:
            ingest = tot

    # ---- the drop-in CLI against the unmodified reference CLI on the same files (rank 0; uses all N devices) ----
    cli = {}
    if rank == 0 and not args.skip_cli and os.path.exists(ob.REF_BIN):
        mine = os.path.join(ROOT, "pomfret_b200", "bin", "pomfret")
        thr = str(max(2, min(ncpu, 16)))
        env = dict(os.environ)
        env.pop("POMFRET_GPU_STAGE_THREADS", None)
        for key, sub, extra, d in (("cli_e2e", "methphase", ["-c", str(cov)], data),
                                   ("cli_report", "report", ["-c", str(cov), "--chunk-size", "50000", "--chunk-stride", "100000"], data),
                                   ("cli_untagged", "methphase", ["-u", "-c", "30"], untagged["data"] if untagged else None),
                                   ("cli_write_bam", "methphase", ["-u", "-c", "30", "--write-bam"], untagged["data"] if untagged else None)):
            if d is None:
                continue
            po, pr = os.path.join(tmp, key + "_ours"), os.path.join(tmp, key + "_ref")
            common = extra + ["--vcf", d["vcf"], d["bam"]]
            # (no --gpus: the front end takes one device per 4 GiB of BAM, i.e. one for these samples; its multi-GPU path is
            #  covered by tests/test_gpu_cli.py on files cut over --gpus 2 / all)
            t_ours, t_start = min(wall([mine, sub, "-t", thr, "-o", po] + common, env, True) for _ in range(2))
            t_ref = wall([ob.REF_BIN, sub, "-t", thr, "-o", pr] + common)
            suffixes = [".report.tsv"] if sub == "report" else [".mp.gtf", ".mp.vcf"] + ([".mp.bam", ".mp.bam.bai"] if "--write-bam" in extra else [])
            same = all(open(po + s_, "rb").read() == open(pr + s_, "rb").read() for s_ in suffixes)
            if not same:
                raise RuntimeError("%s: output files differ from the reference's" % key)
            cli[key] = {"ours_s": t_ours, "reference_s": t_ref, "speedup": t_ref / t_ours, "threads": int(thr), "gpus": 1,
                        "outputs_identical": True, "ours_cuda_startup_s": t_start,
                        "speedup_without_cuda_startup": t_ref / max(t_ours - (t_start or 0.0), 1e-3),
                        "what": "wall time of `pomfret %s %s` on the un-replicated sample files, index, VCF and writers included. "
                                "The reference inflates BGZF through the single-threaded zlib shim on %s threads; this "
                                "front end ships the blocks to the device (inflate_kernel).  ours_s includes the CUDA "
                                "driver/context start-up of the process (ours_cuda_startup_s), which a whole-genome "
                                "run amortises and a sample of this size does not" % (sub, " ".join(extra), thr)}
    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    K = args.steps
    dec_ms = tsum["decode_ms"] / K
    pile_ms = tsum["pileup_ms"] / K
    dec_gbs = tsum["decode_bytes"] / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else 0.0
    pile_gbs = tsum["pileup_bytes"] / (pile_ms * 1e-3) / 1e9 if pile_ms > 0 else 0.0
    kernels = {k: tsum[k] / K for k in ("decode_ms", "readset_ms", "pileup_ms", "methmer_ms", "join_ms")}
    kernels["kernel_sum_ms"] = sum(kernels.values())
    # CPU baseline on a bounded sample of the same windows (rank 0, N = 1 only)
    cpu = None
    if world == 1 and os.path.exists(ob.REF_SO):
        ocfg = ob.make_config(cov)
        sample = data["gaps"][:args.cpu_sample_windows]
        t0 = time.perf_counter()
        for chrom, s, e, _ in sample:
            ob.ref_window(data["bam"], chrom, s, e, ocfg)
        dtc = time.perf_counter() - t0
        n_in = sum(n for (_, n, _, _, _) in wins[:len(sample)])
        cpu = {"value": n_in / dtc, "unit": "reads/s", "cores": 1, "kind": "reference",
               "sample": "first %d windows of the sample through the compiled reference's haplotag_region_given_bam call "
                         "sequence (BAM open + inflate + decode + join), 1 thread, %.1f s" % (len(sample), dtc)}
    line = {"metric": "methphase reads/s", "value": tot_reads / dev_mean, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_mean * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer + f32 scores", "data": "synthetic",
            "bases_per_s": tot_bases / dev_mean,
            "config": {"workload": workload, "windows": len(wins), "replicas": R, "windows_per_gpu": len(batch_windows),
                       "reads_per_step": tot_reads, "bases_per_step": tot_bases, "reads_in_sample": base_reads,
                       "l2": "inputs (%.0f MB per GPU and batch) larger than L2" % (tsum["h2d"] / 1e6),
                       "batches_in_flight": nfl, "timed_region_s": dev_mean * K},
            "latency_ms_per_step": lat_mean * 1e3,
            "e2e": {"value": tot_reads / e2e_mean, "unit": "reads/s", "ms_per_step": e2e_mean * 1e3,
                    "bases_per_s": tot_bases / e2e_mean, "h2d_bytes_per_step": int(h2d_e2e),
                    "d2h_bytes_per_step": int(d2h_e2e), "batches_per_step": nb, "steps": n_e2e,
                    "host_memory": "cudaHostAlloc (torch pinned tensor)" if pinned_alloc else "malloc + cudaHostRegister",
                    "host_buffers": "C ABI from one registered (pinned, mapped) host buffer per rank holding already inflated "
                                    "BAM records; payloads gathered by the device over PCIe; BGZF inflate is NOT in this number "
                                    "(cli_e2e has it)",
                    "host_stage_ms_per_step": 1e3 * e2e_prof["stage"] / n_e2e, "device_calls_ms_per_step": 1e3 * e2e_prof["device"] / n_e2e},
            "gpu_launches": int(launches) * K,
            "kernel_ms": kernels,
            "roofline": {"kernel": "decode_kernel", "bound": "hbm", "achieved": dec_gbs, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": dec_gbs / peaks["hbm_gbs"],
                         "traffic": measured_traffic("decode_kernel", reads), "algorithmic_bytes": int(tsum["decode_bytes"]),
                         "launch_ms": dec_ms, "peak_kind": peak_kind},
            "roofline_pileup": {"kernel": "pileup_tile_kernel+sites_finalize_kernel", "bound": "hbm", "achieved": pile_gbs,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": pile_gbs / peaks["hbm_gbs"],
                                "algorithmic_bytes": int(tsum["pileup_bytes"]), "launch_ms": pile_ms},
            "windows_per_s": len(batch_windows) * world / dev_mean,
            "clocks": sampler.summary()}
    if strong_mean is not None:
        line["strong"] = {"value": reads / strong_mean, "unit": "reads/s", "ms_per_step": strong_mean * 1e3,
                          "what": "the same batch (%d windows) cut into %d contiguous region sets, one per rank; decisions "
                                  "gathered on the host and equal to the single-rank run" % (len(batch_windows), world),
                          "efficiency_vs_one_rank_latency": lat_mean / (strong_mean * world)}
    if untagged:
        hap_gbs = untagged["bytes"] / (untagged["haptag_ms"] * 1e-3) / 1e9
        line["untagged"] = {"value": untagged["reads"] / (untagged["haptag_ms"] * 1e-3), "unit": "reads/s",
                            "bases_per_s": untagged["bases"] / (untagged["haptag_ms"] * 1e-3), "haptag_ms": untagged["haptag_ms"],
                            "what": "haptag_kernel over every primary record of an untagged 30x sample (config 4), records "
                                    "resident, CUDA-event time"}
        line["roofline_haptag"] = {"kernel": "haptag_kernel", "bound": "hbm", "achieved": hap_gbs, "peak": peaks["hbm_gbs"],
                                   "unit": "GB/s", "frac": hap_gbs / peaks["hbm_gbs"], "algorithmic_bytes": int(untagged["bytes"])}
    if ingest:
        gbs_out = ingest["out"] / (ingest["inflate_ms"] * 1e-3) / 1e9
        line["ingest"] = {"inflate_out_gbs": gbs_out, "inflate_in_gbs": ingest["in"] / (ingest["inflate_ms"] * 1e-3) / 1e9,
                          "inflate_ms": ingest["inflate_ms"], "slice_ms": ingest["slice_ms"], "blocks": ingest["blocks"],
                          "records": ingest["records"], "compressed_bytes": int(ingest["in"]), "inflated_bytes": int(ingest["out"]),
                          "call_s": ingest["wall"],
                          "what": "BGZF blocks of the sample's region queries from host memory: H2D, inflate_kernel (one warp per block, "
                                  "ISIZE + CRC-32 checked), record walk and slicing; one launch per contig; call_s is the wall time of the "
                                  "C-ABI calls incl. the D2H of the record headers"}
        line["roofline_inflate"] = {"kernel": "inflate_kernel", "bound": "hbm", "achieved": gbs_out + ingest["in"] / (ingest["inflate_ms"] * 1e-3) / 1e9,
                                    "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": (gbs_out + ingest["in"] / (ingest["inflate_ms"] * 1e-3) / 1e9) / peaks["hbm_gbs"],
                                    "algorithmic_bytes": int(ingest["in"] + ingest["out"]),
                                    "note": "bit-serial Huffman decoding: bound by the latency of one lane per block, not by HBM"}
    line.update(cli)
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


if __name__ == "__main__":
    main()
