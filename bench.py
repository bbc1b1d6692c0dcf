#!/usr/bin/env python
"""bench.py — methphase hot-path throughput (BASELINE.json metric: reads/s and bases/s, HBM GB/s vs peak).

A "step" is one pass of the hot path (decode -> read sets -> pileup -> methmers -> greedy join -> collect) over
one batch of synthetic windows: the whole chr20 30x workload, 98 windows, ~21 k
records, ~306 MB staged per GPU.

  value / ms_per_step   K steps with the records resident in HBM, `--in-flight` batches (default 2) driven by as
                        many host threads: the latency-bound join of one batch overlaps the other batch's kernels
  latency_ms_per_step   one step alone on the device (launches, the pool-size round trip, D2H + Fisher included)
  kernel_ms             CUDA-event time of every stage of that single step; roofline = decode_kernel
  e2e                   the same step through the C ABI from a registered host buffer: descriptors, device gather
                        over PCIe, kernels, D2H; `--e2e-batches` region chunks pipelined (producer + consumers)
  e2e_host_copy         the same from unregistered host buffers (host gather copy into pinned memory + H2D)
  cpu_baseline          the compiled reference's per-window call sequence on a bounded sample, 1 thread (N = 1 only)

`--impl reference` times the reference's own CPU implementation (oracle/_ref/pomfret methphase -t N) on the
same synthetic BAM.  Under torchrun every rank owns its own region set (weak scaling, no data-path collective).
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


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_traffic(kernel, reads_per_launch):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/traffic.json, written from the .ncu-rep by profiles/summarize_ncu.py); None if there is no capture
    of a launch of this size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p)).get(kernel)
        if t and abs(t["reads_per_launch"] - reads_per_launch) <= 0.02 * reads_per_launch:
            return t["dram_bytes"]
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


def make_workload(tmp, region_mb, cov, seed, tagged=True):
    """Synthetic chr20-like haplotagged BAM + phased VCF (the default 63.5 Mb is the
    whole contig from 1 Mb to its end)."""
    import conftest
    beg = 1000000
    end = min(64444167, beg + int(region_mb * 1e6))
    args = ["-c", str(cov), "-s", str(seed), "-C", "chr20:64444167:%d-%d" % (beg, end), "-F", "19"]
    if not tagged:
        args.append("--untagged")
    return conftest.run_synth(os.path.join(tmp, "bench_s%d" % seed), args)


def run_reference(args, data, cov):
    """The unmodified reference CLI built against the hts shim (oracle/_ref), all host threads it can use."""
    import oracle_bindings as ob
    if not os.path.exists(ob.REF_BIN):
        return {"impl": "reference", "unavailable": "oracle/_ref/pomfret not built"}
    ncpu = os.cpu_count() or 1
    times = []
    reads = bases = None
    n_steps = args.warmup + args.steps
    for it in range(n_steps):
        out = os.path.join(os.path.dirname(data["bam"]), "ref_out")
        t0 = time.perf_counter()
        subprocess.run([ob.REF_BIN, "methphase", "-t", str(ncpu), "-c", str(cov), "-o", out, "--vcf", data["vcf"],
                        data["bam"]], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    return times


def count_units(host, hb, data, cfg):
    import parity
    wins = parity.load_windows(host, hb, data["gaps"], cfg)
    reads = sum(n for _, n, _, _, _ in wins)
    bases = sum(host.window_bases(w) for w, _, _, _, _ in wins)
    return wins, reads, bases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--region-mb", type=float, default=float(os.environ.get("POMFRET_BENCH_MB", "63.5")))
    ap.add_argument("--cov", type=int, default=30)
    ap.add_argument("--cpu-sample-windows", type=int, default=12)
    ap.add_argument("--e2e-batches", type=int, default=3, help="region chunks per step on the end-to-end path")
    ap.add_argument("--in-flight", type=int, default=2, help="batches in flight when measuring device-resident throughput")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    import pomfret_b200 as pb
    import oracle_bindings as ob
    import parity
    from pomfret_b200 import build
    build.build_host()

    tmp = tempfile.mkdtemp(prefix="pomfret_bench_")
    import atexit
    import shutil
    atexit.register(shutil.rmtree, tmp, True)
    cov = args.cov
    cfg = pb.make_config(cov)
    workload = "synthetic chr20-like %dx ONT reads (MM/ML+MD, haplotagged), %.0f Mb region, phased VCF; methphase -c %d" % (
        cov, args.region_mb, cov)

    if args.impl == "reference":
        if rank != 0:
            return
        data = make_workload(tmp, args.region_mb, cov, seed=100)
        host = pb.load_host()
        hb = host.bam_open(data["bam"])
        wins, reads, bases = count_units(host, hb, data, cfg)
        times = run_reference(args, data, cov)
        if isinstance(times, dict):
            print(json.dumps(times))
            return
        mean = sum(times) / len(times)
        ncpu = os.cpu_count() or 1
        v = reads / mean
        line = {"impl": "reference", "metric": "methphase reads/s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer + f32 scores", "data": "synthetic",
                "bases_per_s": bases / mean,
                "config": {"workload": workload, "windows": len(wins), "reads_per_step": reads, "bases_per_step": bases},
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": min(ncpu, 1), "kind": "reference",
                                 "sample": "pomfret methphase -t %d on the whole workload (1 contig => 1 worker thread, "
                                           "kt_for over contigs); includes BGZF inflate and per-window BAM open" % ncpu},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    gpu = pb.load_gpu()
    if gpu.device_count() < 1:
        raise RuntimeError("no CUDA device: pomfret_b200 has no CPU fallback")
    # the staging helper threads of all ranks share the host's cores
    os.environ.setdefault("POMFRET_GPU_STAGE_THREADS", str(max(1, min(12, (os.cpu_count() or 1) // max(world, 1)))))
    host = pb.load_host()
    # weak scaling: every rank owns its own contiguous region set (different seed), no data-path collective
    data = make_workload(tmp, args.region_mb, cov, seed=100 + rank)
    hb = host.bam_open(data["bam"])
    wins, reads, bases = count_units(host, hb, data, cfg)
    ctx = gpu.init([local_rank])
    b = gpu.batch_begin(ctx, 0, local_rank)

    def stage():
        b.reset()
        for w, n, chrom, s, e in wins:
            first = b.add_reads(host.window_descs(w), n)
            b.add_window(s, e, first, n)

    def device_pass():
        b.decode(cfg.lo, cfg.hi)
        b.pileup(cfg)
        b.join(cfg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    # device-resident timing: stage + H2D outside, kernels inside
    dev_times, kern_times, e2e_times, launches = [], [], [], 0
    tsum = {}
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            sampler.start()
        stage()
        b.submit()
        barrier()
        t0 = time.perf_counter()
        device_pass()
        res, tags, ids, rc = b.collect()
        barrier()
        dt = time.perf_counter() - t0
        tm = b.timing()
        kern = tm.decode_ms + tm.readset_ms + tm.pileup_ms + tm.methmer_ms + tm.join_ms
        if it >= args.warmup:
            dev_times.append(dt)          # whole device pass: launches, the pool-size round trip, D2H + Fisher in collect
            kern_times.append(kern / 1e3)  # sum of the kernels' own CUDA-event intervals
            launches = tm.launches
            for k in ("decode_ms", "readset_ms", "pileup_ms", "methmer_ms", "join_ms", "h2d_ms"):
                tsum[k] = tsum.get(k, 0.0) + getattr(tm, k)
            tsum["decode_bytes"] = tm.decode_bytes
            tsum["pileup_bytes"] = tm.pileup_bytes
            tsum["h2d"] = tm.bytes_h2d
            tsum["d2h"] = tm.bytes_d2h
    # Throughput with the inputs resident in HBM: the K timed steps are issued by `--in-flight` host threads, each
    # with its own batch (own stream, own resident copy of the step's records), the way the front end's workers
    # keep several region chunks on the device at once.  The join kernel is latency bound (one read tagged per
    # iteration); a second batch in flight fills the SMs it leaves idle.  Every step still runs every stage
    # (rewind -> decode -> pileup -> methmers -> join -> collect) on all of its records.
    nfl = max(1, min(args.in_flight, args.steps))  # 1: the timed steps run one after the other
    fl_batches = [b]
    for i in range(1, nfl):
        bt = gpu.batch_begin(ctx, 100 + i, local_rank)
        for w, n, chrom, s, e in wins:
            first = bt.add_reads(host.window_descs(w), n)
            bt.add_window(s, e, first, n)
        bt.submit()
        fl_batches.append(bt)
    fl_dec = [None] * nfl

    def fl_worker(j, n_steps):
        bt = fl_batches[j]
        for _ in range(n_steps):
            bt.rewind()
            bt.decode(cfg.lo, cfg.hi)
            bt.pileup(cfg)
            bt.join(cfg)
            r = bt.collect()
        fl_dec[j] = [x.decision for x in r[0]]

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
        if dj is not None and dj != [r.decision for r in res]:
            raise RuntimeError("a batch in flight disagrees with the single-batch run")
    for bt in fl_batches[1:]:
        bt.end()
    # end-to-end timing through the C ABI with host buffers.  The step's windows go through the engine as
    # `--e2e-batches` region chunks, the way the front end's workers drive it: a producer thread stages and
    # submits chunk i+1 (gather copy into pinned memory + H2D) while the device works on chunk i.
    import numpy as np
    nb = max(1, min(args.e2e_batches, len(wins)))
    cuts = [round(i * len(wins) / nb) for i in range(nb + 1)]
    chunks = [wins[cuts[i]:cuts[i + 1]] for i in range(nb)]
    batches = [b] + [gpu.batch_begin(ctx, i, local_rank) for i in range(1, nb)]
    e2e_results = [None] * nb
    # the loader's descriptors of a chunk as one array (72-byte records pointing at the host copies of the BAM
    # records): one add_reads() call per chunk
    from pomfret_b200 import _ffi
    chunk_descs = []
    for ws in chunks:
        tot = sum(n for _, n, _, _, _ in ws)
        arr = (_ffi.ReadDesc * max(tot, 1))()
        o = 0
        for w, n, chrom, s, e in ws:
            C.memmove(C.byref(arr, o * C.sizeof(_ffi.ReadDesc)), host.window_descs(w), n * C.sizeof(_ffi.ReadDesc))
            o += n
        import numpy as np
        firsts = np.cumsum([0] + [n for _, n, _, _, _ in ws][:-1]).astype(np.uint32) if ws else np.zeros(0, np.uint32)
        chunk_descs.append((arr, tot, np.array([s for _, _, _, s, _ in ws], dtype=np.uint32),
                            np.array([e for _, _, _, _, e in ws], dtype=np.uint32), firsts,
                            np.array([n for _, n, _, _, _ in ws], dtype=np.uint32)))

    e2e_prof = {"stage": 0.0, "device": 0.0}

    def produce(q):
        t_p = time.perf_counter()
        for i, (bt, ws) in enumerate(zip(batches, chunks)):
            ta = time.perf_counter()
            bt.reset()
            tb = time.perf_counter()
            arr, tot, w_s, w_e, w_first, w_n = chunk_descs[i]
            bt.add_reads(arr, tot)
            tc = time.perf_counter()
            bt.add_windows(w_s, w_e, w_first, w_n)
            bt.submit()
            td = time.perf_counter()
            e2e_prof["reset"] = e2e_prof.get("reset", 0.0) + tb - ta
            e2e_prof["add_reads"] = e2e_prof.get("add_reads", 0.0) + tc - tb
            e2e_prof["submit"] = e2e_prof.get("submit", 0.0) + td - tc
            q.put(i)
        e2e_prof["stage"] += time.perf_counter() - t_p

    def consume(i, ev):
        ev.wait()
        t_c = time.perf_counter()
        bt = batches[i]
        bt.decode(cfg.lo, cfg.hi)
        bt.pileup(cfg)
        bt.join(cfg)
        e2e_results[i] = bt.collect()
        e2e_prof["device"] = max(e2e_prof["device"], 0.0) + (time.perf_counter() - t_c) / nb
        e2e_prof["transfer_ms_evt"] = e2e_prof.get("transfer_ms_evt", 0.0) + bt.timing().h2d_ms

    def e2e_step():
        # one producer (the gather copy saturates host memory bandwidth), one consumer per chunk: the device
        # stages of consecutive chunks overlap on their own streams
        evs = [threading.Event() for _ in range(nb)]

        class Q:
            def put(self, i):
                evs[i].set()
        ths = [threading.Thread(target=produce, args=(Q(),))] + [threading.Thread(target=consume, args=(i, evs[i])) for i in range(nb)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    def run_e2e():
        times = []
        h2d = d2h = 0
        for it in range(args.warmup + args.steps):
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            barrier()
            dt = time.perf_counter() - t0
            if it < args.warmup:
                for k_ in list(e2e_prof):
                    e2e_prof[k_] = 0.0
            else:
                times.append(dt)
                h2d = sum(bt.timing().bytes_h2d for bt in batches)
                d2h = sum(bt.timing().bytes_d2h for bt in batches)
        # the chunked run must give what the single batch gave
        if [r.decision for r in res] != [r.decision for i in range(nb) for r in e2e_results[i][0]]:
            raise RuntimeError("chunked end-to-end run disagrees with the single-batch run")
        if os.environ.get("POMFRET_BENCH_VERBOSE"):
            sys.stderr.write("e2e profile per step (ms): %r\n" % {k_: round((1e3 if k_ != "transfer_ms_evt" else 1.0) * v_ / args.steps, 3)
                                                                  for k_, v_ in e2e_prof.items()})
        return dict(times=times, h2d=h2d, d2h=d2h, stage=1e3 * e2e_prof["stage"] / args.steps,
                    device=1e3 * e2e_prof["device"] / args.steps)

    # (1) host buffers as the loader left them: the engine gathers the payloads into its pinned arena (host copy)
    e2e_copy = run_e2e()
    # (2) the same records in one host buffer that is registered with the engine once (pinned + mapped, the way a
    #     loader would keep its inflate buffers): no host copy, the device gathers the payloads over PCIe
    sizes = [host.window_arena(w)[1] for w, _, _, _, _ in wins]
    big = np.empty(sum((x + 63) & ~63 for x in sizes) + 8192, dtype=np.uint8)
    big_base = (big.ctypes.data + 4095) & ~4095
    deltas, o = {}, 0
    for (w, _, _, _, _), nbytes in zip(wins, sizes):
        ptr, _ = host.window_arena(w)
        C.memmove(big_base + o, ptr, nbytes)
        deltas[w] = (big_base + o) - ptr
        o += (nbytes + 63) & ~63
    for (arr, tot, *_), ws in zip(chunk_descs, chunks):
        k = 0
        for w, n, _, _, _ in ws:
            dl = deltas[w]
            for j in range(k, k + n):
                d = arr[j]
                for f in ("cigar", "seq", "mm", "ml", "md"):
                    v = getattr(d, f)
                    if v:
                        setattr(d, f, v + dl)
            k += n
    gpu.host_register(ctx, big_base, o + 4096)
    e2e_reg = run_e2e()
    gpu.host_unregister(ctx, big_base)
    e2e_times = e2e_reg["times"]
    h2d_e2e, d2h_e2e = e2e_reg["h2d"], e2e_reg["d2h"]
    sampler.stop_flag = True

    def maxr(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def sumr(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return float(x)

    lat_mean = maxr(sum(dev_times) / len(dev_times))  # one step alone on the device
    dev_mean = maxr(fl_mean)                          # per step with `nfl` batches in flight
    e2e_mean = maxr(sum(e2e_times) / len(e2e_times))
    e2e_copy_mean = maxr(sum(e2e_copy["times"]) / len(e2e_copy["times"]))
    tot_reads = sumr(reads)
    tot_bases = sumr(bases)
    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    K = args.steps
    dec_ms = tsum["decode_ms"] / K
    pile_ms = tsum["pileup_ms"] / K
    dec_gbs = tsum["decode_bytes"] / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else 0.0
    pile_gbs = tsum["pileup_bytes"] / (pile_ms * 1e-3) / 1e9 if pile_ms > 0 else 0.0
    kernels = {k: tsum[k] / K for k in ("decode_ms", "readset_ms", "pileup_ms", "methmer_ms", "join_ms", "h2d_ms")}
    kernels["kernel_sum_ms"] = 1e3 * sum(kern_times) / len(kern_times)
    # CPU baseline on a bounded sample of the same windows (rank 0, N = 1 only)
    cpu = None
    if world == 1 and os.path.exists(ob.REF_SO):
        ocfg = ob.make_config(cov)
        sample = data["gaps"][:args.cpu_sample_windows]
        t0 = time.perf_counter()
        r_reads = 0
        for chrom, s, e, _ in sample:
            r = ob.ref_window(data["bam"], chrom, s, e, ocfg)
            r_reads += r["n_reads_loaded"]
        dtc = time.perf_counter() - t0
        n_in = sum(n for (_, n, _, _, _) in wins[:len(sample)])
        cpu = {"value": n_in / dtc, "unit": "reads/s", "cores": 1, "kind": "reference",
               "sample": "first %d windows through the compiled reference's haplotag_region_given_bam call sequence "
                         "(BAM open + inflate + decode + join), 1 thread" % len(sample)}
    line = {"metric": "methphase reads/s", "value": tot_reads / dev_mean, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_mean * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer + f32 scores", "data": "synthetic",
            "bases_per_s": tot_bases / dev_mean,
            "config": {"workload": workload, "windows_per_gpu": len(wins), "reads_per_step": tot_reads,
                       "bases_per_step": tot_bases, "l2": "inputs (%.0f MB per GPU and batch) larger than L2" % (tsum["h2d"] / 1e6),
                       "batches_in_flight": nfl},
            "latency_ms_per_step": lat_mean * 1e3,
            "e2e": {"value": tot_reads / e2e_mean, "unit": "reads/s", "ms_per_step": e2e_mean * 1e3,
                    "bases_per_s": tot_bases / e2e_mean, "h2d_bytes_per_step": int(h2d_e2e),
                    "d2h_bytes_per_step": int(d2h_e2e), "batches_per_step": nb,
                    "host_buffers": "one registered (pinned, mapped) buffer per rank; payloads gathered by the device over PCIe",
                    "host_stage_ms_per_step": e2e_reg["stage"], "device_calls_ms_per_step": e2e_reg["device"]},
            "e2e_host_copy": {"value": tot_reads / e2e_copy_mean, "unit": "reads/s", "ms_per_step": 1e3 * e2e_copy_mean,
                              "host_buffers": "unregistered: payloads copied into the engine's pinned arena by host threads",
                              "host_stage_ms_per_step": e2e_copy["stage"], "h2d_bytes_per_step": int(e2e_copy["h2d"])},
            "gpu_launches": int(launches) * K,
            "kernel_ms": kernels,
            "roofline": {"kernel": "decode_kernel", "bound": "hbm", "achieved": dec_gbs, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": dec_gbs / peaks["hbm_gbs"],
                         "traffic": measured_traffic("decode_kernel", reads), "algorithmic_bytes": int(tsum["decode_bytes"]),
                         "launch_ms": dec_ms, "peak_kind": peak_kind},
            "roofline_pileup": {"kernel": "pileup_tile_kernel+sites_finalize_kernel", "bound": "hbm", "achieved": pile_gbs,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": pile_gbs / peaks["hbm_gbs"]},
            "clocks": sampler.summary()}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


if __name__ == "__main__":
    main()
