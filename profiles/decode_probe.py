"""Decode kernel probe (GPU box): share of records on the lean path, decode time with and without it.
usage: python profiles/decode_probe.py [contigs] [contig_mb] [extra pomfret-synth arguments, e.g. --implicit 0.05]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, conftest, parity
import pomfret_b200 as pb
import tempfile
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mb = float(sys.argv[2]) if len(sys.argv) > 2 else 2.5
tmp = tempfile.mkdtemp()
extra = sys.argv[3:]   # e.g. --implicit 0.05: non-CpG C+m entries => implicit canonical calls => the general (lane 0) path
data = conftest.run_synth(os.path.join(tmp, "s"), bench.synth_args(nc, mb, 60, 100) + extra)
host = pb.load_host(); gpu = pb.load_gpu()
cfg = pb.make_config(60)
hb = host.bam_open(data["bam"])
wins = parity.load_windows(host, hb, data["gaps"], cfg)
ctx = gpu.init([0])
for lean in ("1", "0"):
    os.environ["POMFRET_GPU_DECODE_LEAN"] = lean
    b = gpu.batch_begin(ctx)
    for w, n, chrom, s, e in wins:
        first = b.add_reads(host.window_descs(w), n)
        b.add_window(s, e, first, n)
    b.submit()
    for it in range(4):
        if it: b.rewind()
        b.decode(cfg.lo, cfg.hi); b.pileup(cfg); b.join(cfg); b.collect()
    t = b.timing()
    cnt = collections.Counter()
    for i in range(0, b.n_reads, 7):
        st, ncalls, end = b.read_info(i)
        cnt["lean" if st & 128 else ("generic" if st & 16 else "stream")] += 1
    print("lean=%s reads=%d decode_ms=%.3f GB/s=%.0f paths=%s" % (lean, b.n_reads, t.decode_ms, t.decode_bytes / t.decode_ms / 1e6, dict(cnt)))
    b.end()
