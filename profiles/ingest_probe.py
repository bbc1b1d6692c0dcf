"""Compressed-ingest probe (GPU box): inflate / slice kernel times and rates on the 60x sample's windows."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, conftest
import pomfret_b200 as pb
from pomfret_b200 import _ffi
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tmp = tempfile.mkdtemp()
data = conftest.run_synth(os.path.join(tmp, "s"), bench.synth_args(nc, 2.5, 60, 100))
host = pb.load_host(); gpu = pb.load_gpu()
cfg = pb.make_config(60)
hb = host.bam_open(data["bam"])
ctx = gpu.init([0])
b = gpu.batch_begin(ctx)
for chrom in sorted({g[0] for g in data["gaps"]}):
    regions = [(max(max(s - 50000, 0) - 1, 0), e + 50000) for c, s, e, _ in data["gaps"] if c == chrom]
    t0 = time.perf_counter()
    plan = host.ingest_plan(hb, chrom, regions)
    t1 = time.perf_counter()
    flt = _ffi.IngestFilter(cfg.min_mapq, cfg.readlen_threshold, 2, 1, 0.1)
    for it in range(3):
        b.reset()
        t2 = time.perf_counter()
        rc, recs, n = b.ingest_bgzf(plan["comp"], plan["comp_bytes"], plan["blocks"], plan["n_blocks"], plan["streams"], plan["n_streams"], flt)
        t3 = time.perf_counter()
    t = b.timing()
    print("%s: %d regions, %d blocks, %d streams, comp %.1f MB -> %.1f MB, %d records | plan+read %.3fs, ingest call %.3fs, inflate %.3f ms (%.1f GB/s out), slice %.3f ms"
          % (chrom, len(regions), plan["n_blocks"], plan["n_streams"], plan["comp_bytes"] / 1e6, t.inflate_out_bytes / 1e6, n, t1 - t0, t3 - t2,
             t.inflate_ms, t.inflate_out_bytes / max(t.inflate_ms, 1e-9) / 1e6, t.slice_ms))
    host.ingest_free(plan)
