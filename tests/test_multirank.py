"""N > 1 path on CPU: two gloo ranks shard the windows of one data set, run their region sets through
the engine (the SIMT-emulated build of the kernels, there is no GPU here) and gather the decisions on
the host; the result must equal the single-rank run window by window."""
import os
import sys

import pytest
import torch.multiprocessing as mp

from pomfret_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_contiguous_and_balanced():
    wins = [("chr1", 100000 * i, 100000 * i + 20000 + 7000 * (i % 5)) for i in range(37)]
    for world in (1, 2, 3, 4, 8):
        parts = shard.partition_windows(wins, world, 30)
        assert parts[0][0] == 0 and parts[-1][1] == len(wins)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        loads = [sum(shard.estimate_reads(w[1], w[2], 30) for w in wins[a:b]) for a, b in parts]
        assert max(loads) <= 1.5 * sum(loads) / world + shard.estimate_reads(0, 48000, 30)
    assert shard.partition_windows(wins[:2], 4, 30) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def _rank_main(rank, world, port, data, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
    import torch.distributed as dist
    import build_emu
    import parity
    import pomfret_b200 as pb
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    gpu = pb.load_gpu(build_emu.build())
    host = pb.load_host()
    cfg = pb.make_config(36, readlen=2000)
    a, b = shard.partition_windows(data["gaps"], world, 36, 4000)[rank]
    hb = host.bam_open(data["bam"])
    wins = parity.load_windows(host, hb, data["gaps"][a:b], cfg)
    ctx = gpu.init()
    local = []
    if wins:
        bt, layout, res, tags, ids, rc = parity.run_gpu_batch(gpu, ctx, host, wins, cfg)
        assert rc == 0
        for (w, n, chrom, s, e), (first, _), r in zip(wins, layout, res):
            local.append((chrom, s, e, r.decision, r.n_reads, bytes(tags[first:first + n])))
        bt.end()
    merged = shard.gather_results(local, rank, world, dist)
    dist.barrier()
    if rank == 0:
        import pickle
        pickle.dump(merged, open(out_path, "wb"))
    dist.destroy_process_group()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.emu
def test_two_gloo_ranks_match_single_rank(built, synth_small, tmp_path):
    import pickle
    import build_emu
    build_emu.build()
    outs = {}
    for world in (1, 2):
        out = str(tmp_path / ("w%d.pkl" % world))
        mp.spawn(_rank_main, args=(world, _free_port(), synth_small, out), nprocs=world, join=True)
        outs[world] = pickle.load(open(out, "rb"))
    assert len(outs[1]) == len(synth_small["gaps"])
    assert outs[1] == outs[2]
