"""Region sharding for multi-GPU runs: windows are independent, so the ordered
(contig, window) list is cut into `world` contiguous region sets balanced by the estimated number of
records each window pulls in (gap + 2 x 50 kb halo, reference blockjoin.c:19,1053-1054); no data-path
collective, the host gathers one decision per window and one tag byte per read."""

READBACK = 50000  # blockjoin.c:19


def estimate_reads(start, end, cov, mean_readlen=20000):
    """records of the region query chrom:(start-50k)-(end+50k) at coverage `cov`"""
    span = (end - start) + 2 * READBACK + mean_readlen
    return max(1.0, span * float(cov) / mean_readlen)


def partition_windows(windows, world, cov, mean_readlen=20000):
    """windows: ordered list of (chrom, start, end, ...).  Returns `world` (begin, end) index ranges,
    contiguous and in order, with near-equal estimated record counts."""
    w = [estimate_reads(x[1], x[2], cov, mean_readlen) for x in windows]
    n, total = len(w), sum(w)
    cuts, i, acc = [0], 0, 0.0
    for r in range(world - 1):
        target = total * (r + 1) / world
        if i < n:  # a shard is never empty while windows are left
            acc += w[i]
            i += 1
        # take windows while that brings the shard closer to its share, leaving one for every later shard
        while i < n and n - i > world - 1 - r and acc + w[i] / 2 <= target:
            acc += w[i]
            i += 1
        cuts.append(i)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_results(local, rank, world, dist=None):
    """Host-side gather of per-window results (rank order == window order).  `dist` is torch.distributed
    (any backend); with world == 1 the list is returned as is."""
    if world == 1 or dist is None:
        return list(local)
    out = [None] * world
    dist.all_gather_object(out, list(local))
    merged = []
    for part in out:
        merged.extend(part)
    return merged
