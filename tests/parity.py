"""Shared parity checks: CUDA path (through the C ABI) against the oracle port / compiled reference."""
import numpy as np

import oracle_bindings as ob
import pomfret_b200 as pb

REF_KEYS = ["sites_fwd", "starts_fwd", "lens_fwd", "sites_bwd", "starts_bwd", "lens_bwd", "calls_off", "calls_pos",
            "calls_cat", "strand", "revbuf", "ids_left", "ids_left_strict", "ids_right", "ids_right_strict",
            "mmr_off_fwd", "mmr_fwd", "mmr_start_fwd", "mmr_off_bwd", "mmr_bwd", "mmr_start_bwd", "tags_fwd",
            "tags_bwd", "tags_final"]
REF_SCALARS = ["n_reads", "n_reads_loaded", "decision", "join_fwd", "join_bwd"]


def diff_dicts(a, b, keys=REF_KEYS, scalars=REF_SCALARS):
    bad = [k for k in keys if not (len(a[k]) == len(b[k]) and np.array_equal(a[k], b[k]))]
    bad += [k for k in scalars if a[k] != b[k]]
    return bad


def load_windows(host, bam_handle, gaps, cfg, readlen=None):
    """Host half of load_reads_given_interval for every gap; returns [(handle, n, chrom, s, e)]."""
    out = []
    for chrom, s, e, _ in gaps:
        w = host.window_load(bam_handle, chrom, s, e, cfg.readlen_threshold if readlen is None else readlen,
                             cfg.min_mapq)
        out.append((w, host.window_n(w), chrom, s, e))
    return out


def run_gpu_batch(gpu, ctx, host, wins, cfg):
    b = gpu.batch_begin(ctx)
    layout = []
    for w, n, chrom, s, e in wins:
        first = b.add_reads(host.window_descs(w), n)
        b.add_window(s, e, first, n)
        layout.append((first, n))
    b.submit()
    b.decode(cfg.lo, cfg.hi)
    b.pileup(cfg)
    b.join(cfg)
    res, tags, ids, rc = b.collect(check=False)
    return b, layout, res, tags, ids, rc


def compare_window(b, wi, first, n, res, tags, ids, p, deep=True):
    """Compare one window of a collected batch with the oracle port's dict `p`. Returns a list of diffs."""
    bad = []
    r = res[wi]
    for i in range(n):
        st, nc, end = b.read_info(first + i)
        pst = int(p["status"][i])
        if (st & 15) != (pst & 15):
            bad.append(("status", i, st, pst))
        rid = int(p["read_ids"][i])
        if rid != int(ids[first + i]):
            bad.append(("read_id", i, rid, int(ids[first + i])))
        if rid >= 0 and deep:
            pos, cat = b.calls(first + i)
            a, z = int(p["calls_off"][rid]), int(p["calls_off"][rid + 1])
            if not (np.array_equal(pos, p["calls_pos"][a:z]) and np.array_equal(cat, p["calls_cat"][a:z])):
                bad.append(("calls", i, len(pos), z - a))
            if end != int(p["end_pos"][rid]):
                bad.append(("end_pos", i, end, int(p["end_pos"][rid])))
    for d, dn in ((0, "fwd"), (1, "bwd")):
        pos, st, ln = b.sites(wi, d)
        if p["n_reads"] > 0:
            if not (np.array_equal(pos, p["sites_" + dn]) and np.array_equal(st, p["starts_" + dn])
                    and np.array_equal(ln, p["lens_" + dn])):
                bad.append(("sites", dn, len(pos), len(p["sites_" + dn])))
        if deep:
            for i in range(n):
                rid = int(p["read_ids"][i])
                if rid < 0 or rid >= p["n_reads"]:
                    continue
                m, sti = b.mmrs(first + i, d)
                a, z = int(p["mmr_off_" + dn][rid]), int(p["mmr_off_" + dn][rid + 1])
                if not np.array_equal(m, p["mmr_" + dn][a:z]) or sti != int(p["mmr_start_" + dn][rid]):
                    bad.append(("mmr", dn, i, len(m), z - a, sti, int(p["mmr_start_" + dn][rid])))
        order = b.tag_order(wi, d)
        if not np.array_equal(order, p["order_" + dn]):
            bad.append(("order", dn, len(order), len(p["order_" + dn])))
        if p["n_reads"] > 0 and len(p["sites_fwd"]) > 0:
            pt = b.prop_tags(d)[first:first + p["n_reads"]]
            if not np.array_equal(pt, p["prop_" + dn]):
                bad.append(("prop_tags", dn))
    if p["n_reads"] > 0 and len(p["sites_fwd"]) > 0:
        if list(r.table_fwd) != p["table_fwd"] or list(r.table_bwd) != p["table_bwd"]:
            bad.append(("tables", list(r.table_fwd), p["table_fwd"], list(r.table_bwd), p["table_bwd"]))
        if r.which_way_fwd != p["which_way_fwd"] or r.which_way_bwd != p["which_way_bwd"]:
            bad.append(("which_way", r.which_way_fwd, p["which_way_fwd"], r.which_way_bwd, p["which_way_bwd"]))
        # fp32 scores: ratios of small integers computed with the same IEEE ops -> exact
        if r.score_fwd != p["score_fwd"] or r.score_bwd != p["score_bwd"]:
            bad.append(("scores", r.score_fwd, p["score_fwd"], r.score_bwd, p["score_bwd"]))
    for k in ("decision", "join_fwd", "join_bwd", "n_reads", "n_reads_loaded"):
        if getattr(r, k) != p[k]:
            bad.append((k, getattr(r, k), p[k]))
    kept = [i for i in range(n) if ids[first + i] >= 0]
    ft = np.array([tags[first + i] for i in kept], dtype=np.uint8)[:p["n_reads"]]
    if not np.array_equal(ft, p["tags_final"]):
        bad.append(("final_tags",))
    return bad


def run_gpu_batch_shared(gpu, ctx, host, wins, cfg):
    """The windows of `wins` in one batch through the decode-once entry point: a record that an earlier window of
    the batch already staged (same position, length, CIGAR size and name) is added as a reference to that read."""
    import ctypes as C
    from pomfret_b200 import _ffi
    dsz = C.sizeof(_ffi.ReadDesc)
    b = gpu.batch_begin(ctx)
    seen = {}
    layout, n_shared = [], 0
    for w, n, chrom, s, e in wins:
        same = np.full(max(n, 1), -1, dtype=np.int64)
        names = host.window_qnames(w)
        base = host.window_descs(w)
        for i in range(n):
            d = _ffi.ReadDesc.from_address(base + i * dsz)
            key = (names[i], d.pos, d.l_qseq, d.n_cigar)
            if key in seen:
                same[i] = seen[key]
                n_shared += 1
            else:
                seen[key] = b.n_reads + i
        first = b.add_reads_shared(base, n, same)
        b.add_window(s, e, first, n)
        layout.append((first, n))
    b.submit()
    b.decode(cfg.lo, cfg.hi)
    b.pileup(cfg)
    b.join(cfg)
    res, tags, ids, rc = b.collect(check=False)
    return b, layout, res, tags, ids, rc, n_shared
