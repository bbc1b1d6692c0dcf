"""ctypes bindings for the checkers under oracle/ (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpomfret_ref.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "pomfret")
PORT_SO = os.path.join(ROOT, "oracle", "liboracle_port.so")

u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)


class ReadDesc(C.Structure):
    """pomfret_gpu_read_desc (include/pomfret_gpu.h)"""
    _fields_ = [("pos", C.c_uint32), ("l_qseq", C.c_uint32), ("n_cigar", C.c_uint32), ("flag", C.c_uint16),
                ("mapq", C.c_uint8), ("tags_malformed", C.c_uint8), ("hp", C.c_int32), ("mn", C.c_int32),
                ("cigar", C.c_void_p), ("seq", C.c_void_p), ("mm", C.c_void_p), ("mm_len", C.c_uint32),
                ("ml_len", C.c_int32), ("ml", C.c_void_p), ("md", C.c_void_p), ("md_len", C.c_uint32),
                ("reserved", C.c_uint32)]


class Config(C.Structure):
    """pomfret_gpu_config"""
    _fields_ = [(n, C.c_int32) for n in ("k", "k_span", "lo", "hi", "cov_known", "cov_for_selection",
                                         "cov_for_runtime", "readlen_threshold", "min_mapq",
                                         "n_candidates_per_iter")]


def make_config(cov, k=3, k_span=5000, lo=100, hi=156, readlen=15000, mapq=10, report=False):
    """cli.c:270-275 (-c COV) and blockjoin.c:4657 / 5045-5051 (report adds +1)"""
    sel = cov // 10 + (1 if report else 0)
    ncand = cov // 4 + (1 if report else 0)
    if not report:
        if sel <= 0:
            sel = 1  # blockjoin.c:4381-4385 (cov_for_runtime keeps the unclamped product)
        if ncand <= 1:
            ncand = 2
    run = (cov // 10 + (1 if report else 0)) * 2
    return Config(k, k_span, lo, hi, cov, sel, run, readlen, mapq, ncand)


class RefWin(C.Structure):
    _fields_ = [("n_reads_loaded", C.c_int), ("n_reads", C.c_int), ("decision", C.c_int), ("join1", C.c_int),
                ("join2", C.c_int), ("skipped", C.c_int), ("n_sites_fwd", C.c_int), ("n_sites_bwd", C.c_int),
                ("sites_fwd", u32p), ("starts_fwd", u32p), ("sites_bwd", u32p), ("starts_bwd", u32p),
                ("lens_fwd", u8p), ("lens_bwd", u8p), ("hp_init", C.POINTER(C.c_int)), ("strand", u8p),
                ("len", u32p), ("calls_off", u32p), ("calls_pos", u32p), ("calls_cat", u8p),
                ("mmr_off_fwd", u32p), ("mmr_fwd", u32p), ("mmr_start_fwd", u32p),
                ("mmr_off_bwd", u32p), ("mmr_bwd", u32p), ("mmr_start_bwd", u32p),
                ("tags_fwd", u8p), ("tags_bwd", u8p), ("tags_final", u8p), ("revbuf", u64p),
                ("n_left", C.c_uint32), ("n_left_strict", C.c_uint32), ("n_right", C.c_uint32),
                ("n_right_strict", C.c_uint32), ("ids_left", u32p), ("ids_left_strict", u32p),
                ("ids_right", u32p), ("ids_right_strict", u32p), ("qnames", C.c_void_p),
                ("qnames_len", C.c_uint32), ("qname_off", u32p)]


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.refwin_run.restype = C.POINTER(RefWin)
        lib.refwin_run.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32] + [C.c_int] * 10 + [C.c_void_p]
        lib.refwin_free.argtypes = [C.POINTER(RefWin)]
        lib.refwin_run_whole.restype = C.c_int
        lib.refwin_run_whole.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32] + [C.c_int] * 10 + \
            [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        lib.refh_pre_haplotag.restype = C.c_void_p
        lib.refh_pre_haplotag.argtypes = [C.c_char_p, C.c_char_p]
        lib.refh_tags_hash.restype = C.c_void_p
        lib.refh_tags_hash.argtypes = [C.c_void_p]
        lib.refh_tag_lookup.argtypes = [C.c_void_p, C.c_char_p]
        lib.refh_tags_count.argtypes = [C.c_void_p]
        lib.refh_tags_free.argtypes = [C.c_void_p]
        lib.refh_load_intervals.restype = C.c_void_p
        lib.refh_load_intervals.argtypes = [C.c_char_p, C.c_int]
        for f in ("refh_intervals_nref",):
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.refh_intervals_refname.restype = C.c_char_p
        lib.refh_intervals_refname.argtypes = [C.c_void_p, C.c_int]
        lib.refh_intervals_n.argtypes = [C.c_void_p, C.c_int]
        lib.refh_intervals_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.refh_intervals_decide.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.refh_intervals_finish.argtypes = [C.c_void_p]
        lib.refh_intervals_nblocks.argtypes = [C.c_void_p, C.c_int]
        lib.refh_intervals_blocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.refh_intervals_free.argtypes = [C.c_void_p]
        lib.refh_fisher_two_sided.restype = C.c_double
        lib.refh_fisher_two_sided.argtypes = [C.c_int] * 4
        lib.refh_evaluate_separation1.restype = C.c_float
        lib.refh_evaluate_separation1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        lib.refh_get_mod_poss_on_ref.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                                 C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.POINTER(C.c_int)]
        lib.refh_load_variants.argtypes = [C.c_char_p, C.c_char_p] + [C.c_void_p] * 6 + [C.c_int, C.c_int]
        _ref = lib
    return _ref


class quiet_reference:
    """The reference logs heavily to stderr; silence it only while it runs (REFH_QUIET=0 keeps it)."""

    def __enter__(self):
        self.on = os.environ.get("REFH_QUIET", "1") != "0"
        if self.on:
            ref_lib().refh_quiet(1)

    def __exit__(self, *a):
        if self.on:
            ref_lib().refh_quiet(0)
        return False


def ref_window(bam, chrom, start, end, cfg, raw_tags=None):
    """Run the compiled reference on one window; returns a dict of numpy arrays."""
    lib = ref_lib()
    with quiet_reference():
        p = lib.refwin_run(bam.encode(), chrom.encode(), start, end, cfg.k, cfg.k_span, cfg.lo, cfg.hi,
                           cfg.cov_known, cfg.cov_for_selection, cfg.cov_for_runtime, cfg.readlen_threshold,
                           cfg.min_mapq, cfg.n_candidates_per_iter, raw_tags)
    w = p.contents
    nl, n = w.n_reads_loaded, w.n_reads
    out = dict(n_reads_loaded=nl, n_reads=n, decision=w.decision, join_fwd=w.join1, join_bwd=w.join2,
               skipped=w.skipped)
    out["sites_fwd"] = _arr(w.sites_fwd, w.n_sites_fwd, np.uint32)
    out["starts_fwd"] = _arr(w.starts_fwd, w.n_sites_fwd, np.uint32)
    out["lens_fwd"] = _arr(w.lens_fwd, w.n_sites_fwd, np.uint8)
    out["sites_bwd"] = _arr(w.sites_bwd, w.n_sites_bwd, np.uint32)
    out["starts_bwd"] = _arr(w.starts_bwd, w.n_sites_bwd, np.uint32)
    out["lens_bwd"] = _arr(w.lens_bwd, w.n_sites_bwd, np.uint8)
    out["hp_init"] = _arr(w.hp_init, nl, np.int32)
    out["strand"] = _arr(w.strand, nl, np.uint8)
    out["len"] = _arr(w.len, nl, np.uint32)
    off = _arr(w.calls_off, nl + 1, np.uint32)
    out["calls_off"] = off
    tot = int(off[-1]) if nl else 0
    out["calls_pos"] = _arr(w.calls_pos, tot, np.uint32)
    out["calls_cat"] = _arr(w.calls_cat, tot, np.uint8)
    for d in ("fwd", "bwd"):
        mo = _arr(getattr(w, "mmr_off_" + d), n + 1, np.uint32)
        out["mmr_off_" + d] = mo
        out["mmr_" + d] = _arr(getattr(w, "mmr_" + d), int(mo[-1]) if n else 0, np.uint32)
        out["mmr_start_" + d] = _arr(getattr(w, "mmr_start_" + d), n, np.uint32)
        out["tags_" + d] = _arr(getattr(w, "tags_" + d), n, np.uint8)
    out["tags_final"] = _arr(w.tags_final, n, np.uint8)
    out["revbuf"] = _arr(w.revbuf, nl, np.uint64)
    for s in ("left", "left_strict", "right", "right_strict"):
        out["ids_" + s] = _arr(getattr(w, "ids_" + s), getattr(w, "n_" + s), np.uint32)
    qo = _arr(w.qname_off, nl + 1, np.uint32)
    raw = C.string_at(w.qnames, w.qnames_len) if w.qnames_len else b""
    out["qnames"] = [raw[qo[i]:qo[i + 1]].split(b"\0")[0].decode() for i in range(nl)]
    lib.refwin_free(p)
    return out


class PortCalls(C.Structure):
    _fields_ = [("pos", u32p), ("cat", u8p), ("n", C.c_uint32), ("m", C.c_uint32)]


class PortSites(C.Structure):
    _fields_ = [("n", C.c_int), ("real_pos", u32p), ("starts", u32p), ("lens", u8p)]


class PortReadset(C.Structure):
    _fields_ = [("n", C.c_int), ("n_loaded", C.c_int), ("ref_start", C.c_uint32), ("ref_end", C.c_uint32),
                ("hp", C.POINTER(C.c_int)), ("strand", u8p), ("start_pos", u32p), ("end_pos", u32p),
                ("calls", C.POINTER(PortCalls)), ("revbuf", u64p),
                ("n_left", C.c_uint32), ("n_left_strict", C.c_uint32), ("n_right", C.c_uint32),
                ("n_right_strict", C.c_uint32), ("ids_left", u32p), ("ids_left_strict", u32p),
                ("ids_right", u32p), ("ids_right_strict", u32p),
                ("mmr", C.POINTER(u32p)), ("mmr_n", C.POINTER(C.c_int)), ("mmr_start_i", u32p)]


class PortWindow(C.Structure):
    _fields_ = [("decision", C.c_int), ("join_fwd", C.c_int), ("join_bwd", C.c_int), ("n_reads", C.c_int),
                ("n_reads_loaded", C.c_int), ("n_sites_fwd", C.c_int), ("n_sites_bwd", C.c_int),
                ("table_fwd", C.c_int * 4), ("table_bwd", C.c_int * 4), ("score_fwd", C.c_float),
                ("score_bwd", C.c_float), ("which_way_fwd", C.c_int), ("which_way_bwd", C.c_int),
                ("sites", PortSites * 2), ("tags_final", u8p), ("tags_fwd", u8p), ("tags_bwd", u8p),
                ("read_ids", i32p), ("status", u32p), ("rs", C.POINTER(PortReadset)),
                ("mmr_bwd", C.POINTER(u32p)), ("mmr_n_bwd", C.POINTER(C.c_int)), ("mmr_start_bwd", u32p),
                ("order_fwd", u32p), ("order_bwd", u32p), ("n_order_fwd", C.c_int), ("n_order_bwd", C.c_int),
                ("status_code", C.c_int), ("prop_fwd", u8p), ("prop_bwd", u8p)]


_port = None


def port_lib():
    global _port
    if _port is None:
        lib = C.CDLL(PORT_SO)
        lib.port_window_run.restype = C.POINTER(PortWindow)
        lib.port_window_run.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(Config)]
        lib.port_window_free.argtypes = [C.POINTER(PortWindow)]
        lib.port_fisher_two_sided.restype = C.c_double
        lib.port_fisher_two_sided.argtypes = [C.c_int] * 4
        lib.port_evaluate_separation.restype = C.c_float
        lib.port_evaluate_separation.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        lib.port_map_mods_to_ref.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_uint32, C.POINTER(PortCalls)]
        lib.port_decode_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(PortCalls), u32p, u32p]
        lib.port_haptag_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.port_haptag_cursors.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
        _port = lib
    return _port


def port_window(descs_ptr, n, start, end, cfg):
    """Run the oracle port on packed records; returns a dict shaped like ref_window()."""
    lib = port_lib()
    p = lib.port_window_run(descs_ptr, n, start, end, C.byref(cfg))
    w = p.contents
    rs = w.rs.contents
    nl, nr = w.n_reads_loaded, w.n_reads
    out = dict(n_reads_loaded=nl, n_reads=nr, decision=w.decision, join_fwd=w.join_fwd, join_bwd=w.join_bwd,
               status_code=w.status_code, table_fwd=list(w.table_fwd), table_bwd=list(w.table_bwd),
               score_fwd=w.score_fwd, score_bwd=w.score_bwd, which_way_fwd=w.which_way_fwd,
               which_way_bwd=w.which_way_bwd)
    for di, d in enumerate(("fwd", "bwd")):
        s = w.sites[di]
        out["sites_" + d] = _arr(s.real_pos, s.n, np.uint32)
        out["starts_" + d] = _arr(s.starts, s.n, np.uint32)
        out["lens_" + d] = _arr(s.lens, s.n, np.uint8)
        out["tags_" + d] = _arr(getattr(w, "tags_" + d), nl, np.uint8)[:nr]
        out["prop_" + d] = _arr(getattr(w, "prop_" + d), nl, np.uint8)[:nr]
    out["tags_final"] = _arr(w.tags_final, nl, np.uint8)[:nr]
    out["read_ids"] = _arr(w.read_ids, n, np.int32)
    out["status"] = _arr(w.status, n, np.uint32)
    out["hp_init"] = None
    pos, cat, off = [], [], [0]
    for i in range(nl):
        c = rs.calls[i]
        pos.append(_arr(c.pos, c.n, np.uint32))
        cat.append(_arr(c.cat, c.n, np.uint8))
        off.append(off[-1] + c.n)
    out["calls_off"] = np.array(off, dtype=np.uint32)
    out["calls_pos"] = np.concatenate(pos) if pos else np.zeros(0, np.uint32)
    out["calls_cat"] = np.concatenate(cat) if cat else np.zeros(0, np.uint8)
    out["strand"] = _arr(rs.strand, nl, np.uint8)
    out["end_pos"] = _arr(rs.end_pos, nl, np.uint32)
    out["revbuf"] = _arr(rs.revbuf, nl, np.uint64)
    for s in ("left", "left_strict", "right", "right_strict"):
        out["ids_" + s] = _arr(getattr(rs, "ids_" + s), getattr(rs, "n_" + s), np.uint32)
    # methmers: fwd are the ones left in the read set, bwd were kept aside
    for d, (mm, mn, ms) in (("fwd", (rs.mmr, rs.mmr_n, rs.mmr_start_i)),
                            ("bwd", (w.mmr_bwd, w.mmr_n_bwd, w.mmr_start_bwd))):
        arrs, moff = [], [0]
        for i in range(nr):
            k = mn[i]
            arrs.append(_arr(mm[i], k, np.uint32) if k else np.zeros(0, np.uint32))
            moff.append(moff[-1] + k)
        out["mmr_off_" + d] = np.array(moff, dtype=np.uint32)
        out["mmr_" + d] = np.concatenate(arrs) if arrs else np.zeros(0, np.uint32)
        out["mmr_start_" + d] = _arr(ms, nr, np.uint32)
    out["order_fwd"] = _arr(w.order_fwd, w.n_order_fwd, np.uint32)
    out["order_bwd"] = _arr(w.order_bwd, w.n_order_bwd, np.uint32)
    lib.port_window_free(p)
    return out
