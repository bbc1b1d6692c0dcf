"""ctypes bindings of the C ABI (include/pomfret_gpu.h) and of the host front end's test hooks."""
import ctypes as C
import os

import numpy as np

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(PACKAGE_DIR, "lib")

u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class ReadDesc(C.Structure):
    """pomfret_gpu_read_desc"""
    _fields_ = [("pos", C.c_uint32), ("l_qseq", C.c_uint32), ("n_cigar", C.c_uint32), ("flag", C.c_uint16),
                ("mapq", C.c_uint8), ("tags_malformed", C.c_uint8), ("hp", C.c_int32), ("mn", C.c_int32),
                ("cigar", C.c_void_p), ("seq", C.c_void_p), ("mm", C.c_void_p), ("mm_len", C.c_uint32),
                ("ml_len", C.c_int32), ("ml", C.c_void_p), ("md", C.c_void_p), ("md_len", C.c_uint32),
                ("reserved", C.c_uint32)]


class Config(C.Structure):
    """pomfret_gpu_config (mirrors mmr_config_t, reference blockjoin.h:7-16)"""
    _fields_ = [(n, C.c_int32) for n in ("k", "k_span", "lo", "hi", "cov_known", "cov_for_selection",
                                         "cov_for_runtime", "readlen_threshold", "min_mapq",
                                         "n_candidates_per_iter")]


class Variant(C.Structure):
    """pomfret_gpu_variant"""
    _fields_ = [("pos", C.c_uint32), ("len", C.c_uint32), ("op", C.c_uint8), ("haptag", C.c_uint8),
                ("reserved", C.c_uint16), ("bases_off", C.c_uint32)]


class WindowResult(C.Structure):
    """pomfret_gpu_window_result"""
    _fields_ = [("decision", C.c_int32), ("join_fwd", C.c_int32), ("join_bwd", C.c_int32), ("n_reads", C.c_int32),
                ("n_reads_loaded", C.c_int32), ("n_sites_fwd", C.c_int32), ("n_sites_bwd", C.c_int32),
                ("n_left", C.c_int32), ("n_left_strict", C.c_int32), ("n_right", C.c_int32),
                ("n_right_strict", C.c_int32), ("table_fwd", C.c_int32 * 4), ("table_bwd", C.c_int32 * 4),
                ("score_fwd", C.c_float), ("score_bwd", C.c_float), ("which_way_fwd", C.c_int32),
                ("which_way_bwd", C.c_int32), ("status", C.c_int32)]


class BgzfBlock(C.Structure):
    """pomfret_gpu_bgzf_block"""
    _fields_ = [("comp_off", C.c_uint64), ("csize", C.c_uint32), ("isize", C.c_uint32), ("out_off", C.c_uint64)]


class BgzfStream(C.Structure):
    """pomfret_gpu_bgzf_stream"""
    _fields_ = [("out_off", C.c_uint64), ("out_bytes", C.c_uint64), ("ubeg", C.c_uint32), ("tid", C.c_int32), ("end0", C.c_uint32),
                ("first_block", C.c_uint32), ("n_blocks", C.c_uint32), ("reserved", C.c_uint32)]


class IngestFilter(C.Structure):
    """pomfret_gpu_ingest_filter"""
    _fields_ = [("min_mapq", C.c_uint32), ("min_len", C.c_uint32), ("min_len_floor", C.c_uint32), ("check_de", C.c_uint32),
                ("max_de", C.c_float), ("keep_all_flags", C.c_uint32)]


class SlicedRecord(C.Structure):
    """pomfret_gpu_sliced_record"""
    _fields_ = [("pos", C.c_uint32), ("end_pos", C.c_uint32), ("l_qseq", C.c_uint32), ("n_cigar", C.c_uint32), ("flag", C.c_uint16),
                ("mapq", C.c_uint8), ("tags_malformed", C.c_uint8), ("hp", C.c_int32), ("mn", C.c_int32), ("mm_len", C.c_uint32),
                ("ml_len", C.c_int32), ("md_len", C.c_uint32), ("stream", C.c_uint32), ("keep", C.c_uint8), ("bad", C.c_uint8),
                ("has_mm", C.c_uint8), ("hp_irregular", C.c_uint8), ("l_qname", C.c_uint8), ("hp_type", C.c_uint8), ("cg_cigar", C.c_uint8), ("pad", C.c_uint8),
                ("cigar", C.c_uint64), ("seq", C.c_uint64), ("mm", C.c_uint64), ("ml", C.c_uint64), ("md", C.c_uint64),
                ("qname_dev", C.c_uint64), ("rec_bytes", C.c_uint32), ("hp_off", C.c_uint32), ("tid", C.c_int32),
                ("reserved", C.c_uint32), ("qname", C.c_char * 48)]


class Timing(C.Structure):
    """pomfret_gpu_timing"""
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "decode_ms", "haptag_ms", "readset_ms", "pileup_ms",
                                         "methmer_ms", "join_ms", "d2h_ms")] + \
               [(n, C.c_uint64) for n in ("bytes_h2d", "bytes_d2h", "decode_bytes", "pileup_bytes",
                                          "methmer_bytes", "haptag_bytes")] + [("launches", C.c_uint32)] + \
               [("inflate_ms", C.c_float), ("slice_ms", C.c_float), ("inflate_in_bytes", C.c_uint64),
                ("inflate_out_bytes", C.c_uint64)]


def make_config(cov, k=3, k_span=5000, lo=100, hi=156, readlen=15000, mapq=10, report=False):
    """`-c COV` as the front end derives it: cli.c:270-275, blockjoin.c:4381-4390, 4657 (methphase) and
    blockjoin.c:5045-5051 (report: +1)."""
    extra = 1 if report else 0
    sel = cov // 10 + extra
    ncand = cov // 4 + extra
    run = sel * 2
    if not report:
        if sel <= 0:
            sel = 1
        if ncand <= 1:
            ncand = 2
    return Config(k, k_span, lo, hi, cov, sel, run, readlen, mapq, ncand)


class GpuError(RuntimeError):
    def __init__(self, lib, rc, where):
        self.rc = rc
        super().__init__("%s failed: %d (%s)" % (where, rc, lib.strerror(rc)))


class GpuLib:
    """Loaded libpomfret_gpu.so.  There is no fallback: a missing library raises."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "pomfret_b200 has no CPU fallback" % path)
        self.path = path
        lib = C.CDLL(path)
        self.lib = lib
        vp = C.c_void_p
        lib.pomfret_gpu_strerror.restype = C.c_char_p
        lib.pomfret_gpu_strerror.argtypes = [C.c_int]
        lib.pomfret_gpu_version.restype = C.c_char_p
        lib.pomfret_gpu_init.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int]
        lib.pomfret_gpu_destroy.argtypes = [vp]
        lib.pomfret_gpu_batch_begin.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
        lib.pomfret_gpu_host_register.argtypes = [vp, vp, C.c_size_t]
        lib.pomfret_gpu_host_unregister.argtypes = [vp, vp]
        lib.pomfret_gpu_batch_reset.argtypes = [vp]
        lib.pomfret_gpu_batch_add_read.argtypes = [vp, vp]
        lib.pomfret_gpu_batch_add_reads.argtypes = [vp, vp, C.c_uint32]
        lib.pomfret_gpu_batch_add_reads_shared.argtypes = [vp, vp, C.c_uint32, vp]
        lib.pomfret_gpu_batch_add_reads_device.argtypes = [vp, vp, C.c_uint32, vp]
        lib.pomfret_gpu_batch_ingest_buffer.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        lib.pomfret_gpu_batch_ingest_bgzf.argtypes = [vp, vp, C.c_size_t, vp, C.c_uint32, vp, C.c_uint32, vp, u32p]
        lib.pomfret_gpu_batch_ingest_records.argtypes = [vp, vp, C.c_uint32]
        lib.pomfret_gpu_debug_get_inflated.argtypes = [vp, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.pomfret_gpu_batch_ingest_qname.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
        lib.pomfret_gpu_batch_ingest_coverage.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        lib.pomfret_gpu_batch_ingest_retag.argtypes = [vp, vp, vp, C.c_uint64, vp]
        lib.pomfret_gpu_batch_add_window.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.pomfret_gpu_batch_add_windows.argtypes = [vp, vp, vp, vp, vp, C.c_uint32]
        lib.pomfret_gpu_batch_submit.argtypes = [vp]
        lib.pomfret_gpu_batch_rewind.argtypes = [vp]
        lib.pomfret_gpu_decode.argtypes = [vp, C.c_uint8, C.c_uint8]
        lib.pomfret_gpu_haptag.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp]
        lib.pomfret_gpu_pileup.argtypes = [vp, C.POINTER(Config)]
        lib.pomfret_gpu_join.argtypes = [vp, C.POINTER(Config)]
        lib.pomfret_gpu_batch_collect.argtypes = [vp, vp, vp, vp]
        lib.pomfret_gpu_batch_collect_haptags.argtypes = [vp, vp, vp]
        lib.pomfret_gpu_batch_end.argtypes = [vp]
        lib.pomfret_gpu_batch_timing.argtypes = [vp, C.POINTER(Timing)]
        lib.pomfret_gpu_debug_read_info.argtypes = [vp, C.c_uint32, u32p, u32p, u32p]
        lib.pomfret_gpu_debug_get_calls.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint32, u32p]
        lib.pomfret_gpu_debug_get_sites.argtypes = [vp, C.c_uint32, C.c_int, vp, vp, vp, C.c_uint32, u32p]
        lib.pomfret_gpu_debug_get_mmrs.argtypes = [vp, C.c_uint32, C.c_int, vp, C.c_uint32, u32p, u32p]
        lib.pomfret_gpu_debug_get_tags.argtypes = [vp, C.c_int, vp]
        lib.pomfret_gpu_debug_get_tag_order.argtypes = [vp, C.c_uint32, C.c_int, vp, C.c_uint32, u32p]

    def strerror(self, rc):
        return self.lib.pomfret_gpu_strerror(rc).decode()

    def check(self, rc, where):
        if rc != 0:
            raise GpuError(self, rc, where)

    def device_count(self):
        return self.lib.pomfret_gpu_device_count()

    def init(self, devices=None, n_workers=1):
        ctx = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.pomfret_gpu_init(C.byref(ctx), arr, len(devices), n_workers)
        else:
            rc = self.lib.pomfret_gpu_init(C.byref(ctx), None, 0, n_workers)
        self.check(rc, "pomfret_gpu_init")
        return ctx

    def host_register(self, ctx, ptr, nbytes):
        self.check(self.lib.pomfret_gpu_host_register(ctx, ptr, nbytes), "host_register")

    def host_unregister(self, ctx, ptr):
        self.check(self.lib.pomfret_gpu_host_unregister(ctx, ptr), "host_unregister")

    def destroy(self, ctx):
        self.lib.pomfret_gpu_destroy(ctx)

    def batch_begin(self, ctx, worker=0, device=0):
        b = C.c_void_p()
        self.check(self.lib.pomfret_gpu_batch_begin(ctx, worker, device, C.byref(b)), "batch_begin")
        return Batch(self, b)


class Batch:
    """One pomfret_gpu_batch; methods mirror the C entry points one to one."""

    def __init__(self, gpu, handle):
        self.gpu = gpu
        self.h = handle
        self.n_reads = 0
        self.n_windows = 0

    def reset(self):
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_reset(self.h), "batch_reset")
        self.n_reads = 0
        self.n_windows = 0

    def add_reads(self, descs_ptr, n):
        """descs_ptr: address of an array of pomfret_gpu_read_desc"""
        base = descs_ptr if isinstance(descs_ptr, int) else C.cast(descs_ptr, C.c_void_p).value
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_add_reads(self.h, base, n), "batch_add_reads")
        first = self.n_reads
        self.n_reads += n
        return first

    def add_reads_shared(self, descs_ptr, n, same_as):
        """same_as: int64 numpy array, >= 0 names the earlier batch read that is the same alignment record"""
        base = descs_ptr if isinstance(descs_ptr, int) else C.cast(descs_ptr, C.c_void_p).value
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_add_reads_shared(self.h, base, n, same_as.ctypes.data if same_as is not None else None),
                       "batch_add_reads_shared")
        first = self.n_reads
        self.n_reads += n
        return first

    def ingest_bgzf(self, comp_ptr, comp_bytes, blocks_ptr, n_blocks, streams_ptr, n_streams, flt=None, check=True):
        """compressed ingest; returns (rc, records) with records a ctypes array of SlicedRecord"""
        n = C.c_uint32()
        rc = self.gpu.lib.pomfret_gpu_batch_ingest_bgzf(self.h, comp_ptr, comp_bytes, blocks_ptr, n_blocks, streams_ptr, n_streams,
                                                        C.byref(flt) if flt is not None else None, C.byref(n))
        if check:
            self.gpu.check(rc, "ingest_bgzf")
        recs = (SlicedRecord * max(n.value, 1))()
        if rc == 0:
            self.gpu.check(self.gpu.lib.pomfret_gpu_batch_ingest_records(self.h, recs, n.value), "ingest_records")
        return rc, recs, n.value

    def ingest_coverage(self, min_pos, bin_size, n_bins):
        """bin increments of the ingest's kept records (estimate_read_coverage_dirtyfast, blockjoin.c:1016-1021)"""
        v = C.c_uint64()
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_ingest_coverage(self.h, min_pos, bin_size, n_bins, C.byref(v)), "ingest_coverage")
        return v.value

    def ingest_retag(self, dst_off, hp_val, out_bytes):
        """re-tagged uncompressed BAM stream of the ingest's records (output_modify_bam, blockjoin.c:3022-3103)"""
        out = np.zeros(max(int(out_bytes), 1), dtype=np.uint8)
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_ingest_retag(self.h, dst_off.ctypes.data, hp_val.ctypes.data, int(out_bytes), out.ctypes.data),
                       "ingest_retag")
        return out[:int(out_bytes)]

    def inflated(self, stream, cap=1 << 28):
        n = C.c_size_t()
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_get_inflated(self.h, stream, None, 0, C.byref(n)), "debug_get_inflated")
        buf = np.zeros(max(n.value, 1), dtype=np.uint8)
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_get_inflated(self.h, stream, buf.ctypes.data, n.value, C.byref(n)), "debug_get_inflated")
        return buf[:n.value]

    def add_reads_device(self, descs_ptr, n, same_as=None):
        base = descs_ptr if isinstance(descs_ptr, int) else C.cast(descs_ptr, C.c_void_p).value
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_add_reads_device(self.h, base, n, same_as.ctypes.data if same_as is not None else None),
                       "batch_add_reads_device")
        first = self.n_reads
        self.n_reads += n
        return first

    def add_window(self, ref_start, ref_end, first_read, n_reads):
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_add_window(self.h, ref_start, ref_end, first_read, n_reads),
                       "batch_add_window")
        self.n_windows += 1

    def add_windows(self, ref_start, ref_end, first_read, n_reads):
        """parallel uint32 numpy arrays, one entry per window"""
        n = len(ref_start)
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_add_windows(self.h, ref_start.ctypes.data, ref_end.ctypes.data,
                                                                  first_read.ctypes.data, n_reads.ctypes.data, n),
                       "batch_add_windows")
        self.n_windows += n

    def submit(self):
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_submit(self.h), "batch_submit")

    def rewind(self):
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_rewind(self.h), "batch_rewind")

    def decode(self, lo, hi):
        self.gpu.check(self.gpu.lib.pomfret_gpu_decode(self.h, lo, hi), "decode")

    def pileup(self, cfg):
        self.gpu.check(self.gpu.lib.pomfret_gpu_pileup(self.h, C.byref(cfg)), "pileup")

    def join(self, cfg):
        self.gpu.check(self.gpu.lib.pomfret_gpu_join(self.h, C.byref(cfg)), "join")

    def collect(self, check=True):
        res = (WindowResult * max(self.n_windows, 1))()
        tags = np.full(max(self.n_reads, 1), 255, dtype=np.uint8)
        ids = np.full(max(self.n_reads, 1), -1, dtype=np.int32)
        rc = self.gpu.lib.pomfret_gpu_batch_collect(self.h, res, tags.ctypes.data, ids.ctypes.data)
        if check:
            self.gpu.check(rc, "batch_collect")
        return list(res)[:self.n_windows], tags[:self.n_reads], ids[:self.n_reads], rc

    def haptag(self, variants, n_variants, bases, known_first):
        """variants: numpy byte buffer holding n_variants pomfret_gpu_variant records"""
        self.gpu.check(self.gpu.lib.pomfret_gpu_haptag(self.h, variants.ctypes.data if n_variants else None,
                                                       n_variants, bases.ctypes.data if len(bases) else None,
                                                       len(bases), known_first.ctypes.data), "haptag")

    def collect_haptags(self):
        tags = np.zeros(max(self.n_reads, 1), dtype=np.uint8)
        status = np.zeros(max(self.n_reads, 1), dtype=np.int32)
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_collect_haptags(self.h, tags.ctypes.data, status.ctypes.data),
                       "collect_haptags")
        return tags[:self.n_reads], status[:self.n_reads]

    def timing(self):
        t = Timing()
        self.gpu.check(self.gpu.lib.pomfret_gpu_batch_timing(self.h, C.byref(t)), "batch_timing")
        return t

    def end(self):
        if self.h:
            self.gpu.lib.pomfret_gpu_batch_end(self.h)
            self.h = None

    # ---- parity getters ----
    def read_info(self, i):
        st, nc, end = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_read_info(self.h, i, C.byref(st), C.byref(nc), C.byref(end)),
                       "debug_read_info")
        return st.value, nc.value, end.value

    def calls(self, i, cap=1 << 16):
        pos = np.zeros(cap, dtype=np.uint32)
        cat = np.zeros(cap, dtype=np.uint8)
        n = C.c_uint32()
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_get_calls(self.h, i, pos.ctypes.data, cat.ctypes.data, cap,
                                                               C.byref(n)), "debug_get_calls")
        if n.value > cap:
            return self.calls(i, n.value)
        return pos[:n.value].copy(), cat[:n.value].copy()

    def sites(self, w, direction, cap=1 << 16):
        pos = np.zeros(cap, dtype=np.uint32)
        st = np.zeros(cap, dtype=np.uint32)
        ln = np.zeros(cap, dtype=np.uint8)
        n = C.c_uint32()
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_get_sites(self.h, w, direction, pos.ctypes.data, st.ctypes.data,
                                                               ln.ctypes.data, cap, C.byref(n)), "debug_get_sites")
        if n.value > cap:
            return self.sites(w, direction, n.value)
        return pos[:n.value].copy(), st[:n.value].copy(), ln[:n.value].copy()

    def mmrs(self, i, direction, cap=1 << 14):
        m = np.zeros(cap, dtype=np.uint32)
        n, st = C.c_uint32(), C.c_uint32()
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_get_mmrs(self.h, i, direction, m.ctypes.data, cap, C.byref(n),
                                                              C.byref(st)), "debug_get_mmrs")
        if n.value > cap:
            return self.mmrs(i, direction, n.value)
        return m[:n.value].copy(), st.value

    def prop_tags(self, direction):
        t = np.zeros(max(self.n_reads, 1), dtype=np.uint8)
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_get_tags(self.h, direction, t.ctypes.data), "debug_get_tags")
        return t[:self.n_reads]

    def tag_order(self, w, direction, cap=1 << 16):
        ids = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32()
        self.gpu.check(self.gpu.lib.pomfret_gpu_debug_get_tag_order(self.h, w, direction, ids.ctypes.data, cap,
                                                                   C.byref(n)), "debug_get_tag_order")
        return ids[:n.value].copy()


class HostLib:
    """libpomfret_host.so: BAM window loader (the htslib half of load_reads_given_interval) and friends."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError("%s not found — run __graft_entry__.build()" % path)
        lib = C.CDLL(path)
        self.lib = lib
        vp = C.c_void_p
        lib.pomfret_host_bam_open.restype = vp
        lib.pomfret_host_bam_open.argtypes = [C.c_char_p]
        lib.pomfret_host_bam_close.argtypes = [vp]
        lib.pomfret_host_window_load.restype = vp
        lib.pomfret_host_window_load.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                                 C.POINTER(C.c_int)]
        lib.pomfret_host_window_n.argtypes = [vp]
        lib.pomfret_host_window_descs.restype = vp
        lib.pomfret_host_window_descs.argtypes = [vp]
        lib.pomfret_host_window_qname.restype = C.c_char_p
        lib.pomfret_host_window_qname.argtypes = [vp, C.c_int]
        lib.pomfret_host_window_bases.restype = C.c_uint64
        lib.pomfret_host_window_bases.argtypes = [vp]
        lib.pomfret_host_window_arena.restype = vp
        lib.pomfret_host_window_arena.argtypes = [vp, C.POINTER(C.c_uint64)]
        lib.pomfret_host_window_free.argtypes = [vp]

    def ingest_plan(self, bam, chrom, regions):
        """regions: list of (beg0, end0).  Returns a dict with the compressed bytes and the block / stream tables."""
        lib = self.lib
        vp = C.c_void_p
        lib.pomfret_host_ingest_plan.restype = vp
        lib.pomfret_host_ingest_plan.argtypes = [vp, C.c_char_p, vp, C.c_int]
        for f, rt in (("comp", vp), ("blocks", vp), ("streams", vp)):
            getattr(lib, "pomfret_host_ingest_" + f).restype = rt
        lib.pomfret_host_ingest_comp.argtypes = [vp, C.POINTER(C.c_uint64)]
        lib.pomfret_host_ingest_blocks.argtypes = [vp, C.POINTER(C.c_uint32)]
        lib.pomfret_host_ingest_streams.argtypes = [vp, C.POINTER(C.c_uint32)]
        lib.pomfret_host_ingest_stream_run.restype = C.c_uint32
        lib.pomfret_host_ingest_stream_run.argtypes = [vp, C.c_uint32]
        lib.pomfret_host_ingest_free.argtypes = [vp]
        arr = np.array(regions, dtype=np.int64).reshape(-1)
        h = lib.pomfret_host_ingest_plan(bam, chrom.encode(), arr.ctypes.data, len(regions))
        if not h:
            raise RuntimeError("ingest plan failed")
        n64, nb, ns = C.c_uint64(), C.c_uint32(), C.c_uint32()
        comp = lib.pomfret_host_ingest_comp(h, C.byref(n64))
        blocks = lib.pomfret_host_ingest_blocks(h, C.byref(nb))
        streams = lib.pomfret_host_ingest_streams(h, C.byref(ns))
        return dict(handle=h, comp=comp, comp_bytes=n64.value, blocks=blocks, n_blocks=nb.value, streams=streams, n_streams=ns.value,
                    stream_run=[lib.pomfret_host_ingest_stream_run(h, s) for s in range(ns.value)])

    def ingest_free(self, plan):
        self.lib.pomfret_host_ingest_free(plan["handle"])

    def bam_open(self, path):
        h = self.lib.pomfret_host_bam_open(path.encode())
        if not h:
            raise IOError("cannot open BAM %s" % path)
        return h

    def bam_close(self, h):
        self.lib.pomfret_host_bam_close(h)

    def window_load(self, bam, chrom, start, end, readlen_threshold, min_mapq):
        rc = C.c_int()
        w = self.lib.pomfret_host_window_load(bam, chrom.encode(), start, end, readlen_threshold, min_mapq,
                                              C.byref(rc))
        if not w:
            raise RuntimeError("window load failed: %d" % rc.value)
        return w

    def window_n(self, w):
        return self.lib.pomfret_host_window_n(w)

    def window_descs(self, w):
        return self.lib.pomfret_host_window_descs(w)

    def window_qnames(self, w):
        return [self.lib.pomfret_host_window_qname(w, i).decode() for i in range(self.window_n(w))]

    def window_bases(self, w):
        return self.lib.pomfret_host_window_bases(w)

    def window_arena(self, w):
        n = C.c_uint64()
        p = self.lib.pomfret_host_window_arena(w, C.byref(n))
        return p, n.value

    def window_free(self, w):
        self.lib.pomfret_host_window_free(w)


_gpu = {}
_host = None


def load_gpu(path=None):
    """Load libpomfret_gpu.so (the nvcc sm_100a build).  `path` is for tests that load another build."""
    path = path or os.environ.get("POMFRET_GPU_LIB") or os.path.join(LIB_DIR, "libpomfret_gpu.so")
    if path not in _gpu:
        _gpu[path] = GpuLib(path)
    return _gpu[path]


def load_host(path=None):
    global _host
    if path:
        return HostLib(path)
    if _host is None:
        _host = HostLib(os.path.join(LIB_DIR, "libpomfret_host.so"))
    return _host
