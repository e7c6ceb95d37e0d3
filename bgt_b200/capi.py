"""ctypes binding of include/bgt_b200.h.  No computation happens in this file."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCAN_COUNTS, SCAN_HAP_BITS, SCAN_HAP_BYTES, SCAN_DEVICE_OUT, SCAN_NO_SPLIT = 0x01, 0x02, 0x04, 0x10, 0x20
SCAN_NO_COMPOSE, SCAN_NO_SEGMENTS, SCAN_NO_PIECES = 0x40, 0x80, 0x1000
MAX_GROUPS = 32

_lib = None


class B200Error(RuntimeError):
    pass


def lib_path():
    return os.path.join(HERE, "lib", "libbgt_b200.so")


class ScanOut(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("passed", C.c_void_p), ("hap_bits", C.c_void_p * 2), ("hap_bytes", C.c_void_p * 2),
                ("totals", C.c_int64 * 4)]


class SynthCfg(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_rows", C.c_int64), ("shift", C.c_int32), ("seed", C.c_uint64),
                ("r_max", C.c_int32), ("p1_one_in", C.c_int32), ("p1_max_iv", C.c_int32), ("p1_max_len", C.c_int32)]


# every symbol include/bgt_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _int = C.c_void_p, C.c_int64, C.c_int
SIGNATURES = {
    "b200_abi_version": (_int, []),
    "b200_device_count": (_int, []),
    "b200_strerror": (C.c_char_p, []),
    "b200_errcode": (_int, []),
    "b200_ctx_create": (_vp, [_int]),
    "b200_ctx_destroy": (None, [_vp]),
    "b200_ctx_sync": (_int, [_vp]),
    "b200_host_alloc": (_vp, [C.c_size_t]),
    "b200_host_free": (None, [_vp]),
    "b200_host_register": (_int, [_vp, C.c_size_t]),
    "b200_host_unregister": (_int, [_vp]),
    "b200_pbf_load": (_vp, [_vp, _vp, C.c_size_t, _i64, _i64]),
    "b200_pbf_load_ex": (_vp, [_vp, _vp, C.c_size_t, _i64, _i64, C.c_uint]),
    "b200_pbf_open": (_vp, [_vp, C.c_char_p, _i64, _i64]),
    "b200_pbf_close": (None, [_vp]),
    "b200_pbf_m": (_int, [_vp]),
    "b200_pbf_g": (_int, [_vp]),
    "b200_pbf_shift": (_int, [_vp]),
    "b200_pbf_n": (_i64, [_vp]),
    "b200_pbf_row_beg": (_i64, [_vp]),
    "b200_pbf_row_end": (_i64, [_vp]),
    "b200_pbf_row_bytes": (_i64, [_vp, _i64, _i64, _int]),
    "b200_pbf_bad_rows": (_i64, [_vp]),
    "b200_pbf_split_blocks": (_int, [_vp]),
    "b200_query_create": (_vp, [_vp, _vp, _int, _vp, _vp, _int, C.c_char_p, C.POINTER(_int)]),
    "b200_query_create_m": (_vp, [_vp, _int, _int, _vp, _vp, _int, C.c_char_p, C.POINTER(_int)]),
    "b200_query_create_cols": (_vp, [_vp, _vp, _int, _vp]),
    "b200_query_destroy": (None, [_vp]),
    "b200_query_n_track": (_int, [_vp]),
    "b200_query_filter_needs_host": (_int, [_vp]),
    "b200_query_hap_words": (_int, [_vp]),
    "b200_query_counts_stride": (_int, [_vp]),
    "b200_scan": (_i64, [_vp, _vp, _vp, _i64, _i64, C.c_uint, C.POINTER(ScanOut)]),
    "b200_scan_regions": (_i64, [_vp, _vp, _vp, _int, C.POINTER(_i64), C.POINTER(_i64), C.c_uint, C.POINTER(ScanOut)]),
    "b200_scan_collect": (_int, [_vp, C.POINTER(_i64)]),
    "b200_pbf_peek": (_int, [_vp, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(_i64)]),
    "b200_pbf_load_scan": (_vp, [_vp, _vp, C.c_size_t, _i64, _i64, _vp, C.POINTER(ScanOut), C.POINTER(_i64)]),
    "b200_last_totals": (_int, [_vp, C.POINTER(_i64)]),
    "b200_allreduce_i64": (_int, [C.POINTER(_vp), _int, C.POINTER(_i64), _int]),
    "b200_last_ms": (C.c_double, [_vp, _int]),
    "b200_kernel_launches": (_i64, [_vp]),
    "b200_mark": (_int, [_vp, _int]),
    "b200_mark_elapsed_ms": (C.c_double, [_vp, _int, _int]),
    "b200_flt_eval_host": (_int, [C.c_char_p, _int, _vp, _i64, _vp]),
    "b200_pbf_plan": (_int, [_vp, C.c_size_t, _i64, _i64, C.POINTER(_i64)]),
    "b200_synth_generate": (_vp, [_vp, C.POINTER(SynthCfg)]),
    "b200_pbf_image_size": (C.c_size_t, [_vp]),
    "b200_pbf_image_download": (_int, [_vp, _vp, C.c_size_t]),
    "b200_pbf_image_download_range": (_int, [_vp, _vp, C.c_uint64, C.c_size_t]),
    "b200_pbf_block_bytes": (_int, [_vp, _i64, _i64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b200_enc_create": (_vp, [_vp, _int, _int]),
    "b200_enc_write_bytes": (_int, [_vp, _vp, _vp, _i64]),
    "b200_enc_write_bits": (_int, [_vp, _vp, _i64]),
    "b200_enc_rows": (_i64, [_vp]),
    "b200_enc_drain": (_i64, [_vp, C.POINTER(_vp)]),
    "b200_enc_finish": (_i64, [_vp, C.POINTER(_vp)]),
    "b200_enc_destroy": (None, [_vp]),
    "b200_bgzf_inflate": (_i64, [_vp, _vp, C.c_size_t, _vp, C.c_size_t]),
    "b200_sites_load": (_vp, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _int]),
    "b200_sites_n": (_i64, [_vp]),
    "b200_sites_rows_sorted": (_int, [_vp]),
    "b200_sites_rec_range": (_int, [_vp, _i64, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "b200_sites_header": (C.c_void_p, [_vp, C.POINTER(_i64)]),
    "b200_sites_rows": (_int, [_vp, _vp, _vp]),
    "b200_sites_destroy": (None, [_vp]),
    "b200_view_text": (_i64, [_vp, _vp, _vp, _vp, _int, _vp, _int, C.POINTER(_vp), C.POINTER(_i64)]),
    "b200_view_text_ex": (_i64, [_vp, _vp, _vp, _vp, C.c_uint, _i64, _i64, _vp, _int, C.POINTER(_vp), C.POINTER(_i64)]),
}


def load_library(path=None):
    """dlopen libbgt_b200.so and type every entry point.  Raises if the library was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or lib_path()
    if not os.path.exists(path):
        raise B200Error("libbgt_b200.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                        "there is no CPU fallback" % path)
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def lib():
    return load_library()


def _err():
    return lib().b200_strerror().decode()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One GPU (b200_ctx_t)."""

    def __init__(self, device=0):
        self.h = lib().b200_ctx_create(device)
        if not self.h:
            raise B200Error(_err())
        self.device = device

    def sync(self):
        if lib().b200_ctx_sync(self.h) != 0:
            raise B200Error(_err())

    def last_ms(self, which):
        return lib().b200_last_ms(self.h, which)

    def mark(self, slot):
        if lib().b200_mark(self.h, slot) != 0:
            raise B200Error(_err())

    def mark_elapsed_ms(self, a, b):
        return lib().b200_mark_elapsed_ms(self.h, a, b)

    @property
    def launches(self):
        return lib().b200_kernel_launches(self.h)

    def close(self):
        if self.h:
            lib().b200_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Pbf:
    """A .pbf (or a row shard) resident in HBM (b200_pbf_t); the stand-in for pbf_open_r (pbwt.c:221-262)."""

    def __init__(self, ctx, handle):
        if not handle:
            raise B200Error(_err())
        self.ctx, self.h = ctx, handle
        L = lib()
        self.m, self.g, self.shift, self.n = L.b200_pbf_m(handle), L.b200_pbf_g(handle), L.b200_pbf_shift(handle), L.b200_pbf_n(handle)
        self.row_beg, self.row_end = L.b200_pbf_row_beg(handle), L.b200_pbf_row_end(handle)
        self.bad_rows = L.b200_pbf_bad_rows(handle)

    @classmethod
    def from_bytes(cls, ctx, data, row_beg=0, row_end=-1, prepare_count_scan=False):
        buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        return cls(ctx, lib().b200_pbf_load_ex(ctx.h, _ptr(buf), buf.size, row_beg, row_end, 1 if prepare_count_scan else 0))

    @classmethod
    def open(cls, ctx, fn, row_beg=0, row_end=-1):
        return cls(ctx, lib().b200_pbf_open(ctx.h, os.fsencode(fn), row_beg, row_end))

    def row_bytes(self, beg, end, with_snapshots=True):
        v = lib().b200_pbf_row_bytes(self.h, beg, end, int(with_snapshots))
        if v < 0:
            raise B200Error(_err())
        return v

    def image(self, out=None):
        """Download the complete file image (only when fully resident, e.g. generated cohorts)."""
        n = lib().b200_pbf_image_size(self.h)
        if n == 0:
            raise B200Error("no complete file image resident")
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        if lib().b200_pbf_image_download(self.h, _ptr(out), out.size) != 0:
            raise B200Error(_err())
        return out

    def close(self):
        if self.h:
            lib().b200_pbf_close(self.h)
            self.h = None


class Encoder:
    """The PBWT encoder on the device (b200_enc_t): pbf_open_w / pbf_write / pbf_close (pbwt.c:199-219, 288-311, 264-286)."""

    def __init__(self, ctx, m, shift=13):
        self.h = lib().b200_enc_create(ctx.h, m, shift)
        if not self.h:
            raise B200Error(_err())
        self.m, self.shift = m, shift

    def write(self, a0, a1):
        """rows as pbf_write takes them: two [n_rows][m] uint8 matrices (bit planes 0 and 1, one byte per haplotype)."""
        a0 = np.ascontiguousarray(a0, dtype=np.uint8).reshape(-1, self.m)
        a1 = np.ascontiguousarray(a1, dtype=np.uint8).reshape(-1, self.m)
        if lib().b200_enc_write_bytes(self.h, _ptr(a0), _ptr(a1), a0.shape[0]) != 0:
            raise B200Error(_err())

    def write_bits(self, bits):
        """rows as bit planes in column order: uint32 [n_rows][2][(m+31)//32]."""
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        if lib().b200_enc_write_bits(self.h, _ptr(bits), bits.shape[0]) != 0:
            raise B200Error(_err())

    def drain(self):
        """The file bytes assembled since the last drain (streaming writers, b200_enc_drain)."""
        p = C.c_void_p()
        n = lib().b200_enc_drain(self.h, C.byref(p))
        if n < 0:
            raise B200Error(_err())
        return C.string_at(p, n) if n else b""

    def finish(self):
        """The complete .pbf image as bytes (after drain() calls: the rest of it)."""
        p = C.c_void_p()
        n = lib().b200_enc_finish(self.h, C.byref(p))
        if n < 0:
            raise B200Error(_err())
        return C.string_at(p, n)

    def close(self):
        if self.h:
            lib().b200_enc_destroy(self.h)
            self.h = None


class Sites:
    """The site side of a BGT database on the device (b200_sites_t): .bcf (+ .csi) inflated, indexed and parsed by kernels."""

    def __init__(self, ctx, bcf, csi=None, row_key=-1):
        b = bcf if isinstance(bcf, np.ndarray) else np.frombuffer(bcf, dtype=np.uint8)
        c = None if csi is None else (csi if isinstance(csi, np.ndarray) else np.frombuffer(csi, dtype=np.uint8))
        self.h = lib().b200_sites_load(ctx.h, _ptr(b), b.size, _ptr(c), 0 if c is None else c.size, row_key)
        if not self.h:
            raise B200Error(_err())
        self.ctx, self.n = ctx, lib().b200_sites_n(self.h)

    def header(self):
        n = C.c_int64(0)
        p = lib().b200_sites_header(self.h, C.byref(n))
        return C.string_at(p, n.value)

    def rows(self):
        rows, pos = np.empty(self.n, np.int64), np.empty(self.n, np.int32)
        if lib().b200_sites_rows(self.h, _ptr(rows), _ptr(pos)) != 0:
            raise B200Error(_err())
        return rows, pos

    def view_text(self, pbf, query, with_counts=False, genotypes=False, rec_beg=0, rec_end=-1):
        """The record lines of `bgt view [-G] [-C] [-f ..] [-s ..]` (b200_view_text_ex); returns (bytes, n_lines)."""
        p, nl = C.c_void_p(), C.c_int64(0)
        n = lib().b200_view_text_ex(self.ctx.h, self.h, pbf.h, query.h, (1 if with_counts else 0) | (2 if genotypes else 0), rec_beg, rec_end,
                                    None, 0, C.byref(p), C.byref(nl))
        if n < 0:
            raise B200Error(_err())
        return C.string_at(p, n), nl.value

    def close(self):
        if self.h:
            lib().b200_sites_destroy(self.h)
            self.h = None


def bgzf_inflate(ctx, data):
    """Inflate a BGZF file image on the device (b200_bgzf_inflate); returns bytes."""
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    n = lib().b200_bgzf_inflate(ctx.h, _ptr(buf), buf.size, None, 0)
    if n < 0:
        raise B200Error(_err())
    out = np.empty(max(n, 1), dtype=np.uint8)
    got = lib().b200_bgzf_inflate(ctx.h, _ptr(buf), buf.size, _ptr(out), out.size)
    if got < 0:
        raise B200Error(_err())
    return out[:got].tobytes()


def synth_cohort(ctx, n_samples, n_rows, seed=1, shift=13, r_max=64, p1_one_in=16, p1_max_iv=3, p1_max_len=64):
    """Generate a truthful synthetic cohort on the device and return it resident."""
    cfg = SynthCfg(n_samples, n_rows, shift, seed, r_max, p1_one_in, p1_max_iv, p1_max_len)
    return Pbf(ctx, lib().b200_synth_generate(ctx.h, C.byref(cfg)))


class Query:
    """Sample selection, groups and site filter of one `bgt view` (bgt_prepare + bgtm_add_group + bgtm_set_flt_site)."""

    def __init__(self, ctx, pbf, out_samples=None, group=None, n_groups=1, flt=None):
        self.out_samples = None if out_samples is None else np.ascontiguousarray(out_samples, dtype=np.int32)
        self.group = None if group is None else np.ascontiguousarray(group, dtype=np.uint32)
        m = pbf if isinstance(pbf, int) else pbf.m             # an int: the columns of a PBF that is not resident yet (load_scan)
        n_out = m // 2 if self.out_samples is None else self.out_samples.size
        if self.group is not None and self.group.size != n_out:
            raise ValueError("group must have one entry per selected sample")
        err = C.c_int(0)
        self.h = lib().b200_query_create_m(ctx.h, m, n_out, _ptr(self.out_samples), _ptr(self.group), n_groups,
                                           flt.encode() if flt is not None else None, C.byref(err))
        self.flt_err = err.value
        if not self.h:
            raise B200Error(_err())
        self.n_track = lib().b200_query_n_track(self.h)
        self.words = lib().b200_query_hap_words(self.h)
        self.stride = lib().b200_query_counts_stride(self.h)
        self.n_groups = n_groups

    @classmethod
    def columns(cls, ctx, pbf, cols=None):
        """Column-level selection in list order (pbf_subset, pbwt.c:374-388)."""
        self = cls.__new__(cls)
        self.cols = None if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
        self.h = lib().b200_query_create_cols(ctx.h, pbf.h, 0 if self.cols is None else self.cols.size, _ptr(self.cols))
        if not self.h:
            raise B200Error(_err())
        self.n_track = lib().b200_query_n_track(self.h)
        self.words = lib().b200_query_hap_words(self.h)
        self.stride = lib().b200_query_counts_stride(self.h)
        self.n_groups, self.flt_err = 1, 0
        return self

    def close(self):
        if self.h:
            lib().b200_query_destroy(self.h)
            self.h = None


def scan(ctx, pbf, query, row_beg=0, n_rows=None, counts=True, hap_bits=False, hap_bytes=False, out=None, cols_per_thread=0, no_split=False, no_compose=False, no_segments=False, no_pieces=False):
    """b200_scan with host outputs.  Returns dict(n, counts, passed, hap_bits, hap_bytes, totals)."""
    if n_rows is None:
        n_rows = pbf.row_end - row_beg
    n_alloc = max(int(n_rows), 0)
    res = out or {}
    if counts and "counts" not in res:
        res["counts"] = np.empty((n_alloc, query.stride), dtype=np.int32)
    if "passed" not in res:
        res["passed"] = np.empty(n_alloc, dtype=np.uint8)
    if hap_bits and "hap_bits" not in res:
        res["hap_bits"] = [np.empty((n_alloc, query.words), dtype=np.uint32) for _ in range(2)]
    if hap_bytes and "hap_bytes" not in res:
        res["hap_bytes"] = [np.empty((n_alloc, query.n_track), dtype=np.uint8) for _ in range(2)]
    so = ScanOut()
    so.counts = _ptr(res["counts"]) if counts else None
    so.passed = _ptr(res["passed"])
    flags = (SCAN_COUNTS if counts else 0) | (int(cols_per_thread) << 8) | (SCAN_NO_SPLIT if no_split else 0) | (SCAN_NO_COMPOSE if no_compose else 0) | (SCAN_NO_SEGMENTS if no_segments else 0) | (SCAN_NO_PIECES if no_pieces else 0)
    if hap_bits:
        flags |= SCAN_HAP_BITS
        so.hap_bits[0], so.hap_bits[1] = _ptr(res["hap_bits"][0]), _ptr(res["hap_bits"][1])
    if hap_bytes:
        flags |= SCAN_HAP_BYTES
        so.hap_bytes[0], so.hap_bytes[1] = _ptr(res["hap_bytes"][0]), _ptr(res["hap_bytes"][1])
    done = lib().b200_scan(ctx.h, pbf.h, query.h, row_beg, n_rows, flags, C.byref(so))
    if done < 0:
        raise B200Error(_err())
    res["n"] = done
    res["totals"] = [so.totals[i] for i in range(4)]
    for k in ("counts", "passed"):
        if k in res and res[k] is not None:
            res[k] = res[k][:done]
    return res


def pbf_peek(data):
    """(m, shift, rows) of a .pbf image (b200_pbf_peek)."""
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    m, sh, n = C.c_int32(0), C.c_int32(0), C.c_int64(0)
    if lib().b200_pbf_peek(_ptr(buf), buf.size, C.byref(m), C.byref(sh), C.byref(n)) != 0:
        raise B200Error(_err())
    return m.value, sh.value, n.value


def load_scan(ctx, data, query, row_beg=0, row_end=-1, out=None):
    """b200_pbf_load_scan: host .pbf image -> resident PBF + per-site AC/AN and verdicts in ONE pipeline.  Returns (Pbf, dict)."""
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    m, sh, n = pbf_peek(buf)
    end = n if row_end < 0 or row_end > n else row_end
    rows = max(end - row_beg, 0)
    res = out or {}
    if "counts" not in res:
        res["counts"] = np.empty((rows, query.stride), dtype=np.int32)
    if "passed" not in res:
        res["passed"] = np.empty(rows, dtype=np.uint8)
    so = ScanOut()
    so.counts, so.passed = _ptr(res["counts"]), _ptr(res["passed"])
    done = C.c_int64(0)
    h = lib().b200_pbf_load_scan(ctx.h, _ptr(buf), buf.size, row_beg, row_end, query.h, C.byref(so), C.byref(done))
    if not h:
        raise B200Error(_err())
    res["n"] = done.value
    res["totals"] = [so.totals[i] for i in range(4)]
    return Pbf(ctx, h), res


def scan_regions(ctx, pbf, query, regions, counts=True, hap_bits=False, hap_bytes=False):
    """b200_scan_regions: regions = [(row_beg, n_rows), ...]; outputs are the regions' rows back to back."""
    n = len(regions)
    beg = (C.c_int64 * max(n, 1))(*[int(r[0]) for r in regions])
    cnt = (C.c_int64 * max(n, 1))(*[int(r[1]) for r in regions])
    total = sum(int(r[1]) for r in regions)
    res = {"counts": np.empty((total, query.stride), dtype=np.int32) if counts else None, "passed": np.empty(total, dtype=np.uint8)}
    so = ScanOut()
    so.counts, so.passed = _ptr(res["counts"]), _ptr(res["passed"])
    flags = SCAN_COUNTS if counts else 0
    if hap_bits:
        res["hap_bits"] = [np.empty((total, query.words), dtype=np.uint32) for _ in range(2)]
        flags |= SCAN_HAP_BITS
        so.hap_bits[0], so.hap_bits[1] = _ptr(res["hap_bits"][0]), _ptr(res["hap_bits"][1])
    if hap_bytes:
        res["hap_bytes"] = [np.empty((total, query.n_track), dtype=np.uint8) for _ in range(2)]
        flags |= SCAN_HAP_BYTES
        so.hap_bytes[0], so.hap_bytes[1] = _ptr(res["hap_bytes"][0]), _ptr(res["hap_bytes"][1])
    done = lib().b200_scan_regions(ctx.h, pbf.h, query.h, n, beg, cnt, flags, C.byref(so))
    if done < 0:
        raise B200Error(_err())
    res["n"] = done
    res["totals"] = [so.totals[i] for i in range(4)]
    return res


def scan_device(ctx, pbf, query, row_beg, n_rows, d_counts=0, d_pass=0, d_hap_bits=(0, 0), no_split=False):
    """b200_scan with B200_SCAN_DEVICE_OUT: outputs are device pointers (ints), the call returns without syncing.
    Follow with collect(ctx) to wait, check device error flags and fetch totals / kernel timings."""
    so = ScanOut()
    so.counts = d_counts or None
    so.passed = d_pass or None
    flags = SCAN_DEVICE_OUT | (SCAN_COUNTS if d_counts else 0) | (SCAN_NO_SPLIT if no_split else 0)
    if d_hap_bits[0]:
        flags |= SCAN_HAP_BITS
        so.hap_bits[0], so.hap_bits[1] = d_hap_bits
    done = lib().b200_scan(ctx.h, pbf.h, query.h, row_beg, n_rows, flags, C.byref(so))
    if done < 0:
        raise B200Error(_err())
    return done


def collect(ctx):
    tot = (C.c_int64 * 4)()
    if lib().b200_scan_collect(ctx.h, tot) != 0:
        raise B200Error(_err())
    return [tot[i] for i in range(4)]


def host_alloc(n_bytes):
    """Pinned host memory as a uint8 numpy array (b200_host_alloc); keep the returned array alive, free with host_free."""
    p = lib().b200_host_alloc(n_bytes)
    if not p:
        raise B200Error(_err())
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n_bytes,))


def host_free(arr):
    lib().b200_host_free(arr.ctypes.data_as(C.c_void_p))


def flt_eval_host(flt, n_groups, counts):
    """Host evaluation of a site filter with the library's own compiler + byte-code (b200_flt_eval_host)."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    out = np.zeros(counts.shape[0], dtype=np.uint8)
    err = lib().b200_flt_eval_host(flt.encode(), n_groups, _ptr(counts), counts.shape[0], _ptr(out))
    if err:
        raise B200Error("filter parse error 0x%x" % err if err > 0 else _err())
    return out


def pbf_plan(data, row_beg=0, row_end=-1):
    """Host-side index walk + tile plan of a .pbf image (no device needed)."""
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    info = (C.c_int64 * 8)()
    if lib().b200_pbf_plan(_ptr(buf), buf.size, row_beg, row_end, info) != 0:
        raise B200Error(_err())
    keys = ("m", "shift", "n", "blocks", "tiles", "big_tiles", "max_tile_bytes", "max_row_bytes")
    return dict(zip(keys, [info[i] for i in range(8)]))
