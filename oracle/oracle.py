"""ctypes front end of oracle/liboracle.so and helpers to run the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under bgt_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_BGT = os.path.join(REF_DIR, "bgt")
REF_PBFVIEW = os.path.join(REF_DIR, "pbfview")
REF_LIB = os.path.join(REF_DIR, "libbgtref.so")
MKSITES = os.path.join(REF_DIR, "mksites")

_lib = None


def build():
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    u8p, i32p, u32p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_void_p
    L.orc_rle_len.restype = C.c_uint32
    L.orc_rle_len.argtypes = [C.c_uint8]
    L.orc_rle_encode.argtypes = [C.c_int, vp, vp]
    L.orc_pbf_open_mem.restype = vp
    L.orc_pbf_open_mem.argtypes = [vp, C.c_size_t]
    L.orc_pbf_close.argtypes = [vp]
    for f in ("orc_pbf_m", "orc_pbf_g", "orc_pbf_shift"):
        getattr(L, f).argtypes = [vp]
    L.orc_pbf_n.restype = C.c_int64
    L.orc_pbf_n.argtypes = [vp]
    L.orc_pbf_subset.argtypes = [vp, C.c_int, vp]
    L.orc_pbf_seek.argtypes = [vp, C.c_int64]
    L.orc_pbf_read.restype = C.POINTER(u8p)
    L.orc_pbf_read.argtypes = [vp]
    L.orc_pbf_row_bytes.restype = C.c_int64
    L.orc_pbf_row_bytes.argtypes = [vp, C.c_int64, C.c_int64, C.c_int]
    L.orc_pbfw_new.restype = vp
    L.orc_pbfw_new.argtypes = [C.c_int, C.c_int, C.c_int]
    L.orc_pbfw_write.argtypes = [vp, vp]
    L.orc_pbfw_write_rle.argtypes = [vp, vp, vp]
    L.orc_pbfw_finish.restype = C.c_size_t
    L.orc_pbfw_finish.argtypes = [vp, C.POINTER(vp)]
    L.orc_pbfw_free.argtypes = [vp]
    L.orc_expr_parse.restype = vp
    L.orc_expr_parse.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.orc_expr_free.argtypes = [vp]
    L.orc_expr_set_int.argtypes = [vp, C.c_char_p, C.c_int64]
    L.orc_expr_eval.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.orc_scan.restype = C.c_int64
    L.orc_scan.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, vp, vp, C.c_int, C.c_char_p, vp, vp, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------- PBF encode / decode

def encode_pbf(mat, shift=13):
    """mat: uint8 [rows, m] of 2-bit codes -> bytes of a .pbf written by the restated encoder."""
    L = lib()
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    n, m = mat.shape
    w = L.orc_pbfw_new(m, 2, shift)
    p0 = np.empty(m, np.uint8)
    p1 = np.empty(m, np.uint8)
    arr = (C.c_void_p * 2)(_ptr(p0), _ptr(p1))
    for k in range(n):
        np.bitwise_and(mat[k], 1, out=p0)
        np.right_shift(mat[k], 1, out=p1)
        p1 &= 1
        L.orc_pbfw_write(w, arr)
    buf = C.c_void_p()
    ln = L.orc_pbfw_finish(w, C.byref(buf))
    out = C.string_at(buf, ln)
    L.orc_pbfw_free(w)
    return out


def encode_pbf_rle(m, rows, shift=13):
    """rows: list of (rle_plane0: bytes, rle_plane1: bytes); raw RLE strings, possibly non-canonical."""
    L = lib()
    w = L.orc_pbfw_new(m, 2, shift)
    for r0, r1 in rows:
        b0 = C.create_string_buffer(r0, len(r0) + 1)
        b1 = C.create_string_buffer(r1, len(r1) + 1)
        arr = (C.c_void_p * 2)(C.cast(b0, C.c_void_p), C.cast(b1, C.c_void_p))
        ls = (C.c_int32 * 2)(len(r0), len(r1))
        L.orc_pbfw_write_rle(w, arr, ls)
    buf = C.c_void_p()
    ln = L.orc_pbfw_finish(w, C.byref(buf))
    out = C.string_at(buf, ln)
    L.orc_pbfw_free(w)
    return out


class Pbf:
    """Memory-resident PBF reader (restated pbf_open_r/pbf_read/pbf_seek/pbf_subset)."""

    def __init__(self, data):
        self._buf = np.frombuffer(data, dtype=np.uint8)
        self._h = lib().orc_pbf_open_mem(_ptr(self._buf), self._buf.size)
        if not self._h:
            raise ValueError("not a PBF")
        self.m = lib().orc_pbf_m(self._h)
        self.g = lib().orc_pbf_g(self._h)
        self.shift = lib().orc_pbf_shift(self._h)
        self.n = lib().orc_pbf_n(self._h)
        self.n_sub = 0

    def close(self):
        if self._h:
            lib().orc_pbf_close(self._h)
            self._h = None

    def subset(self, cols):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        self.n_sub = cols.size if 0 < cols.size < self.m else 0
        return lib().orc_pbf_subset(self._h, cols.size, _ptr(cols))

    def seek(self, k):
        return lib().orc_pbf_seek(self._h, k)

    def read(self):
        r = lib().orc_pbf_read(self._h)
        if not r:
            return None
        w = self.n_sub if self.n_sub else self.m
        return [np.ctypeslib.as_array(r[g], shape=(w,)).copy() for g in range(self.g)]

    def row_bytes(self, beg, end, with_snapshots=True):
        return lib().orc_pbf_row_bytes(self._h, beg, end, int(with_snapshots))

    def scan(self, row_beg, n_rows, out_samples=None, group=None, n_groups=1, flt=None, want_hap=False):
        """Restated `bgt view` scan. Returns dict(counts[n,3+3G], passed[n], hap0, hap1)."""
        n_samples = self.m // 2
        if out_samples is None:
            out_samples = np.arange(n_samples, dtype=np.int32)
        out_samples = np.ascontiguousarray(out_samples, dtype=np.int32)
        n_out = out_samples.size
        if group is None:
            group = np.ones(n_out, dtype=np.uint32)
        group = np.ascontiguousarray(group, dtype=np.uint32)
        counts = np.zeros((n_rows, 3 + 3 * n_groups), dtype=np.int32)
        passed = np.zeros(n_rows, dtype=np.uint8)
        hap0 = np.zeros((n_rows, 2 * n_out), dtype=np.uint8) if want_hap else None
        hap1 = np.zeros((n_rows, 2 * n_out), dtype=np.uint8) if want_hap else None
        done = lib().orc_scan(self._h, row_beg, n_rows, n_out, _ptr(out_samples), _ptr(group), n_groups,
                              flt.encode() if flt is not None else None,
                              _ptr(counts), _ptr(passed), _ptr(hap0), _ptr(hap1))
        if done < 0:
            raise ValueError("orc_scan failed: %d" % done)
        return dict(n=done, counts=counts[:done], passed=passed[:done],
                    hap0=None if hap0 is None else hap0[:done], hap1=None if hap1 is None else hap1[:done])


def decode_all(data, cols=None, row_beg=0, n_rows=None):
    """Decode rows to a [rows, width] uint8 matrix of 2-bit codes with the restated reader."""
    p = Pbf(data)
    try:
        if row_beg:
            p.seek(row_beg)
        if cols is not None:
            p.subset(cols)
        out = []
        while n_rows is None or len(out) < n_rows:
            r = p.read()
            if r is None:
                break
            out.append(r[0] | (r[1] << 1))
        w = (len(cols) if cols is not None and 0 < len(cols) < p.m else p.m)
        return np.array(out, dtype=np.uint8).reshape(len(out), w)
    finally:
        p.close()


# ----------------------------------------------------------------------------- filter expression

class Expr:
    def __init__(self, s):
        err = C.c_int(0)
        self._h = lib().orc_expr_parse(s.encode(), C.byref(err))
        self.err = err.value
        if not self._h:
            raise ValueError("parse error 0x%x" % self.err)

    def set_int(self, name, v):
        return lib().orc_expr_set_int(self._h, name.encode(), int(v))

    def eval(self):
        iv, rv, vt = C.c_int64(0), C.c_double(0), C.c_int(0)
        err = lib().orc_expr_eval(self._h, C.byref(iv), C.byref(rv), C.byref(vt))
        return err, iv.value, rv.value, vt.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_expr_free(self._h)
            self._h = None


# ----------------------------------------------------------------------------- the compiled reference

def have_ref():
    return os.path.exists(REF_BGT) and os.path.exists(REF_PBFVIEW)


def ref_run(argv, stdin=None, cwd=None, seekable_stdout=False):
    """Run a reference binary (oracle/_ref/<argv[0]>) and return stdout bytes.

    seekable_stdout: send stdout to a regular file -- `pbfview -b` records ftell() offsets in the PBF index
    (pbwt.c:270,297), which are -1 on a pipe."""
    exe = os.path.join(REF_DIR, argv[0])
    if seekable_stdout:
        import tempfile
        with tempfile.NamedTemporaryFile() as tf:
            r = subprocess.run([exe] + list(argv[1:]), input=stdin, stdout=tf, stderr=subprocess.PIPE, cwd=cwd)
            tf.seek(0)
            out = tf.read()
    else:
        r = subprocess.run([exe] + list(argv[1:]), input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=cwd)
        out = r.stdout
    if r.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (argv, r.returncode, r.stderr.decode()[-400:]))
    return out


def pim_text(mat, g=2):
    mat = np.asarray(mat)
    lines = ["PIM1 %d %d" % (mat.shape[1], g)]
    lines += [" ".join(str(int(x)) for x in row) for row in mat]
    return ("\n".join(lines) + "\n").encode()


def parse_pim(txt):
    lines = txt.decode().strip().split("\n")
    hdr = lines[0].split()
    m = int(hdr[1])
    if len(lines) == 1:
        return np.zeros((0, m), np.uint8)
    return np.array([[int(x) for x in ln.split()] for ln in lines[1:]], dtype=np.uint8).reshape(-1, m)


def vcf_text(mat, sample_names=None, chrom="11", contig_len=135006516, pos0=1000, step=10, multi=None):
    """A phased/unphased-agnostic VCF whose import (-S off: every row is biallelic or has <M>) yields `mat`.
    Codes: 0 REF, 1 ALT, 2 missing, 3 second ALT.  Rows containing code 3 are written with two ALTs."""
    mat = np.asarray(mat)
    n, m = mat.shape
    ns = m // 2
    if sample_names is None:
        sample_names = ["S%07d" % i for i in range(ns)]
    out = ["##fileformat=VCFv4.1", '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
           "##contig=<ID=%s,length=%d>" % (chrom, contig_len),
           "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(sample_names)]
    sym = {0: "0", 1: "1", 2: ".", 3: "2"}
    bases = "ACGT"
    for k in range(n):
        row = mat[k]
        has3 = bool((row == 3).any()) if multi is None else bool(multi[k])
        ref = bases[k % 4]
        alt = bases[(k + 1) % 4] + ("," + bases[(k + 2) % 4] if has3 else "")
        gts = ["%s|%s" % (sym[int(row[2 * s])], sym[int(row[2 * s + 1])]) for s in range(ns)]
        out.append("%s\t%d\t.\t%s\t%s\t0\t.\t.\tGT\t%s" % (chrom, pos0 + step * k, ref, alt, "\t".join(gts)))
    return ("\n".join(out) + "\n").encode()
