"""CPU tests (-m "not gpu"): the C ABI surface, the host logic of the library (filter compiler/evaluator, PBF index
walk and tile plan), the loud failure without a device, and the multi-rank sharding logic over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from cohorts import haplo_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FILTERS = ["AC>0", "AN>0&&AC/AN>.05", "AC1/AN1>0.1&&AC2==0", "AC1/AN1>=0.1&&AC2==0", "AC/AN", "AC3>0", "AC*2+1-AN%7",
           "AC//3==AN>>2", "-AC+AN", "!AC", "~AC&255", "AC**2>AN", "abs(AC-AN)>3", "log(AC)>0", "(AC+1)*(AN-1)/3.0>=AC",
           "AC==AN||AC<2", "AC<>AN", "AC1+AC2==AC", "AN1/AN2", "0x10+010+AC", "1e2<AN", "AC^AN|3", "AC<<2>=AN",
           "'a'=='a'&&AC", "\"ab\"<\"b\"", "AC>=1&&AC<=2||AN==0", "AC/0", "AC/AN>1e-3", "2**3**2==512", "AC-(-AC)", "+AC>0",
           "AC>0 && AN >= 4", "AC1/AN1-AC2/AN2>0.05", "AC*1.5>AN", "AC%4==3", ".5*AN>AC", "-'a'=='a'", "AN10>0", "AC01>0",
           "abs(AC1-AC2)>=2", "abs(1,2)", "AC>0&&foo(AC)", "AN2-AC2<3||AC1>AN1/2", "AC/AN>=0.5==1"]
BAD = ["AC>", "(AC>0", "AC>0)", "AC=0", "AC#1", "'abc", "f(1,)", ",1", "AC,AN"]


def header_functions(path, prefix):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w*)\s*\(" % prefix, txt)))


def test_abi_exports_every_declared_symbol():
    import bgt_b200
    from bgt_b200 import capi
    L = bgt_b200.load_library()
    names = header_functions(os.path.join(ROOT, "include", "bgt_b200.h"), "b200_")
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libbgt_b200.so does not export %s" % n
        assert n in capi.SIGNATURES, "python binding does not type %s" % n
    assert L.b200_abi_version() == 2


def test_python_binding_flags_match_the_header():
    """The scan flags of the ctypes binding are the header's (include/bgt_b200.h), bit for bit, and do not overlap."""
    from bgt_b200 import capi
    txt = open(os.path.join(ROOT, "include", "bgt_b200.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"#define\s+B200_(SCAN_[A-Z_]+)\s+(0x[0-9a-fA-F]+|\d+)\b", txt)}
    assert {"SCAN_COUNTS", "SCAN_HAP_BITS", "SCAN_HAP_BYTES", "SCAN_DEVICE_OUT", "SCAN_NO_SPLIT", "SCAN_NO_COMPOSE", "SCAN_NO_SEGMENTS", "SCAN_NO_PIECES"} <= set(defs)
    for name, val in defs.items():
        assert getattr(capi, name) == val, name
    vals = list(defs.values())
    assert all(v and v & (v - 1) == 0 for v in vals) and len(set(vals)) == len(vals)   # single, distinct bits
    assert all(v & 0xf00 == 0 for v in vals)                                           # bits 8-11 carry B200_SCAN_COLS_PER_THREAD


def test_seam_a_library_exports_pbwt_api():
    path = os.path.join(ROOT, "bgt_b200", "lib", "libpbwt_b200.so")
    if not os.path.exists(path):
        pytest.skip("libpbwt_b200.so not built")
    L = C.CDLL(path)
    for n in header_functions(os.path.join(ROOT, "include", "pbwt_b200.h"), "pbf_"):
        assert hasattr(L, n), n
    L.pbf_open_r.restype = C.c_void_p
    assert not L.pbf_open_r(b"/nonexistent/file.pbf")            # NULL on open failure, pbwt.c:228-229
    bad = os.path.join(ROOT, "README.md")
    assert not L.pbf_open_r(bad.encode())                         # NULL on bad magic, pbwt.c:232-235


def test_fails_loudly_without_a_device():
    import bgt_b200
    if bgt_b200.load_library().b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(bgt_b200.B200Error, match="no CPU fallback"):
        bgt_b200.Context(0)


def test_product_never_touches_the_oracle():
    bad = []
    for top in ("bgt_b200", "include", "integration"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, top)):
            if "_build" in dp or "__pycache__" in dp:
                continue
            for f in fn:
                if f.endswith((".py", ".c", ".h", ".cu", ".cuh", ".cpp")) or f == "Makefile":
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle[/.]|liboracle|from oracle|import oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, "product files reference oracle/: %s" % bad


def test_filter_compiler_matches_oracle(oracle):
    import bgt_b200
    rng = np.random.default_rng(17)
    rows = [[0, 0, 0, 0, 0, 0, 0, 0, 0]]
    for _ in range(300):
        an1, an2 = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        a1, a2 = int(rng.integers(0, an1 + 1)), int(rng.integers(0, an2 + 1))
        rows.append([an1 + an2, a1 + a2, 0, an1, a1, 0, an2, a2, 0])
    counts = np.array(rows, dtype=np.int32)
    for flt in FILTERS:
        got = bgt_b200.flt_eval_host(flt, 2, counts)
        e = oracle.Expr(flt)
        want = []
        for r in rows:
            for k, v in (("AN", r[0]), ("AC", r[1]), ("AN1", r[3]), ("AC1", r[4]), ("AN2", r[6]), ("AC2", r[7])):
                e.set_int(k, v)
            err, iv, rv, vt = e.eval()
            want.append(0 if err else int(iv != 0))
        assert got.tolist() == want, flt
    for flt in BAD:
        err = C.c_int(0)
        assert not oracle.lib().orc_expr_parse(flt.encode(), C.byref(err))
        got = bgt_b200.lib().b200_flt_eval_host(flt.encode(), 2, None, 0, None)
        assert got == err.value, (flt, got, err.value)


def test_filter_compiler_matches_reference_kexpr(ref):
    """The same verdicts straight from the UNMODIFIED reference's kexpr + bgtm_pass_site_flt semantics."""
    import bgt_b200
    L = C.CDLL(ref.REF_LIB)
    L.ke_parse.restype = C.c_void_p
    L.ke_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.ke_eval_int.restype = C.c_int64
    L.ke_eval_int.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.ke_destroy.argtypes = [C.c_void_p]
    rng = np.random.default_rng(3)
    rows = [[an1 + an2, a1 + a2, 0, an1, a1, 0, an2, a2, 0] for an1, an2, a1, a2 in
            [(int(x), int(y), int(rng.integers(0, x + 1)), int(rng.integers(0, y + 1))) for x, y in rng.integers(0, 40, size=(120, 2))]]
    counts = np.array(rows, dtype=np.int32)
    for flt in FILTERS:
        if "%" in flt or "//" in flt:
            continue  # the reference divides unguarded (SIGFPE on 0)
        err = C.c_int(0)
        ke = L.ke_parse(flt.encode(), C.byref(err))
        assert ke and err.value == 0
        want = []
        for r in rows:
            for k, v in ((b"AN", r[0]), (b"AC", r[1]), (b"AN1", r[3]), (b"AC1", r[4]), (b"AN2", r[6]), (b"AC2", r[7])):
                L.ke_set_int(ke, k, v)
            e = C.c_int(0)
            iv = L.ke_eval_int(ke, C.byref(e))
            want.append(0 if e.value else int(iv != 0))       # bgt.c:717-718
        L.ke_destroy(ke)
        assert bgt_b200.flt_eval_host(flt, 2, counts).tolist() == want, flt


def test_index_walk_and_tile_plan(oracle):
    import bgt_b200
    mat = haplo_matrix(700, 90, 3)
    pbf = oracle.encode_pbf(mat, shift=5)
    info = bgt_b200.pbf_plan(pbf)
    assert (info["m"], info["shift"], info["n"], info["blocks"]) == (90, 5, 700, 22)
    assert info["big_tiles"] == 0 and 0 < info["max_tile_bytes"] <= 8192 and info["tiles"] >= 22
    assert bgt_b200.pbf_plan(pbf, 96, 200)["blocks"] == 4      # rows 96..199 live in blocks 3..6
    # rows of > 8 KB RLE become single-row streamed tiles
    m = 30000
    rows = np.array([(np.arange(m) & 1), np.zeros(m), (np.arange(m) & 1) ^ 1], dtype=np.uint8)
    info = bgt_b200.pbf_plan(oracle.encode_pbf(rows, shift=2))
    assert info["big_tiles"] == 1 and info["max_row_bytes"] > 30000   # the PBWT sorts after row 0: only that row is long
    # corrupt images are rejected, not walked
    for bad in (pbf[:40], b"XXXX" + pbf[4:], pbf[:-9], pbf[:745] + b"\xff" * 15 + pbf[760:]):
        with pytest.raises(bgt_b200.B200Error):
            bgt_b200.pbf_plan(bad)
    p = oracle.Pbf(pbf)
    assert p.row_bytes(0, 700) > p.row_bytes(0, 700, False) > 0
    p.close()


def test_shard_rows_partition():
    from bgt_b200.shard import shard_rows
    for n, shift, world in [(1000000, 13, 8), (10000000, 13, 8), (8193, 13, 4), (5, 13, 2), (700, 5, 3), (0, 13, 2)]:
        cover = []
        for r in range(world):
            b, e = shard_rows(n, shift, r, world)
            assert (b % (1 << shift) == 0 or b == n) and b <= e <= n
            cover.append((b, e))
        assert cover[0][0] == 0 and cover[-1][1] == n
        for (b0, e0), (b1, e1) in zip(cover, cover[1:]):
            assert e0 == b1 or (e0 == n and b1 == n)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch.distributed as dist
from bgt_b200.shard import sharded_scan
from oracle import oracle as orc
from cohorts import haplo_matrix
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
mat = haplo_matrix(700, 60, 9)
pbf = orc.encode_pbf(mat, shift=5)
grp = (np.arange(30) % 2 + 1).astype(np.uint32)
flt = "AC1/AN1>0.1&&AC2==0"
def scan_rows(beg, end):            # stand-in for the GPU scan of this rank's shard (host logic under test, not kernels)
    r = orc.Pbf(pbf).scan(beg, end - beg, group=grp, n_groups=2, flt=flt)
    c = r["counts"].astype(np.int64)
    return dict(counts=r["counts"], passed=r["passed"], totals=[int(c[:, 0].sum()), int(c[:, 1].sum()), int(c[:, 2].sum()), int(r["passed"].sum())])
out = sharded_scan(scan_rows, 700, 5, dist.get_rank(), 2)
whole = orc.Pbf(pbf).scan(0, 700, group=grp, n_groups=2, flt=flt)
c = whole["counts"].astype(np.int64)
assert out["totals"] == [int(c[:, 0].sum()), int(c[:, 1].sum()), int(c[:, 2].sum()), int(whole["passed"].sum()), 700], out["totals"]
if dist.get_rank() == 0:
    assert (out["counts"] == whole["counts"]).all() and (out["passed"] == whole["passed"]).all()
    assert out["rows"] == (0, 352)
dist.barrier(); dist.destroy_process_group()
print("rank ok")
'''


def test_two_rank_sharding_over_gloo(oracle, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0 and b"rank ok" in out, out.decode()[-800:]


def test_bgzf_block_walk_without_a_device():
    """b200_bgzf_inflate(out=NULL) only walks the BGZF block headers (bgzf.c:259-281, 318-351) and sums ISIZE: host logic."""
    import ctypes as C
    import struct
    import zlib
    import bgt_b200
    L = bgt_b200.load_library()

    def block(data, extra=b""):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        raw = c.compress(data) + c.flush()
        xlen = 6 + len(extra)
        total = 12 + xlen + len(raw) + 8
        return (b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", xlen) + extra + b"BC" + struct.pack("<HH", 2, total - 1) + raw +
                struct.pack("<II", zlib.crc32(data), len(data)))

    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    f = block(b"a" * 1000) + block(b"hello", extra=b"XY\x01\x00z") + block(b"") + block(bytes(65536)) + eof
    buf = (C.c_uint8 * len(f)).from_buffer_copy(f)
    assert L.b200_bgzf_inflate(None, buf, len(f), None, 0) == 1000 + 5 + 65536
    for bad in (f[:50], b"\x1f\x8b\x08\x00" + f[4:], f[:-10]):
        b2 = (C.c_uint8 * len(bad)).from_buffer_copy(bad)
        assert L.b200_bgzf_inflate(None, b2, len(bad), None, 0) < 0
        assert L.b200_strerror()


# ------------------------------------------------------------------------------------------------ seam A: in-memory codec + library hygiene

class _Pbc(C.Structure):   # pbwt.h:6-9
    _fields_ = [("m", C.c_int32), ("l", C.c_int32), ("S0", C.POINTER(C.c_int32)), ("S", C.POINTER(C.c_int32)), ("u", C.POINTER(C.c_uint8))]


def _codec(L):
    L.pbc_init.restype = C.POINTER(_Pbc)
    L.pbc_init.argtypes = [C.c_int]
    L.pbc_enc.argtypes = [C.POINTER(_Pbc), C.c_void_p]
    L.pbc_dec.argtypes = [C.POINTER(_Pbc), C.c_void_p]
    L.pbs_dec.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def test_seam_a_exports_the_whole_of_pbwt_h():
    """pbwt.h:35-130 completely: the file API and the in-memory codec, so that pbwt.o can be dropped from a host application."""
    L = C.CDLL(os.path.join(ROOT, "bgt_b200", "lib", "libpbwt_b200.so"))
    for n in ("pbf_open_w", "pbf_open_r", "pbf_close", "pbf_write", "pbf_read", "pbf_seek", "pbf_subset", "pbf_get_g", "pbf_get_m", "pbf_get_n",
              "pbf_get_shift", "pbc_init", "pbc_enc", "pbc_dec", "pbs_dec", "pbc_enc_core", "pbc_dec_core"):
        assert hasattr(L, n), n
    for n in header_functions(os.path.join(ROOT, "include", "pbwt_b200.h"), "pb[cfs]_"):
        assert hasattr(L, n), n


def test_host_codec_matches_the_reference(ref):
    """pbc_enc / pbc_dec / pbs_dec of libpbwt_b200.so against the UNMODIFIED reference's (oracle/_ref/libbgtref.so) on the same rows:
    same RLE bytes, same permutations, same decoded bits, same (rank, slot) lists after every row."""
    mine = _codec(C.CDLL(os.path.join(ROOT, "bgt_b200", "lib", "libpbwt_b200.so")))
    theirs = _codec(C.CDLL(ref.REF_LIB))
    rng = np.random.default_rng(5)
    for m, n_rows in ((1, 5), (2, 8), (37, 60), (300, 120), (70001, 12)):
        mat = (haplo_matrix(n_rows, m, 100 + m) & 1).astype(np.uint8)
        mat[n_rows // 2] = 0
        mat[n_rows // 2 + 1] = 1                                   # constant rows (pbwt.c:75-77)
        if m > 20:
            mat[1, :] = (np.arange(m) // 17) & 1                   # runs of 17: two code bytes each
        ea, eb = mine.pbc_init(m), theirs.pbc_init(m)
        da, db = mine.pbc_init(m), theirs.pbc_init(m)
        n_sub = min(m, 9)
        cols = rng.choice(m, size=n_sub, replace=False)
        sub_a = np.zeros((n_sub, 2), np.uint32)
        sub_a[:, 0] = np.sort(cols); sub_a[:, 1] = np.argsort(cols)   # S = identity before row 0: rank = column (pbwt.c:343)
        sub_b = sub_a.copy()
        for k in range(n_rows):
            row = np.ascontiguousarray(mat[k])
            mine.pbc_enc(ea, row.ctypes.data); theirs.pbc_enc(eb, row.ctypes.data)
            la, lb = ea.contents.l, eb.contents.l
            ua, ub = bytes(ea.contents.u[:la + 1]), bytes(eb.contents.u[:lb + 1])
            assert ua == ub and ua[-1] == 0, (m, k)
            assert ea.contents.S[:m] == eb.contents.S[:m], (m, k)
            buf = C.create_string_buffer(ua, la + 1)
            mine.pbc_dec(da, buf); theirs.pbc_dec(db, buf)
            assert bytes(da.contents.u[:m]) == bytes(db.contents.u[:m]) == row.tobytes(), (m, k)
            assert da.contents.S[:m] == db.contents.S[:m]
            if n_sub < m:                                           # pbf_subset falls back to full decoding otherwise (pbwt.c:377)
                oa, ob = np.zeros(n_sub, np.uint8), np.zeros(n_sub, np.uint8)
                mine.pbs_dec(m, n_sub, sub_a.ctypes.data, buf, oa.ctypes.data)
                theirs.pbs_dec(m, n_sub, sub_b.ctypes.data, buf, ob.ctypes.data)
                assert (oa == ob).all() and (sub_a == sub_b).all(), (m, k)
                assert (oa == row[cols]).all()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        for p in (ea, eb, da, db):
            libc.free(C.cast(p, C.c_void_p))                        # "It should be freed with free()", pbwt.h:103


def test_library_code_never_terminates_the_process():
    """The seams are library code (bgt-server links them): no exit()/abort() -- failures return NULL / negative like the
    reference's (pbwt.c:228-235, bgt.c:880-888)."""
    for rel in ("bgt_b200/host/pbwt_shim.c", "bgt_b200/host/pbc_host.c", "integration/bgtm_shim.c", "bgt_b200/csrc/api.cu"):
        txt = re.sub(r"/\*.*?\*/|//[^\n]*", "", open(os.path.join(ROOT, rel)).read(), flags=re.S)
        assert not re.search(r"\b(exit|abort|_exit)\s*\(", txt), rel


def test_seam_a_rejects_files_without_an_index(tmp_path, oracle):
    """pbf_open_r: a PBF whose trailing index offset is missing or hostile is refused with a message (no silent empty result,
    no read outside the image: the offset is untrusted)."""
    L = C.CDLL(os.path.join(ROOT, "bgt_b200", "lib", "libpbwt_b200.so"))
    L.pbf_open_r.restype = C.c_void_p
    L.pbf_b200_strerror.restype = C.c_char_p
    good = oracle.encode_pbf(haplo_matrix(20, 30, 3), shift=3)
    for name, data in (("trunc", good[:-8]), ("wrap", good[:-8] + (0xFFFFFFFFFFFFFFF8).to_bytes(8, "little")), ("zero", good[:-8] + bytes(8))):
        fn = tmp_path / (name + ".pbf")
        fn.write_bytes(data)
        assert not L.pbf_open_r(str(fn).encode()), name
    import bgt_b200
    for data in (good[:-8] + (0xFFFFFFFFFFFFFFF8).to_bytes(8, "little"), good[:-8] + (len(good) - 5).to_bytes(8, "little")):
        with pytest.raises(bgt_b200.B200Error):
            bgt_b200.pbf_plan(data)
