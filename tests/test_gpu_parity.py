"""GPU parity tests: the CUDA path (through the C ABI of libbgt_b200.so) against the oracle on the same inputs.

Bar: bit-exact (integer / byte work).  Sizes are what the oracle finishes in seconds; the full BASELINE shapes
are covered through size-independent properties in test_gpu_fullsize.py.
"""
import os

import numpy as np
import pytest

from cohorts import edge_rows, haplo_matrix, random_matrix

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
EX1 = np.array([[0, 1, 2, 0], [2, 0, 1, 1], [1, 0, 1, 1], [0, 1, 0, 1], [1, 2, 0, 0], [1, 0, 1, 2], [0, 1, 1, 1]], np.uint8)

FILTERS = ["AC>0", "AN>0&&AC/AN>.05", "AC1/AN1>0.1&&AC2==0", "AC1/AN1>=0.1&&AC2==0", "AC/AN", "AC3>0", "AC*2+1-AN%7",
           "AC//3==AN>>2", "-AC+AN", "!AC", "~AC&255", "AC**2>AN", "abs(AC-AN)>3", "log(AC)>0", "AC==AN||AC<2",
           "AC1+AC2==AC", "AN1/AN2", "0x10+010+AC>30", "'a'=='a'&&AC", "AC/0", "AC-(-AC)", "AC1/AN1-AC2/AN2>0.05", ".5*AN>AC",
           "AC%4==3", "AN2-AC2<3||AC1>AN1/2"]


@pytest.fixture(scope="module")
def b200():
    import bgt_b200
    return bgt_b200


@pytest.fixture(scope="module")
def ctx(b200):
    c = b200.Context(0)
    yield c
    c.close()


def codes(res, which="hap_bytes"):
    return res[which][0] | (res[which][1] << 1)


def unpack_bits(bits, n_track):
    b = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")
    return b[:, :n_track]


def check_scan(b200, ctx, orc, pbf_bytes, out_samples=None, group=None, n_groups=1, flt=None, row_beg=0, n_rows=None, hap=True):
    """Run the same query through the GPU library and the oracle and require identical results."""
    want_p = orc.Pbf(pbf_bytes)
    if n_rows is None:
        n_rows = want_p.n - row_beg
    want = want_p.scan(row_beg, n_rows, out_samples=out_samples, group=group, n_groups=n_groups, flt=flt, want_hap=hap)
    want_p.close()
    pb = b200.Pbf.from_bytes(ctx, pbf_bytes)
    q = b200.Query(ctx, pb, out_samples=out_samples, group=group, n_groups=n_groups, flt=flt)
    got = b200.scan(ctx, pb, q, row_beg, n_rows, counts=True, hap_bits=hap, hap_bytes=hap)
    q.close()
    pb.close()
    assert got["n"] == want["n"]
    assert (got["counts"] == want["counts"]).all(), "AC/AN differ"
    assert (got["passed"] == want["passed"]).all(), "filter verdicts differ for %r" % flt
    if hap:
        assert (got["hap_bytes"][0][:got["n"]] == want["hap0"]).all() and (got["hap_bytes"][1][:got["n"]] == want["hap1"]).all()
        nt = want["hap0"].shape[1]
        assert (unpack_bits(got["hap_bits"][0][:got["n"]], nt) == want["hap0"]).all()
        assert (unpack_bits(got["hap_bits"][1][:got["n"]], nt) == want["hap1"]).all()
    assert got["totals"][0] == int(want["counts"][:, 0].astype(np.int64).sum())
    assert got["totals"][1] == int(want["counts"][:, 1].astype(np.int64).sum())
    assert got["totals"][2] == int(want["counts"][:, 2].astype(np.int64).sum())
    assert got["totals"][3] == int(want["passed"].sum())
    return got


def test_ex1_fixture(b200, ctx, oracle):
    pbf = oracle.encode_pbf(EX1)
    got = check_scan(b200, ctx, oracle, pbf)
    assert (codes(got) == EX1).all()
    with open(os.path.join(GOLD, "ex1.pbf"), "rb") as f:
        got = check_scan(b200, ctx, oracle, f.read(), flt="AC>0")
    assert (codes(got) == EX1).all()
    # reference fixture semantics: one sample = columns 2s, 2s+1
    got = check_scan(b200, ctx, oracle, pbf, out_samples=[1])
    assert (codes(got) == EX1[:, 2:4]).all()


@pytest.mark.parametrize("name", ["hap_200x96", "rnd_300x37"])
def test_golden_reference_files(b200, ctx, oracle, name):
    mat = np.load(os.path.join(GOLD, name + ".npy"))
    with open(os.path.join(GOLD, name + ".s5.pbf"), "rb") as f:
        pbf = f.read()
    got = check_scan(b200, ctx, oracle, pbf, flt="AC>0")
    nt = 2 * (mat.shape[1] // 2)     # an odd column count leaves the last column untracked (subset mode, pbwt.c:377)
    assert (codes(got) == mat[:, :nt]).all()


@pytest.mark.parametrize("shape", [(700, 96, 6), (300, 38, 5), (2100, 334, 9), (64, 2, 3), (130, 4098, 4), (40, 70002, 3)])
def test_full_decode_and_counts(b200, ctx, oracle, shape):
    n, m, shift = shape
    mat = haplo_matrix(n, m, 7 + n)
    pbf = oracle.encode_pbf(mat, shift=shift)
    got = check_scan(b200, ctx, oracle, pbf, flt="AC>0")
    assert (codes(got) == mat).all()


def test_random_dense_planes(b200, ctx, oracle):
    mat = random_matrix(400, 1000, 11)
    pbf = oracle.encode_pbf(mat, shift=6)
    got = check_scan(b200, ctx, oracle, pbf, flt="AN>0&&AC/AN>.05")
    assert (codes(got) == mat).all()


@pytest.mark.parametrize("m", [5000, 70002])
def test_rle_alphabet_edges(b200, ctx, oracle, m):
    mat = edge_rows(m)
    pbf = oracle.encode_pbf(mat, shift=3)
    got = check_scan(b200, ctx, oracle, pbf)
    assert (codes(got) == mat).all()


def test_rows_larger_than_staging_buffer(b200, ctx, oracle):
    # alternating 0101... rows are m bytes of RLE each: > 8 KB, so the kernel streams them in pieces
    m = 30000
    rng = np.random.default_rng(5)
    rows = [(np.arange(m) & 1).astype(np.uint8), haplo_matrix(1, m, 1)[0], ((np.arange(m) // 2) & 1).astype(np.uint8) * 3,
            (rng.random(m) < 0.5).astype(np.uint8), (rng.random(m) < 0.3).astype(np.uint8) * 2, haplo_matrix(1, m, 2)[0],
            (np.arange(m) & 1).astype(np.uint8) ^ 1]
    mat = np.array(rows * 3, dtype=np.uint8)
    pbf = oracle.encode_pbf(mat, shift=2)
    grp = (np.arange(m // 2) % 2 + 1).astype(np.uint32)
    got = check_scan(b200, ctx, oracle, pbf, group=grp, n_groups=2, flt="AC1/AN1>0.1&&AC2==0")
    assert (codes(got) == mat).all()


@pytest.mark.parametrize("n_rows", [8191, 8192, 8193, 16385])
def test_checkpoint_boundaries_shift13(b200, ctx, oracle, n_rows):
    mat = haplo_matrix(n_rows, 64, 13, switch=0.05)
    pbf = oracle.encode_pbf(mat, shift=13)
    got = check_scan(b200, ctx, oracle, pbf, flt="AC>0", hap=True)
    assert (codes(got) == mat).all()


def test_groups_and_filters(b200, ctx, oracle):
    n, m = 900, 600
    mat = haplo_matrix(n, m, 21)
    pbf = oracle.encode_pbf(mat, shift=7)
    ns = m // 2
    two = (np.arange(ns) % 2 + 1).astype(np.uint32)             # 50/50 groups A/B as in BASELINE config 3
    for flt in FILTERS:
        check_scan(b200, ctx, oracle, pbf, group=two, n_groups=2, flt=flt, hap=False)
    for flt in ("AC>0", "AC1>0", "AN1==AN", "AC2>0"):
        check_scan(b200, ctx, oracle, pbf, flt=flt, hap=False)     # single implicit group
    rng = np.random.default_rng(9)
    for G in (3, 5, 32):
        grp = rng.integers(1, G + 1, size=ns).astype(np.uint32)
        check_scan(b200, ctx, oracle, pbf, group=grp, n_groups=G, flt="AC%d>0&&AN1>0" % G, hap=False)
    # groups over a subset of samples: the rest is not tracked at all (bgt.c:214-220)
    sel = np.sort(rng.choice(ns, size=120, replace=False)).astype(np.int32)
    grp = rng.integers(1, 3, size=sel.size).astype(np.uint32)
    got = check_scan(b200, ctx, oracle, pbf, out_samples=sel, group=grp, n_groups=2, flt="AC1/AN1>0.1&&AC2==0")
    cols = np.stack([2 * sel, 2 * sel + 1], axis=1).ravel()
    assert (codes(got) == mat[:, cols]).all()


def test_subset_extraction(b200, ctx, oracle):
    # BASELINE config 4 shape, scaled: a 200-sample subset out of many, crossing checkpoints
    n, m = 1500, 6000
    mat = haplo_matrix(n, m, 33)
    pbf = oracle.encode_pbf(mat, shift=8)
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(m // 2, size=200, replace=False)).astype(np.int32)
    got = check_scan(b200, ctx, oracle, pbf, out_samples=sel)
    cols = np.stack([2 * sel, 2 * sel + 1], axis=1).ravel()
    assert (codes(got) == mat[:, cols]).all()
    got = check_scan(b200, ctx, oracle, pbf, out_samples=np.array([17], np.int32), flt="AC>0")
    assert (codes(got) == mat[:, 34:36]).all()


@pytest.mark.parametrize("rng_seed", [0, 1])
def test_row_ranges_seek(b200, ctx, oracle, rng_seed):
    # pbf_seek semantics (pbwt.c:349-372): start mid-block, cross blocks, end anywhere
    n, m, shift = 1000, 128, 6
    mat = haplo_matrix(n, m, 40 + rng_seed)
    pbf = oracle.encode_pbf(mat, shift=shift)
    rng = np.random.default_rng(rng_seed)
    for beg, cnt in [(0, 1), (63, 2), (64, 64), (65, 300), (999, 1), (500, 500)] + [(int(rng.integers(0, n - 1)), int(rng.integers(1, 200))) for _ in range(4)]:
        cnt = min(cnt, n - beg)
        got = check_scan(b200, ctx, oracle, pbf, row_beg=beg, n_rows=cnt, flt="AC>0")
        assert (codes(got) == mat[beg:beg + cnt]).all()


def test_row_shards_resident(b200, ctx, oracle):
    # region sharding: make only some checkpoint blocks resident (b200_pbf_load row range) and scan inside them
    n, m, shift = 700, 90, 5
    mat = haplo_matrix(n, m, 51)
    pbf = oracle.encode_pbf(mat, shift=shift)
    want = oracle.Pbf(pbf).scan(0, n, flt="AC>0", want_hap=True)
    for beg, end in [(0, 700), (0, 32), (96, 200), (333, 700), (640, 700)]:
        pb = b200.Pbf.from_bytes(ctx, pbf, beg, end)
        assert pb.row_beg <= beg and pb.row_end >= end and pb.row_beg % 32 == 0
        q = b200.Query(ctx, pb, flt="AC>0")
        got = b200.scan(ctx, pb, q, beg, end - beg, hap_bytes=True)
        assert (got["counts"] == want["counts"][beg:end]).all()
        assert (got["passed"] == want["passed"][beg:end]).all()
        assert (got["hap_bytes"][0] == want["hap0"][beg:end]).all()
        with pytest.raises(b200.B200Error):
            if pb.row_beg > 0:
                b200.scan(ctx, pb, q, 0, 1)
            else:
                raise b200.B200Error("n/a")
        q.close()
        pb.close()


def test_open_from_a_file_path(b200, ctx, oracle, tmp_path):
    """b200_pbf_open (pbf_open_r's file-name form, pbwt.c:221-262): whole file and a row range; a missing file fails."""
    mat = haplo_matrix(500, 64, 9)
    fn = tmp_path / "x.pbf"
    fn.write_bytes(oracle.encode_pbf(mat, shift=6))
    for beg, end in ((0, -1), (130, 400)):
        pb = b200.Pbf.open(ctx, str(fn), beg, end)
        q = b200.Query.columns(ctx, pb)
        lo, hi = (0, 500) if end < 0 else (beg, end)
        got = b200.scan(ctx, pb, q, lo, hi - lo, hap_bytes=True)
        assert ((got["hap_bytes"][0] | got["hap_bytes"][1] << 1) == mat[lo:hi]).all()
        q.close(); pb.close()
    with pytest.raises(b200.B200Error):
        b200.Pbf.open(ctx, str(tmp_path / "missing.pbf"))


def test_noncanonical_rle_streams(b200, ctx, oracle):
    m = 40
    rows = [(bytes([5 << 1, 1, 3 << 1, 12 << 1 | 1, 32, (16 + 1) << 1, 4 << 1]), bytes([(16 + 2) << 1, 8 << 1 | 1])),
            (bytes([10 << 1 | 1, 10 << 1 | 1, 33, 10 << 1, 10 << 1]), bytes([(16 + 2) << 1 | 1, 33, 8 << 1])),
            (bytes([1, 7 << 1, 1, 33, 13 << 1 | 1, (16 + 1) << 1, 4 << 1 | 1]), bytes([(16 + 2) << 1, 8 << 1]))] * 5
    pbf = oracle.encode_pbf_rle(m, rows, shift=2)
    check_scan(b200, ctx, oracle, pbf, flt="AC>0")
    # the same plane-0 streams over an empty plane 1: grouped count-only queries take the split scan, whose per-group
    # marginals merge the bytes of a run and skip the empty ones (marginal.cu)
    quiet = bytes([(16 + 2) << 1, 8 << 1])
    pbf = oracle.encode_pbf_rle(m, [(r0, quiet) for r0, _ in rows], shift=3)
    grp = (np.arange(m // 2) % 3 + 1).astype(np.uint32)
    want = oracle.Pbf(pbf).scan(0, len(rows), group=grp, n_groups=3, flt="AC1>0")
    pb = b200.Pbf.from_bytes(ctx, pbf)
    q = b200.Query(ctx, pb, group=grp, n_groups=3, flt="AC1>0")
    for kw in (dict(), dict(no_split=True)):
        got = b200.scan(ctx, pb, q, 0, len(rows), **kw)
        assert (got["counts"] == want["counts"]).all() and (got["passed"] == want["passed"]).all(), kw
    q.close(); pb.close()


def test_error_behaviour(b200, ctx, oracle):
    with pytest.raises(b200.B200Error):
        b200.Pbf.from_bytes(ctx, b"not a pbf file at all, but long enough to carry a header.....")
    pbf = oracle.encode_pbf(EX1)
    pb = b200.Pbf.from_bytes(ctx, pbf)
    for bad in ["AC>", "(AC>0", "AC>0)", "AC=0", "'abc"]:
        with pytest.raises(b200.B200Error):
            b200.Query(ctx, pb, flt=bad)
    with pytest.raises(b200.B200Error):
        b200.Query(ctx, pb, out_samples=[1, 0])           # not ascending
    with pytest.raises(b200.B200Error):
        b200.Query(ctx, pb, group=[1, 3], n_groups=2)     # group id out of range
    q = b200.Query(ctx, pb)
    assert b200.scan(ctx, pb, q, 7, 5)["n"] == 0          # past the end: nothing, like pbf_read returning NULL
    assert b200.scan(ctx, pb, q, 5, 50)["n"] == 2
    q.close()
    pb.close()
    # truncated file: index record missing
    with pytest.raises(b200.B200Error):
        b200.Pbf.from_bytes(ctx, pbf[:-9])


def test_row_index_chase_among_fake_tags(b200, ctx, oracle):
    """Blocks of >= 768 rows are indexed by twelve chasers that start in the middle of the block (index.cu); here the record
    bytes are full of 'B' bytes that are not record tags (tests/test_rankwalk_model.py has the CPU model of the scheme)."""
    from cohorts import fake_tag_matrix
    mat = fake_tag_matrix(1900, 2048, 9)
    pbf = oracle.encode_pbf(mat, shift=10)                      # 1024 + 876 rows: both blocks take the team chase
    assert pbf.count(b"B") > 4 * 1900
    ref_pb = oracle.Pbf(pbf)
    pb = b200.Pbf.from_bytes(ctx, pbf)
    assert pb.bad_rows == 0 and pb.row_bytes(0, 1900) == ref_pb.row_bytes(0, 1900)
    pb.close()
    check_scan(b200, ctx, oracle, pbf, flt="AC>0")


def test_device_index_reports_corruption(b200, ctx, oracle):
    """The row index is built on the device (index.cu); damaged tags / lengths / snapshots inside a block must fail the load
    (pbf_read would run off the rails: pbwt.c:318-328 has no checks), a row whose run lengths do not sum to m is counted."""
    mat = haplo_matrix(300, 96, 4)
    good = oracle.encode_pbf(mat, shift=5)
    ref_pb = oracle.Pbf(good)
    pb = b200.Pbf.from_bytes(ctx, good)
    assert pb.bad_rows == 0 and pb.row_bytes(0, 300) == ref_pb.row_bytes(0, 300) and pb.row_bytes(37, 201, False) == ref_pb.row_bytes(37, 201, False)
    pb.close()
    first_rec = 16 + 1 + 8 * 96                               # header, 'S', two snapshots: the first 'B' record
    assert good[first_rec:first_rec + 1] == b"B"
    for off, val in ((first_rec, ord("X")),                    # record tag
                     (first_rec + 3, 0x7f),                    # plane-0 length far beyond the block
                     (first_rec + 4, 0x80)):                   # negative length
        bad = bytearray(good); bad[off] = val
        with pytest.raises(b200.B200Error):
            b200.Pbf.from_bytes(ctx, bytes(bad))
    bad = bytearray(good); bad[17:21] = (96).to_bytes(4, "little")    # snapshot names column m
    with pytest.raises(b200.B200Error):
        b200.Pbf.from_bytes(ctx, bytes(bad))
    # lengths that parse but do not sum to m: the row decodes as all-REF and is counted (the reference has undefined behaviour)
    bad = bytearray(good)
    l0 = int.from_bytes(good[first_rec + 1:first_rec + 5], "little")
    assert l0 >= 1
    bad[first_rec + 5] ^= 0x02                                 # change the first run's length digit
    pb = b200.Pbf.from_bytes(ctx, bytes(bad))
    assert pb.bad_rows >= 1
    pb.close()


def test_synth_generator_is_truthful_and_canonical(b200, ctx, oracle):
    # the device generator must produce exactly the file the reference encoder writes for the decoded matrix
    for n_samples, n_rows, shift, seed in [(40, 300, 5, 1), (333, 200, 6, 2), (1500, 70, 4, 3)]:
        pb = b200.synth_cohort(ctx, n_samples, n_rows, seed=seed, shift=shift, r_max=16, p1_one_in=4)
        img = pb.image().tobytes()
        mat = oracle.decode_all(img)
        assert mat.shape == (n_rows, 2 * n_samples)
        assert oracle.encode_pbf(mat, shift=shift) == img, "generated file is not what the encoder writes for its own matrix"
        q = b200.Query(ctx, pb, flt="AC>0")
        got = b200.scan(ctx, pb, q, 0, n_rows, hap_bytes=True)
        assert (codes(got) == mat).all()
        want = oracle.Pbf(img).scan(0, n_rows, flt="AC>0")
        assert (got["counts"] == want["counts"]).all() and (got["passed"] == want["passed"]).all()
        # and the image loads back through the host path
        pb2 = b200.Pbf.from_bytes(ctx, img)
        got2 = b200.scan(ctx, pb2, q, 0, n_rows)
        assert (got2["counts"] == want["counts"]).all()
        assert pb2.row_bytes(0, n_rows) == oracle.Pbf(img).row_bytes(0, n_rows)
        q.close(); pb.close(); pb2.close()


@pytest.mark.parametrize("case", ["sparse", "dense", "mixed", "allones", "wide", "noisy0", "deep", "noisydeep"])
def test_split_scan_equals_general_and_oracle(b200, ctx, oracle, case):
    """Count-only full-cohort scans take the split path (plane-0 marginal + walk of the columns that carry plane-1 codes);
    it must agree with the general walk and with the oracle whatever the density of plane 1."""
    rng = np.random.default_rng(12)
    if case == "sparse":
        mat = haplo_matrix(900, 8200, 5, p_missing_row=0.1, p_multi_row=0.1)
    elif case == "dense":
        mat = random_matrix(300, 4400, 6)
    elif case == "mixed":
        mat = haplo_matrix(600, 4400, 7)
        mat[300:420] = random_matrix(120, 4400, 8)
    elif case == "allones":
        mat = haplo_matrix(200, 4400, 9)
        mat[50] = 2; mat[51] = 3; mat[120] = 1; mat[121] = 0
    elif case == "noisy0":
        # plane 0 incompressible (thousands of runs per row: the composite maps overflow and the walk falls back to
        # row-by-row), plane 1 sparse (so the split path is still taken)
        mat = (rng.random((400, 4400)) < 0.5).astype(np.uint8)
        for k in range(0, 400, 7):
            mat[k, rng.integers(0, 4400, size=5)] = 2 + (k & 1)
    elif case == "noisydeep":
        # the same with 32 row groups per checkpoint block: the segmented marginals find no composite maps and hand the
        # block to a single CTA; the quiet stretch in the second block has them
        mat = (rng.random((1700, 1300)) < 0.5).astype(np.uint8)
        mat[1024:1700] = haplo_matrix(676, 1300, 13)
        for k in range(0, 1700, 31):
            mat[k, rng.integers(0, 1300, size=3)] = 2 + (k & 1)
    elif case == "deep":
        # long blocks (many 32-row groups per checkpoint) with plane-1 codes on most rows: exercises the composite
        # maps of both planes and the per-group fallbacks at block ends
        mat = haplo_matrix(2300, 1200, 11, p_missing_row=0.45, p_multi_row=0.45)
    else:
        mat = haplo_matrix(70, 70002, 10, p_missing_row=0.3, p_multi_row=0.3)
    shift = 4 if case == "wide" else 11 if case == "deep" else 10 if case == "noisydeep" else 8 if case == "noisy0" else 6
    pbf = oracle.encode_pbf(mat, shift=shift)
    n = mat.shape[0]
    want = oracle.Pbf(pbf).scan(0, n, flt="AC>0")
    pb = b200.Pbf.from_bytes(ctx, pbf)
    q = b200.Query(ctx, pb, flt="AC>0")
    for kw in (dict(), dict(no_split=True), dict(no_compose=True), dict(cols_per_thread=1), dict(cols_per_thread=8)):
        got = b200.scan(ctx, pb, q, 0, n, **kw)
        assert (got["counts"] == want["counts"]).all(), (case, kw)
        assert (got["passed"] == want["passed"]).all()
    for beg, cnt in ((37, 200), (63, 2), (n - 5, 5)):
        cnt = min(cnt, n - beg)
        got = b200.scan(ctx, pb, q, beg, cnt)
        assert (got["counts"] == want["counts"][beg:beg + cnt]).all(), (case, beg)
    q.close()
    # grouped queries: per-group marginals (margpiece.cu piece lists; marginal.cu bit-vector partition with no_pieces / no_segments) + the plane-1 queries
    ns = mat.shape[1] // 2
    for G in (2, 3, 8):
        grp = rng.integers(1, G + 1, size=ns).astype(np.uint32)
        flt = "AC1/AN1>0.1&&AC2==0"
        wantg = oracle.Pbf(pbf).scan(0, n, group=grp, n_groups=G, flt=flt)
        qg = b200.Query(ctx, pb, group=grp, n_groups=G, flt=flt)
        for kw in (dict(), dict(no_split=True), dict(no_segments=True), dict(no_pieces=True)):
            got = b200.scan(ctx, pb, qg, 0, n, **kw)
            assert (got["counts"] == wantg["counts"]).all(), (case, G, kw)
            assert (got["passed"] == wantg["passed"]).all()
        for beg, cnt in ((41, 150), (n // 2 + 3, n // 2 - 20), (n - 9, 9)):
            cnt = min(cnt, n - beg)
            got = b200.scan(ctx, pb, qg, beg, cnt)
            assert (got["counts"] == wantg["counts"][beg:beg + cnt]).all(), (case, G, beg)
        qg.close()
    pb.close()


@pytest.mark.parametrize("shape", [(300, 4480, 6, "sparse"), (300, 2100, 5, "sparse"), (200, 1500, 7, "dense"), (1200, 20000, 13, "sparse")])
def test_load_scan_pipeline_equals_load_then_scan(b200, ctx, oracle, shape):
    """b200_pbf_load_scan queues the pair walk, the per-site AC/AN + verdict and the copy home behind every chunk of the load
    (the last block of a chunk waits for the next chunk: its backward pairs start from the next snapshot).  Results must
    equal the oracle and the two-call path for whole files, row ranges that start and end inside blocks, blocks that are
    off the split path (dense plane 1), several groups and host-evaluated filters (both not fused)."""
    n_samples, n_rows, shift, kind = shape
    pb0 = b200.synth_cohort(ctx, n_samples, n_rows, seed=n_rows, shift=shift, r_max=12, p1_one_in=8 if kind == "sparse" else 1,
                            p1_max_iv=3 if kind == "sparse" else 30, p1_max_len=8 if kind == "sparse" else 64)
    img = pb0.image()
    pb0.close()
    op = oracle.Pbf(img.tobytes())
    m = 2 * n_samples
    grp = (np.arange(n_samples) % 3 == 0).astype(np.uint32) + 1
    cases = [dict(flt="AC>0"), dict(flt=None), dict(flt="AN<%d" % m), dict(group=grp, n_groups=2, flt="AC1>AC2"), dict(flt="AC**2>AN")]
    BS = 1 << shift
    ranges = [(0, n_rows), (BS + 5, min(n_rows, BS * 9 + 1)), (n_rows - BS - 3, n_rows), (BS, BS + 1)]
    for kw in cases:
        q = b200.Query(ctx, m, **kw)
        for beg, end in ranges:
            pb, got = b200.load_scan(ctx, img, q, beg, end)
            want = op.scan(beg, end - beg, **kw)
            assert got["n"] == end - beg
            assert (got["counts"] == want["counts"]).all(), (kw, beg, end)
            assert (got["passed"] == want["passed"]).all(), (kw, beg, end)
            c = want["counts"].astype(np.int64)
            assert got["totals"][:3] == [int(c[:, 0].sum()), int(c[:, 1].sum()), int(c[:, 2].sum())] and got["totals"][3] == int(want["passed"].sum())
            again = b200.scan(ctx, pb, q, beg, end - beg)                     # the handle it returns is an ordinary resident PBF
            assert (again["counts"] == want["counts"]).all()
            pb.close()
        q.close()
    op.close()


def test_batched_regions_equal_one_scan_per_region(b200, ctx, oracle):
    """b200_scan_regions (`-r` / `-B` / server-style short queries: one pbf_seek + read loop per region in the reference,
    pbwt.c:349-372) queues all regions and synchronises once: outputs are the regions' rows back to back and must equal the
    oracle region by region -- counts, verdicts, decoded planes, totals; full cohort (split path), groups, subsets, `**`."""
    mat = haplo_matrix(3000, 260, 21, switch=0.01)
    img = oracle.encode_pbf(mat, shift=8)                                       # 12 checkpoint blocks of 256 rows
    pb = b200.Pbf.from_bytes(ctx, img)
    regions = [(5, 40), (255, 3), (256, 1), (700, 0), (1000, 600), (2990, 10), (10, 20)]   # mid-block, across checkpoints, empty, unordered
    grp = (np.arange(130) % 2 + 1).astype(np.uint32)
    sel = np.array([3, 17, 64, 100, 129], dtype=np.int32)
    for kw, hap in ((dict(flt="AC>0"), False), (dict(group=grp, n_groups=2, flt="AC1>=AC2"), False), (dict(out_samples=sel, flt="AN>2"), True),
                    (dict(flt="AC**2>AN"), False), (dict(), True)):
        q = b200.Query(ctx, pb, **kw)
        got = b200.scan_regions(ctx, pb, q, regions, hap_bytes=hap)
        assert got["n"] == sum(r[1] for r in regions)
        o = 0
        tot = np.zeros(4, np.int64)
        for beg, n in regions:
            # (a fresh reader per region: like the reference's, a subset reader takes its ranks from the snapshot of the block its
            # cursor stands in when pbf_subset is called, pbwt.c:374-388 -- a cursor left mid-block by an earlier scan would not do)
            op = oracle.Pbf(img)
            want = op.scan(beg, n, want_hap=hap, **kw)
            op.close()
            assert (got["counts"][o:o + n] == want["counts"]).all(), (kw, beg)
            assert (got["passed"][o:o + n] == want["passed"]).all(), (kw, beg)
            if hap:
                assert (got["hap_bytes"][0][o:o + n] == want["hap0"]).all() and (got["hap_bytes"][1][o:o + n] == want["hap1"]).all(), (kw, beg)
            c = want["counts"].astype(np.int64)
            tot += np.array([c[:, 0].sum(), c[:, 1].sum(), c[:, 2].sum(), want["passed"].sum()], dtype=np.int64)
            o += n
        assert got["totals"] == [int(x) for x in tot], kw
        q.close()
    pb.close()


def test_blocks_between_the_two_pair_capacities_stay_on_the_split_path(b200, ctx, oracle):
    """Plane 1 with more ones per block than the launches of the load pipeline are sized for (3 per column) but within what the pair
    lists hold (up to 24 per column): such blocks are flagged 2, keep the split scan, and get extension launches of the select and
    of the pair walk sized from their real counts; b200_pbf_load_scan must notice and redo them.  One group, groups, row ranges."""
    n_samples, n_rows, shift = 1500, 1200, 8
    pb0 = b200.synth_cohort(ctx, n_samples, n_rows, seed=77, shift=shift, r_max=12, p1_one_in=1, p1_max_iv=6, p1_max_len=40)
    img = pb0.image()
    n_blk = (n_rows + (1 << shift) - 1) >> shift
    assert b200.lib().b200_pbf_split_blocks(pb0.h) == n_blk                    # every block on the split path ...
    op = oracle.Pbf(img.tobytes())
    m = 2 * n_samples
    want = op.scan(0, n_rows, flt="AC>0")
    plane1 = (m * n_rows - int(want["counts"][:, 0].astype(np.int64).sum())) + int(want["counts"][:, 2].astype(np.int64).sum())
    assert plane1 / n_blk > 3.5 * m                                             # ... although they hold more pairs than 3 per column
    q = b200.Query(ctx, m, flt="AC>0")
    for kw in (dict(), dict(no_split=True)):
        got = b200.scan(ctx, pb0, q, 0, n_rows, **kw)
        assert (got["counts"] == want["counts"]).all() and (got["passed"] == want["passed"]).all(), kw
    for beg, cnt in ((300, 500), (255, 2), (n_rows - 100, 100)):
        got = b200.scan(ctx, pb0, q, beg, cnt)
        assert (got["counts"] == want["counts"][beg:beg + cnt]).all(), beg
    pb, got = b200.load_scan(ctx, img, q, 0, n_rows)                            # the pipeline's launches do not reach all pairs: redone
    assert (got["counts"] == want["counts"]).all() and (got["passed"] == want["passed"]).all()
    assert got["totals"][3] == int(want["passed"].sum())
    pb.close(); q.close()
    grp = (np.arange(n_samples) % 3 == 0).astype(np.uint32) + 1
    wantg = op.scan(0, n_rows, group=grp, n_groups=2, flt="AC1>AC2")
    qg = b200.Query(ctx, m, group=grp, n_groups=2, flt="AC1>AC2")
    got = b200.scan(ctx, pb0, qg, 0, n_rows)
    assert (got["counts"] == wantg["counts"]).all() and (got["passed"] == wantg["passed"]).all()
    qg.close(); pb0.close(); op.close()
