"""Parity at BASELINE.json's full shapes through size-independent properties (the oracle cannot finish 1M x 100k
sites in test time): shard-invariance, tiling-invariance, group additivity, subset == columns of the full decode,
filter verdict == recomputation from the counts, plus oracle spot checks of row windows (block starts and the
rows just before a checkpoint) on the downloaded image."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SAMPLES, ROWS = 100000, 1000000      # BASELINE configs 2-4
SEED = 20261017


@pytest.fixture(scope="module")
def b200():
    import bgt_b200
    return bgt_b200


@pytest.fixture(scope="module")
def ctx(b200):
    c = b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def cohort(b200, ctx):
    pb = b200.synth_cohort(ctx, SAMPLES, ROWS, seed=SEED)
    yield pb
    pb.close()


@pytest.fixture(scope="module")
def full_scan(b200, ctx, cohort):
    q = b200.Query(ctx, cohort, flt="AC>0")
    res = b200.scan(ctx, cohort, q, 0, ROWS)
    q.close()
    return res


def test_config2_counts_are_consistent(b200, ctx, cohort, full_scan):
    c = full_scan["counts"].astype(np.int64)
    m = 2 * SAMPLES
    assert full_scan["n"] == ROWS
    assert (c[:, 0] <= m).all() and (c[:, 1] >= 0).all() and (c[:, 1] + c[:, 2] <= c[:, 0]).all()
    assert (c[:, 3:6] == c[:, 0:3]).all()                                  # one group: AN1/AC1 mirror AN/AC
    assert (full_scan["passed"] == (c[:, 1] > 0)).all()                    # -f'AC>0' recomputed from the counts
    assert full_scan["totals"] == [int(c[:, 0].sum()), int(c[:, 1].sum()), int(c[:, 2].sum()), int((c[:, 1] > 0).sum())]
    assert cohort.bad_rows == 0


def test_config2_shard_and_tiling_invariance(b200, ctx, cohort, full_scan):
    """Region shards (any cut, also mid-block) and any columns-per-thread tiling must reproduce the one-shot scan."""
    q = b200.Query(ctx, cohort, flt="AC>0")
    cuts = [0, 8192 * 17, 8192 * 40 + 123, 8192 * 41, 700001, ROWS]
    for beg, end in zip(cuts, cuts[1:]):
        part = b200.scan(ctx, cohort, q, beg, end - beg)
        assert (part["counts"] == full_scan["counts"][beg:end]).all(), (beg, end)
        assert (part["passed"] == full_scan["passed"][beg:end]).all()
    beg, n = 8192 * 50 - 100, 8192 + 200
    for C in (1, 2, 4, 8):
        part = b200.scan(ctx, cohort, q, beg, n, cols_per_thread=C)
        assert (part["counts"] == full_scan["counts"][beg:beg + n]).all(), C
    q.close()


def test_config3_two_groups_add_up(b200, ctx, cohort, full_scan):
    grp = (np.arange(SAMPLES) % 2 + 1).astype(np.uint32)                   # 50/50 as BASELINE config 3
    flt = "AC1/AN1>0.1&&AC2==0"
    q = b200.Query(ctx, cohort, group=grp, n_groups=2, flt=flt)
    res = b200.scan(ctx, cohort, q, 0, ROWS)
    q.close()
    c = res["counts"].astype(np.int64)
    assert (c[:, 0:3] == full_scan["counts"][:, 0:3]).all()                # totals do not depend on the grouping
    assert (c[:, 3] + c[:, 6] == c[:, 0]).all() and (c[:, 4] + c[:, 7] == c[:, 1]).all() and (c[:, 5] + c[:, 8] == c[:, 2]).all()
    with np.errstate(divide="ignore", invalid="ignore"):
        want = (c[:, 4] / c[:, 3] > 0.1) & (c[:, 7] == 0)                  # IEEE double division as kexpr (kexpr.c:144)
    assert (res["passed"] == want).all()
    assert res["totals"][3] == int(want.sum())


def test_config3_two_groups_against_the_oracle_at_full_width(b200, ctx, cohort, oracle):
    """BASELINE config 3 at 100 000 samples x 1 000 000 sites: all nine count columns (AN, AC, AC<M>, then AN#/AC#/AC#<M> of both
    groups) and the -f verdict of the grouped split scan -- the segmented marginal path (seed vectors pushed through the composite
    maps, one CTA per 256-row segment), the one-CTA-per-block variant and the general walk -- against the oracle's restatement of
    bgtm_cal_info + bgtm_pass_site_flt (bgt.c:735-757, 712-719) on row windows at the start of the file, around a checkpoint deep
    inside it (rows 8192*77 +- 150: both sides of the snapshot) and at its end."""
    grp = (np.arange(SAMPLES) % 2 + 1).astype(np.uint32)
    flt = "AC1/AN1>0.1&&AC2==0"
    q = b200.Query(ctx, cohort, group=grp, n_groups=2, flt=flt)
    full = b200.scan(ctx, cohort, q, 0, ROWS)
    noseg = b200.scan(ctx, cohort, q, 0, ROWS, no_segments=True)
    assert (noseg["counts"] == full["counts"]).all() and (noseg["passed"] == full["passed"]).all()
    nopc = b200.scan(ctx, cohort, q, 0, ROWS, no_pieces=True)                 # the row loop over bit vectors instead of the piece lists
    assert (nopc["counts"] == full["counts"]).all() and (nopc["passed"] == full["passed"]).all()
    img = cohort.image()
    p = oracle.Pbf(img.tobytes())
    for beg, n in ((0, 600), (8192 * 77 - 150, 300), (ROWS - 64, 64)):
        want = p.scan(beg, n, group=grp, n_groups=2, flt=flt)
        assert want["counts"].shape[1] == 9
        assert (full["counts"][beg:beg + n] == want["counts"]).all(), beg
        assert (full["passed"][beg:beg + n] == want["passed"]).all(), beg
        part = b200.scan(ctx, cohort, q, beg, n)                              # the same window as its own scan (mid-block start)
        assert (part["counts"] == want["counts"]).all() and (part["passed"] == want["passed"]).all(), beg
        gen = b200.scan(ctx, cohort, q, beg, n, no_split=True)                # general walk over all columns
        assert (gen["counts"] == want["counts"]).all() and (gen["passed"] == want["passed"]).all(), beg
    # a group assignment that is NOT symmetric (30 % / 70 %, interleaved irregularly): a per-group mix-up cannot cancel out
    rng = np.random.default_rng(3)
    grp2 = (rng.random(SAMPLES) < 0.7).astype(np.uint32) + 1
    q2 = b200.Query(ctx, cohort, group=grp2, n_groups=2, flt="AC1>AC2")
    beg, n = 8192 * 30 - 100, 260
    got = b200.scan(ctx, cohort, q2, beg, n)
    want = p.scan(beg, n, group=grp2, n_groups=2, flt="AC1>AC2")
    assert (got["counts"] == want["counts"]).all() and (got["passed"] == want["passed"]).all()
    q2.close(); q.close(); p.close()


def test_config4_subset_is_columns_of_full_decode(b200, ctx, cohort):
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(SAMPLES, size=200, replace=False)).astype(np.int32)
    qs = b200.Query(ctx, cohort, out_samples=sel)
    sub = b200.scan(ctx, cohort, qs, 0, ROWS, hap_bits=True)               # the pbs_dec-shaped path over all 1M sites
    qs.close()
    qf = b200.Query(ctx, cohort)
    cols = np.stack([2 * sel, 2 * sel + 1], axis=1).ravel()
    for beg in (0, 8192 * 61 - 40, ROWS - 64):
        full = b200.scan(ctx, cohort, qf, beg, 64, hap_bits=True)
        for p in range(2):
            fb = np.unpackbits(full["hap_bits"][p].view(np.uint8), axis=1, bitorder="little")[:, cols]
            sb = np.unpackbits(sub["hap_bits"][p][beg:beg + 64].view(np.uint8), axis=1, bitorder="little")[:, :400]
            assert (fb == sb).all(), (beg, p)
    qf.close()
    # AC/AN of the subset recomputed from its own decoded planes (bgt.c:735-757)
    b0 = np.unpackbits(sub["hap_bits"][0].view(np.uint8), axis=1, bitorder="little")[:, :400]
    b1 = np.unpackbits(sub["hap_bits"][1].view(np.uint8), axis=1, bitorder="little")[:, :400]
    c = sub["counts"]
    assert (c[:, 1] == (b0 & ~b1 & 1).sum(axis=1)).all() and (c[:, 2] == (b0 & b1).sum(axis=1)).all()
    assert (c[:, 0] == 400 - (~b0 & b1 & 1).sum(axis=1)).all()


def test_oracle_spot_checks_on_the_full_cohort(b200, ctx, cohort, full_scan, oracle):
    img = cohort.image()
    p = oracle.Pbf(img.tobytes())
    assert (p.m, p.n) == (2 * SAMPLES, ROWS)
    for beg, n in ((0, 600), (8192 * 77, 300)):
        want = p.scan(beg, n, flt="AC>0")
        assert (full_scan["counts"][beg:beg + n] == want["counts"]).all(), beg
        assert (full_scan["passed"][beg:beg + n] == want["passed"]).all()
    assert cohort.row_bytes(0, 8192 * 3) == p.row_bytes(0, 8192 * 3)
    p.close()


def test_config5_width_one_million_haplotypes(b200, ctx, oracle):
    """BASELINE config 5 width (500k samples, m = 1M) on four checkpoint blocks, oracle-checked at both ends of a block."""
    pb = b200.synth_cohort(ctx, 500000, 4 * 8192, seed=SEED + 5)
    q = b200.Query(ctx, pb, flt="AC>0")
    res = b200.scan(ctx, pb, q, 0, 4 * 8192)
    img = pb.image()
    p = oracle.Pbf(img.tobytes())
    for beg, n in ((0, 150), (8192, 120)):
        want = p.scan(beg, n, flt="AC>0")
        assert (res["counts"][beg:beg + n] == want["counts"]).all(), beg
        assert (res["passed"][beg:beg + n] == want["passed"]).all()
    p.close()
    parts = [b200.scan(ctx, pb, q, b, e - b)["counts"] for b, e in ((0, 9000), (9000, 20000), (20000, 4 * 8192))]
    assert (np.concatenate(parts) == res["counts"]).all()
    q.close()
    pb.close()


def test_encoder_reproduces_the_generated_image_at_full_width(b200, ctx, cohort):
    """decode the first 12 000 rows of the 100 k-sample cohort on the device (crossing the checkpoint at row 8192), encode
    them on the device: the generator's image is canonical and truthful, so the file must come out byte for byte -- 'S'
    snapshots, run-length bytes, index."""
    n = 12000
    small = b200.synth_cohort(ctx, SAMPLES, n, seed=SEED)          # rows are seeded per row: the same first rows
    want = small.image().tobytes()
    q = b200.Query.columns(ctx, small)
    enc = b200.Encoder(ctx, 2 * SAMPLES, 13)
    for beg in range(0, n, 3000):
        r = b200.scan(ctx, small, q, beg, 3000, counts=False, hap_bits=True)
        enc.write_bits(np.ascontiguousarray(np.stack([r["hap_bits"][0], r["hap_bits"][1]], axis=1)))
    assert enc.finish() == want
    enc.close(); q.close(); small.close()


def test_view_text_at_full_width_against_the_reference(b200, ctx, ref, tmp_path):
    """the whole device pipeline of `view -f'AC>0' -G` (inflate, BCF parse, scan, text) on 100 k samples x 16 384 sites against
    the unmodified reference's record lines (the reference needs ~8 s for them)."""
    import os
    import subprocess
    n = 16384
    small = b200.synth_cohort(ctx, SAMPLES, n, seed=SEED)
    prefix = os.path.join(str(tmp_path), "f.bgt")
    with open(prefix + ".pbf", "wb") as f:
        f.write(memoryview(small.image()))
    subprocess.run([ref.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)
    sites = b200.Sites(ctx, open(prefix + ".bcf", "rb").read(), open(prefix + ".bcf.csi", "rb").read())
    q = b200.Query(ctx, small, flt="AC>0")
    got, n_lines = sites.view_text(small, q)
    out = subprocess.run([ref.REF_BGT, "view", "-f", "AC>0", "-G", prefix], stdout=subprocess.PIPE, check=True).stdout
    want = b"".join(ln + b"\n" for ln in out.split(b"\n") if ln and not ln.startswith(b"#"))
    assert got == want and n_lines == want.count(b"\n")
    q.close(); sites.close(); small.close()
