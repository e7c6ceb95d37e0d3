"""GPU parity of the site side of `bgt view` on the device (sites.cu): the .bcf / .csi of a BGT database
inflated, indexed (RNI stretches) and parsed by kernels, and the VCF text of `view -G` assembled on the device, against the
record lines the unmodified reference prints for the same query."""
import os
import subprocess

import numpy as np
import pytest

from cohorts import haplo_matrix, random_matrix

pytestmark = pytest.mark.gpu

INDEL_VCF = """##fileformat=VCFv4.1
##INFO=<ID=END,Number=1,Type=Integer,Description="end">
##ALT=<ID=DEL,Description="Deletion">
##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">
##contig=<ID=7,length=159138663>
##contig=<ID=11,length=135006516>
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3
7\t100\trs1\tC\tT\t50\tPASS\t.\tGT\t0|1\t1|1\t.|0
7\t210\t.\tCATAT\tC,CAT\t100\tPASS\t.\tGT\t2|1\t0|0\t1|0
7\t300\t.\tA\tG,T\t100\tPASS\t.\tGT\t1|2\t0|2\t.|.
11\t19\t.\tGT\tG,CT\t100\tPASS\t.\tGT\t1|2\t0|1\t0|0
11\t22\t.\tGTATATAGCGA\tGTATA\t100\tPASS\t.\tGT\t1|0\t1|1\t0|0
11\t5000\t.\tT\tTAAA\t100\tPASS\t.\tGT\t0|0\t0|1\t1|1
11\t6000\t.\tT\t<DEL>\t100\tPASS\tEND=6100\tGT\t0|1\t0|1\t1|1
11\t7000\t.\tT\tC\t100\tPASS\t.\tGT\t0|1\t.|1\t1|1
"""


@pytest.fixture(scope="module")
def b200():
    import bgt_b200
    return bgt_b200


@pytest.fixture(scope="module")
def ctx(b200):
    c = b200.Context(0)
    yield c
    c.close()


def ref_lines(ref, args, prefix):
    r = subprocess.run([ref.REF_BGT, "view"] + args + [prefix], stdout=subprocess.PIPE, check=True).stdout
    body = b"".join(ln + b"\n" for ln in r.split(b"\n") if ln and not ln.startswith(b"#"))
    return body


def make_bgt(ref, tmp, name, mat, shift=13):
    prefix = os.path.join(str(tmp), name + ".bgt")
    with open(prefix + ".pbf", "wb") as f:
        f.write(ref.encode_pbf(mat, shift=shift))
    subprocess.run([ref.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)
    return prefix


def open_db(b200, ctx, prefix, with_csi=True):
    bcf = open(prefix + ".bcf", "rb").read()
    csi = open(prefix + ".bcf.csi", "rb").read() if with_csi else None
    sites = b200.Sites(ctx, bcf, csi)
    pb = b200.Pbf.from_bytes(ctx, open(prefix + ".pbf", "rb").read())
    return sites, pb


@pytest.mark.parametrize("with_csi", [True, False])
def test_site_table(b200, ctx, ref, tmp_path, with_csi):
    """3000 records = three RNI stretches (rec_shift 10): row numbers and positions of every record."""
    mat = haplo_matrix(3000, 40, 3)
    prefix = make_bgt(ref, tmp_path, "t", mat)
    sites, pb = open_db(b200, ctx, prefix, with_csi)
    assert sites.n == 3000
    rows, pos = sites.rows()
    assert (rows == np.arange(3000)).all() and (pos == 999 + 10 * np.arange(3000)).all()   # mksites: POS = 1000 + 10*row, 0-based in BCF
    assert b"##INFO=<ID=_row" in sites.header()
    sites.close(); pb.close()


def test_view_text_matches_reference(b200, ctx, ref, tmp_path):
    mat = haplo_matrix(9000, 400, 77, switch=0.01)          # crosses the checkpoint at row 8192; rows with <M> and missing calls
    prefix = make_bgt(ref, tmp_path, "v", mat)
    sites, pb = open_db(b200, ctx, prefix)
    ns = mat.shape[1] // 2
    grp = (np.arange(ns) % 2 + 1).astype(np.uint32)
    cases = [(["-G", "-C"], dict(), True), (["-f", "AC>0", "-G"], dict(flt="AC>0"), False), (["-G"], dict(), False),
             (["-f", "AN>0&&AC/AN>.05", "-G"], dict(flt="AN>0&&AC/AN>.05"), False),
             (["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>0.1&&AC2==0", "-G"], dict(group=grp, n_groups=2, flt="AC1/AN1>0.1&&AC2==0"), False),
             (["-s", 'grp=="A"', "-s", 'grp=="B"', "-G"], dict(group=grp, n_groups=2), False),
             (["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC3>0", "-G"], dict(group=grp, n_groups=2, flt="AC3>0"), False)]
    for args, qa, with_counts in cases:
        q = b200.Query(ctx, pb, **qa)
        got, n_lines = sites.view_text(pb, q, with_counts=with_counts)
        want = ref_lines(ref, args, prefix)
        assert got == want, args
        assert n_lines == want.count(b"\n")
        q.close()
    # a sample subset (pbs_dec path of the reference): 10 samples, counts over them only
    sel = np.array([3, 17, 18, 50, 99, 100, 101, 150, 198, 199], np.int32)
    q = b200.Query(ctx, pb, out_samples=sel, flt="AC>0")
    got, _ = sites.view_text(pb, q)
    assert got == ref_lines(ref, ["-s", "," + ",".join("S%07d" % s for s in sel), "-f", "AC>0", "-G"], prefix)
    q.close(); sites.close(); pb.close()


def test_view_text_indels_two_contigs_via_reference_import(b200, ctx, ref, tmp_path):
    """records with END= (rlen != len(REF), bgt.c:824-827), atomized multi-allelic sites, two contigs, non-empty input IDs:
    imported by the reference's own `bgt import`, printed by both."""
    vcf = tmp_path / "in.vcf"
    vcf.write_text(INDEL_VCF)
    prefix = str(tmp_path / "i.bgt")
    subprocess.run([ref.REF_BGT, "import", "-S", prefix, str(vcf)], check=True, stderr=subprocess.DEVNULL)
    sites, pb = open_db(b200, ctx, prefix)
    for args, qa, wc in ((["-G", "-C"], dict(), True), (["-G"], dict(), False), (["-f", "AC>0", "-G"], dict(flt="AC>0"), False)):
        q = b200.Query(ctx, pb, **qa)
        got, _ = sites.view_text(pb, q, with_counts=wc)
        assert got == ref_lines(ref, args, prefix), args
        q.close()
    sites.close(); pb.close()


def test_wide_cohort_and_random_codes(b200, ctx, ref, tmp_path):
    mat = random_matrix(1500, 2000, 9, probs=(0.6, 0.3, 0.05, 0.05))
    prefix = make_bgt(ref, tmp_path, "w", mat)
    sites, pb = open_db(b200, ctx, prefix)
    q = b200.Query(ctx, pb, flt="AC*3>AN")
    got, _ = sites.view_text(pb, q)
    assert got == ref_lines(ref, ["-f", "AC*3>AN", "-G"], prefix)
    q.close(); sites.close(); pb.close()


def test_view_text_with_genotype_columns(b200, ctx, ref, tmp_path):
    """`bgt view` WITHOUT -G (config 4: sample-subset extraction to VCF): FORMAT GT + one unphased a/b per selected sample
    from the device-decoded bit planes (bgt_gen_gt bgt.c:290-313), in record windows."""
    mat = haplo_matrix(9000, 400, 78, switch=0.01)
    prefix = make_bgt(ref, tmp_path, "g", mat)
    sites, pb = open_db(b200, ctx, prefix)
    ns = mat.shape[1] // 2
    sel = np.array([0, 3, 17, 18, 50, 99, 100, 101, 150, 198, 199], np.int32)
    names = "," + ",".join("S%07d" % s for s in sel)
    grp = (np.arange(ns) % 2 + 1).astype(np.uint32)
    cases = [(["-s", names], dict(out_samples=sel), False), (["-s", names, "-f", "AC>0"], dict(out_samples=sel, flt="AC>0"), False),
             (["-C"], dict(), True), ([], dict(), False), (["-s", ",S0000017", "-C"], dict(out_samples=np.array([17], np.int32)), True),
             (["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>=0.1&&AC2==0"], dict(group=grp, n_groups=2, flt="AC1/AN1>=0.1&&AC2==0"), False)]
    for args, qa, wc in cases:
        q = b200.Query(ctx, pb, **qa)
        want = ref_lines(ref, args, prefix)
        got, n_lines = sites.view_text(pb, q, with_counts=wc, genotypes=True)
        assert got == want, args
        assert n_lines == want.count(b"\n")
        # the same in three record windows (one of them crossing the checkpoint at row 8192)
        parts = [sites.view_text(pb, q, with_counts=wc, genotypes=True, rec_beg=a, rec_end=b)[0] for a, b in ((0, 1000), (1000, 8500), (8500, 9000))]
        assert b"".join(parts) == want, args
        q.close()
    sites.close(); pb.close()
