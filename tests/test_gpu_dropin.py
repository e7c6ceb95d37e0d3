"""Drop-in tests: the reference's own host application linked against the B200 seams (integration/_build/bgt and
pbfview) must print byte-identical output to the unmodified reference (oracle/_ref) for the same arguments.

This is the differential matrix of SURVEY section 4 (seam A: pbfview; seam B: bgt view).
"""
import os
import subprocess

import numpy as np
import pytest

from cohorts import haplo_matrix, random_matrix

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEW_BGT = os.path.join(ROOT, "integration", "_build", "bgt")
NEW_PBFVIEW = os.path.join(ROOT, "integration", "_build", "pbfview")


def run(exe, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    assert r.returncode == 0, (exe, args, r.stderr.decode()[-500:])
    return r.stdout


@pytest.fixture(scope="module")
def tools(ref):
    if not (os.path.exists(NEW_BGT) and os.path.exists(NEW_PBFVIEW)):
        pytest.skip("integration/_build not present (built where /root/reference is available)")
    return ref


def make_bgt(orc, tmp, name, mat, shift=13):
    prefix = os.path.join(str(tmp), name + ".bgt")
    with open(prefix + ".pbf", "wb") as f:
        f.write(orc.encode_pbf(mat, shift=shift))
    subprocess.run([orc.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)
    return prefix


@pytest.fixture(scope="module")
def cohort_small(tools, tmp_path_factory):
    """400 haplotypes x 9000 sites: crosses the checkpoint at row 8192 with shift 13 (import.c:68)."""
    tmp = tmp_path_factory.mktemp("small")
    mat = haplo_matrix(9000, 400, 77, switch=0.01)
    return make_bgt(tools, tmp, "small", mat), mat


@pytest.fixture(scope="module")
def cohort_wide(tools, tmp_path_factory):
    tmp = tmp_path_factory.mktemp("wide")
    mat = haplo_matrix(500, 20000, 78)
    return make_bgt(tools, tmp, "wide", mat), mat


VIEW_ARGS = [
    ["-C"], ["-G", "-C"], ["-f", "AC>0", "-G"], ["-f", "AN>0&&AC/AN>.05", "-G"], ["-G"], [],
    ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>0.1&&AC2==0", "-G"],
    ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>=0.1&&AC2==0"],
    ["-s", ",S0000001,S0000002", "-s", ",S0000002,S0000003", "-C"],            # overlapping groups: last -s wins (bgt.c:154)
    ["-s", ",S0000004,S0000001"],                                                # output order = sample index order
    ["-s", ",S0000017", "-f", "AC>0"],
    ["-s", 'idx<120&&grp=="B"', "-C"],
    ["-r", "11:2000-30000", "-f", "AC>0", "-G"], ["-r", "11:81000-84000", "-C"],  # region crossing the checkpoint (POS=1000+10*row)
    ["-i", "8193", "-n", "3", "-C"], ["-i", "100", "-n", "50", "-f", "AC>0", "-G"], ["-n", "2"],
    ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC3>0", "-G"],                   # unbound variable: nothing passes
    ["-t", "CHROM,POS,AC,AN,AC/AN", "-G"], ["-t", "POS,AC1,AN1,AC2", "-s", 'grp=="A"', "-s", 'grp=="B"'],
    ["-b", "-C", "-G"], ["-u", "-f", "AC>0"],
    ["-f", "AC**2>AN", "-G"], ["-f", "abs(AC-40)<10&&AN%2==0", "-G"],
]


@pytest.mark.parametrize("args", VIEW_ARGS, ids=[" ".join(a) or "plain" for a in VIEW_ARGS])
def test_view_matches_reference(tools, cohort_small, args):
    prefix, _ = cohort_small
    want = run(tools.REF_BGT, ["view"] + args + [prefix])
    got = run(NEW_BGT, ["view"] + args + [prefix])
    assert got == want
    assert len(want) > 0


def test_view_subset_file_and_wide_cohort(tools, cohort_wide, tmp_path):
    prefix, mat = cohort_wide
    rng = np.random.default_rng(1)
    sel = sorted(rng.choice(mat.shape[1] // 2, size=200, replace=False).tolist())
    lst = tmp_path / "sub200.txt"
    lst.write_text("".join("S%07d\n" % s for s in sel))
    for args in (["-s", str(lst)], ["-s", str(lst), "-f", "AC>0"], ["-f", "AC>0", "-G"], ["-C", "-n", "40"],
                 ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>0.1&&AC2==0", "-G"]):
        want = run(tools.REF_BGT, ["view"] + args + [prefix])
        got = run(NEW_BGT, ["view"] + args + [prefix])
        assert got == want, args
    # small batches / windows exercise the refill logic of the seams
    env = {"BGT_B200_BATCH_BYTES": "100000", "BGT_B200_WINDOW_BLOCKS": "1"}
    assert run(NEW_BGT, ["view", "-s", str(lst)] + [prefix], env) == run(tools.REF_BGT, ["view", "-s", str(lst)] + [prefix])


def test_view_multiallelic_import(tools, tmp_path):
    """A VCF with second ALT alleles and missing calls imported by the reference's own `bgt import`, then viewed by both."""
    mat = random_matrix(60, 24, 5, probs=(0.5, 0.3, 0.1, 0.1))
    vcf = tmp_path / "in.vcf"
    vcf.write_bytes(tools.vcf_text(mat))
    prefix = str(tmp_path / "imp.bgt")
    run(tools.REF_BGT, ["import", "-S", prefix, str(vcf)])
    for args in (["-C"], ["-f", "AC>0", "-G"], ["-s", ",S0000001,S0000005", "-C"]):
        assert run(NEW_BGT, ["view"] + args + [prefix]) == run(tools.REF_BGT, ["view"] + args + [prefix])
    # the same import through the drop-in binary: pbf_open_w / pbf_write of seam A, i.e. the GPU encoder (encode.cu)
    prefix2 = str(tmp_path / "imp2.bgt")
    run(NEW_BGT, ["import", "-S", prefix2, str(vcf)])
    for ext in (".pbf", ".spl"):
        assert open(prefix + ext, "rb").read() == open(prefix2 + ext, "rb").read()


def test_pbfview_matches_reference(tools, cohort_small, tmp_path):
    prefix, mat = cohort_small
    pbf = prefix + ".pbf"
    cases = [[], ["-c", "1", "-c", "3"], ["-c", "399", "-c", "0", "-c", "7", "-c", "7"], ["-r", "0", "-n", "5"], ["-r", "1", "-n", "3"],
             ["-r", "8191", "-n", "4"], ["-r", "8192", "-n", "2"], ["-r", "8193", "-n", "10", "-c", "5", "-c", "2"], ["-r", "8999"],
             ["-r", "4000", "-n", "10", "-c", "11"]]
    for args in cases:
        want = run(tools.REF_PBFVIEW, args + [pbf])
        got = run(NEW_PBFVIEW, args + [pbf])
        assert got == want, args
    # PBF -> PBF re-encode through the drop-in (GPU decode + GPU encode through seam A)
    a, b = tmp_path / "a.pbf", tmp_path / "b.pbf"
    with open(a, "wb") as f:
        subprocess.run([tools.REF_PBFVIEW, "-b", "-r", "100", "-n", "300", pbf], stdout=f, check=True)
    with open(b, "wb") as f:
        subprocess.run([NEW_PBFVIEW, "-b", "-r", "100", "-n", "300", pbf], stdout=f, check=True)
    assert a.read_bytes() == b.read_bytes()


def test_pbfview_pim_to_pbf_on_the_gpu_encoder(tools, tmp_path):
    """PIM text -> PBF (`pbfview -Sb [-s shift]`, pbfview.c:40-71): the drop-in writes through seam A's pbf_open_w /
    pbf_write (device encoder); files must equal the reference's byte for byte, and decode back to the same text."""
    mat = haplo_matrix(700, 150, 31)
    pim = tmp_path / "x.pim"
    pim.write_bytes(tools.pim_text(mat))
    for shift in ("1", "2", "6", "13"):
        a, b = tmp_path / ("ref%s.pbf" % shift), tmp_path / ("new%s.pbf" % shift)
        with open(a, "wb") as f:
            subprocess.run([tools.REF_PBFVIEW, "-Sb", "-s", shift, str(pim)], stdout=f, check=True)
        with open(b, "wb") as f:
            subprocess.run([NEW_PBFVIEW, "-Sb", "-s", shift, str(pim)], stdout=f, check=True)
        assert a.read_bytes() == b.read_bytes(), shift
        assert run(NEW_PBFVIEW, [str(b)]) == run(tools.REF_PBFVIEW, [str(a)])
