"""Drop-in tests: the reference's own host application linked against the B200 seams (integration/_build/bgt and
pbfview) must print byte-identical output to the unmodified reference (oracle/_ref) for the same arguments.

This is the differential matrix of SURVEY section 4 (seam A: pbfview; seam B: bgt view).
"""
import os
import re
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


def run_routed(exe, args, env=None):
    """stdout plus the route counters the seams print at exit with BGT_B200_ROUTE=1 (include/pbwt_b200.h): which path served
    the output -- the device `view` pipeline (view_fast.c), seam B batches (bgtm_shim.c), seam A batches (pbwt_shim.c), or
    the reference's own CPU loop (ref_bgtm_read)."""
    e = dict(os.environ)
    e.update(env or {})
    e["BGT_B200_ROUTE"] = "1"
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    err = r.stderr.decode()
    if r.returncode != 0:
        print(err)                                                # (pytest shortens the assertion's repr)
    assert r.returncode == 0, (exe, args, err[-500:])
    m = re.search(r"\[b200 route\] (.*)", err)
    assert m, "no route line on stderr: %r" % err[-300:]
    route = {k: int(v) for k, v in (kv.split("=") for kv in m.group(1).split())}
    return r.stdout, route, err


def on_device_pipeline(args):
    """view_fast.c takes `view [-G] [-C] [-f ..] [-s ..] prefix`; everything else is seam B.  Filters with `**` are handed over
    (host libm) before the first output byte."""
    takes = all(a in ("-C", "-G", "-f", "-s") or not a.startswith("-") for a in args)
    return takes, takes and any("**" in a for a in args)


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
    got, route, _ = run_routed(NEW_BGT, ["view"] + args + [prefix])
    assert got == want
    assert len(want) > 0
    # ... and it was the device that produced it, not the reference's CPU loop behind a silent fall-back
    fast, handed_over = on_device_pipeline(args)
    assert route["ref_bgtm_read"] == 0, route
    if fast and not handed_over:
        assert route["view_fast"] == 1 and route["view_fast_to_ref"] == 0 and route["seamB_batches"] == 0 and route["seamA_batches"] == 0, route
    elif handed_over:
        assert route["view_fast"] == 1 and route["view_fast_to_ref"] == 1 and route["seamB_batches"] > 0, route
    else:
        assert route["view_fast"] == 0 and route["seamB_batches"] > 0, route


def test_view_bed_regions_are_batched(tools, cohort_small, tmp_path):
    """`-B` with scattered intervals: seam B pulls the records through the reference's BED iterator and hands the GPU one batch
    of REGIONS (b200_scan_regions) instead of every row from the first to the last record."""
    prefix, _ = cohort_small
    bed = tmp_path / "r.bed"
    ivs = [(1500, 1700), (30000, 30400), (30900, 31000), (60000, 60010), (82900, 83100), (90000, 90990)]   # POS = 1000 + 10*row; crosses row 8192
    bed.write_text("".join("11\t%d\t%d\n" % iv for iv in ivs))
    for args in (["-B", str(bed), "-C"], ["-B", str(bed), "-f", "AC>0", "-G"], ["-B", str(bed), "-s", ",S0000003,S0000010"]):
        want = run(tools.REF_BGT, ["view"] + args + [prefix])
        got, route, _ = run_routed(NEW_BGT, ["view"] + args + [prefix])
        assert got == want, args
        assert want.count(b"\n") > 60
        assert route["seamB_batches"] > 0 and route["region_launches"] > 0 and route["ref_bgtm_read"] == 0, route


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


def test_disabled_seam_is_visible_in_the_route(tools, cohort_small):
    """the route counters do tell the paths apart: with the seams switched off the reference's own bgtm_read serves the records
    (its row decode still goes through seam A)."""
    prefix, _ = cohort_small
    got, route, _ = run_routed(NEW_BGT, ["view", "-f", "AC>0", "-G", "-n", "50", prefix], {"BGT_B200_DISABLE": "1"})
    assert got == run(tools.REF_BGT, ["view", "-f", "AC>0", "-G", "-n", "50", prefix])
    assert route["ref_bgtm_read"] > 0 and route["view_fast"] == 0 and route["seamB_batches"] == 0 and route["seamA_batches"] > 0


@pytest.fixture(scope="module")
def cohort_100k(tools, tmp_path_factory):
    """BASELINE width: 100 000 samples x 16 384 sites (two checkpoint blocks) of the bench's synthetic cohort, on disk."""
    import bgt_b200
    tmp = tmp_path_factory.mktemp("c100k")
    prefix = os.path.join(str(tmp), "c.bgt")
    with bgt_b200.Context(0) as ctx:
        pb = bgt_b200.synth_cohort(ctx, 100000, 16384, seed=20261017)
        with open(prefix + ".pbf", "wb") as f:
            f.write(memoryview(pb.image()))
        pb.close()
    subprocess.run([tools.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)
    return prefix


def test_baseline_configs_2_3_4_at_full_width_against_the_reference(tools, cohort_100k, tmp_path):
    """BASELINE configs 2, 3 and 4 through the drop-in CLI at 100 000 samples: byte-identical VCF to the unmodified reference,
    produced by the device pipeline (route counters)."""
    prefix = cohort_100k
    rng = np.random.default_rng(1)
    sel = sorted(rng.choice(100000, size=200, replace=False).tolist())
    lst = tmp_path / "sub200.txt"
    lst.write_text("".join("S%07d\n" % s for s in sel))
    cases = {
        "config2": ["-f", "AC>0", "-G"],
        "config3": ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>0.1&&AC2==0", "-G"],     # (passes no site of this cohort: header only)
        "config3 passing": ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>0.01&&AC2>0&&AC1!=AC2", "-G"],
        "config4": ["-s", str(lst)],
        "config4 -f": ["-s", str(lst), "-f", "AC>0"],
        "config3 counts": ["-s", 'grp=="A"', "-s", 'grp=="B"', "-G"],
    }
    for name, args in cases.items():
        want = run(tools.REF_BGT, ["view"] + args + [prefix])
        got, route, _ = run_routed(NEW_BGT, ["view"] + args + [prefix])
        assert got == want, name
        assert want.count(b"\n") > (100 if name != "config3" else 10), name
        assert route["view_fast"] == 1 and route["view_fast_to_ref"] == 0 and route["ref_bgtm_read"] == 0 and route["seamB_batches"] == 0, (name, route)


def vcf_totals(vcf):
    n = an = ac = 0
    for ln in vcf.split(b"\n"):
        if ln and ln[:1] != b"#":
            info = dict(kv.split(b"=") for kv in ln.split(b"\t")[7].split(b";") if b"=" in kv)
            n += 1; an += int(info[b"AN"]); ac += int(info[b"AC"].split(b",")[0])
    return n, an, ac


@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0,0", "all"])
def test_sharded_view_is_identical_and_totals_add_up(tools, tmp_path_factory, devices):
    """Multi-GPU in the C product (the loop of view.c:150-155 cut into region shards): one host thread + context per
    listed device over block-aligned row ranges, shard texts in order behind the single header, totals summed (ncclAllReduce
    over distinct GPUs).  A device listed several times runs the same shard logic on one GPU."""
    import bgt_b200
    if devices == "all" and bgt_b200.lib().b200_device_count() < 2:
        pytest.skip("one GPU visible")
    tmp = tmp_path_factory.mktemp("shard")
    mat = haplo_matrix(2600, 300, 9, switch=0.01)
    prefix = make_bgt(tools, tmp, "five", mat, shift=9)          # 6 checkpoint blocks of 512 rows
    for args in (["-f", "AC>0", "-G"], ["-C"], ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1>AC2", "-G"], ["-s", ",S0000004,S0000001,S0000100"]):
        want = run(tools.REF_BGT, ["view"] + args + [prefix])
        got, route, err = run_routed(NEW_BGT, ["view"] + args + [prefix], {"BGT_B200_DEVICES": devices, "BGT_B200_TOTALS": "1"})
        assert got == want, (devices, args)
        n_dev = bgt_b200.lib().b200_device_count() if devices == "all" else devices.count(",") + 1
        assert route["view_fast"] == 1 and route["gpus"] == min(n_dev, 6) and route["ref_bgtm_read"] == 0, route
        m = re.search(r"\[b200 totals\] gpus=(\d+) sites=(\d+) passed=(\d+) sum_AN=(\d+) sum_AC=(\d+)", err)
        assert m, err[-300:]
        assert int(m.group(2)) == 2600 and int(m.group(3)) == want.count(b"\n") - want.count(b"\n#") - (1 if want.startswith(b"#") else 0)
        if "-C" in args or "-f" in args:
            n, an, ac = vcf_totals(want)
            if n == 2600:                                         # every site printed: the VCF carries all the counts
                assert (int(m.group(4)), int(m.group(5))) == (an, ac)
        if devices == "all":
            assert "ncclAllReduce" in err


def test_view_merges_several_files(tools, tmp_path):
    """Several BGT files in one query (bgtm_read_core, bgt.c:797-878): sites are merged by (contig, pos, rlen, alleles), a file
    that lacks a site contributes missing calls, AC/AN are taken over all files' selected samples.  Seam B serves it with one
    feeder per file (device counts per file, summed on the host); nothing may come from the reference's CPU loop."""
    rng = np.random.default_rng(31)
    specs = [("A", 300, 40, 1000, 10, (0.6, 0.3, 0.05, 0.05)), ("B", 200, 24, 1000, 20, (0.5, 0.3, 0.1, 0.1)), ("C", 90, 10, 1490, 30, (0.7, 0.3, 0.0, 0.0))]
    prefixes = []
    for tag, n, m, pos0, step, probs in specs:
        mat = random_matrix(n, m, int(rng.integers(1 << 30)), probs=probs)
        vcf = tmp_path / (tag + ".vcf")
        vcf.write_bytes(tools.vcf_text(mat, sample_names=["%s%07d" % (tag, i) for i in range(m // 2)], pos0=pos0, step=step))
        prefix = str(tmp_path / (tag + ".bgt"))
        run(tools.REF_BGT, ["import", "-S", prefix, str(vcf)])
        prefixes.append(prefix)
    cases = [["-C"], [], ["-f", "AC>0", "-G"], ["-f", "AC>5&&AN>30"], ["-t", "POS,AC,AN,ALT", "-G"], ["-r", "11:1500-2500", "-C"], ["-n", "7"],
             ["-s", ",A0000001,B0000002", "-s", ",B0000003,C0000001,A0000005", "-f", "AC1>0", "-C"],
             ["-s", ",A0000001,A0000003", "-C"],                   # no sample of files B and C selected: they never read (bgt.c:338)
             ["-s", ",B0000001", "-f", "AC>0"]]
    for args in cases:
        for files in (prefixes, prefixes[::-1], prefixes[:2]):
            want = run(tools.REF_BGT, ["view"] + args + files)
            got, route, _ = run_routed(NEW_BGT, ["view"] + args + files)
            assert got == want, (args, files)
            assert want.count(b"\n") > 12, args
            assert route["seamB_batches"] >= 1 and route["ref_bgtm_read"] == 0 and route["view_fast"] == 0, (args, route)
    # sanity of the fixture: some sites are shared, some are not (total < sum of the files' sites, > the largest file)
    single = [run(tools.REF_BGT, ["view", "-G", p]).count(b"\n11\t") for p in prefixes]
    assert max(single) < run(tools.REF_BGT, ["view", "-G"] + prefixes).count(b"\n11\t") < sum(single)


def allele_names(vcf, picks):
    """`chr:1basedPos:refLen:seq` (view.c:69) of the ALT allele of the picks-th records of a VCF text"""
    recs = [ln.split(b"\t") for ln in vcf.split(b"\n") if ln and ln[:1] != b"#"]
    return [b"%s:%s:%d:%s" % (recs[i][0], recs[i][1], len(recs[i][3]), recs[i][4].split(b",")[0]) for i in picks]


def test_view_allele_queries(tools, cohort_small, tmp_path):
    """`-a` with and without `-S` / `-H` (bgt.c:844-848, 859-876: sites that carry a listed allele, samples that carry all of them,
    haplotype patterns over them), one file and several: served by seam B from device-decoded planes, byte-identical."""
    prefix, _ = cohort_small
    sites = run(tools.REF_BGT, ["view", "-f", "AC>30", "-G", prefix])         # (alleles that many samples carry, so that -S has something to print)
    n_site = sites.count(b"\n11\t")
    assert n_site > 200
    names = allele_names(sites, [3, 4, n_site // 3, n_site // 2, n_site // 2 + 1, n_site - 2, n_site - 1])   # rows on both sides of the checkpoint at 8192
    al = "," + b",".join(names).decode()
    lst = tmp_path / "alleles.txt"
    lst.write_bytes(b"\n".join(names[:4]) + b"\n")
    cases = [["-a", al, "-C"], ["-a", al], ["-a", al, "-S"], ["-a", al, "-H"], ["-a", str(lst), "-H"], ["-a", al, "-f", "AC>0", "-G"],
             ["-a", al, "-s", 'grp=="A"', "-s", 'grp=="B"', "-H"], ["-a", al, "-s", ",S0000001,S0000007,S0000100", "-S"],
             ["-a", "," + names[0].decode(), "-S"], ["-a", "," + (names[1] + b"," + names[5]).decode(), "-S"], ["-a", al, "-r", "11:1000-50000", "-H"]]
    n_out = 0
    for args in cases:
        want = run(tools.REF_BGT, ["view"] + args + [prefix])
        got, route, _ = run_routed(NEW_BGT, ["view"] + args + [prefix])
        assert got == want, args
        assert route["seamB_batches"] >= 1 and route["ref_bgtm_read"] == 0 and route["view_fast"] == 0, (args, route)
        n_out += len(want)
    assert n_out > 2000
    # several files: alleles of sites that one, two or all files hold
    rng = np.random.default_rng(32)
    prefixes, vcfs = [], []
    for tag, n, m, pos0, step in (("A", 120, 40, 1000, 10), ("B", 80, 24, 1000, 20), ("C", 40, 10, 1090, 30)):
        mat = random_matrix(n, m, int(rng.integers(1 << 30)), probs=(0.55, 0.4, 0.05, 0.0))
        vcf = tmp_path / (tag + ".vcf")
        vcf.write_bytes(tools.vcf_text(mat, sample_names=["%s%07d" % (tag, i) for i in range(m // 2)], pos0=pos0, step=step))
        p = str(tmp_path / (tag + ".bgt"))
        run(tools.REF_BGT, ["import", "-S", p, str(vcf)])
        prefixes.append(p)
    merged = run(tools.REF_BGT, ["view", "-G"] + prefixes)
    names = allele_names(merged, [0, 1, 2, 9, 10, 40, 77])
    al = "," + b",".join(names).decode()
    for args in (["-a", al, "-C"], ["-a", al, "-S"], ["-a", al, "-H"], ["-a", al, "-s", ",A0000001,B0000002,C0000003", "-s", ",A0000005,B0000001", "-H"],
                 ["-a", "," + names[3].decode(), "-S"]):
        want = run(tools.REF_BGT, ["view"] + args + prefixes)
        got, route, _ = run_routed(NEW_BGT, ["view"] + args + prefixes)
        assert got == want, args
        assert len(want) > 20 or "-S" in args                    # (no sample need carry all the listed alleles)
        assert route["seamB_batches"] >= 1 and route["ref_bgtm_read"] == 0, (args, route)
