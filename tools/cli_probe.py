"""Full-size drop-in CLI timing: `bgt view -f'AC>0' -G` on a 100k-sample x N-site database, the reference host
application linked against seams A+B (integration/_build/bgt), stdout piped to md5; and the unmodified reference
on the first SLICE sites for the rate beside it."""
import os, subprocess, sys, tempfile, time, shutil, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
from oracle import oracle as orc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
samples = 100000
tmp = tempfile.mkdtemp(prefix="bgtcli_")
try:
    with bgt_b200.Context(0) as ctx:
        c = bgt_b200.synth_cohort(ctx, samples, n, seed=20261017)
        prefix = os.path.join(tmp, "c.bgt")
        with open(prefix + ".pbf", "wb") as f:
            f.write(memoryview(c.image()))
        c.close()
    t = time.perf_counter(); subprocess.run([orc.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL); print("mksites %.1f s" % (time.perf_counter() - t))
    exe = os.path.join(ROOT, "integration", "_build", "bgt")
    for args in (["-f", "AC>0", "-G"], ["-G", "-C"], ["-s", 'grp=="A"', "-s", 'grp=="B"', "-f", "AC1/AN1>0.1&&AC2==0", "-G"]):
        for rep in range(2):
            t = time.perf_counter()
            out = subprocess.run([exe, "view"] + args + [prefix], stdout=subprocess.PIPE, check=True).stdout
            dt = time.perf_counter() - t
        print("drop-in bgt view %s: %.2f s -> %.0f sites/s, %d bytes, md5 %s" % (" ".join(args), dt, n / dt, len(out), hashlib.md5(out).hexdigest()))
    slice_rows = 8192
    t = time.perf_counter()
    ref = subprocess.run([orc.REF_BGT, "view", "-f", "AC>0", "-G", "-r", "11:%d-%d" % (1000, 1000 + 10 * (slice_rows - 1)), prefix], stdout=subprocess.PIPE, check=True).stdout
    dt = time.perf_counter() - t
    mine = subprocess.run([exe, "view", "-f", "AC>0", "-G", "-r", "11:%d-%d" % (1000, 1000 + 10 * (slice_rows - 1)), prefix], stdout=subprocess.PIPE, check=True).stdout
    print("reference on %d sites: %.2f s -> %.0f sites/s; identical to drop-in: %s" % (slice_rows, dt, slice_rows / dt, ref == mine))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
