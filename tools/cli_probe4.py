"""Config 4 through the drop-in CLI: 200-sample subset extraction to VCF on the 100k x 1M database."""
import os, subprocess, sys, tempfile, time, shutil, hashlib, random
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
    subprocess.run([orc.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)
    random.seed(1)
    sel = sorted(random.sample(range(samples), 200))
    lst = os.path.join(tmp, "sub200.txt")
    open(lst, "w").write("".join("S%07d\n" % s for s in sel))
    exe = os.path.join(ROOT, "integration", "_build", "bgt")
    for args in (["-s", lst], ["-s", lst, "-f", "AC>0"]):
        t = time.perf_counter()
        p = subprocess.run("%s view %s %s | md5sum" % (exe, " ".join("'%s'" % a for a in args), prefix), shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, BGT_B200_TRACE="1"))
        dt = time.perf_counter() - t
        print("drop-in bgt view %s: %.2f s -> %.0f sites/s, md5 %s" % (" ".join(args[:1] + args[2:]), dt, n / dt, p.stdout.decode().split()[0]))
        print("".join(l + "\n" for l in p.stderr.decode().split("\n") if l.startswith("[view_fast]")))
    slice_rows = 65536
    reg = "11:%d-%d" % (1000, 1000 + 10 * (slice_rows - 1))
    t = time.perf_counter()
    ref = subprocess.run([orc.REF_BGT, "view", "-s", lst, "-r", reg, prefix], stdout=subprocess.PIPE, check=True).stdout
    dt = time.perf_counter() - t
    mine = subprocess.run([exe, "view", "-s", lst, "-r", reg, prefix], stdout=subprocess.PIPE, check=True).stdout
    print("reference on %d sites: %.2f s -> %.0f sites/s; identical to drop-in: %s" % (slice_rows, dt, slice_rows / dt, ref == mine))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
