import os, subprocess, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
from oracle import oracle as orc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n, samples = 1000000, 100000
tmp = tempfile.mkdtemp(prefix="bgtcli_")
try:
    with bgt_b200.Context(0) as ctx:
        c = bgt_b200.synth_cohort(ctx, samples, n, seed=20261017)
        prefix = os.path.join(tmp, "c.bgt")
        with open(prefix + ".pbf", "wb") as f:
            f.write(memoryview(c.image()))
        c.close()
    subprocess.run([orc.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "integration", "_build", "bgt")
    env = dict(os.environ, BGT_B200_TRACE="1")
    for args in (["-f", "AC>0", "-G"], ["-f", "AC>0", "-G"]):
        t = time.perf_counter()
        r = subprocess.run([exe, "view"] + args + [prefix], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        print(args, "%.2f s" % (time.perf_counter() - t)); print(r.stderr.decode()[-3000:])
finally:
    shutil.rmtree(tmp, ignore_errors=True)
