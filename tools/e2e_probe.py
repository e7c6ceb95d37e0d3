"""End-to-end step (b200_pbf_load_scan from a pinned host image) with the load pipeline's trace.
python tools/e2e_probe.py [rows] [samples] [groups]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bgt_b200
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = bgt_b200.Context(0)
pb = bgt_b200.synth_cohort(ctx, samples, rows, seed=20261017)
n = bgt_b200.lib().b200_pbf_image_size(pb.h)
host = bgt_b200.host_alloc(n)
pb.image(out=host)
pb.close()
if groups > 1:
    q = bgt_b200.Query(ctx, 2 * samples, group=(np.arange(samples) % groups + 1).astype(np.uint32), n_groups=groups, flt="AC1/AN1>0.1&&AC2==0")
else:
    q = bgt_b200.Query(ctx, 2 * samples, flt="AC>0")
hc = bgt_b200.host_alloc(rows * 4 * q.stride).view(np.int32).reshape(rows, q.stride)
hp = bgt_b200.host_alloc(rows)
out = {"counts": hc, "passed": hp}
for i in range(6):
    if i == 5:
        os.environ["BGT_B200_TRACE"] = "1"
    t0 = time.perf_counter()
    p, r = bgt_b200.load_scan(ctx, host, q, 0, rows, out=out)
    t1 = time.perf_counter()
    p.close()
    print("step %d: %.2f ms (+ close %.2f ms)  totals %s  scan kernels %.2f ms" % (i, (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3, r["totals"], ctx.last_ms(1)), flush=True)
