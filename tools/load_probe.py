"""Load-path timing: bare H2D copy (BGT_B200_COPYONLY=1) vs the full load (index + composites)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
ctx = bgt_b200.Context(0)
n, samples = 1000000, 100000
cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=7)
sz = bgt_b200.lib().b200_pbf_image_size(cohort.h)
host = bgt_b200.host_alloc(sz); cohort.image(out=host)
for it in range(4):
    t1 = time.perf_counter()
    pb = bgt_b200.Pbf.from_bytes(ctx, host, prepare_count_scan=True)
    t2 = time.perf_counter()
    print("load %.2f ms (h2d event %.2f)" % (1e3 * (t2 - t1), ctx.last_ms(2)))
    pb.close()
