import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
ctx = bgt_b200.Context(0)
n, samples = 1000000, 100000
t=time.perf_counter(); cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=7); print("synth %.3f s" % (time.perf_counter()-t))
sz = bgt_b200.lib().b200_pbf_image_size(cohort.h)
host = bgt_b200.host_alloc(sz); cohort.image(out=host)
q0 = bgt_b200.Query(ctx, cohort, flt="AC>0")
h_counts = bgt_b200.host_alloc(n*6*4).view(np.int32).reshape(n,6); h_pass = bgt_b200.host_alloc(n)
for it in range(3):
    t0=time.perf_counter(); info = bgt_b200.pbf_plan(host); t1=time.perf_counter()
    pb = bgt_b200.Pbf.from_bytes(ctx, host, prepare_count_scan=True); t2=time.perf_counter()
    q = bgt_b200.Query(ctx, pb, flt="AC>0"); t3=time.perf_counter()
    r = bgt_b200.scan(ctx, pb, q, 0, n, out={"counts":h_counts,"passed":h_pass}); t4=time.perf_counter()
    q.close(); pb.close(); t5=time.perf_counter()
    print("plan(host walk only) %.1f ms | load %.1f ms (h2d event %.1f) | query %.1f | scan %.1f (walk %.1f all-kernels %.1f d2h %.1f) | close %.1f" % (
        1e3*(t1-t0), 1e3*(t2-t1), ctx.last_ms(2), 1e3*(t3-t2), 1e3*(t4-t3), ctx.last_ms(0), ctx.last_ms(1), ctx.last_ms(3), 1e3*(t5-t4)))
