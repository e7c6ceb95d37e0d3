import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
ctx = bgt_b200.Context(0)
n, samples = 262144, 100000
cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=20261017)
grp = (np.arange(samples) % 2 + 1).astype(np.uint32)
q = bgt_b200.Query(ctx, cohort, group=grp, n_groups=2, flt="AC1/AN1>0.1&&AC2==0")
for _ in range(2):
    r = bgt_b200.scan(ctx, cohort, q, 0, n)
print("scan %.2f ms marginal %.2f" % (ctx.last_ms(1), ctx.last_ms(5)))
