import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
ctx = bgt_b200.Context(0)
n, samples = 1000000, 100000
cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=20261017)
q = bgt_b200.Query(ctx, cohort, flt="AC>0")
for C in (0, 8, 4, 2, 1):
    for _ in range(2):
        r = bgt_b200.scan(ctx, cohort, q, 0, n, cols_per_thread=C)
    print("C=%d: all decode %.2f ms, select %.2f ms, walk %.2f ms, scan %.2f ms" % (C, ctx.last_ms(0), ctx.last_ms(4), ctx.last_ms(0) - ctx.last_ms(4), ctx.last_ms(1)))
