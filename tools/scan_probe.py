"""One resident config-2 scan (for ncu captures of the scan kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
ctx = bgt_b200.Context(0)
n, samples = 1000000, 100000
cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=20261017)
q = bgt_b200.Query(ctx, cohort, flt="AC>0")
for _ in range(3):
    r = bgt_b200.scan(ctx, cohort, q, 0, n)
print("scan %.2f ms (select %.2f)" % (ctx.last_ms(1), ctx.last_ms(4)), r["totals"])
