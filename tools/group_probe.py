"""Resident grouped scan (BASELINE config 3 shape): piece-list marginals (margpiece.cu) against the row loop over bit vectors
(marginal.cu) -- same digest, kernel times.  python tools/group_probe.py [rows] [samples] [groups]"""
import hashlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bgt_b200

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = bgt_b200.Context(0)
pb = bgt_b200.synth_cohort(ctx, samples, rows, seed=20261017)
q = bgt_b200.Query(ctx, pb, group=(np.arange(samples) % groups + 1).astype(np.uint32), n_groups=groups, flt="AC1/AN1>0.1&&AC2==0")
for kw in (dict(), dict(no_pieces=True)):
    ms = []
    for i in range(5):
        res = bgt_b200.scan(ctx, pb, q, 0, rows, **kw)
        ms.append((ctx.last_ms(1), ctx.last_ms(5)))
    h = hashlib.md5(res["counts"].tobytes() + res["passed"].tobytes()).hexdigest()
    ms = ms[1:]
    print("%-16s scan_ms=%.3f marginal_ms=%.3f md5=%s totals=%s" % (kw or "pieces", sum(m[0] for m in ms) / len(ms), sum(m[1] for m in ms) / len(ms), h, res["totals"]), flush=True)
