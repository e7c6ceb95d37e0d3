"""Resident scan timing + result digest at the bench workload (A/B of kernel variants: run once per variant, e.g. with
BGT_B200_LEGACY_QUERY=1, and compare the md5).  python tools/scan_probe.py [rows] [samples] [cols_per_thread]"""
import hashlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bgt_b200

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
cpt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = bgt_b200.Context(0)
pb = bgt_b200.synth_cohort(ctx, samples, rows, seed=20261017)
q = bgt_b200.Query(ctx, pb, flt="AC>0")
res = bgt_b200.scan(ctx, pb, q, 0, rows, cols_per_thread=cpt)
h = hashlib.md5(res["counts"].tobytes() + res["passed"].tobytes()).hexdigest()
d_counts = torch.empty((rows, q.stride), dtype=torch.int32, device="cuda:0")
d_pass = torch.empty((rows,), dtype=torch.uint8, device="cuda:0")
ms = []
for i in range(12):
    bgt_b200.scan_device(ctx, pb, q, 0, rows, d_counts.data_ptr(), d_pass.data_ptr())
    bgt_b200.collect(ctx)
    ms.append((ctx.last_ms(0), ctx.last_ms(1)))
ms = ms[2:]
print("variant=%s rows=%d walk_ms=%.3f scan_ms=%.3f md5=%s totals=%s" % ("legacy" if os.environ.get("BGT_B200_LEGACY_QUERY") else "pair", rows,
      sum(m[0] for m in ms) / len(ms), sum(m[1] for m in ms) / len(ms), h, res["totals"]))
