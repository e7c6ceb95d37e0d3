"""Time the BASELINE configs 2-4 on the resident 100k x 1M cohort (device events of the whole scan)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
ctx = bgt_b200.Context(0)
n, samples = 1000000, 100000
cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=20261017)
grp = (np.arange(samples) % 2 + 1).astype(np.uint32)
sel = np.sort(np.random.default_rng(1).choice(samples, size=200, replace=False)).astype(np.int32)
cases = [("config2 view -f'AC>0' -G", dict(flt="AC>0"), {}),
         ("config2 general walk (no split)", dict(flt="AC>0"), dict(no_split=True)),
         ("config3 two groups 50/50", dict(group=grp, n_groups=2, flt="AC1/AN1>0.1&&AC2==0"), {}),
         ("config3 general walk (no split)", dict(group=grp, n_groups=2, flt="AC1/AN1>0.1&&AC2==0"), dict(no_split=True)),
         ("config4 200-sample subset + genotype bit planes", dict(out_samples=sel), dict(hap_bits=True))]
cases.insert(3, ("config3 one CTA per block (no segments)", cases[2][1], dict(no_segments=True)))
only = sys.argv[1:]
ref = {}
for name, qa, sa in cases:
    if only and not any(o in name for o in only): continue
    q = bgt_b200.Query(ctx, cohort, **qa)
    for _ in range(2):
        t = time.perf_counter(); r = bgt_b200.scan(ctx, cohort, q, 0, n, **sa); dt = time.perf_counter() - t
    print("marginal %.2f ms" % ctx.last_ms(5), end="  ")
    print("%-50s kernels %.1f ms  (call %.1f ms)  %.2f M sites/s  passed %d" % (name, ctx.last_ms(1), dt * 1e3, n / ctx.last_ms(1) / 1e3, r["totals"][3]))
    key = name.split()[0]
    if key in ref: assert (ref[key] == r["counts"]).all(), "counts differ between the paths of " + key
    else: ref[key] = r["counts"]
    q.close()
