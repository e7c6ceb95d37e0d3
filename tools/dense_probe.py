"""Resident scan of cohorts with a dense plane 1 (the last points of bench.py's sweep).  python tools/dense_probe.py [rows] [samples]"""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bgt_b200
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
ctx = bgt_b200.Context(0)
for one_in, max_iv in ((16, 30), (4, 30), (1, 30)):
    pb = bgt_b200.synth_cohort(ctx, samples, rows, seed=20261017 + 1000 + one_in * 64 + max_iv, p1_one_in=one_in, p1_max_iv=max_iv)
    q = bgt_b200.Query(ctx, pb, flt="AC>0")
    ms = []
    for i in range(4):
        res = bgt_b200.scan(ctx, pb, q, 0, rows)
        ms.append(ctx.last_ms(1))
    gen = bgt_b200.scan(ctx, pb, q, 0, min(rows, 16384), no_split=True)
    same = (gen["counts"] == res["counts"][:len(gen["counts"])]).all()
    print("one_in=%d intervals<=%d: %.2f ms (%.1f M sites/s), split blocks %d of %d, equals general walk on the first rows: %s" %
          (one_in, max_iv, min(ms[1:]), rows / min(ms[1:]) / 1e3, bgt_b200.lib().b200_pbf_split_blocks(pb.h), (rows + 8191) // 8192, same), flush=True)
    q.close(); pb.close()
