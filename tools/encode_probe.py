"""Encoder throughput at the config-2 width: decode rows of a generated cohort on the device, re-encode them on the
device, compare the image with the generator's (truthful, canonical => must be identical), time the encoder."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgt_b200
samples = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
ctx = bgt_b200.Context(0)
cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=5)
want = cohort.image().tobytes()
q = bgt_b200.Query.columns(ctx, cohort)
enc = bgt_b200.Encoder(ctx, 2 * samples, 13)
t_enc = 0.0
step = 2048
for beg in range(0, n, step):
    r = bgt_b200.scan(ctx, cohort, q, beg, min(step, n - beg), counts=False, hap_bits=True)
    bits = np.ascontiguousarray(np.stack([r["hap_bits"][0], r["hap_bits"][1]], axis=1))
    t = time.perf_counter(); enc.write_bits(bits); t_enc += time.perf_counter() - t
img = enc.finish()
print("encoded %d rows x %d haplotypes in %.2f s -> %.0f rows/s (%.1f us/row); image identical to the generator's: %s (%d bytes)" % (
    n, 2 * samples, t_enc, n / t_enc, 1e6 * t_enc / n, img == want, len(img)))
