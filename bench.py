#!/usr/bin/env python
"""bench.py -- sites/sec of the `bgt view -f'AC>0' -G` full-cohort scan (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference --gpus N ...          # the reference's own CPU path on the box's host cores

A "step" is one pass of the hot path (PBWT decode + per-site AC/AN + filter) over one synthetic cohort shard of
`--samples` x `--rows` (default 100k x 1M = BASELINE configs[1]); every rank owns one region shard (weak scaling,
no data-path collective; one NCCL all-reduce of the per-shard totals after the timed region).
  value : sites/s with the .pbf image resident in HBM, timed with CUDA events on the launching stream.
  e2e   : sites/s through the C ABI with HOST buffers: H2D of the .pbf image, row-index walk, kernels, D2H of the
          per-site AC/AN + verdicts -- all inside the timed region.
N = 1 adds: configs 3 and 4 (resident + e2e, oracle-checked), the drop-in CLI end to end on the full database on disk
(`e2e_cli`), a plane-1 density sweep (`p1_sweep`) and the reference on one core (`cpu_baseline`).
N > 1 adds: `config5` -- ONE 500 000-sample x 10 M-site cohort cut into region shards (BASELINE configs[4]) -- and
`h2d_probe`, the host->device bandwidth the ranks get when they copy at the same time (the end-to-end ceiling of the box).
Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
import json
import mmap
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FILTER = "AC>0"
METRIC = "sites/sec for `bgt view -f` full-cohort scan"
DROPIN = os.path.join(ROOT, "integration", "_build", "bgt")


def workload_name(samples, rows):
    return "synthetic %d samples x %d sites per GPU, `view -f'%s' -G` full scan" % (samples, rows, FILTER)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.out = open(self.path, "w")
        self.proc = subprocess.Popen([exe, "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                     stdout=self.out, stderr=subprocess.DEVNULL)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for ln in f:
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0])); mx.append(float(p[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference CLI helpers (oracle/_ref)

def write_bgt(prefix, image):
    """Write <prefix>.pbf from a generated image and its site side through oracle/_ref/mksites (reference library)."""
    from oracle import oracle as orc
    with open(prefix + ".pbf", "wb") as f:
        f.write(memoryview(image))
    subprocess.run([orc.MKSITES, prefix], check=True, stderr=subprocess.DEVNULL)


def view_args(row_beg=None, row_end=None, extra=()):
    cmd = ["view", "-f", FILTER, "-G"] + list(extra)
    if row_beg is not None:  # POS = 1000 + 10*row (mksites.c), inclusive 1-based region
        cmd += ["-r", "11:%d-%d" % (1000 + 10 * row_beg, 1000 + 10 * (row_end - 1))]
    return cmd


def parse_view_counts(vcf_bytes):
    """(POS -> (AN, AC)) of the records `bgt view -G -f` printed."""
    out = {}
    for ln in vcf_bytes.split(b"\n"):
        if not ln or ln[:1] == b"#":
            continue
        f = ln.split(b"\t")
        info = dict(kv.split(b"=") for kv in f[7].split(b";") if b"=" in kv)
        out[int(f[1])] = (int(info[b"AN"]), int(info[b"AC"].split(b",")[0]))
    return out


def tmp_root():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None


def cpu_baseline_sample(ctx, samples, seed, sample_rows, gpu_counts, gpu_pass):
    """Time the UNMODIFIED reference (oracle/_ref/bgt view -f'AC>0' -G, one thread) on the first sample_rows sites of
    the same cohort, and check its VCF against the GPU results of those sites and the drop-in CLI's md5."""
    import bgt_b200
    from oracle import oracle as orc
    if not (orc.have_ref() and os.path.exists(orc.MKSITES)):
        return {"value": None, "unit": "sites/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}, None
    tmp = tempfile.mkdtemp(prefix="bgtb200_", dir=tmp_root())
    try:
        small = bgt_b200.synth_cohort(ctx, samples, sample_rows, seed=seed)   # rows are seeded per row: identical to the big cohort's first rows
        prefix = os.path.join(tmp, "s.bgt")
        write_bgt(prefix, small.image())
        small.close()
        best, out = None, b""
        for _ in range(2):
            t0 = time.perf_counter()
            out = subprocess.run([orc.REF_BGT] + view_args() + [prefix], stdout=subprocess.PIPE, check=True).stdout
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        got = parse_view_counts(out)
        ok = True
        n_pass = 0
        for k in range(sample_rows):
            pos = 1000 + 10 * k
            if gpu_pass[k]:
                n_pass += 1
                ok = ok and got.get(pos) == (int(gpu_counts[k][0]), int(gpu_counts[k][1]))
            else:
                ok = ok and pos not in got
        ok = ok and n_pass == len(got)
        res = {"value": sample_rows / best, "unit": "sites/s", "cores": 1, "kind": "reference",
               "sample": "first %d sites of the same cohort, `bgt view -f'%s' -G`, 1 thread, best of 2, %.1f s" % (sample_rows, FILTER, best),
               "gpu_matches_reference_vcf": bool(ok)}
        cli = None
        if os.path.exists(DROPIN):   # the same query through the drop-in CLI: whole VCF, md5 against the reference's
            env = dict(os.environ, BGT_B200_ROUTE="1")
            mine = subprocess.run([DROPIN] + view_args() + [prefix], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            cli = {"md5_equals_reference_on_sample": bool(mine.returncode == 0 and hashlib.md5(mine.stdout).hexdigest() == hashlib.md5(out).hexdigest()),
                   "sample_sites": sample_rows, "served_by_device_pipeline": b"view_fast=1 view_fast_to_ref=0" in mine.stderr}
        return res, cli
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cli_end_to_end(host_img, rows, sample_check):
    """`e2e_cli`: the drop-in `bgt view -f'AC>0' -G <prefix>` (reference host application linked against the B200 seams) on the
    FULL database on disk (page cache warm), stdout piped to md5sum, wall clock including process start and CUDA context
    creation -- the number a user of the CLI sees."""
    from oracle import oracle as orc
    if not (os.path.exists(DROPIN) and os.path.exists(orc.MKSITES)):
        return {"unavailable": "integration/_build/bgt or oracle/_ref/mksites not built"}
    tmp = tempfile.mkdtemp(prefix="bgtb200_cli_", dir=tmp_root())
    try:
        prefix = os.path.join(tmp, "full.bgt")
        write_bgt(prefix, host_img)
        cmd = "'%s' view -f'%s' -G '%s' | md5sum" % (DROPIN, FILTER, prefix)
        best, digest, trace = None, None, ""
        for i in range(3):
            env = dict(os.environ, BGT_B200_TRACE="1") if i == 2 else dict(os.environ)
            t0 = time.perf_counter()
            r = subprocess.run(["bash", "-c", "set -o pipefail; " + cmd], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return {"error": r.stderr.decode()[-300:]}
            if i < 2:
                best = dt if best is None or dt < best else best
                digest = r.stdout.split()[0].decode()
            else:
                trace = " | ".join(" ".join(ln.split()[1:]) for ln in r.stderr.decode().splitlines() if ln.startswith("[view_fast]"))
        out = {"value": rows / best, "unit": "sites/s", "seconds": round(best, 3), "md5": digest, "sites": rows,
               "how": "`integration/_build/bgt view -f'%s' -G <prefix> | md5sum`, database on disk (page cache warm), wall clock of the whole pipeline incl. process start + CUDA context creation, best of 2" % FILTER,
               "phases_ms": trace}
        if sample_check:
            out.update(sample_check)
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ reference arm

def make_cohort_child(prefix, samples, rows, seed):
    """child process of the reference arm: the cohort is generated on the GPU here so that the arm's own process never loads
    this repo's library"""
    import bgt_b200
    with bgt_b200.Context(0) as ctx:
        cohort = bgt_b200.synth_cohort(ctx, samples, rows, seed=seed)
        write_bgt(prefix, cohort.image())
        cohort.close()


def run_reference(args, rank, world):
    """The reference's own CPU implementation on all host cores: P concurrent `bgt view` processes on disjoint,
    checkpoint-aligned row ranges of the same cohort shape."""
    if rank != 0:
        return
    from oracle import oracle as orc
    if not (orc.have_ref() and os.path.exists(orc.MKSITES)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present"}))
        return
    cores = os.cpu_count() or 1
    P = max(1, min(cores, 64, (args.rows + 8191) // 8192))
    total_steps = args.steps + args.warmup
    rpp = 8192
    while rpp > 512 and total_steps * rpp / 1300.0 > 150.0:   # keep the whole run within a few minutes (~1.3k sites/s/core)
        rpp //= 2
    tmp = tempfile.mkdtemp(prefix="bgtb200_ref_", dir=tmp_root())
    try:
        prefix = os.path.join(tmp, "c.bgt")
        # data generation only (untimed, in a child process); the timed path below is the reference CLI alone
        subprocess.run([sys.executable, os.path.abspath(__file__), "--make-cohort", prefix, "--samples", str(args.samples), "--rows", str(P * 8192),
                        "--seed", str(args.seed)], check=True)

        def one_step():
            procs = [subprocess.Popen([orc.REF_BGT] + view_args(i * 8192, i * 8192 + rpp) + [prefix], stdout=subprocess.DEVNULL) for i in range(P)]
            for p in procs:
                if p.wait() != 0:
                    raise RuntimeError("reference bgt view failed")
        for _ in range(args.warmup):
            one_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_step()
        dt = time.perf_counter() - t0
        value = P * rpp * args.steps / dt
        sample = "%d concurrent `bgt view -f'%s' -G -r` processes x %d sites each (checkpoint-aligned ranges of the same %d-sample cohort shape) per step" % (P, FILTER, rpp, args.samples)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic (device generator in a child process, untimed setup; seed %d)" % args.seed,
            "config": {"workload": workload_name(args.samples, args.rows), "sample": sample},
            "cpu_baseline": {"value": value, "unit": "sites/s", "cores": P, "kind": "reference", "sample": sample, "host_cores": cores},
            "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ this repo's arm

class Rig:
    """one rank: context, process group helpers"""

    def __init__(self, rank, world, local_rank):
        import torch
        import bgt_b200
        self.torch, self.b = torch, bgt_b200
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
        self.dev = torch.device("cuda", local_rank)
        self.ctx = bgt_b200.Context(local_rank)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, xs):
        t = self.torch.tensor([float(x) for x in xs], dtype=self.torch.float64, device=self.dev)
        if self.dist is None:
            return [t.tolist()]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.tolist() for o in out]

    def sum_i64(self, xs):
        t = self.torch.tensor([int(x) for x in xs], dtype=self.torch.int64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t.tolist()]


def e2e_step_fn(rig, host_img, n, q_kwargs, h_counts, h_pass, row_beg=0, hap_bits=None):
    """The end-to-end step through the public API: host .pbf image -> b200_pbf_load_ex (H2D in chunks; row index, row meta,
    start ranks, plane-1 view and composite maps built on the device behind each chunk) -> b200_scan -> host results."""
    b = rig.b

    m = b.pbf_peek(host_img)[0]

    def step():
        out = {"counts": h_counts, "passed": h_pass}
        qq = b.Query(rig.ctx, m, **q_kwargs)
        if hap_bits is None:   # count-only queries: load and scan as one pipeline (b200_pbf_load_scan)
            pb, _ = b.load_scan(rig.ctx, host_img, qq, row_beg, row_beg + n, out=out)
        else:
            pb = b.Pbf.from_bytes(rig.ctx, host_img, row_beg, row_beg + n)
            out["hap_bits"] = hap_bits
            b.scan(rig.ctx, pb, qq, row_beg, n, hap_bits=True, out=out)
        qq.close()
        pb.close()
    return step


def timed_wall(rig, step, steps, warm, median=False):
    """wall-clock seconds per end-to-end step (every step ends with its results on the host): the mean over `steps` steps, max
    over ranks -- or, for the side records of one GPU, the median step (one-off re-allocations of grow-only scratch do not count)"""
    for _ in range(warm):
        step()
    rig.barrier()
    per = []
    t0 = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        step()
        per.append(time.perf_counter() - t1)
        if os.environ.get("BENCH_DEBUG"):
            sys.stderr.write("[bench debug] e2e step %.2f ms (scan kernels %.2f ms)\n" % (per[-1] * 1e3, rig.ctx.last_ms(1)))
    rig.ctx.sync()
    s = rig.max_over_ranks(time.perf_counter() - t0)
    rig.barrier()
    if median:
        per.sort()
        return per[len(per) // 2]
    return s / steps


def resident_ms(rig, cohort, q, beg, n, d_counts, d_pass, steps, warm, d_hap=(0, 0)):
    b, ctx = rig.b, rig.ctx
    for _ in range(warm):
        b.scan_device(ctx, cohort, q, beg, n, d_counts, d_pass, d_hap)
        b.collect(ctx)
    ctx.sync()
    ctx.mark(0)
    tot = None
    for _ in range(steps):
        b.scan_device(ctx, cohort, q, beg, n, d_counts, d_pass, d_hap)
        tot = b.collect(ctx)
    ctx.mark(1)
    ctx.sync()
    return ctx.mark_elapsed_ms(0, 1) / steps, tot


def extra_configs(rig, cohort, host_img, samples, n):
    """BASELINE configs 3 and 4 on the same cohort: resident and end-to-end sites/s, first rows checked against the oracle."""
    import numpy as np
    from oracle import oracle as orc
    torch, b, ctx, dev = rig.torch, rig.b, rig.ctx, rig.dev
    out = {}
    check_rows = 192
    small = b.synth_cohort(ctx, samples, check_rows, seed=cohort_seed(rig))
    opb = orc.Pbf(small.image().tobytes())
    small.close()
    # ---- config 3: two sample groups (50/50), -f'AC1/AN1>0.1&&AC2==0'
    grp = (np.arange(samples) % 2 + 1).astype(np.uint32)
    flt3 = "AC1/AN1>0.1&&AC2==0"
    q3 = b.Query(ctx, cohort, group=grp, n_groups=2, flt=flt3)
    d_counts = torch.empty((n, q3.stride), dtype=torch.int32, device=dev)
    d_pass = torch.empty((n,), dtype=torch.uint8, device=dev)
    ms, _ = resident_ms(rig, cohort, q3, 0, n, d_counts.data_ptr(), d_pass.data_ptr(), 5, 2)
    want = opb.scan(0, check_rows, group=grp, n_groups=2, flt=flt3)
    ok = bool((d_counts[:check_rows].cpu().numpy() == want["counts"]).all() and (d_pass[:check_rows].cpu().numpy() == want["passed"]).all())
    h_counts = b.host_alloc(n * q3.stride * 4).view(np.int32).reshape(n, q3.stride)
    h_pass = b.host_alloc(n)
    e2e_s = timed_wall(rig, e2e_step_fn(rig, host_img, n, dict(group=grp, n_groups=2, flt=flt3), h_counts, h_pass), 5, 2, median=True)
    ok = ok and bool((h_counts[:check_rows] == want["counts"]).all())
    out["config3"] = {"workload": "two -s groups (50/50), -f'%s' -G" % flt3, "resident_sites_per_s": n / (ms * 1e-3), "resident_ms": ms,
                      "e2e_sites_per_s": n / e2e_s, "e2e_ms": e2e_s * 1e3, "matches_oracle_first_rows": ok, "oracle_rows": check_rows,
                      "d2h_bytes_per_step": int(h_counts.nbytes + n)}
    b.host_free(h_counts.reshape(-1).view(np.uint8)); b.host_free(h_pass)
    q3.close()
    del d_counts
    # ---- config 4: 200-sample subset extraction with genotypes (pbs_dec-shaped path): counts + two bit planes per site
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(samples, size=200, replace=False)).astype(np.int32)
    q4 = b.Query(ctx, cohort, out_samples=sel)
    d_counts = torch.empty((n, q4.stride), dtype=torch.int32, device=dev)
    d_hap = [torch.empty((n, q4.words), dtype=torch.int32, device=dev) for _ in range(2)]
    ms, _ = resident_ms(rig, cohort, q4, 0, n, d_counts.data_ptr(), d_pass.data_ptr(), 3, 1, (d_hap[0].data_ptr(), d_hap[1].data_ptr()))
    want = opb.scan(0, check_rows, out_samples=sel, want_hap=True)
    bits = [np.unpackbits(d_hap[p][:check_rows].cpu().numpy().view(np.uint8), axis=1, bitorder="little")[:, :400] for p in range(2)]
    ok = bool((d_counts[:check_rows].cpu().numpy() == want["counts"]).all() and (bits[0] == want["hap0"]).all() and (bits[1] == want["hap1"]).all())
    h_counts = b.host_alloc(n * q4.stride * 4).view(np.int32).reshape(n, q4.stride)
    h_pass = b.host_alloc(n)
    h_hap = [b.host_alloc(n * q4.words * 4).view(np.uint32).reshape(n, q4.words) for _ in range(2)]
    e2e_s = timed_wall(rig, e2e_step_fn(rig, host_img, n, dict(out_samples=sel), h_counts, h_pass, hap_bits=h_hap), 5, 2, median=True)
    out["config4"] = {"workload": "200-sample -s subset with genotypes (400 tracked haplotypes, bit planes out)", "resident_sites_per_s": n / (ms * 1e-3), "resident_ms": ms,
                      "e2e_sites_per_s": n / e2e_s, "e2e_ms": e2e_s * 1e3, "matches_oracle_first_rows": ok, "oracle_rows": check_rows,
                      "d2h_bytes_per_step": int(h_counts.nbytes + n + 2 * h_hap[0].nbytes)}
    for a in [h_counts.reshape(-1).view(np.uint8), h_pass] + [h.reshape(-1).view(np.uint8) for h in h_hap]:
        b.host_free(a)
    q4.close()
    opb.close()
    return out


_SEED = [0]


def cohort_seed(rig):
    return _SEED[0] + rig.rank


def p1_sweep(rig, samples, rows):
    """Resident sites/s against the density of plane 1 (missing / other-ALT codes): the split scan walks only the carriers of
    plane-1 bits, so its cost follows their number; blocks that exceed the pair-list capacity take the general walk."""
    torch, b, ctx, dev = rig.torch, rig.b, rig.ctx, rig.dev
    pts = []
    d_counts = torch.empty((rows, 6), dtype=torch.int32, device=dev)
    d_pass = torch.empty((rows,), dtype=torch.uint8, device=dev)
    for one_in, max_iv in ((64, 3), (16, 3), (4, 3), (2, 3), (1, 3), (16, 30), (4, 30), (1, 30)):
        pb = b.synth_cohort(ctx, samples, rows, seed=_SEED[0] + 1000 + one_in * 64 + max_iv, p1_one_in=one_in, p1_max_iv=max_iv)
        q = b.Query(ctx, pb, flt=FILTER)
        ms, tot = resident_ms(rig, pb, q, 0, rows, d_counts.data_ptr(), d_pass.data_ptr(), 3, 2)
        split = b.lib().b200_pbf_split_blocks(pb.h)
        nblk = (rows + 8191) // 8192
        missing = 2 * samples * rows - tot[0]
        pts.append({"p1_one_in": one_in, "p1_intervals": "1-%d" % max_iv, "sites_per_s": rows / (ms * 1e-3), "ms": ms,
                    "blocks_on_split_path": split, "blocks": nblk, "plane1_ones_per_block": round((missing + tot[2]) / nblk, 1)})
        q.close()
        pb.close()
    return {"rows": rows, "points": pts, "note": "plane 1 non-empty in one of p1_one_in rows with 1..k intervals of <= 64 ones; same shape otherwise"}


def h2d_probe(rig, mb=256):
    """host->device GB/s of every rank while ALL ranks copy at the same time (pinned memory): on this pool's 8-GPU boxes four
    GPUs share about 115 GB/s, so the end-to-end step -- which has to move its .pbf image -- cannot scale past it."""
    torch = rig.torch
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=rig.dev)
    d.copy_(h, non_blocking=True)
    rig.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(rig.dev)
    gbs = 6 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    rig.barrier()
    alone = None
    if rig.rank == 0:
        e0.record()
        for _ in range(6):
            d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(rig.dev)
        alone = 6 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    rig.barrier()
    allg = [round(x[0], 1) for x in rig.gather([gbs])]
    return {"concurrent_gbs_per_rank": allg, "rank0_alone_gbs": None if alone is None else round(alone, 1), "mb": mb}


def topo_affinity():
    """CPU / NUMA affinity of every GPU as `nvidia-smi topo -m` reports it: on this pool's boxes all GPUs share one NUMA node
    (0-31 / 0), so there is no NUMA-local placement of the pinned staging buffers to choose."""
    exe = shutil.which("nvidia-smi")
    if not exe:
        return None
    try:
        out = subprocess.run([exe, "topo", "-m"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
    except Exception:
        return None
    import re
    res = {}
    for ln in out.splitlines():
        ln = re.sub(r"\x1b\[[0-9;]*m", "", ln)
        f = ln.split("\t")
        if f and re.match(r"GPU\d+$", f[0].strip()):
            tail = [x.strip() for x in f[1:] if x.strip()]
            aff = [x for x in tail if re.match(r"^[0-9,\-]+$", x)]
            res[f[0].strip()] = {"cpu_affinity": aff[0] if aff else None, "numa_affinity": aff[1] if len(aff) > 1 else None}
    return res or None


def config5_record(rig, args):
    """BASELINE configs[4]: ONE synthetic cohort of 500 000 samples x 10 M sites (1221 checkpoint blocks), region-sharded: rank r
    takes blocks [r*ceil(1221/N), (r+1)*ceil(1221/N)) -- b200_pbf_load_ex(row_beg,row_end) on the file image -- and the totals
    are all-reduced.  Every rank generates the (deterministic) cohort on its own GPU, keeps only its shard's byte range of the
    file image in host memory (sparse mapping: header, shard, index record) and drops the rest."""
    import ctypes as C
    import numpy as np
    from bgt_b200.shard import shard_rows
    torch, b, ctx, dev, L = rig.torch, rig.b, rig.ctx, rig.dev, rig.b.lib()
    samples, rows = args.config5_samples, args.config5_rows
    t0 = time.perf_counter()
    cohort = b.synth_cohort(ctx, samples, rows, seed=args.seed + 5)
    gen_s = time.perf_counter() - t0
    beg, end = shard_rows(rows, 13, rig.rank, rig.world)
    n = end - beg
    fsize = L.b200_pbf_image_size(cohort.h)
    bb, be, ib = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    if L.b200_pbf_block_bytes(cohort.h, beg, end, C.byref(bb), C.byref(be), C.byref(ib)) != 0:
        raise RuntimeError(L.b200_strerror().decode())
    # sparse host image: untouched pages of the anonymous mapping cost nothing
    mm = mmap.mmap(-1, fsize + 4096)
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)     # 2 MB pages under the part that gets touched: the DMA engine walks far fewer translations
    except (AttributeError, OSError, ValueError):
        pass
    img = np.frombuffer(mm, dtype=np.uint8, count=fsize)
    base = img.ctypes.data
    for off, ln in ((0, 16), (bb.value, be.value - bb.value), (ib.value, fsize - ib.value)):
        if ln and L.b200_pbf_image_download_range(cohort.h, base + off, off, ln) != 0:
            raise RuntimeError(L.b200_strerror().decode())
    cohort.close()
    pin_beg = bb.value & ~4095
    pin_len = ((be.value + 4095) & ~4095) - pin_beg
    if L.b200_host_register(base + pin_beg, pin_len) != 0:
        raise RuntimeError(L.b200_strerror().decode())
    h_counts = b.host_alloc(max(n, 1) * 6 * 4).view(np.int32).reshape(max(n, 1), 6)
    h_pass = b.host_alloc(max(n, 1))
    step = e2e_step_fn(rig, img, n, dict(flt=FILTER), h_counts[:n], h_pass[:n], row_beg=beg)
    rec = {"workload": "ONE synthetic cohort of %d samples x %d sites, `view -f'%s' -G`, region-sharded over %d GPUs: %d checkpoint blocks per rank" % (
        samples, rows, FILTER, rig.world, (n + 8191) // 8192 if rig.rank == 0 else 0), "m_haplotypes": 2 * samples, "generate_s": round(gen_s, 1)}
    # (a) one shard alone on rank 0, the other GPUs idle: the 1-GPU figure of the same shape
    alone_e2e = alone_res = None
    if rig.rank == 0:
        step()
        t0 = time.perf_counter()
        for _ in range(2):
            step()
        alone_e2e = (time.perf_counter() - t0) / 2
    rig.barrier()
    # (b) all shards at the same time
    e2e_s = timed_wall(rig, step, 2, 1)
    # resident: this rank's shard loaded once, scans timed on the device
    pb = b.Pbf.from_bytes(ctx, img, beg, end, prepare_count_scan=True)
    q = b.Query(ctx, pb, flt=FILTER)
    d_counts = torch.empty((max(n, 1), 6), dtype=torch.int32, device=dev)
    d_pass = torch.empty((max(n, 1),), dtype=torch.uint8, device=dev)
    if rig.rank == 0:
        alone_res, _ = resident_ms(rig, pb, q, beg, n, d_counts.data_ptr(), d_pass.data_ptr(), 3, 2)
    rig.barrier()
    ms, tot = resident_ms(rig, pb, q, beg, n, d_counts.data_ptr(), d_pass.data_ptr(), 3, 2)
    ms = rig.max_over_ranks(ms)
    same = bool(int(h_counts[:n, 0].astype(np.int64).sum()) == tot[0] and int(h_pass[:n].sum()) == tot[3])
    per_rank = rig.gather(list(tot) + [n])
    summed = rig.sum_i64(list(tot) + [n])
    ok_sum = all(int(sum(p[i] for p in per_rank)) == summed[i] for i in range(5))
    q.close(); pb.close()
    L.b200_host_unregister(base + pin_beg)
    b.host_free(h_counts.reshape(-1).view(np.uint8)); b.host_free(h_pass)
    if rig.rank == 0:
        rec.update({
            "sites_total": summed[4], "resident_sites_per_s": summed[4] / (ms * 1e-3), "resident_ms": ms,
            "e2e_sites_per_s": summed[4] / e2e_s, "e2e_ms": e2e_s * 1e3,
            "one_gpu_same_shape": {"sites": n, "resident_sites_per_s": n / (alone_res * 1e-3), "e2e_sites_per_s": n / alone_e2e,
                                   "note": "rank 0 processing its shard while the other GPUs are idle; the whole cohort on ONE GPU is N such shards one after the other"},
            "speedup_over_one_gpu": {"resident": (summed[4] / ms) / (n / alone_res), "e2e": (summed[4] / e2e_s) / (n / alone_e2e)},
            "h2d_bytes_per_rank_per_step": int(be.value - bb.value), "e2e_matches_resident": same,
            "totals_allreduce": {"sum_AN": summed[0], "sum_AC": summed[1], "sites_passed": summed[3], "sites": summed[4], "equals_sum_of_rank_totals": bool(ok_sum)}})
    del img
    return rec


def read_traffic(kernel_ms):
    """roofline.traffic from the committed ncu capture of the CURRENT build (profiles/roofline_traffic.json, written by
    tools/ncu_traffic.py): dram bytes per launch of the dominant kernel -- refused when that capture's kernel duration is more
    than 25 % away from the CUDA-event time measured here (a capture of older code, or another workload)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, "profiles/roofline_traffic.json missing"
    ncu_ms = t.get("kernel_ms_ncu")
    if ncu_ms is None:
        return None, "capture has no duration"
    # ncu runs kernels cold-cache and serialised at unlocked clocks: allow 25 %
    if abs(ncu_ms - kernel_ms) > 0.25 * kernel_ms:
        return None, "stale capture: ncu %.3f ms vs %.3f ms measured now" % (ncu_ms, kernel_ms)
    return t.get("dram_bytes_per_launch"), t.get("source")


def run_b200(args, rank, world, local_rank):
    import numpy as np
    rig = Rig(rank, world, local_rank)
    torch, bgt_b200, ctx, dev, dist = rig.torch, rig.b, rig.ctx, rig.dev, rig.dist
    _SEED[0] = args.seed
    barrier, max_over_ranks = rig.barrier, rig.max_over_ranks
    n, samples = args.rows, args.samples
    # ---- untimed setup: this rank's region shard of the cohort, generated on the device, plus its host image
    cohort = bgt_b200.synth_cohort(ctx, samples, n, seed=args.seed + rank)
    img_bytes = bgt_b200.lib().b200_pbf_image_size(cohort.h)
    host_img = bgt_b200.host_alloc(img_bytes)
    cohort.image(out=host_img)
    q = bgt_b200.Query(ctx, cohort, flt=FILTER)
    stride = q.stride
    d_counts = torch.empty((n, stride), dtype=torch.int32, device=dev)
    d_pass = torch.empty((n,), dtype=torch.uint8, device=dev)
    algo_bytes = cohort.row_bytes(0, n, True) + 4 * stride * n          # bytes_in + bytes_out per launch
    torch.cuda.synchronize(dev)

    # ---- resident arm
    def step_resident():
        bgt_b200.scan_device(ctx, cohort, q, 0, n, d_counts.data_ptr(), d_pass.data_ptr())
        tot = bgt_b200.collect(ctx)
        return tot, ctx.last_ms(0) - ctx.last_ms(4) - ctx.last_ms(5), ctx.last_ms(4)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    ctx.sync()
    launches0 = ctx.launches
    if rank == 0:
        sampler.start()
    t_clock0 = time.perf_counter()
    ctx.mark(0)
    walk_ms, sel_ms, tot = [], [], None
    for _ in range(args.steps):
        tot, wms, sms = step_resident()
        walk_ms.append(wms)
        sel_ms.append(sms)
    ctx.mark(1)
    ctx.sync()
    dev_ms = ctx.mark_elapsed_ms(0, 1)
    launches = ctx.launches - launches0                                 # kernels of this library inside the timed region
    extra = 0
    if rank == 0:
        # nvidia-smi cannot sample faster than every few tens of ms: when the timed region is shorter than 1.5 s the
        # identical step keeps running (untimed) until the sampler has seen 1.5 s of this load
        t_end = t_clock0 + 1.5
        while time.perf_counter() < t_end:
            step_resident()
            extra += 1
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = "the %d timed steps (%.0f ms) + %d identical untimed steps" % (args.steps, dev_ms, extra)
    launches_per_step = launches // max(1, args.steps)
    dev_ms = max_over_ranks(dev_ms)
    value = world * n * args.steps / (dev_ms * 1e-3)

    # ---- end-to-end arm: host image in, host results out, every step
    h_counts = bgt_b200.host_alloc(n * stride * 4).view(np.int32).reshape(n, stride)
    h_pass = bgt_b200.host_alloc(n)
    step_e2e = e2e_step_fn(rig, host_img, n, dict(flt=FILTER), h_counts, h_pass)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_s = timed_wall(rig, step_e2e, e2e_steps, min(args.warmup, 2)) * e2e_steps
    e2e_value = world * n * e2e_steps / e2e_s
    nblk = (n + 8191) // 8192
    e2e_bytes = {"h2d": img_bytes + nblk * 20, "d2h": h_counts.nbytes + n}   # the image + the block table; the row index is built on the device
    same = bool((h_counts[:, 0].astype(np.int64).sum() == tot[0]) and int(h_pass.sum()) == tot[3])

    # ---- whole-cohort totals: the one collective of the path (tiny NCCL all-reduce, outside the timed region)
    totals = rig.sum_i64(list(tot) + [n])

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = sum(walk_ms) / len(walk_ms)
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = read_traffic(k_ms)
        line = {
            "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic (device generator, seed %d+rank; rows drawn in PBWT-rank space, truthful snapshots)" % args.seed,
            "config": {"workload": workload_name(samples, n),
                       "m_haplotypes": 2 * samples, "sites_per_gpu": n, "shards": world, "l2": "inputs (%.0f MB .pbf image per GPU) larger than L2" % (img_bytes / 1e6),
                       "resident_state": "`value` times b200_scan on a loaded PBF handle: the .pbf image plus what b200_pbf_load_ex derives from it on the device once per handle (row index, run totals, start ranks, two-sided composite maps of 32-row groups, the plane-1 (column,row) pair list) -- the HBM analogue of the warm page cache the reference is timed with. `e2e` builds all of that from the host image inside the timed region.",
                       "totals_allreduce": {"sum_AN": totals[0], "sum_AC": totals[1], "sites_passed": totals[3], "sites": totals[4]}},
            "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": int(e2e_bytes["h2d"]), "d2h_bytes_per_step": int(e2e_bytes["d2h"]),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "matches_resident": same,
                    "how": "host .pbf image (pinned) -> b200_pbf_load_scan: H2D in 16 chunks (short ones first); behind each chunk on the device: row index, row meta, start ranks, plane-1 view + pair select, composite maps, pair walk, AC/AN + verdict, D2H of the chunk's results to pinned host memory -- one pipeline, one synchronisation at the end"},
            "gpu_launches": int(launches), "gpu_launches_per_step": int(launches_per_step),
            "roofline": {"bound": "hbm", "kernel": "pbwt_pair_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": k_ms,
                         "equivalent_rank_updates_per_s": 2.0 * 2 * samples * n / (dev_ms / args.steps * 1e-3),
                         "other_kernels_ms": {"plane1_select_kernel": sum(sel_ms) / len(sel_ms)},
                         "other_kernels_note": "plane1_select_kernel does not depend on the query: it runs once per resident PBF (inside the load) and is 0 in resident steps",
                         "note": "dominant kernel = pbwt_pair_kernel (pairwalk.cu: plane-0 rank of every (column,row) pair that carries a plane-1 code, through two-sided composite maps staged by TMA and warp-built run tables); it is bound by issue slots / shared-memory look-ups, not HBM (SURVEY 8d); the HBM fraction is reported as asked"},
            "clocks": clocks,
        }
    # ---- the other BASELINE configs, the CLI, the density sweep (one GPU), or config 5 + the copy ceiling (several GPUs)
    if world == 1 and not args.quick:
        line["configs"] = extra_configs(rig, cohort, host_img, samples, n)
        if not args.no_cpu_baseline:
            srows = min(n, args.cpu_sample_rows)
            line["cpu_baseline"], cli_check = cpu_baseline_sample(ctx, samples, args.seed, srows, h_counts, h_pass)
            line["e2e_cli"] = cli_end_to_end(host_img, n, cli_check)
    bgt_b200.host_free(host_img)
    bgt_b200.host_free(h_counts.reshape(-1).view(np.uint8)); bgt_b200.host_free(h_pass)
    q.close()
    cohort.close()
    del d_counts, d_pass
    if world == 1 and not args.quick:
        line["p1_sweep"] = p1_sweep(rig, samples, min(n, args.sweep_rows))
    if world > 1:
        probe = h2d_probe(rig)
        rec5 = config5_record(rig, args) if not args.quick else None
        if rank == 0:
            probe["gpu_affinity"] = topo_affinity()
            probe["note"] = "the end-to-end step has to move its .pbf image over this link: e2e at N GPUs cannot exceed N x (concurrent GB/s of the slowest rank / rank0_alone_gbs) of the 1-GPU e2e"
            line["h2d_probe"] = probe
            line["config5"] = rec5
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample-rows", type=int, default=16384)
    ap.add_argument("--sweep-rows", type=int, default=262144)
    ap.add_argument("--config5-samples", type=int, default=500000)
    ap.add_argument("--config5-rows", type=int, default=10000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="only the headline arms (value, e2e)")
    ap.add_argument("--make-cohort", default=None, help="internal: write a synthetic BGT database with this prefix and exit")
    args = ap.parse_args()
    if args.make_cohort:
        make_cohort_child(args.make_cohort, args.samples, args.rows, args.seed)
        return
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.gpus > 1 and "RANK" not in os.environ:   # convenience: relaunch under torchrun, one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:])
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
