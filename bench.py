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
Prints ONE JSON line on rank 0.
"""
import argparse
import json
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


def ref_view_cmd(prefix, row_beg=None, row_end=None):
    from oracle import oracle as orc
    cmd = [orc.REF_BGT, "view", "-f", FILTER, "-G"]
    if row_beg is not None:  # POS = 1000 + 10*row (mksites.c), inclusive 1-based region
        cmd += ["-r", "11:%d-%d" % (1000 + 10 * row_beg, 1000 + 10 * (row_end - 1))]
    return cmd + [prefix]


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


def cpu_baseline_sample(ctx, samples, seed, sample_rows, gpu_counts, gpu_pass):
    """Time the UNMODIFIED reference (oracle/_ref/bgt view -f'AC>0' -G, one thread) on the first sample_rows sites of
    the same cohort, and check its VCF against the GPU results of those sites."""
    import bgt_b200
    from oracle import oracle as orc
    if not (orc.have_ref() and os.path.exists(orc.MKSITES)):
        return {"value": None, "unit": "sites/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    tmp = tempfile.mkdtemp(prefix="bgtb200_")
    try:
        small = bgt_b200.synth_cohort(ctx, samples, sample_rows, seed=seed)   # rows are seeded per row: identical to the big cohort's first rows
        prefix = os.path.join(tmp, "s.bgt")
        write_bgt(prefix, small.image())
        small.close()
        best, out = None, b""
        for _ in range(2):
            t0 = time.perf_counter()
            out = subprocess.run(ref_view_cmd(prefix), stdout=subprocess.PIPE, check=True).stdout
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
        # the same query through the drop-in CLI (reference host application linked against the B200 seams): whole VCF, byte for byte
        dropin = os.path.join(ROOT, "integration", "_build", "bgt")
        if os.path.exists(dropin):
            t0 = time.perf_counter()
            mine = subprocess.run([dropin] + ref_view_cmd(prefix)[1:], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            res["dropin_cli"] = {"vcf_identical_to_reference": bool(mine.returncode == 0 and mine.stdout == out), "seconds": round(time.perf_counter() - t0, 3),
                                 "note": "process start + CUDA context creation included"}
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ arms

def run_reference(args, rank, world):
    """The reference's own CPU implementation on all host cores: P concurrent `bgt view` processes on disjoint,
    checkpoint-aligned row ranges of the same cohort shape."""
    if rank != 0:
        return
    from oracle import oracle as orc
    if not (orc.have_ref() and os.path.exists(orc.MKSITES)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present"}))
        return
    import bgt_b200
    cores = os.cpu_count() or 1
    P = max(1, min(cores, 64, (args.rows + 8191) // 8192))
    total_steps = args.steps + args.warmup
    rpp = 8192
    while rpp > 512 and total_steps * rpp / 1300.0 > 150.0:   # keep the whole run within a few minutes (~1.3k sites/s/core)
        rpp //= 2
    tmp = tempfile.mkdtemp(prefix="bgtb200_ref_")
    try:
        with bgt_b200.Context(0) as ctx:   # data generation only (untimed); the timed path below is the reference CLI
            cohort = bgt_b200.synth_cohort(ctx, args.samples, P * 8192, seed=args.seed)
            prefix = os.path.join(tmp, "c.bgt")
            write_bgt(prefix, cohort.image())
            cohort.close()

        def one_step():
            procs = [subprocess.Popen(ref_view_cmd(prefix, i * 8192, i * 8192 + rpp), stdout=subprocess.DEVNULL) for i in range(P)]
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
        sample = "%d concurrent `bgt view -f'%s' -G -r` processes x %d sites each (checkpoint-aligned ranges of a %d-sample cohort) per step" % (P, FILTER, rpp, args.samples)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic (B200 generator, untimed setup; seed %d)" % args.seed,
            "config": {"workload": "synthetic %d samples x %d sites, view -f'%s' -G full scan" % (args.samples, args.rows, FILTER), "sample": sample},
            "cpu_baseline": {"value": value, "unit": "sites/s", "cores": P, "kind": "reference", "sample": sample, "host_cores": cores},
            "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import bgt_b200

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = bgt_b200.Context(local_rank)
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
    if rank == 0:
        # nvidia-smi cannot sample faster than every few tens of ms: when the timed region is shorter than 1.5 s the
        # identical step keeps running (untimed) until the sampler has seen 1.5 s of this load
        t_end = t_clock0 + 1.5
        extra = 0
        while time.perf_counter() < t_end:
            step_resident()
            extra += 1
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = "the %d timed steps (%.0f ms) + %d identical untimed steps" % (args.steps, dev_ms, extra)
    launches = ctx.launches - launches0
    dev_ms = max_over_ranks(dev_ms)
    value = world * n * args.steps / (dev_ms * 1e-3)

    # ---- end-to-end arm: host image in, host results out, every step
    h_counts = bgt_b200.host_alloc(n * stride * 4).view(np.int32).reshape(n, stride)
    h_pass = bgt_b200.host_alloc(n)
    e2e_bytes = {}

    # The end-to-end step goes through the public API only.  The cohort is processed as region shards of whole
    # checkpoint blocks (b200_pbf_load row ranges) by two contexts (= two CUDA streams) on two host threads, so that
    # the H2D copy + index walk of one shard overlap the kernels of the previous one.
    ctx2 = bgt_b200.Context(local_rank)
    n_chunks = max(1, args.e2e_chunks)
    nblk = (n + 8191) // 8192
    per = (nblk + n_chunks - 1) // n_chunks
    ranges = [(i * per * 8192, min((i + 1) * per * 8192, n)) for i in range(n_chunks) if i * per * 8192 < n]

    def shard(cx, beg, end):
        pb = bgt_b200.Pbf.from_bytes(cx, host_img, beg, end, prepare_count_scan=True)
        qq = bgt_b200.Query(cx, pb, flt=FILTER)
        bgt_b200.scan(cx, pb, qq, beg, end - beg, out={"counts": h_counts[beg:end], "passed": h_pass[beg:end]})
        qq.close()
        pb.close()

    def step_e2e():
        errs = []

        def worker(cx, mine):
            try:
                for beg, end in mine:
                    shard(cx, beg, end)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        if len(ranges) == 1:
            worker(ctx, ranges)
        else:
            ths = [threading.Thread(target=worker, args=(cx, ranges[i::2])) for i, cx in enumerate((ctx, ctx2))]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        if errs:
            raise errs[0]
        e2e_bytes["h2d"] = img_bytes + nblk * 20      # the image + the block table (offsets, ends, rows per block); the row index is built on the device
        e2e_bytes["d2h"] = h_counts.nbytes + n

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    ctx.sync()
    ctx2.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * n * e2e_steps / e2e_s
    same = bool((h_counts[:, 0].astype(np.int64).sum() == tot[0]) and int(h_pass.sum()) == tot[3])

    # ---- whole-cohort totals: the one collective of the path (tiny NCCL all-reduce, outside the timed region)
    totals = torch.tensor(list(tot) + [n], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    totals = [int(x) for x in totals.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = sum(walk_ms) / len(walk_ms)
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get("pbwt_walk_kernel_dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic (device generator, seed %d+rank; rows drawn in PBWT-rank space, truthful snapshots)" % args.seed,
            "config": {"workload": "synthetic %d samples x %d sites per GPU, `view -f'%s' -G` full scan" % (samples, n, FILTER),
                       "m_haplotypes": 2 * samples, "sites_per_gpu": n, "shards": world, "l2": "inputs (%.0f MB .pbf image per GPU) larger than L2" % (img_bytes / 1e6),
                       "resident_state": "`value` times b200_scan on a loaded PBF handle: the .pbf image plus what b200_pbf_load_ex derives from it on the device once per handle (row index, run totals, start ranks, composite maps of 32-row groups, the plane-1 (column,row) pair list). `e2e` builds all of that from the host image inside the timed region.",
                       "totals_allreduce": {"sum_AN": totals[0], "sum_AC": totals[1], "sites_passed": totals[3], "sites": totals[4]}},
            "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": int(e2e_bytes["h2d"]), "d2h_bytes_per_step": int(e2e_bytes["d2h"]),
                    "steps": e2e_steps, "matches_resident": same,
                    "how": "host .pbf image (pinned) -> b200_pbf_load_ex (%d region shard(s); H2D in 16 chunks (short ones first); row index, row meta, start ranks, plane-1 view and composite maps built on the device behind each chunk, no host walk) -> b200_scan -> host AC/AN + verdicts" % len(ranges)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "pbwt_walk_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": k_ms,
                         "equivalent_rank_updates_per_s": 2.0 * 2 * samples * n / (dev_ms / args.steps * 1e-3),
                         "other_kernels_ms": {"plane1_select_kernel": sum(sel_ms) / len(sel_ms)},
                         "other_kernels_note": "plane1_select_kernel does not depend on the query: it runs once per resident PBF (inside the load; about 1.0 ms for the whole file) and is 0 in resident steps",
                         "note": "dominant kernel = pbwt_walk_kernel (QUERY mode of the split scan); it is bound by shared-memory run look-ups / issue slots, not HBM (SURVEY 8d); the HBM fraction is reported as asked"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            srows = min(n, args.cpu_sample_rows)
            line["cpu_baseline"] = cpu_baseline_sample(ctx, samples, args.seed, srows, h_counts, h_pass)
        print(json.dumps(line))
    bgt_b200.host_free(host_img)
    q.close()
    cohort.close()
    ctx2.close()
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
    ap.add_argument("--e2e-chunks", type=int, default=1)
    ap.add_argument("--cpu-sample-rows", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
