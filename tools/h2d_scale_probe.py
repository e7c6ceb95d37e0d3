"""How does host->device bandwidth scale when 1, 2, 4, 8 GPUs of one box copy at the same time?  (SCALE_r01: the end-to-end
curve flattens at 4 GPUs.)  One process per GPU (like bench.py under torchrun) and, for comparison, one thread per GPU in one
process.  Prints GB/s per GPU and aggregate.  python tools/h2d_scale_probe.py [MB]"""
import os, sys, time, threading, subprocess, json
import torch

MB = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 330
ITERS = 8


def worker(dev, n_bytes, barrier, out, chunks=1):
    torch.cuda.set_device(dev)
    h = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h.fill_(dev + 1)
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda:%d" % dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        d.copy_(h, non_blocking=True)
    st.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        for _ in range(ITERS):
            if chunks == 1:
                d.copy_(h, non_blocking=True)
            else:
                per = n_bytes // chunks
                for c in range(chunks):
                    d[c * per:(c + 1) * per].copy_(h[c * per:(c + 1) * per], non_blocking=True)
    st.synchronize()
    dt = time.perf_counter() - t0
    out[dev] = n_bytes * ITERS / dt / 1e9


def threads_mode(n, chunks=1):
    bar = threading.Barrier(n)
    out = {}
    ths = [threading.Thread(target=worker, args=(i, MB << 20, bar, out, chunks)) for i in range(n)]
    for t in ths: t.start()
    for t in ths: t.join()
    return [round(out[i], 1) for i in range(n)]


def proc_child(dev, n, start_at):
    torch.cuda.set_device(dev)
    n_bytes = MB << 20
    h = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda:%d" % dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < start_at:
        pass
    t0 = time.perf_counter()
    for _ in range(ITERS):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    print(json.dumps({"dev": dev, "gbs": n_bytes * ITERS / (time.perf_counter() - t0) / 1e9}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        proc_child(int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]))
        sys.exit(0)
    ng = torch.cuda.device_count()
    print("gpus", ng, "cpus", os.cpu_count(), "MB", MB)
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True).stdout[:1500])
        print(subprocess.run(["lscpu"], stdout=subprocess.PIPE, text=True).stdout[:1800])
        print(subprocess.run(["bash", "-c", "numactl -H 2>/dev/null | head -20; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c"], stdout=subprocess.PIPE, text=True).stdout)
    except Exception as e:
        print("topo failed", e)
    for n in (1, 2, 4, 8):
        if n > ng: break
        r = threads_mode(n)
        print("threads  n=%d per-gpu GB/s %s aggregate %.1f" % (n, r, sum(r)))
    for n in (4, 8):
        if n > ng: break
        r = threads_mode(n, chunks=16)
        print("threads16chunks n=%d per-gpu GB/s %s aggregate %.1f" % (n, r, sum(r)))
    # specific subsets: which GPUs share an uplink?
    if ng >= 8:
        for pair in ((0, 1), (0, 2), (0, 4), (0, 7), (2, 3), (4, 5)):
            bar = threading.Barrier(2); out = {}
            ths = [threading.Thread(target=worker, args=(i, MB << 20, bar, out)) for i in pair]
            for t in ths: t.start()
            for t in ths: t.join()
            print("pair", pair, [round(out[i], 1) for i in pair])
    for n in (1, 2, 4, 8):
        if n > ng: break
        start = time.time() + 25
        ps = [subprocess.Popen([sys.executable, __file__, "child", str(i), str(n), str(start)], stdout=subprocess.PIPE, text=True) for i in range(n)]
        r = []
        for p in ps:
            o = p.communicate()[0].strip().splitlines()
            r.append(round(json.loads(o[-1])["gbs"], 1) if o else None)
        print("procs    n=%d per-gpu GB/s %s aggregate %.1f" % (n, r, sum(x for x in r if x)))
