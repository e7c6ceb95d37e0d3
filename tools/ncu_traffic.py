"""profiles/roofline_traffic.json from an ncu report of the dominant kernel (bench.py reads it for roofline.traffic and refuses it
when its kernel duration disagrees with the CUDA-event time of the run).  python tools/ncu_traffic.py report.ncu-rep [kernel-regex]"""
import csv, json, os, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "pbwt_pair_kernel")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    return v * scale
res = None
for r in rows[2:]:
    if len(r) > col["Kernel Name"] and pat.search(r[col["Kernel Name"]]):
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        res = {"kernel": r[col["Kernel Name"]].split("(")[0], "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
               "kernel_ms_ncu": round(val(r, "gpu__time_duration.sum"), 4), "source": "profiles/%s (ncu --set full --clock-control none, tools/scan_probe.py = bench.py's default workload, one launch)" % os.path.basename(rep).replace(".ncu-rep", "_details.csv")}
        break
if res is None:
    sys.exit("kernel not found in " + rep)
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json"), "w"), indent=1)
