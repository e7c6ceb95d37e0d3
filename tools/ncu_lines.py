"""Per-source-line hot spots of an ncu report: python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []
for x in rows:
    if len(x) > 7 and x[0].isdigit():
        try:
            lines.append((int(x[0]), x[1], int(x[4]), int(x[7])))
        except ValueError:
            pass
ts = sum(l[2] for l in lines) or 1
ti = sum(l[3] for l in lines) or 1
print("stall%%  inst%%  line  source   (total samples %d, warp instructions %d)" % (ts, ti))
for no, src, ws, ie in lines:
    if 100.0 * ws / ts >= minp or 100.0 * ie / ti >= minp:
        print("%5.1f  %5.1f  %4d  %s" % (100.0 * ws / ts, 100.0 * ie / ti, no, src.strip()[:120]))
