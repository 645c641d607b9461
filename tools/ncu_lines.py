"""Aggregate the warp-stall samples of an ncu report by CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top=40]   (report captured with --import-source on, code built with -lineinfo)"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; si = r.index("# Samples"); ie = r.index("Instructions Executed"); continue
    if hdr and len(r) > si and r[0].isdigit() and r[2] == "-":
        key = (fname, int(r[0]))
        s = float(r[si] or 0); i = float(r[ie] or 0)
        a = agg.setdefault(key, [0.0, 0.0, r[1]])
        a[0] += s; a[1] += i
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]/tot*100:5.1f}%  inst {a[1]/toti*100:5.1f}%  {f}:{ln:<4} {a[2][:100]}")
