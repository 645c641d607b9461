"""Opcode mix and stall samples of an ncu report from its SASS source page, normalised per warp per RHS.
usage: python tools/ncu_ops.py report.ncu-rep <warps_per_cta * rhs_total>   (report captured with --import-source on)"""
import csv, collections, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
si, ie, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, samp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
for r in rows:
    if len(r) > ie and r[ie].isdigit():
        src = r[si].strip()
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        n = int(r[ie]); agg[op] += n; samp[op] += int(r[ss] or 0); tot += n; tots += int(r[ss] or 0)
print(f"instructions per unit: {tot / per:.1f}   stall samples: {tots}")
for op, n in agg.most_common(22):
    print(f"{op:8s} {n / per:8.1f}   samples {samp[op] / max(tots, 1) * 100:5.1f}%")
