"""Per CUDA source line: executed warp instructions split by opcode class (from an ncu report's source page).
usage: python tools/ncu_lineops.py report.ncu-rep <divide_by> [top]"""
import csv, collections, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(io.StringIO(raw)))
fname, cur, agg, text = None, None, {}, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] == "Line No":
        continue
    if r[0].isdigit():
        cur = (fname, int(r[0])); text[cur] = r[1]; continue
    if r[0] == "" and cur and r[7].isdigit():
        src = r[3].strip()
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        cls = "fp64" if op in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU") else "ldst" if op in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL") else "ctl" if op in ("BRA", "BSSY", "BSYNC", "BAR", "WARPSYNC", "EXIT", "CALL", "RET") else "int"
        a = agg.setdefault(cur, collections.Counter())
        a[cls] += int(r[7]); a["all"] += int(r[7]); a["samples"] += int(r[6] or 0)
tot = sum(a["all"] for a in agg.values())
print(f"total {tot / per:.1f} per unit; classes: " + ", ".join(f"{c} {sum(a[c] for a in agg.values()) / per:.1f}" for c in ("fp64", "ldst", "ctl", "int")))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["all"])[:top]:
    print(f"{a['all'] / per:7.1f}  fp64 {a['fp64'] / per:6.1f} ldst {a['ldst'] / per:6.1f} ctl {a['ctl'] / per:6.1f} int {a['int'] / per:6.1f}  {k[0]}:{k[1]:<4} {text[k][:90]}")
