"""Per-kernel launch durations (us) from an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` launch list.

    python tools/launch_times.py gpurun_out/x.csv
"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
out = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    out[r[4].split('(')[0].replace('seq::', '')].append(round(float(r[-1]) / 1e3, 1))
for k, v in out.items():
    print(k, v)
