"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py gpurun_out/launches.csv "command line that was profiled"
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    tot[name][0] += 1
    tot[name][1] += float(r[14])
total = sum(v[1] for v in tot.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {cmd} (cold-cache, serialised: compare SHARES)")
print(f"# unit ns; total {total:.0f}; launches {len(rows)}")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} launches {n:4d}  total {t:14.0f}  avg {t / n:12.0f}  share {100 * t / total:5.1f}%")
