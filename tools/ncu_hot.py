"""Top stall-sample instructions of one kernel from an .ncu-rep source page.
    python tools/ncu_hot.py rep.ncu-rep regex [min_pct]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
data = []
for r in rows[2:]:
    if len(r) <= isamp or r[0].startswith("Kernel"):
        break
    try:
        data.append((float(r[isamp]), r[isrc].strip(), float(r[iex]), r[iw] if iw else ""))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for i, (n, s, ex, w) in enumerate(data):
    if n >= tot * minp / 100:
        print(f"{i:4d} {n / tot * 100:5.1f}%  ex={ex:10.0f} wf={w:>10s}  {s}")
