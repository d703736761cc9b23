#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by executed count / stall samples (source page).
usage: python profiles/ncu_hot.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
path = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(H)]
ci = H.index("Instructions Executed"); si = H.index("# Samples"); src = H.index("Source")
tot_i = sum(int(r[ci]) for r in data); tot_s = sum(int(r[si]) for r in data)
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}  sass lines {len(data)}")
print("--- by instructions executed")
for r in sorted(data, key=lambda r: -int(r[ci]))[:N]:
    print(f"{int(r[ci]):>12,} {100*int(r[ci])/tot_i:5.1f}%  samp {100*int(r[si])/max(tot_s,1):5.1f}%  {r[src].strip()[:90]}")
print("--- by stall samples")
for r in sorted(data, key=lambda r: -int(r[si]))[:N]:
    print(f"{int(r[si]):>8,} {100*int(r[si])/max(tot_s,1):5.1f}%  inst {int(r[ci]):>12,}  {r[src].strip()[:90]}")
