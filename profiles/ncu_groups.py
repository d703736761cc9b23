#!/usr/bin/env python
"""Group SASS lines of an .ncu-rep by execution count (≈ by loop) and show each group's share of instructions."""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(H)]
ci = H.index("Instructions Executed"); src = H.index("Source"); si = H.index("# Samples")
tot = sum(int(r[ci]) for r in data); ts = sum(int(r[si]) for r in data)
g = collections.defaultdict(lambda: [0, 0, 0, []])
for r in data:
    c = int(r[ci]); key = round(c, -len(str(c)) + 2) if c > 0 else 0   # 2 significant digits
    g[key][0] += 1; g[key][1] += c; g[key][2] += int(r[si]); g[key][3].append(r[src].strip().split()[0:2])
print(f"total {tot:,} warp-instr, {ts} samples")
for key, (n, c, s, ex) in sorted(g.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    ops = collections.Counter(" ".join(e[:1]) for e in ex).most_common(6)
    print(f"exec≈{key:>10,}  lines {n:4d}  instr {c:>12,} ({100*c/tot:5.1f}%)  samples {100*s/max(ts,1):5.1f}%  {ops}")
