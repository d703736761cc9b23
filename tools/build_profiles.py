#!/usr/bin/env python
"""Turn gpurun_out/final/ (tools/capture_profiles.sh) into the tracked artefacts under profiles/r1_final/ + traffic.json."""
import collections, csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(ROOT, "gpurun_out", "final"), os.path.join(ROOT, "profiles", "r1_final")
os.makedirs(dst, exist_ok=True)
for f in ("bench_n1.json", "bench_reference_arm.json", "launches.csv", "gpu.csv", "cpu.txt"):
    shutil.copy(os.path.join(src, f), dst)
traffic = {}
for k in ("nms_peaks", "limb_score", "limb_match", "assemble"):
    rep = os.path.join(src, f"prof_{k}.ncu-rep")
    with open(os.path.join(dst, f"ncu_{k}.txt"), "w") as fh:
        for tool, n in (("ncu_summary.py", None), ("ncu_groups.py", "10"), ("ncu_hot.py", "20")):
            cmd = [sys.executable, os.path.join(ROOT, "profiles", tool), rep] + ([n] if n else [])
            fh.write(subprocess.run(cmd, capture_output=True, text=True).stdout)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out))); H, U, V = rows[0], rows[1], rows[2]
    d, u = dict(zip(H, V)), dict(zip(H, U))
    b = lambda n: float(d[n].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[n]]
    name = d["Kernel Name"].split("(")[0].replace("void ", "").strip()
    traffic[name] = int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
    print(f"{name:32s} {d['gpu__time_duration.sum']:>10s} {u['gpu__time_duration.sum']}  dram {traffic[name]/1e6:.1f} MB")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
rows = [r for r in csv.reader(open(os.path.join(src, "launches.csv"))) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]; Hh = rows[hdr]
ki, vi = Hh.index("Kernel Name"), Hh.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    agg[r[ki].split("(")[0].replace("void ", "")].append(float(r[vi].replace(",", "")))
full = {k: max(v) for k, v in agg.items()}  # full-batch launches (the e2e leg's chunks are smaller)
tot = sum(full.values())
d = json.load(open(os.path.join(src, "bench_n1.json")))
ev = {k: v["ms"] for k, v in d["kernels"].items()}; evt = sum(ev.values())
print("shares  ncu launch list vs bench events:")
for k in full:
    print(f"  {k:32s} {100*full[k]/tot:5.1f} %   {100*ev.get(k, 0)/evt:5.1f} %")
print("bench:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["roofline"]["traffic"])
