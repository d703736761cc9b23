#!/usr/bin/env python
"""Per-kernel SASS digest of libspgroup.so (CPU only: cuobjdump on the built library) -> profiles/<round>/sass_digest.txt.

Counts the mnemonics that show what the kernels use of the machine: the TMA engine's 1-D bulk copy (UBLKCP), mbarrier
traffic (SYNCS.*), warp reductions (REDUX), sleeps (NANOSLEEP), peer/system-scope memory operations (.SYS), FP64 work."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "improved_body_parts_b200", "libspgroup.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2", "sass_digest.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kernels = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*$", "", name).replace("void ", "")
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        c = kernels[name]
        c["instructions"] += 1
        for key, pat in (("UBLKCP", r"^UBLKCP"), ("UTMALDG", r"^UTMALDG"), ("SYNCS", r"^SYNCS"), ("SYNCS.TRYWAIT", r"^SYNCS.*TRYWAIT"),
                         ("REDUX", r"REDUX"), ("NANOSLEEP", r"^NANOSLEEP"), ("ATOMS", r"^ATOMS"), ("LDS", r"^LDS"), ("LDG", r"^LDG"),
                         ("STG", r"^STG"), ("DFMA/DADD/DMUL", r"^(DFMA|DADD|DMUL)"), ("MUFU", r"^MUFU"), ("SHFL", r"^SHFL"),
                         ("VOTE", r"^VOTE"), ("BAR", r"^BAR"), ("sys-scope", r"\.SYS")):
            if re.search(pat, op):
                c[key] += 1
os.makedirs(os.path.dirname(out_path), exist_ok=True)
cols = ["instructions", "UBLKCP", "UTMALDG", "SYNCS", "SYNCS.TRYWAIT", "NANOSLEEP", "REDUX", "ATOMS", "LDS", "LDG", "STG", "DFMA/DADD/DMUL",
        "MUFU", "SHFL", "VOTE", "BAR", "sys-scope"]
with open(out_path, "w") as fh:
    fh.write(f"# cuobjdump -sass improved_body_parts_b200/libspgroup.so  (static instruction counts per kernel; cubin arch: {', '.join(arch)})\n")
    fh.write("# UBLKCP = 1-D bulk copy on the TMA engine; UTMALDG = tensor-map TMA (not used: every staged tile is one contiguous span);\n")
    fh.write("# SYNCS = mbarrier operations; no HMMA/UTCMMA anywhere: the path has no dense contraction.\n")
    fh.write("kernel".ljust(62) + " ".join(c.rjust(8) for c in cols) + "\n")
    for k, c in kernels.items():
        fh.write(k[:61].ljust(62) + " ".join(str(c.get(col, 0)).rjust(8) for col in cols) + "\n")
    tensor = len(re.findall(r"\b(HMMA|UTCMMA|UTCHMMA|IMMA|QMMA)\b", sass))
    fh.write(f"# tensor-core instructions in the whole library: {tensor}\n")
print(open(out_path).read())
