#!/usr/bin/env python
"""Clock trace of match_assemble_kernel on the bench workload (development aid).  Needs the trace build
(`make -C improved_body_parts_b200/csrc trace`); prints, per traced CTA (= image), when each matcher warp worked on each
limb and when the assembler waited for / consumed it.  usage: python tools/trace_match_assemble.py [persons] [out.json]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from improved_body_parts_b200 import grouping, skeleton, synth

grouping.LIB_PATH = os.path.join(ROOT, "improved_body_parts_b200", "libspgroup_trace.so")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 30
OUT = sys.argv[2] if len(sys.argv) > 2 else None
NB, CTAS, SLOTS, L = 256, 64, 1024, 30
heat, paf = synth.make_batch(20260921, NB, 128, 128, P)
dev = torch.device("cuda:0")
hd, pd = torch.from_numpy(heat).to(dev), torch.from_numpy(paf).to(dev)
prm = skeleton.default_params()
g = grouping.Grouper(max_batch=NB, max_person_rows=64)
lib = grouping.load_library()
lib.spg_trace_read.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
buf = np.zeros(CTAS * SLOTS, dtype=np.uint64)
for _ in range(3):
    g.group_device(hd, pd, 128, prm)
torch.cuda.synchronize()
assert lib.spg_trace_read(buf.ctypes.data, buf.size, 1) == 0
if os.environ.get("TRACE_ALONE"):
    g.match_assemble(NB, prm)      # the fused kernel alone (candidates wherever the last pass left them)
else:
    g.group_device(hd, pd, 128, prm)   # inside the path: right after limb_score
torch.cuda.synchronize()
assert lib.spg_trace_read(buf.ctypes.data, buf.size, 1) == 0
tr = buf.reshape(CTAS, SLOTS).astype(np.int64)
res = []
for b in range(CTAS):
    t = tr[b]
    t0 = t[0]
    rel = lambda s: int(t[s] - t0) if t[s] else None
    limbs = []
    for k in range(L):
        m = [rel(16 + 4 * k + i) for i in range(4)]
        a = [rel(160 + 4 * k + i) for i in range(3)]
        rc = int(t[160 + 4 * k + 3])
        limbs.append({"k": k, "match_start": m[0], "match_loaded": m[1], "match_done": m[2], "match_published": m[3],
                      "asm_wait": a[0], "asm_acquired": a[1], "asm_done": a[2], "rounds": rc >> 8, "conns": rc & 255})
    res.append({"cta": b, "prologue": rel(1), "loop_done": rel(300), "outputs_done": rel(301), "limbs": limbs})
if OUT:
    json.dump(res, open(OUT, "w"))
# summary over the traced CTAs
def col(f):
    return np.array([[f(l) for l in r["limbs"]] for r in res], dtype=np.float64)
nz = lambda x: 0 if x is None else x
wait = col(lambda l: nz(l["asm_acquired"]) - nz(l["asm_wait"]))
work = col(lambda l: nz(l["asm_done"]) - nz(l["asm_acquired"]))
mload = col(lambda l: nz(l["match_loaded"]) - nz(l["match_start"]))
mrun = col(lambda l: nz(l["match_done"]) - nz(l["match_loaded"]))
mpub = col(lambda l: nz(l["match_published"]) - nz(l["match_done"]))
rounds = col(lambda l: l["rounds"]); conns = col(lambda l: l["conns"])
print("cycles (mean over %d CTAs): prologue %.0f  loop_done %.0f  outputs_done %.0f" % (
    CTAS, np.mean([r["prologue"] for r in res]), np.mean([r["loop_done"] for r in res]), np.mean([r["outputs_done"] for r in res])))
print("assembler per image: wait %.0f  work %.0f   | per limb work %.0f, rounds %.2f, conns %.1f, cycles/round %.0f" % (
    wait.sum(1).mean(), work.sum(1).mean(), work.mean(), rounds.mean(), conns.mean(), work.sum() / max(rounds.sum(), 1)))
print("matcher per limb: load %.0f  rounds %.0f  publish %.0f   (conns %.1f -> %.0f cycles per accepted row)" % (
    mload.mean(), mrun.mean(), mpub.mean(), conns.mean(), mrun.sum() / max(conns.sum(), 1)))
print("matcher detail, CTA 0 (cycles from the limb's start): init, after round 1..4, rounds done, rows out | rounds, candidates")
for k in range(L):
    t = tr[0]; b = 400 + 8 * k; st = t[16 + 4 * k]
    f = lambda v: int(v - st) if v else -1
    print("%3d | %6d | %6d %6d %6d %6d | %6d %6d | %d rounds, %d candidates" % (k, f(t[b]), f(t[b + 1]), f(t[b + 2]), f(t[b + 3]), f(t[b + 4]), f(t[b + 5]), f(t[b + 6]), int(t[b + 7]) >> 10, int(t[b + 7]) & 1023))
print("limb  wait  work rounds conns | m_start m_loaded m_done m_pub   (CTA 0)")
for l in res[0]["limbs"]:
    print("%3d %6d %6d %3d %3d | %7s %7s %7s %7s | asm %7s %7s %7s" % (
        l["k"], nz(l["asm_acquired"]) - nz(l["asm_wait"]), nz(l["asm_done"]) - nz(l["asm_acquired"]), l["rounds"], l["conns"],
        l["match_start"], l["match_loaded"], l["match_done"], l["match_published"], l["asm_wait"], l["asm_acquired"], l["asm_done"]))
