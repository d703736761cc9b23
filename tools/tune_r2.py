#!/usr/bin/env python
"""Tuning sweep on cuda:0 (development aid; bench.py is the reference measurement): stage times of the bench workload
under the environment switches libspgroup.so reads at spg_create.  usage: python tools/tune_r2.py [persons]"""
import itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improved_body_parts_b200 import skeleton, synth
from improved_body_parts_b200 import grouping
from improved_body_parts_b200.grouping import Grouper
if os.environ.get("SPG_LIB"):   # a build variant (make variants / make trace), by path
    grouping.LIB_PATH = os.environ["SPG_LIB"]

P = int(sys.argv[1]) if len(sys.argv) > 1 else 30
NB = 256
heat, paf = synth.make_batch(20260921, NB, 128, 128, P)
dev = torch.device("cuda:0")
hd, pd = torch.from_numpy(heat).to(dev), torch.from_numpy(paf).to(dev)
prm = skeleton.default_params()


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def measure(env):
    for k in ("SPG_MA_WARPS", "SPG_EXACT_WARPS", "SPG_FUSE_MA", "SPG_PERSIST"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    g = Grouper(max_batch=NB, max_person_rows=64)
    g.group_device(hd, pd, 128, prm)
    dt = g.device_tensors()
    surv = float(dt["surv_count"][:NB].float().mean().item())
    cand = float(dt["cand_count"][:NB].float().mean().item())
    r = {"env": env, "lib": os.path.basename(grouping.LIB_PATH), "survivors_per_item": round(surv, 1), "candidates_per_item": round(cand, 1),
         "nms": timeit(lambda: g.nms_peaks(hd, prm)), "score": timeit(lambda: g.limb_score(pd, 128, prm)),
         "match": timeit(lambda: g.limb_match(NB, prm)), "assemble": timeit(lambda: g.assemble(NB, prm)),
         "match_assemble": timeit(lambda: g.match_assemble(NB, prm)), "path": timeit(lambda: g.group_device(hd, pd, 128, prm))}
    g.close()
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
    return r


measure({})
if len(sys.argv) > 2:   # "quick": the default only
    measure({"SPG_FUSE_MA": 0})
    sys.exit(0)
for ew in (10, 14):
    measure({"SPG_EXACT_WARPS": ew})
for mw in (2, 3, 4, 5, 6, 10, 15):
    measure({"SPG_MA_WARPS": mw})
measure({"SPG_FUSE_MA": 0})
measure({})
