#!/usr/bin/env python
"""Time one stage of the path in isolation on cuda:0 (development aid; bench.py is the reference measurement).
usage: python tools/time_stage.py [nms|score|match|assemble] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improved_body_parts_b200 import synth, skeleton
from improved_body_parts_b200.grouping import Grouper

stage = sys.argv[1] if len(sys.argv) > 1 else "score"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
NB = int(os.environ.get("TS_BATCH", 256))
heat, paf = synth.make_batch(20260921, NB, 128, 128, int(os.environ.get("TS_PERSONS", 30)))
dev = torch.device("cuda:0")
hd, pd = torch.from_numpy(heat).to(dev), torch.from_numpy(paf).to(dev)
g = Grouper(max_batch=NB)
P = skeleton.default_params()
g.group_device(hd, pd, 128, P)
fn = {"nms": lambda: g.nms_peaks(hd, P), "score": lambda: g.limb_score(pd, 128, P), "match": lambda: g.limb_match(NB, P),
      "assemble": lambda: g.assemble(NB, P)}[stage]
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"{stage} batch {NB} ms {e0.elapsed_time(e1) / iters:.4f}  env " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("SPG_")))
