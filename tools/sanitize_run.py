#!/usr/bin/env python
"""Small fused run for compute-sanitizer (memcheck / racecheck / synccheck): 6 dirty images through all four kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from improved_body_parts_b200 import synth, skeleton
from improved_body_parts_b200.grouping import Grouper
heat, paf = synth.make_batch(99, 6, 128, 128, 12, drop_prob=0.1, edge=True, spikes=10, colocate=2)
dev = torch.device("cuda:0")
g = Grouper(max_batch=6)
g.group_device(torch.from_numpy(heat).to(dev), torch.from_numpy(paf).to(dev), 128, skeleton.default_params())
r = g.fetch()
print("persons", r.n_persons.tolist(), "status", r.status.tolist(), "kernels", g.stage_kernels())
