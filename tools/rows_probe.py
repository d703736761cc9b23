import sys; sys.path.insert(0, ".")
import torch, numpy as np
from improved_body_parts_b200 import synth, skeleton
from improved_body_parts_b200.grouping import Grouper
dev = torch.device("cuda:0")
P = skeleton.default_params()
for rank in range(8):
    heat, paf = synth.make_batch(20260921 + rank * 256, 256, 128, 128, 30)
    hd, pd = torch.from_numpy(heat).to(dev), torch.from_numpy(paf).to(dev)
    for cap in (64, 48):
        g = Grouper(max_batch=256, max_person_rows=cap)
        g.group_device(hd, pd, 128, P)
        r = g.fetch()
        print("rank", rank, "capR", cap, "flagged images", int((r.status != 0).sum()), "max persons", int(r.n_persons.max()))
        g.close()
