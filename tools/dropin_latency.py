#!/usr/bin/env python
"""Single-image latency of the drop-in functions on cuda:0 (VERDICT r1 weak #8 asked for it at 512x512): the three call
sites chained as evaluate.py:509-511 does, and the fused dropin.group, on host maps (HWC, as predict() returns them) and
on DeviceMaps.  usage: python tools/dropin_latency.py [H] [persons]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from improved_body_parts_b200 import dropin, skeleton, synth

H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
P = int(sys.argv[2]) if len(sys.argv) > 2 else 14
k = H / 128.0
heat, paf = synth.make_image(99, H, H, P, scale_range=(0.8 * k, 1.3 * k), sigma_scale=k, noise=0.0 if k > 1 else 0.02)
hw = np.ascontiguousarray(heat.transpose(1, 2, 0)).astype(np.float64)   # predict() returns float64 HWC arrays
pw = np.ascontiguousarray(paf.transpose(1, 2, 0)).astype(np.float64)
params = skeleton.default_params()
dropin.configure(device=0)


def chain(h, p):
    peaks = dropin.find_peaks(h, params)
    conns, special = dropin.find_connections(peaks, p, H, params)
    return dropin.find_people(conns, special, peaks, params)


def timeit(fn, n=20):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


dev_h = dropin.DeviceMaps(torch.from_numpy(heat)[None].cuda(), False)
dev_p = dropin.DeviceMaps(torch.from_numpy(paf)[None].cuda(), True)
out = {"H": H, "persons_found": int(chain(hw, pw)[0].shape[0]),
       "three_call_sites_host_maps_ms": timeit(lambda: chain(hw, pw)),
       "fused_group_host_maps_ms": timeit(lambda: dropin.group(hw, pw, H, params)),
       "three_call_sites_device_maps_ms": timeit(lambda: chain(dev_h, dev_p)),
       "fused_group_device_maps_ms": timeit(lambda: dropin.group(dev_h, dev_p, H, params)),
       "note": "host maps: [H,W,C] float64 arrays as predict() returns them (transpose + H2D inside); includes the Python "
               "reconstruction of the reference's list/ndarray structures"}
print(json.dumps(out))
