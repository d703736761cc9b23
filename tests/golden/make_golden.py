#!/usr/bin/env python
"""Generate the golden fixtures by EXECUTING THE REFERENCE's own functions, unmodified.

Runs only in the build container (needs /root/reference): ``python tests/golden/make_golden.py``.
For every case it renders seeded synthetic maps (improved_body_parts_b200/synth.py), runs
``find_peaks -> find_connections -> find_people`` lifted verbatim from /root/reference/evaluate.py
(oracle/ref_loader.py), and stores inputs + outputs in ``tests/golden/<case>.npz``.
The reference ships no tests or vectors of its own; these files are what pins parity.
Recorded environment: see ``tests/golden/MANIFEST.json`` (numpy / torch versions matter for f32-vs-f64
promotion).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from golden_io import save_case  # noqa: E402
from improved_body_parts_b200 import skeleton, synth  # noqa: E402
from oracle.ref_loader import DemoReference, Reference  # noqa: E402

# name -> (make_image kwargs, param overrides, paf dtype, image_extent override or None)
Q = dict(noise_levels=16)  # quantised noise keeps the fixtures compressible
CASES = {
    "clean_p10_128": (dict(seed=11, H=128, W=128, persons=10, **Q), {}, "f32", None),
    "clean_p30_128": (dict(seed=12, H=128, W=128, persons=30, **Q), {}, "f32", None),
    "clean_p8_f64_96x160": (dict(seed=13, H=96, W=160, persons=8, **Q), {}, "f64", None),
    "rect_p5_72x56": (dict(seed=14, H=72, W=56, persons=5, **Q), {}, "f32", None),
    "dropout_p16": (dict(seed=15, H=112, W=128, persons=16, drop_prob=0.25, **Q), {}, "f32", None),
    "plateau_spikes_p12": (dict(seed=16, H=96, W=96, persons=12, plateau=8, spikes=40, **Q), {}, "f32", None),
    "colocate_edge_p8": (dict(seed=17, H=64, W=80, persons=8, colocate=8, edge=True, **Q), {}, "f32", None),
    "missing_parts_p10": (dict(seed=18, H=96, W=96, persons=10, missing_parts=(4, 16), **Q), {}, "f32", None),
    "negative_stretch_p20": (dict(seed=19, H=128, W=128, persons=20, negative_bias=0.05, stretch=8, edge=True,
                                  drop_prob=0.1, **Q), {}, "f32", None),
    "weak_pruned_p10": (dict(seed=20, H=96, W=96, persons=10, heat_gain=0.3, paf_gain=0.3, drop_prob=0.5, **Q),
                        {}, "f32", None),
    "params_r1_mid10": (dict(seed=21, H=96, W=112, persons=9, **Q),
                        dict(offset_radius=1, mid_num=10, thre1=0.3, thre2=0.05), "f32", None),
    "params_r3_recon": (dict(seed=22, H=96, W=112, persons=12, drop_prob=0.3, spikes=20, **Q),
                        dict(offset_radius=3, connect_ration=0.7, remove_recon=1, len_rate=4.0), "f32", None),
    "crowd_dirty_p40": (dict(seed=23, H=128, W=128, persons=40, drop_prob=0.15, stretch=10, spikes=40, plateau=4,
                             colocate=5, edge=True, **Q), {}, "f32", None),
    "long_limbs_small_extent": (dict(seed=24, H=128, W=128, persons=6, scale_range=(2.2, 2.8), **Q), {}, "f32", 40),
    "empty": (dict(seed=25, H=64, W=64, persons=0, **Q), {}, "f32", None),
    "single_f64": (dict(seed=26, H=64, W=64, persons=1, **Q), {}, "f64", None),
    "tol_lenrate_p14": (dict(seed=27, H=112, W=112, persons=14, drop_prob=0.35, stretch=6, **Q),
                        dict(len_rate=1.5, connection_tole=1.2), "f32", None),
    # the 24-limb skeleton of config/config2.py: the limb table is runtime data everywhere
    "limbs24_p9": (dict(seed=28, H=96, W=104, persons=9, drop_prob=0.1, limbs=skeleton.LIMBS_24, **Q), {}, "f32", None),
    # demo_image.py's INLINED grouping copy (oracle/ref_loader.DemoReference): `>` at :288, length check at :414-415,
    # `count < 4` at :533.  The stored params carry crit1_strict / refresh_len_check / min_parts so that the checker and
    # the CUDA path are driven into the demo's behaviour; each case is also run through evaluate.py's functions and the
    # manifest records that the two really differ on it.
    "demo_default_p20": (dict(seed=32, H=96, W=112, persons=20, drop_prob=0.2, colocate=3, edge=True, **Q), {}, "f64", None,
                         "demo"),
    "demo_crit1_tie_p14": (dict(seed=43, H=96, W=112, persons=14, drop_prob=0.3, stretch=6, spikes=10, **Q),
                           dict(mid_num=10, thre2=0.3), "f32", None, "demo"),
    "demo_refresh_len_p14": (dict(seed=50, H=96, W=112, persons=14, drop_prob=0.3, stretch=6, spikes=10, **Q),
                             dict(len_rate=1.5, connection_tole=1.2), "f32", None, "demo"),
}


def main() -> None:
    import torch

    refs = {}
    ref = refs.setdefault((skeleton.LIMBS, "evaluate"), Reference())
    assert tuple(ref.limbs) == skeleton.LIMBS, "limb table drifted from the reference"
    only = set(sys.argv[1:])
    manifest = {"generated_by": "tests/golden/make_golden.py", "reference": "hellojialee/Improved-Body-Parts",
                "numpy": np.__version__, "torch": torch.__version__, "python": sys.version.split()[0], "cases": {}}
    if only and os.path.exists(os.path.join(HERE, "MANIFEST.json")):
        manifest["cases"] = json.load(open(os.path.join(HERE, "MANIFEST.json")))["cases"]
    for name, case in CASES.items():
        gen, over, dt, extent = case[:4]
        variant = case[4] if len(case) > 4 else "evaluate"
        if only and name not in only:
            continue
        gen = dict(gen)
        limbs = tuple(gen.get("limbs", skeleton.LIMBS))
        ref = refs.setdefault((limbs, variant), (DemoReference if variant == "demo" else Reference)(limbs=limbs))
        seed = gen.pop("seed")
        H, W, P = gen.pop("H"), gen.pop("W"), gen.pop("persons")
        heat, paf = synth.make_image(seed, H, W, P, **gen)
        if dt == "f64":  # the real predict() output is float64 (evaluate.py:85-86,160-161)
            paf = paf.astype(np.float64) * (1.0 + 2.0 ** -30) + 2.0 ** -40
        params = dict(skeleton.default_params(), **over)
        ext = H if extent is None else extent
        t0 = time.time()
        structs = ref.group(np.ascontiguousarray(heat.transpose(1, 2, 0)), np.ascontiguousarray(paf.transpose(1, 2, 0)),
                            ext, params)
        dt_s = time.time() - t0
        differs = None
        if variant == "demo":
            from parity import diff_structures
            ev = refs.setdefault((limbs, "evaluate"), Reference(limbs=limbs)).group(
                np.ascontiguousarray(heat.transpose(1, 2, 0)), np.ascontiguousarray(paf.transpose(1, 2, 0)), ext, params)
            differs = bool(diff_structures(ev, structs, 0.0))
            params = dict(params, crit1_strict=1, refresh_len_check=1, min_parts=4)  # drives checker + CUDA into the demo's behaviour
        path = os.path.join(HERE, name + ".npz")
        gen.pop("limbs", None)
        save_case(path, heat, paf, limbs, ext, params, structs,
                  meta=dict(seed=seed, H=H, W=W, persons=P, gen=gen, paf_dtype=dt, variant=variant))
        peaks, conn, special, subset, cand = structs
        manifest["cases"][name] = dict(
            peaks=int(sum(len(p) for p in peaks)),
            connections=int(sum(0 if isinstance(c, list) else c.shape[0] for c in conn)),
            special_k=len(special), persons=int(subset.shape[0]), reference_seconds=round(dt_s, 3),
            bytes=os.path.getsize(path), **({"source": "demo_image.py:185-536", "differs_from_evaluate_py": differs} if variant == "demo" else {}))
        print(f"{name:28s} {manifest['cases'][name]}")
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print("total bytes", sum(c["bytes"] for c in manifest["cases"].values()))


if __name__ == "__main__":
    main()
