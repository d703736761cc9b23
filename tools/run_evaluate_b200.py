#!/usr/bin/env python
"""Run the reference's ``evaluate.py`` UNCHANGED with its grouping stage on the B200 path.

    python tools/run_evaluate_b200.py --reference /path/to/Improved-Body-Parts [--config utils/config] [--check]

What it does -- the reference checkout is never modified:

1. stubs the modules ``evaluate.py`` imports but this host lacks (``pycocotools``, ``matplotlib``, ``configobj``,
   ``apex``) -- only the missing ones, and only as empty shells;
2. imports ``evaluate`` with a clean ``sys.argv`` (its module body runs argparse, ``/root/reference/evaluate.py:48``,
   and ``GetConfig``, ``:52``) and restores ``CUDA_VISIBLE_DEVICES``, which the module pins to "0" (``:28``) -- one
   process per GPU needs its own device;
3. ``dropin.install(evaluate)``: rebinds ``evaluate.find_peaks / find_connections / find_people`` (looked up by name
   at the call sites ``:509-511``) and takes ``limbSeq`` from the module (``:54``);
4. fills the globals ``evaluate.__main__`` would set (``:643-646``): ``params, model_params`` from the reference's own
   ``utils/config`` through ``skeleton.read_reference_ini`` (``utils/config_reader.py:7`` hard-codes the author's path),
   ``show_eval_speed``;
5. optionally replaces ``format_results`` (``:563-582``) by ``wire.format_results`` (same file contents).

``prepare()`` returns the module; building ``evaluate.posenet`` (``:626-641``: checkpoint + apex amp) and calling
``evaluate.validation(...)`` is then exactly what ``evaluate.__main__`` does.  With ``--check`` the launcher runs one
synthetic image through ``evaluate``'s own call sequence (``:509-511``) on the GPU and prints what it found.
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_OPTIONAL = ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "matplotlib", "matplotlib.pyplot", "configobj",
             "apex")


def _stub_missing() -> list:
    stubbed = []
    for name in _OPTIONAL:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__spg_stub__ = True
            sys.modules[name] = mod
            parent, _, child = name.rpartition(".")
            if parent:
                setattr(sys.modules[parent], child, mod)
            stubbed.append(name)
    for name, attr in (("pycocotools.coco", "COCO"), ("pycocotools.cocoeval", "COCOeval"), ("configobj", "ConfigObj")):
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__spg_stub__", False) and not hasattr(m, attr):
            setattr(m, attr, dict if attr == "ConfigObj" else object)
    return stubbed


def prepare(reference_root: str, config_path: str = None, device: int = None, install: bool = True,
            replace_format_results: bool = False):
    """Import the reference's ``evaluate`` module (unchanged) and put the B200 grouping path behind its call sites."""
    reference_root = os.path.abspath(reference_root)
    if not os.path.isfile(os.path.join(reference_root, "evaluate.py")):
        raise FileNotFoundError(f"no evaluate.py under {reference_root}")
    for p in (ROOT, reference_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    stubbed = _stub_missing()
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    argv, sys.argv = sys.argv, [os.path.join(reference_root, "evaluate.py")]
    cwd = os.getcwd()
    try:
        os.chdir(reference_root)  # the module appends ".." to sys.path and uses relative data paths
        evaluate = importlib.import_module("evaluate")
    finally:
        os.chdir(cwd)
        sys.argv = argv
        if visible is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)  # evaluate.py:28 pinned it to "0"
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = visible
    from improved_body_parts_b200 import dropin, skeleton, wire

    if install:
        if device is not None:
            dropin.configure(device=device)
        dropin.install(evaluate)
    evaluate.params, evaluate.model_params = skeleton.read_reference_ini(
        config_path or os.path.join(reference_root, "utils", "config"))
    evaluate.show_eval_speed = False
    if replace_format_results:
        evaluate.format_results = wire.format_results
    evaluate.__spg_stubbed__ = stubbed
    return evaluate


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--reference", default=os.environ.get("SPG_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--config", default=None, help="the reference's utils/config INI (default: <reference>/utils/config)")
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--check", action="store_true", help="group one synthetic image through evaluate's call sites on the GPU")
    a = ap.parse_args()
    ev = prepare(a.reference, a.config, a.device)
    print(f"evaluate imported from {ev.__file__}; stubbed: {ev.__spg_stubbed__}; limbs: {len(ev.limbSeq)}; "
          f"find_peaks -> {ev.find_peaks.__module__}.{ev.find_peaks.__name__}")
    if a.check:
        import numpy as np
        from improved_body_parts_b200 import synth
        heat, paf = synth.make_image(7, 128, 128, 6)
        hw, pw = np.ascontiguousarray(heat.transpose(1, 2, 0)), np.ascontiguousarray(paf.transpose(1, 2, 0))
        peaks = ev.find_peaks(hw, ev.params)                                   # evaluate.py:509
        conns, special = ev.find_connections(peaks, pw, hw.shape[0], ev.params)  # :510
        subset, candidate = ev.find_people(conns, special, peaks, ev.params)    # :511
        print(f"check: {sum(len(p) for p in peaks)} peaks, {sum(len(c) for c in conns if len(c))} connections, "
              f"{len(subset)} persons")


if __name__ == "__main__":
    main()
