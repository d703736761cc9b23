"""CPU: the Python/numpy port (oracle/grouping_port.py, the travelling "reference arm") against the goldens.

Bit-exact, floats included -- the port is the reference's algorithm in the reference's own language; it is what
bench.py times as the CPU baseline on the GPU box, where /root/reference does not exist.
"""
import os

import numpy as np
import pytest

from conftest import golden_paths
from golden_io import load_case
from oracle import grouping_port as gp
from parity import diff_structures

GOLDENS = golden_paths()


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[:-4] for p in GOLDENS])
def test_port_matches_reference_golden(path):
    case = load_case(path)
    params = dict(case["params"])
    got = gp.group_image(case["heat"], case["paf"], case["image_extent"], params, case["limbs"])
    diffs = diff_structures(case["structs"], got, float_tol=0.0)
    assert not diffs, "\n".join(diffs)


def test_port_to_coco_equals_checker():
    from improved_body_parts_b200 import skeleton
    from oracle import spg_oracle as so

    case = load_case([p for p in GOLDENS if "dropout_p16" in p][0])
    got = gp.group_image(case["heat"], case["paf"], case["image_extent"], case["params"], case["limbs"])
    people = gp.to_coco(got[3], got[4], skeleton.COCO_FROM_PART)
    res = so.group_batch(case["heat"][None], case["paf"][None], case["limbs"], case["image_extent"], case["params"])
    kp, sc = res.to_coco(0, skeleton.COCO_FROM_PART)
    assert len(people) == kp.shape[0]
    for j, (pts, score) in enumerate(people):
        assert np.array_equal(np.array(pts, np.float64), kp[j]) and score == sc[j]


def test_port_pool_runs_images_in_parallel():
    from improved_body_parts_b200 import skeleton, synth

    heat, paf = synth.make_batch(31, 4, 64, 64, 3)
    pool, run = gp.make_pool(heat, paf, 64, skeleton.default_params(), skeleton.LIMBS, workers=2)
    try:
        counts = run(range(4))
    finally:
        pool.close(); pool.join()
    serial = [gp.group_image(heat[i], paf[i], 64, skeleton.default_params(), skeleton.LIMBS)[3].shape[0] for i in range(4)]
    assert counts == serial
