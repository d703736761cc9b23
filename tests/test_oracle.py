"""CPU: the C restatement (oracle/spg_oracle.c) against the golden vectors produced by the reference itself.

Bit-exact, floats included (float_tol = 0): the oracle reproduces the reference's arithmetic, not just its
decisions.  These tests pin the checker; the GPU parity tests then compare the CUDA path with the checker.
"""
import os

import numpy as np
import pytest

from conftest import golden_paths
from golden_io import load_case
from oracle import spg_oracle as so
from parity import diff_structures, structure_stats

GOLDENS = golden_paths()


def test_goldens_present():
    assert len(GOLDENS) >= 15, "golden fixtures missing (tests/golden/*.npz)"


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[:-4] for p in GOLDENS])
def test_oracle_matches_reference_golden(path):
    case = load_case(path)
    res = so.group_batch(case["heat"][None], case["paf"][None], case["limbs"], case["image_extent"], case["params"])
    assert res.status[0] == 0
    got = res.as_reference_structures(0)
    diffs = diff_structures(case["structs"], got, float_tol=0.0)
    assert not diffs, "\n".join(diffs)
    assert structure_stats(got) == structure_stats(case["structs"])


def test_goldens_reach_the_rare_branches():
    """The fixture set as a whole exercises the branches clean skeletons never take."""
    so.cov_reset()
    for path in GOLDENS:
        case = load_case(path)
        so.group_batch(case["heat"][None], case["paf"][None], case["limbs"], case["image_extent"], case["params"])
    cov = so.cov_read()
    for name in ("norm0", "special_k", "mid_num1", "replace", "keep_old", "refresh", "assign_len_reject", "merge",
                 "merge_reject", "overlap", "recon_remove", "new_person", "pruned", "border_peak"):
        assert cov[name] > 0, f"no golden case reaches branch {name}: {cov}"


def test_threads_do_not_change_results():
    from improved_body_parts_b200 import skeleton, synth

    heat, paf = synth.make_batch(300, 6, 64, 64, 4)
    a = so.group_batch(heat, paf, skeleton.LIMBS, 64, skeleton.default_params(), threads=1)
    b = so.group_batch(heat, paf, skeleton.LIMBS, 64, skeleton.default_params(), threads=4)
    for f in ("px", "py", "pscore", "part_count", "conn_ij", "conn_score", "conn_count", "subset", "n_persons"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_capacity_overflow_is_reported():
    from improved_body_parts_b200 import skeleton, synth

    heat, paf = synth.make_batch(5, 1, 96, 96, 8)
    r = so.group_batch(heat, paf, skeleton.LIMBS, 96, skeleton.default_params(), cap_peaks=20)
    assert r.status[0] == so.ERR_CAPACITY
    r = so.group_batch(heat, paf, skeleton.LIMBS, 96, skeleton.default_params(), cap_rows=3)
    assert r.status[0] == so.ERR_CAPACITY


def test_to_coco_matches_process_tail():
    """evaluate.py:523-543 restated in numpy inline (tiny) vs spgo_to_coco."""
    from improved_body_parts_b200 import skeleton

    case = load_case([p for p in GOLDENS if "clean_p10_128" in p][0])
    res = so.group_batch(case["heat"][None], case["paf"][None], case["limbs"], case["image_extent"], case["params"])
    kp, sc = res.to_coco(0, skeleton.COCO_FROM_PART)
    _, _, _, subset, candidate = case["structs"]
    for j, s in enumerate(subset[..., 0]):
        coords = [(0, 0) if idx == -1 else tuple(candidate[int(idx)][:2]) for idx in s[:18]]
        coco = [None] * 17
        for dt, gt in skeleton.DT_GT_MAPPING.items():
            if gt is not None:
                coco[gt] = coords[dt]
        assert np.array_equal(kp[j], np.array(coco, np.float64))
        assert sc[j] == 1 - 1.0 / s[18]
