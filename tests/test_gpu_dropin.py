"""GPU: the drop-in functions (improved_body_parts_b200/dropin.py) return the reference's own Python structures.

Each golden case is pushed through the three replaced call sites exactly as process() chains them
(evaluate.py:509-511), with the reference's HWC map layout, and compared with what the reference produced."""
import os
import types

import numpy as np
import pytest

from conftest import golden_paths
from golden_io import load_case
from parity import diff_structures

pytestmark = pytest.mark.gpu
GOLDENS = golden_paths()


@pytest.fixture(scope="module")
def dropin(cuda_device):
    from improved_body_parts_b200 import dropin as d
    d.configure(device=0)
    yield d
    d.configure()


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[:-4] for p in GOLDENS])
def test_three_call_sites_chain_like_process(dropin, path):
    case = load_case(path)
    dropin.configure(limbs=case["limbs"])  # evaluate.limbSeq (install() takes it from the module)
    heat_hwc = np.ascontiguousarray(case["heat"].transpose(1, 2, 0))
    paf_hwc = np.ascontiguousarray(case["paf"].transpose(1, 2, 0))
    params = case["params"]
    all_peaks = dropin.find_peaks(heat_hwc, params)
    connection_all, special_k = dropin.find_connections(all_peaks, paf_hwc, case["image_extent"], params)
    subset, candidate = dropin.find_people(connection_all, special_k, all_peaks, params)
    d = diff_structures(case["structs"], (all_peaks, connection_all, special_k, subset, candidate), float_tol=0.0)
    assert not d, "\n".join(d)
    # the python types the reference hands on: tuples with np scalars, lists, float64 arrays
    assert isinstance(all_peaks, list) and len(all_peaks) == 18 and isinstance(special_k, list)
    assert subset.dtype == np.float64 and subset.shape[1:] == (20, 2)


def test_fused_group_and_keypoints(dropin):
    case = load_case([p for p in GOLDENS if "clean_p10_128" in p][0])
    dropin.configure(limbs=case["limbs"])
    heat_hwc = np.ascontiguousarray(case["heat"].transpose(1, 2, 0))
    paf_hwc = np.ascontiguousarray(case["paf"].transpose(1, 2, 0))
    got = dropin.group(heat_hwc, paf_hwc, case["image_extent"], case["params"])
    assert not diff_structures(case["structs"], got, float_tol=0.0)
    kp = dropin.keypoints(got[3], got[4])
    assert len(kp) == got[3].shape[0] and all(len(pts) == 17 for pts, _ in kp)
    assert kp[0][1] == 1 - 1.0 / got[3][0, 18, 0]


def test_install_rebinds_an_evaluate_like_module(dropin):
    """Stand-in for `import evaluate` (the reference is not on the GPU box): a module whose process() resolves the
    three names at call time, like evaluate.py:509-511 does."""
    from improved_body_parts_b200 import skeleton

    mod = types.ModuleType("evaluate")
    mod.limbSeq = list(skeleton.LIMBS)

    def _missing(*a, **k):
        raise AssertionError("the original python function was called")

    mod.find_peaks = mod.find_connections = mod.find_people = _missing

    def process(heat, paf, params):
        all_peaks = mod.find_peaks(heat, params)
        connection_all, special_k = mod.find_connections(all_peaks, paf, heat.shape[0], params)
        return mod.find_people(connection_all, special_k, all_peaks, params)

    dropin.install(mod)
    case = load_case([p for p in GOLDENS if "rect_p5_72x56" in p][0])
    subset, candidate = process(np.ascontiguousarray(case["heat"].transpose(1, 2, 0)),
                                np.ascontiguousarray(case["paf"].transpose(1, 2, 0)), case["params"])
    assert np.array_equal(subset, case["structs"][3]) and np.array_equal(candidate, case["structs"][4])


def test_nms_seam_matches_torch_formula(dropin, cuda_device):
    """util.keypoint_heatmap_nms (utils/util.py:177-183) restated with torch ops vs the CUDA seam."""
    import torch
    import torch.nn.functional as F

    case = load_case([p for p in GOLDENS if "plateau_spikes_p12" in p][0])
    heat = torch.from_numpy(case["heat"])[None].to(cuda_device)
    pad = F.pad(heat, (1, 1, 1, 1), mode="reflect")
    hmax = F.max_pool2d(pad, (3, 3), stride=1, padding=0)
    want = heat * ((hmax == heat).float() * (heat >= 0.1).float())
    got = dropin.keypoint_heatmap_nms(heat, kernel=3, thre=0.1)
    assert torch.equal(got, want)
