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


@pytest.mark.parametrize("name", ["dropout_p16", "clean_p8_f64_96x160", "missing_parts_p10"])
def test_call_sites_with_foreign_objects_upload_their_arguments(dropin, name):
    """The stage functions skip the upload only when they are handed the very objects the previous stage returned
    (evaluate.py:509-511); copies -- or results of the reference's own Python functions -- go through the upload path."""
    import copy
    case = load_case([p for p in GOLDENS if name in p][0])
    dropin.configure(limbs=case["limbs"])
    heat_hwc = np.ascontiguousarray(case["heat"].transpose(1, 2, 0))
    paf_hwc = np.ascontiguousarray(case["paf"].transpose(1, 2, 0))
    params = case["params"]
    ref_peaks, ref_conns, ref_special, _, _ = case["structs"]
    # stages fed with the REFERENCE's outputs of the stage before
    conns, special = dropin.find_connections(copy.deepcopy(ref_peaks), paf_hwc, case["image_extent"], params)
    subset, candidate = dropin.find_people(copy.deepcopy(ref_conns), list(ref_special), copy.deepcopy(ref_peaks), params)
    d = diff_structures(case["structs"], (dropin.find_peaks(heat_hwc, params), conns, special, subset, candidate), float_tol=0.0)
    assert not d, "\n".join(d)


def test_device_predict_feeds_the_call_sites_without_leaving_the_gpu(dropin, cuda_device):
    """dropin.predict (evaluate.py:83-166 with the post-network stage on the device) -> DeviceMaps -> the three call
    sites, against the CPU pipeline on the same network output: the reference's own pre-processing lines (cv2 resize,
    padRightDownCorner restated), oracle/postnet_port.py for :126-161, the C checker for the grouping.
    The 'network' is a stand-in that answers with a fixed tensor of the right shape (the IMHN itself is row f-3)."""
    import cv2
    import torch
    from improved_body_parts_b200 import skeleton, synth
    from oracle import postnet_port as pp
    from oracle import spg_oracle as so
    from test_gpu_postnet import _network_like_output

    class E:  # what _network_like_output needs
        pass
    e = E(); e.synth, e.skeleton = synth, skeleton
    rng = np.random.default_rng(3)
    image = rng.integers(0, 255, size=(150, 210, 3), dtype=np.uint8)  # 150x210 image, boxsize 160 -> scale 1.0667
    params = dict(skeleton.default_params(), scale_search=[1.0], rotation_search=[0.0])
    model_params = dict(boxsize=160, stride=4, max_downsample=64, padValue=128)
    scale = 1.0 * 160 / 150
    resized = cv2.resize(image, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    padded, pad = dropin.pad_right_down_corner(resized, 64, 128)
    assert padded.shape[0] % 64 == 0 and padded.shape[1] % 64 == 0 and padded[-1, -1, 0] == 128
    h, w = padded.shape[0] // 4, padded.shape[1] // 4
    out = _network_like_output(e, 4321, 1, h, w, 5, noise=0.004)[0]  # [2,50,h,w]
    seen = {}

    def model(x):
        seen["shape"], seen["max"], seen["flip_ok"] = tuple(x.shape), float(x.max()), bool(torch.equal(x[1], x[0].flip(1)))
        return [[torch.from_numpy(out).to(x.device)]]

    dropin.configure(limbs=skeleton.LIMBS)
    heatmap, paf = dropin.predict(image, params, model, model_params, 20, 30, "synthetic")
    assert seen["shape"] == (2,) + padded.shape and seen["max"] <= 1.0 and seen["flip_ok"]
    hm, pf = pp.post_network_scale(out, 4, padded.shape[:2], pad, image.shape[:2], 30, 48, skeleton.FLIP_PAF_ORD, skeleton.FLIP_HEAT_ORD[:18])
    ref_heat, ref_paf = pp.accumulate(np.zeros(hm.shape), hm, 1), pp.accumulate(np.zeros(pf.shape), pf, 1)
    assert heatmap.shape == (150, 210, 18) and paf.shape == (150, 210, 30) and paf.as_f64
    assert np.array_equal(heatmap.numpy(), ref_heat) and np.array_equal(paf.numpy(), ref_paf)
    # evaluate.py:509-511 on the device maps
    all_peaks = dropin.find_peaks(heatmap, params)
    connection_all, special_k = dropin.find_connections(all_peaks, paf, image.shape[0], params)
    subset, candidate = dropin.find_people(connection_all, special_k, all_peaks, params)
    o = so.group_batch(np.ascontiguousarray(ref_heat.transpose(2, 0, 1)[None]).astype(np.float32),
                       np.ascontiguousarray(ref_paf.transpose(2, 0, 1)[None]), skeleton.LIMBS, image.shape[0], params)
    d = diff_structures(o.as_reference_structures(0), (all_peaks, connection_all, special_k, subset, candidate), float_tol=0.0)
    assert not d, "\n".join(d)
    assert subset.shape[0] >= 3
    # and the fused call on the same device maps
    assert not diff_structures(o.as_reference_structures(0), dropin.group(heatmap, paf, image.shape[0], params), float_tol=0.0)
