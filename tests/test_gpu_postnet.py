"""GPU: the post-network stage (spg_postnet, csrc/postnet.cuh -- evaluate.py:126-161 on the device) against its CPU
checker oracle/postnet_port.py, which tests/test_postnet_port.py pins to cv2.

The kernel spells out the port's float32 operations one by one, so the bar here is BIT-IDENTICAL maps (which implies the
north_star's 1e-4 and identical integer peaks downstream); the downstream check is run anyway, through the
float32-storage / float64-arithmetic mode (SPG_F32_AS_F64) that the single-scale maps allow."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(cuda_device):
    import torch
    from improved_body_parts_b200 import skeleton, synth
    from improved_body_parts_b200.grouping import Grouper
    from oracle import postnet_port as pp
    from oracle import spg_oracle as so

    class Env:
        pass

    e = Env()
    e.torch, e.skeleton, e.synth, e.Grouper, e.pp, e.so, e.dev = torch, skeleton, synth, Grouper, pp, so, cuda_device
    return e


def _network_like_output(env, seed, n, h, w, persons, noise=0.01):
    """[n, 2, 50, h, w] float32: synthetic maps at network resolution in the network's channel layout (body parts 0..29,
    keypoints 30..47, background 48..49), the second of each pair a noisy mirrored + channel-permuted copy, as the
    network would answer for the mirrored image."""
    rng = np.random.default_rng(seed)
    heat, paf = env.synth.make_batch(seed, n, h, w, persons)
    out = np.zeros((n, 2, 50, h, w), np.float32)
    out[:, 0, :30], out[:, 0, 30:48] = paf, heat
    out[:, 0, 48:] = rng.random((n, 2, h, w), dtype=np.float32)
    inv_p = np.argsort(env.skeleton.FLIP_PAF_ORD)
    inv_h = np.argsort(env.skeleton.FLIP_HEAT_ORD[:18])
    out[:, 1, :30] = paf[:, inv_p][..., ::-1]
    out[:, 1, 30:48] = heat[:, inv_h][..., ::-1]
    out[:, 1] += (rng.random((n, 50, h, w), dtype=np.float32) - 0.5) * np.float32(noise)
    return out


def _port_maps(env, outs, crops, stride, image_hw):
    """The checker: evaluate.py:126-161 per scale through oracle/postnet_port.py -> (heat [N,K,H,W] f64, paf [N,L,H,W] f64)."""
    pp, sk = env.pp, env.skeleton
    N = outs[0].shape[0]
    H, W = image_hw
    heat_avg = np.zeros((N, H, W, 18))
    paf_avg = np.zeros((N, H, W, 30))
    for o, (ch, cw) in zip(outs, crops):
        h, w = o.shape[3:]
        padded = (h * stride, w * stride)
        pad = [0, 0, padded[0] - ch, padded[1] - cw]
        for i in range(N):
            hm, pf = pp.post_network_scale(o[i].astype(np.float32), stride, padded, pad, (H, W), 30, 48, sk.FLIP_PAF_ORD,
                                           sk.FLIP_HEAT_ORD[:18])
            heat_avg[i] = pp.accumulate(heat_avg[i], hm, len(outs))
            paf_avg[i] = pp.accumulate(paf_avg[i], pf, len(outs))
    return heat_avg.transpose(0, 3, 1, 2), paf_avg.transpose(0, 3, 1, 2)


CASES = {
    # name: (network sizes per scale, crops per scale, image size)
    "identity_128_to_512": ([(32, 40)], [(128, 160)], (128, 160)),
    "identity_odd": ([(19, 23)], [(75, 90)], (75, 90)),                    # the identity kernel's border tiles: sizes not multiples of 4
    "identity_tall": ([(40, 9)], [(160, 33)], (160, 33)),
    "identity_wide": ([(12, 80)], [(45, 320)], (45, 320)),                 # full-width tiles with 16-byte stores, a partial tile row
    "identity_wide_odd": ([(11, 70)], [(41, 277)], (41, 277)),             # a full tile followed by a partial one, W not a multiple of 4
    "padded_ratio_1.25": ([(32, 48)], [(120, 180)], (96, 144)),            # 640-style box: crop 120x180 of 128x192, image smaller
    "upscale_ratio_0.6": ([(16, 32)], [(60, 110)], (100, 183)),            # image larger than the network input
    "odd_sizes": ([(19, 23)], [(70, 89)], (131, 167)),
    "three_scales": ([(16, 16), (32, 32), (64, 64)], [(64, 64), (128, 128), (250, 250)], (125, 125)),
    "many_tiles_ratio_1.33": ([(48, 64)], [(192, 250)], (144, 188)),       # several tiles per axis, inner and border ones
    "ratio_2_and_0.5": ([(64, 32)], [(256, 128)], (128, 256)),             # crop twice / half the image
    "tiny": ([(3, 5)], [(9, 17)], (7, 23)),
    # more scales than one launch fuses (4): the float64 sums continue through memory in the second launch
    "five_scales": ([(8, 8), (12, 12), (16, 16), (24, 24), (32, 32)], [(30, 30), (48, 48), (64, 64), (90, 90), (128, 128)], (64, 64)),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("net_dtype", ["f32", "f16"])
def test_postnet_maps_are_the_checkers_maps(env, name, net_dtype):
    t = env.torch
    sizes, crops, image_hw = CASES[name]
    if net_dtype == "f16" and name not in ("identity_128_to_512", "padded_ratio_1.25"):
        pytest.skip("f16 input is covered on two geometries")
    N = 2
    outs = [_network_like_output(env, 500 + 7 * k, N, h, w, 4) for k, (h, w) in enumerate(sizes)]
    if net_dtype == "f16":
        outs = [o.astype(np.float16) for o in outs]
    ref_heat, ref_paf = _port_maps(env, outs, crops, 4, image_hw)
    g = env.Grouper(max_batch=N, max_h=image_hw[0], max_w=image_hw[1])
    try:
        dev_outs = [t.from_numpy(o).to(env.dev) for o in outs]
        heat, paf = g.postnet(dev_outs, crops, image_hw, paf_dtype=t.float64)
        assert heat.dtype == t.float32 and paf.dtype == t.float64
        assert np.array_equal(heat.cpu().numpy(), ref_heat.astype(np.float32)), "keypoint maps differ from the checker"
        assert np.array_equal(paf.cpu().numpy(), ref_paf), "body-part maps differ from the checker"
        if len(sizes) == 1:  # single scale: float32 planes hold the same values
            heat2, paf2 = g.postnet(dev_outs, crops, image_hw)
            assert paf2.dtype == t.float32 and np.array_equal(paf2.cpu().numpy().astype(np.float64), ref_paf)
            assert np.array_equal(heat2.cpu().numpy(), heat.cpu().numpy())
    finally:
        g.close()
    assert np.abs(ref_heat).max() > 0.3 and np.abs(ref_paf).max() > 0.3


@pytest.mark.parametrize("stride,hw,crop,image", [(2, (40, 48), (76, 90), (61, 77)), (8, (12, 16), (90, 120), (90, 120)), (4, (24, 24), (96, 96), (96, 96))])
def test_other_strides_take_the_generic_kernel(env, stride, hw, crop, image):
    """stride 4 (the reference's model) runs the four-phase kernel; any other stride the table-driven generic one.  Both
    against the checker, bit for bit."""
    t = env.torch
    out = _network_like_output(env, 321 + stride, 2, hw[0], hw[1], 3)
    ref_heat, ref_paf = _port_maps(env, [out], [crop], stride, image)
    g = env.Grouper(max_batch=2, max_h=image[0], max_w=image[1])
    try:
        heat, paf = g.postnet([t.from_numpy(out).to(env.dev)], [crop], image, stride=stride, paf_dtype=t.float64)
        assert np.array_equal(heat.cpu().numpy(), ref_heat.astype(np.float32)) and np.array_equal(paf.cpu().numpy(), ref_paf)
    finally:
        g.close()


def test_network_tensor_is_consumed_in_place_with_strides(env):
    """Channel / pair / image strides are arbitrary: a [N,2,50,h,w] view into a larger buffer works without a copy."""
    t = env.torch
    out = _network_like_output(env, 77, 2, 24, 28, 3)
    big = t.zeros((2, 2, 64, 24, 28), device=env.dev)
    big[:, :, 7:57] = t.from_numpy(out).to(env.dev)
    ref_heat, ref_paf = _port_maps(env, [out], [(90, 100)], 4, (90, 100))
    g = env.Grouper(max_batch=2, max_h=90, max_w=100)
    try:
        heat, paf = g.postnet([big[:, :, 7:57]], [(90, 100)], (90, 100), paf_dtype=t.float64)
        assert np.array_equal(heat.cpu().numpy(), ref_heat.astype(np.float32)) and np.array_equal(paf.cpu().numpy(), ref_paf)
    finally:
        g.close()


def test_nan_scrub_is_the_demo_behaviour(env):
    t = env.torch
    out = _network_like_output(env, 78, 1, 16, 16, 2)
    out[0, 0, 3, 5, 5] = np.nan
    out[0, 0, 35, 2, 9] = np.nan
    g = env.Grouper(max_batch=1, max_h=64, max_w=64)
    try:
        d = t.from_numpy(out).to(env.dev)
        h0, p0 = g.postnet([d], [(64, 64)], (64, 64))
        h1, p1 = g.postnet([d], [(64, 64)], (64, 64), nan_scrub=True)
        assert t.isnan(h0).any() and t.isnan(p0).any() and not t.isnan(h1).any() and not t.isnan(p1).any()
        keep = ~t.isnan(p0)
        assert t.equal(p0[keep], p1[keep]) and (p1[~keep] == 0).all()
    finally:
        g.close()


@pytest.mark.parametrize("hw,persons,n", [((32, 32), 6, 6), ((40, 56), 8, 3)])
def test_postnet_then_grouping_equals_the_checkers_pipeline(env, hw, persons, n):
    """Network output -> spg_postnet -> grouping with float32 body-part planes in float64 arithmetic (SPG_F32_AS_F64),
    all on the device, against the CPU pipeline: postnet_port (float64 maps, as predict() returns them) -> C checker."""
    from test_gpu_parity import _assert_same
    t = env.torch
    h, w = hw
    out = _network_like_output(env, 900 + h, n, h, w, persons, noise=0.004)
    image_hw = (4 * h, 4 * w)
    ref_heat, ref_paf = _port_maps(env, [out], [image_hw], 4, image_hw)
    params = env.skeleton.default_params()
    o = env.so.group_batch(ref_heat.astype(np.float32), np.ascontiguousarray(ref_paf), env.skeleton.LIMBS, image_hw[0], params)
    g = env.Grouper(max_batch=n, max_h=image_hw[0], max_w=image_hw[1], max_peaks_per_part=128, max_person_rows=128)
    try:
        heat, paf = g.postnet([t.from_numpy(out).to(env.dev)], [image_hw], image_hw)
        assert paf.dtype == t.float32
        g.group_device(heat, paf, image_hw[0], params, paf_as_f64=True)
        r = g.fetch()
        kernels = g.stage_kernels()
    finally:
        g.close()
    assert (r.status == 0).all() and (o.status == 0).all() and r.n_persons.sum() >= n
    for i in range(n):
        _assert_same(o.as_reference_structures(i), r.as_reference_structures(i), f"image {i} ({kernels[1]})")


@pytest.mark.parametrize("H,persons,kw", [(128, 24, {}), (128, 30, dict(drop_prob=0.1, spikes=8, colocate=2, edge=True)),
                                          (176, 12, {}), (512, 6, dict(scale_range=(2.0, 3.0), sigma_scale=2.0))])
def test_f32_storage_f64_arithmetic_equals_f64_planes(env, H, persons, kw):
    """SPG_F32_AS_F64: float32-stored planes evaluated in float64 give what float64 planes with the same values give
    (the checker's float64 path), on the persistent (128), the staged per-item (176) and the L2 (512) schedules --
    and NOT what the float32 path gives (the sums round differently), which the test also shows."""
    from test_gpu_parity import _assert_same
    t = env.torch
    n = 6 if H < 512 else 2
    heat, paf = env.synth.make_batch(4711 + H, n, H, H, persons, **kw)
    params = env.skeleton.default_params()
    o64 = env.so.group_batch(heat, paf.astype(np.float64), env.skeleton.LIMBS, H, params)
    g = env.Grouper(max_batch=n, max_h=H, max_w=H)
    try:
        hd, pd = t.from_numpy(heat).to(env.dev), t.from_numpy(paf).to(env.dev)
        g.group_device(hd, pd, H, params, paf_as_f64=True)
        r = g.fetch()
        name = g.stage_kernels()[1]
        g.group_device(hd, pd, H, params)
        r32 = g.fetch()
    finally:
        g.close()
    assert "double" in name and (r.status == 0).all()
    for i in range(n):
        _assert_same(o64.as_reference_structures(i), r.as_reference_structures(i), f"image {i} ({name})")
    assert not np.array_equal(r.conn_score, r32.conn_score)  # float32 arithmetic is a different (also reference-exact) path
