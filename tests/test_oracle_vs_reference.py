"""CPU, build container only: the checker against the reference's own functions executed live (skipped where
/root/reference is absent, e.g. on the GPU box -- there the committed goldens stand in for it).

Complements tests/golden/: random hyper-parameters and dirty-input knobs the fixed golden cases do not enumerate."""
import numpy as np
import pytest

from oracle import spg_oracle as so
from oracle.ref_loader import Reference, reference_available
from parity import diff_structures

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference checkout not present")


def fuzz_cases(n_trials, seed=20260921):
    """The draw sequence shared with tests/test_gpu_parity.py::test_random_parameters_and_dirt_against_the_checker."""
    from improved_body_parts_b200 import skeleton, synth

    rng = np.random.default_rng(seed)
    for trial in range(n_trials):
        H = int(rng.choice([48, 64, 72, 96, 128])); W = int(rng.choice([48, 64, 80, 96, 128]))
        persons = int(rng.integers(1, 14))
        kw = dict(drop_prob=float(rng.choice([0.0, 0.1, 0.3])), plateau=int(rng.integers(0, 4)), spikes=int(rng.integers(0, 20)),
                  colocate=int(rng.integers(0, 3)), edge=bool(rng.integers(0, 2)), stretch=int(rng.integers(0, 4)),
                  negative_bias=float(rng.choice([0.0, 0.0, 0.03])), heat_gain=float(rng.choice([1.0, 1.0, 0.5])),
                  paf_gain=float(rng.choice([1.0, 1.0, 0.6])))
        params = dict(skeleton.default_params(),
                      thre1=float(rng.choice([0.05, 0.1, 0.2, 0.3])), thre2=float(rng.choice([0.02, 0.05, 0.1, 0.3])),
                      connect_ration=float(rng.choice([0.5, 0.7, 0.8, 0.95, 1.0])), mid_num=int(rng.choice([1, 2, 5, 10, 20, 33, 40])),
                      len_rate=float(rng.choice([1.2, 4.0, 16.0])), connection_tole=float(rng.choice([0.3, 0.7, 1.3])),
                      offset_radius=int(rng.integers(0, 5)), remove_recon=int(rng.integers(0, 2)))
        f64 = bool(rng.integers(0, 4) == 0)
        heat, paf = synth.make_batch(31000 + trial, 3, H, W, persons, **kw)
        if f64:
            paf = paf.astype(np.float64) * (1.0 + 2.0 ** -27)
        extent = int(rng.choice([H, 40, 2 * H]))
        cap = int(rng.choice([64, 128]))
        yield trial, heat, paf, extent, params, cap


def test_checker_equals_live_reference_on_random_parameters():
    from improved_body_parts_b200 import skeleton

    ref = Reference()
    for trial, heat, paf, extent, params, _ in fuzz_cases(24):
        o = so.group_batch(heat, paf, skeleton.LIMBS, extent, params)
        assert (o.status == 0).all()
        for i in range(2):
            want = ref.group(np.ascontiguousarray(heat[i].transpose(1, 2, 0)), np.ascontiguousarray(paf[i].transpose(1, 2, 0)),
                             extent, params)
            d = diff_structures(want, o.as_reference_structures(i), float_tol=0.0)
            assert not d, f"trial {trial} image {i} {params}:\n" + "\n".join(d)
