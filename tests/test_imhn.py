"""CPU: the inference IMHN (improved_body_parts_b200/imhn.py) against the reference's own module.

Where /root/reference is present (build container) the reference's ``NetworkEval`` (models/posenet.py:175-193) is
instantiated, its state dict is loaded into ours with ``strict=True`` (same parameter names and shapes: reference
checkpoints load, evaluate.py:629-630) and the outputs are compared: all ``nstack x 5`` maps of ``forward_all`` and the
single tensor ``forward`` returns (= ``output_tuple[-1][0]``, evaluate.py:126).  Elsewhere only the self-consistency
checks run (BatchNorm folding, pruned vs full forward)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SPG_REFERENCE_ROOT", "/root/reference")


def _small(nstack=2):
    from improved_body_parts_b200.imhn import IMHN
    torch.manual_seed(0)
    m = IMHN(nstack=nstack)
    with torch.no_grad():  # non-trivial BatchNorm statistics, weights large enough for visible outputs
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.uniform_(-0.2, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.1, 0.1)
            elif isinstance(mod, torch.nn.Conv2d):
                mod.weight.normal_(0, (2.0 / (mod.weight[0].numel())) ** 0.5)
    return m.eval()


def test_pruned_forward_is_the_last_stacks_finest_map_and_bn_folding_is_exact_enough():
    m = _small(2)
    x = torch.rand(2, 64, 64, 3)
    with torch.no_grad():
        full = m.forward_all(x)
        one = m(x)
        assert len(full) == 2 and len(full[0]) == 5 and [tuple(p.shape[1:]) for p in full[-1]] == \
            [(50, 16, 16), (50, 8, 8), (50, 4, 4), (50, 2, 2), (50, 1, 1)]
        assert torch.equal(one, full[-1][0])
        n_bn = sum(isinstance(mod, torch.nn.BatchNorm2d) for mod in m.modules())
        m.fold_batchnorm_()
        assert n_bn > 100 and not any(isinstance(mod, torch.nn.BatchNorm2d) for mod in m.modules())
        folded = m(x)
    assert one.abs().max() > 1e-3
    assert torch.allclose(folded, one, rtol=1e-4, atol=1e-5 * float(one.abs().max()))


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "models", "posenet.py")), reason="needs /root/reference")
def test_reference_state_dict_loads_strictly_and_outputs_agree():
    import contextlib
    import io
    import types

    from improved_body_parts_b200.imhn import IMHN
    sys.path.insert(0, REF)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from models.posenet import NetworkEval  # the reference's module, unmodified
    finally:
        sys.path.remove(REF)
    opt = types.SimpleNamespace(nstack=4, hourglass_inp_dim=256, increase=128)
    cfg = types.SimpleNamespace(num_layers=50)
    torch.manual_seed(1)
    ref = NetworkEval(opt, cfg, bn=True).eval()
    with torch.no_grad():
        for mod in ref.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.uniform_(-0.2, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
    ours = IMHN(nstack=4)
    missing = ours.load_state_dict(ref.state_dict(), strict=True)  # same names, same shapes
    assert not missing.missing_keys and not missing.unexpected_keys
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters()) > 100e6
    x = torch.rand(2, 64, 64, 3)
    with torch.no_grad():
        want = ref(x)
        got = ours.forward_all(x)
        worst = 0.0
        for i in range(4):
            for j in range(5):
                scale = float(want[i][j].abs().max())
                worst = max(worst, float((got[i][j] - want[i][j]).abs().max()) / scale)
        # same float32 network; only the association of a few sums differs (x + (a + b) vs (x + a) + b, mean vs avg-pool)
        assert worst < 2e-5, worst
        assert float((ours(x) - want[-1][0]).abs().max()) / float(want[-1][0].abs().max()) < 2e-5
