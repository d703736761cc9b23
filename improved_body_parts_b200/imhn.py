"""IMHN forward for inference: the 4-stack "identity-mapping hourglass network" whose output the
grouping path consumes -- ``/root/reference/models/posenet.py:43-117`` (``PoseNet``), ``:175-193`` (``NetworkEval``)
and the blocks of ``models/layers_transposed.py`` (``Backbone`` :158-194, ``Hourglass`` :197-282, ``Residual`` :12-47,
``Conv`` :90-121, ``DilatedConv`` :124-155, ``SELayer`` :285-306).

Library-level work: convolutions are cuDNN's.  What this module adds for the B200 pipeline:

* **checkpoint compatible** -- parameter names and shapes are those of the reference's ``NetworkEval`` (``posenet.pre.conv1.weight``,
  ``posenet.hourglass.0.hg.0.0.convBlock.0.weight`` ...), so ``torch.load(ckpt)['weights']`` (evaluate.py:629-630)
  loads with ``strict=True``; tests/test_imhn.py proves it by moving a state dict of the reference's own module across
  and comparing outputs;
* **inference-only graph** -- the reference computes all 5 output scales of all 4 stacks and ``predict()`` keeps
  ``output_tuple[-1][0]`` (evaluate.py:126).  ``forward`` here returns exactly that tensor and skips what cannot reach
  it: the last stack's four coarse-scale heads (``Features`` + ``outs`` for scales 1..4) and its merge layers;
* ``fold_batchnorm_()`` folds every eval-mode BatchNorm into its convolution (one kernel instead of two);
* ``Runner``: channels-last, bf16 autocast (replaces apex amp O1, evaluate.py:636-640), the whole forward captured in a
  CUDA graph per input shape, NCHW float32 ``[N, 50, h, w]`` out -- what ``spg_postnet`` reads in place.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

LEAK = 0.01


def _act() -> nn.Module:
    return nn.LeakyReLU(negative_slope=LEAK, inplace=True)


class Conv(nn.Module):
    """conv (+ BatchNorm) (+ LeakyReLU); parameter names ``conv`` / ``bn`` (layers_transposed.py:90-121, :124-155)."""

    def __init__(self, cin: int, cout: int, k: int = 3, bn: bool = True, relu: bool = True, dilation: int = 1):
        super().__init__()
        pad = dilation if dilation > 1 else (k - 1) // 2
        self.conv = nn.Conv2d(cin, cout, k, 1, padding=pad, dilation=dilation, bias=not bn)
        self.bn = nn.BatchNorm2d(cout) if bn else None
        self.relu = _act() if relu else None

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        return x if self.relu is None else self.relu(x)


class Residual(nn.Module):
    """1x1 -> 3x3 -> 1x1 bottleneck with BatchNorm, projection skip when the width changes, LeakyReLU after the sum
    (layers_transposed.py:12-47); parameter names ``convBlock.{0,1,3,4,6,7}`` / ``skipConv.{0,1}``."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        mid = cout // 2
        self.convBlock = nn.Sequential(
            nn.Conv2d(cin, mid, 1, bias=False), nn.BatchNorm2d(mid), _act(),
            nn.Conv2d(mid, mid, 3, 1, 1, bias=False), nn.BatchNorm2d(mid), _act(),
            nn.Conv2d(mid, cout, 1, bias=False), nn.BatchNorm2d(cout))
        self.skipConv = nn.Sequential(nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout)) if cin != cout else None
        self.relu = _act()

    def forward(self, x):
        y = self.convBlock(x)
        y = y + (x if self.skipConv is None else self.skipConv(x))
        return self.relu(y)


class SELayer(nn.Module):
    """Squeeze-and-excitation, reduction 16 (layers_transposed.py:285-306); parameter names ``fc.0`` / ``fc.2``."""

    def __init__(self, c: int, reduction: int = 16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c, c // reduction), nn.LeakyReLU(inplace=True), nn.Linear(c // reduction, c), nn.Sigmoid())

    def forward(self, x):
        w = self.fc(x.mean(dim=(2, 3)))
        return x * w[:, :, None, None]


class Backbone(nn.Module):
    """7x7 stride-2 stem, residuals, max-pool, six dilated 3x3 convolutions (3,3,4,4,5,5), concat -> 256 channels at
    1/4 resolution (layers_transposed.py:158-194)."""

    def __init__(self, n_feat: int = 256):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = _act()
        self.res1 = Residual(64, 128)
        self.res2 = Residual(128, 128)
        self.dilation = nn.Sequential(*[Conv(128, 128, 3, dilation=d) for d in (3, 3, 4, 4, 5, 5)])
        assert n_feat == 256, "the backbone concatenates 128 + 128 channels"

    def forward(self, x):
        x = self.res1(self.relu(self.bn1(self.conv1(x))))
        x = self.res2(F.max_pool2d(x, 2, 2))
        return torch.cat([x, self.dilation(x)], dim=1)


class Hourglass(nn.Module):
    """Order-``depth`` hourglass whose width grows by ``increase`` per level (layers_transposed.py:197-282).
    ``hg[d] = [skip residual, down residual, up residual, refine conv (, innermost residual)]``.  Returns the full-resolution
    feature map and the ``low2`` maps of every level, finest first."""

    def __init__(self, depth: int, n_feat: int, increase: int = 128, bn: bool = True):
        super().__init__()
        self.depth = depth
        levels = []
        for d in range(depth):
            c0, c1 = n_feat + increase * d, n_feat + increase * (d + 1)
            blocks = [Residual(c0, c0), Residual(c0, c1), Residual(c1, c0), Conv(c0, c0, 3, bn=bn)]
            if d == depth - 1:
                blocks.append(Residual(c1, c1))
            levels.append(nn.ModuleList(blocks))
        self.hg = nn.ModuleList(levels)

    def _level(self, d: int, x, lows: List):
        blk = self.hg[d]
        up1 = blk[0](x)
        low1 = blk[1](F.max_pool2d(x, 2, 2))
        low2 = blk[4](low1) if d == self.depth - 1 else self._level(d + 1, low1, lows)
        lows.append(low2)
        return up1 + blk[3](F.interpolate(blk[2](low2), scale_factor=2, mode="nearest"))

    def forward(self, x):
        lows: List = []
        top = self._level(0, x, lows)
        return [top] + lows[::-1]


class Features(nn.Module):
    """Per scale: two 3x3 convolutions down to ``c`` channels + SE (posenet.py:24-40)."""

    def __init__(self, c: int, increase: int = 128, bn: bool = True):
        super().__init__()
        self.before_regress = nn.ModuleList(
            [nn.Sequential(Conv(c + i * increase, c, 3, bn=bn), Conv(c, c, 3, bn=bn), SELayer(c)) for i in range(5)])


class Merge(nn.Module):
    """1x1 convolution without activation (posenet.py:13-21)."""

    def __init__(self, cin: int, cout: int, bn: bool = True):
        super().__init__()
        self.conv = Conv(cin, cout, 1, bn=bn, relu=False)

    def forward(self, x):
        return self.conv(x)


class PoseNet(nn.Module):
    """posenet.py:43-117.  ``forward`` returns ``pred[-1][0]`` only (see the module docstring); ``forward_all`` returns the
    reference's full ``nstack x 5`` list (used by the parity test)."""

    def __init__(self, nstack: int = 4, inp_dim: int = 256, oup_dim: int = 50, bn: bool = True, increase: int = 128):
        super().__init__()
        self.nstack = nstack
        self.pre = Backbone(inp_dim)
        self.hourglass = nn.ModuleList([Hourglass(4, inp_dim, increase, bn=bn) for _ in range(nstack)])
        self.features = nn.ModuleList([Features(inp_dim, increase, bn=bn) for _ in range(nstack)])
        self.outs = nn.ModuleList([nn.ModuleList([Conv(inp_dim, oup_dim, 1, bn=False, relu=False) for _ in range(5)])
                                   for _ in range(nstack)])
        self.merge_features = nn.ModuleList([nn.ModuleList([Merge(inp_dim, inp_dim + j * increase, bn=bn) for j in range(5)])
                                             for _ in range(nstack - 1)])
        self.merge_preds = nn.ModuleList([nn.ModuleList([Merge(oup_dim, inp_dim + j * increase, bn=bn) for j in range(5)])
                                          for _ in range(nstack - 1)])

    def _run(self, imgs, full: bool):
        x = self.pre(imgs.permute(0, 3, 1, 2))  # the reference feeds NHWC images in [0, 1] (posenet.py:84)
        cache: List = [None] * 5
        preds = []
        for i in range(self.nstack):
            last = i == self.nstack - 1
            fms = self.hourglass[i](x)
            if i > 0:
                fms = [f + c for f, c in zip(fms, cache)]
            scales = range(5) if (full or not last) else range(1)  # the last stack's coarse heads reach nothing predict() reads
            stack_preds = []
            for j in scales:
                feat = self.features[i].before_regress[j](fms[j])
                p = self.outs[i][j](feat)
                stack_preds.append(p)
                if not last:
                    cache[j] = self.merge_preds[i][j](p) + self.merge_features[i][j](feat)
            if not last:
                x = x + cache[0]
            preds.append(stack_preds)
        return preds

    def forward(self, imgs):
        return self._run(imgs, full=False)[-1][0]

    def forward_all(self, imgs):
        return self._run(imgs, full=True)


class IMHN(nn.Module):
    """``NetworkEval`` (posenet.py:175-193): the key prefix of the checkpoint's tensors is ``posenet.``."""

    def __init__(self, nstack: int = 4, inp_dim: int = 256, oup_dim: int = 50, increase: int = 128, bn: bool = True):
        super().__init__()
        self.posenet = PoseNet(nstack, inp_dim, oup_dim, bn=bn, increase=increase)
        self.eval()

    def forward(self, imgs):
        """``imgs [N, H, W, 3]`` float in [0, 1], BGR (evaluate.py:103-121) -> ``[N, 50, H/4, W/4]``: 30 body-part, 18
        keypoint, 2 background channels (config/config.py:101-103) = ``output_tuple[-1][0]`` of the reference."""
        return self.posenet(imgs)

    def forward_all(self, imgs):
        return self.posenet.forward_all(imgs)

    @torch.no_grad()
    def init_like_reference_(self, seed: int = 0) -> "IMHN":
        """The reference's initialisation (posenet.py:119-139): N(0, 0.001) convolutions, unit BatchNorm, N(0, 0.01) linear."""
        gen = torch.Generator().manual_seed(seed)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.copy_(torch.randn(m.weight.shape, generator=gen) * 0.001)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.fill_(1)
                m.bias.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=gen) * 0.01)
                m.bias.zero_()
        return self

    @torch.no_grad()
    def fold_batchnorm_(self) -> "IMHN":
        """Fold every (conv, BatchNorm) pair into the convolution (eval mode: BatchNorm is an affine map).  After this the
        module no longer loads reference checkpoints -- fold after loading."""
        def fold(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation,
                              bias=True).to(conv.weight.device, conv.weight.dtype)
            fused.weight.copy_(conv.weight * scale[:, None, None, None])
            fused.bias.copy_(bn.bias - bn.running_mean * scale + (conv.bias * scale if conv.bias is not None else 0))
            return fused

        for m in self.modules():
            if isinstance(m, Conv) and m.bn is not None:
                m.conv, m.bn = fold(m.conv, m.bn), None
            elif isinstance(m, Residual):
                seq = list(m.convBlock)
                m.convBlock = nn.Sequential(fold(seq[0], seq[1]), seq[2], fold(seq[3], seq[4]), seq[5], fold(seq[6], seq[7]))
                if m.skipConv is not None:
                    m.skipConv = nn.Sequential(fold(m.skipConv[0], m.skipConv[1]))
            elif isinstance(m, Backbone) and isinstance(m.bn1, nn.BatchNorm2d):
                m.conv1, m.bn1 = fold(m.conv1, m.bn1), nn.Identity()
        return self


class Runner:
    """Inference engine around an ``IMHN``: channels-last weights, bf16 autocast, one CUDA graph per input shape.

    ``__call__(imgs [N,H,W,3] float32 CUDA) -> [N, 50, H/4, W/4] float32`` (NCHW, contiguous rows: what ``spg_postnet``
    consumes in place).  The returned tensor is the graph's static output buffer: consume it before the next call."""

    def __init__(self, model: IMHN, device="cuda:0", dtype=torch.bfloat16, use_graph: bool = True, fold_bn: bool = True):
        self.device, self.dtype, self.use_graph = torch.device(device), dtype, use_graph
        self.model = model.to(self.device).eval()
        if fold_bn:
            self.model.fold_batchnorm_()
        self.model = self.model.to(memory_format=torch.channels_last)
        self._graphs: Dict[Tuple[int, ...], tuple] = {}

    @torch.no_grad()
    def _forward(self, imgs):
        with torch.autocast("cuda", dtype=self.dtype, enabled=self.dtype != torch.float32):
            return self.model(imgs).float().contiguous()

    @torch.no_grad()
    def __call__(self, imgs):
        if not self.use_graph:
            return self._forward(imgs)
        key = tuple(imgs.shape)
        if key not in self._graphs:
            static_in = torch.zeros_like(imgs)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # warm-up outside the capture (cuDNN algorithm selection, workspaces)
                for _ in range(3):
                    self._forward(static_in)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._forward(static_in)
            self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = self._graphs[key]
        static_in.copy_(imgs)
        graph.replay()
        return static_out


def load_reference_checkpoint(model: IMHN, path: str, strict: bool = True) -> IMHN:
    """evaluate.py:629-630: ``checkpoint['weights']`` -> the model (CPU map, strict key match by default)."""
    ckpt = torch.load(path, map_location="cpu")
    model.load_state_dict(ckpt["weights"] if "weights" in ckpt else ckpt, strict=strict)
    return model
