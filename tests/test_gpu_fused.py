"""GPU parity of the fused trunk kernels (round 2) against a plain PyTorch fp32 reference of the same ops on bf16-rounded
operands, with the same rounding points (every stored activation is rounded to bf16 once).  All calls go through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from airpose_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bf16(x):
    return x.to(torch.bfloat16)


def _tail_reference(t1, w2, sc2, sh2, w3, sc3, sh3, res):
    """relu(bn3(conv3(bf16(relu(bn2(conv2(t1)))))) + res) in fp32 on the bf16 operands (NHWC in, NHWC out)."""
    x = t1.float().permute(0, 3, 1, 2)
    t2 = torch.relu(F.conv2d(x, w2.float(), padding=1) * sc2.view(1, -1, 1, 1) + sh2.view(1, -1, 1, 1))
    t2 = t2.to(torch.bfloat16).float()
    y = F.conv2d(t2, w3.float()) * sc3.view(1, -1, 1, 1) + sh3.view(1, -1, 1, 1) + res.float().permute(0, 3, 1, 2)
    return torch.relu(y).permute(0, 2, 3, 1).contiguous(), t2


@pytest.mark.parametrize("n,H,W", [(1, 56, 56), (3, 56, 56), (2, 54, 56), (2, 7, 40), (151, 8, 56)])
def test_bneck_tail_matches_torch(n, H, W):
    """conv2 3x3 + BN + ReLU -> conv3 1x1 + BN + residual + ReLU in one launch (csrc/bneck.cu): halo slab by one 4-d TMA box
    with out-of-bounds zero fill, nine row-shifted tcgen05 windows, in-place residual, clipped 4-d TMA store.  Cases: the
    layer1 geometry, a height that is not a multiple of the band (clipped last band), a narrow image (RM = 3 rows per
    M-tile) and more tiles than SMs (the persistent loop, barrier phases wrapping)."""
    lib = _lib.load()
    Cm, Co = 64, 256
    g = torch.Generator(device="cpu").manual_seed(1000 * n + H + W)
    t1 = _bf16(torch.relu(torch.randn(n, H, W, Cm, generator=g))).to(DEV)
    w2 = _bf16(torch.randn(Cm, Cm, 3, 3, generator=g) * (2.0 / (9 * Cm)) ** 0.5).to(DEV)
    w3 = _bf16(torch.randn(Co, Cm, 1, 1, generator=g) * (2.0 / Cm) ** 0.5).to(DEV)
    sc2, sh2 = (torch.rand(Cm, generator=g) + 0.5).to(DEV), (torch.randn(Cm, generator=g) * 0.3).to(DEV)
    sc3, sh3 = (torch.rand(Co, generator=g) + 0.5).to(DEV), (torch.randn(Co, generator=g) * 0.3).to(DEV)
    res = _bf16(torch.randn(n, H, W, Co, generator=g)).to(DEV)
    out = torch.full((n, H, W, Co), float("nan"), device=DEV, dtype=torch.bfloat16)
    w2k = w2.permute(0, 2, 3, 1).contiguous().view(Cm, 9 * Cm)
    w3k = w3.view(Co, Cm).contiguous()
    a = _lib.BneckTailArgs()
    a.t1, a.n, a.H, a.W, a.Cm = t1.data_ptr(), n, H, W, Cm
    a.w2, a.scale2, a.shift2 = w2k.data_ptr(), sc2.data_ptr(), sh2.data_ptr()
    a.w3, a.scale3, a.shift3 = w3k.data_ptr(), sc3.data_ptr(), sh3.data_ptr()
    a.residual, a.out = res.data_ptr(), out.data_ptr()
    _lib.check(lib.airpose_bneck_tail_bf16(C.byref(a), _lib.current_stream()), "bneck_tail")
    torch.cuda.synchronize()
    ref, _ = _tail_reference(t1, w2, sc2, sh2, w3, sc3, sh3, res)
    got = out.float()
    assert torch.isfinite(got).all(), "pixels were not written: %d" % int((~torch.isfinite(got)).sum())
    # output rounding (2^-9 relative) plus the effect of one-ulp flips of the bf16 intermediate (fp32 summation order
    # differs from torch's): a flipped t2 element moves an output by |w3| * ulp(t2) ~ 2^-8 * 0.2
    err = (got - ref).abs()
    bad = err > ref.abs() * 2.0 ** -8 + 2e-2
    print("bneck tail n=%d %dx%d: max abs err %.3e, mean abs err %.3e, bad %d / %d" %
          (n, H, W, err.max().item(), err.mean().item(), int(bad.sum()), bad.numel()))
    assert not bad.any()
    assert err.mean().item() < 2e-3


@pytest.mark.parametrize("n", [6, 58, 64])
def test_bneck_tail_is_deterministic(n):
    """No atomics, no data-dependent scheduling: the same launch twice gives the same bits (the trunk's chunk sizes)."""
    lib = _lib.load()
    Cm, Co, H = 64, 256, 56
    g = torch.Generator(device="cpu").manual_seed(n)
    t1 = _bf16(torch.relu(torch.randn(n, H, H, Cm, generator=g))).to(DEV)
    w2k = _bf16(torch.randn(Cm, 9 * Cm, generator=g) * 0.06).to(DEV)
    w3k = _bf16(torch.randn(Co, Cm, generator=g) * 0.17).to(DEV)
    sc2, sh2 = (torch.rand(Cm, generator=g) + 0.5).to(DEV), (torch.randn(Cm, generator=g) * 0.3).to(DEV)
    sc3, sh3 = (torch.rand(Co, generator=g) + 0.5).to(DEV), (torch.randn(Co, generator=g) * 0.3).to(DEV)
    res = _bf16(torch.randn(n, H, H, Co, generator=g)).to(DEV)
    outs = []
    for rep in range(4):
        out = torch.full((n, H, H, Co), float("nan"), device=DEV, dtype=torch.bfloat16)
        a = _lib.BneckTailArgs()
        a.t1, a.n, a.H, a.W, a.Cm = t1.data_ptr(), n, H, H, Cm
        a.w2, a.scale2, a.shift2 = w2k.data_ptr(), sc2.data_ptr(), sh2.data_ptr()
        a.w3, a.scale3, a.shift3 = w3k.data_ptr(), sc3.data_ptr(), sh3.data_ptr()
        a.residual, a.out = res.data_ptr(), out.data_ptr()
        _lib.check(lib.airpose_bneck_tail_bf16(C.byref(a), _lib.current_stream()), "bneck_tail")
        torch.cuda.synchronize()
        outs.append(out)
    for rep in range(1, 4):
        diff = (outs[rep].view(torch.int16) != outs[0].view(torch.int16))
        assert not diff.any(), "run %d differs from run 0 in %d elements (first at %s)" % (
            rep, int(diff.sum()), diff.nonzero()[0].tolist())


@pytest.mark.parametrize("n,H,W,relu", [(1, 28, 28, True), (3, 28, 28, True), (2, 27, 28, False), (2, 14, 14, True), (2, 5, 9, True), (75, 28, 28, True)])
def test_conv3x3_slab_matches_torch(n, H, W, relu):
    """3x3 / stride 1 conv + BN (+ ReLU) of the 128-channel stage with the input band resident in shared memory (csrc/conv3x3.cu),
    through airpose_conv_bf16 (which dispatches exactly like the trunk): the layer2 geometry, a height that is not a multiple of
    the band (clipped last band), small images (one band per image, short rows) and more bands than SMs (barrier phases wrap)."""
    lib = _lib.load()
    Cc = 128
    g = torch.Generator(device="cpu").manual_seed(77 * n + H + W)
    x = _bf16(torch.relu(torch.randn(n, H, W, Cc, generator=g))).to(DEV)
    w = _bf16(torch.randn(Cc, Cc, 3, 3, generator=g) * (2.0 / (9 * Cc)) ** 0.5).to(DEV)
    sc, sh = (torch.rand(Cc, generator=g) + 0.5).to(DEV), (torch.randn(Cc, generator=g) * 0.3).to(DEV)
    wk = w.permute(0, 2, 3, 1).contiguous().view(Cc, 9 * Cc)
    out = torch.full((n, H, W, Cc), float("nan"), device=DEV, dtype=torch.bfloat16)
    a = _lib.ConvArgs()
    a.x, a.n, a.H, a.W, a.Cin = x.data_ptr(), n, H, W, Cc
    a.w, a.Cout, a.ksize, a.stride, a.pad = wk.data_ptr(), Cc, 3, 1, 1
    a.scale, a.shift, a.relu, a.out = sc.data_ptr(), sh.data_ptr(), int(relu), out.data_ptr()
    n0 = _lib.launch_count()
    _lib.check(lib.airpose_conv_bf16(C.byref(a), _lib.current_stream()), "conv")
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 == 1
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    ref = (torch.relu(ref) if relu else ref).permute(0, 2, 3, 1)
    got = out.float()
    assert torch.isfinite(got).all(), "pixels were not written: %d" % int((~torch.isfinite(got)).sum())
    err = (got - ref).abs()
    bad = err > ref.abs() * 2.0 ** -8 + 2e-3
    print("conv3x3 slab n=%d %dx%d: max abs err %.3e, bad %d / %d" % (n, H, W, err.max().item(), int(bad.sum()), bad.numel()))
    assert not bad.any()
    out2 = torch.empty_like(out)
    a.out = out2.data_ptr()
    _lib.check(lib.airpose_conv_bf16(C.byref(a), _lib.current_stream()), "conv")
    torch.cuda.synchronize()
    assert torch.equal(out, out2)                       # deterministic
