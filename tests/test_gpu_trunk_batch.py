"""Trunk parity against something that is NOT the kernel, at the benchmark's batch.

Comparator: a fp32 `F.conv2d` chain in PyTorch on the same GPU with the trunk's rounding points -- bf16-rounded weights and
input, fp32 accumulation, BatchNorm as fp32 `acc * scale + shift` (scale = gamma / sqrt(var + eps)), residual add, ReLU, ONE
bf16 rounding per stored activation (forward_feat_ext / Bottleneck.forward, copenet/src/copenet/models/model_copenet.py:27-47,
161-176).  TF32 is off for the comparator.

  * test_every_conv_teacher_forced[n]: all 53 convs, each fed the COMPARATOR's own bf16 input of that layer, so an error cannot
    hide behind drift: <= 1 bf16 ulp per layer.  n = 2 and n = 128: at 128 images every layer runs the tile shape / stream-K /
    cta_group::2 variant the benchmark dispatches.  The fused kernels (stem+pool, layer1 conv2+conv3) are checked the same way.
  * test_trunk_at_benchmark_batch[B]: `forward_feat_ext_pair` at 64 and 256 pairs (128 / 512 images: chunks of 64, groups of
    128, two streams) against the chain for a sample of images spread over chunks, groups and both views.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from airpose_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LAYERS = (3, 4, 6, 3)


@pytest.fixture(scope="module")
def net(tmp_path_factory, net_state):
    from airpose_b200.model_copenet import getcopenet
    mp = synthetic.write_mean_params(str(tmp_path_factory.mktemp("mean") / "smpl_mean_params.npz"))
    m = getcopenet(mp, pretrained=False)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in net_state.items()}, strict=True)
    return m.to(DEV).eval()


@pytest.fixture(scope="module")
def params(net_state):
    """bf16-rounded conv weights (fp32 storage, OIHW) and folded fp32 BatchNorm per conv name."""
    sd = {k: torch.from_numpy(np.asarray(v)).to(DEV) for k, v in net_state.items()}

    def fold(bn):
        s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + 1e-5)
        return s.contiguous(), (sd[bn + ".bias"] - sd[bn + ".running_mean"] * s).contiguous()

    P = {"stem": (sd["conv1.weight"].to(torch.bfloat16).float(), *fold("bn1"))}
    for li, nb in enumerate(LAYERS, 1):
        for b in range(nb):
            p = "layer%d.%d" % (li, b)
            for c in (1, 2, 3):
                P["%s.conv%d" % (p, c)] = (sd["%s.conv%d.weight" % (p, c)].to(torch.bfloat16).float(), *fold("%s.bn%d" % (p, c)))
            if b == 0:
                P[p + ".down"] = (sd[p + ".downsample.0.weight"].to(torch.bfloat16).float(), *fold(p + ".downsample.1"))
    return P


def _rb(x):
    return x.to(torch.bfloat16).float()


def _ref_conv(x, prm, stride, pad, res=None, relu=True):
    """x NCHW fp32 holding bf16 values -> bf16-rounded fp32 NCHW."""
    w, sc, sh = prm
    y = F.conv2d(x, w, stride=stride, padding=pad) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    return _rb(torch.relu(y) if relu else y)


def _ref_stem(x, prm):
    y = _ref_conv(_rb(x), prm, 2, 3)
    return F.max_pool2d(y, 3, 2, 1)


def chain_features(x, P, hook=None):
    """The comparator's forward_feat_ext on NCHW fp32 images; `hook(name, kind, inp, res, out, stride)` sees every conv."""
    with torch.no_grad():
        y = _ref_stem(x, P["stem"])
        if hook:
            hook("stem", "stem", x, None, y, 2)
        for li, nb in enumerate(LAYERS, 1):
            for b in range(nb):
                p = "layer%d.%d" % (li, b)
                stride = 2 if (b == 0 and li > 1) else 1
                t1 = _ref_conv(y, P[p + ".conv1"], 1, 0)
                t2 = _ref_conv(t1, P[p + ".conv2"], stride, 1)
                res = _ref_conv(y, P[p + ".down"], stride, 0, relu=False) if b == 0 else y
                out = _ref_conv(t2, P[p + ".conv3"], 1, 0, res=res)
                if hook:
                    hook(p + ".conv1", "conv", y, None, t1, 1)
                    hook(p + ".conv2", "conv", t1, None, t2, stride)
                    if b == 0:
                        hook(p + ".down", "conv_norelu", y, None, res, stride)
                    hook(p + ".conv3", "conv", t2, res, out, 1)
                    hook(p + ".tail", "tail", t1, res, out, stride)
                y = out
        return y.mean(dim=(2, 3))


def _nhwc_bf16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _ours_conv(inp, prm, stride, res, relu):
    """One conv + BN (+ residual) (+ ReLU) through the C ABI, the way trunk.cu launches it: 1x1 stride 1 as a plain GEMM over a
    tiled tensor map (airpose_gemm_bf16), everything else through the TMA im2col map (airpose_conv_bf16)."""
    lib = _lib.load()
    w, sc, sh = prm
    Cout, Cin, k, _ = w.shape
    x = _nhwc_bf16(inp)
    n, H, W, _ = x.shape
    pad = k // 2
    Ho = (H + 2 * pad - k) // stride + 1
    wk = w.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().view(Cout, k * k * Cin)
    out = torch.full((n, Ho, Ho, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    r = _nhwc_bf16(res) if res is not None else None
    if k == 1 and stride == 1:
        g = _lib.GemmArgs()
        g.A, g.lda, g.B, g.ldb = x.data_ptr(), Cin, wk.data_ptr(), Cin
        g.M, g.N, g.K = n * H * W, Cout, Cin
        g.scale, g.shift, g.relu = sc.data_ptr(), sh.data_ptr(), int(relu)
        if r is not None:
            g.residual, g.ldr = r.data_ptr(), Cout
        g.out_bf16, g.ldd = out.data_ptr(), Cout
        _lib.check(lib.airpose_gemm_bf16(C.byref(g), _lib.current_stream()), "gemm")
    else:
        a = _lib.ConvArgs()
        a.x, a.n, a.H, a.W, a.Cin = x.data_ptr(), n, H, W, Cin
        a.w, a.Cout, a.ksize, a.stride, a.pad = wk.data_ptr(), Cout, k, stride, pad
        a.scale, a.shift, a.relu, a.out = sc.data_ptr(), sh.data_ptr(), int(relu), out.data_ptr()
        if r is not None:
            a.residual = r.data_ptr()
        _lib.check(lib.airpose_conv_bf16(C.byref(a), _lib.current_stream()), "conv")
    torch.cuda.synchronize()
    return out.float().permute(0, 3, 1, 2)


def _ours_tail(t1, prm2, prm3, res):
    lib = _lib.load()
    (w2, sc2, sh2), (w3, sc3, sh3) = prm2, prm3
    Cm, Co = w2.shape[0], w3.shape[0]
    x, r = _nhwc_bf16(t1), _nhwc_bf16(res)
    n, H, W, _ = x.shape
    out = torch.full((n, H, W, Co), float("nan"), device=DEV, dtype=torch.bfloat16)
    w2k = w2.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().view(Cm, 9 * Cm)
    w3k = w3.to(torch.bfloat16).view(Co, Cm).contiguous()
    a = _lib.BneckTailArgs()
    a.t1, a.n, a.H, a.W, a.Cm = x.data_ptr(), n, H, W, Cm
    a.w2, a.scale2, a.shift2 = w2k.data_ptr(), sc2.data_ptr(), sh2.data_ptr()
    a.w3, a.scale3, a.shift3 = w3k.data_ptr(), sc3.data_ptr(), sh3.data_ptr()
    a.residual, a.out = r.data_ptr(), out.data_ptr()
    _lib.check(lib.airpose_bneck_tail_bf16(C.byref(a), _lib.current_stream()), "bneck_tail")
    torch.cuda.synchronize()
    return out.float().permute(0, 3, 1, 2)


def _tail_ok(H, Cm):
    return Cm == 64 and 2 * (H + 2) <= 128          # bneck_tail_supported (csrc/bneck.cu)


@pytest.mark.parametrize("n", [2, 128])
def test_every_conv_teacher_forced(net, params, n):
    lib = _lib.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(1000 + n)
    x = torch.randn(n, 3, 224, 224, device=DEV, generator=g)
    net.forward_feat_ext(x[:min(n, 64)])            # builds + loads the handle (stem entry point needs it)
    rows, worst = [], [0.0, ""]

    def check(name, got, ref, one_ulp=2.0 ** -8, floor=2e-3, frac_tol=2e-4):       # measured worst fraction: 7e-5 (layer4.2)
        assert torch.isfinite(got).all(), name + ": unwritten output"
        err = (got - ref).abs()
        off = (err > ref.abs() * one_ulp + floor).float().mean().item()          # |err| > 1 bf16 ulp (half-ulp rounding both sides)
        two = (err > ref.abs() * 2.0 ** -6 + 8 * floor).float().mean().item()    # gross errors: none allowed
        rows.append((name, err.max().item(), off))
        if off > worst[0]:
            worst[0], worst[1] = off, name
        assert off < frac_tol, "%s: %.2e of the outputs are off by more than one bf16 ulp" % (name, off)
        assert two == 0.0, "%s: %.2e of the outputs are off by more than four bf16 ulps" % (name, two)

    def hook(name, kind, inp, res, out, stride):
        if kind == "stem":
            for i0 in range(0, n, 64):              # the stem entry point takes one chunk
                m = min(64, n - i0)
                o = torch.full((m, 56, 56, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
                _lib.check(lib.airpose_backbone_stem(net._handle, inp[i0:i0 + m].contiguous().data_ptr(), m, o.data_ptr(), _lib.current_stream()), "stem")
                torch.cuda.synchronize()
                check("stem(fused conv7x7+bn+relu+maxpool)[%d:%d]" % (i0, i0 + m), o.float().permute(0, 3, 1, 2), out[i0:i0 + m])
        elif kind == "tail":
            p = name[:-5]
            Cm = params[p + ".conv2"][0].shape[0]
            if stride == 1 and _tail_ok(inp.shape[2], Cm):
                # one-ulp flips of the unobservable bf16 intermediate move an output by |w3| * ulp(t2): a wider floor
                check(name + "(fused conv2+conv3)", _ours_tail(inp, params[p + ".conv2"], params[p + ".conv3"], res), out, floor=2e-2)
        else:
            check(name, _ours_conv(inp, params[name], stride, res, kind == "conv"), out)

    chain_features(x, params, hook)
    print("teacher-forced, n=%d images: %d kernels checked, worst fraction beyond one bf16 ulp %.2e (%s), worst max abs err %.3e" %
          (n, len(rows), worst[0], worst[1], max(r[1] for r in rows)))
    assert len(rows) >= 53 + 16 - 13      # 53 convs (+ fused tails where supported)


@pytest.mark.parametrize("B", [64, 256])
def test_trunk_at_benchmark_batch(net, params, B):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(B)
    x0 = torch.randn(B, 3, 224, 224, device=DEV, generator=g)
    x1 = torch.randn(B, 3, 224, 224, device=DEV, generator=g)
    xf = net.forward_feat_ext_pair(x0, x1)
    assert xf.shape == (2 * B, 2048)
    # a sample spread over chunks (64), groups (128) and both views: 16 images
    idx = sorted({0, 1, 31, 63, 64 % B, B - 1, B // 2, B // 2 + 1} | {B + i for i in (0, 1, 63 % B, B // 2, B - 2, B - 1)} | {(127) % (2 * B), (128) % (2 * B)})
    xs = torch.stack([(x0[i] if i < B else x1[i - B]) for i in idx])
    ref = chain_features(xs, params)
    got = xf[idx]
    err = (got - ref).abs()
    max_rel = (err.max() / ref.abs().max()).item()
    mean_rel = (err.mean() / ref.abs().mean()).item()
    per_img = (err.amax(dim=1) / ref.abs().amax(dim=1)).max().item()
    print("trunk at %d pairs (%d images) vs fp32 conv chain with bf16 rounding points, %d sampled images: max-rel %.3e, mean-rel %.3e, worst image %.3e" %
          (B, 2 * B, len(idx), max_rel, mean_rel, per_img))
    # two bf16 evaluations of a 53-layer network with different fp32 summation orders; measured values are printed above and
    # quoted in DESIGN.md section 4 -- the bound is 2x the measured maximum
    assert max_rel < TRUNK_MAX_REL and mean_rel < TRUNK_MEAN_REL


# measured on B200 (gpurun r02e2): 64 pairs max-rel 1.54e-3 / mean-rel 1.08e-3, 256 pairs 1.33e-3 / 1.09e-3
TRUNK_MAX_REL = 3.1e-3
TRUNK_MEAN_REL = 2.2e-3
