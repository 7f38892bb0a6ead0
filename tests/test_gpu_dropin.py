"""The UNMODIFIED reference LightningModule (copenet/src/copenet/copenet_twoview.py, from oracle/_ref or /root/reference)
driven over airpose_b200's objects: the import swap of INTEGRATION.md section 1 executed, not asserted.

`copenet.models.model_copenet` and `copenet.smplx.smplx` are replaced in sys.modules before the reference module is imported
(= editing its import lines :18 and :22); everything else -- `fwd_pass_and_loss` (:164-374), `get_loss` (:83-161),
`training_step` (:376-390), `configure_optimizers` (:416-425), `transform_smpl`, `perspective_projection`,
`rot6d_to_rotmat` -- is the reference's own code running on CUDA tensors."""
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def ref_module():
    if not rh.available():
        pytest.skip("reference sources absent: run `python oracle/make_ref.py` in the build container (oracle/_ref travels with the snapshot)")
    import airpose_b200.model_copenet as our_model
    import airpose_b200.smplx as our_smplx
    shim = types.ModuleType("copenet.smplx.smplx")
    shim.SMPLX, shim.lbs = our_smplx.SMPLX, None          # `lbs` is imported by the reference module but never used (:22)
    rt = rh.import_reference("cuda", inject={"copenet.models.model_copenet": our_model, "copenet.smplx.smplx": shim})
    module = rh.make_module(rt, 2, device="cuda")
    assert type(module.model).__module__ == "airpose_b200.model_copenet"
    assert type(rt.smplx_test).__module__ == "airpose_b200.smplx"
    yield rt, module
    for k in [k for k in sys.modules if k == "copenet" or k.startswith("copenet.")]:
        del sys.modules[k]


def test_reference_fwd_pass_and_loss_over_airpose_objects(ref_module):
    rt, module = ref_module
    g = np.load(os.path.join(GOLDEN, "twoview_b2.npz"))
    batch = rh.make_batch(2, int(g["in_seed"]), 321, device="cuda")
    for k in ("smpl_vertices", "smpl_joints", "smpl_joints_2d0", "smplorient_rel0"):      # same ground truth as the golden run
        assert np.allclose(batch[k].cpu().numpy(), g["gt/" + k], atol=2e-4), k
    module.eval()
    with torch.no_grad():
        output, losses, loss = module.fwd_pass_and_loss(batch, is_val=True, is_test=False)
    worst = {}
    for v in (0, 1):
        got = output["pred_vertices_cam%d" % v].cpu().numpy()
        ref = g["fp32/pred_vertices_cam%d" % v]
        worst["verts%d" % v] = float(np.abs(got - ref).max() / np.abs(ref).max())
        assert got.shape == ref.shape == (2, 10475, 3)
    rel_loss = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    print("reference LightningModule over airpose_b200: vertices rel err %s, loss %.6g vs reference %.6g (rel %.2e)" %
          (worst, float(loss), float(g["loss"]), rel_loss))
    # the only deviation is the bf16 trunk (the golden is the reference's fp32 run): 2.7e-3 on the features.
    # measured on B200: vertices 8.2e-4 / 1.1e-3, loss 3.6e-3; bounds = 2x
    assert max(worst.values()) < 2.5e-3
    assert rel_loss < 8e-3
    for k in ("loss_regr_pose", "loss_regr_shape", "loss_regul_betas"):
        assert abs(losses[k] - float(g["loss/" + k])) <= 5e-2 * abs(float(g["loss/" + k])) + 1e-4, k


def test_reference_training_step_over_airpose_objects(ref_module):
    """training_step (:376-390) + loss.backward() + the reference's own Adam(amsgrad) step (:416-425)."""
    rt, module = ref_module
    batch = rh.make_batch(2, 123, 321, device="cuda")
    module.train()
    opt = module.configure_optimizers()
    assert isinstance(opt, torch.optim.Adam)
    before = {n: p.detach().clone() for n, p in module.model.named_parameters()}
    res = module.training_step(batch, 1)
    assert torch.isfinite(res["loss"])
    opt.zero_grad()
    res["loss"].backward()
    missing = [n for n, p in module.model.named_parameters() if p.grad is None]
    assert missing == ["deccam.weight", "deccam.bias"], missing          # unused by the two-view model (model_copenet.py:73)
    for n, p in module.model.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), n
    assert sum(float(p.grad.abs().sum()) for p in module.model.parameters() if p.grad is not None) > 0
    opt.step()
    moved = [n for n, p in module.model.named_parameters() if p.grad is not None and not torch.equal(p.detach(), before[n])]
    assert len(moved) >= 0.95 * (len(before) - 2), "only %d of %d parameters moved" % (len(moved), len(before))
    module.eval()


def _install_bf16_rounding_points(net):
    """The CUDA trunk's forward rounding points on the reference trunk, differentiable (straight-through): conv inputs and conv
    outputs rounded to bf16, block outputs and the pooled stem rounded to bf16; conv weights are rounded by the caller (both
    sides load the same bf16-representable weights)."""
    rb = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()
    for name, mod in net.named_modules():
        if isinstance(mod, torch.nn.Conv2d):
            mod.register_forward_pre_hook(lambda m, a: (rb(a[0]),))
            mod.register_forward_hook(lambda m, a, out: rb(out))
        if type(mod).__name__ == "Bottleneck" or name == "maxpool":
            mod.register_forward_hook(lambda m, a, out: rb(out))


@pytest.mark.parametrize("mode", ["fp32", "bf16_points"])
def test_whole_network_gradient_matches_reference_fp32_backward(tmp_path, mode):
    """One fp32 `loss.backward()` through the UNMODIFIED reference (LightningModule.training_step, :376-390, its own ResNet-50 /
    regressor / SMPLX / loss in train() mode on CUDA) against the flat gradient buffer `airpose_b200`'s hand-scheduled
    `training_step` fills (bf16 trunk forward and backward, fp32 everywhere else) on the same 4-pair batch, same weights,
    dropout disabled on both sides (p = 0 there, mask=False here).  This is the non-self gradient check: every parameter
    tensor's cosine and relative error are printed; the bounds are those of a bf16 data-gradient chain against fp32 autograd."""
    if not rh.available():
        pytest.skip("reference sources absent (oracle/_ref)")
    from argparse import Namespace
    from airpose_b200 import synthetic
    from airpose_b200.copenet_twoview import copenet_twoview
    B = int(os.environ.get("AIRPOSE_TEST_GRAD_PAIRS", "4"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = synthetic.make_network_state(123, dec_gain=0.01)
    tsd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    batch = rh.make_batch(B, 123, 321, device="cuda")
    if mode == "bf16_points":
        # second mode: the reference gets the CUDA trunk's FORWARD rounding points (same bf16-representable conv weights and
        # images on both sides, activations rounded where the kernels round them) and keeps its fp32 autograd backward -- what
        # remains is the error of the backward kernels alone (bf16 data gradients), not the bf16 forward's different ReLU masks
        for k in list(tsd):
            if tsd[k].dim() == 4:
                tsd[k] = tsd[k].to(torch.bfloat16).float()
        for k in ("im0", "im1"):
            batch[k] = batch[k].to(torch.bfloat16).float()
    # ---- the reference, fp32 autograd
    rt = rh.import_reference("cuda")
    ref = rh.make_module(rt, B, device="cuda", load_weights=False)
    ref.model.load_state_dict(tsd, strict=True)
    ref.train()
    if mode == "bf16_points":
        _install_bf16_rounding_points(ref.model)
    for m in ref.model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    res = ref.training_step({k: v.clone() for k, v in batch.items()}, 1)
    res["loss"].backward()
    gref = {n: p.grad.detach().clone() for n, p in ref.model.named_parameters() if p.grad is not None}
    loss_ref = float(res["loss"])
    del ref
    # ---- ours
    mp = synthetic.write_mean_params(str(tmp_path / "m.npz"))
    synthetic.write_smplx_model(str(tmp_path), 0)
    mod = copenet_twoview(Namespace(smpl_mean_params=mp, smplx_model_dir=str(tmp_path), batch_size=B, val_batch_size=B, reg_iters=3, lr=5e-5))
    mod.model.load_state_dict(tsd, strict=True)
    mod = mod.to("cuda").train()
    opt = mod.configure_optimizers()
    loss, _ = mod.training_step(batch, opt, mask1=False, mask2=False)
    torch.cuda.synchronize()
    rel_loss = abs(float(loss) - loss_ref) / abs(loss_ref)
    rows = []
    dots = np.zeros(3)
    for n, p in mod.model.named_parameters():
        if n not in gref:
            continue
        a, b = p.grad.double().flatten(), gref[n].double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
        rel = float((a - b).abs().max() / (b.abs().max() + 1e-300))
        rows.append((n, cos, rel, float(b.abs().max())))
        dots += [float(a @ b), float(a @ a), float(b @ b)]
    cos_all = dots[0] / np.sqrt(dots[1] * dots[2])
    worst_cos = min(rows, key=lambda r: r[1])
    worst_rel = max(rows, key=lambda r: r[2])
    reg = [r for r in rows if r[0].split(".")[0] in ("fc1", "fc2", "decpose", "decshape")]
    conv = [r for r in rows if "conv" in r[0] or "downsample.0" in r[0]]
    bn = [r for r in rows if r not in reg and r not in conv]
    print("[%s] " % mode, end="")
    print("gradient vs reference fp32 backward, %d pairs: loss %.6g vs %.6g (rel %.2e); %d tensors; cosine over all gradients %.5f; "
          "worst cosine %.4f (%s); worst rel err %.3e (%s); regressor tensors: min cosine %.6f, max rel %.2e" %
          (B, float(loss), loss_ref, rel_loss, len(rows), cos_all, worst_cos[1], worst_cos[0], worst_rel[2], worst_rel[0],
           min(r[1] for r in reg), max(r[2] for r in reg)))
    print("   conv weights (%d): min cosine %.4f, median %.4f, max rel %.2e | BatchNorm affine (%d): min cosine %.4f, median %.4f" %
          (len(conv), min(r[1] for r in conv), float(np.median([r[1] for r in conv])), max(r[2] for r in conv),
           len(bn), min(r[1] for r in bn), float(np.median([r[1] for r in bn]))))
    print("   conv cosines in network order: " + " ".join("%.3f" % r[1] for r in conv))
    for r in sorted(conv, key=lambda r: r[1])[:5] + sorted(bn, key=lambda r: r[1])[:5]:
        print("   %-34s cos %.4f  rel %.3e  |g|max %.3e" % r)
    assert len(rows) == 159 + 8 - 2 or len(rows) >= 150            # every trunk / regressor tensor except deccam
    assert rel_loss < GRAD_LOSS_REL
    assert cos_all > GRAD_COS_ALL
    assert min(r[1] for r in reg) > 0.9999 and max(r[2] for r in reg) < 2e-2
    assert min(r[1] for r in conv) > (GRAD_COS_CONV if mode == "fp32" else GRAD_COS_CONV_SAME_FWD)
    assert min(r[1] for r in bn) > (GRAD_COS_BN if mode == "fp32" else GRAD_COS_BN_SAME_FWD)


# measured on B200 (printed by the test, quoted in DESIGN.md section 4); bounds = measured with a 2x margin on 1 - cos / rel
# Measured (gpurun r02y ... tools/debug_grad.py): loss rel 5.8e-5, cosine over all gradients 0.99998, regressor >= 0.999995 / rel 5.2e-3;
# trunk tensors 0.80-0.96 (fp32 reference) / 0.87-0.97 (reference with the same forward rounding points), lowest in layer1, and the
# same at 24 pairs.  tools/debug_grad.py traced it: the backward kernels reproduce a torch recomputation from THEIR OWN tape to cosine
# 0.999992, and rounding dz to bf16 changes the reference's first weight gradient by 2e-6 -- but two bf16 forwards of this 53-layer
# network in train() mode (batch-statistics BatchNorm, synthetic random weights) differ by cosine 0.998 in the layer4 activations and
# in 1.7 % of the block-output ReLU masks, and THAT moves the first weight gradient to cosine 0.96.  The deviation is the bf16
# forward's, not the backward's; the bounds below are on what is measured.
GRAD_LOSS_REL = 1e-3
GRAD_COS_ALL = 0.9995
GRAD_COS_CONV = 0.6          # measured 0.80
GRAD_COS_BN = 0.45           # measured 0.73
GRAD_COS_CONV_SAME_FWD = 0.6 # reference with the same forward rounding points: set from the measurement below
GRAD_COS_BN_SAME_FWD = 0.45


def test_translation_init_branches_match_reference(tmp_path):
    """copenet_twoview.py:178-188: ground-truth translation on aircapdata test runs, ground truth + noise when
    hparams.smpltrans_noise_sigma is set (the reference's own add_noise_input_smpltrans under the same seed), [0, 0, 10] otherwise."""
    if not rh.available():
        pytest.skip("reference sources absent (oracle/_ref)")
    from argparse import Namespace
    from airpose_b200 import synthetic
    from airpose_b200.copenet_twoview import copenet_twoview
    rh.import_reference("cuda")
    from copenet.utils.utils import add_noise_input_smpltrans
    B = 5
    mp = synthetic.write_mean_params(str(tmp_path / "m.npz"))
    synthetic.write_smplx_model(str(tmp_path), 0)
    batch = rh.make_batch(B, 11, 12, device="cuda")
    hp = dict(smpl_mean_params=mp, smplx_model_dir=str(tmp_path), batch_size=B, val_batch_size=B, reg_iters=3)
    mod = copenet_twoview(Namespace(smpltrans_noise_sigma=0.1, testdata="aerialpeople", **hp)).to("cuda").eval()
    torch.manual_seed(99)
    (s0, s1), (u0, u1) = mod._init_translation(B, torch.device("cuda"), batch)
    torch.manual_seed(99)
    r0, _ = add_noise_input_smpltrans(batch["smpltrans_rel0"], 0.1)
    r1, _ = add_noise_input_smpltrans(batch["smpltrans_rel1"], 0.1)
    assert torch.equal(u0, r0) and torch.equal(u1, r1)
    assert torch.equal(s0, r0 * 0.05) and torch.equal(s1, r1 * 0.05)
    out = mod.fwd_pass(batch)                                   # runs end to end with per-view initial translations
    assert torch.isfinite(out["pred_vertices_cam0"]).all() and out["in_smpltrans1"].shape == (B, 3)
    mod = copenet_twoview(Namespace(smpltrans_noise_sigma=None, testdata="aircapdata", **hp)).to("cuda").eval()
    (s0, s1), (u0, u1) = mod._init_translation(B, torch.device("cuda"), batch, is_test=True)
    assert torch.equal(u0, batch["smpltrans_rel0"]) and torch.equal(u1, batch["smpltrans_rel1"])
    (s0, _), (u0, _) = mod._init_translation(B, torch.device("cuda"), batch, is_test=False)
    assert torch.equal(u0, torch.tensor([0.0, 0.0, 10.0], device="cuda").expand(B, 3)) and torch.equal(s0, u0 * 0.05)
    with pytest.raises(KeyError):
        mod._init_translation(B, torch.device("cuda"), {"im0": batch["im0"]}, is_test=True)
