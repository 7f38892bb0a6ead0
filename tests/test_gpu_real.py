"""GPU parity of the copenet_real variant (VPoser-free part): `airpose_real_loss` against the reference's
own get_loss (tests/golden/real_loss.npz, made by oracle/gen_golden_real.py from the UNMODIFIED function) and the numpy oracle;
the backward chain loss -> projection (per-camera focal lengths) -> transform_smpl -> SMPL-X -> rot6d against fp64 autograd over
the PyTorch port; a short regressor fine-tuning run (the reference's train_reg_only mode, copenet_real/.../copenet_twoview.py:357-372)."""
import os
from argparse import Namespace
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import airpose_oracle as orc
from airpose_b200 import synthetic
from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.fixture(scope="module")
def real_module(tmp_path_factory, net_state):
    from airpose_b200.copenet_real import copenet_twoview
    d = tmp_path_factory.mktemp("real")
    mp = synthetic.write_mean_params(str(d / "smpl_mean_params.npz"))
    synthetic.write_smplx_model(str(d), 0)
    mod = copenet_twoview(Namespace(smpl_mean_params=mp, smplx_model_dir=str(d), batch_size=4, val_batch_size=4, reg_iters=3, lr=5e-5))
    mod.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synthetic.make_network_state(123, dec_gain=0.01).items()})
    return mod.to(DEV).eval()


@pytest.mark.parametrize("B", [1, 6])
def test_real_loss_matches_reference_golden(real_module, B):
    g = np.load(os.path.join(GOLDEN, "real_loss.npz"))
    p = "b%d/" % B
    hp = {k[3:]: float(g[k]) for k in g.files if k.startswith("hp/")}
    for k, v in hp.items():
        setattr(real_module.hparams, k, v)
    c = {k[len(p):]: g[k] for k in g.files if k.startswith(p) and "/" not in k[len(p):]}
    pose0 = np.zeros((B, 135), np.float32); pose0[:, :3] = c["pred_smpltrans0"]            # a strided view, like pred_pose[:, :3]
    pose1 = np.zeros((B, 135), np.float32); pose1[:, :3] = c["pred_smpltrans1"]
    batch = {"smpl_joints_2d0": t(c["smpl_joints_2d0"]), "smpl_joints_2d1": t(c["smpl_joints_2d1"])}
    loss, losses, grads = real_module.get_loss(batch, t(pose0)[:, :3], t(pose1)[:, :3], t(c["pred_rotmat0"]), t(c["pred_rotmat1"]),
                                               t(c["pred_betas0"]), t(c["pred_betas1"]), None, None,
                                               t(c["pred_joints_2d_cam0"]), t(c["pred_joints_2d_cam1"]), with_grads=True)
    ref = float(g[p + "loss"])
    assert abs(float(loss) - ref) <= 2e-6 * abs(ref)
    o_loss, o_losses = orc.real_get_loss(hp, c, c)
    assert abs(float(loss) - o_loss) <= 2e-6 * abs(o_loss)
    for k in ("loss_regr_pose", "loss_keypoints", "loss_regul_betas", "loss_regul_vposer"):
        assert abs(float(losses[k]) - float(g[p + "losses/" + k])) <= 2e-6 * abs(float(g[p + "losses/" + k])) + 1e-12, k
    # gradients: the reference's own autograd through its own get_loss
    worst = 0.0
    for ours, theirs in (("joints_2d0", "pred_joints_2d_cam0"), ("joints_2d1", "pred_joints_2d_cam1"), ("rotmat0", "pred_rotmat0"),
                         ("rotmat1", "pred_rotmat1"), ("betas0", "pred_betas0"), ("betas1", "pred_betas1"),
                         ("smpltrans0", "pred_smpltrans0"), ("smpltrans1", "pred_smpltrans1")):
        e = rel_err(grads[ours].cpu().numpy(), g[p + "grad/" + theirs])
        worst = max(worst, e)
        assert e < 1e-5, (ours, e)
    print("real loss B=%d: %.4f vs reference %.4f; worst gradient rel err %.2e" % (B, float(loss), ref, worst))


def _port_real_loss(tp, m, raw, gt, x, B, hp, focal):
    """copenet_real fwd (:205-307) + get_loss (:99-160) without the VPoser term on the PyTorch port, fp64, differentiable."""
    mse = lambda a, b: (a - b) ** 2
    P = {}
    for v in (0, 1):
        pose = raw["pose%d" % v]
        trans = pose[:, :3] / 0.05
        R = tp.rot6d_to_rotmat(pose[:, 3:]).view(B, 22, 3, 3)
        verts, joints = tp.smplx_forward(m, raw["betas%d" % v], R[:, 1:])
        jc = torch.bmm(R[:, 0], joints.permute(0, 2, 1)).permute(0, 2, 1) + trans[:, None]
        c = torch.tensor(x["intr%d" % v][:, :2, 2], dtype=torch.float64)
        fx, fy = focal[v]
        j2d = torch.stack([fx * jc[:, :, 0] / jc[:, :, 2] + c[:, None, 0], fy * jc[:, :, 1] / jc[:, :, 2] + c[:, None, 1]], -1)
        P[v] = dict(trans=trans, R=R, j2d=j2d, betas=raw["betas%d" % v])
    G = [torch.tensor(gt["smpl_joints_2d%d" % v][:, 0], dtype=torch.float64) for v in (0, 1)]
    w = torch.ones(22, dtype=torch.float64); w[[4, 5, 18, 19]] = hp["limbs2d_loss_weight"]; w[[7, 8, 20, 21]] = hp["limbs2d_loss_weight"] ** 2
    lk = sum(mse(P[v]["j2d"][:, :22], G[v][:, :22, :2]) * G[v][:, :22, 2:] for v in (0, 1))
    l_kp = (lk * w.view(1, 22, 1)).mean()
    l_pose = mse(P[0]["R"][:, 1:], P[1]["R"][:, 1:]).mean()
    b0, b1 = P[0]["betas"], P[1]["betas"]
    l_beta = (b0 * b0).mean() + (b1 * b1).mean() + mse(b0, b1).mean()
    return 60 * (hp["keypoint2d_loss_weight"] * l_kp + hp["beta_loss_weight"] * l_beta + hp["pose_loss_weight"] * l_pose
                 + (torch.exp(-P[0]["trans"][:, 2]) ** 2).mean() + (torch.exp(-P[1]["trans"][:, 2]) ** 2).mean())


def _real_batch(B, seed):
    x = synthetic.make_inputs(B, seed)
    rng = np.random.default_rng(seed)
    gt = {"smpl_joints_2d%d" % v: np.concatenate([(rng.standard_normal((B, 1, 25, 2)) * 80 + 800).astype(np.float32),
                                                  rng.uniform(0.2, 1.0, (B, 1, 25, 1)).astype(np.float32)], -1) for v in (0, 1)}
    return x, gt


def test_real_loss_and_head_backward_matches_autograd(real_module, smplx_data):
    import torch_port as tp
    from airpose_b200.copenet_real import FOCAL_LENGTH0, FOCAL_LENGTH1, REAL_LOSS_WEIGHTS
    for k, v in REAL_LOSS_WEIGHTS.items():
        setattr(real_module.hparams, k, v)
    m = tp.Smplx(smplx_data)
    for k in ("v_template", "shapedirs", "J_regressor", "weights", "posedirs", "lmk_bary"):
        setattr(m, k, getattr(m, k).double())
    B = 3
    x, gt = _real_batch(B, 31)
    batch = {k: t(v) for k, v in {**x, **gt}.items()}
    out = real_module.fwd_pass(batch)
    # per-camera focal lengths reach the projection (copenet_real/.../copenet_twoview.py:297-307)
    jc = out["pred_joints_cam0"].cpu().double().numpy()
    c0 = x["intr0"][:, :2, 2]
    j2 = np.stack([FOCAL_LENGTH0[0] * jc[..., 0] / jc[..., 2] + c0[:, None, 0], FOCAL_LENGTH0[1] * jc[..., 1] / jc[..., 2] + c0[:, None, 1]], -1)
    assert rel_err(out["pred_joints_2d_cam0"].cpu().numpy(), j2) < 1e-5
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        raw = {}
        for v in (0, 1):
            p = out["pred_pose%d" % v].clone()
            p[:, :3] *= 0.05
            raw["pose%d" % v] = p.cpu().double().requires_grad_(True)
            raw["betas%d" % v] = out["pred_betas%d" % v].cpu().double().requires_grad_(True)
        ref = _port_real_loss(tp, m, raw, gt, x, B, REAL_LOSS_WEIGHTS, (FOCAL_LENGTH0, FOCAL_LENGTH1))
        ref.backward()
    finally:
        torch.set_default_dtype(old)
    loss, losses, grads = real_module.loss_and_head_backward(batch, out)
    assert abs(float(loss) - float(ref)) <= 2e-4 * abs(float(ref))
    for v in (0, 1):
        ep = rel_err(grads["pred_pose%d" % v].cpu().numpy(), raw["pose%d" % v].grad.numpy())
        eb = rel_err(grads["pred_betas%d" % v].cpu().numpy(), raw["betas%d" % v].grad.numpy())
        print("copenet_real view %d: d loss/d pred_pose rel err %.3e, d loss/d pred_betas rel err %.3e" % (v, ep, eb))
        assert ep < 1e-3 and eb < 1e-3


def test_real_fine_tuning_reduces_the_loss(real_module):
    """train_reg_only on the real-data loss: fc1/fc2/decpose/decshape move, the trunk stays bit-identical, the loss on a
    fixed batch goes down."""
    from airpose_b200.copenet_real import REAL_LOSS_WEIGHTS
    for k, v in REAL_LOSS_WEIGHTS.items():
        setattr(real_module.hparams, k, v)
    real_module.hparams.lr = 2e-4
    B = 4
    x, gt = _real_batch(B, 41)
    batch = {k: t(v) for k, v in {**x, **gt}.items()}
    trunk0 = real_module.model.layer4[2].conv3.weight.detach().clone()
    opt = real_module.configure_optimizers_reg_only()
    hist = []
    for _ in range(12):
        loss, losses = real_module.training_step_reg_only(batch, opt, mask1=False, mask2=False)
        hist.append(float(loss))
    print("copenet_real fine-tuning: loss %.1f -> %.1f over 12 steps" % (hist[0], hist[-1]))
    assert np.isfinite(hist).all() and hist[-1] < 0.9 * hist[0]
    assert torch.equal(trunk0, real_module.model.layer4[2].conv3.weight.detach())
    assert set(losses) == {"loss", "loss_regul_vposer", "loss_regr_pose", "loss_keypoints", "loss_regul_betas"}
