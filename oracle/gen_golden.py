"""Generate tests/golden/*.npz by running the REAL reference (from /root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not exist on
the GPU box):   python oracle/gen_golden.py

The reference ships no golden vectors for this path; these
fixtures are "reference code executed on seeded synthetic inputs", frozen here.  Inputs
are NOT stored when they can be regenerated bit-exactly from
``airpose_b200.synthetic`` (numpy default_rng); outputs are.

Fixtures:
  smplx_lbs.npz   reference SMPLX.forward(pose2rot=False) on seeded betas / body rotations
                  (copenet/src/copenet/smplx/smplx/body_models.py:820-994)
  twoview_b2.npz  the unmodified copenet_twoview LightningModule (import shims from
                  oracle/ref_stubs.py) run through fwd_pass_and_loss(is_val=True) at B=2,
                  plus the intermediates obtained by calling the reference's own
                  functions in the same order (copenet_twoview.py:164-317), in fp32 and
                  with bf16 rounding points injected by hooks on the reference modules.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF_SRC = "/root/reference/copenet/src"

from airpose_b200 import synthetic  # noqa: E402
import ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SMPLX_SEED, NET_SEED, IN_SEED, GT_SEED = 0, 123, 123, 321


def make_home(tmp):
    """Lay out a synthetic ``copenet_home`` (copenet_twoview.py:60-68 path conventions)."""
    data = os.path.join(tmp, "src", "copenet", "data")
    synthetic.write_mean_params(os.path.join(data, "smpl_mean_params.npz"))
    synthetic.write_smplx_model(os.path.join(data, "smplx", "models", "smplx"), SMPLX_SEED)
    rng = np.random.default_rng(5)
    with open(os.path.join(data, "smplx", "MANO_SMPLX_vertex_ids.pkl"), "wb") as f:
        pickle.dump({"left_hand": rng.choice(synthetic.NUM_VERTS, 778, replace=False),
                     "right_hand": rng.choice(synthetic.NUM_VERTS, 778, replace=False)}, f)
    np.save(os.path.join(data, "smplx", "SMPL-X__FLAME_vertex_ids.npy"),
            rng.choice(synthetic.NUM_VERTS, 5023, replace=False))
    return tmp


def round_bf16_t(t):
    import torch
    return t.to(torch.bfloat16).to(torch.float32)


def install_bf16_hooks(net):
    """Inject the CUDA path's rounding points into the reference trunk (DESIGN.md 'numerics'):
    conv operands bf16, fp32 accumulate; every stored activation rounded once."""
    import torch.nn as nn
    handles = []
    saved = {}
    for name, mod in net.named_modules():
        if isinstance(mod, nn.Conv2d):
            saved[name] = mod.weight.data.clone()
            mod.weight.data = round_bf16_t(mod.weight.data)
            handles.append(mod.register_forward_pre_hook(lambda m, a: (round_bf16_t(a[0]),)))
        if type(mod).__name__ == "Bottleneck" or name.endswith("downsample") or name == "maxpool":
            handles.append(mod.register_forward_hook(lambda m, a, out: round_bf16_t(out)))

    def undo():
        for h in handles:
            h.remove()
        for name, mod in net.named_modules():
            if name in saved:
                mod.weight.data = saved[name]
    return undo


def main():
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref_stubs.install()
    sys.path.insert(0, REF_SRC)
    import copenet.config as ref_config
    ref_config.device = "cpu"                     # the reference hard-codes "cuda" (config.py:7)
    import torchvision.models.resnet as tv_resnet
    _orig_resnet50 = tv_resnet.resnet50
    tv_resnet.resnet50 = lambda pretrained=False, **k: _orig_resnet50(weights=None)   # no network here
    from copenet import copenet_twoview as ref_twoview
    from copenet.smplx.smplx import SMPLX
    from copenet.utils.geometry import rot6d_to_rotmat, perspective_projection
    from copenet.utils.utils import transform_smpl

    os.makedirs(GOLDEN, exist_ok=True)
    tmp = make_home(tempfile.mkdtemp(prefix="airpose_home_"))
    smplx_dir = os.path.join(tmp, "src", "copenet", "data", "smplx", "models", "smplx")

    # ---------------- smplx_lbs.npz ----------------
    B = 3
    li = synthetic.make_lbs_inputs(B, seed=7)
    sm = SMPLX(smplx_dir, batch_size=B, create_transl=False)
    with torch.no_grad():
        eye = torch.eye(3).view(1, 1, 3, 3).repeat(B, 1, 1, 1)
        o = sm.forward(betas=torch.from_numpy(li["betas"]), body_pose=torch.from_numpy(li["body_pose"]),
                       global_orient=eye, transl=torch.zeros(B, 3), pose2rot=False)
        # reduced call that falls back to the module's zero betas (copenet_twoview.py:575-582)
        o0 = sm.forward(body_pose=torch.from_numpy(li["body_pose"]), global_orient=eye, pose2rot=False)
        # zero pose, zero betas -> v_template (SURVEY.md 8(c) KAT 1)
        oz = sm.forward(betas=torch.zeros(B, 10), body_pose=eye.repeat(1, 21, 1, 1), global_orient=eye,
                        pose2rot=False)
    np.savez(os.path.join(GOLDEN, "smplx_lbs.npz"),
             lbs_seed=7, batch=B,
             vertices=o.vertices.numpy(), joints=o.joints.numpy(),
             vertices_zero_betas=o0.vertices.numpy()[:1], joints_zero_betas=o0.joints.numpy(),
             joints_rest=oz.joints.numpy()[:1],
             rest_vertex_max_abs_err=np.abs(oz.vertices.numpy() - sm.v_template.numpy()[None]).max())
    print("smplx_lbs.npz written; rest-pose max err", np.abs(oz.vertices.numpy() - sm.v_template.numpy()[None]).max())

    # ---------------- twoview_b2.npz ----------------
    B = 2
    from argparse import Namespace
    hp = Namespace(copenet_home=tmp, batch_size=B, val_batch_size=B, testdata="aerialpeople",
                   smpltrans_noise_sigma=None, reg_iters=3, shape_loss_weight=50, keypoint2d_loss_weight=0.002,
                   keypoint3d_loss_weight=1, limbs3d_loss_weight=3.0, limbstheta_loss_weight=1.0,
                   trans_loss_weight=10, rootrot_loss_weight=1, pose_loss_weight=50, beta_loss_weight=1)
    module = ref_twoview.copenet_twoview(hp)
    sd_np = synthetic.make_network_state(NET_SEED)
    module.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)
    module.eval()

    batch_np = synthetic.make_inputs(B, IN_SEED)
    # ground truth for the loss: SMPL-X forward of an independent seeded sample (section 8(d))
    gt_in = synthetic.make_lbs_inputs(B, seed=GT_SEED)
    rng = np.random.default_rng(GT_SEED)
    gt_orient = [synthetic.rot6d_to_rotmat_np(np.array([1, 0, 0, 1, 0, 0], np.float32)
                                              + rng.standard_normal((B, 6)).astype(np.float32) * 0.3)
                 for _ in (0, 1)]
    sm2 = SMPLX(smplx_dir, batch_size=B, create_transl=False)
    with torch.no_grad():
        eye = torch.eye(3).view(1, 1, 3, 3).repeat(B, 1, 1, 1)
        g = sm2.forward(betas=torch.from_numpy(gt_in["betas"]), body_pose=torch.from_numpy(gt_in["body_pose"]),
                        global_orient=eye, transl=torch.zeros(B, 3), pose2rot=False)
    gt = {"smplpose_rotmat": gt_in["body_pose"],
          "smplorient_rel0": gt_orient[0][:, None], "smplorient_rel1": gt_orient[1][:, None],
          "smpl_vertices": g.vertices.numpy()[:, None], "smpl_joints": g.joints.numpy()[:, None]}
    for v in (0, 1):
        tm = torch.cat([torch.from_numpy(gt_orient[v]), torch.from_numpy(batch_np["smpltrans_rel%d" % v])[:, :, None]], 2)
        _, jc, _, _ = transform_smpl(tm, g.vertices, g.joints)
        j2 = perspective_projection(jc, torch.eye(3).repeat(B, 1, 1), torch.zeros(B, 3), [1475, 1475],
                                    torch.from_numpy(batch_np["intr%d" % v][:, :2, 2]).unsqueeze(0))
        gt["smpl_joints_2d%d" % v] = j2.numpy()[:, None]
    batch_t = {k: torch.from_numpy(v) for k, v in {**batch_np, **gt}.items()}

    with torch.no_grad():
        output, losses, loss = module.fwd_pass_and_loss(batch_t, is_val=True, is_test=False)

    def manual(net):
        """Same sequence with the reference's own functions, to expose intermediates."""
        res = {}
        with torch.no_grad():
            init = torch.tensor([0.0, 0.0, 10.0]).expand(B, -1).clone() * 0.05
            xf0 = net.forward_feat_ext(batch_t["im0"])
            xf1 = net.forward_feat_ext(batch_t["im1"])
            p0, s0, p1, s1 = net(x0=batch_t["im0"], x1=batch_t["im1"], bb0=batch_t["bb0"], bb1=batch_t["bb1"],
                                 init_position0=init, init_position1=init.clone(), iters=3)
            res["xf0"], res["xf1"] = xf0.numpy(), xf1.numpy()
            for v, (p, s) in enumerate(((p0, s0), (p1, s1))):
                p = p.clone()
                p[:, :3] /= 0.05
                R = rot6d_to_rotmat(p[:, 3:]).view(B, 22, 3, 3)
                o = ref_twoview.smplx_test.forward(betas=s, body_pose=R[:, 1:], global_orient=eye,
                                                   transl=torch.zeros(B, 3), pose2rot=False)
                tm = torch.cat([R[:, 0], p[:, :3].unsqueeze(2)], dim=2)
                vc, jc, _, _ = transform_smpl(tm, o.vertices, o.joints)
                j2d = perspective_projection(jc, torch.eye(3).repeat(B, 1, 1), torch.zeros(B, 3), [1475, 1475],
                                             batch_t["intr%d" % v][:, :2, 2].unsqueeze(0))
                res.update({"pred_pose%d" % v: p.numpy(), "pred_betas%d" % v: s.numpy(),
                            "pred_rotmat%d" % v: R.numpy(), "vertices%d" % v: o.vertices.numpy(),
                            "joints%d" % v: o.joints.numpy(), "pred_vertices_cam%d" % v: vc.numpy(),
                            "pred_joints_cam%d" % v: jc.numpy(), "pred_joints_2d_cam%d" % v: j2d.numpy()})
        return res

    fp32 = manual(module.model)
    for v in (0, 1):   # the LightningModule output must agree with the manual sequence
        d = np.abs(output["pred_vertices_cam%d" % v].numpy() - fp32["pred_vertices_cam%d" % v]).max()
        assert d == 0.0, d
    undo = install_bf16_hooks(module.model)
    bf = manual(module.model)
    undo()
    chk = manual(module.model)
    assert np.array_equal(chk["xf0"], fp32["xf0"])

    save = {"batch": B, "net_seed": NET_SEED, "in_seed": IN_SEED, "smplx_seed": SMPLX_SEED,
            "loss": float(loss), **{"loss/" + k: v for k, v in losses.items()}}
    save.update({"gt/" + k: v for k, v in gt.items()})
    save.update({"fp32/" + k: v for k, v in fp32.items()})
    for k in ("xf0", "xf1", "pred_pose0", "pred_pose1", "pred_betas0", "pred_betas1",
              "pred_joints_2d_cam0", "pred_joints_2d_cam1", "pred_joints_cam0", "pred_joints_cam1"):
        save["bf16/" + k] = bf[k]
    np.savez(os.path.join(GOLDEN, "twoview_b2.npz"), **save)
    print("twoview_b2.npz written; loss", float(loss))
    print("xf stats: mean %.4f std %.4f max %.4f" % (fp32["xf0"].mean(), fp32["xf0"].std(), np.abs(fp32["xf0"]).max()))
    print("bf16-vs-fp32 xf rel err: %.4e" % (np.abs(bf["xf0"] - fp32["xf0"]).max() / np.abs(fp32["xf0"]).max()))
    print("pose delta over 3 iters (max |pose - init|): %.4f" %
          np.abs(fp32["pred_pose0"][:, 3:] - sd_np["init_pose"][:, :132]).max())
    print("pred trans", fp32["pred_pose0"][:, :3])
    print("bf16-vs-fp32 pose max abs err: %.4e ; j2d max abs err %.4e px" %
          (np.abs(bf["pred_pose0"] - fp32["pred_pose0"]).max(),
           np.abs(bf["pred_joints_2d_cam0"] - fp32["pred_joints_2d_cam0"]).max()))


if __name__ == "__main__":
    main()
