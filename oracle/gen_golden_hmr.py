"""Generate tests/golden/hmr_b2.npz by running the REAL reference hmr network (model_hmr.py) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container:   python oracle/gen_golden_hmr.py

BASELINE.json configs[0] ("hmr single-view fwd batch=1 224x224 on CPU") is the reference's own
CPU-runnable case.  The fixture holds, for a seeded batch of 2 images (image 0 alone IS the batch-1
case: eval-mode BatchNorm makes images independent), the outputs of the unmodified
``copenet.models.model_hmr.getcopenet`` module and of the reference's own functions called in the order of
``hmr.fwd_pass_and_loss`` (copenet/src/copenet/hmr.py:127-158), in fp32 and with bf16 rounding points
hooked into the trunk (same hooks as gen_golden.py).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (also puts the repo root on sys.path)
from airpose_b200 import synthetic  # noqa: E402
import ref_stubs  # noqa: E402


def main():
    import torch
    torch.manual_seed(0)
    ref_stubs.install()
    sys.path.insert(0, gg.REF_SRC)
    import copenet.config as ref_config
    ref_config.device = "cpu"
    from copenet.models import model_hmr as ref_hmr
    from copenet.smplx.smplx import SMPLX
    from copenet.utils.geometry import perspective_projection
    from copenet.utils.utils import transform_smpl

    tmp = gg.make_home(tempfile.mkdtemp(prefix="airpose_home_"))
    smplx_dir = os.path.join(tmp, "src", "copenet", "data", "smplx", "models", "smplx")
    mean_params = os.path.join(tmp, "src", "copenet", "data", "smpl_mean_params.npz")
    B = 2
    net = ref_hmr.getcopenet(mean_params, pretrained=False)
    sd_np = synthetic.make_network_state(gg.NET_SEED, variant="hmr")
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)
    net.eval()
    x = torch.from_numpy(synthetic.make_inputs(B, gg.IN_SEED)["im0"])
    sm = SMPLX(smplx_dir, batch_size=B, create_transl=False)
    eye = torch.eye(3).view(1, 1, 3, 3).repeat(B, 1, 1, 1)
    focal = [1475, 1475]

    def run():
        with torch.no_grad():
            xf = net.forward_feat_ext(x)
            rotmat, betas, cam = net.forward(x=x, iters=3)                                     # hmr.py:135-136
            o = sm.forward(betas=betas, body_pose=rotmat[:, 1:], global_orient=eye, transl=torch.zeros(B, 3), pose2rot=False)
            tm = torch.cat([rotmat[:, :1].squeeze(1), torch.zeros(B, 3).unsqueeze(2)], dim=2)   # :144
            pv, pj, _, _ = transform_smpl(tm, o.vertices.squeeze(1), o.joints.squeeze(1))       # :146-148
            cam_t = torch.stack([cam[:, 1], cam[:, 2], 2 * focal[0] / (224 * cam[:, 0] + 1e-9)], dim=-1)   # :149-151
            j2d = perspective_projection(pj, rotation=torch.eye(3).unsqueeze(0).repeat(B, 1, 1), translation=cam_t,
                                         focal_length=focal, camera_center=torch.zeros(B, 2))   # :153-157
        return {"xf": xf.numpy(), "pred_rotmat": rotmat.numpy(), "pred_betas": betas.numpy(), "pred_camera": cam.numpy(),
                "pred_cam_t": cam_t.numpy(), "vertices": o.vertices.numpy(), "joints": o.joints.numpy(),
                "pred_vertices": pv.numpy(), "pred_joints": pj.numpy(), "pred_joints_2d_cam": j2d.numpy()}

    fp32 = run()
    one = None
    with torch.no_grad():                          # config 1 proper: batch of one image
        r1, b1, c1 = net.forward(x=x[:1], iters=3)
        one = {"pred_rotmat": r1.numpy(), "pred_betas": b1.numpy(), "pred_camera": c1.numpy()}
    undo = gg.install_bf16_hooks(net)
    bf = run()
    undo()
    save = {"batch": B, "net_seed": gg.NET_SEED, "in_seed": gg.IN_SEED, "smplx_seed": gg.SMPLX_SEED}
    save.update({"fp32/" + k: v for k, v in fp32.items()})
    save.update({"b1/" + k: v for k, v in one.items()})
    for k in ("xf", "pred_rotmat", "pred_betas", "pred_camera", "pred_joints_2d_cam"):
        save["bf16/" + k] = bf[k]
    np.savez(os.path.join(gg.GOLDEN, "hmr_b2.npz"), **save)
    print("hmr_b2.npz written; cam", fp32["pred_camera"], "cam_t", fp32["pred_cam_t"])
    print("batch-1 vs batch-2 image 0: max abs diff rotmat %.3e" % np.abs(one["pred_rotmat"] - fp32["pred_rotmat"][:1]).max())
    print("bf16-vs-fp32: rotmat max abs %.3e, j2d max abs %.3e px" %
          (np.abs(bf["pred_rotmat"] - fp32["pred_rotmat"]).max(), np.abs(bf["pred_joints_2d_cam"] - fp32["pred_joints_2d_cam"]).max()))


if __name__ == "__main__":
    main()
