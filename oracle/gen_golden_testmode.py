"""Generate tests/golden/testmode_b2.npz: the is_test branch of the UNMODIFIED reference LightningModule
(copenet/src/copenet/copenet_twoview.py:258-279,318-350) run on CPU in the build container.

TEST INFRASTRUCTURE ONLY.   python oracle/gen_golden_testmode.py

torchgeometry is not installable offline, so the two calls the branch makes into it (`tgm.rotation_matrix_to_angle_axis`) are served
by the oracle's restatement (airpose_oracle.tgm_rotation_matrix_to_angle_axis) through the import shim -- those four outputs
(pred_angles*, gt_angles*) are therefore NOT pinned by this file and are not stored.  Everything else the branch returns is the
reference's own arithmetic: stored are the zero-beta meshes at the input translation (pred_vertices_cam_in*), the camera-frame
joints, the translations, and the inputs needed to regenerate the batch (seeds).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from airpose_b200 import synthetic  # noqa: E402
import airpose_oracle as orc  # noqa: E402
import gen_golden  # noqa: E402
import ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NET_SEED, IN_SEED, GT_SEED = 123, 41, 9


def main():
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref_stubs.install()
    tgm = sys.modules["torchgeometry"]
    tgm.rotation_matrix_to_angle_axis = lambda R: torch.from_numpy(orc.tgm_rotation_matrix_to_angle_axis(R.detach().cpu().numpy()))
    sys.path.insert(0, gen_golden.REF_SRC)
    import copenet.config as ref_config
    ref_config.device = "cpu"
    import torchvision.models.resnet as tv_resnet
    _orig = tv_resnet.resnet50
    tv_resnet.resnet50 = lambda pretrained=False, **k: _orig(weights=None)
    from copenet import copenet_twoview as ref_twoview
    from argparse import Namespace

    B = 2
    tmp = gen_golden.make_home(tempfile.mkdtemp(prefix="airpose_home_"))
    hp = Namespace(copenet_home=tmp, batch_size=B, val_batch_size=B, testdata="aerialpeople", smpltrans_noise_sigma=None, reg_iters=3,
                   shape_loss_weight=50, keypoint2d_loss_weight=0.002, keypoint3d_loss_weight=1, limbs3d_loss_weight=3.0,
                   limbstheta_loss_weight=1.0, trans_loss_weight=10, rootrot_loss_weight=1, pose_loss_weight=50, beta_loss_weight=1)
    module = ref_twoview.copenet_twoview(hp)
    sd = synthetic.make_network_state(NET_SEED)
    module.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    module.eval()
    x = synthetic.make_inputs(B, IN_SEED)
    li = synthetic.make_lbs_inputs(B, seed=GT_SEED)
    rng = np.random.default_rng(GT_SEED)
    orient = [synthetic.rot6d_to_rotmat_np(np.array([1, 0, 0, 1, 0, 0], np.float32) + rng.standard_normal((B, 6)).astype(np.float32) * 0.3)[:, None]
              for _ in (0, 1)]
    gt = {"smplpose_rotmat": li["body_pose"], "smplorient_rel0": orient[0], "smplorient_rel1": orient[1],
          "smpl_vertices": np.zeros((B, 1, 10475, 3), np.float32), "smpl_joints": np.zeros((B, 1, 127, 3), np.float32),
          "smpl_joints_2d0": np.zeros((B, 1, 127, 2), np.float32), "smpl_joints_2d1": np.zeros((B, 1, 127, 2), np.float32)}
    batch = {k: torch.from_numpy(v) for k, v in {**x, **gt}.items()}
    with torch.no_grad():
        output, losses, loss = module.fwd_pass_and_loss(batch, is_val=True, is_test=True)
        xf0 = module.model.forward_feat_ext(batch["im0"]).numpy()
        xf1 = module.model.forward_feat_ext(batch["im1"]).numpy()
    assert loss is None and losses is None
    keys = sorted(output)
    save = {"batch": B, "net_seed": NET_SEED, "in_seed": IN_SEED, "gt_seed": GT_SEED, "keys": np.array(keys),
            "smplorient_rel0": orient[0], "smplorient_rel1": orient[1], "xf0": xf0, "xf1": xf1}
    for k in ("pred_vertices_cam_in0", "pred_vertices_cam_in1", "pred_j3d_cam0", "pred_j3d_cam1", "pred_smpltrans0", "pred_smpltrans1",
              "in_smpltrans0", "in_smpltrans1", "pred_betas0", "pred_betas1", "gt_smpltrans0", "gt_smpltrans1"):
        save[k] = output[k].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "testmode_b2.npz"), **save)
    print("testmode_b2.npz:", keys)


if __name__ == "__main__":
    main()
