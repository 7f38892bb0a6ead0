"""Host-side mirror of the reference LightningModule's per-frame forward path.

``copenet_twoview.fwd_pass`` follows copenet/src/copenet/copenet_twoview.py:164-317
(``fwd_pass_and_loss`` up to the loss): initial translation [0,0,10]*0.05, the two-view
network, translation un-scaling, 6D -> rotation matrices, SMPL-X with identity global
orientation, rigid transform about the origin, perspective projection.  SMPL-X,
``transform_smpl`` and ``perspective_projection`` run fused in one native call per view.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import model_copenet
from .smplx import SMPLX, rot6d_to_rotmat

FOCAL_LENGTH = [1475, 1475]          # copenet/src/copenet/constants.py:7
TRANS_SCALE = 0.05                   # copenet_twoview.py:199-203


class copenet_twoview(nn.Module):
    """``hparams`` needs: copenet_home-style paths are replaced by explicit ones:
    ``smpl_mean_params`` (npz path), ``smplx_model_dir``, ``batch_size``, ``val_batch_size``,
    ``reg_iters``."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.model = model_copenet.getcopenet(hparams.smpl_mean_params, pretrained=getattr(hparams, "pretrained", False))
        # the reference keeps two SMPLX instances (train / val batch size) as module globals
        # (copenet_twoview.py:33-45); one instance serves any batch here.
        self.smplx = SMPLX(hparams.smplx_model_dir, batch_size=hparams.batch_size, create_transl=False)
        self.focal_length = FOCAL_LENGTH

    def forward(self, **kwargs):
        return self.model(**kwargs)

    @torch.no_grad()
    def fwd_pass(self, input_batch, profile=None):
        """``profile``: optional dict that receives CUDA event pairs around the trunk, the regressor
        and the two SMPL-X calls (bench.py reads the stage times from them)."""

        def mark(name, done=False):
            if profile is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                profile.setdefault(name, []).append(ev)

        im0, im1 = input_batch["im0"].float(), input_batch["im1"].float()
        bb0, bb1 = input_batch["bb0"], input_batch["bb1"]
        intr = (input_batch["intr0"], input_batch["intr1"])
        B = im0.shape[0]
        dev = im0.device
        in_trans = torch.tensor([0.0, 0.0, 10.0], device=dev).expand(B, -1).clone() * TRANS_SCALE      # :184-203
        reg_iters = getattr(self.hparams, "reg_iters", 3)
        mark("trunk")
        xf = self.model.forward_feat_ext_pair(im0, im1)                         # both views, one call (eval-mode BN)
        mark("trunk")
        mark("ief")
        pred = self.model._ief(xf[:B], xf[B:], bb0, bb1, in_trans, in_trans, None, None, None, None, reg_iters)
        mark("ief")
        out = {}
        mark("smplx")
        for v in (0, 1):
            pose, betas = pred[2 * v], pred[2 * v + 1]
            pose[:, :3] /= TRANS_SCALE                                   # in-place on the view, like :214-218
            trans = pose[:, :3]
            rotmat = rot6d_to_rotmat(pose[:, 3:]).view(B, 22, 3, 3)      # :222-223
            mo, cam = self.smplx.forward_camera(
                betas=betas, body_pose=rotmat[:, 1:],
                global_orient=None,                                       # identity (:283)
                transl=None, pose2rot=False,                              # the reference passes zeros (:284); None skips the add
                root_R=rotmat[:, 0], root_t=trans,                        # transform_smpl (:287-292)
                focal_length=self.focal_length, camera_center=intr[v][:, :2, 2])   # :307-317
            out.update({"pred_pose%d" % v: pose, "pred_betas%d" % v: betas, "pred_rotmat%d" % v: rotmat,
                        "pred_smpltrans%d" % v: trans, "in_smpltrans%d" % v: in_trans / TRANS_SCALE,
                        "pred_output_cam%d" % v: mo,
                        "pred_vertices_cam%d" % v: cam["vertices_cam"], "pred_joints_cam%d" % v: cam["joints_cam"],
                        "pred_joints_2d_cam%d" % v: cam["joints_2d"]})
        mark("smplx")
        return out
