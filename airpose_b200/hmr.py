"""Host-side mirror of the reference's single-view ``hmr`` LightningModule forward (BASELINE config 1).

``hmr.fwd_pass`` follows copenet/src/copenet/hmr.py:127-158 (``fwd_pass_and_loss`` up to the loss):
the single-view network (rotation matrices, betas, weak-perspective camera), SMPL-X with identity
global orientation, rotation about the origin by the predicted root rotation (``transform_smpl`` with
zero translation), camera translation ``[cam_y, cam_z... ]`` = ``[s1, s2, 2 f / (img_res * s0 + 1e-9)]``
(:145-147) and the perspective projection with that translation and the principal point at the origin
(:149-153).  SMPL-X, the rigid transform and the projection run fused in one native call.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import model_hmr
from .smplx import SMPLX

FOCAL_LENGTH = [1475, 1475]          # copenet/src/copenet/constants.py:7
IMG_RES = 224                        # hmr.py:196 (--img_res default)


class hmr(nn.Module):
    """``hparams`` needs ``smpl_mean_params`` (npz path), ``smplx_model_dir``, ``batch_size``; optional
    ``reg_iters`` (3), ``img_res`` (224), ``pretrained``."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.model = model_hmr.getcopenet(hparams.smpl_mean_params, pretrained=getattr(hparams, "pretrained", False))
        self.smplx = SMPLX(hparams.smplx_model_dir, batch_size=hparams.batch_size, create_transl=False)
        self.focal_length = FOCAL_LENGTH

    def forward(self, **kwargs):
        return self.model(**kwargs)

    @torch.no_grad()
    def fwd_pass(self, input_batch):
        im = input_batch["im0"].float()
        B = im.shape[0]
        pred_rotmat, pred_betas, pred_camera = self.model.forward(x=im, iters=getattr(self.hparams, "reg_iters", 3))   # :135-136
        img_res = float(getattr(self.hparams, "img_res", IMG_RES))
        pred_cam_t = torch.stack([pred_camera[:, 1], pred_camera[:, 2],
                                  2 * self.focal_length[0] / (img_res * pred_camera[:, 0] + 1e-9)], dim=-1)            # :145-147
        mo, cam = self.smplx.forward_camera(
            betas=pred_betas, body_pose=pred_rotmat[:, 1:], global_orient=None, transl=None, pose2rot=False,           # :139-143
            root_R=pred_rotmat[:, 0], root_t=None,                                                                    # transform_smpl with t = 0
            focal_length=self.focal_length, camera_center=None, proj_translation=pred_cam_t)                          # :149-153
        return {"pred_rotmat": pred_rotmat, "pred_betas": pred_betas, "pred_camera": pred_camera, "pred_cam_t": pred_cam_t,
                "pred_output_cam": mo, "pred_vertices": cam["vertices_cam"], "pred_joints": cam["joints_cam"],
                "pred_joints_2d_cam": cam["joints_2d"]}
