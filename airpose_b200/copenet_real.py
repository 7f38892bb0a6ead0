"""The real-data fine-tuning variant of the two-view module:
``copenet_real/src/copenet_real/copenet_twoview.py`` -- same network and forward, per-camera focal lengths
(``copenet_real/constants.py:12-13``: FOCAL_LENGTH0 / FOCAL_LENGTH1, used at ``copenet_twoview.py:85-86,297-307``) and a loss that
needs no 3D ground truth (``get_loss``, :99-160): confidence-weighted 2D keypoints over the first 22 joints with limb weights,
cross-view consistency of the body rotations and of the betas, a beta regulariser, the exp(-t_z)^2 depth barrier, x60.

The VPoser prior term (:123-135) needs ``human_body_prior`` and its downloaded weights, which do not exist offline: pass
``vposer=callable`` (pose_body_aa [B,63] -> loss_regul_vposer, a 0-d tensor; evaluated in torch, no gradient flows through it in
the hand-scheduled steps) or leave it out (term = 0).  Everything else runs on the native kernels: ``airpose_real_loss``
(csrc/loss.cu: loss + d loss / d prediction in one launch) and, for the backward, the same ``airpose_smplx_bwd`` /
``rot6d`` / regressor / trunk chain as ``copenet_twoview.training_step`` -- this class only overrides the loss.
The AirPose+ bundle adjustment (``copenet_real_data/scripts/bundle_adj.py:301-401``) optimises VPoser latents and is NOT built.
No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .copenet_twoview import copenet_twoview as _copenet_twoview

FOCAL_LENGTH0 = (1537.0, 1517.0)        # copenet_real/constants.py:12
FOCAL_LENGTH1 = (1361.0, 1378.0)        # :13
REAL_LOSS_WEIGHTS = {"limbs2d_loss_weight": 3.0, "keypoint2d_loss_weight": 0.002, "beta_loss_weight": 1.0,
                     "vposer_loss_weight": 0.01, "pose_loss_weight": 50.0}
REAL_LOSS_KEYS = ("loss", "loss_regul_vposer", "loss_regr_pose", "loss_keypoints", "loss_regul_betas")


class copenet_twoview(_copenet_twoview):
    def __init__(self, hparams, vposer=None):
        super().__init__(hparams)
        self.focal_length0, self.focal_length1 = FOCAL_LENGTH0, FOCAL_LENGTH1      # :85-86
        self.vposer = vposer

    def _focal(self, view):
        return self.focal_length1 if view else self.focal_length0

    def _hp(self, name):
        return float(getattr(self.hparams, name, REAL_LOSS_WEIGHTS.get(name, 0.0)))

    @torch.no_grad()
    def get_loss(self, input_batch, pred_smpltrans0, pred_smpltrans1, pred_rotmat0, pred_rotmat1, pred_betas0, pred_betas1,
                 pred_output_cam0, pred_output_cam1, pred_joints_2d_cam0, pred_joints_2d_cam1, with_grads=False):
        """copenet_real's get_loss (:99-160), same argument list.  Returns ``(loss, losses)`` -- ``losses`` are 0-d views of ONE
        5-float device buffer in the reference's dict order -- and with ``with_grads=True`` a third value: d loss / d prediction
        for joints_2d, rotmat, betas and smpltrans of both views."""
        dev = pred_betas0.device
        if dev.type != "cuda":
            raise _lib.AirposeError("get_loss runs on CUDA only; there is no CPU path")
        lib = _lib.load()
        f = lambda t: t.detach().float().contiguous()
        B = pred_betas0.shape[0]
        g0, g1 = f(input_batch["smpl_joints_2d0"].to(dev)[:, 0]), f(input_batch["smpl_joints_2d1"].to(dev)[:, 0])
        if g0.dim() != 3 or g0.shape[0] != B or g0.shape[1] < 22 or g0.shape[2] != 3 or g1.shape != g0.shape:
            raise ValueError("get_loss: smpl_joints_2d{0,1} must be [B, 1, >=22, 3] = (x, y, confidence); got %s / %s" %
                             (tuple(input_batch["smpl_joints_2d0"].shape), tuple(input_batch["smpl_joints_2d1"].shape)))
        j0, j1 = f(pred_joints_2d_cam0), f(pred_joints_2d_cam1)
        r0, r1, b0, b1 = f(pred_rotmat0), f(pred_rotmat1), f(pred_betas0), f(pred_betas1)
        if pred_smpltrans0.stride(-1) != 1 or pred_smpltrans0.stride(0) != pred_smpltrans1.stride(0) or pred_smpltrans0.dtype != torch.float32:
            pred_smpltrans0, pred_smpltrans1 = f(pred_smpltrans0), f(pred_smpltrans1)
        vterm = 0.0
        if self.vposer is not None:          # :123-135, evaluated by the caller's model
            from .copenet_twoview import rotation_matrix_to_angle_axis
            aa = [rotation_matrix_to_angle_axis(r[:, 1:].reshape(-1, 3, 3)).reshape(B, 63) for r in (r0, r1)]
            vterm = float(self.vposer(aa[0]) + self.vposer(aa[1]))
        a = _lib.RealLossArgs()
        a.batch, a.num_joints, a.gt_joints = B, j0.shape[1], g0.shape[1]
        a.trans0, a.trans1, a.trans_stride = pred_smpltrans0.data_ptr(), pred_smpltrans1.data_ptr(), pred_smpltrans0.stride(0)
        a.rotmat0, a.rotmat1, a.betas0, a.betas1 = r0.data_ptr(), r1.data_ptr(), b0.data_ptr(), b1.data_ptr()
        a.j2d0, a.j2d1, a.gt_j2d0, a.gt_j2d1 = j0.data_ptr(), j1.data_ptr(), g0.data_ptr(), g1.data_ptr()
        a.w_kp2d, a.w_limbs2d, a.w_beta = self._hp("keypoint2d_loss_weight"), self._hp("limbs2d_loss_weight"), self._hp("beta_loss_weight")
        a.w_pose, a.w_vposer, a.vposer_term = self._hp("pose_loss_weight"), self._hp("vposer_loss_weight"), vterm
        out = torch.empty(5, device=dev, dtype=torch.float32)
        a.out = out.data_ptr()
        grads = None
        if with_grads:
            grads = {"joints_2d0": torch.empty_like(j0), "joints_2d1": torch.empty_like(j1), "rotmat0": torch.empty_like(r0),
                     "rotmat1": torch.empty_like(r1), "betas0": torch.empty_like(b0), "betas1": torch.empty_like(b1),
                     "smpltrans0": torch.empty(B, 3, device=dev, dtype=torch.float32),
                     "smpltrans1": torch.empty(B, 3, device=dev, dtype=torch.float32)}
            for field, key in (("g_j2d0", "joints_2d0"), ("g_j2d1", "joints_2d1"), ("g_rotmat0", "rotmat0"), ("g_rotmat1", "rotmat1"),
                               ("g_betas0", "betas0"), ("g_betas1", "betas1"), ("g_trans0", "smpltrans0"), ("g_trans1", "smpltrans1")):
                setattr(a, field, grads[key].data_ptr())
        with torch.cuda.device(dev):
            _lib.check(lib.airpose_real_loss(C.byref(a), _lib.current_stream()), "airpose_real_loss")
        losses = {k: out[i] for i, k in enumerate(REAL_LOSS_KEYS)}
        if with_grads:
            # the 3D quantities do not enter this loss: their upstream gradients are absent (None), the SMPL-X backward
            # (airpose_smplx_bwd) then runs on the 2D-joint gradient alone
            grads.update({"vertices0": None, "vertices1": None, "joints0": None, "joints1": None})
            return out[0], losses, grads
        return out[0], losses
