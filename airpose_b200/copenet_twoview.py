"""Host-side mirror of the reference LightningModule's per-frame forward path.

``copenet_twoview.fwd_pass`` follows copenet/src/copenet/copenet_twoview.py:164-317
(``fwd_pass_and_loss`` up to the loss): initial translation [0,0,10]*0.05, the two-view
network, translation un-scaling, 6D -> rotation matrices, SMPL-X with identity global
orientation, rigid transform about the origin, perspective projection.  SMPL-X,
``transform_smpl`` and ``perspective_projection`` run fused in one native call per view.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib, model_copenet
from .smplx import SMPLX, rot6d_to_rotmat, rot6d_to_rotmat_backward, smplx_backward

FOCAL_LENGTH = [1475, 1475]          # copenet/src/copenet/constants.py:7
TRANS_SCALE = 0.05                   # copenet_twoview.py:199-203
# default loss weights, copenet_twoview.py:655-677
LOSS_WEIGHTS = {"shape_loss_weight": 50.0, "keypoint2d_loss_weight": 0.002, "keypoint3d_loss_weight": 1.0,
                "limbs3d_loss_weight": 3.0, "limbstheta_loss_weight": 1.0, "trans_loss_weight": 10.0,
                "rootrot_loss_weight": 1.0, "pose_loss_weight": 50.0, "beta_loss_weight": 1.0}
# keys of the output dict of fwd_pass_and_loss(is_test=True) (copenet_twoview.py:328-350)
TEST_OUTPUT_KEYS = tuple(sorted(
    [k + v for v in "01" for k in ("pred_vertices_cam", "pred_vertices_cam_in", "pred_j2d_cam", "pred_j3d_cam", "pred_smpltrans", "pred_angles",
                                   "pred_betas", "in_smpltrans", "gt_angles", "gt_smpltrans", "smplorient_rel")] + ["smplpose_rotmat"]))
LOSS_NAMES = ("loss", "loss_regr_trans", "loss_keypoints", "loss_keypoints_3d", "loss_regr_shape", "loss_rootrot",
              "loss_regr_pose", "loss_regul_betas")           # order of the reference's `losses` dict (:152-159)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """``tgm.rotation_matrix_to_angle_axis`` as the reference calls it (copenet_twoview.py:323-326): [N,3,4] (or [N,3,3]) -> [N,3].
    torchgeometry 0.1.2 is absent offline: the arithmetic is restated in csrc/testmode.cu, PARITY UNPINNED."""
    if rotation_matrix.device.type != "cuda":
        raise _lib.AirposeError("airpose_b200.rotation_matrix_to_angle_axis runs on CUDA only; there is no CPU path")
    if rotation_matrix.dim() != 3 or rotation_matrix.shape[1] != 3 or rotation_matrix.shape[2] not in (3, 4):
        raise ValueError("Input size must be a N x 3 x 4 (or N x 3 x 3) tensor. Got {}".format(tuple(rotation_matrix.shape)))
    R = rotation_matrix.detach().to(torch.float32).contiguous()
    n, row = R.shape[0], R.shape[2]
    out = torch.empty(n, 3, device=R.device, dtype=torch.float32)
    if n:
        with torch.cuda.device(R.device):
            _lib.check(_lib.load().airpose_rotmat_to_angle_axis(R.data_ptr(), n, 3 * row, row, out.data_ptr(), _lib.current_stream()),
                       "airpose_rotmat_to_angle_axis")
    return out


def angle_axis_to_rotation_matrix(angle_axis):
    """``tgm.angle_axis_to_rotation_matrix`` (copenet_twoview.py:558-559): [N,3] -> [N,4,4] homogeneous, like torchgeometry returns it
    (the 3x3 block comes from the kernel).  PARITY UNPINNED, as above."""
    if angle_axis.device.type != "cuda":
        raise _lib.AirposeError("airpose_b200.angle_axis_to_rotation_matrix runs on CUDA only; there is no CPU path")
    aa = angle_axis.detach().to(torch.float32).reshape(-1, 3).contiguous()
    n = aa.shape[0]
    R = torch.empty(n, 3, 3, device=aa.device, dtype=torch.float32)
    if n:
        with torch.cuda.device(aa.device):
            _lib.check(_lib.load().airpose_angle_axis_to_rotmat(aa.data_ptr(), n, R.data_ptr(), _lib.current_stream()),
                       "airpose_angle_axis_to_rotmat")
    out = torch.eye(4, device=aa.device, dtype=torch.float32).repeat(n, 1, 1)
    out[:, :3, :3] = R
    return out


def mean_distance(a, b, points_used=None):
    """mean over items and their first ``points_used`` points of ||a - b||: [N,P,3] x [N,P,3] -> 0-d device tensor.  MPJPE with
    ``points_used=22`` (copenet_twoview.py:583-586), the mean position error for [N,3] inputs (:541-551)."""
    if a.device.type != "cuda":
        raise _lib.AirposeError("airpose_b200.mean_distance runs on CUDA only; there is no CPU path")
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(device=a.device, dtype=torch.float32).contiguous()
    if a.shape != b.shape or a.shape[-1] != 3:
        raise ValueError("mean_distance expects two [N,P,3] (or [N,3]) tensors of the same shape")
    if a.dim() == 2:
        a, b = a[:, None], b[:, None]
    n, per = a.shape[0], a.shape[1]
    out = torch.zeros(1, device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().airpose_mean_distance(a.data_ptr(), b.data_ptr(), n, per, int(points_used or per), out.data_ptr(),
                                                     _lib.current_stream()), "airpose_mean_distance")
    return out[0]


def dist_initialized():
    return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1


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

    def _focal(self, view):
        """Focal length pair of camera ``view`` (one constant for both in copenet, copenet_twoview.py:30-31; per camera in
        copenet_real, see airpose_b200/copenet_real.py)."""
        return self.focal_length

    def forward(self, **kwargs):
        return self.model(**kwargs)

    @torch.no_grad()
    def fwd_pass(self, input_batch, profile=None, is_test=False):
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
        in_trans, in_trans_unscaled = self._init_translation(B, im0.device, input_batch, is_test)
        reg_iters = getattr(self.hparams, "reg_iters", 3)
        mark("trunk")
        xf = self.model.forward_feat_ext_pair(im0, im1)                         # both views, one call (eval-mode BN)
        mark("trunk")
        mark("ief")
        pred = self.model._ief(xf[:B], xf[B:], bb0, bb1, in_trans[0], in_trans[1], None, None, None, None, reg_iters)
        mark("ief")
        mark("smplx")
        out = self._after_regressor(pred, intr, in_trans_unscaled)
        mark("smplx")
        return out

    def _init_translation(self, B, dev, input_batch=None, is_test=False):
        """The regressor's initial translation per view (copenet_twoview.py:178-203), scaled by 0.05 and unscaled:
        ``((scaled0, scaled1), (unscaled0, unscaled1))``.  Three branches, as in the reference:
          * ``is_test`` on ``testdata == "aircapdata"``: the ground-truth ``smpltrans_rel*`` (:178-180);
          * ``hparams.smpltrans_noise_sigma`` given: ground truth + sigma * N(0, 1), drawn the way ``add_noise_input_smpltrans``
            draws it (utils/utils.py:273-279: two CPU ``torch.randn(B, 3)`` per view, the first one used), so that a seeded
            run consumes the global generator identically;
          * otherwise the constant [0, 0, 10] -- cached per (B, device): building it from a Python list is a synchronous
            host->device copy, i.e. a host sync in every step."""
        sigma = getattr(self.hparams, "smpltrans_noise_sigma", None)
        aircap = is_test and str(getattr(self.hparams, "testdata", "")).lower() == "aircapdata"
        if aircap or sigma is not None:
            if input_batch is None or "smpltrans_rel0" not in input_batch or "smpltrans_rel1" not in input_batch:
                raise KeyError("this configuration initialises the translation from the ground truth: the batch needs "
                               "'smpltrans_rel0' / 'smpltrans_rel1' (copenet_twoview.py:169-170)")
            uns = []
            for v in (0, 1):
                gt = input_batch["smpltrans_rel%d" % v].to(device=dev, dtype=torch.float32)
                if not aircap:
                    noise = torch.randn(B, 3)
                    torch.randn(B, 3)                     # the reference's second, unused draw
                    gt = gt + float(sigma) * noise.to(dev)
                uns.append(gt.contiguous())
            return (uns[0] * TRANS_SCALE, uns[1] * TRANS_SCALE), (uns[0], uns[1])
        key = (B, str(dev))
        cache = self.__dict__.setdefault("_init_cache", {})
        if key not in cache:
            it = torch.tensor([0.0, 0.0, 10.0], dtype=torch.float32).expand(B, -1).clone()
            sc, un = (it * TRANS_SCALE).to(dev), it.to(dev)
            cache[key] = ((sc, sc), (un, un))
        return cache[key]

    def _after_regressor(self, pred, intr, in_trans_unscaled):
        """copenet_twoview.py:214-317 from the regressor outputs on: translation un-scaling, 6D -> rotation matrices,
        SMPL-X, transform_smpl, perspective_projection (fused into one native call).  When the two views' predictions are the
        halves of one buffer (``copenet._ief`` allocates them that way) and share the focal length, BOTH views go through ONE
        set of launches as a batch of 2B meshes; otherwise one call per view."""
        B = pred[0].shape[0]
        out = {}
        pose2, betas2 = pred[0]._base, pred[1]._base
        batched = (B > 0 and pose2 is not None and betas2 is not None and pose2.shape == (2 * B, 135) and betas2.shape == (2 * B, 10)
                   and pred[2]._base is pose2 and pred[3]._base is betas2 and pred[0].data_ptr() == pose2.data_ptr()
                   and pred[2].data_ptr() == pose2[B:].data_ptr() and tuple(self._focal(0)) == tuple(self._focal(1)))
        if batched:
            pose2[:, :3] /= TRANS_SCALE                                   # in-place, like :214-218
            rotmat2 = rot6d_to_rotmat(pose2[:, 3:]).view(2 * B, 22, 3, 3)   # :222-223
            centers = torch.cat([intr[0][:, :2, 2], intr[1][:, :2, 2]])
            mo2, cam2 = self.smplx.forward_camera(betas=betas2, body_pose=rotmat2[:, 1:], global_orient=None, transl=None,
                                                  pose2rot=False, root_R=rotmat2[:, 0], root_t=pose2[:, :3],
                                                  focal_length=self._focal(0), camera_center=centers)
        for v in (0, 1):
            pose, betas = pred[2 * v], pred[2 * v + 1]
            if batched:
                sl = slice(v * B, (v + 1) * B)
                rotmat = rotmat2[sl]
                mo = type(mo2)(**{k: (val[sl] if torch.is_tensor(val) and val.shape[:1] == (2 * B,) else val)
                                  for k, val in mo2._asdict().items()})
                cam = {k: val[sl] for k, val in cam2.items()}
            else:
                pose[:, :3] /= TRANS_SCALE                                   # in-place on the view, like :214-218
                rotmat = rot6d_to_rotmat(pose[:, 3:]).view(B, 22, 3, 3)      # :222-223
                mo, cam = self.smplx.forward_camera(
                    betas=betas, body_pose=rotmat[:, 1:],
                    global_orient=None,                                       # identity (:283)
                    transl=None, pose2rot=False,                              # the reference passes zeros (:284); None skips the add
                    root_R=rotmat[:, 0], root_t=pose[:, :3],                  # transform_smpl (:287-292)
                    focal_length=self._focal(v), camera_center=intr[v][:, :2, 2])   # :307-317
            trans = pose[:, :3]
            out.update({"pred_pose%d" % v: pose, "pred_betas%d" % v: betas, "pred_rotmat%d" % v: rotmat,
                        "pred_smpltrans%d" % v: trans, "in_smpltrans%d" % v: in_trans_unscaled[v],
                        "pred_output_cam%d" % v: mo,
                        "pred_vertices_cam%d" % v: cam["vertices_cam"], "pred_joints_cam%d" % v: cam["joints_cam"],
                        "pred_joints_2d_cam%d" % v: cam["joints_2d"]})
        return out

    def _hp(self, name):
        return float(getattr(self.hparams, name, LOSS_WEIGHTS[name]))

    @torch.no_grad()
    def get_loss(self, input_batch, pred_smpltrans0, pred_smpltrans1, pred_rotmat0, pred_rotmat1, pred_betas0,
                 pred_betas1, pred_output_cam0, pred_output_cam1, pred_joints_2d_cam0, pred_joints_2d_cam1,
                 with_grads=False):
        """copenet_twoview.get_loss (copenet_twoview.py:83-161), same argument list.  Returns
        ``(loss, losses)`` where ``loss`` is a 0-d device tensor and ``losses`` a dict of 0-d device
        tensors (views of ONE 8-float buffer: read them with a single ``.cpu()`` instead of the
        reference's eight ``.item()`` syncs).  ``with_grads=True`` also returns, as a third value, the
        gradient of ``loss`` with respect to every prediction (the backward of this function)."""
        dev = pred_betas0.device
        if dev.type != "cuda":
            raise _lib.AirposeError("get_loss runs on CUDA only; there is no CPU path")
        lib = _lib.load()
        def f(t):                        # float32, contiguous, 16-byte aligned (the kernel reads the vertex tensors as float4;
            t = t.float().contiguous()   # a view's half of a two-view [2B, V, 3] buffer is misaligned when B % 4 != 0)
            return t if t.data_ptr() % 16 == 0 else t.clone()
        B = pred_betas0.shape[0]
        v0, v1 = f(pred_output_cam0.vertices), f(pred_output_cam1.vertices)
        j0, j1 = f(pred_output_cam0.joints), f(pred_output_cam1.joints)
        keep = [v0, v1, j0, j1]
        a = _lib.LossArgs()
        a.batch, a.num_verts, a.num_joints = B, v0.shape[1], j0.shape[1]
        if pred_smpltrans0.stride(-1) != 1 or pred_smpltrans0.stride(0) != pred_smpltrans1.stride(0):
            pred_smpltrans0, pred_smpltrans1 = f(pred_smpltrans0), f(pred_smpltrans1)
        keep += [pred_smpltrans0, pred_smpltrans1]
        a.trans0, a.trans1, a.trans_stride = pred_smpltrans0.data_ptr(), pred_smpltrans1.data_ptr(), pred_smpltrans0.stride(0)
        names = {"rotmat0": pred_rotmat0, "rotmat1": pred_rotmat1, "betas0": pred_betas0, "betas1": pred_betas1,
                 "j2d0": pred_joints_2d_cam0, "j2d1": pred_joints_2d_cam1,
                 "gt_pose_rotmat": input_batch["smplpose_rotmat"], "gt_trans0": input_batch["smpltrans_rel0"],
                 "gt_trans1": input_batch["smpltrans_rel1"], "gt_orient0": input_batch["smplorient_rel0"],
                 "gt_orient1": input_batch["smplorient_rel1"], "gt_verts": input_batch["smpl_vertices"],
                 "gt_joints": input_batch["smpl_joints"], "gt_j2d0": input_batch["smpl_joints_2d0"],
                 "gt_j2d1": input_batch["smpl_joints_2d1"]}
        V, J = v0.shape[1], j0.shape[1]
        # the kernel indexes every tensor with the PREDICTION's vertex / joint counts: check the layouts before the launch
        expect = {"rotmat0": B * 22 * 9, "rotmat1": B * 22 * 9, "betas0": B * 10, "betas1": B * 10, "j2d0": B * J * 2, "j2d1": B * J * 2,
                  "gt_pose_rotmat": B * 21 * 9, "gt_trans0": B * 3, "gt_trans1": B * 3, "gt_orient0": B * 9, "gt_orient1": B * 9,
                  "gt_verts": B * V * 3, "gt_joints": B * J * 3, "gt_j2d0": B * J * 2, "gt_j2d1": B * J * 2}
        for n, t in names.items():
            if t.numel() != expect[n]:
                raise ValueError("get_loss: '{}' has shape {} ({} elements); the predictions have B={}, {} vertices, {} joints, "
                                 "which needs {} elements".format(n, tuple(t.shape), t.numel(), B, V, J, expect[n]))
            t = f(t.to(dev))
            keep.append(t)
            setattr(a, n, t.data_ptr())
        a.verts0, a.verts1, a.joints0, a.joints1 = v0.data_ptr(), v1.data_ptr(), j0.data_ptr(), j1.data_ptr()
        a.w_shape, a.w_kp2d, a.w_kp3d = self._hp("shape_loss_weight"), self._hp("keypoint2d_loss_weight"), self._hp("keypoint3d_loss_weight")
        a.w_limbs3d, a.w_limbstheta, a.w_trans = self._hp("limbs3d_loss_weight"), self._hp("limbstheta_loss_weight"), self._hp("trans_loss_weight")
        a.w_rootrot, a.w_pose, a.w_beta = self._hp("rootrot_loss_weight"), self._hp("pose_loss_weight"), self._hp("beta_loss_weight")
        out = torch.empty(8, device=dev, dtype=torch.float32)
        a.out = out.data_ptr()
        grads = None
        if with_grads:
            grads = {"vertices0": torch.empty_like(v0), "vertices1": torch.empty_like(v1), "joints0": torch.empty_like(j0),
                     "joints1": torch.empty_like(j1), "joints_2d0": torch.empty(B, j0.shape[1], 2, device=dev, dtype=torch.float32),
                     "joints_2d1": torch.empty(B, j0.shape[1], 2, device=dev, dtype=torch.float32), "rotmat0": torch.empty(B, 22, 3, 3, device=dev, dtype=torch.float32),
                     "rotmat1": torch.empty(B, 22, 3, 3, device=dev, dtype=torch.float32), "betas0": torch.empty(B, 10, device=dev, dtype=torch.float32),
                     "betas1": torch.empty(B, 10, device=dev, dtype=torch.float32), "smpltrans0": torch.empty(B, 3, device=dev, dtype=torch.float32),
                     "smpltrans1": torch.empty(B, 3, device=dev, dtype=torch.float32)}
            for field, key in (("g_verts0", "vertices0"), ("g_verts1", "vertices1"), ("g_joints0", "joints0"),
                               ("g_joints1", "joints1"), ("g_j2d0", "joints_2d0"), ("g_j2d1", "joints_2d1"),
                               ("g_rotmat0", "rotmat0"), ("g_rotmat1", "rotmat1"), ("g_betas0", "betas0"),
                               ("g_betas1", "betas1"), ("g_trans0", "smpltrans0"), ("g_trans1", "smpltrans1")):
                setattr(a, field, grads[key].data_ptr())
        with torch.cuda.device(dev):
            _lib.check(lib.airpose_twoview_loss(C.byref(a), _lib.current_stream()), "airpose_twoview_loss")
        del keep
        losses = {n: out[i] for i, n in enumerate(LOSS_NAMES)}
        return (out[0], losses, grads) if with_grads else (out[0], losses)

    @torch.no_grad()
    def fwd_pass_and_loss(self, input_batch, is_test=False, is_val=False):
        """copenet_twoview.fwd_pass_and_loss (copenet_twoview.py:164-374): forward + get_loss + the
        reference's output dict.  (The backward half: ``training_step`` hand-scheduled, or autograd through
        ``copenet.forward`` in train() mode + ``SMPLX.forward``, see INTEGRATION.md.)"""
        out = self.fwd_pass(input_batch, is_test=is_test)
        if is_test:
            return self._test_outputs(input_batch, out), None, None
        else:
            loss, losses = self.get_loss(input_batch, out["pred_smpltrans0"], out["pred_smpltrans1"], out["pred_rotmat0"],
                                         out["pred_rotmat1"], out["pred_betas0"], out["pred_betas1"],
                                         out["pred_output_cam0"], out["pred_output_cam1"],
                                         out["pred_joints_2d_cam0"], out["pred_joints_2d_cam1"])
        output = {"pred_vertices_cam0": out["pred_vertices_cam0"], "pred_vertices_cam1": out["pred_vertices_cam1"],   # :362-372
                  "pred_smpltrans0": out["pred_smpltrans0"], "pred_smpltrans1": out["pred_smpltrans1"],
                  "in_smpltrans0": out["in_smpltrans0"], "in_smpltrans1": out["in_smpltrans1"]}
        return output, losses, loss

    @torch.no_grad()
    def _test_outputs(self, input_batch, out):
        """The ``is_test`` branch of fwd_pass_and_loss (copenet_twoview.py:258-279,318-350): the extra SMPL-X passes with zero
        betas placed at the INPUT translation with identity orientation, angle-axis forms of the predicted and ground-truth
        rotations, and the reference's output dict (every entry ``.detach().cpu()`` like the reference's)."""
        B = out["pred_pose0"].shape[0]
        dev = out["pred_pose0"].device
        cpu = lambda t_: t_.detach().cpu()
        pad = torch.zeros(B, 22, 3, 1, device=dev)
        aa = lambda R: rotation_matrix_to_angle_axis(torch.cat([R, pad], dim=3).view(-1, 3, 4)).view(B, 22, 3)       # :323-326
        res = {}
        for v in (0, 1):
            rot = out["pred_rotmat%d" % v]
            # :258-279 zero betas, predicted body pose, identity global orientation, then [I | in_smpltrans]
            _, cam = self.smplx.forward_camera(betas=torch.zeros(B, 10, device=dev), body_pose=rot[:, 1:], global_orient=None, transl=None,
                                               pose2rot=False, root_R=torch.eye(3, device=dev).expand(B, 3, 3).contiguous(),
                                               root_t=out["in_smpltrans%d" % v])
            gt_rot = torch.cat([input_batch["smplorient_rel%d" % v], input_batch["smplpose_rotmat"]], dim=1).to(dev).float()
            res.update({"pred_vertices_cam%d" % v: cpu(out["pred_vertices_cam%d" % v]), "pred_vertices_cam_in%d" % v: cpu(cam["vertices_cam"]),
                        "pred_j2d_cam%d" % v: cpu(out["pred_joints_2d_cam%d" % v]), "pred_j3d_cam%d" % v: cpu(out["pred_joints_cam%d" % v]),
                        "pred_smpltrans%d" % v: cpu(out["pred_smpltrans%d" % v]), "pred_angles%d" % v: cpu(aa(rot)),
                        "pred_betas%d" % v: cpu(out["pred_betas%d" % v]), "in_smpltrans%d" % v: cpu(out["in_smpltrans%d" % v]),
                        "gt_angles%d" % v: cpu(aa(gt_rot)), "gt_smpltrans%d" % v: cpu(input_batch["smpltrans_rel%d" % v]),
                        "smplorient_rel%d" % v: cpu(input_batch["smplorient_rel%d" % v])})
        res["smplpose_rotmat"] = cpu(input_batch["smplpose_rotmat"])
        return res

    def test_step(self, batch, batch_idx=0, dset_idx=0):
        """copenet_twoview.test_step (:537-546)."""
        output, losses, loss = self.fwd_pass_and_loss(batch, is_val=True, is_test=True)
        return {"test_loss": loss, "output": output}

    @torch.no_grad()
    def test_metrics(self, outputs):
        """The numbers ``test_epoch_end`` prints (copenet_twoview.py:548-586) for ONE list of ``test_step`` results: the mean
        position errors of the two views and the MPJPE over the first 22 joints between the SMPL-X meshes (zero betas -- the
        module's own parameter, :575-582) of the ground-truth and of the predicted rotations, the latter taken through
        angle-axis and back exactly as the reference does (:556-559).  Reductions run on the device; returns a dict of floats."""
        dev = self.smplx.v_template.device
        cat = lambda k: torch.cat([o["output"][k].to(dev) for o in outputs])
        res = {"mpe0": float(mean_distance(cat("pred_smpltrans0"), cat("gt_smpltrans0"))),
               "mpe1": float(mean_distance(cat("pred_smpltrans1"), cat("gt_smpltrans1")))}
        body_gt = cat("smplpose_rotmat")
        N = body_gt.shape[0]
        for v in (0, 1):
            pred = angle_axis_to_rotation_matrix(cat("pred_angles%d" % v).view(-1, 3)).view(N, 22, 4, 4)[:, :, :3, :3].contiguous()
            j_gt = self.smplx.forward(body_pose=body_gt, global_orient=cat("smplorient_rel%d" % v), pose2rot=False).joints
            j_pr = self.smplx.forward(body_pose=pred[:, 1:22].contiguous(), global_orient=pred[:, 0:1].contiguous(), pose2rot=False).joints
            res["mpjpe%d" % v] = float(mean_distance(j_gt, j_pr, points_used=22))
        return res

    @torch.no_grad()
    def loss_and_head_backward(self, input_batch, out):
        """get_loss and the backward pass from the loss down to the regressor outputs: d loss / d pred_pose{0,1}
        [B,135] (as the network returned them, i.e. before the translation un-scaling of :214-218) and
        d loss / d pred_betas{0,1} [B,10] -- the part of ``loss.backward()`` (copenet_twoview.py:378-386) that runs
        through get_loss, perspective_projection, transform_smpl, SMPL-X and rot6d_to_rotmat.  ``out`` is the dict
        ``fwd_pass`` returned.  ``training_step`` continues from here through the regressor and the trunk."""
        loss, losses, g = self.get_loss(input_batch, out["pred_smpltrans0"], out["pred_smpltrans1"], out["pred_rotmat0"],
                                        out["pred_rotmat1"], out["pred_betas0"], out["pred_betas1"], out["pred_output_cam0"],
                                        out["pred_output_cam1"], out["pred_joints_2d_cam0"], out["pred_joints_2d_cam1"],
                                        with_grads=True)
        grads = {}
        for v in (0, 1):
            rotmat, pose, betas = out["pred_rotmat%d" % v], out["pred_pose%d" % v], out["pred_betas%d" % v]
            sg = smplx_backward(self.smplx, betas, rotmat[:, 1:], None, grad_vertices=g["vertices%d" % v],
                                grad_joints=g["joints%d" % v], grad_joints_2d=g["joints_2d%d" % v],
                                joints=out["pred_output_cam%d" % v].joints, root_R=rotmat[:, 0], root_t=pose[:, :3],
                                focal_length=self._focal(v))
            g_rot = g["rotmat%d" % v].clone()
            g_rot[:, 1:] += sg["body_pose"]
            g_rot[:, 0] += sg["root_R"]
            g_pose = torch.empty_like(pose)
            g_pose[:, 3:] = rot6d_to_rotmat_backward(pose[:, 3:], g_rot.view(-1, 3, 3))
            g_pose[:, :3] = (g["smpltrans%d" % v] + sg["root_t"]) / TRANS_SCALE          # pose[:, :3] /= trans_scale (:214-218)
            grads["pred_pose%d" % v] = g_pose
            grads["pred_betas%d" % v] = g["betas%d" % v] + sg["betas"]
        return loss, losses, grads

    # ------------------------------------------------------------------ regressor-only training step
    def configure_optimizers_reg_only(self):
        """The optimizer of copenet_twoview.py:416-425 (Adam, amsgrad) over the parameters the reference's
        ``train_reg_only`` switch leaves trainable (copenet_real/.../copenet_twoview.py:357-372): fc1, fc2, decpose,
        decshape (deccam has no gradient in the two-view model and is left out)."""
        from .optim import Adam
        from .parallel import broadcast_buffers_
        params = dict(self.model.named_parameters())
        if dist_initialized():                    # the frozen trunk must be identical across ranks too (DDP broadcasts the whole module)
            for p in self.model.parameters():
                torch.distributed.broadcast(p.data, src=0)
        broadcast_buffers_(self.model)
        return Adam([params[n] for n in self.model.REG_PARAMS], lr=float(getattr(self.hparams, "lr", 5e-5)), amsgrad=True)

    @torch.no_grad()
    def training_step_reg_only(self, input_batch, optimizer, mask1=None, mask2=None):
        """One data-parallel training step of the regressor with the trunk frozen: the reference's ``train_reg_only``
        fine-tuning mode with copenet's loss.  Forward: trunk features (frozen; with the module in ``train()`` mode, as
        the reference leaves it, BatchNorm normalises with batch statistics and updates its running statistics, one
        trunk call per view; in ``eval()`` mode the folded-BN pair kernel path runs), regressor with dropout, SMPL-X, projection, get_loss.  Backward:
        loss -> SMPL-X -> rot6d -> regressor, parameter gradients written straight into the optimizer's flat
        gradient buffer.  Then ONE all-reduce of that buffer over the data-parallel ranks (NCCL) and one Adam launch.
        Returns ``(loss, losses)`` as device tensors (no host sync)."""
        im0, im1 = input_batch["im0"].float(), input_batch["im1"].float()
        B = im0.shape[0]
        in_trans, in_trans_unscaled = self._init_translation(B, im0.device, input_batch)
        reg_iters = getattr(self.hparams, "reg_iters", 3)
        if self.model.training:       # train() mode, like the reference: batch-statistics BatchNorm, one call per view (:140-141)
            xf = torch.cat([self.model.forward_feat_ext(im0), self.model.forward_feat_ext(im1)])
        else:
            xf = self.model.forward_feat_ext_pair(im0, im1)
        pred, ctx = self.model.ief_train_forward(xf[:B], xf[B:], input_batch["bb0"], input_batch["bb1"], in_trans[0], in_trans[1],
                                                 iters=reg_iters, mask1=mask1, mask2=mask2)
        out = self._after_regressor(pred, (input_batch["intr0"], input_batch["intr1"]), in_trans_unscaled)
        loss, losses, g = self.loss_and_head_backward(input_batch, out)
        optimizer.zero_grad()
        self.model.ief_train_backward(ctx, g["pred_pose0"], g["pred_betas0"], g["pred_pose1"], g["pred_betas1"],
                                      into_param_grads=True)
        scale = optimizer.allreduce_grads()
        optimizer.step(grad_scale=scale)
        return loss, losses

    # ------------------------------------------------------------------ full training step
    def configure_optimizers(self):
        """copenet_twoview.configure_optimizers (copenet_twoview.py:416-425): Adam(lr, weight_decay=0, amsgrad=True) over
        ``self.model.parameters()``, here as one launch over one flat buffer (``airpose_b200.optim.Adam``)."""
        from .optim import Adam
        from .parallel import broadcast_buffers_
        broadcast_buffers_(self.model)            # with Adam's parameter broadcast: every rank starts from rank 0's state, like DDP
        return Adam(self.model.parameters(), lr=float(getattr(self.hparams, "lr", 5e-5)), amsgrad=True)

    @torch.no_grad()
    def training_step(self, input_batch, optimizer, mask1=None, mask2=None):
        """One data-parallel training step of the whole network (copenet_twoview.py:376-386 + Lightning's backward /
        optimizer step / DDP all-reduce), hand-scheduled instead of autograd-driven:

        forward   trunk in train() mode, one call per view, activations kept on a tape (batch-statistics BatchNorm,
                  running statistics updated); regressor with dropout; SMPL-X + transform + projection; get_loss
        backward  loss -> SMPL-X -> rot6d -> regressor (parameter gradients + d loss / d features) -> trunk, view 0 then
                  view 1 accumulating, every gradient written straight into the optimizer's flat gradient buffer
        update    the flat gradient buffer all-reduced over the ranks (NCCL over NVLink) in two parts -- layer3 + layer4 + regressor
                  under the rest of the backward, the remainder at its end --, ONE Adam(amsgrad) launch.
        ``deccam`` takes part with a zero gradient (it is unused by the two-view model, model_copenet.py:73).
        Returns ``(loss, losses)`` as device tensors (no host sync)."""
        if not self.model.training:
            raise RuntimeError("training_step needs the module in train() mode (batch-statistics BatchNorm, dropout)")
        im0, im1 = input_batch["im0"].float(), input_batch["im1"].float()
        B = im0.shape[0]
        in_trans, in_trans_unscaled = self._init_translation(B, im0.device, input_batch)
        reg_iters = getattr(self.hparams, "reg_iters", 3)
        im0, im1 = im0.contiguous(), im1.contiguous()
        paired = 2 <= B and 2 * B <= self.model.PAIR_MAX_IMAGES    # both views through one set of launches (BatchNorm still per view)
        if paired:
            xf = self.model._forward_feat_ext_train_pair(im0, im1, tape=0)
            xf0, xf1 = xf[:B], xf[B:]
        else:
            xf0 = self.model._forward_feat_ext_train(im0, tape=0)
            xf1 = self.model._forward_feat_ext_train(im1, tape=1)
        pred, ctx = self.model.ief_train_forward(xf0, xf1, input_batch["bb0"], input_batch["bb1"], in_trans[0], in_trans[1],
                                                 iters=reg_iters, mask1=mask1, mask2=mask2)
        out = self._after_regressor(pred, (input_batch["intr0"], input_batch["intr1"]), in_trans_unscaled)
        loss, losses, g = self.loss_and_head_backward(input_batch, out)
        optimizer.zero_grad()
        gr = self.model.ief_train_backward(ctx, g["pred_pose0"], g["pred_betas0"], g["pred_pose1"], g["pred_betas1"],
                                           want_feature_grads=True, into_param_grads=True)
        # Data-parallel gradient mean, overlapped with the backward the way DDP's buckets are: the gradients of layer3, layer4 and
        # the regressor (everything behind layer2 in the flat buffer: 95 % of the bytes) are final when the trunk backward has
        # enqueued layer3.0 -- their all-reduce runs on NCCL's stream under the backward of layer2, layer1 and the stem; the
        # rest is reduced at the end.  AIRPOSE_NO_OVERLAP_ALLREDUCE=1: one all-reduce of the whole buffer after the backward.
        world = torch.distributed.get_world_size() if dist_initialized() else 1
        split, works = 0, []
        if world > 1 and not os.environ.get("AIRPOSE_NO_OVERLAP_ALLREDUCE"):
            m = self.model
            late = [p for mod in (m.conv1, m.bn1, m.layer1, m.layer2) for p in mod.parameters()]
            split = optimizer.late_split(late)
            if not 0 < split < optimizer.numel:
                split = 0
        hook = (lambda: works.append(optimizer.allreduce_begin(split))) if split else None
        if paired:
            self.model.backward_feat_ext(im0, 0, torch.cat([gr["xf0"], gr["xf1"]]), accumulate=False, into_param_grads=True, x1=im1,
                                         upper_done=hook)
        else:
            self.model.backward_feat_ext(im0, 0, gr["xf0"], accumulate=False, into_param_grads=True)
            self.model.backward_feat_ext(im1, 1, gr["xf1"], accumulate=True, into_param_grads=True, upper_done=hook)
        if split and works and works[0] is not None:
            works.append(optimizer.allreduce_begin(0, split))
            for w in works:
                if w is not None:
                    w.wait()
            scale = 1.0 / world
        else:
            scale = optimizer.allreduce_grads()
        optimizer.step(grad_scale=scale)
        return loss, losses
