"""Drop-in ``SMPLX`` body model whose forward runs on the sm_100a kernels.

Mirrors the call surface of the reference's forked smplx
(copenet/src/copenet/smplx/smplx/body_models.py:648-994) that AirPose uses:

    SMPLX(model_path, batch_size=B, create_transl=False[, gender=...]).to(device)
    out = model.forward(betas=..., body_pose=[B,21,3,3], global_orient=[B,1,3,3],
                        transl=[B,3], pose2rot=False)          # copenet_twoview.py:281-285
    out = model.forward(body_pose=..., global_orient=..., pose2rot=False)   # :575-582
    out.vertices [B,10475,3], out.joints [B,127,3], out.betas, out.body_pose, out.full_pose

Same buffer / parameter names as the reference module, so ``state_dict`` and attribute
access (``.v_template``, ``.faces``, ``.faces_tensor``, ``.batch_size``) keep working.
Only the rotation-matrix path (``pose2rot=False``, the fork's addition at :878-935) is
implemented; the axis-angle path is used by the reference's tooling only.
Unlike the reference, the runtime batch does not have to equal ``batch_size``.
"""
from __future__ import annotations

import ctypes as C
import os
import os.path as osp
import pickle
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

ModelOutput = namedtuple("ModelOutput",
                         ["vertices", "joints", "full_pose", "betas", "global_orient", "body_pose", "expression",
                          "left_hand_pose", "right_hand_pose", "jaw_pose"])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)

# vertex_ids.py:47-69 ('smplx') in the order VertexJointSelector concatenates them
# (vertex_joint_selector.py:38-68): face, feet, left-hand tips, right-hand tips.
_VERTEX_IDS = {"nose": 9120, "reye": 9929, "leye": 9448, "rear": 616, "lear": 6,
               "LBigToe": 5770, "LSmallToe": 5780, "LHeel": 8846, "RBigToe": 8463, "RSmallToe": 8474, "RHeel": 8635,
               "lthumb": 5361, "lindex": 4933, "lmiddle": 5058, "lring": 5169, "lpinky": 5286,
               "rthumb": 8079, "rindex": 7669, "rmiddle": 7794, "rring": 7905, "rpinky": 8022}
EXTRA_JOINT_ORDER = ["nose", "reye", "leye", "rear", "lear", "LBigToe", "LSmallToe", "LHeel", "RBigToe", "RSmallToe",
                     "RHeel", "lthumb", "lindex", "lmiddle", "lring", "lpinky", "rthumb", "rindex", "rmiddle",
                     "rring", "rpinky"]
SMPL2OP_J14 = [15, 12, 17, 19, 21, 16, 18, 20, 2, 5, 8, 1, 4, 7]   # copenet_real_data/scripts/bundle_adj.py:48


def _to_np(a, dtype=np.float32):
    if "scipy.sparse" in str(type(a)):
        a = a.todense()
    return np.array(a, dtype=dtype)


class VertexJointSelector(nn.Module):
    """Holds ``extra_joints_idxs`` like the reference module (vertex_joint_selector.py:25-71);
    the gather itself happens inside the fused joints kernel."""

    def __init__(self):
        super().__init__()
        idx = np.array([_VERTEX_IDS[k] for k in EXTRA_JOINT_ORDER], dtype=np.int64)
        self.register_buffer("extra_joints_idxs", torch.from_numpy(idx))


class SMPLX(nn.Module):
    NUM_BODY_JOINTS = 21
    NUM_HAND_JOINTS = 15
    NUM_FACE_JOINTS = 3
    NUM_JOINTS = NUM_BODY_JOINTS + 2 * NUM_HAND_JOINTS + NUM_FACE_JOINTS
    NUM_BETAS = 10
    NUM_EXPR_COEFFS = 10

    def __init__(self, model_path, create_expression=True, expression=None, create_jaw_pose=True, jaw_pose=None,
                 create_leye_pose=True, leye_pose=None, create_reye_pose=True, reye_pose=None,
                 use_face_contour=False, batch_size=1, gender="neutral", dtype=torch.float32, ext="npz",
                 create_betas=True, betas=None, create_global_orient=True, global_orient=None,
                 create_body_pose=True, body_pose=None, create_transl=True, transl=None,
                 create_left_hand_pose=True, left_hand_pose=None, create_right_hand_pose=True,
                 right_hand_pose=None, use_pca=True, num_pca_comps=6, flat_hand_mean=False, joint_mapper=None,
                 **kwargs):
        super().__init__()
        if dtype != torch.float32:
            raise ValueError("airpose_b200.SMPLX computes in float32 only")
        if use_face_contour:
            raise NotImplementedError("use_face_contour (dynamic landmarks) is not on the AirPose hot path")
        if osp.isdir(model_path):
            path = os.path.join(model_path, "SMPLX_{}.{}".format(gender.upper(), ext))
        else:
            path = model_path
        assert osp.exists(path), "Path {} does not exist!".format(path)      # body_models.py:705-706
        if ext == "pkl":
            with open(path, "rb") as f:
                data = pickle.load(f, encoding="latin1")
        elif ext == "npz":
            data = np.load(path, allow_pickle=True)
        else:
            raise ValueError("Unknown extension: {}".format(ext))

        self.batch_size = batch_size
        self.dtype = dtype
        self.gender = gender
        self.joint_mapper = joint_mapper
        self.use_pca = use_pca
        self.num_pca_comps = num_pca_comps
        self.flat_hand_mean = flat_hand_mean
        self.use_face_contour = False
        self.vertex_joint_selector = VertexJointSelector()

        self.faces = data["f"]
        self.register_buffer("faces_tensor", torch.from_numpy(_to_np(self.faces, np.int64)))
        self.register_buffer("v_template", torch.from_numpy(_to_np(data["v_template"])))
        self.register_buffer("shapedirs", torch.from_numpy(_to_np(data["shapedirs"])))
        self.register_buffer("J_regressor", torch.from_numpy(_to_np(data["J_regressor"])))
        nb = data["posedirs"].shape[-1]
        self.register_buffer("posedirs", torch.from_numpy(np.reshape(_to_np(data["posedirs"]), [-1, nb]).T.copy()))
        parents = torch.from_numpy(_to_np(data["kintree_table"][0])).long()
        parents[0] = -1
        self.register_buffer("parents", parents)
        self.register_buffer("lbs_weights", torch.from_numpy(_to_np(data["weights"])))
        self.register_buffer("lmk_faces_idx", torch.from_numpy(_to_np(data["lmk_faces_idx"], np.int64)))
        self.register_buffer("lmk_bary_coords", torch.from_numpy(_to_np(data["lmk_bary_coords"])))
        lc = np.asarray(data["hands_componentsl"])[:num_pca_comps]
        rc = np.asarray(data["hands_componentsr"])[:num_pca_comps]
        if use_pca:
            self.register_buffer("left_hand_components", torch.tensor(lc, dtype=dtype))
            self.register_buffer("right_hand_components", torch.tensor(rc, dtype=dtype))
        lm = np.zeros_like(data["hands_meanl"]) if flat_hand_mean else data["hands_meanl"]
        rm = np.zeros_like(data["hands_meanr"]) if flat_hand_mean else data["hands_meanr"]
        self.register_buffer("left_hand_mean", torch.tensor(_to_np(lm)))
        self.register_buffer("right_hand_mean", torch.tensor(_to_np(rm)))
        self.register_buffer("pose_mean", torch.cat([torch.zeros(3 + 63 + 9, dtype=torch.float32), self.left_hand_mean, self.right_hand_mean]))

        def param(name, create, value, shape):
            if not create:
                return
            t = torch.zeros(shape, dtype=dtype) if value is None else torch.as_tensor(value, dtype=dtype).clone()
            self.register_parameter(name, nn.Parameter(t, requires_grad=True))

        hand_dim = num_pca_comps if use_pca else 3 * self.NUM_HAND_JOINTS
        param("betas", create_betas, betas, [batch_size, self.NUM_BETAS])
        param("global_orient", create_global_orient, global_orient, [batch_size, 3])
        param("body_pose", create_body_pose, body_pose, [batch_size, self.NUM_BODY_JOINTS * 3])
        param("left_hand_pose", create_left_hand_pose, left_hand_pose, [batch_size, hand_dim])
        param("right_hand_pose", create_right_hand_pose, right_hand_pose, [batch_size, hand_dim])
        param("transl", create_transl, transl, [batch_size, 3])
        param("jaw_pose", create_jaw_pose, jaw_pose, [batch_size, 3])
        param("leye_pose", create_leye_pose, leye_pose, [batch_size, 3])
        param("reye_pose", create_reye_pose, reye_pose, [batch_size, 3])
        param("expression", create_expression, expression, [batch_size, self.NUM_EXPR_COEFFS])

        self._handle = None
        self._handle_device = None
        self._zero_cache = {}

    # ------------------------------------------------------------------ native handle
    def _get_handle(self, device):
        if self._handle is not None and self._handle_device == device:
            return self._handle
        lib = _lib.load()
        self._release()
        host = {k: getattr(self, k).detach().cpu().contiguous() for k in
                ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights", "faces_tensor",
                 "lmk_faces_idx", "lmk_bary_coords")}
        extra = self.vertex_joint_selector.extra_joints_idxs.detach().cpu().contiguous()
        m = _lib.SmplxModelHost()
        m.num_verts = host["v_template"].shape[0]
        m.num_joints = host["J_regressor"].shape[0]
        m.num_shape = host["shapedirs"].shape[-1]
        m.num_pose_basis = host["posedirs"].shape[0]
        m.num_faces = host["faces_tensor"].shape[0]
        m.num_landmarks = host["lmk_faces_idx"].shape[0]
        m.num_extra = extra.shape[0]
        m.v_template = host["v_template"].data_ptr()
        m.shapedirs = host["shapedirs"].data_ptr()
        m.posedirs = host["posedirs"].data_ptr()
        m.J_regressor = host["J_regressor"].data_ptr()
        m.parents = host["parents"].data_ptr()
        m.lbs_weights = host["lbs_weights"].data_ptr()
        m.faces = host["faces_tensor"].data_ptr()
        m.lmk_faces_idx = host["lmk_faces_idx"].data_ptr()
        m.lmk_bary_coords = host["lmk_bary_coords"].data_ptr()
        m.extra_joint_idx = extra.data_ptr()
        h = C.c_void_p()
        _lib.check(lib.airpose_smplx_create(C.byref(h), C.byref(m), device.index or 0), "airpose_smplx_create")
        self._handle, self._handle_device = h, device
        return h

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().airpose_smplx_destroy(self._handle)
            except Exception:
                pass
            try:
                object.__setattr__(self, "_handle", None)      # also safe during interpreter shutdown
            except Exception:
                pass

    def __del__(self):
        self._release()

    def _is_zero(self, name):
        """True when the module's own parameter ``name`` is all zeros (cached per version):
        batch_rodrigues(0) is exactly the identity (lbs.py:284-299), so the kernel may skip it."""
        p = getattr(self, name, None)
        if p is None:
            return True
        key = (p.data_ptr(), p._version)
        hit = self._zero_cache.get(name)
        if hit is None or hit[0] != key:
            hit = (key, bool((p.detach() == 0).all().item()))
            self._zero_cache[name] = hit
        return hit[1]

    @staticmethod
    def _f32c(t, device):
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    @staticmethod
    def _rows(t, device, width):
        """[B, width] row view of ``t`` without a copy when each sample's block is contiguous (e.g. ``rotmat[:, 1:]``,
        ``pose[:, :3]``: the C ABI takes a row stride); falls back to a contiguous copy.  Returns (tensor, stride)."""
        t = t.detach()
        if t.device != device or t.dtype != torch.float32:
            t = t.to(device=device, dtype=torch.float32)
        B = t.shape[0]
        if t.numel() == B * width and B > 0 and t[0].is_contiguous() and (B == 1 or t.stride(0) >= width):
            return t, (t.stride(0) if B > 1 else width)
        t = t.contiguous().reshape(B, width)
        return t, width

    # ------------------------------------------------------------------ forward
    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None, right_hand_pose=None,
                transl=None, expression=None, jaw_pose=None, leye_pose=None, reye_pose=None, return_verts=True,
                return_full_pose=False, pose2rot=True, **kwargs):
        kw = dict(left_hand_pose=left_hand_pose, right_hand_pose=right_hand_pose, expression=expression,
                  jaw_pose=jaw_pose, leye_pose=leye_pose, reye_pose=reye_pose, return_verts=return_verts,
                  return_full_pose=return_full_pose, pose2rot=pose2rot)
        if torch.is_grad_enabled():
            # the module's own zero ``betas`` Parameter stands in when betas is omitted (body_models.py:875), and it
            # requires grad like the reference's -- so a call without betas under grad mode is differentiable too
            diff = [t for t in (betas if betas is not None else getattr(self, "betas", None), body_pose, global_orient, transl)
                    if t is not None and t.requires_grad]
            if diff:
                return self._forward_autograd(betas, global_orient, body_pose, transl, kw)
        out, _ = self.forward_camera(betas=betas, global_orient=global_orient, body_pose=body_pose, transl=transl, **kw)
        return out

    def _forward_autograd(self, betas, global_orient, body_pose, transl, kw):
        """``forward`` as ONE autograd node (``_SmplxFn``): what makes ``loss.backward()`` of the reference's training step
        (copenet_twoview.py:281-317,378-386) run through the native SMPL-X backward.  Only the hot-path call pattern is
        differentiable: betas / body_pose [B,21,3,3] / global_orient [B,1,3,3] / transl, no expression, jaw, eye or hand poses."""
        for name in ("left_hand_pose", "right_hand_pose", "expression", "jaw_pose", "leye_pose", "reye_pose"):
            if kw.get(name) is not None:
                raise NotImplementedError("airpose_b200.SMPLX: '{}' is not differentiable on the native path (off the AirPose "
                                          "hot path); call under torch.no_grad() or leave it out".format(name))
        if body_pose is None:
            raise NotImplementedError("airpose_b200.SMPLX: the differentiable call needs body_pose [B,21,3,3]")
        B = body_pose.shape[0]
        b_in = betas if betas is not None else self.betas
        if b_in.shape[0] != B:
            b_in = b_in.expand(B, -1)
        tr_in = transl if transl is not None else getattr(self, "transl", None)
        vertices, joints = _SmplxFn.apply(self, b_in, body_pose, global_orient, tr_in)
        full_pose = None
        if kw.get("return_full_pose"):
            eye = torch.eye(3, device=vertices.device, dtype=torch.float32).expand(B, 1, 3, 3)
            full_pose = torch.cat([global_orient.reshape(B, 1, 3, 3) if global_orient is not None else eye,
                                   body_pose.reshape(B, 21, 3, 3), eye.expand(B, 33, 3, 3)], dim=1)
        return ModelOutput(vertices=vertices if kw.get("return_verts", True) else None, joints=joints,
                           betas=betas if betas is not None else self.betas, expression=getattr(self, "expression", None),
                           global_orient=getattr(self, "global_orient", None), body_pose=body_pose,
                           left_hand_pose=getattr(self, "left_hand_pose", None),
                           right_hand_pose=getattr(self, "right_hand_pose", None), jaw_pose=None, full_pose=full_pose)

    def forward_camera(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None,
                       right_hand_pose=None, transl=None, expression=None, jaw_pose=None, leye_pose=None,
                       reye_pose=None, return_verts=True, return_full_pose=False, pose2rot=True,
                       root_R=None, root_t=None, focal_length=None, camera_center=None, proj_translation=None):
        """``forward`` plus, fused into the same kernels, ``transform_smpl`` (utils/utils.py:237-256;
        ``root_R`` [B,3,3], ``root_t`` [B,3]) and ``perspective_projection`` (utils/geometry.py:63-91;
        ``focal_length`` pair, ``camera_center`` [B,2], ``proj_translation`` [B,3] = its ``translation``).  Returns (ModelOutput, dict of camera outputs)."""
        if pose2rot:
            raise NotImplementedError("airpose_b200.SMPLX implements the rotation-matrix path only: pass pose2rot=False")
        if self.joint_mapper is not None:
            raise NotImplementedError("joint_mapper is not used on the AirPose hot path")
        device = self.v_template.device
        if device.type != "cuda":
            raise _lib.AirposeError("airpose_b200.SMPLX runs on CUDA only (module is on {}); there is no CPU path".format(device))
        lib = _lib.load()
        h = self._get_handle(device)

        def rot_or_identity(t, name, nj):
            if t is not None:
                if nj == 21 and t.dim() == 4 and tuple(t.shape[1:]) == (21, 3, 3):
                    return t                          # body_pose: handed over as a strided row view below
                return self._f32c(t, device).reshape(-1, nj, 3, 3)
            if not self._is_zero(name):
                raise NotImplementedError("non-zero module parameter '{}' with pose2rot=False needs batch_rodrigues, "
                                          "which is off the hot path".format(name))
            return None

        go = rot_or_identity(global_orient, "global_orient", 1)
        bp = rot_or_identity(body_pose, "body_pose", 21)
        tail_in = [rot_or_identity(jaw_pose, "jaw_pose", 1), rot_or_identity(leye_pose, "leye_pose", 1),
                   rot_or_identity(reye_pose, "reye_pose", 1)]
        for t, name in ((left_hand_pose, "left_hand_pose"), (right_hand_pose, "right_hand_pose")):
            if t is not None:
                raise NotImplementedError("explicit hand poses are not on the AirPose hot path")
            if not self._is_zero(name):
                raise NotImplementedError("non-zero module parameter '{}' is off the hot path".format(name))
            tail_in.append(None)

        betas_t = self._f32c(betas if betas is not None else self.betas, device)
        if expression is not None:
            shape_comp = torch.cat([betas_t, self._f32c(expression, device)], dim=-1)
        elif not self._is_zero("expression"):
            shape_comp = torch.cat([betas_t, self._f32c(self.expression, device)], dim=-1)
        else:
            shape_comp = betas_t          # zero expression contributes exactly nothing (body_models.py:943)
        pose_B = max(go.shape[0] if go is not None else 0, bp.shape[0] if bp is not None else 0)
        if betas is None and pose_B > 0 and betas_t.shape[0] not in (1, pose_B):
            # the reduced call of copenet_twoview.py:575-582 falls back to the module's own betas Parameter, created for
            # ``batch_size`` meshes; it is all zeros there, so any runtime batch can be served (the reference needs them equal)
            if not (self._is_zero("betas") and (expression is not None or self._is_zero("expression"))):
                raise ValueError("the module's betas Parameter holds {} rows, the poses {}".format(betas_t.shape[0], pose_B))
            shape_comp = torch.zeros(pose_B, shape_comp.shape[1], device=device, dtype=torch.float32)
            betas_t = shape_comp
        B = max(betas_t.shape[0], pose_B)
        if shape_comp.shape[0] != B:
            shape_comp = shape_comp.expand(B, -1).contiguous()

        tail = None
        if any(t is not None for t in tail_in[:3]):
            eye = torch.eye(3, device=device, dtype=torch.float32).expand(B, 1, 3, 3)
            parts = [t if t is not None else eye for t in tail_in[:3]] + [eye.expand(B, 30, 3, 3)]
            tail = torch.cat(parts, dim=1).contiguous()

        apply_trans = transl is not None or hasattr(self, "transl")
        tr = None
        if apply_trans:
            tr = self._f32c(transl if transl is not None else self.transl, device)

        V = self.v_template.shape[0]
        nj = self.J_regressor.shape[0] + self.vertex_joint_selector.extra_joints_idxs.shape[0] + self.lmk_faces_idx.shape[0]
        vertices = torch.empty(B, V, 3, device=device, dtype=torch.float32)
        joints = torch.empty(B, nj, 3, device=device, dtype=torch.float32)
        cam = {}
        if B == 0:                                   # empty batch: nothing to launch
            if root_R is not None or root_t is not None:
                cam["vertices_cam"], cam["joints_cam"] = torch.empty_like(vertices), torch.empty_like(joints)
            if focal_length is not None:
                cam["joints_2d"] = torch.empty(0, nj, 2, device=device, dtype=torch.float32)
            return ModelOutput(vertices=vertices if return_verts else None, joints=joints, betas=betas, body_pose=body_pose), cam
        a = _lib.SmplxFwdArgs()
        a.batch = B
        a.num_betas = shape_comp.shape[1]
        a.betas = shape_comp.data_ptr(); a.betas_stride = shape_comp.stride(0)
        if go is not None:
            a.global_orient = go.data_ptr(); a.global_orient_stride = 9
        if bp is not None:
            bp, bp_stride = self._rows(bp, device, 189)
            a.body_pose = bp.data_ptr(); a.body_pose_stride = bp_stride
        if tail is not None:
            a.tail_pose = tail.data_ptr(); a.tail_pose_stride = 33 * 9
        if tr is not None:
            a.transl = tr.data_ptr()
        keep = [shape_comp, go, bp, tail, tr]
        if root_R is not None or root_t is not None:
            if root_R is not None:
                rR, st = self._rows(root_R, device, 9); a.root_R = rR.data_ptr(); a.root_R_stride = st; keep.append(rR)
            if root_t is not None:
                rt, st = self._rows(root_t, device, 3); a.root_t = rt.data_ptr(); a.root_t_stride = st; keep.append(rt)
            cam["vertices_cam"] = torch.empty(B, V, 3, device=device, dtype=torch.float32)
            cam["joints_cam"] = torch.empty(B, nj, 3, device=device, dtype=torch.float32)
            a.out_vertices_cam = cam["vertices_cam"].data_ptr()
            a.out_joints_cam = cam["joints_cam"].data_ptr()
        if focal_length is not None:
            a.focal_x, a.focal_y = float(focal_length[0]), float(focal_length[1])
            if camera_center is not None:
                cc = self._f32c(camera_center, device).reshape(B, 2); a.center = cc.data_ptr(); a.center_stride = 2; keep.append(cc)
            if proj_translation is not None:
                pt = self._f32c(proj_translation, device).reshape(B, 3); a.proj_t = pt.data_ptr(); a.proj_t_stride = 3; keep.append(pt)
            cam["joints_2d"] = torch.empty(B, nj, 2, device=device, dtype=torch.float32)
            a.out_joints_2d = cam["joints_2d"].data_ptr()
        a.out_vertices = vertices.data_ptr()
        a.out_joints = joints.data_ptr()
        with torch.cuda.device(device):
            _lib.check(lib.airpose_smplx_fwd(h, C.byref(a), _lib.current_stream()), "airpose_smplx_fwd")
        del keep

        full_pose = None
        if return_full_pose:
            eye = torch.eye(3, device=device, dtype=torch.float32).expand(B, 1, 3, 3)
            full_pose = torch.cat([go if go is not None else eye,
                                   (bp.reshape(B, 21, 3, 3) if bp.dim() == 2 else bp) if bp is not None else eye.expand(B, 21, 3, 3),
                                   tail if tail is not None else eye.expand(B, 33, 3, 3)], dim=1)
        out = ModelOutput(vertices=vertices if return_verts else None, joints=joints,
                          betas=betas if betas is not None else self.betas,
                          expression=expression if expression is not None else getattr(self, "expression", None),
                          global_orient=getattr(self, "global_orient", None), body_pose=body_pose,
                          left_hand_pose=getattr(self, "left_hand_pose", None),
                          right_hand_pose=getattr(self, "right_hand_pose", None), jaw_pose=jaw_pose,
                          full_pose=full_pose)
        return out, cam


class _SmplxFn(torch.autograd.Function):
    """``SMPLX.forward(pose2rot=False)`` -> (vertices, joints) as one autograd node: forward = ``airpose_smplx_fwd``,
    backward = ``airpose_smplx_bwd`` (gradients w.r.t. betas, body_pose, global_orient; transl is additive, so its gradient
    is the sum of the upstream gradients over vertices and joints, lbs.py:219 / body_models.py:976-978)."""

    @staticmethod
    def forward(ctx, module, betas, body_pose, global_orient, transl):
        with torch.no_grad():
            out, _ = module.forward_camera(betas=betas, global_orient=global_orient, body_pose=body_pose, transl=transl,
                                           pose2rot=False)
        ctx.module = module
        ctx.has = (global_orient is not None, transl is not None)
        ctx.shapes = (tuple(betas.shape), tuple(body_pose.shape), None if global_orient is None else tuple(global_orient.shape))
        ctx.save_for_backward(betas.detach(), body_pose.detach(), None if global_orient is None else global_orient.detach())
        return out.vertices, out.joints

    @staticmethod
    def backward(ctx, g_vertices, g_joints):
        betas, body_pose, global_orient = ctx.saved_tensors
        g = smplx_backward(ctx.module, betas, body_pose, global_orient, grad_vertices=g_vertices, grad_joints=g_joints)
        sb, sp, so = ctx.shapes
        g_transl = None
        if ctx.has[1] and ctx.needs_input_grad[4]:
            g_transl = 0
            for t in (g_vertices, g_joints):
                if t is not None:
                    g_transl = g_transl + t.sum(dim=1)
        return (None, g["betas"].reshape(sb), g["body_pose"].reshape(sp),
                g["global_orient"].reshape(so) if ctx.has[0] else None, g_transl)


def smplx_backward(module, betas, body_pose, global_orient=None, grad_vertices=None, grad_joints=None,
                   grad_joints_cam=None, grad_joints_2d=None, joints=None, root_R=None, root_t=None, focal_length=None):
    """Gradient of ``SMPLX.forward(pose2rot=False)`` (+ ``transform_smpl`` + ``perspective_projection`` when the
    camera-frame gradients are given) w.r.t. betas, body_pose, global_orient, root_R, root_t -- what autograd derives
    for the reference (copenet_twoview.py:281-317).  Returns a dict of gradient tensors."""
    device = module.v_template.device
    if device.type != "cuda":
        raise _lib.AirposeError("smplx_backward runs on CUDA only; there is no CPU path")
    lib = _lib.load()
    h = module._get_handle(device)
    f = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
    betas, body_pose, global_orient = f(betas), f(body_pose), f(global_orient)
    B = betas.shape[0]
    a = _lib.SmplxBwdArgs()
    a.batch, a.num_betas = B, betas.shape[1]
    a.betas, a.betas_stride = betas.data_ptr(), betas.stride(0)
    keep = [betas, body_pose, global_orient]
    out = {"betas": torch.empty(B, betas.shape[1], device=device, dtype=torch.float32)}
    a.grad_betas = out["betas"].data_ptr()
    if body_pose is not None:
        body_pose = body_pose.reshape(B, 21, 3, 3)
        a.body_pose, a.body_pose_stride = body_pose.data_ptr(), 189
        out["body_pose"] = torch.empty(B, 21, 3, 3, device=device, dtype=torch.float32)
        a.grad_body_pose = out["body_pose"].data_ptr()
    if global_orient is not None:
        global_orient = global_orient.reshape(B, 1, 3, 3)
        a.global_orient, a.global_orient_stride = global_orient.data_ptr(), 9
    out["global_orient"] = torch.empty(B, 1, 3, 3, device=device, dtype=torch.float32)
    a.grad_global_orient = out["global_orient"].data_ptr()
    for name, t in (("grad_vertices", grad_vertices), ("grad_joints", grad_joints), ("grad_joints_cam", grad_joints_cam),
                    ("grad_joints_2d", grad_joints_2d), ("joints", joints)):
        t = f(t)
        if t is not None:
            keep.append(t)
            setattr(a, name, t.data_ptr())
    if root_R is not None:
        rR = f(root_R).reshape(B, 9); keep.append(rR); a.root_R, a.root_R_stride = rR.data_ptr(), 9
        out["root_R"] = torch.empty(B, 3, 3, device=device, dtype=torch.float32); a.grad_root_R = out["root_R"].data_ptr()
    if root_t is not None:
        rt = f(root_t).reshape(B, 3); keep.append(rt); a.root_t, a.root_t_stride = rt.data_ptr(), 3
        out["root_t"] = torch.empty(B, 3, device=device, dtype=torch.float32); a.grad_root_t = out["root_t"].data_ptr()
    if focal_length is not None:
        a.focal_x, a.focal_y = float(focal_length[0]), float(focal_length[1])
    with torch.cuda.device(device):
        _lib.check(lib.airpose_smplx_bwd(h, C.byref(a), _lib.current_stream()), "airpose_smplx_bwd")
    del keep
    return out


def rot6d_to_rotmat_backward(x, grad_R):
    """Backward of ``rot6d_to_rotmat``: x [B, 6k] (may be a strided view such as pred_pose[:, 3:]), grad_R [B*k,3,3]
    -> grad_x [B, 6k]."""
    if x.device.type != "cuda":
        raise _lib.AirposeError("rot6d_to_rotmat_backward runs on CUDA only; there is no CPU path")
    lib = _lib.load()
    x = x.detach().float()
    if x.stride(-1) != 1:
        x = x.contiguous()
    B, w = x.shape
    g = grad_R.detach().to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty(B, w, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.airpose_rot6d_to_rotmat_bwd_strided(x.data_ptr(), B, w // 6, x.stride(0), g.data_ptr(), out.data_ptr(), w,
                                                           _lib.current_stream()), "airpose_rot6d_to_rotmat_bwd_strided")
    return out


def rot6d_to_rotmat(x):
    """geometry.rot6d_to_rotmat (copenet/src/copenet/utils/geometry.py:47-61): [..,6k] -> [N,3,3]."""
    if x.device.type != "cuda":
        raise _lib.AirposeError("airpose_b200.rot6d_to_rotmat runs on CUDA only")
    lib = _lib.load()
    x = x.detach()
    if x.numel() == 0:
        return torch.empty(0, 3, 3, device=x.device, dtype=torch.float32)
    if x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 6 == 0 and x.dtype == torch.float32:
        groups, per, stride = x.shape[0], x.shape[1] // 6, x.stride(0)       # e.g. pred_pose[:, 3:] in place
    else:
        x = x.reshape(-1, 6).to(torch.float32).contiguous()
        groups, per, stride = x.shape[0], 1, 6
    out = torch.empty(groups * per, 3, 3, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.airpose_rot6d_to_rotmat_strided(x.data_ptr(), groups, per, stride, out.data_ptr(),
                                                       _lib.current_stream()), "airpose_rot6d_to_rotmat")
    return out


def joints_to_j14(joints, index_map=None):
    """SMPL joint -> 14 OpenPose joints gather (copenet_real_data/scripts/bundle_adj.py:48,116); bit-exact."""
    if joints.device.type != "cuda":
        raise _lib.AirposeError("airpose_b200.joints_to_j14 runs on CUDA only")
    lib = _lib.load()
    j = joints.detach().to(torch.float32).contiguous()
    B, nj = j.shape[0], j.shape[1]
    out = torch.empty(B, 14, 3, device=j.device, dtype=torch.float32)
    if B == 0:
        return out
    mp = None
    if index_map is not None:
        mp = (C.c_int32 * 14)(*[int(i) for i in index_map])
    with torch.cuda.device(j.device):
        _lib.check(lib.airpose_j14_gather(j.data_ptr(), B, nj, mp, out.data_ptr(), _lib.current_stream()),
                   "airpose_j14_gather")
    return out
