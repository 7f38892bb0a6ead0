"""CPU oracle for the AirPose copenet_twoview hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference algorithm.  It is the checker
for the CUDA path and the ``cpu_baseline`` / ``--impl reference`` arm of bench.py;
nothing under ``airpose_b200/`` may import it (the product path fails loudly when the
CUDA library is missing instead of falling back to this).

Parity pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the *real* reference
modules executed in the build container on seeded synthetic inputs
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``).

Every function cites the reference file:line it restates; paths are relative to
``/root/reference``.  ``bf16=True`` switches the trunk to the rounding points of the
CUDA kernels (bf16 operands, fp32 accumulate, one rounding per stored activation).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------

def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bfloat16 (ties to even), returned as fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def _maybe(x, bf16):
    return round_bf16(x) if bf16 else x


# --------------------------------------------------------------------------------------
# ResNet-50 trunk: copenet/src/copenet/models/model_copenet.py
# --------------------------------------------------------------------------------------

def conv2d(x, w, stride, pad):
    """nn.Conv2d(bias=False) on NCHW fp32 (model_copenet.py:16-21,57-58,98-100)."""
    n, c, h, wd = x.shape
    co, ci, kh, kw = w.shape
    assert ci == c
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (wd + 2 * pad - kw) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad))) if pad else x
    s = xp.strides
    cols = np.lib.stride_tricks.as_strided(
        xp, shape=(n, ho, wo, c, kh, kw),
        strides=(s[0], s[2] * stride, s[3] * stride, s[1], s[2], s[3]), writeable=False)
    a = cols.reshape(n * ho * wo, c * kh * kw)
    y = a @ w.reshape(co, -1).T.astype(np.float32)
    return y.reshape(n, ho, wo, co).transpose(0, 3, 1, 2)


def batchnorm_eval(x, sd, name, eps=1e-5):
    """nn.BatchNorm2d in eval mode (running statistics, eps 1e-5)."""
    scale = sd[name + ".weight"] / np.sqrt(sd[name + ".running_var"] + np.float32(eps))
    shift = sd[name + ".bias"] - sd[name + ".running_mean"] * scale
    return x * scale[None, :, None, None].astype(np.float32) + shift[None, :, None, None].astype(np.float32)


def maxpool_3x3_s2_p1(x):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (model_copenet.py:61)."""
    n, c, h, w = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), constant_values=-np.inf)
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    s = xp.strides
    win = np.lib.stride_tricks.as_strided(
        xp, shape=(n, c, ho, wo, 3, 3), strides=(s[0], s[1], s[2] * 2, s[3] * 2, s[2], s[3]), writeable=False)
    return win.max(axis=(4, 5))


def bottleneck(x, sd, p, stride, has_down, bf16):
    """Bottleneck.forward (model_copenet.py:27-47); stride sits on conv2 (:18-19)."""
    wq = (lambda k: round_bf16(sd[k])) if bf16 else (lambda k: sd[k])
    out = np.maximum(batchnorm_eval(conv2d(x, wq(p + ".conv1.weight"), 1, 0), sd, p + ".bn1"), 0)
    out = _maybe(out, bf16)
    out = np.maximum(batchnorm_eval(conv2d(out, wq(p + ".conv2.weight"), stride, 1), sd, p + ".bn2"), 0)
    out = _maybe(out, bf16)
    out = batchnorm_eval(conv2d(out, wq(p + ".conv3.weight"), 1, 0), sd, p + ".bn3")
    if has_down:
        res = batchnorm_eval(conv2d(x, wq(p + ".downsample.0.weight"), stride, 0), sd, p + ".downsample.1")
        res = _maybe(res, bf16)
    else:
        res = x
    return _maybe(np.maximum(out + res, 0), bf16)


def forward_feat_ext(x, sd, bf16=False, layers=(3, 4, 6, 3)):
    """copenet.forward_feat_ext (model_copenet.py:161-176): NCHW image -> [B,2048]."""
    x = _maybe(np.asarray(x, dtype=np.float32), bf16)
    w = round_bf16(sd["conv1.weight"]) if bf16 else sd["conv1.weight"]
    x = np.maximum(batchnorm_eval(conv2d(x, w, 2, 3), sd, "bn1"), 0)
    x = _maybe(maxpool_3x3_s2_p1(x), bf16)
    for li, blocks in enumerate(layers, start=1):
        for b in range(blocks):
            stride = 2 if (li > 1 and b == 0) else 1
            x = bottleneck(x, sd, "layer{}.{}".format(li, b), stride, b == 0, bf16)
    # nn.AvgPool2d(7, stride=1) on a 7x7 map, then flatten (:173-174)
    return x.mean(axis=(2, 3), dtype=np.float32).astype(np.float32)


# --------------------------------------------------------------------------------------
# IEF regressor: model_copenet.py:112-159,178-204
# --------------------------------------------------------------------------------------

def linear(x, sd, name):
    return x @ sd[name + ".weight"].T + sd[name + ".bias"]


def forward_reg(sd, xf0, xf1, bb0, bb1, pos0, pos1, ori0, ori1, art0, art1, sh0, sh1):
    """copenet.forward_reg in eval mode (dropout = identity) (model_copenet.py:178-204)."""
    xc0 = np.concatenate([xf0, bb0, pos0, ori0, art0, sh0, art1, sh1], axis=1)
    xc0 = linear(linear(xc0, sd, "fc1"), sd, "fc2")
    xc1 = np.concatenate([xf1, bb1, pos1, ori1, art1, sh1, art0, sh0], axis=1)
    xc1 = linear(linear(xc1, sd, "fc1"), sd, "fc2")
    nsh0 = sh0 + linear(xc0, sd, "decshape")
    npose0 = np.concatenate([pos0, ori0, art0], axis=1) + linear(xc0, sd, "decpose")
    nsh1 = sh1 + linear(xc1, sd, "decshape")
    npose1 = np.concatenate([pos1, ori1, art1], axis=1) + linear(xc1, sd, "decpose")
    return npose0, nsh0, npose1, nsh1


def ief_forward(sd, xf0, xf1, bb0, bb1, init_position0, init_position1, iters=3):
    """The regressor half of copenet.forward (model_copenet.py:118-159)."""
    b = xf0.shape[0]
    ori = np.broadcast_to(sd["init_pose"][:, :6], (b, 6))
    art = np.broadcast_to(sd["init_pose"][:, 6:22 * 6], (b, 126))
    sh = np.broadcast_to(sd["init_shape"], (b, 10))
    p0, s0, p1, s1 = forward_reg(sd, xf0, xf1, bb0, bb1, init_position0, init_position1,
                                 ori, ori, art, art, sh, sh)
    for _ in range(int(iters) - 1):
        p0, s0, p1, s1 = forward_reg(sd, xf0, xf1, bb0, bb1, p0[:, :3], p1[:, :3],
                                     p0[:, 3:9], p1[:, 3:9], p0[:, 9:], p1[:, 9:], s0, s1)
    return (p0.astype(np.float32), s0.astype(np.float32),
            p1.astype(np.float32), s1.astype(np.float32))


def copenet_forward(sd, x0, x1, bb0, bb1, init_position0, init_position1, iters=3, bf16=False):
    """copenet.forward (model_copenet.py:112-159): two trunk passes, then the IEF loop."""
    xf0 = forward_feat_ext(x0, sd, bf16)
    xf1 = forward_feat_ext(x1, sd, bf16)
    return ief_forward(sd, xf0, xf1, bb0, bb1, init_position0, init_position1, iters)


# --------------------------------------------------------------------------------------
# hmr baseline: copenet/src/copenet/models/model_hmr.py:112-172, copenet/src/copenet/hmr.py:127-158
# --------------------------------------------------------------------------------------

def hmr_forward_reg(sd, xf, pose, shape, cam):
    """model_hmr.copenet.forward_reg in eval mode (model_hmr.py:160-172)."""
    xc = np.concatenate([xf, pose, shape, cam], axis=1)
    xc = linear(linear(xc, sd, "fc1"), sd, "fc2")
    return (linear(xc, sd, "decpose") + pose, linear(xc, sd, "decshape") + shape, linear(xc, sd, "deccam") + cam)


def hmr_forward(sd, x, iters=3, bf16=False, feats=None):
    """model_hmr.copenet.forward (model_hmr.py:112-141): returns (rotmat [B,22,3,3], betas, cam, pose6d)."""
    xf = forward_feat_ext(x, sd, bf16) if feats is None else feats
    b = xf.shape[0]
    pose = np.broadcast_to(sd["init_pose"][:, :22 * 6], (b, 132))
    shape = np.broadcast_to(sd["init_shape"], (b, 10))
    cam = np.broadcast_to(sd["init_cam"], (b, 3))
    for _ in range(int(iters)):
        pose, shape, cam = hmr_forward_reg(sd, xf, pose, shape, cam)
    pose, shape, cam = pose.astype(np.float32), shape.astype(np.float32), cam.astype(np.float32)
    return rot6d_to_rotmat(pose).reshape(b, 22, 3, 3), shape, cam, pose


def hmr_fwd_pass(sd, m, x, iters=3, bf16=False, feats=None, focal_length=(1475.0, 1475.0), img_res=224):
    """hmr.fwd_pass_and_loss without the loss (hmr.py:127-158)."""
    rotmat, betas, cam, pose = hmr_forward(sd, x, iters, bf16, feats)
    b = rotmat.shape[0]
    verts, joints = smplx_forward(m, betas, rotmat[:, 1:], transl=np.zeros((b, 3), np.float32))             # :139-143
    tm = np.concatenate([rotmat[:, 0], np.zeros((b, 3, 1), np.float32)], axis=2)                             # :144
    pv, pj = transform_smpl(tm, verts, joints)                                                               # :146-148
    cam_t = np.stack([cam[:, 1], cam[:, 2], np.float32(2 * focal_length[0]) / (np.float32(img_res) * cam[:, 0] + np.float32(1e-9))],
                     axis=-1).astype(np.float32)                                                              # :149-151
    j2d = perspective_projection(pj + cam_t[:, None, :], focal_length, np.zeros((b, 2), np.float32))        # :153-157
    return {"pred_rotmat": rotmat, "pred_betas": betas, "pred_camera": cam, "pred_pose6d": pose, "pred_cam_t": cam_t,
            "vertices": verts, "joints": joints, "pred_vertices": pv, "pred_joints": pj, "pred_joints_2d_cam": j2d}


# --------------------------------------------------------------------------------------
# geometry: copenet/src/copenet/utils/geometry.py, utils/utils.py
# --------------------------------------------------------------------------------------

def rot6d_to_rotmat(x):
    """geometry.rot6d_to_rotmat (geometry.py:47-61); F.normalize eps = 1e-12."""
    x = np.asarray(x, dtype=np.float32).reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = a1 / np.maximum(np.sqrt(np.sum(a1 * a1, axis=1, keepdims=True)), np.float32(1e-12))
    u = a2 - np.sum(b1 * a2, axis=1, keepdims=True) * b1
    b2 = u / np.maximum(np.sqrt(np.sum(u * u, axis=1, keepdims=True)), np.float32(1e-12))
    b3 = np.cross(b1, b2)
    return np.stack([b1, b2, b3], axis=-1).astype(np.float32)


def transform_smpl(trans_mat, verts, joints):
    """utils.transform_smpl (utils/utils.py:237-256): x' = R x + t about the origin."""
    R, t = trans_mat[:, :3, :3], trans_mat[:, :3, 3]
    v = np.einsum("bij,bvj->bvi", R, verts) + t[:, None, :]
    j = np.einsum("bij,bvj->bvi", R, joints) + t[:, None, :]
    return v.astype(np.float32), j.astype(np.float32)


def perspective_projection(points, focal_length, camera_center):
    """geometry.perspective_projection with R = I, t = 0 (geometry.py:63-91)."""
    b = points.shape[0]
    K = np.zeros((b, 3, 3), dtype=np.float32)
    K[:, 0, 0] = focal_length[0]
    K[:, 1, 1] = focal_length[1]
    K[:, 2, 2] = 1.0
    K[:, :-1, -1] = camera_center
    proj = points / points[:, :, -1:]
    proj = np.einsum("bij,bkj->bki", K, proj)
    return proj[:, :, :-1].astype(np.float32)


# --------------------------------------------------------------------------------------
# SMPL-X: copenet/src/copenet/smplx/smplx/{lbs,body_models,vertex_joint_selector,vertex_ids}.py
# --------------------------------------------------------------------------------------

# vertex_ids.py:47-69 in the order of vertex_joint_selector.py:38-68
SMPLX_EXTRA_JOINT_VERTS = np.array(
    [9120, 9929, 9448, 616, 6,                      # nose, reye, leye, rear, lear
     5770, 5780, 8846, 8463, 8474, 8635,            # LBigToe LSmallToe LHeel RBigToe RSmallToe RHeel
     5361, 4933, 5058, 5169, 5286,                  # l thumb index middle ring pinky
     8079, 7669, 7794, 7905, 8022], dtype=np.int64)  # r thumb index middle ring pinky

# copenet_real_data/scripts/bundle_adj.py:48 -- SMPL joint -> 14 OpenPose joints
SMPL2OP_J14 = np.array([15, 12, 17, 19, 21, 16, 18, 20, 2, 5, 8, 1, 4, 7], dtype=np.int64)


class SmplxModel:
    """Buffers SMPL.__init__/SMPLX.__init__ register (body_models.py:205-296,727-730)."""

    def __init__(self, data):
        f32 = lambda k: np.asarray(data[k], dtype=np.float32)
        self.v_template = f32("v_template")
        self.shapedirs = f32("shapedirs")
        nb = data["posedirs"].shape[-1]
        self.posedirs = np.reshape(f32("posedirs"), [-1, nb]).T.copy()       # :284-288
        self.J_regressor = f32("J_regressor")
        parents = np.asarray(data["kintree_table"][0]).astype(np.float32).astype(np.int64)
        parents[0] = -1                                                       # :291-292
        self.parents = parents
        self.lbs_weights = f32("weights")
        self.faces = np.asarray(data["f"]).astype(np.int64)
        self.lmk_faces_idx = np.asarray(data["lmk_faces_idx"]).astype(np.int64)
        self.lmk_bary_coords = f32("lmk_bary_coords")


def blend_shapes(betas, shape_disps):
    """lbs.blend_shapes (lbs.py:245-266): einsum('bl,mkl->bmk')."""
    return np.einsum("bl,mkl->bmk", betas, shape_disps, optimize=True)


def vertices2joints(J_regressor, vertices):
    """lbs.vertices2joints (lbs.py:225-242): einsum('bik,ji->bjk')."""
    return np.einsum("bik,ji->bjk", vertices, J_regressor, optimize=True)


def batch_rigid_transform(rot_mats, joints, parents):
    """lbs.batch_rigid_transform (lbs.py:316-370) incl. transform_mat (:303-313)."""
    b, n = joints.shape[:2]
    rel = joints.copy()
    rel[:, 1:] -= joints[:, parents[1:]]
    T = np.zeros((b, n, 4, 4), dtype=np.float32)
    T[:, :, :3, :3] = rot_mats
    T[:, :, :3, 3] = rel
    T[:, :, 3, 3] = 1.0
    chain = [T[:, 0]]
    for i in range(1, n):
        chain.append(np.matmul(chain[parents[i]], T[:, i]))
    G = np.stack(chain, axis=1)
    posed = G[:, :, :3, 3].copy()
    jh = np.concatenate([joints, np.zeros((b, n, 1), dtype=np.float32)], axis=2)[..., None]
    A = G.copy()
    A[:, :, :, 3:4] -= np.matmul(G, jh)
    return posed, A


def lbs(betas, pose, m: SmplxModel):
    """lbs.lbs with pose2rot=False (lbs.py:135-222): pose is [B,55,3,3] rotation matrices."""
    b = max(betas.shape[0], pose.shape[0])
    v_shaped = m.v_template[None] + blend_shapes(betas, m.shapedirs)
    J = vertices2joints(m.J_regressor, v_shaped)
    ident = np.eye(3, dtype=np.float32)
    pose_feature = (pose[:, 1:].reshape(b, -1, 3, 3) - ident).reshape(b, -1)
    rot_mats = pose.reshape(b, -1, 3, 3)
    pose_offsets = (pose_feature @ m.posedirs).reshape(b, -1, 3)
    v_posed = pose_offsets + v_shaped
    J_t, A = batch_rigid_transform(rot_mats, J.astype(np.float32), m.parents)
    nj = m.J_regressor.shape[0]
    T = np.matmul(m.lbs_weights[None], A.reshape(b, nj, 16)).reshape(b, -1, 4, 4)
    vh = np.concatenate([v_posed, np.ones((b, v_posed.shape[1], 1), dtype=np.float32)], axis=2)
    v = np.matmul(T, vh[..., None])[:, :, :3, 0]
    return v.astype(np.float32), J_t.astype(np.float32)


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary_coords):
    """lbs.vertices2landmarks (lbs.py:96-132)."""
    lmk_faces = faces[lmk_faces_idx]                       # [L,3]
    lmk_vertices = vertices[:, lmk_faces]                  # [B,L,3,3]
    return np.einsum("blfi,lf->bli", lmk_vertices, lmk_bary_coords).astype(np.float32)


def smplx_forward(m: SmplxModel, betas, body_pose, global_orient=None, transl=None, expression=None):
    """SMPLX.forward on the pose2rot=False path (body_models.py:820-994).

    Jaw / eye / hand poses come from the module's zero Parameters through
    batch_rodrigues, which is exactly the identity (:878-925); ``pose_mean`` is skipped
    (:934-935); expression defaults to zeros (:918,943).
    Returns (vertices [B,10475,3], joints [B,127,3]).
    """
    b = betas.shape[0]
    eye = np.broadcast_to(np.eye(3, dtype=np.float32), (b, 1, 3, 3))
    if global_orient is None:
        global_orient = eye
    rest = np.broadcast_to(np.eye(3, dtype=np.float32), (b, 33, 3, 3))
    full_pose = np.concatenate([global_orient.reshape(b, 1, 3, 3), body_pose.reshape(b, 21, 3, 3), rest], axis=1)
    if expression is None:
        expression = np.zeros((b, 10), dtype=np.float32)
    shape_components = np.concatenate([betas, expression], axis=-1)
    verts, joints = lbs(shape_components, full_pose.astype(np.float32), m)
    landmarks = vertices2landmarks(verts, m.faces, m.lmk_faces_idx, m.lmk_bary_coords)
    extra = verts[:, SMPLX_EXTRA_JOINT_VERTS]              # vertex_joint_selector.py:73-77
    joints = np.concatenate([joints, extra, landmarks], axis=1)
    if transl is not None:
        joints = joints + transl[:, None, :]
        verts = verts + transl[:, None, :]
    return verts.astype(np.float32), joints.astype(np.float32)


def j14_from_joints(joints):
    """SMPL joint -> 14 OpenPose joints index map (copenet_real_data/scripts/bundle_adj.py:48,116)."""
    return joints[:, SMPL2OP_J14]


# --------------------------------------------------------------------------------------
# the whole per-frame path: copenet/src/copenet/copenet_twoview.py:164-317
# --------------------------------------------------------------------------------------

TRANS_SCALE = 0.05


def twoview_forward(sd, m: SmplxModel, batch, iters=3, bf16=False, focal_length=(1475.0, 1475.0),
                    feats=None):
    """copenet_twoview.fwd_pass_and_loss without the loss (copenet_twoview.py:164-317).

    ``batch`` holds im0/im1, bb0/bb1, intr0/intr1.  ``feats`` optionally supplies
    precomputed trunk features (xf0, xf1) so the expensive trunk can be skipped.
    """
    b = batch["bb0"].shape[0]
    init = np.tile(np.array([0, 0, 10], dtype=np.float32), (b, 1)) * np.float32(TRANS_SCALE)   # :184-203
    if feats is None:
        xf0 = forward_feat_ext(batch["im0"], sd, bf16)
        xf1 = forward_feat_ext(batch["im1"], sd, bf16)
    else:
        xf0, xf1 = feats
    p0, s0, p1, s1 = ief_forward(sd, xf0, xf1, batch["bb0"], batch["bb1"], init, init, iters)
    out = {"xf0": xf0, "xf1": xf1}
    for v, (p, s) in enumerate(((p0, s0), (p1, s1))):
        p = p.copy()
        p[:, :3] /= np.float32(TRANS_SCALE)                                                      # :214-218
        trans = p[:, :3]
        R = rot6d_to_rotmat(p[:, 3:]).reshape(b, 22, 3, 3)                                       # :222-223
        verts, joints = smplx_forward(m, s, R[:, 1:], transl=np.zeros((b, 3), np.float32))       # :281-285
        tm = np.concatenate([R[:, 0], trans[:, :, None]], axis=2)                                # :287-288
        vc, jc = transform_smpl(tm, verts, joints)                                               # :290-292
        j2d = perspective_projection(jc, focal_length, batch["intr%d" % v][:, :2, 2])            # :307-317
        out.update({"pred_pose%d" % v: p, "pred_betas%d" % v: s, "pred_rotmat%d" % v: R,
                    "pred_smpltrans%d" % v: trans, "vertices%d" % v: verts, "joints%d" % v: joints,
                    "pred_vertices_cam%d" % v: vc, "pred_joints_cam%d" % v: jc,
                    "pred_joints_2d_cam%d" % v: j2d})
    return out


def get_loss(hp, gt, out):
    """copenet_twoview.get_loss (copenet_twoview.py:83-161).  ``hp`` = loss weights dict
    (defaults :655-677); ``gt`` = batch dict; ``out`` = twoview_forward result."""
    mse = lambda a, b: (a - b) ** 2
    gv = gt["smpl_vertices"].squeeze(1)
    gj = gt["smpl_joints"].squeeze(1)
    l_kp = sum(mse(out["pred_joints_2d_cam%d" % v][:, :22], gt["smpl_joints_2d%d" % v].squeeze(1)[:, :22]).mean()
               for v in (0, 1))
    l3 = (mse(out["joints0"][:, :22], gj[:, :22]) + mse(out["joints1"][:, :22], gj[:, :22])
          + mse(out["joints0"][:, :22], out["joints1"][:, :22]))
    l3[:, [4, 5, 18, 19]] *= hp["limbs3d_loss_weight"]
    l3[:, [7, 8, 20, 21]] *= hp["limbs3d_loss_weight"] ** 2
    l_kp3d = l3.mean()
    l_shape = (mse(out["vertices0"], gv).mean() + mse(out["vertices1"], gv).mean()
               + mse(out["vertices0"], out["vertices1"]).mean())
    l_trans = sum(mse(out["pred_smpltrans%d" % v], gt["smpltrans_rel%d" % v]).mean() for v in (0, 1))
    l_root = sum(mse(out["pred_rotmat%d" % v][:, :1], gt["smplorient_rel%d" % v]).mean() for v in (0, 1))
    lr = (mse(out["pred_rotmat0"][:, 1:], gt["smplpose_rotmat"]) + mse(out["pred_rotmat1"][:, 1:], gt["smplpose_rotmat"])
          + mse(out["pred_rotmat0"][:, 1:], out["pred_rotmat1"][:, 1:]))
    lr[:, [3, 4, 17, 18]] *= hp["limbstheta_loss_weight"]
    lr[:, [6, 7, 19, 20]] *= hp["limbstheta_loss_weight"] ** 2
    l_pose = lr.mean()
    b0, b1 = out["pred_betas0"], out["pred_betas1"]
    l_beta = (b0 * b0).mean() + (b1 * b1).mean() + mse(b0, b1).mean()
    loss = (hp["trans_loss_weight"] * l_trans + hp["keypoint2d_loss_weight"] * l_kp
            + hp["keypoint3d_loss_weight"] * l_kp3d + hp["shape_loss_weight"] * l_shape
            + hp["rootrot_loss_weight"] * l_root + hp["pose_loss_weight"] * l_pose
            + hp["beta_loss_weight"] * l_beta) * 60
    return float(loss), {"loss_regr_trans": float(l_trans), "loss_keypoints": float(l_kp),
                         "loss_keypoints_3d": float(l_kp3d), "loss_regr_shape": float(l_shape),
                         "loss_rootrot": float(l_root), "loss_regr_pose": float(l_pose),
                         "loss_regul_betas": float(l_beta)}


def real_get_loss(hp, gt, out, vposer_term=0.0):
    """copenet_real's get_loss (copenet_real/src/copenet_real/copenet_twoview.py:99-160) WITHOUT the VPoser prior (its
    weights are an external download; ``vposer_term`` = loss_regul_vposer if the caller has it, 0 otherwise):
    confidence-weighted 2D keypoint loss over the first 22 joints with the limb weights (:115-121), cross-view pose
    consistency (:137), beta regulariser + cross-view beta consistency (:139-141), the exp(-t_z)^2 depth barrier (:148-149),
    x60 (:151).  ``gt['smpl_joints_2d%d']`` is [B,1,J,3] = (x, y, confidence)."""
    mse = lambda a, b: (a - b) ** 2
    g0, g1 = gt["smpl_joints_2d0"][:, 0], gt["smpl_joints_2d1"][:, 0]
    lk = (mse(out["pred_joints_2d_cam0"][:, :22], g0[:, :22, :2]) * g0[:, :22, 2:]
          + mse(out["pred_joints_2d_cam1"][:, :22], g1[:, :22, :2]) * g1[:, :22, 2:])
    lk[:, [4, 5, 18, 19]] *= hp["limbs2d_loss_weight"]
    lk[:, [7, 8, 20, 21]] *= hp["limbs2d_loss_weight"] ** 2
    l_kp = lk.mean()
    l_pose = mse(out["pred_rotmat0"][:, 1:], out["pred_rotmat1"][:, 1:]).mean()
    b0, b1 = out["pred_betas0"], out["pred_betas1"]
    l_beta = (b0 * b0).mean() + (b1 * b1).mean() + mse(b0, b1).mean()
    loss = (hp["keypoint2d_loss_weight"] * l_kp + hp["beta_loss_weight"] * l_beta + hp["vposer_loss_weight"] * vposer_term
            + hp["pose_loss_weight"] * l_pose + (np.exp(-out["pred_smpltrans0"][:, 2]) ** 2).mean()
            + (np.exp(-out["pred_smpltrans1"][:, 2]) ** 2).mean()) * 60
    return float(loss), {"loss": float(loss), "loss_regul_vposer": float(vposer_term), "loss_regr_pose": float(l_pose),
                         "loss_keypoints": float(l_kp), "loss_regul_betas": float(l_beta)}


DEFAULT_LOSS_WEIGHTS = {   # copenet_twoview.py:655-677
    "shape_loss_weight": 50, "keypoint2d_loss_weight": 0.002, "keypoint3d_loss_weight": 1,
    "limbs3d_loss_weight": 3, "limbstheta_loss_weight": 1, "trans_loss_weight": 10,
    "rootrot_loss_weight": 1, "pose_loss_weight": 50, "beta_loss_weight": 1}


# --------------------------------------------------------------------------------------
# input preprocessing and the staged drone-server protocol
# --------------------------------------------------------------------------------------
IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], np.float32)      # server.py:75-76, aerialpeople.py (Normalize)
IMAGENET_STD = np.array([0.229, 0.224, 0.225], np.float32)
SERVER_SIZE = 224                                                # server.py:37
SERVER_BUFFERSIZE = 1 + 3 * 4 + SERVER_SIZE * SERVER_SIZE * 3    # server.py:38: u8 stage, 3 x f32 bb, BGR bytes
SERVER_BUFFERSIZE_STAGES = 1 + (10 + 21 * 6) * 4                 # server.py:39: u8 stage, 136 x f32 (betas | articulated pose)


def server_preprocess(data) -> np.ndarray:
    """Stage-0 image decoding of airpose_server/server.py:91-98: BGR bytes -> RGB -> CHW -> * (1/255) -> (x - mean) / std,
    each step one float32 operation.  Returns [1,3,224,224] float32."""
    npimg = np.frombuffer(data, dtype=np.uint8, count=SERVER_SIZE * SERVER_SIZE * 3, offset=1 + 3 * 4)
    npimg = npimg.reshape(SERVER_SIZE, SERVER_SIZE, 3)[:, :, [2, 1, 0]].transpose(2, 0, 1)
    frame = npimg[None].astype(np.float32) * np.float32(1.0 / 255)
    frame = (frame - IMAGENET_MEAN[None, :, None, None]) / IMAGENET_STD[None, :, None, None]
    return frame.astype(np.float32)


def server_forward_reg(sd, xf0, bb0, pos0, ori0, art0, art1, sh0, sh1):
    """The server model's single-view regressor pass (airpose_server/airpose.py:179-195; eval mode)."""
    xc0 = np.concatenate([xf0, bb0, pos0, ori0, art0, sh0, art1, sh1], axis=1)
    xc0 = linear(linear(xc0, sd, "fc1"), sd, "fc2")
    return (np.concatenate([pos0, ori0, art0], axis=1) + linear(xc0, sd, "decpose")).astype(np.float32), \
           (sh0 + linear(xc0, sd, "decshape")).astype(np.float32)


class ServerState:
    """The globals of airpose_server/server.py:69-73 (one connection's network state)."""

    def __init__(self, sd):
        self.xf = np.zeros((1, 2048), np.float32)
        self.bb = np.zeros((1, 3), np.float32)
        self.curr_pose = np.array(sd["init_pose"], np.float32).copy()        # [1,144]; replaced by pose[:, 3:] ([1,132]) after a stage
        self.curr_shape = np.array(sd["init_shape"], np.float32).copy()
        self.curr_position = np.array([[0, 0, 0.5]], np.float32)


def server_process(sd, state: ServerState, data, stage: int, feat_fn=None, bf16=False) -> np.ndarray:
    """``process(data, metainfo, stage)`` of airpose_server/server.py:78-150: returns the reply as a float32 vector
    (136 floats for stages 0-1: betas | articulated pose; 145 for stage 2: betas | full pose) and advances ``state``.
    ``feat_fn(frame) -> [1,2048]`` replaces the oracle trunk (tests feed the device features to isolate the regressor)."""
    init_pose, init_shape = np.asarray(sd["init_pose"], np.float32), np.asarray(sd["init_shape"], np.float32)
    if stage == 0:
        state.bb = np.frombuffer(data, dtype=np.float32, count=3, offset=1)[None].copy()
        frame = server_preprocess(data)
        state.xf = feat_fn(frame) if feat_fn is not None else forward_feat_ext(frame, sd, bf16=bf16)
        pose, shape = server_forward_reg(sd, state.xf, state.bb, np.array([[0, 0, 0.5]], np.float32), init_pose[:, :6],
                                         init_pose[:, 6:22 * 6], init_pose[:, 6:22 * 6], init_shape, init_shape)
        reply = np.concatenate([shape[0], pose[0, 9:]])
    elif stage in (1, 2):
        shape2 = np.frombuffer(data, dtype=np.float32, count=10, offset=1)[None]
        art2 = np.frombuffer(data, dtype=np.float32, count=126, offset=41)[None]
        pose, shape = server_forward_reg(sd, state.xf, state.bb, state.curr_position, state.curr_pose[:, :6],
                                         state.curr_pose[:, 6:22 * 6], art2, state.curr_shape, shape2)
        reply = np.concatenate([shape[0], pose[0, 9:] if stage == 1 else pose[0]])
    else:
        raise ValueError("Invalid stage number {}".format(stage))            # server.py:141-142 prints; nothing to reply
    state.curr_position, state.curr_pose, state.curr_shape = pose[:, :3], pose[:, 3:], shape     # server.py:144-146
    return reply.astype(np.float32)


def _cv_linear_coef(dst_size: int, src_size: int):
    """Source index and weight per destination index of cv2.resize(INTER_LINEAR) on a CV_64F image (OpenCV imgproc/resize.cpp):
    scale = 1 / (dst / src); f = (d + 0.5) * scale - 0.5; s = floor(f); f -= s; clamped at both ends (s < 0 -> s = 0, f = 0;
    s >= src - 1 -> s = src - 1, f = 0).  All in double: measured against the cv2 4.13.0 of the build container, whose
    sample positions on a ramp image are exact to 1e-14 (the reference pins opencv-python 4.5.1.48, requirements.txt:4,
    which is not installable offline; a float-coefficient variant would differ by <= 4e-5 on [0,1] pixel values)."""
    scale = 1.0 / (float(dst_size) / float(src_size))
    d = np.arange(dst_size, dtype=np.float64)
    f = (d + 0.5) * scale - 0.5
    s = np.floor(f).astype(np.int64)
    f = f - s
    lo = s < 0
    f[lo] = 0; s[lo] = 0
    hi = s >= src_size - 1
    f[hi] = 0; s[hi] = src_size - 1
    return s, f


def cv_resize_linear(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h)) for a float64 [H,W,C] image: bilinear, pixel-centre aligned, no antialiasing,
    horizontal pass then vertical pass in double (what resize_with_pad calls, utils.py:224)."""
    img = np.asarray(img, np.float64)
    h, w = img.shape[:2]
    sx, fx = _cv_linear_coef(dst_w, w)
    sy, fy = _cv_linear_coef(dst_h, h)
    sx1, sy1 = np.minimum(sx + 1, w - 1), np.minimum(sy + 1, h - 1)
    rows = img[:, sx] * (1.0 - fx)[None, :, None] + img[:, sx1] * fx[None, :, None]      # [H, dst_w, C]
    return rows[sy] * (1.0 - fy)[:, None, None] + rows[sy1] * fy[:, None, None]


def resize_with_pad(img: np.ndarray, size: int = 224):
    """utils.resize_with_pad (copenet/src/copenet/utils/utils.py:214-235): longer side -> size, zero letterbox."""
    bigger = img.shape[0] if img.shape[0] > img.shape[1] else img.shape[1]
    scale = size / bigger
    out = cv_resize_linear(img, int(scale * img.shape[1]), int(scale * img.shape[0]))
    pad_top = (size - out.shape[0]) // 2
    pad_left = (size - out.shape[1]) // 2
    full = np.zeros((size, size, img.shape[2]), np.float64)
    full[pad_top:pad_top + out.shape[0], pad_left:pad_left + out.shape[1]] = out
    return full, scale, [pad_left, pad_top]


def dataset_preprocess(frame_bgr_u8: np.ndarray, rect, size: int = 224):
    """The per-camera image path of aerialpeople.__getitem__ (dsets/aerialpeople.py:125-141,174): BGR u8 frame ->
    [:, :, ::-1] / 255. -> crop rect = (y0, y1, x0, x1) -> resize_with_pad -> CHW float32 -> Normalize(mean, std).
    Returns (image [3,size,size] float32, scale, [pad_left, pad_top])."""
    y0, y1, x0, x1 = (int(v) for v in rect)
    img = frame_bgr_u8[:, :, ::-1] / 255.
    img, scale, pad = resize_with_pad(img[y0:y1, x0:x1, :], size)
    chw = img.transpose(2, 0, 1).astype(np.float32)
    chw = (chw - IMAGENET_MEAN[:, None, None]) / IMAGENET_STD[:, None, None]
    return chw.astype(np.float32), scale, pad


# --------------------------------------------------------------------------------------
# test-mode conversions and metrics  --  PARITY UNPINNED
# torchgeometry==0.1.2 (requirements.txt) is not under /root/reference and not installable offline; the two functions below
# restate the published torchgeometry/core/conversions.py of that version, reading `1 - mask` on its boolean masks as logical
# NOT (the literal expression raises under torch >= 1.2).  Pinned by known-answer identities only (tests/test_oracle_golden.py).
# --------------------------------------------------------------------------------------
def tgm_rotation_matrix_to_angle_axis(rotation_matrix, eps=1e-6):
    """tgm.rotation_matrix_to_angle_axis on [N,3,4] / [N,3,3] (copenet_twoview.py:323-326): matrix -> quaternion (w,x,y,z)
    through the transposed matrix and the four-branch trace test, quaternion -> angle-axis with the atan2 form."""
    R = np.asarray(rotation_matrix, np.float32)[:, :3, :3]
    t = np.transpose(R, (0, 2, 1))
    d2 = t[:, 2, 2] < eps
    d0_d1 = t[:, 0, 0] > t[:, 1, 1]
    d0_nd1 = t[:, 0, 0] < -t[:, 1, 1]
    t0 = 1 + t[:, 0, 0] - t[:, 1, 1] - t[:, 2, 2]
    q0 = np.stack([t[:, 1, 2] - t[:, 2, 1], t0, t[:, 0, 1] + t[:, 1, 0], t[:, 2, 0] + t[:, 0, 2]], -1)
    t1 = 1 - t[:, 0, 0] + t[:, 1, 1] - t[:, 2, 2]
    q1 = np.stack([t[:, 2, 0] - t[:, 0, 2], t[:, 0, 1] + t[:, 1, 0], t1, t[:, 1, 2] + t[:, 2, 1]], -1)
    t2 = 1 - t[:, 0, 0] - t[:, 1, 1] + t[:, 2, 2]
    q2 = np.stack([t[:, 0, 1] - t[:, 1, 0], t[:, 2, 0] + t[:, 0, 2], t[:, 1, 2] + t[:, 2, 1], t2], -1)
    t3 = 1 + t[:, 0, 0] + t[:, 1, 1] + t[:, 2, 2]
    q3 = np.stack([t3, t[:, 1, 2] - t[:, 2, 1], t[:, 2, 0] - t[:, 0, 2], t[:, 0, 1] - t[:, 1, 0]], -1)
    c0, c1, c2, c3 = d2 & d0_d1, d2 & ~d0_d1, ~d2 & d0_nd1, ~d2 & ~d0_nd1
    f = lambda m: m[:, None].astype(np.float32)
    q = q0 * f(c0) + q1 * f(c1) + q2 * f(c2) + q3 * f(c3)
    q = q / np.sqrt(t0[:, None] * f(c0) + t1[:, None] * f(c1) + t2[:, None] * f(c2) + t3[:, None] * f(c3))
    q = q * np.float32(0.5)
    sin2 = q[:, 1] ** 2 + q[:, 2] ** 2 + q[:, 3] ** 2
    sn = np.sqrt(sin2)
    two_theta = 2.0 * np.where(q[:, 0] < 0.0, np.arctan2(-sn, -q[:, 0]), np.arctan2(sn, q[:, 0]))
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(sin2 > 0.0, two_theta / sn, 2.0)
    return (q[:, 1:] * k[:, None]).astype(np.float32)


def tgm_angle_axis_to_rotation_matrix(angle_axis, eps=1e-6):
    """tgm.angle_axis_to_rotation_matrix (copenet_twoview.py:558-559): [N,3] -> [N,4,4]; Rodrigues above theta^2 = 1e-6 (axis
    = aa / (theta + 1e-6)), first-order Taylor below."""
    aa = np.asarray(angle_axis, np.float32).reshape(-1, 3)
    theta2 = (aa * aa).sum(1)
    theta = np.sqrt(theta2)
    w = aa / (theta + np.float32(eps))[:, None]
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    c, s_ = np.cos(theta), np.sin(theta)
    k = 1 - c
    normal = np.stack([c + wx * wx * k, wx * wy * k - wz * s_, wy * s_ + wx * wz * k,
                       wz * s_ + wx * wy * k, c + wy * wy * k, -wx * s_ + wy * wz * k,
                       -wy * s_ + wx * wz * k, wx * s_ + wy * wz * k, c + wz * wz * k], 1).reshape(-1, 3, 3)
    one = np.ones_like(theta)
    taylor = np.stack([one, -aa[:, 2], aa[:, 1], aa[:, 2], one, -aa[:, 0], -aa[:, 1], aa[:, 0], one], 1).reshape(-1, 3, 3)
    out = np.tile(np.eye(4, dtype=np.float32), (aa.shape[0], 1, 1))
    out[:, :3, :3] = np.where((theta2 > eps)[:, None, None], normal, taylor)
    return out.astype(np.float32)


def mean_distance(a, b, points_used=None):
    """np.mean(np.sqrt(np.sum((a - b) ** 2, -1))[:, :points_used])  (copenet_twoview.py:541-551,583-586)."""
    d = np.sqrt(np.sum((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2, -1))
    if d.ndim == 1:
        return float(d.mean())
    return float(d[:, :points_used].mean())


def test_mode_outputs(sd, m: SmplxModel, batch, out):
    """The extra outputs of the is_test branch (copenet_twoview.py:258-279,323-326) from ``twoview_forward``'s result."""
    b = out["pred_pose0"].shape[0]
    res = {}
    for v in (0, 1):
        R = out["pred_rotmat%d" % v]
        verts, joints = smplx_forward(m, np.zeros((b, 10), np.float32), R[:, 1:], transl=np.zeros((b, 3), np.float32))
        init = np.tile(np.array([0, 0, 10], dtype=np.float32), (b, 1))
        tm = np.concatenate([np.tile(np.eye(3, dtype=np.float32), (b, 1, 1)), init[:, :, None]], axis=2)
        res["pred_vertices_cam_in%d" % v] = transform_smpl(tm, verts, joints)[0]
        res["pred_angles%d" % v] = tgm_rotation_matrix_to_angle_axis(R.reshape(-1, 3, 3)).reshape(b, 22, 3)
        gt = np.concatenate([batch["smplorient_rel%d" % v], batch["smplpose_rotmat"]], axis=1)
        res["gt_angles%d" % v] = tgm_rotation_matrix_to_angle_axis(gt.reshape(-1, 3, 3)).reshape(b, 22, 3)
    return res
