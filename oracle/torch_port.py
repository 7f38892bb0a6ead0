"""The reference path restated with PyTorch CPU ops  --  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's runtime *is* PyTorch ("no native code on this path"), and
/root/reference does not exist on the GPU box, so the CPU arm of bench.py
(``--impl reference`` and ``cpu_baseline``, kind "port") times this port: the same ATen
ops the reference modules call, in the same order, in fp32, on all host threads.  It is
pinned against the same golden fixtures as the numpy oracle (tests/test_oracle_golden.py).
Nothing under airpose_b200/ imports it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EXTRA = [9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
         5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022]      # vertex_ids.py:47-69


def to_torch(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, 1e-5)


def forward_feat_ext(x, sd):
    """model_copenet.py:161-176 with Bottleneck.forward :27-47."""
    x = F.max_pool2d(F.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1")), 3, 2, 1)
    for li, blocks in enumerate((3, 4, 6, 3), start=1):
        for b in range(blocks):
            p = "layer{}.{}".format(li, b)
            s = 2 if (li > 1 and b == 0) else 1
            out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
            out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=s, padding=1), sd, p + ".bn2"))
            out = _bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            if b == 0:
                x = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=s), sd, p + ".downsample.1")
            x = F.relu(out + x)
    return F.avg_pool2d(x, 7, stride=1).flatten(1)


def ief(sd, xf0, xf1, bb0, bb1, pos0, pos1, iters=3):
    """model_copenet.py:118-159,178-204 (eval: dropout is the identity)."""
    b = xf0.shape[0]
    ori0 = ori1 = sd["init_pose"][:, :6].expand(b, -1)
    art0 = art1 = sd["init_pose"][:, 6:132].expand(b, -1)
    sh0 = sh1 = sd["init_shape"].expand(b, -1)
    lin = lambda x, n: F.linear(x, sd[n + ".weight"], sd[n + ".bias"])
    for _ in range(int(iters)):
        xc0 = lin(lin(torch.cat([xf0, bb0, pos0, ori0, art0, sh0, art1, sh1], 1), "fc1"), "fc2")
        xc1 = lin(lin(torch.cat([xf1, bb1, pos1, ori1, art1, sh1, art0, sh0], 1), "fc1"), "fc2")
        p0 = torch.cat([pos0, ori0, art0], 1) + lin(xc0, "decpose")
        p1 = torch.cat([pos1, ori1, art1], 1) + lin(xc1, "decpose")
        sh0, sh1 = sh0 + lin(xc0, "decshape"), sh1 + lin(xc1, "decshape")
        pos0, ori0, art0 = p0[:, :3], p0[:, 3:9], p0[:, 9:]
        pos1, ori1, art1 = p1[:, :3], p1[:, 3:9], p1[:, 9:]
    return p0, sh0, p1, sh1


def ief_train(sd, xf0, xf1, bb0, bb1, pos0, pos1, mask1, mask2, iters=3):
    """The same loop in training mode (model_copenet.py:186-189: fc1 -> drop1 -> fc2 -> drop2) with the dropout masks
    given explicitly: mask[it, view] is the multiplicative mask (0 or 1/(1-p)) of that Dropout call.  Differentiable:
    the parity tests of the CUDA backward run autograd through it."""
    b = xf0.shape[0]
    ori0 = ori1 = sd["init_pose"][:, :6].expand(b, -1)
    art0 = art1 = sd["init_pose"][:, 6:132].expand(b, -1)
    sh0 = sh1 = sd["init_shape"].expand(b, -1)
    lin = lambda x, n: F.linear(x, sd[n + ".weight"], sd[n + ".bias"])
    for it in range(int(iters)):
        xc0 = lin(lin(torch.cat([xf0, bb0, pos0, ori0, art0, sh0, art1, sh1], 1), "fc1") * mask1[it, 0], "fc2") * mask2[it, 0]
        xc1 = lin(lin(torch.cat([xf1, bb1, pos1, ori1, art1, sh1, art0, sh0], 1), "fc1") * mask1[it, 1], "fc2") * mask2[it, 1]
        p0 = torch.cat([pos0, ori0, art0], 1) + lin(xc0, "decpose")
        p1 = torch.cat([pos1, ori1, art1], 1) + lin(xc1, "decpose")
        sh0, sh1 = sh0 + lin(xc0, "decshape"), sh1 + lin(xc1, "decshape")
        pos0, ori0, art0 = p0[:, :3], p0[:, 3:9], p0[:, 9:]
        pos1, ori1, art1 = p1[:, :3], p1[:, 3:9], p1[:, 9:]
    return p0, sh0, p1, sh1


def rot6d_to_rotmat(x):
    """geometry.py:47-61."""
    x = x.reshape(-1, 3, 2)
    b1 = F.normalize(x[:, :, 0])
    b2 = F.normalize(x[:, :, 1] - torch.einsum("bi,bi->b", b1, x[:, :, 1]).unsqueeze(-1) * b1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=1)), dim=-1)


class Smplx:
    def __init__(self, data):
        f = lambda k: torch.from_numpy(np.asarray(data[k], dtype=np.float32))
        self.v_template, self.shapedirs, self.J_regressor, self.weights = f("v_template"), f("shapedirs"), f("J_regressor"), f("weights")
        nb = data["posedirs"].shape[-1]
        self.posedirs = torch.from_numpy(np.reshape(np.asarray(data["posedirs"], np.float32), [-1, nb]).T.copy())
        par = np.asarray(data["kintree_table"][0]).astype(np.float32).astype(np.int64)
        par[0] = -1
        self.parents = par
        self.faces = torch.from_numpy(np.asarray(data["f"]).astype(np.int64))
        self.lmk_faces_idx = torch.from_numpy(np.asarray(data["lmk_faces_idx"]).astype(np.int64))
        self.lmk_bary = f("lmk_bary_coords")


def smplx_forward(m: Smplx, betas, body_pose):
    """body_models.py:820-994 + lbs.py:135-222 on the pose2rot=False path (identity root / face / hands,
    zero expression, zero transl)."""
    b = betas.shape[0]
    eye = torch.eye(3).expand(b, 1, 3, 3)
    pose = torch.cat([eye, body_pose, eye.expand(b, 33, 3, 3)], 1)
    shape = torch.cat([betas, torch.zeros(b, 10)], -1)
    v_shaped = m.v_template + torch.einsum("bl,mkl->bmk", shape, m.shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, m.J_regressor)
    feat = (pose[:, 1:] - torch.eye(3)).view(b, -1)
    v_posed = v_shaped + torch.matmul(feat, m.posedirs).view(b, -1, 3)
    rel = J.clone()
    rel[:, 1:] -= J[:, m.parents[1:]]
    T = torch.cat([F.pad(pose.reshape(-1, 3, 3), [0, 0, 0, 1]), F.pad(rel.reshape(-1, 3, 1), [0, 0, 0, 1], value=1)], 2).view(b, -1, 4, 4)
    chain = [T[:, 0]]
    for i in range(1, len(m.parents)):                      # the 54 sequential matmuls of lbs.py:350-355
        chain.append(torch.matmul(chain[m.parents[i]], T[:, i]))
    G = torch.stack(chain, 1)
    Jt = G[:, :, :3, 3]
    A = G - F.pad(torch.matmul(G, F.pad(J.unsqueeze(-1), [0, 0, 0, 1])), [3, 0, 0, 0, 0, 0, 0, 0])
    Tv = torch.matmul(m.weights.unsqueeze(0).expand(b, -1, -1), A.view(b, -1, 16)).view(b, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(b, v_posed.shape[1], 1)], 2)
    verts = torch.matmul(Tv, vh.unsqueeze(-1))[:, :, :3, 0]
    lmk = torch.einsum("blfi,lf->bli", verts[:, m.faces[m.lmk_faces_idx]], m.lmk_bary)
    joints = torch.cat([Jt, verts[:, EXTRA], lmk], 1)
    return verts, joints


def twoview_forward(sd, m: Smplx, batch, iters=3, focal=(1475.0, 1475.0), feats=None):
    """copenet_twoview.py:164-317 without the loss.  `feats` = (xf0, xf1) skips the trunk (used when the trunk ran under
    autocast on the GPU baseline)."""
    b = batch["im0"].shape[0]
    init = torch.tensor([0.0, 0.0, 10.0]).expand(b, -1) * 0.05
    xf0, xf1 = feats if feats is not None else (forward_feat_ext(batch["im0"], sd), forward_feat_ext(batch["im1"], sd))
    p0, s0, p1, s1 = ief(sd, xf0, xf1, batch["bb0"], batch["bb1"], init, init, iters)
    out = {"xf0": xf0, "xf1": xf1}
    for v, (p, s) in enumerate(((p0, s0), (p1, s1))):
        p = p.clone()
        p[:, :3] /= 0.05
        R = rot6d_to_rotmat(p[:, 3:]).view(b, 22, 3, 3)
        verts, joints = smplx_forward(m, s, R[:, 1:])
        vc = torch.bmm(R[:, 0], verts.permute(0, 2, 1)).permute(0, 2, 1) + p[:, None, :3]          # utils.py:237-239
        jc = torch.bmm(R[:, 0], joints.permute(0, 2, 1)).permute(0, 2, 1) + p[:, None, :3]
        c = batch["intr%d" % v][:, :2, 2]
        proj = jc / jc[:, :, -1:]
        j2d = torch.stack([focal[0] * proj[:, :, 0] + c[:, None, 0], focal[1] * proj[:, :, 1] + c[:, None, 1]], -1)
        out.update({"pred_pose%d" % v: p, "pred_betas%d" % v: s, "pred_vertices_cam%d" % v: vc,
                    "pred_joints_cam%d" % v: jc, "pred_joints_2d_cam%d" % v: j2d})
    return out
