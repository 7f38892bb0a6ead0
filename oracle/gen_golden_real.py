"""Generate tests/golden/real_loss.npz from the REAL reference, on CPU (VPoser-free part).

TEST INFRASTRUCTURE ONLY.  Run in the build container:   python oracle/gen_golden_real.py

The UNMODIFIED ``copenet_twoview.get_loss`` of /root/reference/copenet_real/src/copenet_real/copenet_twoview.py:99-160 is
extracted with ``ast`` (the module itself imports human_body_prior / VPoser weights at import time, which do not exist
offline) and executed on seeded predictions and ground truth with
  * ``vp_model.encode(x).rsample()`` stubbed to zeros -- the VPoser prior needs external weights and is OUT of scope; its
    term is then exactly 0 whatever ``vposer_loss_weight`` is, and every other term is the reference's own arithmetic;
  * ``tgm.rotation_matrix_to_angle_axis`` stubbed (its result only feeds the stubbed encoder).
Stored: the inputs, ``loss``, the ``losses`` dict and, through the reference's own autograd, d loss / d every prediction.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/copenet_real/src/copenet_real/copenet_twoview.py"
GOLDEN = os.path.join(ROOT, "tests", "golden")
HP = dict(limbs2d_loss_weight=3.0, keypoint2d_loss_weight=0.002, beta_loss_weight=1.0, vposer_loss_weight=0.01, pose_loss_weight=50.0)


def reference_get_loss():
    tree = ast.parse(open(REF).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "copenet_twoview"][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "get_loss"]
    assert len(fn) == 1

    class _Q:
        def __init__(self, n):
            self.n = n

        def rsample(self):
            return torch.zeros(self.n, 32)

    ns = {"torch": torch, "np": np,
          "tgm": types.SimpleNamespace(rotation_matrix_to_angle_axis=lambda m: torch.zeros(m.shape[0], 3)),
          "vp_model": types.SimpleNamespace(encode=lambda x: _Q(x.shape[0]))}
    exec(compile(ast.Module(body=fn, type_ignores=[]), REF, "exec"), ns)
    return ns["get_loss"]


def make_case(B, seed, J=127):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    rot = lambda n: np.linalg.qr(rng.standard_normal((n, 3, 3)))[0].astype(np.float32)
    gt2d = lambda: np.concatenate([f(B, 1, 25, 2) * 200 + 900, rng.uniform(0, 1, (B, 1, 25, 1)).astype(np.float32)], -1)
    return {"pred_smpltrans0": f(B, 3) + np.array([0, 0, 6], np.float32), "pred_smpltrans1": f(B, 3) + np.array([0, 0, 7], np.float32),
            "pred_rotmat0": rot(B * 22).reshape(B, 22, 3, 3), "pred_rotmat1": rot(B * 22).reshape(B, 22, 3, 3),
            "pred_betas0": f(B, 10), "pred_betas1": f(B, 10),
            "pred_joints_2d_cam0": f(B, J, 2) * 200 + 900, "pred_joints_2d_cam1": f(B, J, 2) * 200 + 900,
            "smpl_joints_2d0": gt2d(), "smpl_joints_2d1": gt2d()}


def main():
    get_loss = reference_get_loss()
    self = types.SimpleNamespace(mseloss=torch.nn.MSELoss(reduction="none"), hparams=types.SimpleNamespace(**HP))
    out = {"hp/" + k: np.float32(v) for k, v in HP.items()}
    for B, seed in ((1, 3), (6, 4)):
        case = make_case(B, seed)
        t = {k: torch.from_numpy(v.copy()) for k, v in case.items()}
        preds = ["pred_smpltrans0", "pred_smpltrans1", "pred_rotmat0", "pred_rotmat1", "pred_betas0", "pred_betas1",
                 "pred_joints_2d_cam0", "pred_joints_2d_cam1"]
        for k in preds:
            t[k].requires_grad_(True)
        batch = {"smpl_joints_2d0": t["smpl_joints_2d0"], "smpl_joints_2d1": t["smpl_joints_2d1"]}
        loss, losses = get_loss(self, batch, t["pred_smpltrans0"], t["pred_smpltrans1"], t["pred_rotmat0"], t["pred_rotmat1"],
                                t["pred_betas0"], t["pred_betas1"], None, None, t["pred_joints_2d_cam0"], t["pred_joints_2d_cam1"])
        loss.backward()
        p = "b%d/" % B
        for k, v in case.items():
            out[p + k] = v
        out[p + "loss"] = np.float32(loss.item())
        for k, v in losses.items():
            out[p + "losses/" + k] = np.float32(v)
        for k in preds:
            out[p + "grad/" + k] = t[k].grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "real_loss.npz"), **out)
    print("wrote real_loss.npz:", {k: float(v) for k, v in out.items() if k.endswith("/loss")})


if __name__ == "__main__":
    main()
