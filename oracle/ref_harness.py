"""Run the UNMODIFIED reference (`copenet_twoview` LightningModule and its modules) on synthetic assets.

TEST / BASELINE INFRASTRUCTURE ONLY.  Sources come from /root/reference when it exists (build container) and from
oracle/_ref (oracle/make_ref.py) otherwise (GPU box).  Import shims for pytorch_lightning / torchgeometry / pyrender /
trimesh / imgaug come from ref_stubs.py.  Nothing under airpose_b200/ imports this.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile
from argparse import Namespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from airpose_b200 import synthetic  # noqa: E402
import ref_stubs  # noqa: E402

HPARAMS = dict(testdata="aerialpeople", smpltrans_noise_sigma=None, reg_iters=3, shape_loss_weight=50, keypoint2d_loss_weight=0.002,
               keypoint3d_loss_weight=1, limbs3d_loss_weight=3.0, limbstheta_loss_weight=1.0, trans_loss_weight=10,
               rootrot_loss_weight=1, pose_loss_weight=50, beta_loss_weight=1, lr=5e-5, summary_steps=500, val_summary_steps=50)


def ref_src():
    """(path to put on sys.path, kind): the real tree in the build container, the verbatim copy elsewhere."""
    real = "/root/reference/copenet/src"
    if os.path.isdir(real) and not os.environ.get("AIRPOSE_REF_FORCE_COPY"):      # the env var makes this container behave like the GPU box
        return real, "reference"
    copy = os.path.join(HERE, "_ref", "copenet", "src")
    if os.path.isdir(copy):
        return copy, "reference"
    return None, None


def available():
    return ref_src()[0] is not None


def make_home(tmp=None, smplx_seed=0):
    """A synthetic `copenet_home` (copenet_twoview.py:60-68 path conventions): mean params, SMPL-X npz, hand / face id files."""
    tmp = tmp or tempfile.mkdtemp(prefix="airpose_home_")
    data = os.path.join(tmp, "src", "copenet", "data")
    synthetic.write_mean_params(os.path.join(data, "smpl_mean_params.npz"))
    synthetic.write_smplx_model(os.path.join(data, "smplx", "models", "smplx"), smplx_seed)
    rng = np.random.default_rng(5)
    with open(os.path.join(data, "smplx", "MANO_SMPLX_vertex_ids.pkl"), "wb") as f:
        pickle.dump({"left_hand": rng.choice(synthetic.NUM_VERTS, 778, replace=False),
                     "right_hand": rng.choice(synthetic.NUM_VERTS, 778, replace=False)}, f)
    np.save(os.path.join(data, "smplx", "SMPL-X__FLAME_vertex_ids.npy"), rng.choice(synthetic.NUM_VERTS, 5023, replace=False))
    return tmp


def import_reference(device="cpu", inject=None):
    """Import the reference's `copenet.copenet_twoview` module.  `inject` maps module names inside the reference package to
    replacement modules BEFORE the import (the drop-in test passes airpose_b200's model_copenet / smplx there, which is
    exactly the import swap INTEGRATION.md section 1 describes)."""
    src, _ = ref_src()
    if src is None:
        raise RuntimeError("the reference is neither at /root/reference nor in oracle/_ref (run oracle/make_ref.py in the build container)")
    ref_stubs.install()
    for k in [k for k in sys.modules if k == "copenet" or k.startswith("copenet.")]:
        del sys.modules[k]
    if src not in sys.path:
        sys.path.insert(0, src)
    import copenet.config as ref_config
    ref_config.device = device                    # the reference hard-codes "cuda" (config.py:7)
    import torchvision.models.resnet as tv_resnet
    if not getattr(tv_resnet.resnet50, "_airpose_no_download", False):
        orig = tv_resnet.resnet50

        def resnet50(pretrained=False, **k):      # model_copenet.py:236-238 downloads ImageNet weights: no network here
            return orig(weights=None)
        resnet50._airpose_no_download = True
        tv_resnet.resnet50 = resnet50
    for name, mod in (inject or {}).items():
        sys.modules[name] = mod
    from copenet import copenet_twoview as ref_twoview
    return ref_twoview


def make_module(ref_twoview, pairs, home=None, net_seed=123, device="cpu", load_weights=True):
    import torch
    home = home or make_home()
    hp = Namespace(copenet_home=home, batch_size=pairs, val_batch_size=pairs, **HPARAMS)
    module = ref_twoview.copenet_twoview(hp)
    if load_weights:
        sd = synthetic.make_network_state(net_seed)
        module.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return module.to(device)


def make_batch(pairs, in_seed=123, gt_seed=321, device="cpu"):
    """Inputs + a ground truth for the loss: the SMPL-X forward (numpy oracle) of an independent seeded sample."""
    import torch
    import airpose_oracle as orc
    x = synthetic.make_inputs(pairs, in_seed)
    gt_in = synthetic.make_lbs_inputs(pairs, seed=gt_seed)
    rng = np.random.default_rng(gt_seed)
    orient = [synthetic.rot6d_to_rotmat_np(np.array([1, 0, 0, 1, 0, 0], np.float32) + rng.standard_normal((pairs, 6)).astype(np.float32) * 0.3)
              for _ in (0, 1)]
    model = orc.SmplxModel(synthetic.make_smplx_model(0))
    verts, joints = orc.smplx_forward(model, gt_in["betas"], gt_in["body_pose"], transl=np.zeros((pairs, 3), np.float32))
    gt = {"smplpose_rotmat": gt_in["body_pose"], "smplorient_rel0": orient[0][:, None], "smplorient_rel1": orient[1][:, None],
          "smpl_vertices": verts[:, None].astype(np.float32), "smpl_joints": joints[:, None].astype(np.float32)}
    for v in (0, 1):
        jc = np.einsum("bij,bkj->bki", orient[v], joints) + x["smpltrans_rel%d" % v][:, None]
        c = x["intr%d" % v][:, :2, 2]
        gt["smpl_joints_2d%d" % v] = (1475.0 * jc[..., :2] / jc[..., 2:] + c[:, None])[:, None].astype(np.float32)
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in {**x, **gt}.items()}
