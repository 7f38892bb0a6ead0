"""Time the REAL reference (unmodified copenet_twoview LightningModule from /root/reference, import shims from ref_stubs.py) next to the
PyTorch-CPU port that bench.py times as its reference arm -- build container only (the reference does not travel).
TEST INFRASTRUCTURE.   python oracle/time_reference_vs_port.py"""
import os, sys, time, tempfile
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from airpose_b200 import synthetic
import ref_stubs, gen_golden
torch.set_num_threads(os.cpu_count())
ref_stubs.install()
sys.path.insert(0, "/root/reference/copenet/src")
import copenet.config as ref_config
ref_config.device = "cpu"
import torchvision.models.resnet as tv_resnet
_orig = tv_resnet.resnet50
tv_resnet.resnet50 = lambda pretrained=False, **k: _orig(weights=None)
from copenet import copenet_twoview as ref_twoview
from argparse import Namespace
tmp = gen_golden.make_home(tempfile.mkdtemp(prefix="airpose_home_"))
B = 8
hp = Namespace(copenet_home=tmp, batch_size=B, val_batch_size=B, testdata="aerialpeople", smpltrans_noise_sigma=None, reg_iters=3,
               shape_loss_weight=50, keypoint2d_loss_weight=0.002, keypoint3d_loss_weight=1, limbs3d_loss_weight=3.0,
               limbstheta_loss_weight=1.0, trans_loss_weight=10, rootrot_loss_weight=1, pose_loss_weight=50, beta_loss_weight=1)
module = ref_twoview.copenet_twoview(hp)
module.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synthetic.make_network_state(123).items()}, strict=True)
module.eval()
x = synthetic.make_inputs(B, 123)
li = synthetic.make_lbs_inputs(B, seed=9)
gt = {"smplpose_rotmat": li["body_pose"], "smplorient_rel0": np.tile(np.eye(3, dtype=np.float32), (B, 1, 1, 1)), "smplorient_rel1": np.tile(np.eye(3, dtype=np.float32), (B, 1, 1, 1)),
      "smpl_vertices": np.zeros((B, 1, 10475, 3), np.float32), "smpl_joints": np.zeros((B, 1, 127, 3), np.float32),
      "smpl_joints_2d0": np.zeros((B, 1, 127, 2), np.float32), "smpl_joints_2d1": np.zeros((B, 1, 127, 2), np.float32)}
batch = {k: torch.from_numpy(v) for k, v in {**x, **gt}.items()}
def run_ref():
    with torch.no_grad():
        module.fwd_pass_and_loss(batch, is_val=True, is_test=False)
import torch_port as tp
sd = tp.to_torch(synthetic.make_network_state(123)); m = tp.Smplx(synthetic.make_smplx_model(0))
xt = {k: torch.from_numpy(v) for k, v in x.items()}
def run_port():
    with torch.no_grad():
        tp.twoview_forward(sd, m, xt)
for name, fn in (("reference LightningModule.fwd_pass_and_loss", run_ref), ("port twoview_forward", run_port), ("reference", run_ref), ("port", run_port)):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(4): fn()
    dt = (time.perf_counter() - t0) / 4
    print("%-45s %.1f ms/step  %.1f pairs/s (B=%d, %d threads)" % (name, dt * 1e3, B / dt, B, torch.get_num_threads()))
