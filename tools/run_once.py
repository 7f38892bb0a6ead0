"""Run one piece of the hot path a few times (target for ncu captures).
    python tools/run_once.py trunk 32 | lbs 8192 | twoview 64 | ief 64     [iters]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.gpu_probe import setup_net

what, n = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
tmp = tempfile.mkdtemp(prefix="airpose_once_")
if what == "trunk":
    net = setup_net(tmp)
    x = torch.randn(n, 3, 224, 224, device="cuda")
    fn = lambda: net.forward_feat_ext(x)
elif what == "ief":
    net = setup_net(tmp)
    a, b, c = torch.rand(n, 2048, device="cuda"), torch.rand(n, 3, device="cuda"), torch.rand(n, 3, device="cuda")
    fn = lambda: net._ief(a, a, b, b, c, c, None, None, None, None, 3)
elif what == "lbs":
    from airpose_b200 import synthetic
    from airpose_b200.smplx import SMPLX
    synthetic.write_smplx_model(tmp, 0)
    sm = SMPLX(tmp, batch_size=1, create_transl=False).cuda()
    li = synthetic.make_lbs_inputs(n, seed=1)
    betas, body = torch.from_numpy(li["betas"]).cuda(), torch.from_numpy(li["body_pose"]).cuda()
    fn = lambda: sm.forward(betas=betas, body_pose=body, pose2rot=False)
else:
    raise SystemExit("unknown target")
for _ in range(iters):
    fn()
torch.cuda.synchronize()
print("done", what, n)
