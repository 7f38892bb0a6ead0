"""Generate tests/golden/trainmode_b2.npz: the reference network in train() mode (model_copenet.py:112-204 with BatchNorm batch
statistics and ACTIVE dropout) run on CPU in the build container, with the dropout masks it drew recorded by forward hooks.

TEST INFRASTRUCTURE ONLY.   python oracle/gen_golden_trainmode.py

Pins the semantics the training-mode regressor kernels (csrc/ief_train.cu) and their PyTorch port (oracle/torch_port.py:ief_train)
implement: which Dropout call belongs to which (iteration, view), the 1/(1-p) scaling, the state update between iterations.
Stored: the features the train-mode trunk produced (fp32, batch statistics), the masks as kept-bits, the four outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from airpose_b200 import synthetic  # noqa: E402
import gen_golden  # noqa: E402
import ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NET_SEED, IN_SEED = 123, 55


def main():
    import tempfile
    import torch
    torch.set_num_threads(os.cpu_count())
    ref_stubs.install()
    sys.path.insert(0, gen_golden.REF_SRC)
    import torchvision.models.resnet as tv_resnet
    _orig = tv_resnet.resnet50
    tv_resnet.resnet50 = lambda pretrained=False, **k: _orig(weights=None)
    from copenet.models import model_copenet as ref_model

    B, iters = 2, 3
    mp = synthetic.write_mean_params(os.path.join(tempfile.mkdtemp(), "smpl_mean_params.npz"))
    net = ref_model.getcopenet(mp, pretrained=False)
    sd = synthetic.make_network_state(NET_SEED, dec_gain=0.01)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    net.train()
    masks = {"drop1": [], "drop2": []}

    def hook(name):
        def fn(mod, inp, out):
            x = inp[0]
            masks[name].append(torch.where(x != 0, out / x, torch.full_like(out, 2.0) * (out != 0)).detach().clone())
        return fn

    net.drop1.register_forward_hook(hook("drop1"))
    net.drop2.register_forward_hook(hook("drop2"))
    x = {k: torch.from_numpy(v) for k, v in synthetic.make_inputs(B, IN_SEED).items()}
    init = torch.tensor([0.0, 0.0, 10.0]).expand(B, -1).clone() * 0.05
    torch.manual_seed(7)
    feats = []
    net.avgpool.register_forward_hook(lambda m, i, o: feats.append(o.detach().view(o.size(0), -1).clone()))
    with torch.no_grad():
        p0, s0, p1, s1 = net(x0=x["im0"], x1=x["im1"], bb0=x["bb0"], bb1=x["bb1"], init_position0=init, init_position1=init.clone(), iters=iters)
    assert len(masks["drop1"]) == 2 * iters and len(feats) == 2
    m1 = torch.stack(masks["drop1"]).view(iters, 2, B, 1024)          # call order: iteration-major, view 0 then view 1 (:186-196)
    m2 = torch.stack(masks["drop2"]).view(iters, 2, B, 1024)
    assert set(torch.unique(m1).tolist()) <= {0.0, 2.0} and set(torch.unique(m2).tolist()) <= {0.0, 2.0}
    np.savez_compressed(os.path.join(GOLDEN, "trainmode_b2.npz"), batch=B, iters=iters, net_seed=NET_SEED, in_seed=IN_SEED, dec_gain=0.01,
                        xf0=feats[0].numpy(), xf1=feats[1].numpy(), kept1=(m1 != 0).numpy(), kept2=(m2 != 0).numpy(),
                        pred_pose0=p0.numpy(), pred_betas0=s0.numpy(), pred_pose1=p1.numpy(), pred_betas1=s1.numpy())
    print("trainmode_b2.npz: kept fraction %.3f / %.3f" % (float((m1 != 0).float().mean()), float((m2 != 0).float().mean())))


if __name__ == "__main__":
    main()
