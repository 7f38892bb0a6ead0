"""The reference algorithm (oracle/torch_port.py = the ATen ops the reference modules call) on ONE B200 through stock PyTorch
(cuDNN / cuBLAS): the denominator of north_star's ">= 10x the reference single-GPU PyTorch forward" target.

    python tools/torch_gpu_baseline.py [pairs] [steps]

Prints one JSON line: pairs/s for fp32 (TF32 off, the reference's own precision), fp32 with TF32 allowed, and
autocast(bfloat16) + channels_last weights/inputs (the fastest stock configuration).  CUDA-event timed, eval, no_grad.
Test / baseline infrastructure: nothing under airpose_b200/ imports it.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def time_variant(tp, sd, m, x, steps, warmup, autocast, channels_last):
    import torch
    if channels_last:
        sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
        x = dict(x)
        for k in ("im0", "im1"):
            x[k] = x[k].contiguous(memory_format=torch.channels_last)

    def step():
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                xf0, xf1 = tp.forward_feat_ext(x["im0"], sd), tp.forward_feat_ext(x["im1"], sd)
            return tp.twoview_forward(sd, m, x, feats=(xf0.float(), xf1.float()))
        return tp.twoview_forward(sd, m, x)

    with torch.no_grad():
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def measure(pairs=64, steps=10, warmup=3):
    import torch
    import torch_port as tp
    from airpose_b200 import synthetic
    dev = torch.device("cuda:0")
    torch.set_default_device(dev)          # torch_port builds its constants (eye, zeros) on the default device
    sd = {k: v.to(dev) for k, v in tp.to_torch(synthetic.make_network_state(123)).items()}
    m = tp.Smplx(synthetic.make_smplx_model(0))
    for k, v in list(vars(m).items()):
        if torch.is_tensor(v):
            setattr(m, k, v.to(dev))
    x = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.make_inputs(pairs, 123).items()}
    out = {"pairs": pairs, "steps": steps, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    try:
        torch.backends.cudnn.benchmark = True
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ms = time_variant(tp, sd, m, x, steps, warmup, False, False)
        out["fp32"] = {"ms_per_step": ms, "pairs_per_s": pairs / ms * 1e3}
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        ms = time_variant(tp, sd, m, x, steps, warmup, False, False)
        out["tf32"] = {"ms_per_step": ms, "pairs_per_s": pairs / ms * 1e3}
        ms = time_variant(tp, sd, m, x, steps, warmup, True, True)
        out["bf16_autocast_channels_last"] = {"ms_per_step": ms, "pairs_per_s": pairs / ms * 1e3}
    finally:
        torch.set_default_device("cpu")
    return out


if __name__ == "__main__":
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print(json.dumps(measure(pairs, steps)))
