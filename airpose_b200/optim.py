"""Optimizer of the training configuration: ``torch.optim.Adam(model.parameters(), lr, weight_decay=0,
amsgrad=True)`` (copenet/src/copenet/copenet_twoview.py:416-425) as ONE kernel launch per step.

The parameters are re-homed into a single flat fp32 buffer (each ``p.data`` becomes a view of it, the
module keeps working unchanged), and so are the gradients (``p.grad`` views of a flat gradient buffer):
the optimizer step is one HBM-bound pass (csrc/optim.cu), ``zero_grad`` one memset, and the data-parallel
gradient mean one all-reduce of the flat buffer (``allreduce_grads``) instead of DDP's bucket machinery.
No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0:
            raise NotImplementedError("weight_decay != 0 is not used by the reference (copenet_twoview.py:421)")
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.AirposeError("airpose_b200.optim.Adam runs on CUDA only; there is no CPU path")
        if any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise ValueError("all parameters must be float32 on one CUDA device")
        self.lr, self.betas, self.eps, self.amsgrad = float(lr), (float(betas[0]), float(betas[1])), float(eps), bool(amsgrad)
        self.device = dev
        self.step_count = 0
        # 16-byte aligned slots so that every view can be used by float4 kernels
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.max_exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32) if amsgrad else None
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                g = self.flat_grad[o:o + p.numel()].view_as(p)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
        self.offsets = offs

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    def allreduce_grads(self, group=None):
        """Gradient mean over the data-parallel ranks: one collective on the flat buffer (NCCL over NVLink)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            return 1.0 / dist.get_world_size(group)
        return 1.0

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        lib = _lib.load()
        self.step_count += 1
        a = _lib.AdamArgs()
        a.param, a.grad = self.flat.data_ptr(), self.flat_grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.max_exp_avg_sq = self.max_exp_avg_sq.data_ptr() if self.amsgrad else None
        a.n = self.numel
        a.lr, a.beta1, a.beta2, a.eps = self.lr, self.betas[0], self.betas[1], self.eps
        a.step, a.grad_scale = self.step_count, float(grad_scale)
        with torch.cuda.device(self.device):
            _lib.check(lib.airpose_adam_step(C.byref(a), _lib.current_stream()), "airpose_adam_step")
        for p in self.params:                    # the kernel wrote through raw pointers: tell cached packings they are stale
            p._airpose_gen = self.step_count
