"""Optimizer of the training configuration: ``torch.optim.Adam(model.parameters(), lr, weight_decay=0,
amsgrad=True)`` (copenet/src/copenet/copenet_twoview.py:416-425) as ONE kernel launch per step.

The parameters are re-homed into a single flat fp32 buffer (each ``p.data`` becomes a view of it, the
module keeps working unchanged), and so are the gradients (``p.grad`` views of a flat gradient buffer):
the optimizer step is one HBM-bound pass (csrc/optim.cu), ``zero_grad`` one memset, and the data-parallel
gradient mean one all-reduce of the flat buffer (``allreduce_grads``), or two slices of it overlapped with the backward
(``late_split`` / ``allreduce_begin``), instead of DDP's bucket machinery.
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
        # torch.optim.Optimizer-compatible view of the hyper-parameters (schedulers / Lightning read and write ``lr`` here)
        self.param_groups = [{"params": self.params, "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0,
                              "amsgrad": self.amsgrad}]
        self.sync_from_rank0()

    def sync_from_rank0(self, group=None):
        """What DistributedDataParallel does at construction: every rank starts from rank 0's parameters (one broadcast of
        the flat buffer).  Module BUFFERS (BatchNorm running statistics) are the caller's:
        ``airpose_b200.parallel.broadcast_buffers_(module)``."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)

    def zero_grad(self, set_to_none=False):
        """One memset of the flat gradient buffer.  ``set_to_none`` is accepted for signature compatibility and ignored: the
        gradients must stay views of the flat buffer (the backward kernels write straight into it)."""
        self._reattach_grads()
        self.flat_grad.zero_()

    def _reattach_grads(self):
        """``p.grad`` must be the parameter's slot of the flat gradient buffer.  Code outside this class can rebind it
        (``nn.Module.zero_grad()`` sets grads to None by default; autograd then allocates fresh tensors): a fresh gradient is
        copied into its slot and the view restored; a None gradient counts as zeros -- never silently skipped."""
        for p, o in zip(self.params, self.offsets):
            slot = self.flat_grad[o:o + p.numel()]
            g = p.grad
            if g is not None and g.data_ptr() == slot.data_ptr():
                continue
            with torch.no_grad():
                if g is None:
                    slot.zero_()
                else:
                    if g.shape != p.shape or g.device != self.device:
                        raise _lib.AirposeError("airpose_b200.optim.Adam: a parameter's .grad was replaced by a tensor of another shape or device")
                    slot.copy_(g.reshape(-1).to(torch.float32))
            p.grad = slot.view_as(p)

    def state_dict(self):
        """``torch.optim.Adam.state_dict()`` layout: per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` / ``max_exp_avg_sq``
        (clones, in parameter order) + ``param_groups`` with parameter indices -- loadable by ``torch.optim.Adam`` over the same
        parameters and vice versa."""
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            sl = slice(o, o + p.numel())
            st = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[sl].view_as(p).clone(),
                  "exp_avg_sq": self.exp_avg_sq[sl].view_as(p).clone()}
            if self.amsgrad:
                st["max_exp_avg_sq"] = self.max_exp_avg_sq[sl].view_as(p).clone()
            state[i] = st
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("optimizer state does not match: expected one group of {} parameters".format(len(self.params)))
        g = groups[0]
        if bool(g.get("amsgrad", self.amsgrad)) != self.amsgrad:
            raise ValueError("optimizer state was saved with amsgrad={}".format(g.get("amsgrad")))
        self.lr, self.betas, self.eps = float(g["lr"]), (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"])
        self.param_groups[0].update(lr=self.lr, betas=self.betas, eps=self.eps)
        steps = set()
        with torch.no_grad():
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                st = sd["state"].get(i, sd["state"].get(str(i)))
                sl = slice(o, o + p.numel())
                if st is None:                       # a parameter that never received a gradient (torch keeps no state for it)
                    self.exp_avg[sl].zero_(); self.exp_avg_sq[sl].zero_()
                    if self.amsgrad:
                        self.max_exp_avg_sq[sl].zero_()
                    continue
                self.exp_avg[sl].copy_(st["exp_avg"].reshape(-1)); self.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1))
                if self.amsgrad:
                    self.max_exp_avg_sq[sl].copy_(st["max_exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ ({}): the flat-buffer optimizer keeps one".format(sorted(steps)))
        self.step_count = steps.pop() if steps else 0

    def allreduce_grads(self, group=None):
        """Gradient mean over the data-parallel ranks: one collective on the flat buffer (NCCL over NVLink)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self._reattach_grads()
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            return 1.0 / dist.get_world_size(group)
        return 1.0

    def late_split(self, late_params):
        """Offset (in elements of the flat buffers) after the last of ``late_params``: everything from there on belongs to other
        parameters.  The overlapped all-reduce of ``copenet_twoview.training_step`` reduces ``flat_grad[offset:]`` as soon as
        those gradients are final and ``flat_grad[:offset]`` at the end of the backward."""
        from .parallel import late_split_offset
        late = {id(p) for p in late_params}
        return late_split_offset(self.offsets, [p.numel() for p in self.params], [id(p) in late for p in self.params])

    def allreduce_begin(self, lo, hi=None, group=None):
        """Asynchronous SUM all-reduce of ``flat_grad[lo:hi]`` (ordered after the work already enqueued on the current stream);
        returns the work handle (``.wait()`` orders the current stream after it), or None when there is nothing to reduce."""
        from .parallel import allreduce_begin
        return allreduce_begin(self.flat_grad, lo, self.numel if hi is None else hi, group=group)

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        lib = _lib.load()
        self._reattach_grads()
        self.lr = float(self.param_groups[0]["lr"])          # a scheduler may have changed it
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
