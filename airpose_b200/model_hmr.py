"""Drop-in ``model_hmr.copenet`` (the single-view HMR baseline: ResNet-50 trunk + IEF regressor with a
weak-perspective camera) on the same sm_100a kernels.

Mirrors copenet/src/copenet/models/model_hmr.py of the reference (BASELINE.json configs[0]):

    model = getcopenet(smpl_mean_params_path, pretrained=False)
    pred_rotmat, pred_betas, pred_cam = model(x, iters=3)          # :112-141

``fc1`` takes 2048 + 132 + 10 + 3 inputs (:66), ``decpose`` decodes 22 x 6 numbers, ``deccam`` IS used
(:71,:170), and the 6D pose is converted to rotation matrices inside the model (:140).  State-dict keys
are the reference's.  Eval mode only; no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import model_copenet as _mc
from .smplx import rot6d_to_rotmat

Bottleneck = _mc.Bottleneck


class copenet(_mc.copenet):
    FC1_EXTRA = 22 * 6 + 10 + 3         # model_hmr.py:66
    NPOSE_OUT = 22 * 6                   # :69

    def _weight_tensors(self):
        return super()._weight_tensors() + [self.deccam.weight, self.deccam.bias, self.init_cam]

    def _load_native(self, lib):
        p = self._fill_common(_lib.HmrParams())
        p.deccam_w, p.deccam_b = self.deccam.weight.data_ptr(), self.deccam.bias.data_ptr()
        p.init_cam = self.init_cam.data_ptr()
        _lib.check(lib.airpose_hmr_load(self._handle, C.byref(p), _lib.current_stream()), "airpose_hmr_load")

    def forward_feat_ext_pair(self, x0, x1):
        raise NotImplementedError("model_hmr.copenet is single-view")

    def _ief(self, xf, theta, shape, cam, iters):
        device = self.conv1.weight.device
        f = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
        xf, theta, shape, cam = map(f, (xf, theta, shape, cam))
        B = xf.shape[0]
        lib, h = self._ensure(0, device)
        pose = torch.empty(B, 132, device=device, dtype=torch.float32)
        betas = torch.empty(B, 10, device=device, dtype=torch.float32)
        pcam = torch.empty(B, 3, device=device, dtype=torch.float32)
        a = _lib.HmrIefArgs()
        a.batch, a.iters, a.xf = B, int(iters), xf.data_ptr()
        keep = []
        for name, t, width in (("init_theta", theta, 132), ("init_shape", shape, 10), ("init_cam", cam, 3)):
            if t is None:
                continue
            if t.shape[0] != B:
                t = t.expand(B, -1).contiguous()
            if t.shape[1] < width:
                raise ValueError("{} needs at least {} columns".format(name, width))
            keep.append(t)
            setattr(a, name, t.data_ptr())
            setattr(a, name + "_stride", t.stride(0))
        a.out_pose, a.out_betas, a.out_cam = pose.data_ptr(), betas.data_ptr(), pcam.data_ptr()
        with torch.cuda.device(device):
            _lib.check(lib.airpose_hmr_ief_fwd(h, C.byref(a), _lib.current_stream()), "airpose_hmr_ief_fwd")
        del keep
        return pose, betas, pcam

    def forward_reg(self, xf, pred_pose, pred_shape, pred_cam):
        """One regressor pass (model_hmr.py:160-172), eval mode."""
        return self._ief(xf, pred_pose, pred_shape, pred_cam, 1)

    def forward(self, x, init_cam=None, init_theta=None, init_shape=None, iters=3):
        """model_hmr.py:112-141."""
        B = x.shape[0]
        xf = self.forward_feat_ext(x)
        pose, betas, cam = self._ief(xf, init_theta, init_shape, init_cam, iters)
        return rot6d_to_rotmat(pose).view(B, 22, 3, 3), betas, cam


def getcopenet(smpl_mean_params, pretrained=True, **kwargs):
    """model_hmr.getcopenet (:196-206)."""
    model = copenet(Bottleneck, [3, 4, 6, 3], smpl_mean_params, **kwargs)
    if pretrained:
        import torchvision.models.resnet as resnet
        model.load_state_dict(resnet.resnet50(pretrained=True).state_dict(), strict=False)
    return model
