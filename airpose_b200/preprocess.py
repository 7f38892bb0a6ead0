"""Input preprocessing on the device -- the step immediately before the hot path.

``crop_resize_pad_normalize``  the dataset's per-camera image path (copenet/src/copenet/dsets/aerialpeople.py:125-141,174 with
                               utils.resize_with_pad, utils/utils.py:214-235): u8 BGR frame -> RGB / 255 -> crop ->
                               cv2.resize(INTER_LINEAR) to a longer side of 224 -> zero letterbox -> Normalize -> [n,3,224,224].
``bgr8_to_normalized``         the drone server's stage-0 conversion (airpose_server/server.py:93-98).

In the reference both run on the CPU (OpenCV in 30 DataLoader workers / numpy in the server loop); here they are one
HBM-streaming kernel each (``csrc/preprocess.cu``), so the u8 frame is the only thing that crosses PCIe.  CUDA only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # aerialpeople.py:68-69, server.py:75-76
IMAGENET_STD = (0.229, 0.224, 0.225)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def letterbox_geometry(height: int, width: int, size: int = 224):
    """``scale, (dst_w, dst_h), [pad_left, pad_top]`` of utils.resize_with_pad for a crop of ``height`` x ``width``
    (utils/utils.py:218-229) -- the numbers the dataset puts into ``bb[:, 2]`` and uses to map 2D joints into the crop
    (aerialpeople.py:172,200)."""
    bigger = height if height > width else width
    scale = size / bigger
    dst_w, dst_h = int(scale * width), int(scale * height)
    return scale, (dst_w, dst_h), [(size - dst_w) // 2, (size - dst_h) // 2]


def crop_resize_pad_normalize(frames, rects, size: int = 224, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """``frames``: u8 BGR on the device, [n,H,W,3] or one [H,W,3] frame shared by all crops; ``rects``: n x (y0, y1, x0, x1)
    (a list of tuples or an int tensor), the crop being ``frame[y0:y1, x0:x1]``.  Returns
    ``(images [n,3,size,size] float32, scales [n] list of float, pads [n] list of [pad_left, pad_top])``."""
    if frames.device.type != "cuda":
        raise _lib.AirposeError("airpose_b200.preprocess runs on CUDA only (frames are on {}); there is no CPU path".format(frames.device))
    if frames.dtype != torch.uint8 or frames.shape[-1] != 3 or frames.dim() not in (3, 4):
        raise ValueError("frames must be uint8 [n,H,W,3] or [H,W,3] (BGR), got {} {}".format(frames.dtype, tuple(frames.shape)))
    lib = _lib.load()
    frames = frames.contiguous()
    rects_host = rects.tolist() if torch.is_tensor(rects) else [tuple(int(v) for v in r) for r in rects]
    n = len(rects_host)
    H, W = int(frames.shape[-3]), int(frames.shape[-2])
    if frames.dim() == 4 and frames.shape[0] not in (1, n):
        raise ValueError("{} frames for {} crop rectangles".format(frames.shape[0], n))
    stride = H * W * 3 if (frames.dim() == 4 and frames.shape[0] == n and n > 1) else 0
    scales, pads = [], []
    for y0, y1, x0, x1 in rects_host:
        if not (0 <= y0 < y1 <= H and 0 <= x0 < x1 <= W):
            raise ValueError("crop rectangle {} outside the {}x{} frame".format((y0, y1, x0, x1), H, W))
        s, _, pad = letterbox_geometry(y1 - y0, x1 - x0, size)
        scales.append(s)
        pads.append(pad)
    rects_dev = torch.tensor(rects_host, dtype=torch.int32).to(frames.device, non_blocking=True)
    if out is None:
        out = torch.empty(n, 3, size, size, device=frames.device, dtype=torch.float32)
    with torch.cuda.device(frames.device):
        _lib.check(lib.airpose_preprocess_crop_resize(frames.data_ptr(), stride, H, W, rects_dev.data_ptr(), n, size, _f3(mean), _f3(std),
                                                      out.data_ptr(), _lib.current_stream()), "airpose_preprocess_crop_resize")
    return out, scales, pads


def bgr8_to_normalized(bgr, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None, size=None):
    """u8 BGR [n,S,S,3] (or a flat byte view of it) on the device -> float32 RGB [n,3,S,S], ``x * (1/255)`` then
    ``(x - mean) / std`` (server.py:93-98), bit-exact with the reference's torch ops."""
    if bgr.device.type != "cuda":
        raise _lib.AirposeError("airpose_b200.preprocess runs on CUDA only; there is no CPU path")
    if bgr.dtype != torch.uint8:
        raise ValueError("bgr must be uint8")
    lib = _lib.load()
    if bgr.dim() == 4:
        n, size = int(bgr.shape[0]), int(bgr.shape[1])
    elif bgr.dim() == 3:
        n, size = 1, int(bgr.shape[0])
    else:
        if size is None:
            raise ValueError("a flat byte buffer needs size=")
        n = bgr.numel() // (size * size * 3)
    if bgr.numel() != n * size * size * 3 or not bgr.is_contiguous():
        raise ValueError("bgr must be a contiguous [n,{0},{0},3] byte image".format(size))
    if out is None:
        out = torch.empty(n, 3, size, size, device=bgr.device, dtype=torch.float32)
    with torch.cuda.device(bgr.device):
        _lib.check(lib.airpose_preprocess_bgr8(bgr.data_ptr(), n, size, _f3(mean), _f3(std), out.data_ptr(), _lib.current_stream()),
                   "airpose_preprocess_bgr8")
    return out
