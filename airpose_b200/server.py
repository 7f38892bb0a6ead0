"""Staged, streaming inference of the drone server on the sm_100a kernels.

Mirrors the reference's on-board server (catkin_ws/src/aircap/packages/flight/airpose_server/server.py:69-150 and its model,
airpose_server/airpose.py:179-199): each drone runs the trunk on its own 224 x 224 crop (stage 0) and then three
single-view regressor passes, exchanging ``betas | articulated pose`` (136 floats) with the other drone between passes.

Wire format (server.py:38-39,91-98,110,125,140; airpose_client/AirPoseClient.h:20-31), little-endian, packed:
    request  stage 0     u8 stage | 3 x f32 bb (cx/c_x - 1, cy/c_y - 1, scale) | 224*224*3 u8 BGR          150541 bytes
    request  stage 1, 2  u8 stage | 10 x f32 betas | 126 x f32 articulated 6D pose of the OTHER view          545 bytes
    reply    stage 0, 1  10 x f32 betas | 126 x f32 articulated pose                                           544 bytes
    reply    stage 2     10 x f32 betas | 135 x f32 pose (3 position, 6 root, 126 articulated)                  580 bytes

``StagedServer.process(data, metainfo, stage)`` is the drop-in for the reference's ``process``: same arguments, same
reply bytes, same carried state (``xf``, ``bb``, ``curr_pose``, ``curr_shape``, ``curr_position``).  The TCP select loop
around it is networking and stays the reference's.  Everything between the request bytes and the reply bytes runs on the
device: one H2D copy of the message, ``airpose_preprocess_bgr8`` -> ``airpose_backbone_fwd`` (1 image) ->
``airpose_ief_fwd`` (one pass; only view 0's output is used) and one D2H copy of the reply.  With ``graph=True`` each
stage's device work is captured once into a CUDA graph and replayed per message (latency, not throughput, is what the
45 ms / 2.5 ms slots of airpose.yaml:9-11 budget).  CUDA only; no CPU path.
"""
from __future__ import annotations

import re
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .model_copenet import Bottleneck, copenet
from .preprocess import bgr8_to_normalized

SIZE = 224                                      # server.py:37
BUFFERSIZE = 1 + 3 * 4 + SIZE * SIZE * 3        # server.py:38
BUFFERSIZE_STAGES = 1 + (10 + 21 * 6) * 4       # server.py:39
REPLY_FLOATS = (136, 136, 145)


def fix_state_dict(old_state_dict):
    """server.py:16-22: Lightning checkpoints prefix every network key with ``model.``."""
    new_state_dict = OrderedDict()
    for k, v in old_state_dict.items():
        new_state_dict[re.sub(r"^model\.", "", k)] = v
    return new_state_dict


def getmodel(smpl_mean_params_path):
    """airpose_server/airpose.py:197-199.  The server model has the two-view network's parameters (same ``state_dict``) and a
    single-view ``forward_reg``; ``StagedServer`` calls the two-view module's kernels with view 1 unused."""
    return copenet(Bottleneck, [3, 4, 6, 3], smpl_mean_params_path)


class StagedServer:
    """One connection's network state + ``process`` (server.py:69-150)."""

    def __init__(self, model, device="cuda:0", graph=False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AirposeError("airpose_b200.server runs on CUDA only; there is no CPU path")
        _lib.load()
        self.model = model.to(self.device).eval()
        dev = self.device
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        # server.py:69-73
        self.xf, self.bb = f(1, 2048), f(1, 3)
        self.curr_theta = self.model.init_pose[:, :132].detach().clone().contiguous()     # orient(6) | articulated(126)
        self.curr_shape = self.model.init_shape.detach().clone().contiguous()
        self.curr_position = torch.tensor([[0.0, 0.0, 0.5]], device=dev)
        self._pos0 = torch.tensor([[0.0, 0.0, 0.5]], device=dev)                          # server.py:103
        # staging: pinned host buffers for the request and the reply, device copies of both
        self._req_host = torch.empty(BUFFERSIZE, dtype=torch.uint8).pin_memory()
        self._req_np = self._req_host.numpy()
        self._req_dev = torch.empty(BUFFERSIZE + 3, dtype=torch.uint8, device=dev)       # + 3: the f32 fields start at byte 4
        self._rep_host = torch.empty(145, dtype=torch.float32).pin_memory()
        self._rep_dev = f(145)
        self._other_theta = self.model.init_pose[:, :132].detach().clone().contiguous()  # [:, 6:] <- the other view's pose
        self._other_shape = self.model.init_shape.detach().clone().contiguous()
        self._frame = torch.empty(1, 3, SIZE, SIZE, device=dev, dtype=torch.float32)
        self._graphs = {} if graph else None
        self._stream = torch.cuda.Stream(device=dev)

    # -- device work of one stage, on the current stream, reading self._req_dev and writing self._rep_dev
    def _stage_device(self, stage):
        m = self.model
        # the request is staged at byte offset 3 so that its float fields (message offsets 1, 41) are 4-byte aligned
        payload = self._req_dev[4:]
        if stage == 0:
            self.bb.copy_(payload[:12].view(torch.float32).view(1, 3))
            bgr8_to_normalized(payload[12:12 + SIZE * SIZE * 3], out=self._frame, size=SIZE)
            self.xf.copy_(m.forward_feat_ext(self._frame))
            pos, theta0, shape0 = self._pos0, m.init_pose[:, :132], m.init_shape
            theta1, shape1 = m.init_pose[:, :132], m.init_shape
        else:
            other = payload[:136 * 4].view(torch.float32)
            self._other_shape.copy_(other[:10].view(1, 10))
            self._other_theta[:, 6:].copy_(other[10:].view(1, 126))
            pos, theta0, shape0 = self.curr_position, self.curr_theta, self.curr_shape
            theta1, shape1 = self._other_theta, self._other_shape
        # one regressor pass; view 1 is fed view 0's features (its output is discarded): forward_reg of airpose.py:179-195
        pose, shape, _, _ = m._ief(self.xf, self.xf, self.bb, self.bb, pos, pos, theta0, theta1, shape0, shape1, 1)
        self._rep_dev[:10].copy_(shape[0])
        if stage == 2:
            self._rep_dev[10:145].copy_(pose[0])
        else:
            self._rep_dev[10:136].copy_(pose[0, 9:])
        # server.py:144-146
        self.curr_position.copy_(pose[:, :3])
        self.curr_theta.copy_(pose[:, 3:])
        self.curr_shape.copy_(shape)

    def _run_stage(self, stage):
        if self._graphs is None:
            self._stage_device(stage)
            return
        g = self._graphs.get(stage)
        if g is None:
            # warm up eagerly (weight packing, launch plans and workspaces are created on first use), restore the carried
            # state, then capture; the capture itself does not execute, so the replay below is this message's only pass
            # The warm-up runs on the capture stream: the stream-K GEMMs keep one partial-tile workspace per stream.
            cur = torch.cuda.current_stream()
            saved = [t.clone() for t in (self.xf, self.bb, self.curr_position, self.curr_theta, self.curr_shape)]
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                self._stage_device(stage)
                for t, s in zip((self.xf, self.bb, self.curr_position, self.curr_theta, self.curr_shape), saved):
                    t.copy_(s)
            self._stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._stage_device(stage)
            self._graphs[stage] = g
        g.replay()

    def process(self, data, metainfo=None, stage=None):
        """server.py:78-150.  ``data``: bytes-like request (at least BUFFERSIZE / BUFFERSIZE_STAGES bytes, stage byte first);
        returns a memoryview over the float32 reply, as the reference does."""
        if stage is None:
            stage = int(np.frombuffer(data, dtype=np.uint8, count=1)[0])
        stage = int(stage)
        if stage not in (0, 1, 2):
            raise ValueError("Invalid stage number {}".format(stage))          # server.py:141-142 (prints, then fails on `pose`)
        need = BUFFERSIZE if stage == 0 else BUFFERSIZE_STAGES
        if len(data) < need:
            raise ValueError("stage {} needs {} bytes, got {}".format(stage, need, len(data)))
        self._req_np[:need] = np.frombuffer(data, dtype=np.uint8, count=need)
        n_out = REPLY_FLOATS[stage]
        with torch.cuda.device(self.device):
            self._req_dev[3:3 + need].copy_(self._req_host[:need], non_blocking=True)
            self._run_stage(stage)
            self._rep_host[:n_out].copy_(self._rep_dev[:n_out], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return memoryview(self._rep_host[:n_out].numpy().copy())

    # the reference's globals, for callers that inspect them
    @property
    def curr_pose(self):
        return self.curr_theta
