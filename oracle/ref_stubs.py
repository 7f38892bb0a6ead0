"""Import shims so the UNMODIFIED reference LightningModule can be imported on a box
without pytorch_lightning 1.2 / torchgeometry / pyrender / trimesh / imgaug.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py and the drop-in test that runs
the reference's ``copenet_twoview`` over our replacement modules).  The stubs implement
just the surface the reference touches on the hot path:
  pytorch_lightning.LightningModule  (copenet/src/copenet/copenet_twoview.py:50,59)
  torchgeometry.rotation_matrix_to_angle_axis (:323-326, test mode only -- not stubbed
      numerically: parity for that call is unpinned)
  pyrender.OffscreenRenderer / trimesh (utils/renderer.py:6-19)
  imgaug.augmenters (dsets/aerialpeople.py:17)
"""
from __future__ import annotations

import sys
import types
from argparse import Namespace


def install():
    import torch
    import torch.nn as nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def __init__(self):
                super().__init__()
                self.hparams = Namespace()

            def save_hyperparameters(self, hparams=None):
                if isinstance(hparams, dict):
                    hparams = Namespace(**hparams)
                self.hparams = hparams

            @property
            def device(self):
                for p in self.parameters():
                    return p.device
                return torch.device("cpu")

            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.Trainer = type("Trainer", (), {})
        pl.seed_everything = lambda s: torch.manual_seed(s)
        sys.modules["pytorch_lightning"] = pl

    if "torchgeometry" not in sys.modules:
        tgm = types.ModuleType("torchgeometry")

        def _unpinned(*a, **k):
            raise NotImplementedError("torchgeometry is not installed; test-mode angle-axis output is out of scope")

        tgm.rotation_matrix_to_angle_axis = _unpinned
        tgm.angle_axis_to_rotation_matrix = _unpinned
        sys.modules["torchgeometry"] = tgm

    if "pyrender" not in sys.modules:
        pr = types.ModuleType("pyrender")

        class OffscreenRenderer:
            def __init__(self, *a, **k):
                pass

        pr.OffscreenRenderer = OffscreenRenderer
        sys.modules["pyrender"] = pr
    if "trimesh" not in sys.modules:
        sys.modules["trimesh"] = types.ModuleType("trimesh")
    if "imgaug" not in sys.modules:
        ia = types.ModuleType("imgaug")
        iaa = types.ModuleType("imgaug.augmenters")
        ia.augmenters = iaa
        sys.modules["imgaug"] = ia
        sys.modules["imgaug.augmenters"] = iaa
