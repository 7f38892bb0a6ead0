"""Recipe for oracle/_ref: the reference's OWN hot-path sources, copied verbatim from /root/reference at build time.

TEST / BASELINE INFRASTRUCTURE ONLY.  `/root/reference` exists only in the build container; the GPU box receives the repo
snapshot, so the copy in `oracle/_ref/` (git-ignored: no reference source enters this repo's history; NOT gpurun-ignored:
it travels) is what lets the GPU-box tests and the CPU arm of bench.py run the reference ITSELF:
  * bench.py `--impl reference` / `cpu_baseline`  -> the unmodified `copenet_twoview` LightningModule, kind "reference"
  * tests/test_gpu_dropin.py                      -> the unmodified LightningModule driven over airpose_b200's objects
  * tests/test_gpu_parity.py gradient check       -> fp32 `loss.backward()` through the reference modules
Run by `__graft_entry__.build()` whenever /root/reference is present:   python oracle/make_ref.py
Only the files the path imports are copied (robot-perception-group/AirPose, MIT licence, copied with its LICENSE).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
PKG = "copenet/src/copenet"
FILES = [
    "LICENSE",
    PKG + "/__init__.py", PKG + "/config.py", PKG + "/constants.py", PKG + "/copenet_twoview.py", PKG + "/hmr.py",
    PKG + "/data/smpl_mean_params.npz",
    PKG + "/dsets/aerialpeople.py",
    PKG + "/models/model_copenet.py", PKG + "/models/model_hmr.py",
    PKG + "/smplx/smplx/__init__.py", PKG + "/smplx/smplx/body_models.py", PKG + "/smplx/smplx/joint_names.py",
    PKG + "/smplx/smplx/lbs.py", PKG + "/smplx/smplx/utils.py", PKG + "/smplx/smplx/vertex_ids.py",
    PKG + "/smplx/smplx/vertex_joint_selector.py",
    PKG + "/utils/geometry.py", PKG + "/utils/renderer.py", PKG + "/utils/utils.py",
]


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("make_ref: %s is absent (GPU box): keeping the shipped oracle/_ref" % SRC)
        return os.path.isdir(DST)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "robot-perception-group/AirPose (copied from /root/reference, unmodified)", "sha256": manifest}, f, indent=1)
    if verbose:
        print("make_ref: %d reference files -> %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
