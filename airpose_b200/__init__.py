"""airpose_b200: the AirPose copenet_twoview forward on hand-written sm_100a kernels.

Public surface (mirrors the reference's own):
    airpose_b200.model_copenet.getcopenet / copenet      <- copenet/models/model_copenet.py
    airpose_b200.smplx.SMPLX / ModelOutput               <- copenet/smplx/smplx/body_models.py
    airpose_b200.smplx.rot6d_to_rotmat                   <- copenet/utils/geometry.py
    airpose_b200.copenet_twoview.copenet_twoview         <- copenet/copenet_twoview.py (forward path)
Everything computes through libairpose_b200.so (C ABI in include/airpose_b200.h); there is
no CPU or PyTorch fallback.
"""
__all__ = ["model_copenet", "smplx", "copenet_twoview", "synthetic"]
