"""ctypes binding of libairpose_b200.so (include/airpose_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
product path raises.  The numpy oracle under ``oracle/`` is test infrastructure and is
never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libairpose_b200.so")

c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)


class AirposeError(RuntimeError):
    pass


class SmplxModelHost(C.Structure):
    _fields_ = [("num_verts", C.c_int32), ("num_joints", C.c_int32), ("num_shape", C.c_int32),
                ("num_pose_basis", C.c_int32), ("num_faces", C.c_int32), ("num_landmarks", C.c_int32),
                ("num_extra", C.c_int32),
                ("v_template", C.c_void_p), ("shapedirs", C.c_void_p), ("posedirs", C.c_void_p),
                ("J_regressor", C.c_void_p), ("parents", C.c_void_p), ("lbs_weights", C.c_void_p),
                ("faces", C.c_void_p), ("lmk_faces_idx", C.c_void_p), ("lmk_bary_coords", C.c_void_p),
                ("extra_joint_idx", C.c_void_p)]


class SmplxFwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_betas", C.c_int32),
                ("betas", C.c_void_p), ("betas_stride", C.c_int32),
                ("global_orient", C.c_void_p), ("global_orient_stride", C.c_int32),
                ("body_pose", C.c_void_p), ("body_pose_stride", C.c_int32),
                ("tail_pose", C.c_void_p), ("tail_pose_stride", C.c_int32),
                ("transl", C.c_void_p),
                ("root_R", C.c_void_p), ("root_R_stride", C.c_int32),
                ("root_t", C.c_void_p), ("root_t_stride", C.c_int32),
                ("focal_x", C.c_float), ("focal_y", C.c_float),
                ("center", C.c_void_p), ("center_stride", C.c_int32),
                ("out_vertices", C.c_void_p), ("out_joints", C.c_void_p),
                ("out_vertices_cam", C.c_void_p), ("out_joints_cam", C.c_void_p),
                ("out_joints_2d", C.c_void_p), ("proj_t", C.c_void_p), ("proj_t_stride", C.c_int32)]


class SmplxBwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_betas", C.c_int32),
                ("betas", C.c_void_p), ("betas_stride", C.c_int32),
                ("global_orient", C.c_void_p), ("global_orient_stride", C.c_int32),
                ("body_pose", C.c_void_p), ("body_pose_stride", C.c_int32),
                ("joints", C.c_void_p),
                ("root_R", C.c_void_p), ("root_R_stride", C.c_int32),
                ("root_t", C.c_void_p), ("root_t_stride", C.c_int32),
                ("focal_x", C.c_float), ("focal_y", C.c_float),
                ("grad_vertices", C.c_void_p), ("grad_joints", C.c_void_p), ("grad_joints_cam", C.c_void_p),
                ("grad_joints_2d", C.c_void_p),
                ("grad_betas", C.c_void_p), ("grad_body_pose", C.c_void_p), ("grad_global_orient", C.c_void_p),
                ("grad_root_R", C.c_void_p), ("grad_root_t", C.c_void_p)]


class ConvParams(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bn_weight", C.c_void_p), ("bn_bias", C.c_void_p),
                ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p)]


class NetParams(C.Structure):
    _fields_ = [("conv", ConvParams * 53),
                ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p),
                ("decpose_w", C.c_void_p), ("decpose_b", C.c_void_p),
                ("decshape_w", C.c_void_p), ("decshape_b", C.c_void_p),
                ("init_pose", C.c_void_p), ("init_shape", C.c_void_p), ("bn_eps", C.c_float)]


class BnTrainParams(C.Structure):
    _fields_ = [("bn_weight", C.c_void_p * 53), ("bn_bias", C.c_void_p * 53), ("running_mean", C.c_void_p * 53),
                ("running_var", C.c_void_p * 53), ("momentum", C.c_float), ("eps", C.c_float), ("saved_stats", C.c_void_p),
                ("tape", C.c_int32)]


UPPER_DONE_FN = C.CFUNCTYPE(None, C.c_void_p)


class TrunkGrads(C.Structure):
    _fields_ = [("g_weight", C.c_void_p * 53), ("g_bn_weight", C.c_void_p * 53), ("g_bn_bias", C.c_void_p * 53),
                ("accumulate", C.c_int32), ("upper_done", UPPER_DONE_FN), ("user", C.c_void_p)]


class HmrParams(C.Structure):
    _fields_ = [("conv", ConvParams * 53),
                ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p),
                ("decpose_w", C.c_void_p), ("decpose_b", C.c_void_p),
                ("decshape_w", C.c_void_p), ("decshape_b", C.c_void_p),
                ("deccam_w", C.c_void_p), ("deccam_b", C.c_void_p),
                ("init_pose", C.c_void_p), ("init_shape", C.c_void_p), ("init_cam", C.c_void_p), ("bn_eps", C.c_float)]


class HmrIefArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("iters", C.c_int32), ("xf", C.c_void_p),
                ("init_theta", C.c_void_p), ("init_theta_stride", C.c_int32),
                ("init_shape", C.c_void_p), ("init_shape_stride", C.c_int32),
                ("init_cam", C.c_void_p), ("init_cam_stride", C.c_int32),
                ("out_pose", C.c_void_p), ("out_betas", C.c_void_p), ("out_cam", C.c_void_p)]


class IefTrainArgs(C.Structure):
    _fields_ = ([("batch", C.c_int32), ("iters", C.c_int32)] +
                [(n, C.c_void_p) for n in ("xf0", "xf1", "bb0", "bb1", "pos0", "pos1")] +
                [("init_theta0", C.c_void_p), ("init_theta1", C.c_void_p), ("init_theta_stride", C.c_int32),
                 ("init_shape0", C.c_void_p), ("init_shape1", C.c_void_p), ("init_shape_stride", C.c_int32)] +
                [(n, C.c_void_p) for n in ("fc1_w", "fc1_b", "fc2_w", "fc2_b", "decpose_w", "decpose_b", "decshape_w", "decshape_b",
                                           "init_pose", "init_shape", "mask1", "mask2", "saved", "workspace",
                                           "out_pose0", "out_betas0", "out_pose1", "out_betas1",
                                           "g_pose0", "g_betas0", "g_pose1", "g_betas1",
                                           "g_fc1_w", "g_fc1_b", "g_fc2_w", "g_fc2_b", "g_decpose_w", "g_decpose_b",
                                           "g_decshape_w", "g_decshape_b", "g_xf0", "g_xf1")])


class IefArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("iters", C.c_int32),
                ("xf0", C.c_void_p), ("xf1", C.c_void_p), ("bb0", C.c_void_p), ("bb1", C.c_void_p),
                ("pos0", C.c_void_p), ("pos1", C.c_void_p),
                ("init_theta0", C.c_void_p), ("init_theta1", C.c_void_p), ("init_theta_stride", C.c_int32),
                ("init_shape0", C.c_void_p), ("init_shape1", C.c_void_p), ("init_shape_stride", C.c_int32),
                ("out_pose0", C.c_void_p), ("out_betas0", C.c_void_p),
                ("out_pose1", C.c_void_p), ("out_betas1", C.c_void_p)]


class GemmArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("residual", C.c_void_p), ("ldr", C.c_int64), ("relu", C.c_int32),
                ("out_bf16", C.c_void_p), ("ldd", C.c_int64), ("out_f32", C.c_void_p), ("ldf", C.c_int64),
                ("a_t", C.c_int32), ("b_t", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("n", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
                ("w", C.c_void_p), ("Cout", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
                ("pad", C.c_int32), ("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p),
                ("relu", C.c_int32), ("out", C.c_void_p)]


class BneckTailArgs(C.Structure):
    _fields_ = [("t1", C.c_void_p), ("n", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cm", C.c_int32),
                ("w2", C.c_void_p), ("scale2", C.c_void_p), ("shift2", C.c_void_p),
                ("w3", C.c_void_p), ("scale3", C.c_void_p), ("shift3", C.c_void_p),
                ("residual", C.c_void_p), ("out", C.c_void_p)]


class LossArgs(C.Structure):
    _fields_ = ([("batch", C.c_int32), ("num_verts", C.c_int32), ("num_joints", C.c_int32),
                 ("trans0", C.c_void_p), ("trans1", C.c_void_p), ("trans_stride", C.c_int32)] +
                [(n, C.c_void_p) for n in ("rotmat0", "rotmat1", "betas0", "betas1", "verts0", "verts1", "joints0", "joints1",
                                           "j2d0", "j2d1", "gt_pose_rotmat", "gt_trans0", "gt_trans1", "gt_orient0",
                                           "gt_orient1", "gt_verts", "gt_joints", "gt_j2d0", "gt_j2d1")] +
                [(n, C.c_float) for n in ("w_shape", "w_kp2d", "w_kp3d", "w_limbs3d", "w_limbstheta", "w_trans",
                                          "w_rootrot", "w_pose", "w_beta")] +
                [("out", C.c_void_p)] +
                [(n, C.c_void_p) for n in ("g_verts0", "g_verts1", "g_joints0", "g_joints1", "g_j2d0", "g_j2d1",
                                           "g_rotmat0", "g_rotmat1", "g_betas0", "g_betas1", "g_trans0", "g_trans1")])


class RealLossArgs(C.Structure):
    _fields_ = ([("batch", C.c_int32), ("num_joints", C.c_int32), ("gt_joints", C.c_int32),
                 ("trans0", C.c_void_p), ("trans1", C.c_void_p), ("trans_stride", C.c_int32)] +
                [(n, C.c_void_p) for n in ("rotmat0", "rotmat1", "betas0", "betas1", "j2d0", "j2d1", "gt_j2d0", "gt_j2d1")] +
                [(n, C.c_float) for n in ("w_kp2d", "w_limbs2d", "w_beta", "w_pose", "w_vposer", "vposer_term")] +
                [("out", C.c_void_p)] +
                [(n, C.c_void_p) for n in ("g_j2d0", "g_j2d1", "g_rotmat0", "g_rotmat1", "g_betas0", "g_betas1", "g_trans0", "g_trans1")])


class AdamArgs(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("max_exp_avg_sq", C.c_void_p), ("n", C.c_int64), ("lr", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float), ("step", C.c_int32), ("grad_scale", C.c_float)]


# name -> (restype, argtypes); every symbol include/airpose_b200.h declares
SYMBOLS = {
    "airpose_last_error": (C.c_char_p, []),
    "airpose_abi_version": (C.c_int, []),
    "airpose_launch_count": (C.c_int64, []),
    "airpose_smplx_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(SmplxModelHost), C.c_int]),
    "airpose_smplx_destroy": (C.c_int, [C.c_void_p]),
    "airpose_smplx_skin_nnz": (C.c_int, [C.c_void_p]),
    "airpose_smplx_fwd": (C.c_int, [C.c_void_p, C.POINTER(SmplxFwdArgs), C.c_void_p]),
    "airpose_smplx_bwd": (C.c_int, [C.c_void_p, C.POINTER(SmplxBwdArgs), C.c_void_p]),
    "airpose_rot6d_to_rotmat": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "airpose_rot6d_to_rotmat_strided": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "airpose_rot6d_to_rotmat_bwd_strided": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                                      C.c_int64, C.c_void_p]),
    "airpose_j14_gather": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i32_p, C.c_void_p, C.c_void_p]),
    "airpose_rotmat_to_angle_axis": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "airpose_angle_axis_to_rotmat": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "airpose_mean_distance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "airpose_preprocess_bgr8": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_float_p, c_float_p, C.c_void_p, C.c_void_p]),
    "airpose_preprocess_crop_resize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                                 c_float_p, c_float_p, C.c_void_p, C.c_void_p]),
    "airpose_net_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "airpose_net_destroy": (C.c_int, [C.c_void_p]),
    "airpose_net_load": (C.c_int, [C.c_void_p, C.POINTER(NetParams), C.c_void_p]),
    "airpose_backbone_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "airpose_bn_saved_stats_floats": (C.c_int64, []),
    "airpose_backbone_fwd_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(BnTrainParams), C.c_void_p, C.c_void_p]),
    "airpose_backbone_bwd_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(BnTrainParams), C.c_void_p,
                                             C.POINTER(TrunkGrads), C.POINTER(C.c_void_p * 53), C.c_void_p]),
    "airpose_backbone_fwd_train_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(BnTrainParams), C.c_void_p, C.c_void_p]),
    "airpose_backbone_bwd_train_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(BnTrainParams), C.c_void_p,
                                                  C.POINTER(TrunkGrads), C.POINTER(C.c_void_p * 53), C.c_void_p]),
    "airpose_debug_conv_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_void_p]),
    "airpose_debug_bn_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "airpose_debug_tape_get": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "airpose_net_load_trunk": (C.c_int, [C.c_void_p, C.POINTER(NetParams), C.c_void_p]),
    "airpose_net_load_regressor": (C.c_int, [C.c_void_p, C.POINTER(NetParams), C.c_void_p]),
    "airpose_backbone_fwd_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "airpose_ief_fwd": (C.c_int, [C.c_void_p, C.POINTER(IefArgs), C.c_void_p]),
    "airpose_ief_train_saved_floats": (C.c_int64, [C.c_int32, C.c_int32]),
    "airpose_ief_train_workspace_floats": (C.c_int64, [C.c_int32]),
    "airpose_ief_train_fwd": (C.c_int, [C.POINTER(IefTrainArgs), C.c_void_p]),
    "airpose_ief_train_bwd": (C.c_int, [C.POINTER(IefTrainArgs), C.c_void_p]),
    "airpose_hmr_load": (C.c_int, [C.c_void_p, C.POINTER(HmrParams), C.c_void_p]),
    "airpose_hmr_ief_fwd": (C.c_int, [C.c_void_p, C.POINTER(HmrIefArgs), C.c_void_p]),
    "airpose_adam_step": (C.c_int, [C.POINTER(AdamArgs), C.c_void_p]),
    "airpose_twoview_loss": (C.c_int, [C.POINTER(LossArgs), C.c_void_p]),
    "airpose_real_loss": (C.c_int, [C.POINTER(RealLossArgs), C.c_void_p]),
    "airpose_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "airpose_conv_bf16": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "airpose_bneck_tail_bf16": (C.c_int, [C.POINTER(BneckTailArgs), C.c_void_p]),
    "airpose_backbone_stem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (building nothing here): raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AirposeError(
            "{} is missing: build it with `python -m airpose_b200.build` (nvcc, sm_100a). "
            "airpose_b200 has no CPU or PyTorch fallback.".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.airpose_abi_version() != 1:
        raise AirposeError("libairpose_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().airpose_last_error().decode("utf-8", "replace")
        raise AirposeError("{} failed (rc={}): {}".format(what or "airpose call", rc, msg))


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().airpose_launch_count())
