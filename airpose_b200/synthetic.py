"""Deterministic synthetic stand-ins for the assets AirPose downloads.

The reference needs three licensed / downloaded assets that exist nowhere on disk
(SURVEY.md section 7 "No model assets"): the SMPL-X body model ``SMPLX_NEUTRAL.npz``
(read at copenet/src/copenet/smplx/smplx/body_models.py:205-296,493-514,727-730),
ImageNet ResNet-50 weights (copenet/src/copenet/models/model_copenet.py:236-238) and
``smpl_mean_params.npz`` (model_copenet.py:86-92).  Everything here is generated
with ``numpy.random.default_rng`` (bit-stable across machines), so the golden
fixtures made in the build container can be re-derived on the GPU box.

Only numpy is imported: the oracle, the golden generator, the tests and bench.py
all share these generators.
"""
from __future__ import annotations

import os
import numpy as np

NUM_VERTS = 10475
NUM_JOINTS = 55
NUM_FACES = 20908
NUM_SHAPE_COEFFS = 20          # 10 betas + 10 expression
NUM_POSE_BASIS = (NUM_JOINTS - 1) * 9
NUM_LANDMARKS = 51

# Standard SMPL-X kinematic tree (``parents`` row).
SMPLX_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
     15, 15, 15,
     20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
     21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53], dtype=np.int64)

# Camera constants (copenet/src/copenet/constants.py:7-11).
FOCAL_LENGTH = (1475.0, 1475.0)
IMG_SIZE = (1920, 1080)


def _rest_skeleton() -> np.ndarray:
    """Rough T-pose joint positions (metres) for the 55 SMPL-X joints."""
    J = np.zeros((NUM_JOINTS, 3), dtype=np.float64)
    J[0] = (0.0, -0.35, 0.0)                                  # pelvis
    J[1] = (0.06, -0.44, 0.0); J[2] = (-0.06, -0.44, 0.0)      # hips
    J[3] = (0.0, -0.23, -0.02)                                 # spine1
    J[4] = (0.10, -0.82, 0.0); J[5] = (-0.10, -0.82, 0.0)      # knees
    J[6] = (0.0, -0.09, 0.0)                                   # spine2
    J[7] = (0.09, -1.22, -0.03); J[8] = (-0.09, -1.22, -0.03)  # ankles
    J[9] = (0.0, -0.03, 0.01)                                  # spine3
    J[10] = (0.11, -1.28, 0.09); J[11] = (-0.11, -1.28, 0.09)  # feet
    J[12] = (0.0, 0.17, -0.03)                                 # neck
    J[13] = (0.05, 0.09, -0.02); J[14] = (-0.05, 0.09, -0.02)  # collars
    J[15] = (0.0, 0.27, 0.01)                                  # head
    J[16] = (0.17, 0.12, -0.03); J[17] = (-0.17, 0.12, -0.03)  # shoulders
    J[18] = (0.43, 0.11, -0.05); J[19] = (-0.43, 0.11, -0.05)  # elbows
    J[20] = (0.68, 0.11, -0.05); J[21] = (-0.68, 0.11, -0.05)  # wrists
    J[22] = (0.0, 0.26, 0.04)                                  # jaw
    J[23] = (0.03, 0.31, 0.07); J[24] = (-0.03, 0.31, 0.07)    # eyes
    for side, wrist, base in ((1.0, 20, 25), (-1.0, 21, 40)):
        for f in range(5):                                     # five fingers x 3 phalanges
            z = (f - 2) * 0.02
            for k in range(3):
                J[base + 3 * f + k] = (J[wrist][0] + side * (0.08 + 0.03 * k),
                                       J[wrist][1] - 0.005 * f, z)
    return J


def make_smplx_model(seed: int = 0) -> dict:
    """Synthetic SMPL-X neutral model with the keys body_models.py reads.

    Shapes follow SURVEY.md section 8(c): ``shapedirs`` has 20 columns because the
    fork never slices it (body_models.py:273-277, lbs.py:265).  ``weights`` has at
    most four non-zeros per vertex and ``J_regressor`` is row-stochastic and sparse,
    like the real model; both are stored dense as the reference densifies them
    (smplx/utils.py:36-39).
    """
    rng = np.random.default_rng(seed)
    V, J = NUM_VERTS, NUM_JOINTS
    parents = SMPLX_PARENTS
    Jrest = _rest_skeleton()

    # Vertices: region-coherent blocks per joint (the real mesh numbering is
    # region-coherent too), block order shuffled.
    share = np.ones(J)
    share[:22] = 6.0
    share[15] = 14.0
    counts = np.floor(share / share.sum() * V).astype(np.int64)
    counts[0] += V - counts.sum()
    order = rng.permutation(J)
    owner = np.concatenate([np.full(counts[j], j) for j in order])
    radius = np.where(np.arange(J) < 22, 0.07, 0.012)
    v_template = Jrest[owner] + rng.normal(size=(V, 3)) * radius[owner][:, None]

    # Skinning weights: owner joint + parent + up to two children, <= 4 non-zeros.
    children = [np.nonzero(parents == j)[0] for j in range(J)]
    weights = np.zeros((V, J), dtype=np.float64)
    for v in range(V):
        j = owner[v]
        cand = [j]
        if parents[j] >= 0:
            cand.append(int(parents[j]))
        ch = list(children[j])
        rng.shuffle(ch)
        cand += [int(c) for c in ch[:2]]
        w = rng.random(len(cand)) ** 2
        w[0] += 1.0
        drop = rng.random(len(cand)) < 0.3
        drop[0] = False
        w[drop] = 0.0
        weights[v, cand] = w / w.sum()

    # Joint regressor: each joint is a convex combination of ~40 nearby vertices.
    J_regressor = np.zeros((J, V), dtype=np.float64)
    for j in range(J):
        d = np.linalg.norm(v_template - Jrest[j], axis=1)
        idx = np.argsort(d)[:40]
        w = rng.random(40) + 0.05
        J_regressor[j, idx] = w / w.sum()

    shapedirs = rng.normal(size=(V, 3, NUM_SHAPE_COEFFS)) * 0.012
    shapedirs *= (0.85 ** np.arange(NUM_SHAPE_COEFFS))[None, None, :]
    posedirs = rng.normal(size=(V, 3, NUM_POSE_BASIS)) * 0.004

    # Faces: triples of nearby vertex ids; landmark faces and barycentrics.
    base = rng.integers(0, V - 8, size=NUM_FACES)
    f = np.stack([base, base + rng.integers(1, 4, size=NUM_FACES),
                  base + rng.integers(4, 8, size=NUM_FACES)], axis=1).astype(np.uint32)
    lmk_faces_idx = rng.integers(0, NUM_FACES, size=NUM_LANDMARKS).astype(np.int64)
    bary = rng.random((NUM_LANDMARKS, 3)) + 0.05
    lmk_bary_coords = bary / bary.sum(axis=1, keepdims=True)

    kintree = np.stack([parents.copy(), np.arange(J)], axis=0).astype(np.int64)
    kintree[0, 0] = 2 ** 32 - 1        # the real files store uint32(-1) here (body_models.py:291-292)

    return {
        "v_template": v_template.astype(np.float32),
        "shapedirs": shapedirs.astype(np.float32),
        "posedirs": posedirs.astype(np.float32),
        "J_regressor": J_regressor.astype(np.float32),
        "kintree_table": kintree,
        "weights": weights.astype(np.float32),
        "f": f,
        "hands_componentsl": rng.normal(size=(45, 45)).astype(np.float32),
        "hands_componentsr": rng.normal(size=(45, 45)).astype(np.float32),
        "hands_meanl": (rng.normal(size=45) * 0.1).astype(np.float32),
        "hands_meanr": (rng.normal(size=45) * 0.1).astype(np.float32),
        "lmk_faces_idx": lmk_faces_idx,
        "lmk_bary_coords": lmk_bary_coords.astype(np.float32),
    }


def write_smplx_model(directory: str, seed: int = 0, gender: str = "neutral") -> str:
    """Write ``SMPLX_<GENDER>.npz`` into ``directory`` (created) and return its path."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "SMPLX_{}.npz".format(gender.upper()))
    if not os.path.exists(path):
        tmp = path + ".tmp.{}.npz".format(os.getpid())
        np.savez(tmp, **make_smplx_model(seed))
        os.replace(tmp, path)
    return path


def make_mean_params(seed: int = 1) -> dict:
    """Stand-in for ``smpl_mean_params.npz`` (model_copenet.py:86-92): pose (144,), shape (10,), cam (3,)."""
    rng = np.random.default_rng(seed)
    pose = np.tile(np.array([1, 0, 0, 1, 0, 0], dtype=np.float32), 24)
    pose = pose + rng.normal(size=144).astype(np.float32) * 0.15
    shape = (rng.normal(size=10) * 0.3).astype(np.float32)
    cam = np.array([0.9, 0.0, 0.0], dtype=np.float32)
    return {"pose": pose.astype(np.float32), "shape": shape, "cam": cam}


def write_mean_params(path: str, seed: int = 1) -> str:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    if not os.path.exists(path):
        tmp = path + ".tmp.{}.npz".format(os.getpid())
        np.savez(tmp, **make_mean_params(seed))
        os.replace(tmp, path)
    return path


# --------------------------------------------------------------------------------------
# Network weights.  Layout and names follow the reference state_dict
# (model_copenet.py:53-92: conv1, bn1, layer{1-4}.{i}.{conv,bn}{1-3}, downsample.{0,1},
# fc1, fc2, decpose, decshape, deccam, init_pose, init_shape, init_cam).
# --------------------------------------------------------------------------------------
RESNET_LAYERS = (3, 4, 6, 3)
RESNET_PLANES = (64, 128, 256, 512)
NPOSE = 21 * 6
FC1_IN = 2048 + 3 + 3 + 6 + NPOSE + 10 + NPOSE + 10        # 2332, model_copenet.py:67


def conv_specs():
    """Yield (name, cout, cin, k, stride, pad, bn_name) for the 53 convs in forward order."""
    yield ("conv1", 64, 3, 7, 2, 3, "bn1")
    inplanes = 64
    for li, (blocks, planes) in enumerate(zip(RESNET_LAYERS, RESNET_PLANES), start=1):
        for b in range(blocks):
            stride = 2 if (li > 1 and b == 0) else 1
            p = "layer{}.{}".format(li, b)
            yield (p + ".conv1", planes, inplanes, 1, 1, 0, p + ".bn1")
            yield (p + ".conv2", planes, planes, 3, stride, 1, p + ".bn2")
            yield (p + ".conv3", planes * 4, planes, 1, 1, 0, p + ".bn3")
            if b == 0:
                yield (p + ".downsample.0", planes * 4, inplanes, 1, stride, 0, p + ".downsample.1")
            inplanes = planes * 4


FC1_IN_HMR = 2048 + 22 * 6 + 10 + 3                          # 2193, model_hmr.py:66


def make_network_state(seed: int = 123, dec_gain: float = 0.25, mean_seed: int = 1, variant: str = "twoview") -> dict:
    """Random copenet state_dict (numpy arrays, reference key names).  ``variant="hmr"`` gives the
    single-view model of model_hmr.py (fc1 in 2193, decpose 132 rows; same trunk draw as "twoview").

    Conv init is the reference's He fan-out normal (model_copenet.py:78-81).  BN running
    statistics and affine parameters are randomised and the decoder gains raised as
    SURVEY.md section 8(d) prescribes (the reference's gain 0.01 makes the three IEF
    iterations a near no-op and would hide bugs); bn3 gammas are kept small so the
    residual stream stays O(1) through 16 blocks.
    """
    rng = np.random.default_rng(seed)
    sd = {}

    def bn(name, c, gamma_lo, gamma_hi):
        sd[name + ".weight"] = rng.uniform(gamma_lo, gamma_hi, size=c).astype(np.float32)
        sd[name + ".bias"] = (rng.normal(size=c) * 0.1).astype(np.float32)
        sd[name + ".running_mean"] = (rng.normal(size=c) * 0.1).astype(np.float32)
        sd[name + ".running_var"] = rng.uniform(0.5, 1.5, size=c).astype(np.float32)
        sd[name + ".num_batches_tracked"] = np.array(0, dtype=np.int64)

    for name, cout, cin, k, stride, pad, bn_name in conv_specs():
        std = np.sqrt(2.0 / (k * k * cout))
        sd[name + ".weight"] = (rng.standard_normal(size=(cout, cin, k, k)) * std).astype(np.float32)
        if bn_name.endswith("bn3"):
            bn(bn_name, cout, 0.15, 0.45)
        else:
            bn(bn_name, cout, 0.5, 1.5)

    def linear(name, cout, cin, bound_w, bound_b):
        sd[name + ".weight"] = rng.uniform(-bound_w, bound_w, size=(cout, cin)).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-bound_b, bound_b, size=cout).astype(np.float32)

    fc1_in = FC1_IN if variant == "twoview" else FC1_IN_HMR
    linear("fc1", 1024, fc1_in, 1.0 / np.sqrt(fc1_in), 1.0 / np.sqrt(fc1_in))
    linear("fc2", 1024, 1024, 1.0 / 32.0, 1.0 / 32.0)
    for name, cout in (("decpose", 3 + 6 + NPOSE if variant == "twoview" else 22 * 6), ("decshape", 10), ("deccam", 3)):
        bound = dec_gain * np.sqrt(6.0 / (1024 + cout))
        linear(name, cout, 1024, bound, 1.0 / 32.0)

    mp = make_mean_params(mean_seed)
    sd["init_pose"] = mp["pose"][None, :].astype(np.float32)
    sd["init_shape"] = mp["shape"][None, :].astype(np.float32)
    sd["init_cam"] = mp["cam"][None, :].astype(np.float32)
    return sd


def make_inputs(batch: int, seed: int = 123) -> dict:
    """Seeded synthetic batch for the two-view path."""
    rng = np.random.default_rng(seed)
    out = {}
    for v in (0, 1):
        out["im%d" % v] = rng.standard_normal(size=(batch, 3, 224, 224)).astype(np.float32)
        bb = np.empty((batch, 3), dtype=np.float32)
        bb[:, :2] = rng.uniform(-1, 1, size=(batch, 2))
        bb[:, 2] = rng.uniform(0.1, 2.0, size=batch)
        out["bb%d" % v] = bb
        intr = np.array([[FOCAL_LENGTH[0], 0, IMG_SIZE[0] / 2.0],
                         [0, FOCAL_LENGTH[1], IMG_SIZE[1] / 2.0],
                         [0, 0, 1]], dtype=np.float32)
        out["intr%d" % v] = np.broadcast_to(intr, (batch, 3, 3)).copy()
        out["smpltrans_rel%d" % v] = (np.array([0, 0, 10], dtype=np.float32)
                                      + rng.normal(size=(batch, 3)).astype(np.float32))
    return out


def rot6d_to_rotmat_np(x: np.ndarray) -> np.ndarray:
    """numpy twin of geometry.rot6d_to_rotmat (copenet/src/copenet/utils/geometry.py:47-61)."""
    x = x.reshape(-1, 3, 2).astype(np.float32)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = a1 / np.maximum(np.linalg.norm(a1, axis=1, keepdims=True), 1e-12)
    u = a2 - np.sum(b1 * a2, axis=1, keepdims=True) * b1
    b2 = u / np.maximum(np.linalg.norm(u, axis=1, keepdims=True), 1e-12)
    b3 = np.cross(b1, b2)
    return np.stack([b1, b2, b3], axis=-1).astype(np.float32)


def make_lbs_inputs(batch: int, seed: int = 7, pose_scale: float = 0.35) -> dict:
    """betas ~N(0,1) and random body rotations for the standalone lbs() config (section 8(d))."""
    rng = np.random.default_rng(seed)
    betas = rng.standard_normal(size=(batch, 10)).astype(np.float32)
    six = np.tile(np.array([1, 0, 0, 1, 0, 0], dtype=np.float32), (batch, 21, 1))
    six = six + rng.standard_normal(size=(batch, 21, 6)).astype(np.float32) * pose_scale
    body = rot6d_to_rotmat_np(six.reshape(-1, 6)).reshape(batch, 21, 3, 3)
    return {"betas": betas, "body_pose": body}


# --------------------------------------------------------------------------------------
# Camera frames and drone-server messages
# --------------------------------------------------------------------------------------
def camera_frame(height: int, width: int, seed: int) -> np.ndarray:
    """A seeded u8 BGR frame [H,W,3] as cv2.imread would return it: low-frequency structure plus pixel noise, so that a
    bilinear resize is sensitive to both the sample positions and the weights."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    img = np.empty((height, width, 3), np.float64)
    for c in range(3):
        fy, fx, ph = rng.uniform(0.01, 0.08), rng.uniform(0.01, 0.08), rng.uniform(0, 6.28)
        img[:, :, c] = 127.5 + 80.0 * np.sin(fy * y + fx * x + ph) + rng.normal(0, 25.0, size=(height, width))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def server_messages(seed: int, frames: int, init_pose: np.ndarray, init_shape: np.ndarray):
    """Seeded messages in the drone server's wire format (airpose_server/server.py:38-39,91-98,110,125;
    airpose_client/AirPoseClient.h:20-31), ``frames`` x (stage 0, stage 1, stage 2):
      stage 0:   u8 stage | 3 x f32 bb (cx/c_x-1, cy/c_y-1, scale) | 224*224*3 u8 BGR
      stage 1/2: u8 stage | 10 x f32 betas | 126 x f32 articulated 6D pose   (the OTHER drone's previous reply)
    Returns [(stage, bytes), ...]."""
    rng = np.random.default_rng(seed)
    init_pose = np.asarray(init_pose, np.float32).reshape(-1)
    init_shape = np.asarray(init_shape, np.float32).reshape(-1)
    msgs = []
    for f in range(frames):
        bb = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.1, 2.0)], np.float32)
        img = camera_frame(224, 224, seed * 1000 + f)
        msgs.append((0, bytes([0]) + bb.tobytes() + img.tobytes()))
        for stage in (1, 2):
            betas = (init_shape + rng.normal(0, 0.3, size=10)).astype(np.float32)
            art = (init_pose[6:22 * 6] + rng.normal(0, 0.1, size=126)).astype(np.float32)
            msgs.append((stage, bytes([stage]) + betas.tobytes() + art.tobytes()))
    return msgs
