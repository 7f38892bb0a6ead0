"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/airpose_b200.h declares (no compute without a GPU), the ctypes table matches the
header, the product path refuses to run without CUDA, and the module keeps the reference's
state_dict keys."""
import os
import re

import numpy as np
import pytest
import torch

from airpose_b200 import _lib, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "airpose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(airpose_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 17
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), "libairpose_b200.so does not export " + n
    assert sorted(_lib.SYMBOLS) == names
    assert lib.airpose_abi_version() == 1
    assert lib.airpose_launch_count() == 0
    assert lib.airpose_last_error() == b""


def test_error_channel_reports_bad_arguments():
    lib = _lib.load()
    rc = lib.airpose_rot6d_to_rotmat(None, 4, None, None)     # rejected before any CUDA call
    assert rc != 0
    assert b"airpose_rot6d_to_rotmat" in lib.airpose_last_error()


def test_no_cpu_fallback(tmp_path):
    from airpose_b200.model_copenet import getcopenet
    from airpose_b200.smplx import SMPLX
    mp = synthetic.write_mean_params(str(tmp_path / "smpl_mean_params.npz"))
    net = getcopenet(mp, pretrained=False).eval()
    with pytest.raises(_lib.AirposeError):
        net.forward_feat_ext(torch.zeros(1, 3, 224, 224))
    synthetic.write_smplx_model(str(tmp_path), 0)
    sm = SMPLX(str(tmp_path), batch_size=1, create_transl=False)
    with pytest.raises(_lib.AirposeError):
        sm.forward(betas=torch.zeros(1, 10), body_pose=torch.eye(3).expand(1, 21, 3, 3), pose2rot=False)


def test_state_dict_keys_match_reference_layout(tmp_path, net_state):
    """331 entries, same names as the reference module."""
    from airpose_b200.model_copenet import getcopenet
    mp = synthetic.write_mean_params(str(tmp_path / "smpl_mean_params.npz"))
    net = getcopenet(mp, pretrained=False)
    sd = net.state_dict()
    assert len(sd) == 331
    assert set(sd) == set(net_state)
    assert sum(p.numel() for p in net.parameters()) == 27098324
    assert tuple(sd["fc1.weight"].shape) == (1024, 2332)
    for k, v in net_state.items():
        assert tuple(sd[k].shape) == tuple(np.asarray(v).shape), k
    specs = list(synthetic.conv_specs())
    assert len(specs) == 53
    assert [s[0] for s in specs] == [n[:-7] for n, m in ((n + ".weight", m) for n, m in net.named_modules()
                                                       if isinstance(m, torch.nn.Conv2d))]
    # Lightning checkpoints prefix the network with "model." (airpose_server/server.py:16-22)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in net_state.items()}, strict=True)


def test_smplx_module_surface(tmp_path):
    from airpose_b200.smplx import SMPLX, ModelOutput
    synthetic.write_smplx_model(str(tmp_path), 0)
    sm = SMPLX(str(tmp_path), batch_size=3, create_transl=False)
    assert sm.batch_size == 3 and not hasattr(sm, "transl")
    assert tuple(sm.v_template.shape) == (10475, 3) and sm.faces.shape == (20908, 3)
    assert tuple(sm.posedirs.shape) == (486, 31425) and int(sm.parents[0]) == -1
    assert sm.faces_tensor.dtype == torch.long
    assert tuple(sm.betas.shape) == (3, 10) and tuple(sm.expression.shape) == (3, 10)
    assert sm.vertex_joint_selector.extra_joints_idxs.tolist()[:5] == [9120, 9929, 9448, 616, 6]
    assert ModelOutput._fields[:2] == ("vertices", "joints")
    with pytest.raises(NotImplementedError):
        sm.forward(betas=torch.zeros(3, 10), pose2rot=True)


def test_hmr_state_dict_keys_match_reference_layout(tmp_path):
    """model_hmr.getcopenet: fc1 takes 2193 inputs, decpose decodes 132 numbers, same 331 keys."""
    from airpose_b200.model_hmr import getcopenet
    mp = synthetic.write_mean_params(str(tmp_path / "smpl_mean_params.npz"))
    net = getcopenet(mp, pretrained=False)
    state = synthetic.make_network_state(123, variant="hmr")
    sd = net.state_dict()
    assert len(sd) == 331 and set(sd) == set(state)
    assert tuple(sd["fc1.weight"].shape) == (1024, 2193) and tuple(sd["decpose.weight"].shape) == (132, 1024)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
    with pytest.raises(_lib.AirposeError):
        net.eval()(torch.zeros(1, 3, 224, 224))


def test_server_wire_constants_and_letterbox_geometry():
    """Host logic of the two 'next' rows that needs no GPU: the drone server's message sizes (server.py:37-39) and reply
    lengths, the checkpoint-key fix-up (server.py:16-22), and the letterbox geometry of utils.resize_with_pad
    (utils.py:218-229) against the values the reference itself returned for the golden crops."""
    import airpose_oracle as orc
    from airpose_b200 import server
    from airpose_b200.preprocess import letterbox_geometry
    assert (server.SIZE, server.BUFFERSIZE, server.BUFFERSIZE_STAGES) == (224, 150541, 545)
    assert (server.BUFFERSIZE, server.BUFFERSIZE_STAGES) == (orc.SERVER_BUFFERSIZE, orc.SERVER_BUFFERSIZE_STAGES)
    assert server.REPLY_FLOATS == (136, 136, 145)
    sd = server.fix_state_dict({"model.fc1.weight": 1, "model.model.x": 2, "smplx.betas": 3})
    assert list(sd) == ["fc1.weight", "model.x", "smplx.betas"]
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz")))
    for i, case in enumerate(g["cases"]):
        _, _, _, y0, y1, x0, x1 = (int(v) for v in case)
        scale, (dw, dh), pad = letterbox_geometry(y1 - y0, x1 - x0)
        assert scale == float(g["scale_%d" % i]) and pad == g["pad_%d" % i].tolist()
        assert max(dw, dh) in (223, 224) and min(pad) >= 0
    with pytest.raises(_lib.AirposeError):          # CUDA only, no CPU path
        import tempfile
        server.StagedServer(server.getmodel(synthetic.write_mean_params(os.path.join(tempfile.mkdtemp(), "smpl_mean_params.npz"))), device="cpu")


_STRUCTS = {   # header typedef -> ctypes mirror (airpose_b200/_lib.py)
    "airpose_smplx_model_host": "SmplxModelHost", "airpose_smplx_fwd_args": "SmplxFwdArgs", "airpose_smplx_bwd_args": "SmplxBwdArgs",
    "airpose_conv_params": "ConvParams", "airpose_net_params": "NetParams", "airpose_bn_train_params": "BnTrainParams",
    "airpose_trunk_grads": "TrunkGrads", "airpose_ief_args": "IefArgs", "airpose_ief_train_args": "IefTrainArgs",
    "airpose_hmr_params": "HmrParams", "airpose_hmr_ief_args": "HmrIefArgs", "airpose_twoview_loss_args": "LossArgs",
    "airpose_real_loss_args": "RealLossArgs", "airpose_adam_args": "AdamArgs", "airpose_gemm_args": "GemmArgs",
    "airpose_conv_args": "ConvArgs", "airpose_bneck_tail_args": "BneckTailArgs",
}


def test_struct_sizes_match_the_header(tmp_path):
    """Every argument struct of include/airpose_b200.h has the size its ctypes mirror has (gcc compiles the header as plain C):
    a field added on one side only would shift everything behind it silently."""
    import ctypes
    import shutil
    import subprocess
    from airpose_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sizes.c"
    lines = ['#include <stdio.h>', '#include "airpose_b200.h"', "int main(void) {"]
    lines += ['  printf("%s %%zu\\n", sizeof(%s));' % (name, name) for name in _STRUCTS]
    fields = [("airpose_gemm_args", "GemmArgs", f) for f in ("K", "relu", "out_f32", "a_t", "b_t")] + \
             [("airpose_trunk_grads", "TrunkGrads", f) for f in ("g_bn_bias", "accumulate", "upper_done", "user")] + \
             [("airpose_conv_args", "ConvArgs", f) for f in ("w", "pad", "relu", "out")]
    lines.insert(1, "#include <stddef.h>")
    lines += ['  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (t, f, t, f) for t, _, f in fields]
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "sizes"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes = dict(l.split() for l in out.splitlines())
    for t, mirror, f in fields:                      # the fields added in round 2 and a few old ones sit where ctypes puts them
        assert int(sizes.pop("%s.%s" % (t, f))) == getattr(getattr(_lib, mirror), f).offset, (t, f)
    assert set(sizes) == set(_STRUCTS)
    for name, mirror in _STRUCTS.items():
        assert int(sizes[name]) == ctypes.sizeof(getattr(_lib, mirror)), (name, mirror, sizes[name], ctypes.sizeof(getattr(_lib, mirror)))

