"""Quick device-side timings of the hot-path pieces (development aid; bench.py is the contract).

    python tools/gpu_probe.py [section ...]     sections: lbs trunk ief twoview gemm server preprocess
Each section runs in its own process so a device fault in one does not poison the others.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cuda_time(fn, warmup=3, iters=10):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def setup_net(tmp):
    import numpy as np
    import torch
    from airpose_b200 import synthetic
    from airpose_b200.model_copenet import getcopenet
    mp = synthetic.write_mean_params(os.path.join(tmp, "smpl_mean_params.npz"))
    net = getcopenet(mp, pretrained=False)
    sd = synthetic.make_network_state(123)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return net.cuda().eval()


def sec_lbs(tmp):
    import torch
    from airpose_b200 import synthetic
    from airpose_b200.smplx import SMPLX
    synthetic.write_smplx_model(tmp, 0)
    sm = SMPLX(tmp, batch_size=1, create_transl=False).cuda()
    for B in (64, 1024, 8192):
        li = synthetic.make_lbs_inputs(B, seed=1)
        betas, body = torch.from_numpy(li["betas"]).cuda(), torch.from_numpy(li["body_pose"]).cuda()
        ms = cuda_time(lambda: sm.forward(betas=betas, body_pose=body, pose2rot=False))
        gb = (B * 128420 + 68338900) / 1e9
        print("lbs B=%d: %.3f ms  %.1f k meshes/s  algorithmic %.1f GB/s" % (B, ms, B / ms, gb / (ms * 1e-3)), flush=True)


def sec_trunk(tmp):
    import torch
    net = setup_net(tmp)
    for n in (2, 32, 128):
        x = torch.randn(n, 3, 224, 224, device="cuda")
        ms = cuda_time(lambda: net.forward_feat_ext(x), warmup=3, iters=5)
        print("trunk n=%d images: %.3f ms  %.1f img/s  %.1f TFLOP/s" % (n, ms, n / ms * 1e3, n * 8.174e9 / (ms * 1e-3) / 1e12), flush=True)


def sec_ief(tmp):
    import torch
    net = setup_net(tmp)
    for B in (2, 64, 256):
        xf0, xf1 = torch.rand(B, 2048, device="cuda"), torch.rand(B, 2048, device="cuda")
        bb, pos = torch.rand(B, 3, device="cuda"), torch.rand(B, 3, device="cuda")
        ms = cuda_time(lambda: net._ief(xf0, xf1, bb, bb, pos, pos, None, None, None, None, 3))
        print("ief B=%d: %.3f ms" % (B, ms), flush=True)


def sec_twoview(tmp):
    import torch
    from argparse import Namespace
    import numpy as np
    from airpose_b200 import synthetic
    from airpose_b200.copenet_twoview import copenet_twoview
    mp = synthetic.write_mean_params(os.path.join(tmp, "smpl_mean_params.npz"))
    synthetic.write_smplx_model(tmp, 0)
    mod = copenet_twoview(Namespace(smpl_mean_params=mp, smplx_model_dir=tmp, batch_size=64, val_batch_size=64, reg_iters=3))
    sd = synthetic.make_network_state(123)
    mod.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    mod = mod.cuda().eval()
    for B in (2, 64):
        x = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_inputs(B, 1).items()}
        ms = cuda_time(lambda: mod.fwd_pass(x), warmup=3, iters=5)
        print("twoview B=%d pairs: %.3f ms  %.1f pairs/s" % (B, ms, B / ms * 1e3), flush=True)


def sec_gemm(tmp):
    import ctypes as C
    import torch
    from airpose_b200 import _lib
    lib = _lib.load()
    for M, N, K in ((8192, 8192, 8192), (401408, 64, 64), (401408, 256, 64), (100352, 128, 1152), (25088, 1024, 256), (6272, 512, 4608)):
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        Bm = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        a = _lib.GemmArgs()
        a.A, a.lda, a.B, a.ldb, a.M, a.N, a.K = A.data_ptr(), K, Bm.data_ptr(), K, M, N, K
        a.out_bf16, a.ldd = out.data_ptr(), N
        ms = cuda_time(lambda: _lib.check(lib.airpose_gemm_bf16(C.byref(a), _lib.current_stream())))
        ms_t = cuda_time(lambda: torch.matmul(A, Bm.t()))
        print("gemm %dx%dx%d: %.3f ms %.1f TFLOP/s  (torch.matmul %.3f ms %.1f TFLOP/s)" %
              (M, N, K, ms, 2.0 * M * N * K / ms / 1e9, ms_t, 2.0 * M * N * K / ms_t / 1e9), flush=True)


def sec_server(tmp):
    """Per-message latency of the staged drone server (host wall clock around process(): H2D + device work + D2H + sync),
    eager and CUDA-graph replay, against the slot budgets of airpose.yaml:9-11 (45 ms stage 0, 2.5 ms stages 1-2)."""
    import numpy as np
    import torch
    from airpose_b200 import server, synthetic
    sd = synthetic.make_network_state(123)
    mp = synthetic.write_mean_params(os.path.join(tmp, "smpl_mean_params.npz"))
    msgs = synthetic.server_messages(17, 1, sd["init_pose"], sd["init_shape"])
    for graph in (False, True):
        model = server.getmodel(mp)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        try:
            srv = server.StagedServer(model, graph=graph)
            for _ in range(5):
                for stage, data in msgs:
                    srv.process(data, None, stage)
            lat = {0: [], 1: [], 2: []}
            for _ in range(100):
                for stage, data in msgs:
                    t0 = time.perf_counter()
                    srv.process(data, None, stage)
                    lat[stage].append((time.perf_counter() - t0) * 1e3)
            print("server graph=%s: " % graph + "  ".join("stage %d median %.3f ms p99 %.3f ms" % (k, float(np.median(v)), float(np.percentile(v, 99)))
                                                          for k, v in lat.items()), flush=True)
        except Exception as e:      # report, keep the other mode's numbers
            print("server graph=%s FAILED: %r" % (graph, e), flush=True)


def sec_preprocess(tmp):
    import numpy as np
    import torch
    from airpose_b200 import synthetic
    from airpose_b200.preprocess import crop_resize_pad_normalize
    n = 128
    frames = torch.from_numpy(np.stack([synthetic.camera_frame(1080, 1920, s) for s in range(4)])).cuda()
    frames = frames.repeat(n // 4, 1, 1, 1)                     # 128 full-HD BGR frames = 796 MB
    rects = [(100 + (i % 7) * 10, 900 - (i % 5) * 20, 600 + (i % 3) * 30, 1300 + (i % 4) * 25) for i in range(n)]
    out = torch.empty(n, 3, 224, 224, device="cuda")
    ms = cuda_time(lambda: crop_resize_pad_normalize(frames, rects, out=out))
    print("crop_resize_pad_normalize n=%d crops from 1080p frames: %.3f ms  %.1f k images/s  output %.1f GB/s" %
          (n, ms, n / ms, n * 3 * 224 * 224 * 4 / 1e9 / (ms * 1e-3)), flush=True)


SECTIONS = {"server": sec_server, "preprocess": sec_preprocess, "lbs": sec_lbs, "trunk": sec_trunk, "ief": sec_ief, "twoview": sec_twoview, "gemm": sec_gemm}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        import tempfile
        SECTIONS[sys.argv[2]](tempfile.mkdtemp(prefix="airpose_probe_"))
        sys.exit(0)
    for name in (sys.argv[1:] or list(SECTIONS)):
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], timeout=600)
        print("[%s] exit %d in %.1fs" % (name, r.returncode, time.time() - t0), flush=True)
