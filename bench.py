"""Benchmark of the AirPose hot path: two-view frame-pairs/sec of the copenet_twoview forward
(two ResNet-50 trunks, 3-iteration IEF regressor, SMPL-X, rigid transform, reprojection).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One JSON line on stdout (rank 0).  Workload = BASELINE.json configs[1]: 64 pairs per GPU,
224x224, bf16 trunk / fp32 SMPL-X; with N GPUs the batch of pairs is sharded, 64 per rank,
no data-path collective (weak scaling).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "two-view frame-pairs/sec copenet_twoview fwd+SMPLX"
UNIT = "pairs/s"
PAIRS_PER_GPU = 64
GFLOP_PER_IMAGE = 8.174272512          # 4,087,136,256 MAC
LBS_BYTES_PER_MESH = 128420
LBS_CONST_BYTES = 68338900
CPU_SAMPLE_PAIRS = 8
# dram__bytes_read.sum + dram__bytes_write.sum summed over the 77 conv-GEMM launches of one 128-image trunk
# forward (64 pairs), from the ncu --set full capture profiles/r01s_ncu_full_trunk_128img.csv (cold-cache, serialised)
TRUNK_DRAM_BYTES_PER_64_PAIRS = 4954.3e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(pairs, steps, warmup):
    """The reference algorithm (PyTorch-CPU port, oracle/torch_port.py) on all host threads."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port as tp
    from airpose_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = tp.to_torch(synthetic.make_network_state(123))
    m = tp.Smplx(synthetic.make_smplx_model(0))
    x = {k: torch.from_numpy(v) for k, v in synthetic.make_inputs(pairs, 123).items()}
    with torch.no_grad():
        for _ in range(warmup):
            tp.twoview_forward(sd, m, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            tp.twoview_forward(sd, m, x)
        dt = time.perf_counter() - t0
    return pairs * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores = cpu_reference(CPU_SAMPLE_PAIRS, args.steps, args.warmup)
    sample = "%d pairs per step (bounded sample of the 64-pair workload), fp32, PyTorch-CPU port of the reference" % CPU_SAMPLE_PAIRS
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "copenet_twoview fwd batch=64 pairs 224x224 (BASELINE.json configs[1])", "pairs_per_step": CPU_SAMPLE_PAIRS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from argparse import Namespace
    from airpose_b200 import _lib, synthetic
    from airpose_b200.copenet_twoview import copenet_twoview

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: airpose_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # rank 0 prints ONE JSON line: this image exports NCCL_DEBUG=VERSION, which makes NCCL print "NCCL version ..." on stdout
    # at communicator creation (NCCL_DEBUG_FILE does not catch it).  Any other setting (WARN, INFO) is the caller's choice.
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    B = args.pairs
    tmp = tempfile.mkdtemp(prefix="airpose_bench_%d_" % rank)
    mp = synthetic.write_mean_params(os.path.join(tmp, "smpl_mean_params.npz"))
    synthetic.write_smplx_model(tmp, 0)
    mod = copenet_twoview(Namespace(smpl_mean_params=mp, smplx_model_dir=tmp, batch_size=B, val_batch_size=B, reg_iters=3))
    mod.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synthetic.make_network_state(123).items()})
    mod = mod.to(dev).eval()

    # synthetic inputs: NSETS distinct batches rotated so a step's inputs are never L2-resident
    NSETS = 4
    gen = torch.Generator(device=dev).manual_seed(123 + rank)
    intr = torch.tensor([[1475.0, 0, 960.0], [0, 1475.0, 540.0], [0, 0, 1.0]], device=dev).expand(B, 3, 3).contiguous()

    def make_set():
        s = {"intr0": intr, "intr1": intr}
        for v in (0, 1):
            s["im%d" % v] = torch.randn(B, 3, 224, 224, device=dev, generator=gen)
            bb = torch.rand(B, 3, device=dev, generator=gen)
            bb[:, :2] = bb[:, :2] * 2 - 1
            bb[:, 2] = bb[:, 2] * 1.9 + 0.1
            s["bb%d" % v] = bb
        return s

    sets = [make_set() for _ in range(NSETS)]
    set_bytes = 2 * B * 3 * 224 * 224 * 4

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident throughput
    prof = []
    for i in range(args.warmup):
        mod.fwd_pass(sets[i % NSETS])
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        p = {}
        mod.fwd_pass(sets[i % NSETS], profile=p)
        prof.append(p)
    e1.record()
    sync_all()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1) / args.steps
    trunk_ms = sum(p["trunk"][0].elapsed_time(p["trunk"][1]) for p in prof) / len(prof)
    ief_ms = sum(p["ief"][0].elapsed_time(p["ief"][1]) for p in prof) / len(prof)
    smplx_ms = sum(p["smplx"][0].elapsed_time(p["smplx"][1]) for p in prof) / len(prof)

    # ------------------------------------------------------------------ end to end from pinned host memory
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    out_keys = ("pred_pose", "pred_betas", "pred_vertices_cam", "pred_joints_cam", "pred_joints_2d_cam")

    def run_e2e(host_sets, prepare):
        """Timed pipeline: H2D of step i+1 (copy stream) | prepare + fwd_pass of step i | D2H of step i-1 (third stream).
        Returns (ms per step, h2d bytes per step, d2h bytes per step)."""
        host_out = None
        h2d = sum(v.numel() * v.element_size() for v in host_sets[0].values())

        def stage(k):
            with torch.cuda.stream(copy_stream):
                s = {n: v.to(dev, non_blocking=True) for n, v in host_sets[k % 2].items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return (s, ev)

        def e2e_step(i, staged):
            nonlocal host_out
            nxt = stage(i + 1)                      # stage step i+1 on the copy stream while step i computes
            torch.cuda.current_stream().wait_event(staged[1])
            for tns in staged[0].values():
                tns.record_stream(torch.cuda.current_stream())
            out = mod.fwd_pass(prepare(staged[0]))
            res = {k + str(v): out[k + str(v)] for k in out_keys for v in (0, 1)}
            if host_out is None:
                host_out = [{k: torch.empty(t.shape, dtype=t.dtype).pin_memory() for k, t in res.items()} for _ in range(2)]
            # results leave on their own stream so the D2H of step i overlaps the compute of step i+1
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                for k, t in res.items():
                    t.record_stream(d2h_stream)
                    host_out[i % 2][k].copy_(t, non_blocking=True)
            return nxt

        staged = stage(0)
        for i in range(max(args.warmup, 2)):
            staged = e2e_step(i, staged)
        sync_all()
        d2h = sum(t.numel() * t.element_size() for t in host_out[0].values())
        e0.record()
        for i in range(args.steps):
            staged = e2e_step(i, staged)
        torch.cuda.current_stream().wait_stream(d2h_stream)      # the last step's results must be on the host before the clock stops
        e1.record()
        sync_all()
        return e0.elapsed_time(e1) / args.steps, h2d, d2h

    # (1) the reference's batch format: fp32 normalised images in pinned host memory (copenet_twoview.py:166-183)
    host_sets = [{k: v.cpu().pin_memory() for k, v in s.items()} for s in sets[:2]]
    e2e_ms, h2d, d2h = run_e2e(host_sets, lambda s: s)
    # what the host link of THIS box delivers for exactly that copy (explains e2e when it is PCIe-bound)
    with torch.cuda.stream(copy_stream):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(copy_stream)
        for _ in range(3):
            tmp_dev = {k: v.to(dev, non_blocking=True) for k, v in host_sets[0].items()}
        c1.record(copy_stream)
    copy_stream.synchronize()
    h2d_gbs = 3 * h2d / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del tmp_dev
    # (2) the wire / dataset format: u8 BGR 224 x 224 crops (airpose_server/server.py:38,91-98), normalised on the device by
    # airpose_preprocess_bgr8 inside the timed region -- a quarter of the bytes over the host link
    from airpose_b200.preprocess import bgr8_to_normalized
    host_sets_u8 = []
    for s in sets[:2]:
        hs = {k: v.cpu().pin_memory() for k, v in s.items() if not k.startswith("im")}
        for v in (0, 1):
            hs["im%d" % v] = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(7 + v)).pin_memory()
        host_sets_u8.append(hs)

    def prepare_u8(sd):
        return {k: (bgr8_to_normalized(v) if k.startswith("im") else v) for k, v in sd.items()}

    e2e_u8_ms, h2d_u8, _ = run_e2e(host_sets_u8, prepare_u8)
    clocks = sampler.stop() if rank == 0 else None

    # ------------------------------------------------------------------ SMPL-X lbs() roofline at config 3 (rank 0, N=1)
    lbs = None
    if world == 1:
        nb = 8192
        li = synthetic.make_lbs_inputs(nb, seed=1)
        betas, body = torch.from_numpy(li["betas"]).to(dev), torch.from_numpy(li["body_pose"]).to(dev)
        for _ in range(3):
            mod.smplx.forward(betas=betas, body_pose=body, pose2rot=False)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            mod.smplx.forward(betas=betas, body_pose=body, pose2rot=False)
        e1.record()
        torch.cuda.synchronize()
        lbs_ms = e0.elapsed_time(e1) / 5
        lbs = (nb, lbs_ms)

    from airpose_b200 import parallel
    ms, e2e_ms, trunk_ms, e2e_u8_ms = parallel.max_over_ranks([ms, e2e_ms, trunk_ms, e2e_u8_ms], device=dev)
    if rank == 0:
        peaks = load_peaks()
        value = world * B / (ms * 1e-3)
        tf = 2 * B * GFLOP_PER_IMAGE / (trunk_ms * 1e-3) / 1e3
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "copenet_twoview fwd batch=64 pairs 224x224 bf16 trunk / fp32 SMPL-X (BASELINE.json configs[1])",
                       "pairs_per_gpu": B, "global_pairs": world * B, "reg_iters": 3, "parallelism": "batch-sharded x%d, no collective" % world,
                       "l2": "inputs rotate over %d distinct batches (%.0f MB) so no step's inputs are L2-resident" % (NSETS, NSETS * set_bytes / 1e6)},
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "h2d_gbs_measured": h2d_gbs, "h2d_bound_value": world * B / (h2d / (h2d_gbs * 1e9)),
                    "note": "pinned host fp32 images -> H2D (double-buffered on a copy stream) -> fwd_pass -> D2H of pose/betas/vertices_cam/joints_cam/joints_2d into pinned buffers on a third stream; the timed region ends with a device-wide synchronize, so every copy is inside it. h2d_gbs_measured = this box's host->device rate for the same copy; h2d_bound_value = the pairs/s that rate alone allows (e2e is PCIe-bound when it is below `value`)"},
            "e2e_u8": {"value": world * B / (e2e_u8_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                       "note": "same pipeline from u8 BGR crops (the drone server's wire format), normalised on the device by airpose_preprocess_bgr8 inside the timed region"},
            "gpu_launches": launches,
            "stage_ms": {"trunk": trunk_ms, "ief": ief_ms, "smplx_x2": smplx_ms},
            "roofline": {"bound": "tensor", "kernel": "gemm_tma_kernel / gemm_sk_kernel (tcgen05 implicit-GEMM convs of the ResNet-50 trunk, both views; 77 launches per 128 images)",
                         "achieved": tf, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tflops_sustained"],
                         "traffic": TRUNK_DRAM_BYTES_PER_64_PAIRS * B / 64.0,
                         "traffic_note": "bytes per step summed over the trunk's GEMM launches (ncu, profiles/r01s_ncu_full_trunk_128img.csv); algorithmic HBM bytes 56.4 MB/image",
                         "peak_source": peaks["source"] + ", sustained bf16"},
            "clocks": clocks,
        }
        if lbs:
            nb, lbs_ms = lbs
            gbs = (nb * LBS_BYTES_PER_MESH + LBS_CONST_BYTES) / (lbs_ms * 1e-3) / 1e9
            out["roofline_lbs"] = {"bound": "hbm", "kernel": "smplx_vertex_tc_kernel (+pose/joints kernels), lbs() batch=%d" % nb,
                                   "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                   "meshes_per_s": nb / (lbs_ms * 1e-3), "traffic": 1095.5e6,
                                   "traffic_note": "dram read+write of smplx_vertex_tc_kernel per launch at B=8192 (ncu, profiles/r01s_ncu_full_lbs_b8192.csv)",
                                   "peak_source": peaks["source"]}
            if not args.no_cpu_baseline:
                v, cms, cores = cpu_reference(CPU_SAMPLE_PAIRS, 3, 1)
                out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                       "sample": "%d pairs per step x 3 steps (1 warm-up), fp32, PyTorch-CPU port of the reference (oracle/torch_port.py)" % CPU_SAMPLE_PAIRS}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
