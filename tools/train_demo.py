"""Data-parallel training on N GPUs -- the whole network (--full: BASELINE config 4) or the regressor only (the
reference's `train_reg_only` mode with copenet's loss):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_demo.py [--pairs 32] [--steps 20]
Every rank trains on its own shard of a synthetic batch; the one collective of the step is the all-reduce of the
optimizer's flat gradient buffer (NCCL over NVLink).  Prints one JSON line on rank 0: pairs/s (device time, max over
ranks), first/last loss, and whether the parameters are still bit-identical across ranks."""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from argparse import Namespace
from airpose_b200 import parallel, synthetic
from airpose_b200.copenet_twoview import copenet_twoview


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32, help="pairs per GPU per step (BASELINE config 4: 32)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--full", action="store_true", help="train the whole network (trunk backward included), not only the regressor")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.pairs
    tmp = tempfile.mkdtemp(prefix="airpose_train_%d_" % rank)
    mp = synthetic.write_mean_params(os.path.join(tmp, "m.npz"))
    synthetic.write_smplx_model(tmp, 0)
    mod = copenet_twoview(Namespace(smpl_mean_params=mp, smplx_model_dir=tmp, batch_size=B, val_batch_size=B, reg_iters=3, lr=5e-5))
    mod.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synthetic.make_network_state(123, dec_gain=0.01).items()})
    mod = mod.to(dev)
    mod = mod.train() if args.full else mod.eval()
    opt = mod.configure_optimizers() if args.full else mod.configure_optimizers_reg_only()
    step = mod.training_step if args.full else mod.training_step_reg_only
    # global synthetic batch, sharded by pair (no overlap between ranks); GT = SMPL-X forward of an independent sample
    x = synthetic.make_inputs(B * world, 123)
    li = synthetic.make_lbs_inputs(B * world, seed=9)
    rng = np.random.default_rng(5)
    b0, b1 = parallel.shard_range(B * world, world, rank)
    with torch.no_grad():      # ground-truth meshes: the SMPL-X forward (the product's own kernels) of an independent seeded sample
        gt_out = mod.smplx.forward(betas=torch.from_numpy(li["betas"][b0:b1]).to(dev), body_pose=torch.from_numpy(li["body_pose"][b0:b1]).to(dev),
                                   pose2rot=False)
    gv, gj = gt_out.vertices.cpu().numpy(), gt_out.joints.cpu().numpy()
    r6 = lambda: synthetic.rot6d_to_rotmat_np(np.array([1, 0, 0, 1, 0, 0], np.float32) + rng.standard_normal((B, 6)).astype(np.float32) * 0.3)[:, None]
    gt = {"smplpose_rotmat": li["body_pose"][b0:b1], "smplorient_rel0": r6(), "smplorient_rel1": r6(), "smpl_vertices": gv[:, None],
          "smpl_joints": gj[:, None], "smpl_joints_2d0": (rng.standard_normal((B, 1, 127, 2)) * 50 + 500).astype(np.float32),
          "smpl_joints_2d1": (rng.standard_normal((B, 1, 127, 2)) * 50 + 500).astype(np.float32)}
    batch = {k: torch.from_numpy(np.ascontiguousarray(v[b0:b1])).to(dev) for k, v in x.items()}
    batch.update({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in gt.items()})
    losses = []
    for i in range(args.warmup):
        loss, _ = step(batch, opt)
        losses.append(loss)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss, _ = step(batch, opt)
        losses.append(loss)
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks([e0.elapsed_time(e1) / args.steps], device=dev)[0]
    # parameters must stay identical across ranks: compare a checksum of the flat buffer
    chk = torch.stack([opt.flat.double().sum(), opt.flat.double().abs().sum()])
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
    lv = torch.stack(losses).cpu().tolist()
    if rank == 0:
        print(json.dumps({"mode": "train_full" if args.full else "train_reg_only", "n_gpus": world, "pairs_per_gpu": B, "steps": args.steps, "ms_per_step": ms,
                          "pairs_per_s": world * B / (ms * 1e-3), "loss_first": lv[0], "loss_last": lv[-1],
                          "params_identical_across_ranks": same, "trainable_params": int(opt.numel),
                          "collective": "one all-reduce of the flat gradient buffer (%.1f MB) per step" % (opt.numel * 4 / 1e6)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
