"""One-rank NCCL init + all-reduce with bench.py's handling of the image's NCCL_DEBUG=VERSION: stdout must stay empty
(bench.py prints exactly one JSON line at N > 1).   python tools/nccl_banner_check.py | wc -c   -> 0"""
import os, sys
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29541")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ.pop("NCCL_DEBUG")
import torch, torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
dist.destroy_process_group()
