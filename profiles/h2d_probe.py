#!/usr/bin/env python
"""Host->device bandwidth per rank with all ranks copying at once, before and after binding the
process to its GPU's NUMA node (pangenomenem_b200.sharded.bind_to_gpu_numa).  Explains the e2e
load time of the row-sharded bench at 4-8 GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 profiles/h2d_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pangenomenem_b200 import sharded  # noqa: E402


def measure(dev, nbytes, reps=6):
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)                                   # touch every page
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(host, non_blocking=True); torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        d.copy_(host, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    return reps * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nbytes = 512 << 20
    before = measure(dev, nbytes)
    info = sharded.bind_to_gpu_numa(local)
    after = measure(dev, nbytes)
    rec = torch.tensor([before, after], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    infos = [None] * dist.get_world_size()
    dist.all_gather_object(infos, info)
    if dist.get_rank() == 0:
        print(json.dumps({"world": dist.get_world_size(), "bytes": nbytes,
                          "h2d_gbs_unbound": [round(float(o[0]), 1) for o in out],
                          "h2d_gbs_bound": [round(float(o[1]), 1) for o in out],
                          "binding": infos}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
