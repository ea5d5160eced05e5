#!/usr/bin/env python
"""Phase profile of the row-sharded persistent kernel (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 profiles/pk_shard_probe.py [families_per_gpu] [genomes]
Weak-scaling pangenome built like bench.py's (own graph per shard + 2000 chords per boundary)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pangenomenem_b200 import capi, sharded, synth, synth_gpu  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    xdev, _ = synth_gpu.make_packed_on_device(n, d, seed=42 + rank, device=dev)
    xh = xdev.cpu().numpy()
    row_ptr, col, wgt = bench.build_global_graph(torch, dist, dev, rank, world, n, xh, 42, "pangenome")
    eng, comm = sharded.make_engine(dist, local)
    p = sharded.plan(n * world, world, rank)
    eng.load_shard_device(xdev.data_ptr(), n * world, p.row0, n, d, xdev.shape[1], row_ptr, col, wgt)
    theta0 = synth.default_theta(3, d)
    opts = dict(k=3, algo="ncem", update="seq", beta=0.5, conv="clas", conv_thr=1e-8, it_max=100,
                prop="pk", disp="sk_", sweep_impl="auto")
    for _ in range(3):
        eng.fit(*theta0, **opts)
    ms = []
    for _ in range(5):
        dist.barrier(); torch.cuda.synchronize()
        f = eng.fit(*theta0, **opts)
        ms.append(f.fit_ms)
    prof = eng.persist_profile()
    tr = eng.persist_trace()
    if rank == 0:
        print(json.dumps({"world": world, "families_per_gpu": n, "fit_ms": round(float(np.median(ms)), 4),
                          "iters": f.iters, "launches": f.kernel_launches, "pk": f.pk,
                          "cut_edges": sharded.cut_edges(row_ptr, col, world),
                          "phase_us": {k: round(v, 1) for k, v in prof.items()}}))
        print("   iter:   scan  delta  final margin   eval  fixup | active rounds   (us, rank 0, last launch)")
        for i, row in enumerate(tr):
            if row[:6].sum() > 0:
                print("   %4d: %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f | %7d %3d" % ((i,) + tuple(row[:6]) + (int(row[6]), int(row[7]))))
    eng.close()
    capi.comm_destroy(comm)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
