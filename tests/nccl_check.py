"""Real multi-GPU check (not collected by pytest; run under torchrun on N >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/nccl_check.py

Every rank fits the same seeded pangenome row-sharded over NCCL; rank 0 also
fits it on one GPU.  The sharded labels, theta, iteration count and criteria must equal the
single-GPU fit exactly (integer statistics, exact sequential sweep across ranks)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import torch.distributed as dist
    from pangenomenem_b200 import capi, sharded, synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    report = []
    # 64000 and 40000 families split on 16-family boundaries: the persistent row-sharded kernel
    # (labels and requests through peer memory); 60001 does not: the launch-per-stage loop (NCCL)
    for n, d, graph, algo, update in [(64000, 120, "pangenome", "ncem", "seq"),
                                      (60001, 200, "pangenome", "ncem", "seq"),
                                      (40000, 96, "random", "ncem", "seq"),
                                      (256000, 64, "pangenome", "ncem", "seq"),
                                      (30000, 64, "pangenome", "nem", "para")]:
        only = os.environ.get("NCCL_CHECK_ONLY")
        if only and str(n) not in only.split(","):
            continue
        print(f"[rank {rank}] case n={n} d={d} {graph} {algo} {update}", file=sys.stderr, flush=True)
        pg = synth.make_pangenome(n, d, seed=7, graph=graph)
        theta = synth.default_theta(3, d)
        kw = dict(k=3, algo=algo, update=update, disp="sk_", prop="pk", beta=0.5,
                  it_max=30 if algo == "nem" else 100)
        eng, comm = sharded.make_engine(dist, local)
        xp = synth.pack_rows(pg.x)
        p = sharded.plan(n, world, rank)
        fit = sharded.fit_sharded(eng, xp[p.rows], n, d, pg.row_ptr, pg.col, pg.wgt, theta, rank, world, **kw)
        print(f"[rank {rank}] fitted: iters={fit.iters} pk={fit.pk} ms={fit.fit_ms:.2f}", file=sys.stderr, flush=True)
        lab, t = eng.labels(), eng.posteriors()
        eng.close()
        capi.comm_destroy(comm)
        ok = True
        if rank == 0:
            one = capi.Engine(local)
            one.load_packed(xp, d, pg.row_ptr, pg.col, pg.wgt)
            ref = one.fit(*theta, **kw)
            rlab, rt = one.labels(), one.posteriors()
            one.close()
            same_t = np.array_equal(t, rt) if algo == "ncem" else bool(np.allclose(t, rt, rtol=1e-6, atol=1e-12))
            ok = (fit.iters == ref.iters and np.array_equal(lab, rlab) and same_t
                  and np.array_equal(fit.center, ref.center)
                  and all(abs(fit.crit[c] - ref.crit[c]) <= 1e-9 * abs(ref.crit[c]) for c in "UDL"))
            report.append(dict(n=n, d=d, graph=graph, algo=algo, update=update, world=world,
                               iters=fit.iters, ref_iters=ref.iters, exchanges=fit.exchanges,
                               persistent_launches=fit.pk.get("launches"), barriers=fit.pk.get("barriers"),
                               fit_ms=round(fit.fit_ms, 3), single_gpu_ms=round(ref.fit_ms, 3),
                               label_mismatches=int((lab != rlab).sum()), ok=bool(ok)))
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) != 1:
            if rank == 0:
                print(json.dumps(report))
            raise SystemExit(1)
    if rank == 0:
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
