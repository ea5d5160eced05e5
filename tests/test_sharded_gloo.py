"""CPU, world_size 2 (gloo): host-side logic of the row-sharded path -- the shard plan, the
bootstrap of the engine's NCCL id over torch.distributed, the exactness of the cross-rank sweep
protocol and of the rank-ordered statistic sums -- checked against the oracle.  The CUDA side of
the same path is covered on one device by tests/test_gpu_sharded.py."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    from oracle import nemo
    from pangenomenem_b200 import sharded, synth
    import sharded_protocol as proto
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        # (1) the NCCL id drawn by rank 0 reaches every rank unchanged
        uid = sharded.exchange_unique_id(dist, 0)
        gathered = [None] * world
        dist.all_gather_object(gathered, uid)
        res["uid_same"] = all(g == gathered[0] for g in gathered) and len(uid) == 128 and any(uid)

        # (2) sharded EM = oracle fit (ncem, sequential sweep, sk_ pk), protocol in numpy
        pg = synth.make_pangenome(1501, 24, seed=3, graph="pangenome")
        k, beta = 3, 0.7
        p = sharded.plan(pg.n, world, rank)
        full = nemo.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, k=k, algo="ncem", beta=beta)
        local = nemo.Problem(pg.x[p.rows], k=k, algo="ncem")
        ref = full.fit(*nemo.default_theta(k, pg.d))
        prop, center, disp = nemo.default_theta(k, pg.d)
        lab = np.full(pg.n, 255, dtype=np.uint8)
        logpf = local.logpf(prop, center, disp) * 0.2      # weaken the data term: labels must move
        # blind sweep (beta 0) then coupled sweep, as ComputePartitionFromPara does
        lab, _ = proto.sharded_seq_sweep(dist, p, logpf, lab, pg.row_ptr, pg.col, pg.wgt, 0.0, k)
        t_full0 = np.zeros((pg.n, k), dtype=np.float32)
        logpf_full = full.logpf(prop, center, disp) * 0.2
        t1, lab1 = full.sweep(logpf_full, 0.0, t_full0)
        res["blind_equal"] = bool(np.array_equal(lab, lab1))
        total_x = 0
        pend_log = []
        for it in range(3):
            # cap 64 then 4: the second value forces the whole-slice fallback
            lab, nx = proto.sharded_seq_sweep(dist, p, logpf, lab, pg.row_ptr, pg.col, pg.wgt, beta, k,
                                              cap=64 if it != 1 else 4, check=pend_log)
            total_x += nx
            t1, lab1 = full.sweep(logpf_full, beta, t1)
            res[f"sweep{it}_equal"] = bool(np.array_equal(lab, lab1))
            res[f"sweep{it}_moved"] = int((lab1 != np.argmax(logpf_full, axis=1)).sum())
        res["exchanges"] = total_x
        logs = [None] * world
        dist.all_gather_object(logs, pend_log)              # test only: the engine exchanges no counters
        res["pending_agree"] = all(l == logs[0] for l in logs) and any(v > 0 for v in logs[0])

        # (3) M-step statistics: rank-ordered sum of the per-rank integer counts == full counts
        t_loc = np.eye(k, dtype=np.float32)[lab[p.rows]]
        _, _, _, _, nk_loc, s_loc = local.mstep(t_loc, prop, center, disp)
        stat = torch.from_numpy(np.concatenate([s_loc.ravel(), nk_loc]))
        parts = [torch.zeros_like(stat) for _ in range(world)]
        dist.all_gather(parts, stat)
        tot = sum(parts[1:], parts[0].clone()).numpy()       # ranks added in order
        _, _, _, _, nk_full, s_full = full.mstep(np.eye(k, dtype=np.float32)[lab], prop, center, disp)
        res["stats_equal"] = bool(np.array_equal(tot[:-k], s_full.ravel()) and np.array_equal(tot[-k:], nk_full))
        res["cut_edges"] = sharded.cut_edges(pg.row_ptr, pg.col, world)
        res["ref_iters"] = ref.iters
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_world2_protocol_matches_oracle():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = dict(q.get(timeout=200) for _ in procs)
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank in (0, 1):
        r = out[rank]
        assert r["uid_same"] and r["blind_equal"] and r["stats_equal"] and r["pending_agree"], r
        assert all(r[f"sweep{i}_equal"] for i in range(3)), r
        assert r["sweep0_moved"] > 0, "the coupled sweep should move labels in this test"
        assert r["exchanges"] >= 3 and r["cut_edges"] > 0


def test_plan_covers_every_family_once():
    from pangenomenem_b200 import capi, sharded
    for n, world in [(1, 1), (9, 4), (1000, 3), (1_000_000, 8), (7, 8)]:
        seen = np.zeros(n, dtype=np.int32)
        for r in range(world):
            p = sharded.plan(n, world, r)
            assert (p.shard_len, p.row0, p.n_loc) == capi.shard_range(n, world, r)
            assert p.shard_len * world >= n
            seen[p.rows] += 1
        assert (seen == 1).all()
    with pytest.raises(ValueError):
        sharded.plan(10, 2, 2)
