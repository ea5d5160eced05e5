"""Row-sharded NEM fits: one process per GPU, plumbing by ``torch.distributed``.

The engine (libnem_b200.so) does the sharded EM itself -- density and M-step statistics on the
rank's rows of X, speculative sequential sweep with label exchanges, rank-ordered sums -- through
ONE primitive, an all-gather on device pointers (``nemb_comm`` in include/nem_b200.h).  This
module only

* splits the families into the contiguous id ranges the engine expects (:func:`plan`),
* bootstraps the engine's NCCL communicator: rank 0 draws the 128-byte NCCL id and
  ``torch.distributed`` (any backend; NCCL on GPUs, gloo in the CPU tests) broadcasts it
  (:func:`exchange_unique_id`, :func:`make_engine`),
* and offers :func:`fit_sharded` as the one call a user makes per rank.

The reference has no counterpart: its engine is a single-threaded process and PPanGGOLiN's only
parallelism is process-level replicas of whole runs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass(frozen=True)
class ShardPlan:
    n_glob: int
    world: int
    rank: int
    shard_len: int
    row0: int
    n_loc: int

    @property
    def rows(self) -> slice:
        return slice(self.row0, self.row0 + self.n_loc)


def plan(n_glob: int, world: int, rank: int) -> ShardPlan:
    """Contiguous id ranges of ceil(N/world) families, rounded up to 16 when world > 1 and a shard has 1024 families or more (the last
    ranks may hold fewer, even none).

    Family ids follow chromosome order in PPanGGOLiN (node insertion order, ppanggolin.py:481-517),
    so contiguous ranges keep neighbours together and the cut is a few edges per boundary: this is
    the graph-aware partition for this input order.  Pure Python mirror of nemb_shard_range."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    shard_len = (n_glob + world - 1) // world
    if world > 1 and shard_len >= 1024:
        shard_len = (shard_len + 15) // 16 * 16      # shards start on 16-family boundaries
    row0 = min(rank * shard_len, n_glob)
    n_loc = max(0, min(shard_len, n_glob - rank * shard_len))
    return ShardPlan(n_glob, world, rank, shard_len, row0, n_loc)


def cut_edges(row_ptr: np.ndarray, col: np.ndarray, world: int) -> int:
    """Directed CSR entries whose endpoints live on different ranks (halo size diagnostic)."""
    n = row_ptr.shape[0] - 1
    shard_len = plan(n, world, 0).shard_len
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    return int(np.count_nonzero(src // shard_len != col // shard_len))


def bind_to_gpu_numa(gpu_index: int) -> dict:
    """Pin this process to the CPU cores next to its GPU (NVML's ideal CPU set), BEFORE it
    allocates pinned host buffers: cudaHostAlloc places pages on the calling thread's NUMA node,
    and on a two-socket 8-GPU box every rank that stages its shard on the wrong socket pulls it
    over the inter-socket link -- measured: host->device 54 GB/s per GPU at 1-2 ranks, 23 GB/s at
    8 ranks without binding.  Returns what was done (for the bench line); never raises."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = ideal & allowed
        info.update(ideal_cpus=len(ideal), allowed_cpus=len(allowed))
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus), first_cpu=min(cpus))
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
    except Exception as exc:                      # no NVML / restricted container: run unbound
        info["error"] = f"{type(exc).__name__}: {exc}"[:120]
    return info


def exchange_unique_id(dist, src: int = 0, device=None) -> bytes:
    """Rank `src` draws the NCCL unique id; every rank returns the same 128 bytes."""
    import torch
    rank = dist.get_rank()
    if rank == src:
        buf = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8)
    else:
        buf = torch.zeros(128, dtype=torch.uint8)
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().tolist())


def make_engine(dist, local_device: int):
    """Engine + NCCL communicator of this rank (call after torch.cuda.set_device)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", local_device) if dist.get_backend() == "nccl" else None
    uid = exchange_unique_id(dist, 0, dev)
    eng = capi.Engine(local_device)
    comm = capi.nccl_comm(uid, rank, world)
    eng.set_comm(comm)
    return eng, comm


def fit_sharded(eng, x_packed_local, n_glob: int, d: int, row_ptr, col, wgt, theta0, rank: int,
                world: int, **options):
    """Load this rank's rows + the global graph and run the fit; every rank returns the same Fit
    and, through eng.labels(), the labels of ALL families."""
    p = plan(n_glob, world, rank)
    assert x_packed_local.shape[0] == p.n_loc, (x_packed_local.shape, p)
    eng.load_shard(x_packed_local, n_glob, p.row0, d, row_ptr, col, wgt)
    return eng.fit(*theta0, **options)
