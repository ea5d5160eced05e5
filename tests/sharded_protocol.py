"""TEST INFRASTRUCTURE: a numpy model of the engine's row-sharded E-step protocol
(pangenomenem_b200/csrc/nem_fit.c run_sweep, world > 1), rank-local work in Python and the
exchanges through torch.distributed (gloo on CPU).  It exists to check, without a GPU, that the
protocol -- Jacobi round, local fix-ups, label all-gather, re-queue the readers of every remote
label that moved, stop when no rank queued anything -- lands on the SEQUENTIAL sweep's labels."""
from __future__ import annotations

import numpy as np


def site_argmax(logpf_i, ctx, beta):
    sc = logpf_i + beta * ctx
    return int(np.argmax(sc))          # first maximum, like ComputeMAP TIE_FIRST


def ctx_of(i, lab_lo, lab_hi, row_ptr, col, wgt, k):
    """ctx_k = sum_j w_ij [label_j = k], label_j = lab_lo[j] for j < i else lab_hi[j]."""
    ctx = np.zeros(k)
    for e in range(row_ptr[i], row_ptr[i + 1]):
        j = col[e]
        l = lab_lo[j] if j < i else lab_hi[j]
        if l < k:
            ctx[l] += float(wgt[e])
    return ctx


def sharded_seq_sweep(dist, plan, logpf_local, lab_old, row_ptr, col, wgt, beta, k):
    """Returns (new labels of ALL families, number of label exchanges)."""
    n, row0, row1 = plan.n_glob, plan.row0, plan.row0 + plan.n_loc
    import torch
    cur = lab_old.copy()
    seen = lab_old.copy()
    dirty = set()

    def mark_readers(i):
        for e in range(row_ptr[i], row_ptr[i + 1]):     # symmetric graph: readers = neighbours
            j = col[e]
            if j > i and row0 <= j < row1:
                dirty.add(j)

    # round 0: Jacobi on the old labels
    for i in range(row0, row1):
        km = site_argmax(logpf_local[i - row0], ctx_of(i, lab_old, lab_old, row_ptr, col, wgt, k), beta)
        cur[i] = km
        if km != lab_old[i]:
            mark_readers(i)

    def local_fixups():
        while dirty:
            work = sorted(dirty)
            dirty.clear()
            for i in work:
                km = site_argmax(logpf_local[i - row0], ctx_of(i, cur, lab_old, row_ptr, col, wgt, k), beta)
                if km != cur[i]:
                    cur[i] = km
                    mark_readers(i)

    local_fixups()
    exchanges = 0
    while True:
        # label exchange: every rank's slice (padded to shard_len)
        mine = np.full(plan.shard_len, 255, dtype=np.uint8)
        mine[:plan.n_loc] = cur[row0:row1]
        parts = [torch.zeros(plan.shard_len, dtype=torch.uint8) for _ in range(plan.world)]
        dist.all_gather(parts, torch.from_numpy(mine))
        allv = torch.cat(parts).numpy()[:n]
        cur[:row0] = allv[:row0]
        cur[row1:] = allv[row1:]
        exchanges += 1
        for j in np.flatnonzero(cur != seen):
            if row0 <= j < row1:
                continue
            seen[j] = cur[j]
            for e in range(row_ptr[j], row_ptr[j + 1]):
                i = col[e]
                if i > j and row0 <= i < row1:
                    dirty.add(i)
        pending = torch.tensor([len(dirty)])
        dist.all_reduce(pending)
        if int(pending) == 0:
            return cur, exchanges
        local_fixups()
