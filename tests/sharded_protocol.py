"""TEST INFRASTRUCTURE: a numpy model of the engine's row-sharded E-step protocol
(pangenomenem_b200/csrc/nem_fit.c run_sweep, world > 1), rank-local work in Python and the
exchanges through torch.distributed (gloo on CPU).  It exists to check, without a GPU, that the
protocol -- Jacobi round, local fix-ups, exchange of the labels that moved, re-queue the readers of
every remote label that moved, stop when every rank finds no cross-rank work left -- lands on the
SEQUENTIAL sweep's labels."""
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


def sharded_seq_sweep(dist, plan, logpf_local, lab_old, row_ptr, col, wgt, beta, k, cap=64, check=None):
    """Returns (new labels of ALL families, number of label exchanges).

    Mirrors run_sweep for world > 1: after the local rounds every rank publishes the labels of ITS
    rows that moved since the previous exchange -- a block of at most `cap` (family, label) pairs
    (k_delta_pack), or its whole slice when some rank moved more (fallback, k_mark_remote).  Every
    rank applies all blocks: remote labels go into its copy, `seen` is refreshed for every moved
    label, the later readers it owns are queued, and the cross-rank (reader, label) pairs are
    counted -- from data every rank holds alike, so NO counter is exchanged to agree on
    termination.  `check` (a list) receives the pending count of every round so that the test can
    assert the ranks computed the same numbers."""
    n, row0, row1 = plan.n_glob, plan.row0, plan.row0 + plan.n_loc
    import torch
    cur = lab_old.copy()
    seen = lab_old.copy()          # labels all ranks last saw (== the input labels at sweep start)
    dirty = set()
    sl = plan.shard_len

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

    def apply_moved(j, lab):
        """label j moved (on whatever rank owns it): same bookkeeping on every rank"""
        pend = 0
        mine_j = row0 <= j < row1
        if not mine_j:
            cur[j] = lab
        seen[j] = lab
        own = j // sl
        for e in range(row_ptr[j], row_ptr[j + 1]):
            i = col[e]
            if i <= j or i // sl == own:
                continue
            pend += 1
            if not mine_j and row0 <= i < row1:
                dirty.add(i)
        return pend

    local_fixups()
    exchanges = 0
    while True:
        # sparse block: (family, label) pairs of my rows that moved since the last exchange
        moved = [i for i in range(row0, row1) if cur[i] != seen[i]]
        block = np.full(2 + 2 * cap, -1, dtype=np.int64)
        block[0] = len(moved)
        for q, i in enumerate(moved[:cap]):
            block[2 + 2 * q], block[3 + 2 * q] = i, cur[i]
        parts = [torch.zeros(block.size, dtype=torch.int64) for _ in range(plan.world)]
        dist.all_gather(parts, torch.from_numpy(block))
        exchanges += 1
        blocks = [p_.numpy() for p_ in parts]
        pending = 0
        if any(b[0] > cap for b in blocks):
            # overflow somewhere: nothing applied, exchange whole slices instead
            mine = np.full(sl, 255, dtype=np.uint8)
            mine[:plan.n_loc] = cur[row0:row1]
            slices = [torch.zeros(sl, dtype=torch.uint8) for _ in range(plan.world)]
            dist.all_gather(slices, torch.from_numpy(mine))
            exchanges += 1
            allv = torch.cat(slices).numpy()[:n]
            for j in np.flatnonzero(allv != seen):
                pending += apply_moved(int(j), allv[j])
        else:
            for b in blocks:
                for q in range(int(b[0])):
                    pending += apply_moved(int(b[2 + 2 * q]), np.uint8(b[3 + 2 * q]))
        if check is not None:
            check.append(pending)
        if pending == 0:
            return cur, exchanges
        local_fixups()
