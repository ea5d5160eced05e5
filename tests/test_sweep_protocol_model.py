"""Host model of the speculative sequential sweep (csrc/nem_kernels.cu: k_sweep_ncem_jacobi round 0,
then fixup_site / mark_readers rounds separated by barriers) under RANDOM interleavings of the
threads' memory operations inside a round.

The reference sweep is in place and in index order (UPDATE_SEQ, nem_alg.c:2370-2392):
    cur_i = F_i(cur_j for j < i, old_j for j >= i)
The kernels reach the same labels without serialising: round 0 evaluates every site from the old
labels, every later round re-evaluates the sites one of whose lower-index inputs moved.  The
protocol the model checks, operation by operation:
  * a site clears its dirty flag BEFORE it reads its inputs;
  * a site that changes publishes the label BEFORE it claims (flag 0 -> 1) its later readers;
  * a claimed reader goes to the NEXT round's list (one evaluator per site per round).
Whatever the schedule, the fixed point must be the sequential sweep's labels.  The last test shows
that the model is sharp: "chasing" a claimed reader inside the same round (removed from the kernels
after bench.py's full-size property check caught it) does produce wrong labels under some schedules.
"""
import numpy as np
import pytest


def make_problem(n, k, rng, chord=0.3):
    nbrs = [dict() for _ in range(n)]

    def edge(a, b):
        if a != b:
            w = float(rng.integers(1, 6))
            nbrs[a][b] = w
            nbrs[b][a] = w
    for i in range(n - 1):
        if rng.random() < 0.85:
            edge(i, i + 1)
    for _ in range(int(chord * n)):
        edge(int(rng.integers(0, n)), int(rng.integers(0, n)))
    adj = [sorted(d.items()) for d in nbrs]
    lp = rng.normal(0.0, 1.0, size=(n, k))       # weak data term: long dependency chains
    old = rng.integers(0, k, size=n)
    return adj, lp, old


def site_label(i, adj, lp, beta, read):
    ctx = np.zeros(lp.shape[1])
    for j, w in adj[i]:
        ctx[read(j)] += w
    return int(np.argmax(lp[i] + beta * ctx))     # first maximum, like ComputeMAP with TIE_FIRST


def sequential(adj, lp, old, beta):
    cur = old.copy()
    for i in range(len(adj)):
        cur[i] = site_label(i, adj, lp, beta, lambda j: cur[j])   # in place: j > i still holds old_j
    return cur


def speculative(adj, lp, old, beta, rng, chase=False):
    n = len(adj)
    cur = np.array([site_label(i, adj, lp, beta, lambda j: old[j]) for i in range(n)])   # round 0
    dirty = np.zeros(n, dtype=int)
    work = []
    for i in range(n):
        if cur[i] != old[i]:
            for j, _ in adj[i]:
                if j > i and dirty[j] == 0:
                    dirty[j] = 1
                    work.append(j)
    rounds = 0
    while work:
        rounds += 1
        nxt = []

        def thread(i):
            while True:
                dirty[i] = 0                                    # atomicExch(&dirty[i], 0); fence
                yield
                ctx = np.zeros(lp.shape[1])
                for j, w in adj[i]:
                    ctx[cur[j] if j < i else old[j]] += w       # one label gather per step
                    yield
                km = int(np.argmax(lp[i] + beta * ctx))
                if km == cur[i]:
                    return
                cur[i] = km                                     # publish; fence
                yield
                follow = None
                for j, _ in adj[i]:
                    if j > i:
                        was, dirty[j] = dirty[j], 1             # atomicExch(&dirty[j], 1)
                        if was == 0:
                            if chase and follow is None:
                                follow = j                      # the racy shortcut
                            else:
                                nxt.append(j)
                        yield
                if follow is None:
                    return
                i = follow

        live = [thread(i) for i in work]
        while live:
            t = int(rng.integers(0, len(live)))
            try:
                next(live[t])
            except StopIteration:
                live[t] = live[-1]
                live.pop()
        work = nxt                                              # barrier between rounds
    return cur, rounds


@pytest.mark.parametrize("seed", range(12))
def test_any_interleaving_reaches_the_sequential_sweep(seed):
    rng = np.random.default_rng(seed)
    adj, lp, old = make_problem(int(rng.integers(20, 120)), int(rng.integers(2, 5)), rng)
    beta = float(rng.choice([0.3, 0.5, 1.0]))
    ref = sequential(adj, lp, old, beta)
    for _ in range(6):                                          # six schedules per problem
        got, rounds = speculative(adj, lp, old, beta, rng)
        assert np.array_equal(got, ref), rounds


def test_the_model_catches_the_chase_race():
    bad = 0
    for seed in range(40):
        rng = np.random.default_rng(1000 + seed)
        adj, lp, old = make_problem(60, 3, rng, chord=0.6)
        ref = sequential(adj, lp, old, 1.0)
        for _ in range(5):
            got, _ = speculative(adj, lp, old, 1.0, rng, chase=True)
            bad += not np.array_equal(got, ref)
    assert bad > 0


# ---- the margin-cached dense round looks at a hub twice (k_sweep_ncem_jacobi): its warp in the hub
# blocks and the light thread of its index.  Memory operations of the two visitors, in program
# order; every interleaving of the two sequences is enumerated.
def _hub_visitors(guard):
    """Generators over one hub's shared cells.  mem: stale, margin, lab_in, lab_out."""
    def hub_warp(mem, evaluate, thr):
        st = mem["stale"]; yield                      # noqa: E702  (one operation per step)
        m = mem["margin"]; yield                      # noqa: E702
        if not st and m > thr:
            mem["lab_out"] = mem["lab_in"]; yield     # noqa: E702  kept: copy
            return
        km, margin = evaluate()
        mem["lab_out"] = km; yield                    # noqa: E702
        mem["margin"] = margin; yield                 # noqa: E702
        mem["stale"] = 0; yield                       # noqa: E702

    def light_thread(mem, evaluate, thr):
        st = mem["stale"]; yield                      # noqa: E702
        m = mem["margin"]; yield                      # noqa: E702
        lb = mem["lab_in"]; yield                     # noqa: E702
        keep = (not st) and m > thr
        if keep and not guard:                        # JAC_HUB_GUARD: a kept hub is not copied here
            mem["lab_out"] = lb; yield                # noqa: E702
        # not kept: the site is a hub, the light paths skip it (is_heavy / hv)
    return hub_warp, light_thread


def _interleavings(na, nb):
    from itertools import combinations
    for pos in combinations(range(na + nb), na):
        s = ["b"] * (na + nb)
        for p in pos:
            s[p] = "a"
        yield s


def _run_hub_round(guard, schedule, stale0, margin0, old, new, new_margin, thr):
    mem = {"stale": stale0, "margin": margin0, "lab_in": old, "lab_out": None}
    hub_warp, light_thread = _hub_visitors(guard)
    gens = {"a": hub_warp(mem, lambda: (new, new_margin), thr),
            "b": light_thread(mem, lambda: (new, new_margin), thr)}
    done = set()
    for who in schedule + ["a"] * 8 + ["b"] * 8:                # the tail drains whoever is left
        if who in done:
            continue
        try:
            next(gens[who])
        except StopIteration:
            done.add(who)
    return mem["lab_out"]


@pytest.mark.parametrize("stale0,margin0", [(1, 5.0), (0, 0.5), (1, 0.5), (0, 5.0)])
def test_hub_guard_every_interleaving_keeps_the_hub_warps_label(stale0, margin0):
    """With the guard (the default build) the label in lab_out is the hub warp's in every schedule:
    the new label when the warp had to evaluate, the old one when the margin test kept the site."""
    thr, old, new = 1.0, 0, 2
    expect = old if (not stale0 and margin0 > thr) else new
    for sched in _interleavings(6, 5):
        assert _run_hub_round(True, sched, stale0, margin0, old, new, 9.0, thr) == expect


def test_the_model_catches_the_hub_copy_race():
    """Without the guard a light thread that looks after the hub warp has stored a large margin
    and cleared the flag copies the OLD label over the new one (the hole DESIGN.md 2.2 described)."""
    thr, old, new = 1.0, 0, 2
    bad = sum(_run_hub_round(False, s, 1, 5.0, old, new, 9.0, thr) != new
              for s in _interleavings(6, 5))
    assert bad > 0
