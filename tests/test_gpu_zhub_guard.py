"""Hubs at the HIGHEST family ids, labels that keep moving: the layout where the hub copy race of
the margin-cached dense sweep round could fire (the host model of the race is in
test_sweep_protocol_model.py).  A hub is looked at by its warp in the first blocks of
k_sweep_ncem_jacobi and by the light thread of its index in the LAST wave of CTAs; with
JAC_HUB_GUARD (default build) only the hub warp writes its label.  Every fit must give the
sequential sweep's partition (ComputePartitionNEM, nem_alg.c:2330-2405), fit after fit.

Runs last of the GPU files on purpose (file name): it is the first GPU run of the guard.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def hub_case(synth, n, d, n_hubs, hub_deg, seed):
    """A chain of families with a weak data term (runs of 50 families per latent class) plus hubs
    of degree ~hub_deg at the highest ids; on the oracle 2-24 hubs change class in each of
    iterations 3-11 of a 14-iteration fit (60 000 families, 600 hubs, beta 0.5)."""
    rng = np.random.default_rng(seed)
    runs = np.repeat(rng.integers(0, 3, size=n // 50 + 1), 50)[:n]
    p = np.array([0.7, 0.5, 0.3])[runs]
    x = (rng.random((n, d)) < p[:, None]).astype(np.uint8)
    x[x.sum(axis=1) == 0, 0] = 1
    chain = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
    hubs = np.arange(n - n_hubs, n)
    he = np.stack([np.repeat(hubs, hub_deg), rng.integers(0, n - n_hubs, size=n_hubs * hub_deg)], axis=1)
    edges = np.unique(np.sort(np.concatenate([chain, he]), axis=1), axis=0)
    edges = edges[edges[:, 0] != edges[:, 1]]
    row_ptr, col, wgt = synth.edges_to_csr(n, edges, np.ones(edges.shape[0], dtype=np.float32))
    return x, row_ptr, col, wgt, hubs


@pytest.mark.parametrize("seed", [0, 1])
def test_hubs_at_the_highest_ids_match_the_sequential_sweep(engine, oracle, synth, seed):
    n, d = 60_000, 32
    x, row_ptr, col, wgt, hubs = hub_case(synth, n, d, 600, 40, seed)
    assert int(np.diff(row_ptr)[hubs].min()) > 16            # hub warps of the dense round
    theta = oracle.default_theta(3, d)
    kw = dict(k=3, algo="ncem", beta=0.5, disp="sk_", prop="pk", it_max=40)
    ref = oracle.Problem(x, row_ptr, col, wgt, **kw).fit(*theta)
    assert ref.converged and ref.iters > 8                   # labels move for many sweeps
    engine.load_dense(x, row_ptr, col, wgt)
    for rep in range(3):                                     # schedules differ from fit to fit
        got = engine.fit(*theta, **kw)
        lab = engine.labels()
        bad = np.flatnonzero(lab != ref.label)
        assert bad.size == 0, (f"fit {rep}: {bad.size} labels differ "
                               f"({int(np.isin(bad, hubs).sum())} on hubs), kept {got.n_kept}")
        assert got.iters == ref.iters and got.converged == ref.converged


# ---- the other hub paths.  The generated pangenomes of the smaller parity cases have no family
# with more than 16 neighbours (hubs only appear from ~100 000 families on, at the lowest ids), so
# the warp-per-hub code of the criteria, of the file-order sums and of the fuzzy sweep is checked
# here on the graph above, at a size the oracle finishes in a second.
def fractional(synth, n, row_ptr, col, seed):
    """Symmetric non-integer weights on an existing CSR (the same value on i->j and j->i)."""
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    lo, hi = np.minimum(src, col).astype(np.int64), np.maximum(src, col).astype(np.int64)
    h = (lo * 1_000_003 + hi * 7919 + seed) % 9973                 # a function of the undirected pair
    return (0.05 + 0.95 * h / 9973.0).astype(np.float32)


@pytest.mark.parametrize("algo", ["ncem", "nem"])
def test_criteria_with_hubs(engine, oracle, synth, algo):
    from conftest import rel_close
    n, d = 20_000, 32
    x, row_ptr, col, wgt, hubs = hub_case(synth, n, d, 200, 40, 3)
    wgt = fractional(synth, n, row_ptr, col, 1)
    engine.load_dense(x, row_ptr, col, wgt)
    pb = oracle.Problem(x, row_ptr, col, wgt, algo=algo)
    logpf = pb.logpf(*oracle.default_theta(3, d))
    t1, _ = pb.sweep(logpf, 0.0, np.zeros((n, 3), dtype=np.float32))
    t2, _ = pb.sweep(logpf, 0.5, t1)
    want = pb.criteria(logpf, t2, 0.5)
    got = engine.stage_criteria(logpf, t2, 0.5, k=3, algo=algo)
    for key in "UDLMZG":
        assert rel_close(got[key], want[key], 1e-6), (key, got[key], want[key])


def test_hub_sums_keep_the_file_order(engine, oracle, synth):
    """Non-integer weights: a hub's fp64 context sum must be added in file order by its warp
    (SumNeighsOfClass, nem_alg.c:2865-2875), in the dense round and in the fix-up rounds."""
    from conftest import rel_close
    n, d = 20_000, 32
    x, row_ptr, col, wgt, hubs = hub_case(synth, n, d, 200, 40, 4)
    wgt = fractional(synth, n, row_ptr, col, 2)
    theta = oracle.default_theta(3, d)
    kw = dict(k=3, algo="ncem", beta=0.7, disp="sk_", prop="pk", it_max=60)
    ref = oracle.Problem(x, row_ptr, col, wgt, **kw).fit(*theta)
    engine.load_dense(x, row_ptr, col, wgt)
    got = engine.fit(*theta, **kw)
    assert got.iters == ref.iters and got.converged == ref.converged
    lab = engine.labels()
    assert np.array_equal(lab, ref.label), int((lab != ref.label).sum())
    assert np.array_equal(got.disp, ref.disp) and got.n_ties == ref.n_ties
    for key in "UDLMZG":
        assert rel_close(got.crit[key], ref.crit[key], 1e-6), (key, got.crit[key], ref.crit[key])


def test_fuzzy_nem_with_hubs(engine, oracle, synth):
    """Fuzzy `nem` (sequential update) on the hub graph: posteriors within the north-star
    tolerance, partitions identical where the posterior margin is not a near-tie."""
    from conftest import rel_close
    n, d = 20_000, 32
    x, row_ptr, col, wgt, hubs = hub_case(synth, n, d, 200, 40, 5)
    theta = oracle.default_theta(3, d)
    kw = dict(k=3, algo="nem", update="seq", beta=0.5, disp="sk_", prop="pk", it_max=8)
    ref = oracle.Problem(x, row_ptr, col, wgt, **kw).fit(*theta)
    engine.load_dense(x, row_ptr, col, wgt)
    got = engine.fit(*theta, **kw)
    assert got.status == ref.status == 0 and got.iters == ref.iters
    t = engine.posteriors()
    assert rel_close(t, ref.t, 1e-6, atol=1e-30)
    margin = np.sort(ref.t, axis=1)
    safe = (margin[:, -1] - margin[:, -2]) >= 1e-4
    assert np.array_equal(engine.labels()[safe], ref.label[safe])
