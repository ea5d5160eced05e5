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
