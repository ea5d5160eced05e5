"""CPU side of tests/test_gpu_zhub_guard.py: the hub graph must really have hubs whose class keeps
changing in steady-state sweeps (otherwise the GPU test of the hub guard would be vacuous), and
the small generated pangenomes of the other parity tests really have none."""
import numpy as np

from conftest import make_case
from test_gpu_zhub_guard import hub_case


def test_hub_labels_move_in_steady_state_sweeps(oracle, synth):
    n, d = 20_000, 32
    x, row_ptr, col, wgt, hubs = hub_case(synth, n, d, 200, 40, 0)
    assert int(np.diff(row_ptr)[hubs].min()) > 16
    assert np.array_equal(hubs, np.arange(n - 200, n))            # last CTAs of the dense round
    theta = oracle.default_theta(3, d)
    kw = dict(k=3, algo="ncem", beta=0.5, disp="sk_", prop="pk")
    full = oracle.Problem(x, row_ptr, col, wgt, it_max=40, **kw).fit(*theta)
    assert full.converged and full.iters >= 8
    prev, moving = None, 0
    for it in range(2, full.iters):
        lab = oracle.Problem(x, row_ptr, col, wgt, it_max=it, **kw).fit(*theta).label
        if prev is not None:
            moving += bool((lab[hubs] != prev[hubs]).any())
        prev = lab
    assert moving >= 3, moving                                    # hubs change class in >= 3 sweeps


def test_small_parity_pangenomes_have_no_hub():
    for n, d, seed in [(6000, 50, 42), (30000, 64, 17)]:
        pg = make_case(n, d, seed=seed)
        assert int(np.diff(pg.row_ptr).max()) <= 16
