"""CPU: oracle #2 (oracle/nem_oracle.c) against the committed outputs of the unmodified
reference (tests/golden/*.npz, written by tests/golden/make_golden.py from oracle/_ref).
This is what pins the oracle; the GPU parity tests then compare the CUDA engine with the oracle
at 1e-6 and with the same golden vectors directly (tests/test_gpu_golden.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, Golden, check_against_reference


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference(oracle, name):
    g = Golden(name)
    pb = oracle.Problem(g.x, g.row_ptr, g.col, g.wgt, **g.opt)
    fit = pb.fit(*oracle.default_theta(3, g.d))
    assert fit.status == 0
    check_against_reference(g, fit.t, fit.label, fit.prop, fit.center, fit.disp, fit.crit,
                            fit.iters, fit.converged)


def test_golden_set_covers_the_path():
    algos = {Golden(n).opt["algo"] for n in GOLDEN_CASES}
    updates = {Golden(n).opt["update"] for n in GOLDEN_CASES}
    disps = {Golden(n).opt["disp"] for n in GOLDEN_CASES}
    assert algos == {"nem", "ncem"} and updates == {"seq", "para"}
    assert disps == {"s__", "sk_", "s_d", "skd"}
    assert any(not Golden(n).spatial for n in GOLDEN_CASES)


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if Golden(n).text("uf")])
def test_reference_uf_text_matches_its_own_posteriors(name):
    """The .uf text the reference's nem() wrote (TIE_RANDOM, wall-clock seed) agrees with the
    harness dump (TIE_FIRST) wherever there is no exact tie: same partition."""
    g = Golden(name)
    uf = np.array(g.text("uf").split(), dtype=np.float64).reshape(g.n, 3)
    if g.opt["algo"] == "ncem":
        assert (uf.argmax(axis=1) != g.label).sum() <= 2     # exact ties only
    else:
        assert np.abs(uf - g.cm).max() < 2e-3
