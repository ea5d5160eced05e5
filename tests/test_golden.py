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


# ------------------------------------------------------------------ beta estimation
def _beta_fit(oracle, g):
    pb = oracle.Problem(g.x, g.row_ptr, g.col, g.wgt, **g.opt)
    th = oracle.default_theta(3, g.d)
    if g.beta_mode == "psgrad":
        nit, conv, step = g.beta_params
        return pb.fit_ex(*th, psgrad=(int(nit), conv, step))
    step, bmax, ddrop, dloss, lloss = g.beta_params
    return pb.fit_heuristic(*th, mode=g.beta_mode, step=step, bmax=bmax, ddrop=ddrop, dloss=dloss,
                            lloss=lloss)


from conftest import BETA_GOLDEN_CASES  # noqa: E402


@pytest.mark.parametrize("name", BETA_GOLDEN_CASES)
def test_oracle_reproduces_reference_beta_estimation(oracle, name):
    """psgrad (EstimBeta, nem_alg.c:2120-2230) and the two heuristics (ClassifyByNemHeuBeta,
    nem_alg.c:731-992) against the reference harness run with BetaModel set: the estimated beta
    (1e-5: the reference sums the gradient in float32), the betas it tested, and the final fit."""
    g = Golden(name)
    fit = _beta_fit(oracle, g)
    assert fit.status == 0
    assert abs(fit.beta - g.ref_beta) <= 1e-5 * max(1.0, abs(g.ref_beta)), (fit.beta, g.ref_beta)
    if g.beta_mode != "psgrad":
        assert np.allclose(fit.beta_tested, g.ref_beta_tested, atol=6e-3)   # "%5.2f" in the log
    if g.opt["algo"] == "nem":
        assert fit.iters == g.iters and fit.converged == g.converged   # it_max cuts these runs
        assert np.abs(fit.t - g.cm).max() < 2e-3
    else:
        check_against_reference(g, fit.t, fit.label, fit.prop, fit.center, fit.disp, fit.crit,
                                fit.iters, fit.converged)


def test_beta_golden_set_covers_the_modes():
    modes = {Golden(n).beta_mode for n in BETA_GOLDEN_CASES}
    assert modes == {"psgrad", "heu_d", "heu_l"}
