"""GPU: beta estimation -- psgrad (EstimBeta, nem_alg.c:2120-2230), the two
heuristics (ClassifyByNemHeuBeta, nem_alg.c:731-992) and the start from a given classification
they end with (INIT_FILE, nem_alg.c:1091-1113) -- through the C ABI, against the committed runs
of the unmodified reference (tests/golden/beta_*.npz) and against the float64 oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import BETA_GOLDEN_CASES, Golden, check_against_reference, make_case, rel_close

pytestmark = pytest.mark.gpu


def engine_beta_fit(engine, oracle, g):
    th = oracle.default_theta(3, g.d)
    if g.beta_mode == "psgrad":
        nit, conv, step = g.beta_params
        return engine.fit(*th, psgrad=(int(nit), conv, step), **g.opt)
    step, bmax, ddrop, dloss, lloss = g.beta_params
    return engine.fit(*th, heuristic=dict(mode=g.beta_mode, step=step, max=bmax, ddrop=ddrop,
                                          dloss=dloss, lloss=lloss), **g.opt)


@pytest.mark.parametrize("name", BETA_GOLDEN_CASES)
def test_engine_reproduces_reference_beta_estimation(engine, oracle, name):
    g = Golden(name)
    engine.load_dense(g.x, g.row_ptr, g.col, g.wgt)
    fit = engine_beta_fit(engine, oracle, g)
    assert fit.status == 0 and fit.kernel_launches > 0
    # the reference sums the gradient and the criteria in float32
    assert abs(fit.beta - g.ref_beta) <= 1e-5 * max(1.0, abs(g.ref_beta)), (fit.beta, g.ref_beta)
    if g.beta_mode != "psgrad":
        assert np.allclose(fit.beta_tested, g.ref_beta_tested, atol=6e-3)   # "%5.2f" in its log
        assert fit.n_beta_tested == len(g.ref_beta_tested)
    if g.opt["algo"] == "nem":
        assert fit.iters == g.iters and fit.converged == g.converged
        assert np.abs(engine.posteriors() - g.cm).max() < 2e-3
    else:
        check_against_reference(g, engine.posteriors(), engine.labels(), fit.prop, fit.center,
                                fit.disp, fit.crit, fit.iters, fit.converged)


@pytest.mark.parametrize("algo", ["ncem", "nem"])
@pytest.mark.parametrize("weighted", [False, True])
def test_estim_beta_stage_matches_oracle(engine, oracle, algo, weighted):
    """One EstimBeta call on a given classification: the three site sums within 1e-9 relative of
    the oracle's index-order float64 sums, the new beta bit-equal (float update of float inputs),
    Newton-like and fixed-step variants, several gradient iterations.  Weighted graphs (hubs with
    co-presence weights ~D) are where the reference's float exp() overflows; both sides use the
    shifted soft-max there."""
    pg = make_case(5000, 60, seed=3, weighted=weighted)
    kw = dict(k=3, algo=algo, it_max=3)
    pb = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, beta=0.5, **kw)
    o = pb.fit(*oracle.default_theta(3, pg.d))
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    for beta0, nit, step in [(0.5, 1, 0.0), (0.0, 4, 0.0), (1.0, 3, 0.5), (-0.2, 2, 0.0)]:
        want_b, want = pb.estim_beta(o.t, beta0, n_iter=nit, conv_thr=0.001, step=step)
        got_b, got = engine.estim_beta(o.t, beta0, psgrad=(nit, 0.001, step), **kw)
        for key in ("crit", "grad", "dsec"):
            assert rel_close(got[key], want[key], 1e-9, atol=1e-7), (key, got[key], want[key])
        assert got_b == want_b, (beta0, nit, step, got_b, want_b)


@pytest.mark.parametrize("algo,disp,weighted", [("ncem", "sk_", True), ("ncem", "skd", True),
                                                ("nem", "sk_", True), ("ncem", "s__", False)])
def test_psgrad_fit_matches_oracle(engine, oracle, algo, disp, weighted, monkeypatch):
    pg = make_case(6000, 50, seed=5, weighted=weighted)
    kw = dict(k=3, algo=algo, beta=0.3, disp=disp, it_max=30 if algo == "ncem" else 8)
    th = oracle.default_theta(3, pg.d)
    o = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit_ex(*th, psgrad=(2, 0.001, 0.0))
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    fit = engine.fit(*th, psgrad=(2, 0.001, 0.0), **kw)
    assert fit.status == o.status == 0 and fit.iters == o.iters and fit.converged == o.converged
    assert rel_close(fit.beta, o.beta, 1e-6) and fit.beta != pytest.approx(0.3)
    if algo == "ncem":
        assert np.array_equal(engine.labels(), o.label)
        assert np.array_equal(fit.center, o.center)
        # margins of the dense sweep must be void whenever beta moved: same result without them
        monkeypatch.setenv("NEM_B200_NO_MARGINS", "1")
        fit2 = engine.fit(*th, psgrad=(2, 0.001, 0.0), **kw)
        assert np.array_equal(engine.labels(), o.label) and fit2.beta == fit.beta
    else:
        assert rel_close(engine.posteriors(), o.t, 1e-6, atol=1e-9)
    assert rel_close(fit.disp, o.disp, 1e-6) and rel_close(fit.prop, o.prop, 1e-6)
    for key in "UDLM":
        assert rel_close(fit.crit[key], o.crit[key], 1e-6), key


def test_psgrad_nonspatial_keeps_beta(engine, oracle):
    pg = make_case(2000, 40, seed=6)
    engine.load_dense(pg.x)                 # type N: no graph, beta forced to 0, EstimBeta a no-op
    fit = engine.fit(*oracle.default_theta(3, pg.d), psgrad=(3, 0.001, 0.0), beta=0.7)
    o = oracle.Problem(pg.x).fit(*oracle.default_theta(3, pg.d))
    assert fit.beta == 0.0 and fit.iters == o.iters and np.array_equal(engine.labels(), o.label)


@pytest.mark.parametrize("algo", ["ncem", "nem"])
def test_fit_from_partition_matches_oracle(engine, oracle, algo):
    """INIT_FILE: NemAlgo from a given classification (here: the latent classes, then a fit's own
    result -- a fixed point, which must converge at once)."""
    pg = make_case(5000, 64, seed=8)
    kw = dict(k=3, algo=algo, beta=0.5, it_max=40 if algo == "ncem" else 6)
    t0 = np.zeros((pg.n, 3), dtype=np.float32)
    t0[np.arange(pg.n), pg.latent % 3] = 1.0
    if algo == "nem":
        t0 = (0.8 * t0 + 0.2 / 3).astype(np.float32)
    pb = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw)
    th = oracle.default_theta(3, pg.d)
    o = pb.fit_ex(*th, t_init=t0)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    junk = [np.full_like(a, 0.25) for a in th]          # theta0 is ignored
    fit = engine.fit(*junk, t_init=t0, **kw)
    assert fit.status == o.status == 0 and fit.iters == o.iters and fit.converged == o.converged
    if algo == "ncem":
        assert np.array_equal(engine.labels(), o.label)
        again = engine.fit(*junk, t_init=engine.posteriors(), **kw)
        assert again.iters == 1 and again.converged and np.array_equal(engine.labels(), o.label)
    else:
        assert rel_close(engine.posteriors(), o.t, 1e-6, atol=1e-9)
    assert np.array_equal(fit.center, o.center) and rel_close(fit.disp, o.disp, 1e-6)
    for key in "UDLM":
        assert rel_close(fit.crit[key], o.crit[key], 1e-6), key


def test_fit_from_partition_edge_cases(engine, oracle):
    from pangenomenem_b200 import capi
    pg = make_case(1500, 33, seed=9)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    th = oracle.default_theta(3, pg.d)
    t0 = np.zeros((pg.n, 3), dtype=np.float32)
    t0[:, 0] = 1.0                                      # classes 2 and 3 have no family
    fit = engine.fit(*th, t_init=t0)
    assert fit.status == 1 and fit.iters == 0 and fit.empty_class == 3   # the last empty class, nem_mod.c:1404-1409
    o = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt).fit_ex(*th, t_init=t0)
    assert o.status == 1 and o.iters == 0
    t0[5] = (0.5, 0.5, 0.0)                             # ncem wants hardened rows
    with pytest.raises(capi.NemError):
        engine.fit(*th, t_init=t0)
    t0[5] = (0, 1, 0); t0[7] = (0, 0, 1)
    fit = engine.fit(*th, t_init=t0, it_max=0)          # it_max 0: M-step + criteria only
    o = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, it_max=0).fit_ex(*th, t_init=t0)
    assert fit.status == o.status == 0 and fit.iters == 0
    assert np.array_equal(fit.center, o.center) and rel_close(fit.crit["L"], o.crit["L"], 1e-6)
    with pytest.raises(capi.NemError):
        engine.fit(*th, t_init=t0, param_fixed=True)


@pytest.mark.parametrize("mode,algo,weighted,hz", [
    ("heu_d", "ncem", True, dict(step=0.2, max=1.6)),
    ("heu_l", "ncem", True, dict()),
    ("heu_l", "ncem", False, dict(step=0.25, max=1.0, lloss=0.002)),
    ("heu_d", "nem", False, dict(step=0.25, max=1.0)),
    ("heu_d", "ncem", False, dict(step=0.5, max=1.0, ddrop=1e-3)),
])
def test_beta_heuristics_match_oracle(engine, oracle, mode, algo, weighted, hz):
    pg = make_case(4000, 48, seed=12, weighted=weighted)
    kw = dict(k=3, algo=algo, beta=0.5, it_max=100 if algo == "ncem" else 5)
    th = oracle.default_theta(3, pg.d)
    full = dict(dict(step=0.1, max=2.0, ddrop=0.8, dloss=0.5, lloss=0.02), **hz)
    o = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit_heuristic(
        *th, mode=mode, step=full["step"], bmax=full["max"], ddrop=full["ddrop"],
        dloss=full["dloss"], lloss=full["lloss"])
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    fit = engine.fit(*th, heuristic=dict(mode=mode, **hz), **kw)
    assert fit.status == o.status == 0
    assert fit.beta == pytest.approx(o.beta, abs=1e-7)
    assert np.array_equal(fit.beta_tested, o.beta_tested)
    assert rel_close(fit.crit_tested, o.crit_tested, 1e-6)
    assert fit.iters == o.iters and fit.converged == o.converged
    if algo == "ncem":
        assert np.array_equal(engine.labels(), o.label)
    else:
        assert rel_close(engine.posteriors(), o.t, 1e-6, atol=1e-9)
    assert np.array_equal(fit.center, o.center) and rel_close(fit.disp, o.disp, 1e-6)
    for key in "UDL":
        assert rel_close(fit.crit[key], o.crit[key], 1e-6), key


def test_cli_beta_options(tmp_path, oracle):
    """`nem_exe file 3 ... -B heu_l -H ...` and `-B psgrad -G ...`: the historic syntax
    (nem_hlp.c:220-245); the .mf names the mode and carries the estimate (nem_exe.c:1714-1715)."""
    from pangenomenem_b200 import build, synth
    g = Golden("beta_heu_l")
    base = str(tmp_path / "run" / "nem_file")
    g.write_files(base)
    common = [build.CLI, base, "3", "-a", "ncem", "-b", "0.5", "-c", "clas", "1e-8", "-f", "fuzzy",
              "-i", "100", "-m", "bern", "pk", "sk_", "-s", "m", "x", "-l", "y"]
    cp = subprocess.run(common + ["-B", "heu_l", "-H", "0.1", "2.0", "0.8", "0.5", "0.02"],
                        capture_output=True, text=True, timeout=600)
    assert cp.returncode == 0, cp.stderr
    mf = open(base + ".mf").read()
    assert "Beta (heuristic mixture likelihood)" in mf
    assert float(mf.split("Beta (")[1].split("\n")[1]) == pytest.approx(g.ref_beta, abs=1e-4)
    assert "Estimated beta" in open(base + ".stderr").read()
    uf = synth.read_uf(base + ".uf", 3)
    assert np.array_equal(uf.argmax(axis=1), g.label)
    g = Golden("beta_psgrad_ncem")
    base = str(tmp_path / "run2" / "nem_file")
    g.write_files(base)
    common[1] = base
    cp = subprocess.run(common + ["-B", "psgrad", "-G", "1", "0.001", "0.0", "0"],
                        capture_output=True, text=True, timeout=600)
    assert cp.returncode == 0, cp.stderr
    mf = open(base + ".mf").read()
    assert "Beta (pseudo-likelihood gradient)" in mf
    assert float(mf.split("Beta (")[1].split("\n")[1]) == pytest.approx(g.ref_beta, abs=1e-4)
    assert np.array_equal(synth.read_uf(base + ".uf", 3).argmax(axis=1), g.label)
    # random initial beta / unknown mode are refused, not ignored
    assert subprocess.run(common + ["-B", "psgrad", "-G", "1", "0.001", "0.0", "1"],
                          capture_output=True).returncode == 2
    assert subprocess.run(common + ["-B", "nope"], capture_output=True).returncode == 2
