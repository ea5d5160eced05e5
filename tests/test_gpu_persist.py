"""The persistent EM kernel (csrc/nem_persist.cuh: one cooperative launch per fit, device-wide
barriers between the phases) against the float64 oracle AND against the launch-per-stage loop
(NEM_B200_NO_PERSIST=1), for every way it can be entered and left:

  * whole fit inside the kernel (X and X^T fit the L2: in-kernel X pass and recount),
  * NEM_B200_PK_XLIMIT=0: the kernel leaves for the TMA density pass and the X^T recount and
    re-enters (what a 1M x 5000 pangenome does),
  * it_max cutting the fit at every iteration, empty class, K != 3, update=para, no graph,
    margin cache on/off, small grids (NEM_B200_PK_GRID).
Reference control flow: nem_alg.c:1151-1169, 1746-1879, 1951-1989, 2056-2112.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import make_case, rel_close  # noqa: E402

pytestmark = pytest.mark.gpu
KNOBS = ("NEM_B200_NO_PERSIST", "NEM_B200_PK_XLIMIT", "NEM_B200_PK_GRID", "NEM_B200_NO_MARGINS")


def set_knobs(monkeypatch, **kv):
    for name in KNOBS:
        monkeypatch.delenv(name, raising=False)
    for name, val in kv.items():
        monkeypatch.setenv("NEM_B200_" + name, str(val))


def check(got, lab, ref, crit_keys="UDLM"):
    assert got.status == ref.status
    assert got.iters == ref.iters and got.converged == ref.converged, (got.iters, ref.iters)
    assert np.array_equal(lab, ref.label), f"{int((lab != ref.label).sum())} labels differ"
    assert np.array_equal(got.center, ref.center)
    assert np.array_equal(got.disp, ref.disp) and np.array_equal(got.prop, ref.prop)
    for key in crit_keys:
        assert rel_close(got.crit[key], ref.crit[key], 1e-6), (key, got.crit[key], ref.crit[key])
    assert got.n_ties == ref.n_ties and got.n_allnul == ref.n_allnul


MODES = [dict(), dict(PK_XLIMIT=0), dict(NO_PERSIST=1), dict(PK_GRID=3), dict(NO_MARGINS=1)]


@pytest.mark.parametrize("graph,beta,disp,update", [
    ("pangenome", 0.5, "sk_", "seq"), ("random", 1.0, "s__", "seq"), ("chain", 2.0, "sk_", "seq"),
    ("pangenome", 0.5, "sk_", "para"), ("none", 0.0, "sk_", "seq"), ("pangenome", 0.0, "sk_", "seq")])
def test_persistent_fit_equals_oracle_and_legacy(engine, oracle, monkeypatch, graph, beta, disp, update):
    pg = make_case(30000, 70, seed=12, graph=graph)
    theta = oracle.default_theta(3, pg.d, low_disp=0.2)
    kw = dict(k=3, algo="ncem", update=update, disp=disp, prop="pk", beta=beta, it_max=60)
    args = (pg.x,) if graph == "none" else (pg.x, pg.row_ptr, pg.col, pg.wgt)
    ref = oracle.Problem(*args, **kw).fit(*theta)
    engine.load_dense(*args)
    for mode in MODES:
        set_knobs(monkeypatch, **mode)
        for rep in range(2):            # twice: the kernel must leave its counters and lists clean
            got = engine.fit(*theta, **kw)
            check(got, engine.labels(), ref)
            if "NO_PERSIST" in mode:
                assert got.pk["launches"] == 0
            elif "PK_XLIMIT" in mode:
                # left for the first recount and for every X pass, re-entered each time
                assert got.pk["launches"] >= 3 and got.pk["barriers"] > 0, got.pk
            else:
                assert got.pk["launches"] == 1 and got.kernel_launches <= 6, (got.pk, got.kernel_launches)


@pytest.mark.parametrize("it_max", [0, 1, 2, 3, 4, 6, 9, 40])
def test_persistent_fit_cut_at_every_iteration(engine, oracle, monkeypatch, it_max):
    pg = make_case(20000, 50, seed=42)
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=it_max)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    for mode in (dict(), dict(PK_XLIMIT=0)):
        set_knobs(monkeypatch, **mode)
        got = engine.fit(*theta, **kw)
        check(got, engine.labels(), ref, crit_keys="UDL")
        assert got.pk["launches"] >= 1


def test_persistent_fit_general_masks_k5_and_empty_class(engine, oracle, monkeypatch):
    """theta0 whose centres are arbitrary data rows (X pass at the start, no class of constant
    centre), K = 5, and a start that empties a class (nem_alg.c:1831-1838)."""
    pg = make_case(12000, 96, seed=4)
    rng = np.random.default_rng(5)
    k = 5
    prop = (np.arange(1, k + 1) / np.arange(1, k + 1).sum()).astype(np.float32)
    center = pg.x[rng.choice(pg.n, size=k, replace=False)].astype(np.float32)
    disp = np.repeat((0.11 + 0.037 * np.arange(k, dtype=np.float32))[:, None], pg.d, axis=1)
    kw = dict(k=k, algo="ncem", disp="sk_", it_max=30)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(prop, center, disp)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    for mode in (dict(), dict(PK_XLIMIT=0), dict(NO_PERSIST=1)):
        set_knobs(monkeypatch, **mode)
        got = engine.fit(prop, center, disp, **kw)
        assert got.status == ref.status and got.iters == ref.iters
        if ref.status == 0:
            assert np.array_equal(engine.labels(), ref.label)
            assert np.array_equal(got.disp, ref.disp)
    # empty class
    pg = make_case(4000, 40, seed=3, graph="random")
    prop = np.array([0.4, 0.4, 0.2], dtype=np.float32)
    center = np.repeat(np.array([1.0, 0.0, 1.0], dtype=np.float32)[:, None], pg.d, axis=1)
    center[2, ::2] = 0.0
    disp = np.repeat(np.array([0.1, 0.1, 1e-30], dtype=np.float32)[:, None], pg.d, axis=1)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=50)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(prop, center, disp)
    assert ref.status == 1
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    for mode in (dict(), dict(PK_XLIMIT=0)):
        set_knobs(monkeypatch, **mode)
        got = engine.fit(prop, center, disp, **kw)
        assert got.status == ref.status and got.iters == ref.iters and got.empty_class == 3
        assert np.array_equal(engine.labels(), ref.label)
        ok = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*oracle.default_theta(3, pg.d))
        again = engine.fit(*oracle.default_theta(3, pg.d), **kw)
        assert again.status == 0 and again.iters == ok.iters
        assert np.array_equal(engine.labels(), ok.label)


def test_persistent_fit_long_and_repeatable(engine, oracle, monkeypatch):
    """Strong coupling, labels moving for many iterations, hubs: six fits, same labels each time."""
    set_knobs(monkeypatch)
    pg = make_case(300000, 40, seed=33, graph="pangenome")
    theta = oracle.default_theta(3, pg.d, low_disp=0.25)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=1.0, it_max=30)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    for rep in range(6):
        got = engine.fit(*theta, **kw)
        lab = engine.labels()
        assert got.pk["launches"] == 1
        assert got.iters == ref.iters, (rep, got.iters, ref.iters)
        assert np.array_equal(lab, ref.label), (rep, int((lab != ref.label).sum()))
        assert got.n_kept > 0
