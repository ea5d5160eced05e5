"""Whole-fit parity AT THE BENCHMARKED SHAPES (BASELINE.json configs 2-4), CUDA engine vs the
float64 oracle on the same seeded pangenome: labels, iteration count, centres, dispersions and
proportions bit-exact (ncem: the statistics are integer counts), U D L M within 1e-6 relative.

  C2  100 000 x 500,  beta 0 (type N file: no graph)            -- config[1]
  C3  250 000 x 1000, beta 0.5, pangenome graph (hubs)          -- config[2]
  C4s 120 000 x 5000, beta 0.5, pangenome graph (hubs)          -- config[3]'s D, N cut to what the
      oracle fits in ~10 s; exercises the 8-CTA-cluster finalize, k_mstep_ncem through X^T,
      k_mstep_delta and the margin cache's drift bound |da_k|*D at D = 5000
plus the M-step stage at D = 1000 / 5000, the margin cache on/off at D = 5000 and the
classification of every family that differs from the unmodified reference at 20 000 x 500.
Follows nem_alg.c:1746-1879 (NemAlgo), nem_mod.c:1275-1479, 1646-1704 (M-step).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import make_case, rel_close  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def whole_fit(engine, oracle, pg, beta, spatial=True, **extra):
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=beta, it_max=100)
    kw.update(extra)
    theta = oracle.default_theta(3, pg.d)
    if spatial:
        ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
        engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    else:
        ref = oracle.Problem(pg.x, **kw).fit(*theta)
        engine.load_dense(pg.x)
    got = engine.fit(*theta, **kw)
    lab = engine.labels()
    assert got.status == ref.status == 0
    assert got.iters == ref.iters and got.converged == ref.converged, (got.iters, ref.iters)
    assert np.array_equal(lab, ref.label), f"{int((lab != ref.label).sum())} labels differ"
    assert np.array_equal(got.center, ref.center)
    assert np.array_equal(got.disp, ref.disp) and np.array_equal(got.prop, ref.prop)
    for key in "UDLM":
        assert rel_close(got.crit[key], ref.crit[key], RTOL), (key, got.crit[key], ref.crit[key])
    assert got.n_ties == ref.n_ties and got.n_allnul == ref.n_allnul
    assert got.kernel_launches > 0
    return got, ref


def test_c2_shape_whole_fit(engine, oracle):
    pg = make_case(100_000, 500, seed=42, graph="none")
    got, _ = whole_fit(engine, oracle, pg, 0.0, spatial=False)
    assert got.converged


def test_c3_shape_whole_fit(engine, oracle):
    pg = make_case(250_000, 1000, seed=42)
    assert int((np.diff(pg.row_ptr) > 16).sum()) > 1000      # the warp-per-hub paths are on
    got, _ = whole_fit(engine, oracle, pg, 0.5)
    assert got.converged


@pytest.mark.parametrize("disp", ["sk_", "skd"])
def test_c4_shape_whole_fit(engine, oracle, disp, monkeypatch):
    pg = make_case(120_000, 5000, seed=42)
    assert int((np.diff(pg.row_ptr) > 16).sum()) > 500
    monkeypatch.delenv("NEM_B200_NO_MARGINS", raising=False)
    got, ref = whole_fit(engine, oracle, pg, 0.5, disp=disp)
    assert got.converged
    if disp == "sk_":
        # margin cache off: same partition, and the cache did skip evaluations when it was on
        assert got.n_kept > 0
        monkeypatch.setenv("NEM_B200_NO_MARGINS", "1")
        off = engine.fit(*oracle.default_theta(3, pg.d), k=3, algo="ncem", update="seq", disp=disp,
                         prop="pk", beta=0.5, it_max=100)
        assert off.n_kept == 0 and off.iters == ref.iters
        assert np.array_equal(engine.labels(), ref.label)
        assert np.array_equal(off.disp, ref.disp)


@pytest.mark.parametrize("n,d", [(60_000, 1000), (40_000, 5000)])
@pytest.mark.parametrize("disp", ["sk_", "skd", "s_d"])
def test_mstep_stage_at_baseline_genome_counts(engine, oracle, n, d, disp):
    """S = X^T T through the transposed bit matrix and the finalize kernel's cluster sums at the
    genome counts of C3 / C4 (the round-1 stage test stopped at D = 90)."""
    pg = make_case(n, d, seed=5, graph="none")
    engine.load_dense(pg.x)
    pb = oracle.Problem(pg.x, algo="ncem", disp=disp, prop="pk")
    rng = np.random.default_rng(2)
    t = np.eye(3, dtype=np.float32)[rng.choice(3, size=pg.n, p=[0.45, 0.2, 0.35])]
    theta0 = oracle.default_theta(3, pg.d)
    st, p_ref, c_ref, d_ref, nk_ref, s_ref = pb.mstep(t, *theta0)
    empty, p_got, c_got, d_got, nk_got, s_got = engine.stage_mstep(
        t, *theta0, k=3, algo="ncem", disp=disp, prop="pk")
    assert st == 0 and empty == 0
    assert np.array_equal(nk_got, nk_ref) and np.array_equal(s_got, s_ref)
    assert np.array_equal(c_got, c_ref.reshape(3, -1))
    assert np.array_equal(p_got, p_ref) and np.array_equal(d_got, d_ref.reshape(3, -1))


def test_every_family_differing_from_the_reference_is_classified(engine, oracle, synth, tmp_path):
    import test_reference_differences as trd
    pg, ref = trd.reference_run(oracle, synth, tmp_path)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    got = engine.fit(*oracle.default_theta(3, pg.d), k=3, algo="ncem", update="seq", disp="sk_",
                     prop="pk", beta=trd.BETA, it_max=100)
    counts = trd.check_classified(oracle, pg, ref, engine.labels(), got.prop, got.center, got.disp)
    assert sum(counts.values()) < trd.N // 500, counts
