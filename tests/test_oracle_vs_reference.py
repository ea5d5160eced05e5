"""CPU: oracle #2 (float64 restatement) against oracle #1, the UNMODIFIED reference compiled in
place (oracle/_ref, built by `make -C oracle ref` where /root/reference is mounted; the binaries
travel to the GPU box, the sources do not).  Skipped when oracle/_ref is absent -- the committed
golden vectors (tests/test_golden.py) cover that case."""
import os

import numpy as np
import pytest

from conftest import make_case, rel_close


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle


CASES = [
    # n, d, graph, kwargs
    (20000, 50, "pangenome", dict(algo="ncem", beta=0.5, disp="sk_", prop="pk")),   # BASELINE config 1
    (4000, 120, "pangenome", dict(algo="ncem", beta=0.5, disp="skd", prop="pk")),
    (4000, 64, "random", dict(algo="ncem", beta=1.0, disp="s_d", prop="p_")),
    (3000, 48, "chain", dict(algo="ncem", beta=0.5, disp="s__", prop="pk", update="para", it_max=12)),
    (3000, 40, "pangenome", dict(algo="nem", beta=0.5, disp="sk_", prop="pk", it_max=10)),
    (3000, 40, "pangenome", dict(algo="nem", beta=0.5, disp="skd", prop="pk", update="para", it_max=6)),
]


@pytest.mark.parametrize("n,d,graph,kw", CASES)
def test_fit_matches_reference_harness(tmp_path, ref, n, d, graph, kw):
    from pangenomenem_b200 import synth
    pg = make_case(n, d, seed=42, graph=graph)
    base = str(tmp_path / "nem_file")
    synth.write_nem_files(base, pg)
    kw = dict(dict(it_max=100, update="seq"), **kw)
    r = ref.run_ref_harness(base, str(tmp_path / "out"), k=3, tie="first", **kw)
    o = ref.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, k=3, **kw).fit(*ref.default_theta(3, d))
    assert r["status"] == 0 and not r["density_zero"]
    assert o.iters == r["iters"] and o.converged == r["converged"]
    assert np.array_equal(o.center, r["center"])
    if kw["algo"] == "ncem":
        assert np.array_equal(o.label, r["cm"].argmax(axis=1))
        assert rel_close(o.disp, r["disp"], 1e-6) and rel_close(o.prop, r["prop"], 1e-6)
    else:
        # the reference accumulates the log-density in float (nem_mod.c:631,661)
        assert np.abs(o.t - r["cm"]).max() < 2e-3
        assert rel_close(o.disp, r["disp"], 1e-4) and rel_close(o.prop, r["prop"], 1e-4)
    for key in "UDL":
        assert rel_close(o.crit[key], r["crit"][key], 2e-3), key


def test_nem_cli_as_ppanggolin_calls_it(tmp_path, ref):
    """nem() through its 13 arguments (ppanggolin.py:1814-1826): parse .uf/.mf like
    run_partitioning (ppanggolin.py:1890-1972) and compare the P/S/C partition."""
    from pangenomenem_b200 import synth
    pg = make_case(6000, 50, seed=1)
    base = str(tmp_path / "nem_file")
    synth.write_nem_files(base, pg)
    rc, _, _ = ref.run_ref_cli(base, dolog=1)
    assert rc == 0
    uf = synth.read_uf(base + ".uf", 3)
    mf = synth.read_mf(base + ".mf", 3, pg.d)
    o = ref.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt).fit(*ref.default_theta(3, pg.d))
    ties = int((np.sort(uf, axis=1)[:, -1] == np.sort(uf, axis=1)[:, -2]).sum())
    assert (uf.argmax(axis=1) != o.label).sum() <= o.n_ties + ties
    assert np.array_equal(mf["mu"], o.center)
    assert rel_close(mf["eps"], o.disp, 1e-5)            # %10g prints 6 digits
    assert "NEM converged after %d iterations" % o.iters in open(base + ".stderr").read()
    names = synth.classify_psc(uf, mf)
    assert set(names) <= {"P", "S", "C"} and names.count("P") > 0


def test_reference_underflow_is_gated_not_copied(tmp_path, ref):
    """At D >~ 1000 the reference's linear-domain p_k f_k underflows (nem_alg.c:2589-2613) and it
    warns 'density = 0'; the restatement stays finite (SURVEY.md 'five things' #4)."""
    from pangenomenem_b200 import synth
    pg = make_case(400, 2000, seed=2, graph="chain")
    base = str(tmp_path / "nem_file")
    synth.write_nem_files(base, pg)
    r = ref.run_ref_harness(base, str(tmp_path / "out"), k=3, tie="first", it_max=3, conv="none")
    o = ref.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, it_max=3, conv="none").fit(*ref.default_theta(3, pg.d))
    assert r["density_zero"]
    assert o.n_allnul == 0 and np.isfinite(o.crit["L"])


def test_reference_context_overflow_is_gated_not_copied(tmp_path, ref):
    """The reference evaluates p_k f_k exp(beta ctx_k) in the linear domain (nem_alg.c:2589-2613):
    with co-presence weights of a few hundred genomes, beta*ctx passes log(DBL_MAX) = 709.8, the
    numerator is inf, the posterior inf/inf = NaN and ComputeMAP (nem_alg.c:590-645) keeps class 0
    whatever the data.  With PPanGGOLiN's class order the overflowing class IS class 0 (persistent),
    so the accident is invisible; with the persistent class last the reference mislabels those
    families and never converges.  The log-domain restatement (and the CUDA engine) stay finite:
    found by a randomised oracle-vs-reference campaign at the end of round 1."""
    from pangenomenem_b200 import synth
    pg = make_case(600, 40, seed=9, graph="chain")
    pg = synth.Pangenome(x=pg.x, row_ptr=pg.row_ptr, col=pg.col,
                         wgt=(pg.wgt * np.float32(12)).astype(np.float32), latent=pg.latent)
    kw = dict(algo="ncem", beta=1.0, disp="sk_", prop="pk", it_max=30, update="seq")

    def both(theta, m_text, sub):
        base = str(tmp_path / sub / "nem_file")
        synth.write_nem_files(base, pg, m_text=m_text)
        r = ref.run_ref_harness(base, str(tmp_path / sub / "out"), k=3, tie="first", **kw)
        o = ref.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, k=3, **kw).fit(*theta)
        assert r["status"] == 0 and not r["density_zero"]       # the reference does not notice
        ctx = np.zeros((pg.n, 3))
        for i in range(pg.n):
            for e in range(pg.row_ptr[i], pg.row_ptr[i + 1]):
                ctx[i, o.label[pg.col[e]]] += pg.wgt[e]
        over = kw["beta"] * ctx.max(axis=1) > 709.0
        return r, o, over

    p, c, s = ref.default_theta(3, pg.d)
    r, o, over = both((p, c, s), None, "first")
    assert over.sum() > 50                                        # the overflow is there ...
    assert np.array_equal(o.label, r["cm"].argmax(axis=1))       # ... and lands on class 0 by luck
    assert o.iters == r["iters"] and o.converged and r["converged"]

    p2, c2, s2 = p[::-1].copy(), c[::-1].copy(), s[::-1].copy()   # cloud, shell, persistent
    p2[2] = np.float32(np.float32(1.0) - p2[0]) - p2[1]           # as ReadParamFile rebuilds the last one
    r, o, over = both((p2, c2, s2), synth.m_line(1, p2, c2, s2), "last")
    diff = o.label != r["cm"].argmax(axis=1)
    assert o.converged and not r["converged"]
    assert diff.sum() > 50 and (diff & over).sum() > 0.7 * diff.sum()
    # the restatement's partition is the first run's with the class names permuted
    first = ref.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, k=3, **kw).fit(p, c, s)
    assert np.array_equal(2 - o.label, first.label)
