"""GPU: the CUDA engine, through the C ABI, against the committed outputs of the unmodified
reference (tests/golden/*.npz) -- both the in-memory API and the drop-in nem() file route that
ppanggolin.py:1814-1826 + 1886-1972 drives."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, Golden, check_against_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_engine_reproduces_reference(engine, oracle, name):
    g = Golden(name)
    engine.load_dense(g.x, g.row_ptr, g.col, g.wgt)
    fit = engine.fit(*oracle.default_theta(3, g.d), **g.opt)
    assert fit.status == 0 and fit.kernel_launches > 0
    check_against_reference(g, engine.posteriors(), engine.labels(), fit.prop, fit.center, fit.disp,
                            fit.crit, fit.iters, fit.converged)


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if Golden(n).text("uf")])
@pytest.mark.parametrize("dolog", [False, True])
def test_nem_dropin_on_the_reference_input_files(tmp_path, name, dolog):
    """nem() with the reference's 13 arguments on the very files the reference consumed; outputs
    parsed like run_partitioning (ppanggolin.py:1890-1972)."""
    from pangenomenem_b200 import capi, synth
    g = Golden(name)
    base = str(tmp_path / "run" / "nem_file")
    g.write_files(base)
    rc = capi.nem(Fname=base.encode(), nk=3, algo=g.opt["algo"].encode(), beta=g.opt["beta"],
                  convergence=b"clas", convergence_th=1e-8, format=b"fuzzy", it_max=g.opt["it_max"],
                  dolog=dolog, model_family=b"bern", proportion=g.opt["prop"].encode(),
                  dispersion=g.opt["disp"].encode(), init_mode=2)
    assert rc == 0
    uf = synth.read_uf(base + ".uf", 3)
    mf = synth.read_mf(base + ".mf", 3, g.d)
    uf_ref = np.array(g.text("uf").split(), dtype=np.float64).reshape(g.n, 3)
    mf_ref = synth.read_mf_text(g.text("mf"), 3, g.d)
    if g.opt["algo"] == "ncem":
        # the reference's nem() breaks exact ties at random (wall-clock seed): allow those rows
        assert (uf.argmax(axis=1) != uf_ref.argmax(axis=1)).sum() <= 2
        assert np.array_equal(uf.argmax(axis=1), g.label)
    else:
        assert np.abs(uf - uf_ref).max() <= 2e-3
    assert np.array_equal(mf["mu"], mf_ref["mu"])
    assert np.allclose(mf["eps"], mf_ref["eps"], rtol=2e-4) and np.allclose(mf["p"], mf_ref["p"], rtol=2e-3, atol=1e-3)
    for key in "UDL":
        assert abs(mf[key] - mf_ref[key]) <= 2e-3 * abs(mf_ref[key])
    assert synth.classify_psc(uf, mf) == synth.classify_psc(uf_ref, mf_ref) or g.opt["algo"] == "nem"
    assert os.path.exists(base + ".log") == dolog and os.path.exists(base + ".stderr") == dolog
    if dolog:
        txt = open(base + ".stderr").read()
        assert ("NEM converged after %d iterations" % g.iters in txt) == g.converged


def test_nem_hard_format_and_empty_class(tmp_path):
    from pangenomenem_b200 import capi, synth
    g = Golden("ppanggolin_ncem_sk")
    base = str(tmp_path / "run" / "nem_file")
    g.write_files(base)
    args = dict(nk=3, algo=b"ncem", beta=0.5, convergence=b"clas", convergence_th=1e-8,
                it_max=100, dolog=False, model_family=b"bern", proportion=b"pk", dispersion=b"sk_",
                init_mode=2)
    assert capi.nem(Fname=base.encode(), format=b"hard", **args) == 0
    cf = synth.read_cf(base + ".cf")
    assert np.array_equal(cf - 1, g.label)                       # 1-based MAP, nem_exe.c:1683-1700
    # a starting point that empties a class: EXIT_W_RESULT and no result files (nem_exe.c:624-631)
    os.remove(base + ".cf"); os.remove(base + ".mf")
    d = g.d
    line = "1 0.98 0.01 " + " ".join(["1"] * d + ["1"] * d + ["1"] * d) + " " + \
           " ".join(["0.4"] * d + ["1e-9"] * d + ["1e-9"] * d)
    open(base + ".m", "w").write(line)
    assert capi.nem(Fname=base.encode(), format=b"fuzzy", **args) == 1
    assert not os.path.exists(base + ".uf") and not os.path.exists(base + ".mf")


def test_cython_nem_module_drives_the_engine(tmp_path):
    """`from nem import *` + the exact call of ppanggolin.py:1814-1826."""
    import importlib
    import sys
    from pangenomenem_b200 import build_pyx, synth
    so = build_pyx.module_path() or build_pyx.build()
    sys.path.insert(0, os.path.dirname(so))
    try:
        nem = importlib.import_module("nem").nem
        g = Golden("ppanggolin_ncem_sk")
        nem_dir_path = str(tmp_path / "NEM_results")
        g.write_files(nem_dir_path + "/nem_file")
        rc = nem(Fname=nem_dir_path.encode("ascii") + b"/nem_file", nk=3, algo=b"ncem", beta=0.5,
                 convergence=b"clas", convergence_th=0.00000001, format=b"fuzzy", it_max=100,
                 dolog=True, model_family=b"bern", proportion=b"pk", dispersion=b"sk_", init_mode=2)
        assert rc == 0
        uf = synth.read_uf(nem_dir_path + "/nem_file.uf", 3)
        assert np.array_equal(uf.argmax(axis=1), g.label)
    finally:
        sys.path.pop(0)
        sys.modules.pop("nem", None)


def _log_rows(text):
    """(header lines, table header, {iteration: tokens}) of a NEM .log; the first line is dated."""
    lines = text.splitlines()
    rows, head, table_head = {}, [], None
    for ln in lines[1:]:
        tok = ln.split()
        if tok and tok[0] == "It":
            table_head = ln
        elif tok and tok[0].lstrip("-").isdigit() and len(tok) > 8:
            rows[int(tok[0])] = tok
        else:
            head.append(ln)
    return head, table_head, rows


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if Golden(n).text("log")])
def test_nem_log_matches_the_reference_table(tmp_path, name):
    """PPanGGOLiN always passes dolog=True (ppanggolin.py:1822) and points its users at
    nem_file.log: the per-iteration table -- U and M before / after the sweep, error, beta, p_k,
    mu_kd, eps_kd, n_kd (nem_alg.c:1478-1498, 1883-1946, 1995-2052, 2620-2646) -- must be the
    reference's: same lines, same header, same columns; criteria within the reference's own
    float32 rounding (its M is -inf: float zi overflows)."""
    from pangenomenem_b200 import capi
    g = Golden(name)
    base = str(tmp_path / "run" / "nem_file")
    g.write_files(base)
    rc = capi.nem(Fname=base.encode(), nk=3, algo=g.opt["algo"].encode(), beta=g.opt["beta"],
                  convergence=b"clas", convergence_th=1e-8, format=b"fuzzy", it_max=g.opt["it_max"],
                  dolog=True, model_family=b"bern", proportion=g.opt["prop"].encode(),
                  dispersion=g.opt["disp"].encode(), init_mode=2)
    assert rc == 0
    ours, ref = open(base + ".log").read(), g.text("log")
    assert len(ours.splitlines()) == len(ref.splitlines())
    h_o, th_o, r_o = _log_rows(ours)
    h_r, th_r, r_r = _log_rows(ref)
    assert h_o == h_r                      # "Criteria are multiplied by", "Initializing parameters ..." and blanks
    assert th_o == th_r                    # the column header, character for character
    assert sorted(r_o) == sorted(r_r) == list(range(0, g.iters + 1))
    k, d = 3, g.d
    for it in r_r:
        a, b = r_o[it], r_r[it]
        assert len(a) == len(b) == 1 + 6 + 1 + k + 3 * k * d, (it, len(a), len(b))
        for col in (1, 4):                 # U * mult, %5.0f
            assert abs(float(a[col]) - float(b[col])) <= max(1.5, 2e-3 * abs(float(b[col]))), (it, col, a[col], b[col])
        for col in (2, 5):                 # M * mult: finite here, -inf (float overflow) in the reference
            assert b[col] in ("-inf", "inf", "nan", "-nan") or \
                abs(float(a[col]) - float(b[col])) <= max(1.5, 2e-3 * abs(float(b[col])))
        assert a[3] == b[3] and a[6] == b[6]          # error rate: nan (no reference labels)
        assert a[7] == b[7]                           # beta
        p_o, p_r = np.array(a[8:8 + k], float), np.array(b[8:8 + k], float)
        assert np.abs(p_o - p_r).max() <= 1.1e-3
        o = 8 + k
        assert a[o:o + k * d] == b[o:o + k * d]                                   # centres
        e_o, e_r = np.array(a[o + k * d:o + 2 * k * d], float), np.array(b[o + k * d:o + 2 * k * d], float)
        assert np.abs(e_o - e_r).max() <= 1.1e-3
        n_o, n_r = np.array(a[o + 2 * k * d:], float), np.array(b[o + 2 * k * d:], float)
        assert np.abs(n_o - n_r).max() <= (0.0 if g.opt["algo"] == "ncem" else 0.2)


@pytest.mark.parametrize("q", [2, 4])
def test_nem_random_starts_through_the_files(tmp_path, q, monkeypatch):
    """partition_shell's call (ppanggolin.py:1207, 1826): nem(init_mode=1) with Q != 3 classes, no
    .m file -- RandNemAlgo's 50 random starts (nem_alg.c:1574-1742).  The file route must give the
    partition, theta and criteria of the in-memory random-start fit with the same seed."""
    from pangenomenem_b200 import capi, synth
    pg = synth.make_pangenome(3000, 60, seed=23)
    base = str(tmp_path / "shell" / "nem_file")
    synth.write_nem_files(base, pg)
    if os.path.exists(base + ".m"):
        os.remove(base + ".m")                     # init_mode=1 must not need it
    monkeypatch.setenv("NEM_B200_SEED", "17")
    rc = capi.nem(Fname=base.encode(), nk=q, algo=b"ncem", beta=0.5, convergence=b"clas",
                  convergence_th=1e-8, format=b"fuzzy", it_max=100, dolog=True, model_family=b"bern",
                  proportion=b"pk", dispersion=b"sk_", init_mode=1)
    assert rc == 0
    uf = synth.read_uf(base + ".uf", q)
    mf = synth.read_mf(base + ".mf", q, pg.d)
    assert uf.shape == (pg.n, q) and np.array_equal(uf.sum(axis=1), np.ones(pg.n))
    eng = capi.Engine(0)
    eng.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    theta = (np.full(q, 1.0 / q, np.float32), np.zeros((q, pg.d), np.float32), np.ones((q, pg.d), np.float32))
    fit = eng.fit(*theta, k=q, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=100,
                  n_random_starts=50, seed=17)
    assert fit.status == 0 and fit.n_success >= 1
    assert np.array_equal(uf.argmax(axis=1), eng.labels())
    assert np.array_equal(mf["mu"], fit.center)
    assert np.allclose(mf["eps"], fit.disp, rtol=1e-5) and np.allclose(mf["p"], fit.prop, rtol=2e-3, atol=1e-3)
    for key in "UDL":
        assert abs(mf[key] - fit.crit[key]) <= 1e-5 * abs(fit.crit[key])
    assert len(np.unique(eng.labels())) == q       # every class is used (else W_EMPTYCLASS)
    txt = open(base + ".stderr").read()
    assert "Random initial partitions (50 starts)" in txt
    eng.close()
