"""Parity of the CUDA path (through the C ABI of libnem_b200.so) against oracle #2
(oracle/nem_oracle.c, itself pinned against the compiled reference in test_oracle_*.py).

Bars (BASELINE.json north_star): packing / CSR / levels / Hamming counts bit-exact; posteriors
and criteria within 1e-6 relative; hard partitions identical (exact ties are counted)."""
import numpy as np
import pytest

from conftest import make_case, rel_close

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def theta_default(oracle, d):
    return oracle.default_theta(3, d)


def random_theta(k, d, seed, uniform):
    rng = np.random.default_rng(seed)
    prop = rng.dirichlet(np.ones(k)).astype(np.float32)
    center = rng.choice(np.array([0.0, 0.5, 1.0], dtype=np.float32), size=(k, d))
    if uniform:
        disp = np.repeat(rng.uniform(0.02, 0.5, size=(k, 1)).astype(np.float32), d, axis=1)
    else:
        disp = rng.uniform(0.02, 0.5, size=(k, d)).astype(np.float32)
    return prop, center, disp


# ------------------------------------------------------------------ loader
@pytest.mark.parametrize("n,d", [(1, 1), (33, 31), (100, 32), (257, 129), (1000, 500), (64, 1000)])
def test_pack_and_transpose_bit_exact(engine, oracle, synth, n, d):
    rng = np.random.default_rng(n * 1000 + d)
    x = (rng.random((n, d)) < 0.4).astype(np.uint8)
    engine.load_dense(x)
    dm = engine.dims()
    assert dm["wpr"] % 4 == 0 and dm["wpr"] * 32 >= d
    got = engine.packed()
    want = oracle.pack(x, dm["wpr"])
    assert np.array_equal(got, want)
    assert np.array_equal(want, synth.pack_rows(x, dm["wpr"]))
    # host-packed path gives the same device image, also with an unpadded row length
    w0 = (d + 31) // 32
    engine.load_packed(oracle.pack(x, w0), d)
    assert np.array_equal(engine.packed(), want)
    # transposed bits: XT[d][i/32] bit i%32 == x[i][d]
    xt = engine.transposed()
    bits = np.unpackbits(xt.view(np.uint8), axis=1, bitorder="little")[:, :n]
    assert np.array_equal(bits, x.T)


def test_levels_match_oracle(engine, oracle, small_pg):
    pg = small_pg
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    lv, depth = oracle.levels(pg.row_ptr, pg.col)
    assert np.array_equal(engine.levels(), lv)
    assert engine.dims()["depth"] == depth
    assert engine.dims()["nnz"] == pg.col.shape[0]


# ------------------------------------------------------------------ density
@pytest.mark.parametrize("d", [1, 31, 64, 70, 500, 1000, 5000])
def test_density_popcount_path(engine, oracle, d):
    n = 777
    pg = make_case(n, d, seed=d, graph="none")
    engine.load_dense(pg.x)
    pb = oracle.Problem(pg.x)
    for seed, theta in enumerate([theta_default(oracle, d), random_theta(3, d, d + 1, True)]):
        got, ham, used = engine.stage_density(*theta, k=3, want_hamming=True)
        assert used, "popcount path expected for per-class-constant dispersions"
        assert np.array_equal(ham, pb.hamming(theta[1], theta[2]))
        assert rel_close(got, pb.logpf(*theta), RTOL)


@pytest.mark.parametrize("k,d", [(3, 70), (2, 33), (5, 500), (3, 1000), (9, 40), (3, 5000), (4, 2500), (8, 3000)])
def test_density_general_path(engine, oracle, k, d):
    n = 500
    pg = make_case(n, d, seed=d + k, graph="none")
    engine.load_dense(pg.x)
    pb = oracle.Problem(pg.x, k=k)
    prop, center, disp = random_theta(k, d, 3, False)
    disp[0, : d // 3] = 0.0            # forbidden cells (eps <= EPSILON => zero density)
    center[k - 1, : d // 2] = 0.3      # centres off {0, 1/2, 1}: (int)(x-mu) == 0
    got, _, used = engine.stage_density(prop, center, disp, k=k)
    assert not used
    want = pb.logpf(prop, center, disp)
    assert np.isinf(want).any()
    assert rel_close(got, want, RTOL)
    # the general kernel also reproduces the popcount-eligible case
    theta = random_theta(k, d, 5, True)
    got2, _, used2 = engine.stage_density(*theta, k=k, force_general=True)
    assert not used2
    assert rel_close(got2, pb.logpf(*theta), RTOL)


# ------------------------------------------------------------------ sweep
def _state_after_blind(pb, oracle, logpf, algo):
    t0 = np.zeros((pb.n, pb.k), dtype=np.float32)
    t1, _ = pb.sweep(logpf, 0.0, t0)
    return t1


@pytest.mark.parametrize("algo,update,impl", [
    ("ncem", "seq", "spec"), ("ncem", "seq", "level"), ("ncem", "para", "auto"),
    ("nem", "seq", "auto"), ("nem", "para", "auto")])
@pytest.mark.parametrize("graph", ["pangenome", "chain", "random"])
def test_sweep(engine, oracle, algo, update, impl, graph):
    pg = make_case(2500, 60, seed=11, graph=graph)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    pb = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, algo=algo, update=update)
    theta = theta_default(oracle, pg.d)
    logpf = pb.logpf(*theta)
    # neighbour term must matter: scale the data term down so beta*ctx competes
    logpf = logpf * 0.05
    t0 = np.zeros((pg.n, 3), dtype=np.float32)
    kw = dict(k=3, algo=algo, update=update, sweep_impl=impl)
    # blind sweep from the unlabelled state, then two coupled sweeps
    t_ref, lab_ref = pb.sweep(logpf, 0.0, t0)
    t_got, lab_got, _ = engine.stage_sweep(logpf, 0.0, t0, **kw)
    assert np.array_equal(lab_got, lab_ref)
    for beta in (0.5, 1.5):
        t_ref2, lab_ref2 = pb.sweep(logpf, beta, t_ref)
        t_got2, lab_got2, rounds = engine.stage_sweep(logpf, beta, t_ref, **kw)
        assert (lab_ref2 != lab_ref).any(), "test should move labels"
        if algo == "ncem":
            assert np.array_equal(lab_got2, lab_ref2)
            assert np.array_equal(t_got2, t_ref2)
        else:
            assert rel_close(t_got2, t_ref2, RTOL, atol=1e-30)
        t_ref, lab_ref = t_ref2, lab_ref2


def test_sweep_asymmetric_graph_and_self_loops(engine, oracle):
    """Directed .nei files and self loops (ppanggolin.py:515) keep the sequential semantics."""
    rng = np.random.default_rng(5)
    n, d = 1500, 40
    pg = make_case(n, d, seed=3, graph="random")
    keep = rng.random(pg.col.shape[0]) < 0.6           # drop 40 % of the directed entries
    rows = np.repeat(np.arange(n), np.diff(pg.row_ptr))[keep]
    col, wgt = pg.col[keep], pg.wgt[keep]
    loops = rng.choice(n, size=50, replace=False)       # add self loops
    rows = np.concatenate([rows, loops]); col = np.concatenate([col, loops])
    wgt = np.concatenate([wgt, np.full(50, 3.0, dtype=np.float32)])
    order = np.lexsort((col, rows))
    rows, col, wgt = rows[order], col[order].astype(np.int32), wgt[order].astype(np.float32)
    row_ptr = np.zeros(n + 1, dtype=np.int32); np.add.at(row_ptr, rows + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    engine.load_dense(pg.x, row_ptr, col, wgt)
    lv, depth = oracle.levels(row_ptr, col)
    assert np.array_equal(engine.levels(), lv)
    for algo, impl in [("ncem", "spec"), ("ncem", "level"), ("nem", "auto")]:
        pb = oracle.Problem(pg.x, row_ptr, col, wgt, algo=algo)
        logpf = pb.logpf(*theta_default(oracle, d)) * 0.05
        t0 = np.zeros((n, 3), dtype=np.float32)
        t1, _ = pb.sweep(logpf, 0.0, t0)
        t_ref, lab_ref = pb.sweep(logpf, 1.0, t1)
        t_got, lab_got, _ = engine.stage_sweep(logpf, 1.0, t1, k=3, algo=algo, sweep_impl=impl)
        if algo == "ncem":
            assert np.array_equal(lab_got, lab_ref)
        else:
            assert rel_close(t_got, t_ref, RTOL, atol=1e-30)


# ------------------------------------------------------------------ M-step
@pytest.mark.parametrize("algo", ["ncem", "nem"])
@pytest.mark.parametrize("disp", ["s__", "sk_", "s_d", "skd"])
@pytest.mark.parametrize("prop", ["pk", "p_"])
def test_mstep(engine, oracle, algo, disp, prop):
    pg = make_case(5000, 90, seed=21, graph="none")
    engine.load_dense(pg.x)
    pb = oracle.Problem(pg.x, algo=algo, disp=disp, prop=prop)
    rng = np.random.default_rng(1)
    if algo == "ncem":
        t = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=pg.n)]
    else:
        t = rng.dirichlet(np.ones(3) * 0.3, size=pg.n).astype(np.float32)
    theta0 = theta_default(oracle, pg.d)
    st, p_ref, c_ref, d_ref, nk_ref, s_ref = pb.mstep(t, *theta0)
    empty, p_got, c_got, d_got, nk_got, s_got = engine.stage_mstep(
        t, *theta0, k=3, algo=algo, disp=disp, prop=prop)
    assert st == 0 and empty == 0
    if algo == "ncem":
        assert np.array_equal(nk_got, nk_ref) and np.array_equal(s_got, s_ref)
        assert np.array_equal(c_got, c_ref.reshape(3, -1))
        assert np.array_equal(p_got, p_ref) and np.array_equal(d_got, d_ref.reshape(3, -1))
    else:
        assert rel_close(nk_got, nk_ref, 1e-12) and rel_close(s_got, s_ref, 1e-12)
        assert np.array_equal(c_got, c_ref.reshape(3, -1))
        assert rel_close(p_got, p_ref, RTOL) and rel_close(d_got, d_ref.reshape(3, -1), RTOL)


def test_mstep_half_centres_and_empty_class(engine, oracle):
    # a column exactly balanced inside a class gives mu = 1/2 (nem_mod.c:1468-1477)
    x = np.array([[1, 0], [0, 0], [1, 1], [0, 1], [1, 1], [1, 0]], dtype=np.uint8)
    t = np.eye(3, dtype=np.float32)[[0, 0, 1, 1, 1, 1]]
    engine.load_dense(x)
    pb = oracle.Problem(x, algo="ncem", disp="skd")
    theta0 = oracle.default_theta(3, 2)
    st, p_ref, c_ref, d_ref, *_ = pb.mstep(t, *theta0)
    empty, p_got, c_got, d_got, *_ = engine.stage_mstep(t, *theta0, k=3, algo="ncem", disp="skd")
    assert st == 1 and empty == 3            # class 3 is empty -> W_EMPTYCLASS
    assert c_ref.reshape(3, 2)[0, 0] == 0.5 and c_ref.reshape(3, 2)[1, 1] == 1.0
    assert np.array_equal(c_got, c_ref.reshape(3, 2))
    assert np.array_equal(d_got, d_ref.reshape(3, 2)) and np.array_equal(p_got, p_ref)


# ------------------------------------------------------------------ criteria
@pytest.mark.parametrize("algo", ["ncem", "nem"])
def test_criteria(engine, oracle, small_pg, algo):
    pg = small_pg
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    pb = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, algo=algo)
    logpf = pb.logpf(*theta_default(oracle, pg.d))
    t0 = np.zeros((pg.n, 3), dtype=np.float32)
    t1, _ = pb.sweep(logpf, 0.0, t0)
    t2, _ = pb.sweep(logpf, 0.5, t1)
    want = pb.criteria(logpf, t2, 0.5)
    got = engine.stage_criteria(logpf, t2, 0.5, k=3, algo=algo)
    for key in "UDLMZG":
        assert rel_close(got[key], want[key], RTOL), (key, got[key], want[key])


# ------------------------------------------------------------------ whole fit
FIT_CASES = [
    # algo, update, disp, prop, beta, impl
    ("ncem", "seq", "sk_", "pk", 0.5, "spec"),     # exactly PPanGGOLiN's call (ppanggolin.py:1814-1826)
    ("ncem", "seq", "sk_", "pk", 0.5, "level"),
    ("ncem", "seq", "skd", "pk", 0.5, "spec"),     # -fd free dispersion
    ("ncem", "seq", "s_d", "p_", 1.0, "spec"),
    ("ncem", "seq", "s__", "pk", 0.5, "spec"),
    ("ncem", "para", "sk_", "pk", 0.5, "auto"),
    ("ncem", "seq", "sk_", "pk", 0.0, "auto"),     # pure mixture
    ("nem", "seq", "sk_", "pk", 0.5, "auto"),
    ("nem", "para", "skd", "pk", 0.5, "auto"),
]


@pytest.mark.parametrize("algo,update,disp,prop,beta,impl", FIT_CASES)
def test_fit_matches_oracle(engine, oracle, algo, update, disp, prop, beta, impl):
    pg = make_case(6000, 50, seed=42)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    it_max = 100 if algo == "ncem" else 12
    kw = dict(k=3, algo=algo, update=update, disp=disp, prop=prop, beta=beta, it_max=it_max)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*oracle.default_theta(3, pg.d))
    got = engine.fit(*oracle.default_theta(3, pg.d), sweep_impl=impl, **kw)
    assert got.status == ref.status == 0
    assert got.iters == ref.iters and got.converged == ref.converged
    t = engine.posteriors()
    lab = engine.labels()
    if algo == "ncem":
        if update == "seq":
            assert got.iters < it_max, "sequential ncem should converge"
        assert np.array_equal(lab, ref.label), f"{int((lab != ref.label).sum())} labels differ"
        assert np.array_equal(t, ref.t)
        assert np.array_equal(got.center, ref.center)
        assert np.array_equal(got.disp, ref.disp) and np.array_equal(got.prop, ref.prop)
    else:
        assert rel_close(t, ref.t, RTOL, atol=1e-30)
        margin = np.sort(ref.t, axis=1)
        safe = (margin[:, -1] - margin[:, -2]) >= 1e-4
        assert np.array_equal(lab[safe], ref.label[safe])
        assert np.array_equal(got.center, ref.center)
        assert rel_close(got.disp, ref.disp, RTOL) and rel_close(got.prop, ref.prop, RTOL)
    for key in "UDLMZG":
        assert rel_close(got.crit[key], ref.crit[key], RTOL), (key, got.crit[key], ref.crit[key])
    assert got.n_ties == ref.n_ties and got.n_allnul == ref.n_allnul
    assert got.kernel_launches > 0


def test_fit_nonspatial_and_fixed_params(engine, oracle):
    pg = make_case(4000, 64, seed=9, graph="none")
    engine.load_dense(pg.x)                          # type N: beta forced to 0 (nem_exe.c:570-574)
    for fixed in (False, True):
        kw = dict(k=3, algo="ncem", beta=0.7, param_fixed=fixed)
        ref = oracle.Problem(pg.x, **kw).fit(*oracle.default_theta(3, pg.d))
        got = engine.fit(*oracle.default_theta(3, pg.d), **kw)
        assert got.iters == ref.iters and np.array_equal(engine.labels(), ref.label)
        assert np.array_equal(got.disp, ref.disp)
        for key in "UDLM":
            assert rel_close(got.crit[key], ref.crit[key], RTOL)


def test_fit_it_max_zero_and_k_other_than_3(engine, oracle):
    pg = make_case(3000, 48, seed=4)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", it_max=0)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    got = engine.fit(*theta, **kw)
    assert got.iters == ref.iters == 0 and np.array_equal(engine.labels(), ref.label)
    assert np.array_equal(got.disp, ref.disp)
    for k in (2, 5):
        rng = np.random.default_rng(k)
        # distinct proportions / dispersions per class: no MATHEMATICAL ties, which no two
        # implementations (not even the reference with itself, float dk) break alike
        prop = (np.arange(1, k + 1) / np.arange(1, k + 1).sum()).astype(np.float32)
        center = pg.x[rng.choice(pg.n, size=k, replace=False)].astype(np.float32)
        disp = np.repeat((0.11 + 0.037 * np.arange(k, dtype=np.float32))[:, None], pg.d, axis=1)
        kw = dict(k=k, algo="ncem", disp="skd", it_max=30)
        ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(prop, center, disp)
        got = engine.fit(prop, center, disp, **kw)
        assert got.status == ref.status and got.iters == ref.iters
        if ref.status == 0:
            assert np.array_equal(engine.labels(), ref.label)


@pytest.mark.parametrize("it_max", [1, 2, 3, 5, 8, 9, 10, 30])
def test_speculative_iterations_change_nothing(engine, oracle, it_max, monkeypatch):
    """The EM loop enqueues iteration i+1 before the status of iteration i is known (device halt
    flag, nem_fit.c em_core), and the dense sweep copies the label of sites whose cached margin
    exceeds what theta can have moved (nemk_margins).  Both are exact shortcuts: whatever it_max
    cuts, with either switched off, the fit must equal the oracle: same iteration count, labels,
    theta, criteria."""
    pg = make_case(20000, 50, seed=42)
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=it_max)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    out = []
    for no_spec, no_margins in (("", ""), ("1", ""), ("", "1"), ("1", "1")):
        for name, val in (("NEM_B200_NO_SPEC", no_spec), ("NEM_B200_NO_MARGINS", no_margins)):
            if val:
                monkeypatch.setenv(name, val)
            else:
                monkeypatch.delenv(name, raising=False)
        for _ in range(2):                      # twice: a halted fit must leave the handle clean
            got = engine.fit(*theta, **kw)
            out.append((got, engine.labels()))
    for got, lab in out:
        assert got.iters == ref.iters and got.converged == ref.converged
        assert np.array_equal(lab, ref.label)
        assert np.array_equal(got.center, ref.center) and np.array_equal(got.disp, ref.disp)
        assert np.array_equal(got.prop, ref.prop)
        for c in "UDL":
            assert abs(got.crit[c] - ref.crit[c]) <= 1e-6 * abs(ref.crit[c])


def test_empty_class_then_the_handle_recovers(engine, oracle):
    """An M-step that empties a class aborts the loop (nem_alg.c:1831-1838: E-step not run, no
    further iteration); the device halt flag it raises must not leak into the next fit."""
    pg = make_case(4000, 40, seed=3, graph="random")
    prop = np.array([0.4, 0.4, 0.2], dtype=np.float32)
    center = np.repeat(np.array([1.0, 0.0, 1.0], dtype=np.float32)[:, None], pg.d, axis=1)
    center[2, ::2] = 0.0                                   # a pattern no family matches exactly
    disp = np.repeat(np.array([0.1, 0.1, 1e-30], dtype=np.float32)[:, None], pg.d, axis=1)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=50)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(prop, center, disp)
    assert ref.status == 1                                 # W_EMPTYCLASS
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    got = engine.fit(prop, center, disp, **kw)
    assert got.status == ref.status and got.iters == ref.iters and got.empty_class == 3
    assert np.array_equal(engine.labels(), ref.label)
    theta = oracle.default_theta(3, pg.d)
    ok = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    again = engine.fit(*theta, **kw)
    assert again.status == 0 and again.iters == ok.iters
    assert np.array_equal(engine.labels(), ok.label)


@pytest.mark.parametrize("graph,beta,disp", [("pangenome", 0.5, "sk_"), ("random", 1.0, "sk_"),
                                             ("pangenome", 2.5, "s__"), ("random", 0.3, "sk_")])
def test_margin_cache_is_exact_on_long_fits(engine, oracle, graph, beta, disp, monkeypatch):
    """Many iterations with slowly moving theta and labels that keep flipping near the class
    borders (strong beta, noisy shell): the margin cache must never keep a label the full
    evaluation would change."""
    pg = make_case(30000, 64, seed=17, graph=graph)
    theta = oracle.default_theta(3, pg.d, low_disp=0.3)
    kw = dict(k=3, algo="ncem", update="seq", disp=disp, prop="pk", beta=beta, it_max=60)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    for off, ordered in (("", ""), ("1", ""), ("", "1")):
        # ordered: hubs add their neighbours' weights in file order even though the weights are
        # integers (the default then is an order-free parallel sum, exact for integers)
        for name, val in (("NEM_B200_NO_MARGINS", off), ("NEM_B200_ORDERED_SUMS", ordered)):
            if val:
                monkeypatch.setenv(name, val)
            else:
                monkeypatch.delenv(name, raising=False)
        engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
        got = engine.fit(*theta, **kw)
        assert got.iters == ref.iters and got.converged == ref.converged, (got.iters, ref.iters)
        assert np.array_equal(engine.labels(), ref.label), int((engine.labels() != ref.label).sum())
        assert np.array_equal(got.center, ref.center) and np.array_equal(got.disp, ref.disp)
        assert got.n_ties == ref.n_ties and got.n_allnul == ref.n_allnul


def test_label_rows_getter(engine, oracle):
    """nemb_get_labels_rows: any row range equals the slice of the full read-back; ranges outside
    the pangenome are refused; a caller-owned buffer is filled in place."""
    from pangenomenem_b200 import capi
    pg = make_case(3001, 40, seed=31)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    engine.fit(*oracle.default_theta(3, pg.d), k=3, algo="ncem", beta=0.5, it_max=20)
    full = engine.labels()
    assert full.shape == (pg.n,) and full.min() >= 0 and full.max() <= 2
    for first, count in [(0, 1), (1, 31), (1000, 2001), (3000, 1), (17, 0)]:
        assert np.array_equal(engine.labels(first, count), full[first:first + count])
    buf = np.full(64, -7, dtype=np.int32)
    got = engine.labels(100, 50, out=buf)
    assert np.array_equal(got, full[100:150]) and np.all(buf[50:] == -7)
    for first, count in [(-1, 5), (3000, 2), (0, pg.n + 1)]:
        with pytest.raises(capi.NemError):
            engine.labels(first, count)


@pytest.mark.parametrize("graph,beta", [("pangenome", 0.5), ("chain", 2.0), ("random", 1.0)])
def test_fixup_rounds_on_domino_chains(engine, oracle, graph, beta, monkeypatch):
    """Long dependency chains (a chain graph with a strong beta is one long domino line): the
    fix-up rounds must walk them to the sequential sweep's labels, with the margin cache on or
    off."""
    pg = make_case(40000, 48, seed=29, graph=graph)
    theta = oracle.default_theta(3, pg.d, low_disp=0.25)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=beta, it_max=40)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    for margins_off in ("", "1"):
        if margins_off:
            monkeypatch.setenv("NEM_B200_NO_MARGINS", "1")
        else:
            monkeypatch.delenv("NEM_B200_NO_MARGINS", raising=False)
        got = engine.fit(*theta, **kw)
        assert got.iters == ref.iters and got.converged == ref.converged
        assert np.array_equal(engine.labels(), ref.label), int((engine.labels() != ref.label).sum())
        assert np.array_equal(got.disp, ref.disp) and got.n_ties == ref.n_ties


def test_large_fit_is_exact_and_repeatable(engine, oracle):
    """Races in the speculative sweep need scale to show (hundreds of thousands of concurrent
    re-evaluations): a 400 000-family pangenome with a strong coupling, fitted six times, must
    give the oracle's labels every time."""
    pg = make_case(400000, 40, seed=33, graph="pangenome")
    theta = oracle.default_theta(3, pg.d, low_disp=0.25)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=1.0, it_max=30)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    for rep in range(6):
        got = engine.fit(*theta, **kw)
        lab = engine.labels()
        assert got.iters == ref.iters, (rep, got.iters, ref.iters)
        assert np.array_equal(lab, ref.label), (rep, int((lab != ref.label).sum()))


def test_fractional_weights_keep_the_file_order(engine, oracle):
    """Non-integer edge weights: the fp64 context sums depend on the order, so the engine must add
    them in file order like SumNeighsOfClass (nem_alg.c:2865-2875).  (This pangenome has no family
    with more than 16 neighbours; the warp-per-hub sums are in test_gpu_zhub_guard.py.)"""
    pg = make_case(12000, 48, seed=23, graph="pangenome")
    rng = np.random.default_rng(5)
    wgt = (pg.wgt * rng.uniform(0.05, 1.0, size=pg.wgt.shape)).astype(np.float32)
    src = np.repeat(np.arange(pg.n), np.diff(pg.row_ptr))          # keep the graph symmetric in value
    lo, hi = np.minimum(src, pg.col), np.maximum(src, pg.col)
    key = lo.astype(np.int64) * pg.n + hi
    order = np.argsort(key, kind="stable")
    first = np.ones(key.size, dtype=bool); first[1:] = key[order][1:] != key[order][:-1]
    canon = np.empty(key.size, dtype=np.float32)
    canon[order] = np.repeat(wgt[order][first], np.diff(np.append(np.flatnonzero(first), key.size)))
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.7, it_max=60)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, canon, **kw).fit(*theta)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, canon)
    got = engine.fit(*theta, **kw)
    assert got.iters == ref.iters
    assert np.array_equal(engine.labels(), ref.label)
    assert np.array_equal(got.disp, ref.disp)


@pytest.mark.parametrize("algo,conv,thr,it_max", [("ncem", "crit", 1e-4, 40), ("ncem", "none", 0.01, 6),
                                                 ("nem", "crit", 1e-3, 25), ("nem", "clas", 0.05, 40)])
def test_convergence_tests(engine, oracle, algo, conv, thr, it_max):
    """HasConverged (nem_alg.c:2056-2112): `crit` = relative change of criterion M, `none` = run
    it_max iterations, `clas` on fuzzy posteriors = max |t - t_old| < thr."""
    pg = make_case(5000, 40, seed=12)
    engine.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo=algo, update="seq" if algo == "ncem" else "para", conv=conv, conv_thr=thr,
              disp="sk_", prop="pk", beta=0.5, it_max=it_max)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    got = engine.fit(*theta, **kw)
    assert got.iters == ref.iters and got.converged == ref.converged
    if conv == "none":
        assert got.iters == it_max and not got.converged
    if algo == "ncem":
        assert np.array_equal(engine.labels(), ref.label)
    else:
        assert rel_close(engine.posteriors(), ref.t, RTOL, atol=1e-30)
    for key in "UDLM":
        assert rel_close(got.crit[key], ref.crit[key], RTOL), key


@pytest.mark.parametrize("n,d", [(1, 1), (2, 33), (31, 1), (33, 32), (257, 65)])
def test_tiny_and_ragged_shapes(oracle, n, d):
    """Shapes around the packing granules: one family, one genome, N and D just past 32/64-bit
    word and 256-row tile boundaries, families without neighbours."""
    from pangenomenem_b200 import capi
    rng = np.random.default_rng(n * 100 + d)
    q = np.array([0.95, 0.5, 0.05])[np.arange(n) % 3]                     # persistent / shell / cloud rows
    x = (rng.random((n, d)) < q[:, None]).astype(np.uint8)
    x[x.sum(axis=1) == 0, 0] = 1
    # a chain graph with integer weights; the last family stays isolated when n > 2
    src, dst = [], []
    for i in range(max(0, n - 2)):
        src += [i, i + 1]; dst += [i + 1, i]
    order = np.lexsort((dst, src)) if src else np.zeros(0, dtype=np.int64)
    src, dst = np.asarray(src, dtype=np.int64)[order], np.asarray(dst, dtype=np.int32)[order]
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    wgt = (1 + (np.arange(dst.size) % 3)).astype(np.float32)
    wgt = np.where(src < dst, wgt, 0).astype(np.float32)                 # symmetric values
    key = {(int(a), int(b)): float(w) for a, b, w in zip(src, dst, wgt) if a < b}
    wgt = np.asarray([key[(min(a, b), max(a, b))] for a, b in zip(src, dst)], dtype=np.float32)
    prop = np.array([0.5, 0.3, 0.2], dtype=np.float32)
    center = np.repeat(np.array([1.0, 0.5, 0.0], dtype=np.float32)[:, None], d, axis=1)
    disp = np.repeat(np.array([0.1, 0.45, 0.2], dtype=np.float32)[:, None], d, axis=1)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=20)
    ref = oracle.Problem(x, row_ptr, dst, wgt, **kw).fit(prop, center, disp)
    eng = capi.Engine(0)
    eng.load_dense(x, row_ptr, dst, wgt)
    assert np.array_equal(eng.packed()[:, :(d + 31) // 32],
                          np.packbits(np.pad(x, ((0, 0), (0, (-d) % 32))).reshape(n, -1, 32), axis=2,
                                      bitorder="little").view(np.uint32).reshape(n, -1))
    got = eng.fit(prop, center, disp, **kw)
    assert got.status == ref.status and got.iters == ref.iters
    if ref.status == 0:
        assert np.array_equal(eng.labels(), ref.label)
        assert np.array_equal(got.center, ref.center) and np.array_equal(got.disp, ref.disp)
    eng.close()
