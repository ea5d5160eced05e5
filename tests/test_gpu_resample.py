"""GPU: the resample driver (include/nem_b200.h layer 4).  The subsample the device builds must be
the one PPanGGOLiN writes for an organism subset (ppanggolin.py:821-930: families without a
selected organism dropped and renumbered in order, edge weight = number of selected organisms
holding the edge, zero => dropped) -- bit-exact packing and CSR -- and the fit on it must equal the
oracle's fit of the host-built subsample.  The batch runner's votes are checked against a host loop
over the same runs (class -> P/S/C map of ppanggolin.py:1925-1957)."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def host_subsample(pg, mask):
    """numpy restatement of __write_nem_input_files for the genome subset `mask` (co-presence edges)."""
    cols = np.flatnonzero(mask)
    xs = pg.x[:, cols]
    active = xs.any(axis=1)
    idx = np.flatnonzero(active)
    new_id = np.full(pg.n, -1, dtype=np.int64)
    new_id[idx] = np.arange(idx.size)
    src = np.repeat(np.arange(pg.n), np.diff(pg.row_ptr))
    w = (xs[src].astype(np.int32) * xs[pg.col].astype(np.int32)).sum(axis=1)
    keep = active[src] & active[pg.col] & (w > 0)
    rows = new_id[src[keep]]
    row_ptr = np.zeros(idx.size + 1, dtype=np.int32)
    np.add.at(row_ptr, rows + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    return xs[idx], idx.astype(np.int32), row_ptr, new_id[pg.col[keep]].astype(np.int32), w[keep].astype(np.float32)


def random_mask(d, frac, seed):
    rng = np.random.default_rng(seed)
    m = np.zeros(d, dtype=bool)
    m[rng.choice(d, size=max(1, int(d * frac)), replace=False)] = True
    return m


@pytest.mark.parametrize("n,d,frac", [(3001, 70, 0.5), (2500, 200, 0.1), (1800, 33, 1.0)])
def test_subsample_is_bit_exact(synth, n, d, frac):
    from pangenomenem_b200 import capi
    pg = make_case(n, d, seed=11)
    mask = random_mask(d, frac, seed=n)
    xs, idx, rp, cl, wg = host_subsample(pg, mask)
    src, dst = capi.Engine(0), capi.Engine(0)
    src.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    n_eff, d_eff = src.subsample_into(dst, mask)
    assert (n_eff, d_eff) == (xs.shape[0], int(mask.sum()))
    assert np.array_equal(dst.family_index(), idx)
    got = dst.packed()
    want = synth.pack_rows(xs, got.shape[1])
    assert np.array_equal(got, want)
    grp, gcl, gwg = dst.graph()
    assert np.array_equal(grp, rp) and np.array_equal(gcl, cl) and np.array_equal(gwg, wg)
    # a second, different subsample into the same handle (buffers are reused)
    mask2 = random_mask(d, 0.3, seed=n + 1)
    xs2, idx2, rp2, cl2, wg2 = host_subsample(pg, mask2)
    src.subsample_into(dst, mask2)
    assert np.array_equal(dst.family_index(), idx2)
    assert np.array_equal(dst.packed(), synth.pack_rows(xs2, dst.packed().shape[1]))
    g2 = dst.graph()
    assert np.array_equal(g2[0], rp2) and np.array_equal(g2[1], cl2) and np.array_equal(g2[2], wg2)
    src.close(); dst.close()


def test_subsample_drops_absent_families(synth):
    from pangenomenem_b200 import capi
    pg = make_case(2000, 64, seed=5)
    mask = np.zeros(64, dtype=bool)
    mask[[3, 17]] = True                      # two genomes: most cloud families disappear
    xs, idx, rp, cl, wg = host_subsample(pg, mask)
    assert idx.size < pg.n
    src, dst = capi.Engine(0), capi.Engine(0)
    src.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    assert src.subsample_into(dst, mask) == (idx.size, 2)
    assert np.array_equal(dst.family_index(), idx)
    assert np.array_equal(dst.graph()[1], cl)
    with pytest.raises(capi.NemError):
        src.subsample_into(dst, np.zeros(64, dtype=bool))
    src.close(); dst.close()


@pytest.mark.parametrize("beta", [0.0, 0.5, 1.0])
def test_fit_on_subsample_matches_oracle(oracle, synth, beta):
    from pangenomenem_b200 import capi
    pg = make_case(6000, 120, seed=42)
    mask = random_mask(120, 0.5, seed=3)
    xs, idx, rp, cl, wg = host_subsample(pg, mask)
    theta = synth.default_theta(3, xs.shape[1])
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=beta, it_max=100)
    ref = oracle.Problem(xs, rp, cl, wg, **kw).fit(*theta)
    src, dst = capi.Engine(0), capi.Engine(0)
    src.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    src.subsample_into(dst, mask)
    got = dst.fit(*theta, **kw)
    assert got.iters == ref.iters and got.converged == ref.converged
    assert np.array_equal(dst.labels(), ref.label)
    assert np.array_equal(got.center, ref.center) and np.array_equal(got.disp, ref.disp)
    for c in "UDL":
        assert abs(got.crit[c] - ref.crit[c]) <= 1e-6 * abs(ref.crit[c])   # north-star tolerance
    src.close(); dst.close()


def psc_consistent(center, disp):
    sum_mu = (center != 0).sum(axis=1)
    sum_eps = disp.astype(np.float64).sum(axis=1)
    return int(np.argmax(sum_mu)) == 0 and int(np.argmax(sum_eps)) == 1


def test_resample_batch_votes(synth):
    from pangenomenem_b200 import capi
    pg = make_case(5000, 96, seed=9)
    runs = 14
    masks = np.stack([random_mask(96, 0.25 + 0.05 * (r % 5), seed=100 + r) for r in range(runs)])
    betas = np.linspace(0.0, 1.0, runs).astype(np.float32)
    src, dst = capi.Engine(0), capi.Engine(0)
    src.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", it_max=100)
    votes, iters, st = src.resample_batch(masks, betas, n_workers=3, **kw)
    want = np.zeros((pg.n, 4), dtype=np.int32)
    want_iters = []
    for r in range(runs):
        src.subsample_into(dst, masks[r])
        fit = dst.fit(*synth.default_theta(3, dst.d), beta=float(betas[r]), **kw)
        idx, lab = dst.family_index(), dst.labels()
        want_iters.append(fit.iters)
        if fit.status != 0 or not psc_consistent(fit.center, fit.disp):
            want[idx, 3] += 1
        else:
            np.add.at(want, (idx, lab), 1)
    assert np.array_equal(votes, want)
    assert list(iters) == want_iters
    assert st.n_runs == runs and st.n_ok + st.n_inconsistent + st.n_failed == runs
    assert votes.sum() == sum(int((pg.x[:, m].any(axis=1)).sum()) for m in masks)
    src.close(); dst.close()
