"""GPU: init_mode = 1, the random-start entry RandNemAlgo (nem_alg.c:1574-1742) that PPanGGOLiN
uses for the shell sub-partition (ppanggolin.py:1207, 1826).  The reference draws from libc
random() seeded with the wall clock, so parity is checked piece by piece: the whole-sample
dispersion (InitPara), the shape of every random start (MakeRandomPara), every start's fit against
the oracle from the same theta0 (where no exact score tie makes the labels rounding dependent),
the best-of selection (first maximum of criterion M) and the final EstimPara on the best
partition -- for any number of worker streams."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def count_based_fit(pb, prop, center, disp, beta, it_max):
    """The oracle's EM loop (nem_alg.c:1151-1169, 1746-1879) with the densities written the way the
    engine's popcount path writes them: log p_k - (a_k H_ik + D c_k) from the INTEGER Hamming
    counts (nem_mod.c:649-674 for a dispersion that is constant over the genomes).  Families at the
    same distance from two centres of equal p and eps are then exact ties in both implementations
    (first class wins), so a random start is comparable label for label."""
    n, k, d = pb.n, pb.k, pb.d
    prop, center, disp = prop.copy(), center.reshape(k, d).copy(), disp.reshape(k, d).copy()

    def logpf():
        e = disp[:, 0].astype(np.float32)
        if not np.array_equal(disp, np.repeat(e[:, None], d, 1)):
            return pb.logpf(prop, center, disp)      # per-genome dispersions (the random start itself)
        h = pb.hamming(center, disp).astype(np.float64)
        ratio = ((np.float32(1.0) - e) / e).astype(np.float32)
        a = np.log(ratio.astype(np.float64))
        c = -np.log((np.float32(1.0) - e).astype(np.float64))
        return np.log(prop.astype(np.float64))[None, :] - (h * a[None, :] + (d * c)[None, :])

    t = np.zeros((n, k), dtype=np.float32)
    lp = logpf()
    t, _ = pb.sweep(lp, 0.0, t)
    t, lab = pb.sweep(lp, beta, t)
    iters = 0
    for it in range(1, it_max + 1):
        told = t
        st, prop, center, disp, _, _ = pb.mstep(t, prop, center, disp)
        center, disp = center.reshape(k, d), disp.reshape(k, d)
        iters = it
        if st != 0:
            break
        lp = logpf()
        t, lab = pb.sweep(lp, beta, t)
        if np.array_equal(t, told):
            break
    return iters, lab


@pytest.mark.parametrize("k,disp,graph", [(3, "sk_", "pangenome"), (4, "skd", "random")])
def test_random_starts_piece_by_piece(oracle, k, disp, graph):
    from pangenomenem_b200 import capi
    pg = make_case(5000, 40, seed=29, graph=graph)
    kw = dict(k=k, algo="ncem", update="seq", disp=disp, prop="pk", beta=0.5, it_max=60)
    eng = capi.Engine(0)
    eng.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)
    pb = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw)

    # InitPara: M-step with every family in the first class
    sam = eng.sample_dispersion(**kw)
    t_all0 = np.zeros((pg.n, k), dtype=np.float32); t_all0[:, 0] = 1.0
    half = np.full((k, pg.d), 0.5, dtype=np.float32)
    _, _, _, md, _, _ = pb.mstep(t_all0, np.full(k, 1.0 / k, dtype=np.float32), half, half)
    assert np.array_equal(sam, md.reshape(k, pg.d)[0])

    n_starts, seed = 7, 11
    rows = {tuple(r) for r in pg.x}
    thetas, own, agree, exempt = [], [], 0, 0
    for s in range(n_starts):
        prop, center, dsp = eng.random_start(k, seed, s, sam)
        assert np.allclose(prop, 1.0 / k) and np.array_equal(dsp, np.repeat((sam / np.float32(k))[None], k, 0))
        assert all(tuple(c.astype(np.uint8)) in rows for c in center)          # centres are data rows
        assert len({tuple(c) for c in center}) == k                            # ... all different
        _, c2, _ = eng.random_start(k, seed, s, sam)                           # reproducible
        assert np.array_equal(center, c2)
        thetas.append((prop, center, dsp))
        got = eng.fit(prop, center, dsp, **kw)
        lab = eng.labels()
        own.append((got, lab))
        # A random start gives every class the same proportion and dispersion, so a family at the
        # same Hamming distance from two centres is an EXACT tie of the scores.  The engine's
        # count-based density keeps it a tie (first class wins); the oracle -- like the reference --
        # adds the per-genome terms in order and its rounding breaks the tie either way.  Parity is
        # therefore required only for the starts that meet no such near-tie (north star: families
        # with a top-two margin < 1e-4 are exempt); the others are counted.
        ref = pb.fit(prop, center, dsp)
        lo = np.sort(pb.logpf(prop, center, dsp), axis=1)
        near_ties = int((lo[:, -1] - lo[:, -2] < 1e-4).sum())
        same = got.iters == ref.iters and np.array_equal(lab, ref.label)
        if not same and disp == "sk_" and got.status == 0:
            # near-ties: compare with the oracle's loop on count-based densities (ties stay ties)
            it_c, lab_c = count_based_fit(pb, prop, center, dsp, 0.5, 60)
            same = got.iters == it_c and np.array_equal(lab, lab_c)
        agree += same
        exempt += (not same) and near_ties > 0
        assert same or near_ties > 0, (s, int((lab != ref.label).sum()))
    assert len({tuple(map(tuple, th[1])) for th in thetas}) > 1                # the starts differ
    # every start agrees with the oracle, except starts that meet a near-tie (a random start gives all
    # classes the same p and per-genome eps: equidistant families are rounding-dependent in BOTH
    # implementations); those are counted, and nothing else may differ
    assert agree + exempt == n_starts, (agree, exempt)
    print(f"random starts: {agree} of {n_starts} identical to the oracle, {exempt} with near-ties")

    ok = [s for s in range(n_starts) if own[s][0].status == 0]
    assert ok
    best = ok[0]
    for s in ok[1:]:
        if own[s][0].crit["M"] > own[best][0].crit["M"]:          # strict: first maximum
            best = s
    bfit, blab = own[best]
    t_best = np.zeros((pg.n, k), dtype=np.float32)
    t_best[np.arange(pg.n), blab] = 1.0
    _, fp, fc, fd, _, _ = pb.mstep(t_best, bfit.prop, bfit.center, bfit.disp)   # final EstimPara

    for workers in (1, 3):
        got = eng.fit(*thetas[0], n_random_starts=n_starts, seed=seed, random_workers=workers, **kw)
        assert got.status == 0 and got.n_success == len(ok) and got.best_start == best + 1
        assert np.array_equal(eng.labels(), blab)
        assert got.crit["M"] == bfit.crit["M"] and got.iters == bfit.iters
        assert np.array_equal(got.center, fc.reshape(k, pg.d))
        assert np.array_equal(got.disp, fd.reshape(k, pg.d))
        assert np.array_equal(got.prop, fp)
    eng.close()
