"""GPU: the row-sharded fit (DESIGN.md "Multi-GPU") on ONE device -- `world` ranks are `world`
host threads driving `world` engine handles joined by the in-process test communicator
(nemb_comm_create_local: staged device copies ordered by events; no kernel waits on another
rank's kernel).  The sharded path must reproduce the single-GPU fit and the oracle exactly:
X is sharded, the sweep stays the reference's SEQUENTIAL sweep (nem_alg.c:2370-2392) through the
cross-rank speculative fixed point, and the M-step statistics are integer sums."""
import threading

import numpy as np
import pytest

from conftest import make_case, rel_close

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def run_sharded(pg, world, theta, check_graph=False, **kw):
    from pangenomenem_b200 import capi, sharded, synth
    xp = synth.pack_rows(pg.x)
    comms = capi.local_comms(world)
    out, errs = [None] * world, []

    def work(rank):
        try:
            eng = capi.Engine(0)
            eng.set_comm(comms[rank])
            p = sharded.plan(pg.n, world, rank)
            fit = sharded.fit_sharded(eng, xp[p.rows], pg.n, pg.d, pg.row_ptr, pg.col, pg.wgt, theta,
                                      rank, world, **kw)
            if check_graph:                      # the CSR as resident on this rank
                rp, cl, wg = eng.graph()
                assert np.array_equal(rp, pg.row_ptr) and np.array_equal(cl, pg.col)
                assert np.array_equal(wg, pg.wgt)
            out[rank] = (fit, eng.labels(), eng.posteriors())
            eng.close()
        except Exception as exc:  # a dead rank would deadlock the others at the next barrier
            errs.append((rank, exc))
            import os
            os._exit(3)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for c in comms:
        capi.comm_destroy(c)
    assert not errs, errs
    return out


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("graph", ["pangenome", "random"])
def test_sharded_ncem_seq_is_exact(oracle, world, graph):
    """PPanGGOLiN's configuration (ncem, sequential sweep) on row shards == oracle == one GPU."""
    pg = make_case(6001, 50, seed=42, graph=graph)      # 6001: uneven shards
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=100)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    res = run_sharded(pg, world, theta, **kw)
    for fit, lab, t in res:
        assert fit.status == 0 and fit.iters == ref.iters and fit.converged == ref.converged
        assert np.array_equal(lab, ref.label), int((lab != ref.label).sum())
        assert np.array_equal(t, ref.t)
        assert np.array_equal(fit.center, ref.center) and np.array_equal(fit.disp, ref.disp)
        assert np.array_equal(fit.prop, ref.prop)
        for key in "UDLMZG":
            assert rel_close(fit.crit[key], ref.crit[key], 1e-6), key
    # every rank reports the same numbers, bit for bit
    for fit, lab, t in res[1:]:
        assert fit.crit == res[0][0].crit and np.array_equal(lab, res[0][1])


def test_sharded_sweep_needs_cross_rank_rounds(oracle):
    """A strong coupling (beta 2, weak data term through high dispersions) makes label changes
    propagate along the chain across the shard boundaries: the exchange loop must iterate."""
    pg = make_case(4000, 30, seed=5, graph="chain")
    prop, center, disp = oracle.default_theta(3, pg.d)
    disp = np.full_like(disp, 0.45); disp[1] = 0.5
    kw = dict(k=3, algo="ncem", update="seq", disp="skd", prop="pk", beta=2.0, it_max=15)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(prop, center, disp)
    for world in (2, 4):
        for fit, lab, _ in run_sharded(pg, world, (prop, center, disp), **kw):
            assert fit.iters == ref.iters and fit.status == ref.status
            if ref.status == 0:
                assert np.array_equal(lab, ref.label), int((lab != ref.label).sum())


@pytest.mark.parametrize("algo,disp", [("ncem", "skd"), ("nem", "sk_")])
def test_sharded_para_update(oracle, algo, disp):
    """The parallel (Jacobi) update with a halo exchange per sweep: labels (ncem) or float
    posteriors (nem) all-gathered; float64 statistics summed in rank order."""
    pg = make_case(5000, 40, seed=8)
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo=algo, update="para", disp=disp, prop="pk", beta=0.5, it_max=8)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    for fit, lab, t in run_sharded(pg, 3, theta, **kw):
        assert fit.iters == ref.iters
        if algo == "ncem":
            assert np.array_equal(lab, ref.label) and np.array_equal(fit.disp, ref.disp)
        else:
            assert rel_close(t, ref.t, 1e-6, atol=1e-30)
            assert rel_close(fit.disp, ref.disp, 1e-6) and rel_close(fit.prop, ref.prop, 1e-6)
        for key in "UDL":
            assert rel_close(fit.crit[key], ref.crit[key], 1e-6), key


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_graph_upload_in_slices(oracle, world, monkeypatch):
    """The replicated graph goes up in one slice per rank + an all-gather (nem_fit.c
    upload_replicated) once it is large; forced here for small arrays with lengths that do not
    divide by the rank count: the resident CSR and the fit must be what the plain upload gives."""
    monkeypatch.setenv("NEM_B200_SLICED_UPLOAD_MIN", "1")
    pg = make_case(5003, 40, seed=14)
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.5, it_max=50)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit(*theta)
    for fit, lab, _ in run_sharded(pg, world, theta, check_graph=True, **kw):
        assert fit.status == 0 and fit.iters == ref.iters
        assert np.array_equal(lab, ref.label)
        assert np.array_equal(fit.disp, ref.disp)


def test_sharded_psgrad(oracle):
    """BETA_PSGRAD on row shards: the pseudo-likelihood sums are per-rank partial rows gathered
    and added in rank order, so every rank derives the same beta as one GPU and the oracle."""
    pg = make_case(6001, 40, seed=10)
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", update="seq", disp="sk_", prop="pk", beta=0.3, it_max=30)
    ref = oracle.Problem(pg.x, pg.row_ptr, pg.col, pg.wgt, **kw).fit_ex(*theta, psgrad=(2, 0.001, 0.0))
    for fit, lab, _ in run_sharded(pg, 3, theta, psgrad=(2, 0.001, 0.0), **kw):
        assert fit.status == 0 and fit.iters == ref.iters and fit.converged == ref.converged
        assert rel_close(fit.beta, ref.beta, 1e-6)
        assert np.array_equal(lab, ref.label)


def test_sharded_nonspatial_and_empty_rank(oracle):
    """Type N data (no graph) and more ranks than rows-per-shard allows (an empty last rank)."""
    pg = make_case(9, 20, seed=1, graph="none")
    theta = oracle.default_theta(3, pg.d)
    kw = dict(k=3, algo="ncem", beta=0.5, it_max=20)
    ref = oracle.Problem(pg.x, **kw).fit(*theta)
    from pangenomenem_b200 import sharded
    assert sharded.plan(9, 4, 3).n_loc == 0
    for fit, lab, _ in run_sharded(pg, 4, theta, **kw):
        assert fit.status == ref.status and fit.iters == ref.iters
        if ref.status == 0:
            assert np.array_equal(lab, ref.label)


def test_sharded_rejects_single_gpu_only_modes(oracle):
    from pangenomenem_b200 import capi, synth
    pg = make_case(300, 16, seed=2)
    comms = capi.local_comms(2)
    eng = capi.Engine(0)
    eng.set_comm(comms[0])
    with pytest.raises(capi.NemError):
        eng.load_dense(pg.x, pg.row_ptr, pg.col, pg.wgt)          # single-GPU loader
    with pytest.raises(capi.NemError):                             # wrong row range for rank 0
        eng.load_shard(synth.pack_rows(pg.x)[:100], pg.n, 0, pg.d, pg.row_ptr, pg.col, pg.wgt)
    eng.close()
    for c in comms:
        capi.comm_destroy(c)
