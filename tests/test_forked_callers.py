"""nem() from forked workers and from threads -- how PPanGGOLiN really calls it: in the parent
(ppanggolin.py:1125, 1207), then from a multiprocessing.Pool (ppanggolin.py:1039) and from
ProcessPoolExecutor workers (command_line.py:262-281, 618), all created by fork().  A CUDA context
does not survive fork(), and the caller never looks at nem()'s return value (a missing .uf silently
becomes "all families undefined"), so a process that inherited an initialised CUDA state forwards
its calls to a helper process (`nem_exe --serve`, csrc/nem_api.c).

CPU part: the forwarding mechanism itself (NEM_B200_FORCE_HELPER=1; without a GPU the helper's
nem() answers 5 = no usable device, which must come back through the pipe).  GPU part: the real
parent-fit / fork / child-fit sequence, a fork Pool and two threads."""
import multiprocessing as mp
import os
import signal
import sys
import threading
import time

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import Golden  # noqa: E402

CALL = dict(nk=3, algo=b"ncem", beta=0.5, convergence=b"clas", convergence_th=1e-8, format=b"fuzzy",
            it_max=100, dolog=True, model_family=b"bern", proportion=b"pk", dispersion=b"sk_", init_mode=2)


def _files(tmp_path, tag):
    g = Golden("ppanggolin_ncem_sk")
    base = str(tmp_path / tag / "nem_file")
    g.write_files(base)
    return g, base


def test_calls_are_forwarded_to_one_helper_process(tmp_path, monkeypatch):
    from pangenomenem_b200 import capi
    has_gpu = os.path.exists("/dev/nvidiactl")
    g, base = _files(tmp_path, "a")
    monkeypatch.setenv("NEM_B200_FORCE_HELPER", "1")
    assert capi.helper_pid() == 0 or True
    rc = capi.nem(Fname=base.encode(), **CALL)
    pid = capi.helper_pid()
    assert pid > 0 and pid != os.getpid()
    os.kill(pid, 0)                                         # alive
    assert rc == (0 if has_gpu else 5)                      # the helper's answer, through the pipe
    assert os.path.exists(base + ".uf") == has_gpu
    rc2 = capi.nem(Fname=base.encode(), **CALL)
    assert rc2 == rc and capi.helper_pid() == pid           # ONE helper per calling process
    # bad arguments are answered by the helper like by the library (2, before any file is touched)
    bad = dict(CALL, algo=b"gibbs")
    assert capi.nem(Fname=base.encode(), **bad) == 2
    # a helper that died is replaced
    os.kill(pid, signal.SIGKILL)
    time.sleep(0.2)
    rc3 = capi.nem(Fname=base.encode(), **CALL)
    assert rc3 == rc and capi.helper_pid() not in (0, pid)


def _child_fit(base, q):
    from pangenomenem_b200 import capi
    rc = capi.nem(Fname=base.encode(), **CALL)
    q.put((rc, capi.helper_pid(), os.getpid()))


@pytest.mark.gpu
def test_parent_fit_then_forked_children_fit(tmp_path):
    """run_partitioning in the parent, then forked workers: every child must produce the parent's
    partition, whatever CUDA state it inherited."""
    from pangenomenem_b200 import capi, synth
    g, base = _files(tmp_path, "parent")
    assert capi.nem(Fname=base.encode(), **CALL) == 0       # CUDA is now initialised in this process
    want = synth.read_uf(base + ".uf", 3)
    assert np.array_equal(want.argmax(axis=1), g.label)
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = []
    for c in range(2):
        _, cb = _files(tmp_path, f"child{c}")
        p = ctx.Process(target=_child_fit, args=(cb, q))
        p.start()
        procs.append((p, cb))
    got = [q.get(timeout=120) for _ in procs]
    for p, _ in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rc, hpid, cpid in got:
        assert rc == 0 and hpid > 0 and hpid != cpid        # served by the child's own helper
    for _, cb in procs:
        assert np.array_equal(synth.read_uf(cb + ".uf", 3), want)
        assert open(cb + ".mf").read() == open(base + ".mf").read()


def _pool_fit(base):
    from pangenomenem_b200 import capi
    rc1 = capi.nem(Fname=base.encode(), **CALL)
    rc2 = capi.nem(Fname=base.encode(), **CALL)             # pool workers are long-lived: same helper
    return rc1, rc2, os.path.exists(base + ".uf")


@pytest.mark.gpu
def test_fork_pool_and_threads(tmp_path):
    from pangenomenem_b200 import capi, synth
    g, base = _files(tmp_path, "p0")
    assert capi.nem(Fname=base.encode(), **CALL) == 0
    bases = [_files(tmp_path, f"w{i}")[1] for i in range(4)]
    with mp.get_context("fork").Pool(2) as pool:
        out = pool.map(_pool_fit, bases)
    assert out == [(0, 0, True)] * 4
    for b in bases:
        assert np.array_equal(synth.read_uf(b + ".uf", 3).argmax(axis=1), g.label)
    # two threads of one process at once (the second gets a private engine for the call)
    res = {}
    tb = [_files(tmp_path, f"t{i}")[1] for i in range(2)]

    def run(i):
        res[i] = capi.nem(Fname=tb[i].encode(), **CALL)
    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert res == {0: 0, 1: 0}
    for b in tb:
        assert np.array_equal(synth.read_uf(b + ".uf", 3).argmax(axis=1), g.label)
