"""ctypes binding of libnem_b200.so (include/nem_b200.h) -- the host-side mirror used by the
tests, bench.py and the Python ``nem`` shim.  There is deliberately no fallback: if the CUDA
library is missing or no GPU is visible, calls raise."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# NEM_B200_LIB: an alternative build of the same library (tuning experiments); never a fallback
LIB_PATH = os.environ.get("NEM_B200_LIB") or os.path.join(PKG, "libnem_b200.so")

ALGO = {"nem": 0, "ncem": 1}
UPDATE = {"seq": 0, "para": 1}
CONV = {"none": 0, "clas": 1, "crit": 2}
PROP = {"p_": 0, "pk": 1}
DISP = {"s__": 0, "sk_": 1, "s_d": 2, "skd": 3}
SWEEP = {"auto": 0, "level": 1, "spec": 2}
STATUS = {0: "OK", 1: "W_EMPTYCLASS", 2: "E_ARG", 3: "E_FILE", 4: "E_MEMORY", 5: "E_CUDA", 6: "E_BUG"}


class NemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nem_b200 error {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("k", C.c_int32), ("algo", C.c_int32), ("update", C.c_int32), ("conv", C.c_int32),
                ("prop", C.c_int32), ("disp", C.c_int32), ("it_max", C.c_int32),
                ("param_fixed", C.c_int32), ("dolog", C.c_int32), ("sweep_impl", C.c_int32),
                ("profile", C.c_int32), ("beta", C.c_float), ("conv_thr", C.c_float),
                ("beta_mode", C.c_int32), ("grad_n_iter", C.c_int32), ("grad_conv", C.c_float),
                ("grad_step", C.c_float), ("reserved", C.c_int32 * 4)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("converged", C.c_int32),
                ("empty_class", C.c_int32),
                ("U", C.c_double), ("D", C.c_double), ("L", C.c_double), ("M", C.c_double),
                ("Z", C.c_double), ("G", C.c_double),
                ("n_allnul", C.c_int64), ("n_ties", C.c_int64), ("fixup_rounds", C.c_int64),
                ("kernel_launches", C.c_int64), ("fit_ms", C.c_float),
                ("ms_density", C.c_float), ("ms_sweep", C.c_float), ("ms_mstep", C.c_float),
                ("ms_criteria", C.c_float),
                ("n_density", C.c_int32), ("n_sweep", C.c_int32), ("n_mstep", C.c_int32),
                ("n_criteria", C.c_int32), ("best_start", C.c_int32), ("n_success", C.c_int32),
                ("ms_density_cached", C.c_float), ("ms_mstep_delta", C.c_float),
                ("n_density_cached", C.c_int32), ("n_mstep_delta", C.c_int32),
                ("exchanges", C.c_int64), ("n_kept", C.c_int64),
                ("beta", C.c_float), ("n_beta_tested", C.c_int32),
                ("pk_launches", C.c_int32), ("pk_barriers", C.c_int32),
                ("pk_x_passes", C.c_int32), ("pk_recounts", C.c_int32)]


class BatchStats(C.Structure):
    _fields_ = [("n_runs", C.c_int32), ("n_ok", C.c_int32), ("n_inconsistent", C.c_int32),
                ("n_failed", C.c_int32), ("family_iterations", C.c_int64),
                ("kernel_launches", C.c_int64), ("fit_ms_sum", C.c_double)]


def pack_mask(mask, d: int) -> np.ndarray:
    """bool[D] -> uint32[ceil(D/32)], genome j = bit j%32 of word j/32 (the layout of X's rows)."""
    m = np.asarray(mask)
    wm = (d + 31) // 32
    if m.dtype == np.uint32 and m.shape == (wm,):
        return np.ascontiguousarray(m)
    b = np.zeros(wm * 32, dtype=np.uint8)
    b[:d] = m.astype(bool)[:d]
    return np.packbits(b.reshape(wm, 32), axis=1, bitorder="little").view(np.uint32).reshape(wm).copy()


class Extra(C.Structure):
    _fields_ = [("update", C.c_int32), ("sweep_impl", C.c_int32), ("device", C.c_int32),
                ("n_random_inits", C.c_int32), ("seed", C.c_int64),
                ("beta_mode", C.c_int32), ("grad_n_iter", C.c_int32), ("grad_conv", C.c_float),
                ("grad_step", C.c_float), ("heu_step", C.c_float), ("heu_max", C.c_float),
                ("heu_ddrop", C.c_float), ("heu_dloss", C.c_float), ("heu_lloss", C.c_float),
                ("reserved", C.c_int32 * 7)]


class BetaHeuristic(C.Structure):
    _fields_ = [("step", C.c_float), ("max", C.c_float), ("ddrop", C.c_float),
                ("dloss", C.c_float), ("lloss", C.c_float)]


class Comm(C.Structure):
    """nemb_comm: the all-gather vtable of a row-sharded fit (include/nem_b200.h)."""
    _fields_ = [("ctx", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32),
                ("allgather", C.c_void_p), ("destroy", C.c_void_p)]


class HostProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("d", C.c_int32), ("words_per_row", C.c_int32),
                ("spatial", C.c_int32), ("nnz", C.c_int32), ("max_neigh", C.c_int32),
                ("m_flag", C.c_int32),
                ("x_packed", C.POINTER(C.c_uint32)), ("row_ptr", C.POINTER(C.c_int32)),
                ("col", C.POINTER(C.c_int32)), ("wgt", C.POINTER(C.c_float)),
                ("prop", C.POINTER(C.c_float)), ("center", C.POINTER(C.c_float)),
                ("disp", C.POINTER(C.c_float))]


_lib = None


def load_library():
    """Load the in-tree CUDA library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m pangenomenem_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.nemb_last_error.restype = C.c_char_p
        lib.nemb_version.restype = C.c_char_p
        lib.nemb_free_host_problem.restype = None
        lib.nemb_comm_destroy.restype = None
        lib.nemb_comm_destroy.argtypes = [C.c_void_p]
        lib.nemb_shard_range.restype = None
        lib.nemb_set_comm.argtypes = [C.c_void_p, C.c_void_p]
        sig = [C.c_char_p, C.c_int, C.c_char_p, C.c_float, C.c_char_p, C.c_float, C.c_char_p,
               C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        if hasattr(lib, "nem"):
            lib.nem.argtypes = sig
            lib.nem_b200_ex.argtypes = sig + [C.c_void_p]
        lib.nemb_write_mf.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def make_options(k=3, algo="ncem", update="seq", conv="clas", conv_thr=1e-8, prop="pk",
                 disp="sk_", it_max=100, beta=0.5, param_fixed=False, dolog=False,
                 sweep_impl="auto", profile=False, psgrad=None) -> Options:
    """psgrad = (n_iter, conv, step): beta re-estimated after every M-step (BETA_PSGRAD)."""
    o = Options(k, ALGO[algo], UPDATE[update], CONV[conv], PROP[prop], DISP[disp], int(it_max),
                int(param_fixed), int(dolog), SWEEP[sweep_impl], int(profile), float(beta),
                float(conv_thr))
    if psgrad is not None:
        o.beta_mode, o.grad_n_iter = 1, int(psgrad[0])
        o.grad_conv, o.grad_step = float(psgrad[1]), float(psgrad[2])
    return o


@dataclass
class Fit:
    status: int
    iters: int
    converged: bool
    prop: np.ndarray
    center: np.ndarray
    disp: np.ndarray
    crit: dict
    n_allnul: int
    n_ties: int
    fixup_rounds: int
    kernel_launches: int
    fit_ms: float
    stage_ms: dict = field(default_factory=dict)
    stage_launches: dict = field(default_factory=dict)
    empty_class: int = 0
    exchanges: int = 0
    n_kept: int = 0
    best_start: int = 0
    n_success: int = 0
    beta: float = 0.0
    n_beta_tested: int = 0
    beta_tested: np.ndarray | None = None
    crit_tested: np.ndarray | None = None
    pk: dict = field(default_factory=dict)   # persistent EM kernel: launches, barriers, x_passes, recounts


class Engine:
    """One CUDA context + one resident pangenome (nemb_handle)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.nemb_create(C.byref(self.h), int(device))
        if rc != 0:
            raise NemError(rc, "nemb_create failed (no CUDA device?)")
        self.n = self.d = 0

    def close(self):
        if self.h:
            self.lib.nemb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, ok=(0,)):
        if rc not in ok:
            raise NemError(rc, self.lib.nemb_last_error(self.h).decode())
        return rc

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.nemb_set_stream(self.h, C.c_void_p(cuda_stream)))

    # ---- loaders
    def _graph(self, row_ptr, col, wgt):
        if row_ptr is None:
            return None, None, None
        return (np.ascontiguousarray(row_ptr, dtype=np.int32),
                np.ascontiguousarray(col, dtype=np.int32),
                np.ascontiguousarray(wgt, dtype=np.float32))

    def load_dense(self, x, row_ptr=None, col=None, wgt=None):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        rp, cl, wg = self._graph(row_ptr, col, wgt)
        self.n, self.d = x.shape
        self._check(self.lib.nemb_load_dense_u8(self.h, self.n, self.d, _p(x), _p(rp), _p(cl), _p(wg)))

    def load_packed(self, xp, d, row_ptr=None, col=None, wgt=None):
        xp = np.ascontiguousarray(xp, dtype=np.uint32)
        rp, cl, wg = self._graph(row_ptr, col, wgt)
        self.n, self.d = xp.shape[0], int(d)
        self._check(self.lib.nemb_load_packed(self.h, self.n, self.d, xp.shape[1], _p(xp), _p(rp),
                                              _p(cl), _p(wg)))

    def load_packed_device(self, dev_ptr: int, n: int, d: int, wpr: int, row_ptr=None, col=None,
                           wgt=None):
        rp, cl, wg = self._graph(row_ptr, col, wgt)
        self.n, self.d = int(n), int(d)
        self._check(self.lib.nemb_load_packed_device(self.h, self.n, self.d, int(wpr),
                                                     C.c_void_p(dev_ptr), _p(rp), _p(cl), _p(wg)))

    # ---- row shards (one Engine per rank)
    def set_comm(self, comm_ptr):
        """comm_ptr: a nemb_comm* (int / c_void_p) from nccl_comm() or local_comms(); None detaches."""
        self._check(self.lib.nemb_set_comm(self.h, C.c_void_p(comm_ptr) if comm_ptr else None))

    def load_shard(self, xp_local, n_glob, row0, d, row_ptr=None, col=None, wgt=None):
        xp = np.ascontiguousarray(xp_local, dtype=np.uint32)
        rp, cl, wg = self._graph(row_ptr, col, wgt)
        self.n, self.d = int(n_glob), int(d)
        self._check(self.lib.nemb_load_shard(self.h, int(n_glob), int(row0), xp.shape[0], self.d,
                                             xp.shape[1], _p(xp), _p(rp), _p(cl), _p(wg)))

    def load_shard_device(self, dev_ptr, n_glob, row0, n_loc, d, wpr, row_ptr=None, col=None, wgt=None):
        rp, cl, wg = self._graph(row_ptr, col, wgt)
        self.n, self.d = int(n_glob), int(d)
        self._check(self.lib.nemb_load_shard_device(self.h, int(n_glob), int(row0), int(n_loc),
                                                    self.d, int(wpr), C.c_void_p(dev_ptr), _p(rp),
                                                    _p(cl), _p(wg)))

    def dims(self):
        v = [C.c_int() for _ in range(6)]
        self._check(self.lib.nemb_get_dims(self.h, *[C.byref(a) for a in v]))
        return dict(zip(("n", "d", "wpr", "nwt", "depth", "nnz"), [a.value for a in v]))

    def packed(self):
        dm = self.dims()
        out = np.zeros((dm["n"], dm["wpr"]), dtype=np.uint32)
        self._check(self.lib.nemb_get_packed(self.h, _p(out)))
        return out

    def graph(self):
        dm = self.dims()
        rp = np.zeros(self.n + 1, dtype=np.int32)      # the GLOBAL graph, also on a row shard
        cl = np.zeros(dm["nnz"], dtype=np.int32)
        wg = np.zeros(dm["nnz"], dtype=np.float32)
        self._check(self.lib.nemb_get_graph(self.h, _p(rp), _p(cl), _p(wg)))
        return rp, cl, wg

    def transposed(self):
        dm = self.dims()
        out = np.zeros((dm["d"], dm["nwt"]), dtype=np.uint32)
        self._check(self.lib.nemb_get_transposed(self.h, _p(out)))
        return out

    def levels(self):
        out = np.zeros(self.n, dtype=np.int32)
        self._check(self.lib.nemb_get_levels(self.h, _p(out)))
        return out

    # ---- fit
    def sample_dispersion(self, **kw):
        o = make_options(**kw)
        out = np.zeros(self.d, dtype=np.float32)
        self._check(self.lib.nemb_sample_dispersion(self.h, C.byref(o), _p(out)))
        self.k = o.k
        return out

    def random_start(self, k, seed, start, disp_sample):
        prop = np.zeros(k, dtype=np.float32)
        center = np.zeros((k, self.d), dtype=np.float32)
        disp = np.zeros((k, self.d), dtype=np.float32)
        self._check(self.lib.nemb_random_start(self.h, int(k), C.c_int64(seed), int(start),
                                               _p(_f32(disp_sample)), _p(prop), _p(center), _p(disp)))
        return prop, center, disp

    def fit(self, prop0, center0, disp0, n_random_starts=0, seed=42, random_workers=0,
            t_init=None, heuristic=None, **kw) -> Fit:
        """theta0 = (prop0[K], center0[K,D], disp0[K,D]); kw = make_options() fields.
        t_init[N,K]: start from this classification (nemb_fit_from_partition, theta0 ignored).
        heuristic = dict(mode="heu_d"|"heu_l", step=, max=, ddrop=, dloss=, lloss=): estimate beta
        with nemb_fit_beta_heuristic."""
        o = make_options(**kw)
        prop, center, disp = _f32(prop0).copy(), _f32(center0).copy(), _f32(disp0).copy()
        r = Result()
        bt = ct = None
        if heuristic is not None:
            hz = dict(heuristic)
            mode = {"heu_d": 2, "heu_l": 3}[hz.pop("mode", "heu_d")]
            hp = BetaHeuristic(float(hz.get("step", 0)), float(hz.get("max", 0)),
                               float(hz.get("ddrop", 0)), float(hz.get("dloss", 0)),
                               float(hz.get("lloss", 0)))
            cap = int((hp.max or 2.0) / (hp.step or 0.1)) + 3
            bt = np.zeros(cap, dtype=np.float32)
            ct = np.zeros(cap, dtype=np.float32)
            rc = self.lib.nemb_fit_beta_heuristic(self.h, C.byref(o), mode, C.byref(hp), _p(prop),
                                                  _p(center), _p(disp), C.byref(r), _p(bt), _p(ct),
                                                  cap)
        elif t_init is not None:
            t0 = _f32(t_init)
            assert t0.shape == (self.n, o.k)
            rc = self.lib.nemb_fit_from_partition(self.h, C.byref(o), _p(t0), _p(prop), _p(center),
                                                  _p(disp), C.byref(r))
        elif n_random_starts and random_workers:
            rc = self.lib.nemb_fit_random_workers(self.h, C.byref(o), int(n_random_starts), C.c_int64(seed),
                                                  int(random_workers), _p(prop), _p(center), _p(disp),
                                                  C.byref(r))
        elif n_random_starts:
            rc = self.lib.nemb_fit_random(self.h, C.byref(o), int(n_random_starts), C.c_int64(seed),
                                          _p(prop), _p(center), _p(disp), C.byref(r))
        else:
            rc = self.lib.nemb_fit(self.h, C.byref(o), _p(prop), _p(center), _p(disp), C.byref(r))
        self._check(rc, ok=(0, 1))
        self.k = o.k
        return Fit(r.status, r.iters, bool(r.converged), prop, center.reshape(o.k, self.d),
                   disp.reshape(o.k, self.d), dict(U=r.U, D=r.D, L=r.L, M=r.M, Z=r.Z, G=r.G),
                   r.n_allnul, r.n_ties, r.fixup_rounds, r.kernel_launches, r.fit_ms,
                   dict(density=r.ms_density, sweep=r.ms_sweep, mstep=r.ms_mstep,
                        criteria=r.ms_criteria, density_cached=r.ms_density_cached,
                        mstep_delta=r.ms_mstep_delta),
                   dict(density=r.n_density, sweep=r.n_sweep, mstep=r.n_mstep,
                        criteria=r.n_criteria, density_cached=r.n_density_cached,
                        mstep_delta=r.n_mstep_delta), r.empty_class, r.exchanges, r.n_kept, r.best_start,
                   r.n_success, r.beta, r.n_beta_tested,
                   None if bt is None else bt[:min(r.n_beta_tested, bt.shape[0])].copy(),
                   None if ct is None else ct[:min(r.n_beta_tested, ct.shape[0])].copy(),
                   dict(launches=r.pk_launches, barriers=r.pk_barriers, x_passes=r.pk_x_passes,
                        recounts=r.pk_recounts))

    def estim_beta(self, t, beta, **kw):
        """EstimBeta on the classification t[N,K] -> (new beta, dict(crit, grad, dsec))."""
        o = make_options(**kw)
        t = _f32(t)
        b = C.c_float(float(beta))
        out = np.zeros(3, dtype=np.float64)
        self._check(self.lib.nemb_stage_estim_beta(self.h, C.byref(o), _p(t), C.byref(b), _p(out)))
        return float(b.value), dict(zip(("crit", "grad", "dsec"), out))

    def posteriors(self, k=None):
        k = k or self.k
        out = np.zeros((self.n, k), dtype=np.float32)
        self._check(self.lib.nemb_get_posteriors(self.h, _p(out)))
        return out

    PK_PHASES = ("init", "scan", "delta", "recount", "finalize", "xpass", "margin_test", "eval_list",
                 "eval_dense", "fixup")

    def persist_profile(self) -> dict:
        """Per-phase microseconds of the persistent EM kernel during the last fit (CTA 0's clock,
        barrier waits included) + the number of fix-up rounds."""
        buf = (C.c_ulonglong * 12)()
        self._check(self.lib.nemb_get_persist_profile(self.h, buf))
        d = {name: buf[i] / 1e3 for i, name in enumerate(self.PK_PHASES)}
        d["fixup_rounds"] = int(buf[10])
        return d

    def persist_trace(self) -> np.ndarray:
        """[12, 8] per-iteration trace of the last persistent launch: us of scan, delta/recount,
        finalize, margin test, evaluation, fix-up; active sites (-1 dense); fix-up rounds."""
        buf = (C.c_longlong * 96)()
        self._check(self.lib.nemb_get_persist_trace(self.h, buf))
        t = np.array(buf[:], dtype=np.float64).reshape(12, 8)
        t[:, :6] /= 1e3
        return t

    def labels(self, first: int = 0, count: int | None = None, out: np.ndarray | None = None):
        """MAP labels of families [first, first + count) (default: all); `out` (int32, C order)
        is reused when given -- a long-lived caller avoids a fresh 4N-byte array per call."""
        count = self.n - first if count is None else count
        if out is None:
            out = np.empty(count, dtype=np.int32)
        assert out.dtype == np.int32 and out.flags.c_contiguous and out.shape[0] >= count
        self._check(self.lib.nemb_get_labels_rows(self.h, int(first), int(count), _p(out)))
        return out[:count]

    # ---- resample driver (include/nem_b200.h layer 4)
    def subsample_into(self, dst: "Engine", genome_mask, edge_presence_dev: int = 0):
        """Build the genome subsample `genome_mask` (bool[D] or packed uint32 words) of this
        pangenome into `dst` on the device; returns (n_eff, d_eff)."""
        words = pack_mask(genome_mask, self.d)
        n_eff, d_eff = C.c_int(), C.c_int()
        rc = self.lib.nemb_subsample(self.h, dst.h, _p(words), C.c_void_p(edge_presence_dev or None),
                                     C.byref(n_eff), C.byref(d_eff))
        if rc:
            raise NemError(rc, self.lib.nemb_last_error(dst.h).decode())
        dst.n, dst.d = n_eff.value, d_eff.value
        return n_eff.value, d_eff.value

    def family_index(self):
        out = np.zeros(self.n, dtype=np.int32)
        self._check(self.lib.nemb_get_family_index(self.h, _p(out)))
        return out

    def resample_batch(self, genome_masks, betas=None, n_workers=4, edge_presence_dev: int = 0, **kw):
        """Fit every genome subset of `genome_masks` (bool[R][D] or uint32[R][ceil(D/32)]);
        returns (votes int32[N][4] = P,S,C,U counts per family, iters int32[R], BatchStats)."""
        gm = np.asarray(genome_masks)
        if gm.dtype != np.uint32:
            gm = np.stack([pack_mask(m, self.d) for m in gm])
        gm = np.ascontiguousarray(gm, dtype=np.uint32)
        assert gm.shape[1] == (self.d + 31) // 32, gm.shape
        runs = gm.shape[0]
        o = make_options(**kw)
        bt = None if betas is None else _f32(betas)
        votes = np.zeros((self.n, 4), dtype=np.int32)
        iters = np.zeros(runs, dtype=np.int32)
        st = BatchStats()
        self._check(self.lib.nemb_resample_batch(self.h, runs, _p(gm), _p(bt), C.byref(o), int(n_workers),
                                                 C.c_void_p(edge_presence_dev or None), _p(votes),
                                                 _p(iters), C.byref(st)))
        return votes, iters, st

    # ---- stages
    def stage_density(self, prop, center, disp, k=3, force_general=False, want_hamming=False):
        prop, center, disp = _f32(prop), _f32(center), _f32(disp)
        out = np.zeros((self.n, k), dtype=np.float64)
        ham = np.full((self.n, k), -1, dtype=np.int32) if want_hamming else None
        used = C.c_int()
        self._check(self.lib.nemb_stage_density(self.h, k, _p(prop), _p(center), _p(disp),
                                                int(force_general), _p(out), _p(ham), C.byref(used)))
        self.k = k
        return out, ham, bool(used.value)

    def stage_sweep(self, logpf, beta, t, **kw):
        o = make_options(**kw)
        t = _f32(t).copy()
        lab = np.zeros(self.n, dtype=np.int32)
        rounds = C.c_int64()
        self._check(self.lib.nemb_stage_sweep(self.h, C.byref(o),
                                              _p(np.ascontiguousarray(logpf, dtype=np.float64)),
                                              C.c_float(beta), _p(t), _p(lab), C.byref(rounds)))
        self.k = o.k
        return t, lab, rounds.value

    def stage_mstep(self, t, prop0, center0, disp0, **kw):
        o = make_options(**kw)
        prop, center, disp = _f32(prop0).copy(), _f32(center0).copy(), _f32(disp0).copy()
        nk = np.zeros(o.k); skd = np.zeros((o.k, self.d)); empty = C.c_int()
        self._check(self.lib.nemb_stage_mstep(self.h, C.byref(o), _p(_f32(t)), _p(prop), _p(center),
                                              _p(disp), _p(nk), _p(skd), C.byref(empty)))
        self.k = o.k
        return empty.value, prop, center.reshape(o.k, self.d), disp.reshape(o.k, self.d), nk, skd

    def stage_criteria(self, logpf, t, beta, **kw):
        o = make_options(**kw)
        out = np.zeros(6)
        self._check(self.lib.nemb_stage_criteria(self.h, C.byref(o),
                                                 _p(np.ascontiguousarray(logpf, dtype=np.float64)),
                                                 _p(_f32(t)), C.c_float(beta), _p(out)))
        return dict(zip("UDLMZG", out))


# ------------------------------------------------------------------ communicators
def shard_range(n_glob: int, world: int, rank: int):
    """(shard_len, row0, n_loc) exactly as the engine computes them (nemb_shard_range)."""
    sl, r0, nl = C.c_int(), C.c_int(), C.c_int()
    load_library().nemb_shard_range(int(n_glob), int(world), int(rank), C.byref(sl), C.byref(r0),
                                    C.byref(nl))
    return sl.value, r0.value, nl.value


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    rc = load_library().nemb_nccl_unique_id(buf)
    if rc:
        raise NemError(rc, "ncclGetUniqueId (is libnccl.so.2 loadable?)")
    return bytes(buf)


def nccl_comm(uid: bytes, rank: int, world: int) -> int:
    """ncclCommInitRank on the CURRENT CUDA device; returns a nemb_comm* (int)."""
    out = C.c_void_p()
    buf = (C.c_uint8 * 128).from_buffer_copy(uid)
    rc = load_library().nemb_comm_create_nccl(C.byref(out), buf, int(rank), int(world))
    if rc:
        raise NemError(rc, f"nemb_comm_create_nccl(rank {rank} of {world})")
    return out.value


def local_comms(world: int) -> list:
    """In-process test double: `world` nemb_comm* for `world` threads on one device."""
    arr = (C.c_void_p * world)()
    rc = load_library().nemb_comm_create_local(arr, int(world))
    if rc:
        raise NemError(rc, "nemb_comm_create_local")
    return [arr[i] for i in range(world)]


def comm_destroy(comm_ptr) -> None:
    if comm_ptr:
        load_library().nemb_comm_destroy(C.c_void_p(comm_ptr))


# ------------------------------------------------------------------ host loader / writers (no GPU)
def read_files(base: str, k: int = 0) -> dict:
    """Parse <base>.str/.dat/.nei(/.m) with the engine's C loader; numpy copies of its buffers."""
    lib = load_library()
    hp = HostProblem()
    rc = lib.nemb_read_files(base.encode(), int(k), C.byref(hp))
    if rc != 0:
        raise NemError(rc, f"nemb_read_files({base})")
    try:
        n, d, wpr = hp.n, hp.d, hp.words_per_row
        out = dict(n=n, d=d, wpr=wpr, spatial=bool(hp.spatial), nnz=hp.nnz, max_neigh=hp.max_neigh,
                   m_flag=hp.m_flag,
                   x_packed=np.ctypeslib.as_array(hp.x_packed, shape=(n, wpr)).copy())
        if hp.spatial:
            out["row_ptr"] = np.ctypeslib.as_array(hp.row_ptr, shape=(n + 1,)).copy()
            nnz = max(hp.nnz, 1)
            out["col"] = np.ctypeslib.as_array(hp.col, shape=(nnz,)).copy()[:hp.nnz]
            out["wgt"] = np.ctypeslib.as_array(hp.wgt, shape=(nnz,)).copy()[:hp.nnz]
        if k > 0:
            out["prop"] = np.ctypeslib.as_array(hp.prop, shape=(k,)).copy()
            out["center"] = np.ctypeslib.as_array(hp.center, shape=(k, d)).copy()
            out["disp"] = np.ctypeslib.as_array(hp.disp, shape=(k, d)).copy()
        return out
    finally:
        lib.nemb_free_host_problem(C.byref(hp))


def write_uf(path: str, t) -> None:
    t = _f32(t)
    rc = load_library().nemb_write_uf(path.encode(), t.shape[0], t.shape[1], _p(t))
    if rc:
        raise NemError(rc, path)


def write_cf(path: str, label) -> None:
    label = np.ascontiguousarray(label, dtype=np.int32)
    rc = load_library().nemb_write_cf(path.encode(), label.shape[0], _p(label))
    if rc:
        raise NemError(rc, path)


def write_mf(path: str, crit: dict, beta: float, prop, center, disp) -> None:
    center = _f32(center); k, d = center.shape
    rc = load_library().nemb_write_mf(path.encode(), k, d, C.c_double(crit["U"]), C.c_double(crit["D"]),
                                      C.c_double(crit["L"]), C.c_double(crit["M"]), C.c_float(beta),
                                      _p(_f32(prop)), _p(center), _p(_f32(disp)))
    if rc:
        raise NemError(rc, path)


def nem(Fname, nk, algo, beta, convergence, convergence_th, format, it_max, dolog, model_family,
        proportion, dispersion, init_mode) -> int:
    """Keyword-callable mirror of the reference's Cython ``nem.nem`` (NEM/nem.pyx:1-14): bytes
    strings, same argument names and meaning, returns the ExitET code."""
    def b(v):
        return v if isinstance(v, bytes) else str(v).encode()
    return load_library().nem(b(Fname), int(nk), b(algo), float(beta), b(convergence),
                              float(convergence_th), b(format), int(it_max), int(bool(dolog)),
                              b(model_family), b(proportion), b(dispersion), int(init_mode))


def helper_pid() -> int:
    """pid of the helper process that serves this process's nem() calls (0: none) -- a process
    forked after CUDA was initialised cannot use CUDA itself (csrc/nem_api.c "forked callers")."""
    return int(load_library().nem_b200_helper_pid())


def nem_ex(Fname, nk, algo, beta, convergence, convergence_th, format, it_max, dolog, model_family,
           proportion, dispersion, init_mode, update="seq", sweep_impl="auto", device=-1,
           n_random_inits=0, seed=0, beta_mode="fix", psgrad=(0, 0.0, 0.0),
           heuristic=(0.0, 0.0, 0.0, 0.0, 0.0)) -> int:
    """nem_b200_ex: nem() plus the knobs the reference's CLI has (-U, -S, -B fix|psgrad|heu_d|heu_l,
    -G nit conv step, -H bstep bmax ddrop dloss lloss; zeros = the reference's defaults)."""
    def b(v):
        return v if isinstance(v, bytes) else str(v).encode()
    ex = Extra(UPDATE[update], SWEEP[sweep_impl], int(device), int(n_random_inits), int(seed),
               {"fix": 0, "psgrad": 1, "heu_d": 2, "heu_l": 3}[beta_mode], int(psgrad[0]),
               float(psgrad[1]), float(psgrad[2]), *[float(v) for v in heuristic])
    return load_library().nem_b200_ex(b(Fname), int(nk), b(algo), float(beta), b(convergence),
                                      float(convergence_th), b(format), int(it_max),
                                      int(bool(dolog)), b(model_family), b(proportion),
                                      b(dispersion), int(init_mode), C.byref(ex))
