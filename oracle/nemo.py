"""ctypes bridge to the oracles -- TEST INFRASTRUCTURE, never imported by the product package.

* oracle #2: ``oracle/_build/libnem_oracle.so`` (our float64 restatement, nem_oracle.c)
* oracle #1: ``oracle/_ref/nem_ref_cli`` and ``oracle/_ref/nem_ref_harness`` (the unmodified
  reference compiled in place from /root/reference by ``make -C oracle ref``)

Allowed importers: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libnem_oracle.so")
REF_CLI = os.path.join(HERE, "_ref", "nem_ref_cli")
REF_HARNESS = os.path.join(HERE, "_ref", "nem_ref_harness")

ALGO = {"nem": 0, "ncem": 1}
UPDATE = {"seq": 0, "para": 1}
CONV = {"none": 0, "clas": 1, "crit": 2}
PROP = {"p_": 0, "pk": 1}
DISP = {"s__": 0, "sk_": 1, "s_d": 2, "skd": 3}


def build(ref: bool = True) -> None:
    """Compile the restatement and, when /root/reference exists, the reference itself."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/ppanggolin/NEM"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.access(REF_CLI, os.X_OK) and os.access(REF_HARNESS, os.X_OK)


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("d", C.c_int), ("k", C.c_int),
                ("x", C.c_void_p), ("row_ptr", C.c_void_p), ("col", C.c_void_p),
                ("wgt", C.c_void_p),
                ("algo", C.c_int), ("update", C.c_int), ("conv", C.c_int),
                ("prop", C.c_int), ("disp", C.c_int),
                ("beta", C.c_float), ("conv_thr", C.c_float),
                ("it_max", C.c_int), ("param_fixed", C.c_int), ("dolog", C.c_int)]


class _Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("converged", C.c_int),
                ("U", C.c_double), ("D", C.c_double), ("L", C.c_double), ("M", C.c_double),
                ("Z", C.c_double), ("G", C.c_double),
                ("n_allnul", C.c_int64), ("n_ties", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        _lib = C.CDLL(LIB_PATH)
        _lib.nemo_levels.restype = C.c_int
        _lib.nemo_mstep.restype = C.c_int
        _lib.nemo_fit.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Fit:
    status: int
    iters: int
    converged: bool
    t: np.ndarray
    label: np.ndarray
    prop: np.ndarray
    center: np.ndarray
    disp: np.ndarray
    crit: dict
    n_allnul: int
    n_ties: int


class Problem:
    """Holds the numpy buffers alive next to the C struct."""

    def __init__(self, x, row_ptr=None, col=None, wgt=None, k=3, algo="ncem", update="seq",
                 beta=0.5, conv="clas", conv_thr=1e-8, it_max=100, prop="pk", disp="sk_",
                 param_fixed=False, dolog=False):
        self.x = np.ascontiguousarray(x, dtype=np.uint8)
        self.row_ptr = None if row_ptr is None else np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.col = None if col is None else np.ascontiguousarray(col, dtype=np.int32)
        self.wgt = None if wgt is None else np.ascontiguousarray(wgt, dtype=np.float32)
        n, d = self.x.shape
        self.n, self.d, self.k = n, d, k
        self.c = _Problem(n, d, k, _p(self.x), _p(self.row_ptr), _p(self.col), _p(self.wgt),
                          ALGO[algo], UPDATE[update], CONV[conv], PROP[prop], DISP[disp],
                          float(beta), float(conv_thr), int(it_max), int(param_fixed), int(dolog))

    # ---- stages
    def hamming(self, center, disp):
        h = np.zeros((self.n, self.k), dtype=np.int32)
        lib().nemo_hamming(C.byref(self.c), _p(_f32(center)), _p(_f32(disp)), _p(h))
        return h

    def logpf(self, prop, center, disp):
        out = np.zeros((self.n, self.k), dtype=np.float64)
        lib().nemo_logpf(C.byref(self.c), _p(_f32(prop)), _p(_f32(center)), _p(_f32(disp)), _p(out))
        return out

    def sweep(self, logpf, beta, t):
        t = np.ascontiguousarray(t, dtype=np.float32).copy()
        label = np.zeros(self.n, dtype=np.int32)
        lib().nemo_sweep(C.byref(self.c), _p(np.ascontiguousarray(logpf, dtype=np.float64)),
                         C.c_double(float(beta)), _p(t), _p(label))
        return t, label

    def mstep(self, t, prop, center, disp):
        prop, center, disp = _f32(prop).copy(), _f32(center).copy(), _f32(disp).copy()
        nk = np.zeros(self.k); skd = np.zeros((self.k, self.d))
        st = lib().nemo_mstep(C.byref(self.c), _p(np.ascontiguousarray(t, dtype=np.float32)),
                              _p(prop), _p(center), _p(disp), _p(nk), _p(skd))
        return st, prop, center, disp, nk, skd

    def criteria(self, logpf, t, beta):
        out = np.zeros(6)
        lib().nemo_criteria(C.byref(self.c), _p(np.ascontiguousarray(logpf, dtype=np.float64)),
                            _p(np.ascontiguousarray(t, dtype=np.float32)),
                            C.c_double(float(beta)), _p(out))
        return dict(zip("UDLMZG", out))

    def fit(self, prop, center, disp) -> Fit:
        prop, center, disp = _f32(prop).copy(), _f32(center).copy(), _f32(disp).copy()
        t = np.zeros((self.n, self.k), dtype=np.float32)
        label = np.zeros(self.n, dtype=np.int32)
        r = _Result()
        lib().nemo_fit(C.byref(self.c), _p(prop), _p(center), _p(disp), _p(t), _p(label),
                       C.byref(r))
        return Fit(r.status, r.iters, bool(r.converged), t, label, prop,
                   center.reshape(self.k, self.d), disp.reshape(self.k, self.d),
                   dict(U=r.U, D=r.D, L=r.L, M=r.M, Z=r.Z, G=r.G), r.n_allnul, r.n_ties)


    # -- beta estimation ------------------------------------------------------
    def estim_beta(self, t, beta, n_iter=1, conv_thr=0.001, step=0.0):
        """EstimBeta BETA_PSGRAD (nem_alg.c:2120-2230) -> (new beta, dict(crit, grad, dsec))."""
        g = _PsGrad(int(n_iter), float(conv_thr), float(step))
        t = _f32(t)
        out = np.zeros(3, dtype=np.float64)
        f = lib().nemo_estim_beta
        f.restype = C.c_float
        b = f(C.byref(self.c), C.byref(g), _p(t), C.c_float(float(beta)), _p(out))
        return float(b), dict(zip(("crit", "grad", "dsec"), out))

    def fit_ex(self, prop, center, disp, t_init=None, psgrad=None) -> Fit:
        """nemo_fit_ex: optional starting classification (INIT_FILE) and psgrad=(nit, conv, step)."""
        prop, center, disp = _f32(prop).copy(), _f32(center).copy(), _f32(disp).copy()
        t = np.zeros((self.n, self.k), dtype=np.float32) if t_init is None else _f32(t_init).copy()
        label = np.zeros(self.n, dtype=np.int32)
        r = _Result()
        g = None if psgrad is None else _PsGrad(int(psgrad[0]), float(psgrad[1]), float(psgrad[2]))
        bo = C.c_float(0.0)
        lib().nemo_fit_ex(C.byref(self.c), int(t_init is not None),
                          None if g is None else C.byref(g), _p(prop), _p(center), _p(disp),
                          _p(t), _p(label), C.byref(r), C.byref(bo))
        f = Fit(r.status, r.iters, bool(r.converged), t, label, prop,
                center.reshape(self.k, self.d), disp.reshape(self.k, self.d),
                dict(U=r.U, D=r.D, L=r.L, M=r.M, Z=r.Z, G=r.G), r.n_allnul, r.n_ties)
        f.beta = float(bo.value)
        return f

    def fit_heuristic(self, prop, center, disp, mode="heu_d", step=0.1, bmax=2.0, ddrop=0.8,
                      dloss=0.5, lloss=0.02) -> Fit:
        """ClassifyByNemHeuBeta (nem_alg.c:731-992); Fit gains beta, beta_tested, crit_tested."""
        prop, center, disp = _f32(prop).copy(), _f32(center).copy(), _f32(disp).copy()
        t = np.zeros((self.n, self.k), dtype=np.float32)
        label = np.zeros(self.n, dtype=np.int32)
        r = _Result()
        hp = _Heu(float(step), float(bmax), float(ddrop), float(dloss), float(lloss))
        cap = int(bmax / step) + 3
        bt = np.zeros(cap, dtype=np.float32)
        ct = np.zeros(cap, dtype=np.float32)
        be, nt = C.c_float(0.0), C.c_int(0)
        lib().nemo_fit_heuristic(C.byref(self.c), {"heu_d": 0, "heu_l": 1}[mode], C.byref(hp),
                                 _p(prop), _p(center), _p(disp), _p(t), _p(label), C.byref(r),
                                 C.byref(be), C.byref(nt), _p(bt), _p(ct), cap)
        f = Fit(r.status, r.iters, bool(r.converged), t, label, prop,
                center.reshape(self.k, self.d), disp.reshape(self.k, self.d),
                dict(U=r.U, D=r.D, L=r.L, M=r.M, Z=r.Z, G=r.G), r.n_allnul, r.n_ties)
        f.beta = float(be.value)
        f.beta_tested = bt[:nt.value].copy()
        f.crit_tested = ct[:nt.value].copy()
        return f


class _PsGrad(C.Structure):
    _fields_ = [("n_iter", C.c_int), ("conv_thr", C.c_float), ("step", C.c_float)]


class _Heu(C.Structure):
    _fields_ = [("step", C.c_float), ("max", C.c_float), ("ddrop", C.c_float),
                ("dloss", C.c_float), ("lloss", C.c_float)]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pack(x, words_per_row):
    x = np.ascontiguousarray(x, dtype=np.uint8)
    out = np.zeros((x.shape[0], words_per_row), dtype=np.uint32)
    lib().nemo_pack(_p(x), x.shape[0], x.shape[1], words_per_row, _p(out))
    return out


def levels(row_ptr, col):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    lv = np.zeros(row_ptr.shape[0] - 1, dtype=np.int32)
    depth = lib().nemo_levels(lv.shape[0], _p(row_ptr), _p(col), _p(lv))
    return lv, depth


def default_theta(k: int, d: int, low_disp: float = 0.1):
    """PPanGGOLiN's default .m (ppanggolin.py:893-901) as the reference reads it
    (nem_exe.c:1022-1034: last proportion = 1 - sum of the others, in float)."""
    assert k == 3
    prop = np.array([0.33333, 0.33333, 0.0], dtype=np.float32)
    prop[2] = np.float32(np.float32(1.0) - prop[0]) - prop[1]
    center = np.repeat(np.array([1.0, 0.5, 0.0], dtype=np.float32)[:, None], d, axis=1)
    disp = np.repeat(np.array([low_disp, 0.5, low_disp], dtype=np.float32)[:, None], d, axis=1)
    return prop, center, disp


# ------------------------------------------------------------------ oracle #1 runners

def run_ref_cli(base, k=3, algo="ncem", beta=0.5, conv="clas", thr=1e-8, fmt="fuzzy",
                it_max=100, dolog=0, family="bern", prop="pk", disp="sk_", init=2, timeout=3600):
    """The reference's nem() exactly as PPanGGOLiN calls it (ppanggolin.py:1814-1826)."""
    args = [REF_CLI, base, str(k), algo, repr(float(beta)), conv, repr(float(thr)), fmt,
            str(it_max), str(int(dolog)), family, prop, disp, str(init)]
    cp = subprocess.run(args, capture_output=True, text=True, timeout=timeout)
    return cp.returncode, cp.stdout, cp.stderr


def run_ref_harness(base, out_prefix, k=3, algo="ncem", beta=0.5, conv="clas", thr=1e-8,
                    it_max=100, family="bern", prop="pk", disp="sk_", init=2, update="seq",
                    tie="first", seed=42, timeout=3600, beta_mode="fix", beta_params=()):
    """ClassifyByNem with the hidden knobs; returns dict(cm, prop, center, disp, crit, iters).
    beta_mode fix|psgrad|heu_d|heu_l with beta_params (nit, conv, step) or (bstep, bmax, ddrop,
    dloss, lloss) -- the CLI's -B/-G/-H (nem_hlp.c:220-245)."""
    args = [REF_HARNESS, base, str(k), algo, repr(float(beta)), conv, repr(float(thr)),
            str(it_max), family, prop, disp, str(init), update, tie, str(seed), out_prefix]
    if beta_mode != "fix":
        args += [beta_mode] + [repr(float(v)) if not isinstance(v, int) else str(v)
                               for v in beta_params]
    subprocess.run(args, check=True, timeout=timeout, capture_output=True)
    with open(base + ".str") as f:
        _, n, d = f.read().split()[:3]
    n, d = int(n), int(d)
    cm = np.fromfile(out_prefix + ".cm.f32", dtype=np.float32).reshape(n, k)
    par = np.fromfile(out_prefix + ".par.f32", dtype=np.float32)
    vals = open(out_prefix + ".txt").read().split()
    txt = open(out_prefix + ".stderr").read()
    m = re.findall(r"NEM (converged|did not converge) after (\d+) iterations", txt)
    return dict(cm=cm, prop=par[:k], center=par[k:k + k * d].reshape(k, d),
                disp=par[k + k * d:].reshape(k, d), status=int(vals[0]),
                crit=dict(zip("UDLMZG", [float(v) for v in vals[1:7]])),
                beta=float(vals[7]) if len(vals) > 7 else None,
                beta_tested=[float(v) for v in re.findall(r"Testing beta = +([-0-9.]+)", txt)],
                iters=int(m[-1][1]) if m else None,
                converged=(m[-1][0] == "converged") if m else None,
                density_zero=("density = 0" in txt))
