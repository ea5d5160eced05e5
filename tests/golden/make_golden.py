#!/usr/bin/env python
"""Generates tests/golden/*.npz from OUTPUTS OF THE REFERENCE ITSELF (oracle #1).

The reference ships no tests, fixtures or golden vectors, so parity is
pinned on what the unmodified reference computes: this script writes the NEM input files of a
few small seeded pangenomes exactly like PPanGGOLiN does (ppanggolin.py:829-930), runs

  * oracle/_ref/nem_ref_harness  -- ClassifyByNem (nem_alg.c:546-584) with TIE_FIRST and a fixed
    seed, dumping ClassifM / parameters / criteria at full float32 precision, and
  * oracle/_ref/nem_ref_cli      -- nem() (nem_exe.c:239-704) exactly as ppanggolin.py:1814-1826
    calls it, keeping the .uf / .mf text it writes,

and stores inputs + reference outputs in one compressed .npz per case.  It needs
/root/reference (to build oracle/_ref) and therefore only runs in the build container:

    python tests/golden/make_golden.py

The committed .npz files are what tests/test_golden.py (CPU: oracle #2 vs reference) and
tests/test_gpu_golden.py (GPU: CUDA engine vs reference) read; neither needs the reference.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import nemo  # noqa: E402
from pangenomenem_b200 import synth  # noqa: E402

# name: (n, d, seed, graph, weighted, dict(harness/cli options))
CASES = {
    # exactly PPanGGOLiN's call (ppanggolin.py:1814-1826), BASELINE config 1 in miniature
    "ppanggolin_ncem_sk": (1200, 50, 42, "pangenome", True,
                           dict(algo="ncem", beta=0.5, disp="sk_", prop="pk", it_max=100)),
    # -fd free dispersion
    "ncem_skd": (900, 40, 7, "pangenome", True,
                 dict(algo="ncem", beta=0.5, disp="skd", prop="pk", it_max=100)),
    # fuzzy NEM, sequential update, fixed number of iterations
    "nem_seq_sk": (800, 36, 11, "pangenome", True,
                   dict(algo="nem", beta=0.5, disp="sk_", prop="pk", it_max=8)),
    # parallel (Jacobi) update, the semantics that shards across GPUs
    "ncem_para_s_d": (1000, 33, 5, "random", False,
                      dict(algo="ncem", beta=1.0, disp="s_d", prop="p_", it_max=15, update="para")),
    "nem_para_s__": (700, 64, 9, "chain", True,
                     dict(algo="nem", beta=0.3, disp="s__", prop="pk", it_max=6, update="para")),
    # pure Bernoulli mixture, type N file (beta forced to 0, nem_exe.c:570-574): config 2 shape
    "mixture_nonspatial": (1500, 96, 3, "none", True,
                           dict(algo="ncem", beta=0.5, disp="sk_", prop="pk", it_max=100)),
}

# beta estimation (the reference reaches it through its CLI's -B/-G/-H only):
# name: (n, d, seed, graph, weighted, options, beta_mode, beta_params)
# psgrad cases are unweighted: with co-presence weights the reference's float exp(beta * ctx)
# overflows, its gradient is NaN and it resets beta to 0 (nem_alg.c:2172-2191, 2222-2225)
BETA_CASES = {
    "beta_psgrad_ncem": (1000, 40, 21, "pangenome", False,
                         dict(algo="ncem", beta=0.5, disp="sk_", prop="pk", it_max=100),
                         "psgrad", (1, 0.001, 0.0)),
    "beta_psgrad_nem": (800, 36, 22, "pangenome", False,
                        dict(algo="nem", beta=0.3, disp="sk_", prop="pk", it_max=10),
                        "psgrad", (5, 0.001, 0.0)),
    "beta_psgrad_step": (900, 33, 23, "random", False,
                         dict(algo="ncem", beta=0.2, disp="skd", prop="pk", it_max=100),
                         "psgrad", (3, 0.001, 0.5)),
    "beta_heu_d": (1000, 40, 24, "pangenome", True,
                   dict(algo="ncem", beta=0.5, disp="sk_", prop="pk", it_max=100),
                   "heu_d", (0.1, 2.0, 0.8, 0.5, 0.02)),
    "beta_heu_l": (1000, 40, 25, "pangenome", True,
                   dict(algo="ncem", beta=0.5, disp="sk_", prop="pk", it_max=100),
                   "heu_l", (0.1, 2.0, 0.8, 0.5, 0.02)),
    "beta_heu_d_nem": (800, 36, 26, "pangenome", False,
                       dict(algo="nem", beta=0.5, disp="sk_", prop="pk", it_max=6),
                       "heu_d", (0.25, 1.0, 0.8, 0.5, 0.002)),
}


def main() -> None:
    nemo.build(ref=True)
    if not nemo.have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference)")
    todo = {k: v + ("fix", ()) for k, v in CASES.items()}
    todo.update(BETA_CASES)
    for name, (n, d, seed, graph, weighted, opt, beta_mode, beta_params) in todo.items():
        pg = synth.make_pangenome(n, d, seed=seed, graph=graph, weighted=weighted)
        spatial = graph != "none"
        with tempfile.TemporaryDirectory() as tmp:
            base = os.path.join(tmp, "nem_file")
            synth.write_nem_files(base, pg, spatial=spatial, weighted_flag=1 if weighted else 0)
            ref = nemo.run_ref_harness(base, os.path.join(tmp, "out"), k=3, tie="first", seed=42,
                                       beta_mode=beta_mode, beta_params=beta_params, **opt)
            cli = {}
            if beta_mode != "fix":
                cli = dict(beta_mode=beta_mode, beta_params=np.array(beta_params, dtype=np.float64),
                           ref_beta=ref["beta"], ref_beta_tested=np.array(ref["beta_tested"]))
            elif opt.get("update", "seq") == "seq":
                rc, _, _ = nemo.run_ref_cli(base, k=3, algo=opt["algo"], beta=opt["beta"],
                                            it_max=opt["it_max"], dolog=1, prop=opt["prop"],
                                            disp=opt["disp"])
                assert rc == 0, rc
                # .log of the reference's own nem(dolog=1) call: header + one row per EM iteration
                # (nem_alg.c:1478-1498, 1883-1946, 1995-2052, 2620-2646); its first line carries the date
                cli = dict(uf_text=np.frombuffer(open(base + ".uf", "rb").read(), dtype=np.uint8),
                           mf_text=np.frombuffer(open(base + ".mf", "rb").read(), dtype=np.uint8),
                           log_text=np.frombuffer(open(base + ".log", "rb").read(), dtype=np.uint8))
            files = {ext: np.frombuffer(open(base + "." + ext, "rb").read(), dtype=np.uint8)
                     for ext in (("str", "dat", "nei", "m") if spatial else ("str", "dat", "m"))}
        assert ref["status"] == 0 and not ref["density_zero"], name
        out = dict(
            x_bits=np.packbits(pg.x, axis=1, bitorder="little"), n=n, d=d,
            row_ptr=pg.row_ptr, col=pg.col, wgt=pg.wgt, spatial=spatial,
            opt_keys=np.array(sorted(opt)), opt_vals=np.array([str(opt[k]) for k in sorted(opt)]),
            ref_cm=ref["cm"], ref_prop=ref["prop"], ref_center=ref["center"], ref_disp=ref["disp"],
            ref_crit=np.array([ref["crit"][c] for c in "UDLMZG"]), ref_iters=ref["iters"],
            ref_converged=bool(ref["converged"]), **cli,
            **{"file_" + k: v for k, v in files.items()})
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: n={n} d={d} iters={ref['iters']} converged={ref['converged']} "
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
