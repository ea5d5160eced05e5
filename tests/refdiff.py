"""Classify every family whose hard label differs from the unmodified reference's.

north_star: "final hard partitions must be identical, except for families whose top-two posterior
margin is below 1e-4; the count of those families is reported" -- and SURVEY.md "five things" #4:
at D >= 500 the reference's LINEAR-domain products p_k f_k(x_i) (double, nem_alg.c:2282, 2589-2613)
underflow, which the log-domain engine gates instead of copying.  A differing family must fall in
one of these classes, anything else fails the test that calls this:

  margin     top-two posterior margin (softmax of the final scores) below 1e-4
  tie        exact score tie (the reference breaks it with random(), nem_alg.c:617-640)
  underflow  the class the float64 evaluation picks -- or every class -- has p_k f_k(x_i) below
             DBL_MIN in the reference's linear domain (log < -708.396): its numerator is 0 or a
             denormal there whatever beta * ctx adds afterwards
  cascade    downstream of the families above: a neighbour is itself a differing family, or the
             reference's final theta differs (its M-step counted the families above in another
             class), and the float64 evaluation of the site under the REFERENCE's theta and the
             REFERENCE's neighbour labels picks the reference's label -- both partitions are
             self-consistent there.  Only accepted when a root cause (one of the three classes
             above) exists somewhere in the pangenome.

Test infrastructure (CPU only, numpy).
"""
from __future__ import annotations

import numpy as np

LOG_DBL_MIN = -708.3964185322641


def classify(pg, beta, logpf, label_ours, label_ref, logpf_ref=None):
    """logpf: float64 [N, K] under OUR final theta, logpf_ref under the reference's (None: same
    theta); returns dict(category -> list of families) with a key 'unexplained' that must stay
    empty."""
    if logpf_ref is None:
        logpf_ref = logpf
    k = logpf.shape[1]
    out = {"margin": [], "tie": [], "underflow": [], "cascade": [], "unexplained": []}
    differ = np.flatnonzero(label_ours != label_ref)
    dset = set(differ.tolist())

    def ctx_of(i, lab):
        c = np.zeros(k)
        if pg.row_ptr is not None:
            for e in range(pg.row_ptr[i], pg.row_ptr[i + 1]):
                c[lab[pg.col[e]]] += float(pg.wgt[e])
        return c

    for i in differ.tolist():
        sc = logpf[i] + beta * ctx_of(i, label_ours)
        order = np.argsort(-sc, kind="stable")
        gap = sc[order[0]] - sc[order[1]]
        post = np.exp(sc - sc.max())
        post /= post.sum()
        ps = np.sort(post)
        if gap == 0.0:
            out["tie"].append(i)
        elif ps[-1] - ps[-2] < 1e-4:
            out["margin"].append(i)
        elif logpf[i, label_ours[i]] < LOG_DBL_MIN or logpf[i].max() < LOG_DBL_MIN:
            out["underflow"].append(i)
        else:
            has_diff_nb = pg.row_ptr is not None and any(
                int(pg.col[e]) in dset for e in range(pg.row_ptr[i], pg.row_ptr[i + 1]))
            theta_moved = not np.array_equal(logpf_ref[i], logpf[i])
            sc_ref = logpf_ref[i] + beta * ctx_of(i, label_ref)
            # the reference's own products may underflow for the class it would otherwise pick
            lin_ok = logpf_ref[i] >= LOG_DBL_MIN
            cand = np.where(lin_ok, sc_ref, -np.inf)
            if (has_diff_nb or theta_moved) and int(np.argmax(cand)) == int(label_ref[i]):
                out["cascade"].append(i)
            else:
                out["unexplained"].append(i)
    if out["cascade"] and not (out["margin"] or out["tie"] or out["underflow"]):
        out["unexplained"] += out["cascade"]      # a cascade needs a root cause
        out["cascade"] = []
    return out
