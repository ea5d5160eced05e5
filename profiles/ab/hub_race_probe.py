#!/usr/bin/env python
"""Probe for the hub copy race of the margin-cached dense sweep round (DESIGN.md section 2.2,
knob JAC_HUB_GUARD).  NOT part of the test suite: written without a GPU at the end of round 1, to
be run on a B200 against a -DJAC_HUB_GUARD=0 variant and the default build (guard on).

    python profiles/ab/hub_race_probe.py [--families 600000] [--seeds 6]

The graph is built so that the race CAN fire: a chain of families with a weak data term (labels
follow the neighbourhood, so fits run for many iterations and labels keep moving) plus hubs
(degree ~40) placed at the HIGHEST ids, i.e. in the light CTAs of the last wave of the dense round,
long after the hub warps of the first blocks finished (the margin cache must be active for the
race to exist: watch `kept` in the output).  Every fit is compared with the oracle's
sequential sweep; any differing label is reported with the hub / non-hub split.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import nemo                                  # noqa: E402  (checker only)
from pangenomenem_b200 import capi, synth                # noqa: E402


def build(n, d, n_hubs, hub_deg, seed):
    rng = np.random.default_rng(seed)
    # latent classes in runs of 50 families with weakly separated densities: labels follow the
    # neighbourhood and keep moving for ~13 iterations, hubs included (tuned on the CPU oracle:
    # 20 000 families, 200 hubs, beta 0.5 -> 1-4 hubs change class in each of iterations 3-11)
    runs = np.repeat(rng.integers(0, 3, size=n // 50 + 1), 50)[:n]
    p = np.array([0.7, 0.5, 0.3])[runs]
    x = (rng.random((n, d)) < p[:, None]).astype(np.uint8)
    x[x.sum(axis=1) == 0, 0] = 1
    chain = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
    hubs = np.arange(n - n_hubs, n)
    he = np.stack([np.repeat(hubs, hub_deg), rng.integers(0, n - n_hubs, size=n_hubs * hub_deg)], axis=1)
    edges = np.unique(np.sort(np.concatenate([chain, he]), axis=1), axis=0)
    edges = edges[edges[:, 0] != edges[:, 1]]
    w = np.ones(edges.shape[0], dtype=np.float32)
    row_ptr, col, wgt = synth.edges_to_csr(n, edges, w)
    return x, row_ptr, col, wgt, hubs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=600_000)
    ap.add_argument("--genomes", type=int, default=32)
    ap.add_argument("--hubs", type=int, default=3000)
    ap.add_argument("--seeds", type=int, default=6)
    ap.add_argument("--beta", type=float, default=0.5)
    a = ap.parse_args()
    eng = capi.Engine(0)
    total_bad = 0
    for seed in range(a.seeds):
        x, row_ptr, col, wgt, hubs = build(a.families, a.genomes, a.hubs, 40, seed)
        theta = nemo.default_theta(3, a.genomes)
        kw = dict(k=3, algo="ncem", beta=a.beta, disp="sk_", prop="pk", it_max=40)
        ref = nemo.Problem(x, row_ptr, col, wgt, **kw).fit(*theta)
        eng.load_dense(x, row_ptr, col, wgt)
        for rep in range(3):                              # schedules differ from fit to fit
            got = eng.fit(*theta, **kw)
            lab = eng.labels()
            bad = np.flatnonzero(lab != ref.label)
            on_hubs = int(np.isin(bad, hubs).sum())
            total_bad += bad.size
            print(f"seed {seed} fit {rep}: iters {got.iters} (oracle {ref.iters}), kept {got.n_kept}, "
                  f"labels differing {bad.size} (hubs {on_hubs})", flush=True)
    eng.close()
    print("RACE OBSERVED" if total_bad else "no difference from the oracle")
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
