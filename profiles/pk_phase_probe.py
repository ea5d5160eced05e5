#!/usr/bin/env python
"""Per-phase profile of the persistent EM kernel (CTA 0's globaltimer, barrier waits included).
    python profiles/pk_phase_probe.py [c1 c2 c3 c4] [--grid G]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pangenomenem_b200 import capi, synth, synth_gpu  # noqa: E402

W = {"c1": (20_000, 50, 0.5, "pangenome"), "c2": (100_000, 500, 0.0, "none"),
     "c3": (250_000, 1000, 0.5, "pangenome"), "c4": (1_000_000, 5000, 0.5, "pangenome")}


def main():
    names = [a for a in sys.argv[1:] if a in W] or ["c1", "c2", "c3", "c4"]
    dev = torch.device("cuda", 0)
    for w in names:
        n, d, beta, graph = W[w]
        xdev, _ = synth_gpu.make_packed_on_device(n, d, seed=42, device=dev)
        xh = xdev.cpu().numpy()
        if beta != 0 and graph != "none":
            row_ptr, col, wgt = synth_gpu.make_graph(n, xh, seed=42, kind=graph)
        else:
            row_ptr = col = wgt = None
        eng = capi.Engine(0)
        eng.load_shard_device(xdev.data_ptr(), n, 0, n, d, xdev.shape[1], row_ptr, col, wgt)
        theta0 = synth.default_theta(3, d)
        opts = dict(k=3, algo="ncem", update="seq", beta=beta, conv="clas", conv_thr=1e-8, it_max=100,
                    prop="pk", disp="sk_", sweep_impl="auto")
        for _ in range(3):
            f = eng.fit(*theta0, **opts)
        ms = []
        for _ in range(5):
            f = eng.fit(*theta0, **opts)
            ms.append(f.fit_ms)
        prof = eng.persist_profile()
        tot = sum(v for k, v in prof.items() if k != "fixup_rounds")
        print(json.dumps({"workload": w, "fit_ms": round(float(np.median(ms)), 4), "iters": f.iters,
                          "kernel_launches": f.kernel_launches, "pk": f.pk,
                          "phase_us": {k: round(v, 1) for k, v in prof.items()},
                          "phase_total_us": round(tot, 1), "kept": f.n_kept}))
        tr = eng.persist_trace()
        print("   iter:   scan  delta  final margin   eval  fixup | active rounds   (us, last launch of the fit)")
        for i, row in enumerate(tr):
            if row[:6].sum() > 0:
                print("   %4d: %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f | %7d %3d" % ((i,) + tuple(row[:6]) + (int(row[6]), int(row[7]))))
        eng.close()


if __name__ == "__main__":
    main()
