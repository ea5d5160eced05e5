"""C5 (resample driver) probe: how the time of a run splits into subsample build and fit, and how
it moves with the number of concurrent workers.  Run on a GPU box:  python profiles/c5_probe.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from pangenomenem_b200 import capi, synth, synth_gpu

dev = torch.device("cuda", 0)
n, d, beta, graph = bench.WORKLOADS["c5"]
xdev, _ = synth_gpu.make_packed_on_device(n, d, seed=42, device=dev)
xh = xdev.cpu().numpy()
row_ptr, col, wgt = synth_gpu.make_graph(n, xh, seed=42, kind=graph)
runs = int(os.environ.get("RUNS", "256"))
masks, betas, sizes = bench.c5_plan(d, runs)
eng = capi.Engine(0)
eng.load_packed(xh.view(np.uint32), d, row_ptr, col, wgt)
opts = dict(k=3, algo="ncem", update="seq", conv="clas", conv_thr=1e-8, it_max=100, prop="pk", disp="sk_")
for w in [int(a) for a in os.environ.get("WORKERS", "1,2,4,8,16").split(",")]:
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        votes, iters, st = eng.resample_batch(masks, betas, n_workers=w, **opts)
        torch.cuda.synchronize(); wall = time.time() - t0
    print(json.dumps({"workers": w, "runs": runs, "wall_ms_per_run": 1e3 * wall / runs,
                      "fit_ms_mean": st.fit_ms_sum / max(1, st.n_runs), "value": st.family_iterations / wall,
                      "iters_mean": float(np.mean(iters)), "launches_per_run": st.kernel_launches / max(1, st.n_runs)}), flush=True)
# build only: subsample into a scratch engine, timed with events
dst = capi.Engine(0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if True:
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        for r in range(32):
            eng.subsample_into(dst, masks[r % runs])
        torch.cuda.synchronize(); wall = time.time() - t0
    print(json.dumps({"build_only_ms_per_run": 1e3 * wall / 32}))
    # one fit alone on the last subsample
    th = synth.default_theta(3, dst.d)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        f = dst.fit(*th, beta=float(betas[31 % runs]), **opts)
        torch.cuda.synchronize(); wall = time.time() - t0
    print(json.dumps({"fit_alone_ms": 1e3 * wall, "n": dst.n, "d": dst.d, "iters": f.iters, "fit_ms_engine": f.fit_ms if hasattr(f, 'fit_ms') else None}))
