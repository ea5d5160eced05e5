"""C5 concurrency sweep: workers x fits in flight x CTAs per fit x builders in flight
(NEM_B200_BATCH_FITS / _GRID / _BUILDS).  python profiles/c5_sweep.py  on a GPU box."""
import itertools, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from pangenomenem_b200 import capi, synth_gpu

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
ref_votes = None
res = []
W = [int(a) for a in os.environ.get("W", "8,12").split(",")]
F = [int(a) for a in os.environ.get("F", "2,3,4,6").split(",")]
G = [int(a) for a in os.environ.get("G", "24,37,74,148").split(",")]
B = [int(a) for a in os.environ.get("B", "0,1").split(",")]
for w, f, g, b in itertools.product(W, F, G, B):
    os.environ["NEM_B200_BATCH_FITS"] = str(f)
    os.environ["NEM_B200_BATCH_GRID"] = str(g)
    os.environ["NEM_B200_BATCH_BUILDS"] = str(b)
    best = None
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        votes, iters, st = eng.resample_batch(masks, betas, n_workers=w, **opts)
        torch.cuda.synchronize(); wall = time.time() - t0
        best = wall if best is None else min(best, wall)
    if ref_votes is None:
        ref_votes = votes
    same = bool(np.array_equal(votes, ref_votes))
    rec = {"workers": w, "fits": f, "grid": g, "builds": b, "ms_per_run": 1e3 * best / runs,
           "value": st.family_iterations / best, "fit_ms_mean": st.fit_ms_sum / max(1, st.n_runs), "votes_same": same}
    res.append(rec)
    print(json.dumps(rec), flush=True)
res.sort(key=lambda r: r["ms_per_run"])
print("BEST", json.dumps(res[:5]))
