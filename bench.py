#!/usr/bin/env python
"""bench.py -- NEM family-iterations/s on B200 (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c1|c5|dropin] [--impl reference]

One "step" = one complete NEM fit (blind sweep, beta sweep, EM iterations until the `clas`
convergence test, final criteria) of the synthetic pangenome, run exactly as PPanGGOLiN runs it
(ncem, sequential sweep, bern pk sk_, K=3; reference ppanggolin.py:1814-1826).
value = N x (EM iterations executed) / fit seconds, inputs resident in HBM.
e2e   = same metric through the C ABI with HOST buffers (H2D of packed X + CSR, graph
        preprocessing, fit, D2H of the labels and theta inside the timed region).
N > 1 : one process per GPU (torchrun).  Default --mode sharded: ONE pangenome of N x families
        row-sharded over the GPUs (BASELINE config 4: X sharded, NCCL all-gathers of the M-step
        statistics and of the labels inside the exact sequential sweep) => weak scaling with real
        collectives.  --mode replicas: every rank fits an independent replica (BASELINE config 5).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (families, genomes, beta, graph)
    "c1": (20_000, 50, 0.5, "pangenome"),       # BASELINE config 1 (reference's CPU-runnable case)
    "c2": (100_000, 500, 0.0, "none"),          # pure Bernoulli mixture
    "c3": (250_000, 1000, 0.5, "pangenome"),    # full NEM, 1 GPU
    "c4": (1_000_000, 5000, 0.5, "pangenome"),  # large pangenome, the metric's 1/2/4/8-GPU config
    # BASELINE config 5: 1024 independent fits on genome subsamples of a C3-like pangenome
    # (100..900 of the 1000 genomes) with beta swept over 0..1, replicas spread over the GPUs
    "c5": (250_000, 1000, 0.5, "pangenome"),
}
C5_RUNS = 1024
K = 3


def measured_traffic(kernel_prefix):
    """DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture
    (profiles/roofline_traffic.json, written from profiles/*_full.csv); None when absent."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    for name, rec in t.items():
        if name.startswith(kernel_prefix):
            return rec.get("dram_bytes_per_launch"), rec.get("source")
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm
def write_sample_files(base, x_rows, row_ptr, col, wgt):
    from pangenomenem_b200 import synth
    pg = synth.Pangenome(x=x_rows, row_ptr=row_ptr, col=col, wgt=wgt,
                         latent=np.zeros(x_rows.shape[0], dtype=np.int8))
    synth.write_nem_files(base, pg)


def make_cpu_sample(n_s, d, beta, graph, seed):
    """A bounded sample of the workload: n_s families of the same D, its own pangenome-like graph."""
    from pangenomenem_b200 import synth
    pg = synth.make_pangenome(n_s, d, seed=seed, graph=graph if beta != 0 else "none")
    return pg


def run_reference_pairs(bases, beta, n_s, spatial):
    """All samples in parallel, it_max=1 then it_max=3 with convergence=none; returns
    (family-iterations/s aggregate, seconds of the 2 extra iterations, T1, T3)."""
    from oracle import nemo

    def launch(itmax):
        t0 = time.time()
        procs = [subprocess.Popen([nemo.REF_CLI, b, str(K), "ncem", repr(float(beta)), "none", "0.01",
                                   "fuzzy", str(itmax), "0", "bern", "pk", "sk_", "2"],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                 for b in bases]
        for p in procs:
            p.wait()
        return time.time() - t0
    t1 = launch(1)
    t3 = launch(3)
    dt = max(t3 - t1, 1e-9)
    return len(bases) * n_s * 2 / dt, dt, t1, t3


def run_port(pg, beta):
    from oracle import nemo
    pb = nemo.Problem(pg.x, pg.row_ptr if beta else None, pg.col if beta else None,
                      pg.wgt if beta else None, k=K, algo="ncem", beta=beta, conv="none", it_max=3)
    t0 = time.time(); pb.fit(*nemo.default_theta(K, pg.d)); t3 = time.time() - t0
    pb = nemo.Problem(pg.x, pg.row_ptr if beta else None, pg.col if beta else None,
                      pg.wgt if beta else None, k=K, algo="ncem", beta=beta, conv="none", it_max=1)
    t0 = time.time(); pb.fit(*nemo.default_theta(K, pg.d)); t1 = time.time() - t0
    return pg.n * 2 / max(t3 - t1, 1e-9)


def cpu_baseline(workload, cores, budget_cells=6.0e7):
    """Times the reference (oracle/_ref, kind "reference") or, if it is absent, the C port on a
    bounded sample: `cores` independent samples in parallel, one process per core."""
    from oracle import nemo
    n, d, beta, graph = WORKLOADS[workload]
    n_s = int(max(500, min(n, budget_cells / d / 3)))
    tmp = tempfile.mkdtemp(prefix="nem_cpu_")
    if nemo.have_ref():
        bases = []
        for c in range(cores):
            pg = make_cpu_sample(n_s, d, beta, graph, seed=100 + c)
            base = os.path.join(tmp, f"s{c}", "nem_file")
            from pangenomenem_b200 import synth
            synth.write_nem_files(base, pg, spatial=beta != 0)
            bases.append(base)
        rate, dt, t1, t3 = run_reference_pairs(bases, beta, n_s, beta != 0)
        kind = "reference"
        note = (f"{cores} x ({n_s} families x {d} genomes) samples of {workload}, unmodified reference "
                f"nem() via oracle/_ref/nem_ref_cli, ncem beta={beta} sk_ pk, per-iteration time = "
                f"(T[it_max=3]-T[it_max=1])/2 = {dt / 2:.3f}s (file parsing and the per-genome qsort "
                f"excluded; T1={t1:.1f}s T3={t3:.1f}s)")
        if d >= 2000:
            note += "; at this D the reference's linear-domain densities underflow (results degenerate, timing still representative)"
    else:
        pg = make_cpu_sample(n_s, d, beta, graph, seed=100)
        rate = run_port(pg, beta)
        kind, cores = "port", 1
        note = f"{n_s} families x {d} genomes sample of {workload}, oracle/nem_oracle.c (float64 C port), 1 thread"
    return {"value": rate, "unit": "family-iterations/s", "cores": cores, "kind": kind, "sample": note}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, last, t_steps = [], None, 0.0
    for step in range(args.warmup + args.steps):
        t0 = time.time()
        last = cpu_baseline(args.workload, cores, budget_cells=2.0e7)
        if step >= args.warmup:
            vals.append(last["value"])
            t_steps += time.time() - t0
    n, d, beta, graph = WORKLOADS[args.workload]
    v = float(np.mean(vals))
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "NEM family-iterations/s", "value": v,
        "unit": "family-iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_steps * 1e3 / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload) + " -- measured on a BOUNDED SAMPLE of it: "
                               + last["sample"].split(",")[0],
                   "families": n, "genomes": d, "K": K, "beta": beta, "algo": "ncem", "update": "seq",
                   "same_config_as_gpu_arm": False,
                   "why": "the reference stores X as float[N*D] plus an int[N*D] sort index (40 GB at 1M x 5000) and "
                          "takes ~7 s per iteration at 100k x 500 on one core: the full workload does not finish; "
                          "per-family-iteration cost is what the sample measures"},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "family-iterations/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main_dropin(args):
    """--workload dropin: the call PPanGGOLiN makes (ppanggolin.py:1814-1826) -- nem() on the text
    files it writes (C2 shape: 100 000 families x 500 genomes, with a pangenome-like .nei,
    beta = 0.5), wall seconds of the whole call: parse .str/.dat/.nei/.m, upload, fit, write
    .uf/.mf.  Ours in-process through the C ABI (first call = CUDA context creation included, then
    warm calls); the unmodified reference (oracle/_ref/nem_ref_cli) once on the same files.  Not
    the round's headline metric: it shows what a drop-in user sees, where text I/O dominates."""
    from oracle import nemo
    from pangenomenem_b200 import capi, synth
    n, d, beta = (args.rows or 100_000), 500, 0.5
    tmp = tempfile.mkdtemp(prefix="nem_dropin_")
    base = os.path.join(tmp, "nem_file")
    pg = synth.make_pangenome(n, d, seed=42)
    synth.write_nem_files(base, pg)
    call = dict(Fname=base.encode(), nk=3, algo=b"ncem", beta=beta, convergence=b"clas",
                convergence_th=1e-8, format=b"fuzzy", it_max=100, dolog=False, model_family=b"bern",
                proportion=b"pk", dispersion=b"sk_", init_mode=2)
    times = []
    for _ in range(1 + args.warmup + args.steps):
        for ext in (".uf", ".mf"):
            if os.path.exists(base + ext):
                os.remove(base + ext)
        t0 = time.time()
        rc = capi.nem(**call)
        times.append(time.time() - t0)
        assert rc == 0 and os.path.exists(base + ".uf"), rc
    uf = synth.read_uf(base + ".uf", 3)
    warm = times[1 + args.warmup:]
    line = {"metric": "nem() drop-in call on PPanGGOLiN's files, wall seconds", "value": float(np.mean(warm)),
            "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": False, "data": "synthetic",
            "config": {"workload": "dropin: %d families x %d genomes, K=3, beta=%.1f, ncem seq bern pk sk_, "
                                   "text files .str/.dat/.nei/.m in, .uf/.mf out" % (n, d, beta),
                       "dat_bytes": os.path.getsize(base + ".dat"), "nei_bytes": os.path.getsize(base + ".nei")},
            "first_call_s": times[0], "warm_calls_s": warm, "host_threads": os.cpu_count()}
    if nemo.have_ref() and not args.no_cpu:
        os.rename(base + ".uf", base + ".ours.uf")
        t0 = time.time()
        rc, _, _ = nemo.run_ref_cli(base, beta=beta, dolog=0)
        t_ref = time.time() - t0
        line["reference"] = {"seconds": t_ref, "rc": rc, "cores": 1,
                             "what": "oracle/_ref/nem_ref_cli (unmodified reference nem()), same files and arguments"}
        if rc == 0 and os.path.exists(base + ".uf"):
            ref = synth.read_uf(base + ".uf", 3)
            line["reference"]["labels_differ"] = int((ref.argmax(axis=1) != uf.argmax(axis=1)).sum())
    print(json.dumps(line))


def workload_name(w):
    n, d, beta, graph = WORKLOADS[w]
    return f"{w}: {n} families x {d} genomes, K={K}, beta={beta}, graph={graph}, ncem seq bern pk sk_"


# ----------------------------------------------------------------------------- config 5
def c5_plan(d, runs, seed=42):
    """The 1024 runs: genome-subset masks (sizes 100..900 step 100) and betas 0, 0.1 .. 1."""
    rng = np.random.default_rng(seed)
    wm = (d + 31) // 32
    masks = np.zeros((runs, wm), dtype=np.uint32)
    betas = np.zeros(runs, dtype=np.float32)
    sizes = np.zeros(runs, dtype=np.int32)
    for r in range(runs):
        size = 100 * (1 + r % 9)
        sel = rng.choice(d, size=min(size, d), replace=False)
        np.bitwise_or.at(masks[r], sel >> 5, (np.uint32(1) << (sel & 31).astype(np.uint32)))
        betas[r] = 0.1 * ((r // 9) % 11)
        sizes[r] = size
    return masks, betas, sizes


def main_c5(args):
    """Resample driver: the pangenome is resident on every GPU, rank r
    takes the runs r, r+world, ...; no data-path communication (replicas only); one all-reduce
    of the vote table at the end of a step."""
    import torch
    import torch.distributed as dist
    from pangenomenem_b200 import capi, synth_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()                   # settle the communicator first (see main_ours)
        torch.cuda.synchronize()
    n, d, beta, graph = WORKLOADS["c5"]
    if args.rows:
        n = args.rows
    runs = args.runs or C5_RUNS
    xdev, _ = synth_gpu.make_packed_on_device(n, d, seed=42, device=dev)      # same pangenome on every rank
    xh = xdev.cpu().numpy()
    row_ptr, col, wgt = synth_gpu.make_graph(n, xh, seed=42, kind=graph)
    masks, betas, sizes = c5_plan(d, runs)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if local_world * args.workers * 2 > (os.cpu_count() or 1):
        os.environ.setdefault("NEM_B200_POLL", "relaxed")   # ranks x workers pollers share the host cores
    mine = np.arange(rank, runs, world)
    eng = capi.Engine(local)
    eng.load_packed(xh.view(np.uint32), d, row_ptr, col, wgt)
    opts = dict(k=K, algo="ncem", update="seq", conv="clas", conv_thr=1e-8, it_max=100, prop="pk", disp="sk_")
    m_mine, b_mine = np.ascontiguousarray(masks[mine]), np.ascontiguousarray(betas[mine])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        votes, iters, st = eng.resample_batch(m_mine, b_mine, n_workers=args.workers, **opts)
        vt = torch.from_numpy(votes).to(dev)
        if world > 1:
            dist.all_reduce(vt)
        return vt, iters, st

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fam_it = launches = 0
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        vt, iters, st = step()
        fam_it += st.family_iterations
        launches += st.kernel_launches
    barrier()
    wall = time.time() - t0
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([wall], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(fam_it), float(launches), float(st.n_ok), float(st.n_inconsistent),
                        float(st.n_failed)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        wall_max = float(tt[0])
        fam_all, launches_all, n_ok, n_inc, n_fail = [float(v) for v in tot.tolist()]
        v = vt.cpu().numpy()
        decided = v[:, :3].argmax(axis=1)
        line = {
            "metric": "NEM family-iterations/s", "value": fam_all / wall_max, "unit": "family-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_max * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 popcount + f64 log-domain posteriors", "data": "synthetic",
            "config": {"workload": f"c5: {runs} independent fits on genome subsamples (100..900 of {d} genomes) of a "
                                   f"{n}-family pangenome, beta swept 0..1, K={K}, ncem seq bern pk sk_",
                       "families": n, "genomes": d, "runs": runs, "workers_per_gpu": args.workers,
                       "multi_gpu": "replicas only: runs dealt round-robin to the ranks, one all-reduce of the vote table per step",
                       "fits_ok": int(n_ok), "fits_inconsistent": int(n_inc), "fits_empty_class": int(n_fail),
                       "em_iterations_mean": float(np.mean(iters)),
                       "vote_summary": {"persistent": int((decided == 0).sum()), "shell": int((decided == 1).sum()),
                                        "cloud": int((decided == 2).sum())},
                       "timing": "wall clock between device synchronisations (the batch runs on worker streams)",
                       "l2": "every run re-reads the 32 MB pangenome through fresh subsample buffers; not flushed"},
            "clocks": clocks, "gpu_launches": int(launches_all),
            "e2e": {"value": fam_all / wall_max, "unit": "family-iterations/s",
                    "h2d_bytes_per_step": int(m_mine.nbytes + b_mine.nbytes) * world,
                    "d2h_bytes_per_step": int(n * 16) * world,
                    "what": "nemb_resample_batch: host masks in, vote table out (already the user-facing call)"},
            "roofline": None,
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------- extra measurements
def timed_fits(torch, eng, theta0, opts, steps, warmup, stream, flush=None):
    """W untimed fits then K fits between CUDA events on the engine's stream; returns
    (device ms, EM iterations, kernel launches, last Fit)."""
    for _ in range(warmup):
        eng.fit(*theta0, **opts)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    fits = []
    for s in range(steps):
        if flush is not None:
            flush.fill_(s & 0xff)
        ev[s][0].record(stream)
        fits.append(eng.fit(*theta0, **opts))
        ev[s][1].record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    return ms, sum(f.iters for f in fits), sum(f.kernel_launches for f in fits), fits[-1]


def small_workload_line(torch, capi, synth, synth_gpu, dev, local, name, steps=10, warmup=3):
    """value + e2e of one of the L2-sized BASELINE configs (C1, C2, C3) on one GPU."""
    n, d, beta, graph = WORKLOADS[name]
    xdev, _ = synth_gpu.make_packed_on_device(n, d, seed=42, device=dev)
    xh = xdev.cpu().numpy()
    if beta != 0 and graph != "none":
        row_ptr, col, wgt = synth_gpu.make_graph(n, xh, seed=42, kind=graph)
    else:
        row_ptr = col = wgt = None
    theta0 = synth.default_theta(K, d)
    opts = dict(k=K, algo="ncem", update="seq", beta=beta, conv="clas", conv_thr=1e-8, it_max=100,
                prop="pk", disp="sk_", sweep_impl="auto")
    stream = torch.cuda.current_stream()
    eng = capi.Engine(local)
    eng.set_stream(stream.cuda_stream)
    eng.load_shard_device(xdev.data_ptr(), n, 0, n, d, xdev.shape[1], row_ptr, col, wgt)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ms, iters, launches, f = timed_fits(torch, eng, theta0, opts, steps, warmup, stream, flush)
    lab = eng.labels()
    xhu = xh.view(np.uint32)
    eng.load_shard(xhu, n, 0, d, row_ptr, col, wgt); eng.fit(*theta0, **opts)      # first-touch pass
    lab_host = np.empty(n, dtype=np.int32)
    torch.cuda.synchronize()
    t0 = time.time(); e_iters = 0
    for _ in range(3):
        eng.load_shard(xhu, n, 0, d, row_ptr, col, wgt)
        fe = eng.fit(*theta0, **opts)
        eng.labels(0, n, out=lab_host)
        e_iters += fe.iters
    e_wall = time.time() - t0
    eng.close()
    del flush, xdev
    wb, tk = 4 * ((d + 31) // 32), 4 * K
    nnz = 0 if col is None else int(col.shape[0])
    b_iter = 2 * n * wb + 5 * n * tk + 8 * nnz + 4 * (n + 1) + 2 * 4 * K * d
    peak, _ = peaks()
    return {"workload": workload_name(name), "value": n * iters / (ms * 1e-3), "unit": "family-iterations/s",
            "ms_per_fit": ms / steps, "em_iterations_per_fit": f.iters, "kernel_launches_per_fit": launches / steps,
            "persistent_kernel": f.pk, "steps": steps, "warmup": warmup,
            "iteration_frac_of_hbm_roofline": b_iter * iters / (ms * 1e-3) / 1e9 / peak,
            "e2e": {"value": n * e_iters / e_wall, "unit": "family-iterations/s", "ms_per_step": e_wall * 1e3 / 3,
                    "labels_equal_resident_fit": bool(np.array_equal(lab_host, lab))},
            "l2": "X fits the L2; 512 MB flush between timed fits"}

# ----------------------------------------------------------------------------- our arm
def build_global_graph(torch, dist, dev, rank, world, n_loc, xh, seed, kind):
    """Weak-scaling pangenome of world x n_loc families: every rank's shard carries its own
    pangenome-like graph (co-presence weights from its own rows); consecutive shards are joined
    like consecutive replicons -- a chain link plus 2000 seeded chords per boundary, weight =
    min(popcount_i, popcount_j).  Every rank assembles the same global CSR (on its GPU)."""
    from pangenomenem_b200 import synth
    rng = np.random.default_rng(seed + 1 + rank)
    edges = synth.pangenome_edges(n_loc, rng, kind)
    wts = synth.copresence(xh.view(np.uint32), edges) if edges.shape[0] else np.zeros(0, np.float32)
    if world == 1:
        return synth.edges_to_csr(n_loc, edges, wts)
    pop = torch.from_numpy(synth._POP8[xh.view(np.uint8)].sum(axis=1, dtype=np.int64).astype(np.int32)).to(dev)
    m = torch.tensor([edges.shape[0]], device=dev)
    ms = [torch.zeros_like(m) for _ in range(world)]
    dist.all_gather(ms, m)
    mmax = int(max(int(v) for v in ms))
    e_pad = torch.zeros((mmax, 2), dtype=torch.int64, device=dev)
    w_pad = torch.zeros(mmax, dtype=torch.float32, device=dev)
    e_pad[:edges.shape[0]] = torch.from_numpy(edges).to(dev)
    w_pad[:edges.shape[0]] = torch.from_numpy(wts).to(dev)
    e_all = [torch.zeros_like(e_pad) for _ in range(world)]
    w_all = [torch.zeros_like(w_pad) for _ in range(world)]
    p_all = [torch.zeros_like(pop) for _ in range(world)]
    dist.all_gather(e_all, e_pad); dist.all_gather(w_all, w_pad); dist.all_gather(p_all, pop)
    pop_g = torch.cat(p_all).to(torch.float32)
    src, dst, w = [], [], []
    for r in range(world):
        k = int(ms[r])
        src.append(e_all[r][:k, 0] + r * n_loc); dst.append(e_all[r][:k, 1] + r * n_loc)
        w.append(w_all[r][:k])
    g = torch.Generator(device="cpu"); g.manual_seed(seed + 777)
    for r in range(world - 1):
        a = torch.randint(r * n_loc, (r + 1) * n_loc, (2000,), generator=g)
        bb = torch.randint((r + 1) * n_loc, (r + 2) * n_loc, (2000,), generator=g)
        a[0], bb[0] = (r + 1) * n_loc - 1, (r + 1) * n_loc          # the chain link
        a, bb = a.to(dev), bb.to(dev)
        src.append(a); dst.append(bb)
        w.append(torch.clamp(torch.minimum(pop_g[a], pop_g[bb]), min=1.0))
    src, dst, w = torch.cat(src), torch.cat(dst), torch.cat(w)
    n_glob = world * n_loc
    key = torch.cat([src * n_glob + dst, dst * n_glob + src])
    ww = torch.cat([w, w])
    key, order = torch.sort(key)
    keep = torch.ones_like(key, dtype=torch.bool); keep[1:] = key[1:] != key[:-1]   # unique
    key, ww = key[keep], ww[order][keep]
    rows = key // n_glob
    col = (key % n_glob).to(torch.int32)
    row_ptr = torch.zeros(n_glob + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n_glob), 0)
    return (row_ptr.to(torch.int32).cpu().numpy(), col.cpu().numpy(), ww.cpu().numpy())


def main_ours(args):
    import torch
    import torch.distributed as dist
    from pangenomenem_b200 import capi, sharded, synth, synth_gpu

    def trace(msg):      # NEM_BENCH_TRACE=1: where a multi-rank run is (stderr, with the device drained)
        if os.environ.get("NEM_BENCH_TRACE"):
            torch.cuda.synchronize()
            print("[bench rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host staging buffers must sit on the socket next to this rank's GPU (see the helper)
    numa = sharded.bind_to_gpu_numa(local) if not os.environ.get("NEM_BENCH_NO_NUMA") else {"bound": False, "off": True}
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner/diagnostics go to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        # settle the communicator (connections, peer mappings) before anything else touches the GPUs:
        # at 8 ranks the first collective racing the data generation below died with an illegal
        # memory access inside NCCL on this pool (three runs of three; never with these two lines)
        dist.barrier()
        torch.cuda.synchronize()
    mode = args.mode if world > 1 else "single"

    n, d, beta, graph = WORKLOADS[args.workload]
    if args.rows:
        n = args.rows
    wb = 4 * ((d + 31) // 32)            # unpadded packed row bytes
    tk = 4 * K
    n_glob = n * world if mode == "sharded" else n      # weak scaling: n families PER GPU

    # ---- synthetic pangenome: X drawn + packed on the device, graph on the host
    t0 = time.time()
    trace("process group up")
    xdev, _ = synth_gpu.make_packed_on_device(n, d, seed=42 + rank, device=dev)
    torch.cuda.synchronize()
    trace("X drawn")
    wpr = xdev.shape[1]
    xhost = torch.empty(xdev.shape, dtype=torch.int32, pin_memory=True)
    xhost.copy_(xdev)
    torch.cuda.synchronize()
    xh = xhost.numpy()

    def pinned(a):      # e2e inputs live in pinned host memory (bench contract)
        if a is None:
            return None
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    if beta != 0 and graph != "none":
        if mode == "sharded":
            row_ptr, col, wgt = build_global_graph(torch, dist, dev, rank, world, n, xh, 42, graph)
        else:
            row_ptr, col, wgt = synth_gpu.make_graph(n, xh, seed=42 + rank, kind=graph)
    else:
        row_ptr = col = wgt = None
    trace("graph built")
    if os.environ.get("NEM_BENCH_TRACE") == "graph":
        return
    row_ptr, col, wgt = pinned(row_ptr), pinned(col), pinned(wgt)
    gen_s = time.time() - t0
    nnz = 0 if col is None else int(col.shape[0])
    theta0 = synth.default_theta(K, d)
    opts = dict(k=K, algo=args.algo, update=args.update, beta=beta, conv="clas", conv_thr=1e-8,
                it_max=100, prop="pk", disp=args.disp, sweep_impl="auto")
    off_default = (args.algo, args.update, args.disp) != ("ncem", "seq", "sk_")
    if off_default:
        args.no_extras = True

    stream = torch.cuda.current_stream()
    comm = None
    if mode == "sharded":
        eng, comm = sharded.make_engine(dist, local)
        plan = sharded.plan(n_glob, world, rank)
        assert plan.n_loc == n and plan.row0 == rank * n
    else:
        eng = capi.Engine(local)
        plan = sharded.plan(n, 1, 0)
    eng.set_stream(stream.cuda_stream)
    trace("engine + communicator up")
    eng.load_shard_device(xdev.data_ptr(), n_glob, plan.row0, n, d, wpr, row_ptr, col, wgt)
    trace("shard loaded")
    x_bytes = n * wpr * 4
    flush = None
    if x_bytes < 256 << 20:              # X fits the 126 MB L2: flush between timed steps
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_fit(profile):
        return eng.fit(*theta0, profile=profile, **opts)

    for _ in range(args.warmup):
        one_fit(False)
        trace("warm-up fit done")
    lab_warm = eng.labels()              # the timed fits below must reproduce these labels
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    fits = []
    barrier()
    w0 = time.time()
    timed = []
    for s in range(args.steps):
        if flush is not None:
            flush.fill_(s & 0xff)
        ev[s][0].record(stream)
        timed.append(one_fit(False))
        ev[s][1].record(stream)
    barrier()
    wall = time.time() - w0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    iters = sum(f.iters for f in timed)
    launches = sum(f.kernel_launches for f in timed)
    lab_resident = eng.labels()          # untimed: the e2e path below must reproduce these labels
    # per-stage CUDA-event timings (roofline): separate, untimed-for-`value` fits with profile=1
    for s in range(min(args.steps, 3)):
        if flush is not None:
            flush.fill_(s & 0xff)
        fits.append(one_fit(True))
    pk_prof = eng.persist_profile() if (fits and fits[-1].pk and fits[-1].pk.get("launches")) else None
    pk_trace = eng.persist_trace().tolist() if pk_prof else None
    barrier()

    # ---- e2e through the C ABI with host buffers (pinned X), H2D + graph upload/validation +
    #      fit + D2H of the labels inside the timed region; same handle => buffers are reused
    #      like a long-lived caller (ppanggolin's chunk loop, ppanggolin.py:1045-1095) would
    e2e_steps = max(1, min(args.steps, 3))
    xhu = xh.view(np.uint32)
    for _ in range(1):                   # one untimed pass: first-touch allocations
        eng.load_shard(xhu, n_glob, plan.row0, d, row_ptr, col, wgt)
        eng.fit(*theta0, **opts)
    # every rank reads back the labels of ITS OWN families (all of them on one GPU), into a buffer
    # it keeps across steps like the long-lived caller above
    lab_host = np.empty(plan.n_loc, dtype=np.int32)
    barrier()
    e0 = time.time()
    e_iters = 0
    t_load = t_fit = 0.0
    for _ in range(e2e_steps):
        ta = time.time()
        eng.load_shard(xhu, n_glob, plan.row0, d, row_ptr, col, wgt)
        tb = time.time()
        f = eng.fit(*theta0, **opts)
        eng.labels(plan.row0, plan.n_loc, out=lab_host)
        t_load += tb - ta; t_fit += time.time() - tb
        e_iters += f.iters
    barrier()
    e_wall = time.time() - e0
    e2e_same = bool(np.array_equal(lab_host, lab_resident[plan.rows]))
    # ---- full-size property check (untimed): a converged ncem fit is a fixed point of NemAlgo --
    #      restarted from its own partition (nemb_fit_from_partition, the reference's INIT_FILE
    #      branch) it must stop after ONE iteration with every label unchanged
    props = None
    if world == 1:
        f2 = eng.fit(*theta0, t_init=eng.posteriors(K), **opts)
        lab2 = eng.labels()
        props = {"refit_from_final_partition": {"iters": f2.iters, "converged": bool(f2.converged),
                                                "labels_unchanged": bool(np.array_equal(lab2, lab_resident))},
                 "labels_repeatable_across_fits": bool(np.array_equal(lab_warm, lab_resident)),
                 "class_sizes": np.bincount(lab_resident, minlength=K).tolist(),
                 "class_sizes_sum_to_n": bool(np.bincount(lab_resident, minlength=K).sum() == n)}
    # ---- extra lines of the default single-GPU run: the exact shortcuts switched off,
    #      a pangenome whose shell centre keeps moving, and the three L2-sized BASELINE configs
    extras = {}
    if world == 1 and not args.rows and not args.no_extras:
        b_iter_x = 2 * n * wb + 5 * n * tk + 8 * nnz + 4 * (n_glob + 1) + 2 * 4 * K * d
        peak_x, _ = peaks()
        os.environ["NEM_B200_NO_SHORTCUTS"] = "1"
        ms_ns, it_ns, ln_ns, f_ns = timed_fits(torch, eng, theta0, opts, 5, 2, stream, flush)
        del os.environ["NEM_B200_NO_SHORTCUTS"]
        same_ns = bool(np.array_equal(eng.labels(), lab_resident))
        extras["value_no_shortcuts"] = {
            "value": n * it_ns / (ms_ns * 1e-3), "unit": "family-iterations/s", "ms_per_fit": ms_ns / 5,
            "em_iterations_per_fit": f_ns.iters, "x_passes_per_fit": f_ns.pk.get("x_passes", 0) if f_ns.pk else None,
            "xt_recounts_per_fit": f_ns.pk.get("recounts", 0) if f_ns.pk else None,
            "site_evaluations_saved_per_fit": f_ns.n_kept, "labels_equal_default_fit": same_ns,
            "iteration_frac_of_hbm_roofline": b_iter_x * it_ns / (ms_ns * 1e-3) / 1e9 / peak_x,
            "what": "NEM_B200_NO_SHORTCUTS=1: every EM iteration reads X (Hamming counts recomputed), recounts "
                    "S = X^T T through the transposed bits and evaluates every site (margin cache off); same "
                    "pangenome, same labels; the fraction is (iterations x the per-iteration algorithmic bytes of "
                    "SURVEY 8d) / time / measured HBM peak"}
        if args.workload == "c4":
            xmv, _ = synth_gpu.make_packed_on_device(n, d, seed=43, device=dev, shell="moving")
            xmh = xmv.cpu().numpy()
            if beta != 0 and graph != "none":
                rp2, c2, w2 = synth_gpu.make_graph(n, xmh, seed=43, kind=graph)
            else:
                rp2 = c2 = w2 = None
            em = capi.Engine(local)
            em.set_stream(stream.cuda_stream)
            em.load_shard_device(xmv.data_ptr(), n, 0, n, d, xmv.shape[1], rp2, c2, w2)
            ms_mv, it_mv, ln_mv, f_mv = timed_fits(torch, em, theta0, opts, 5, 2, stream, None)
            nnz2 = 0 if c2 is None else int(c2.shape[0])
            b_iter_m = 2 * n * wb + 5 * n * tk + 8 * nnz2 + 4 * (n + 1) + 2 * 4 * K * d
            extras["moving_centres"] = {
                "value": n * it_mv / (ms_mv * 1e-3), "unit": "family-iterations/s", "ms_per_fit": ms_mv / 5,
                "em_iterations_per_fit": f_mv.iters, "converged": bool(f_mv.converged),
                "x_passes_per_fit": f_mv.pk.get("x_passes", 0) if f_mv.pk else None,
                "site_evaluations_saved_per_fit": f_mv.n_kept,
                "iteration_frac_of_hbm_roofline": b_iter_m * it_mv / (ms_mv * 1e-3) / 1e9 / peak_x,
                "what": "default engine (all shortcuts on) on a pangenome generated so that they pay least: every "
                        "shell family has presence probability 1/2, the shell centre's majority votes sit on the "
                        "boundary and flip from iteration to iteration (synth_gpu shell='moving', seed 43)"}
            em.close()
            del xmv
            extras["other_workloads"] = {w: small_workload_line(torch, capi, synth, synth_gpu, dev, local, w)
                                         for w in ("c1", "c2", "c3")}

    # ---- N > 1: BASELINE config 4 as written -- ONE pangenome of `n` families in total, row-sharded
    #      over the N GPUs (strong scaling), checked on rank 0 against the same fit on one GPU
    strong = None
    if mode == "sharded" and world > 1 and not args.no_extras:
        xfull, _ = synth_gpu.make_packed_on_device(n, d, seed=4242, device=dev)       # same on every rank
        xfh = xfull.cpu().numpy()
        if beta != 0 and graph != "none":
            rps, cs, ws = synth_gpu.make_graph(n, xfh, seed=4242, kind=graph)
        else:
            rps = cs = ws = None
        ps = sharded.plan(n, world, rank)
        es, comm_s = sharded.make_engine(dist, local)
        es.set_stream(stream.cuda_stream)
        xs = xfull[ps.row0:ps.row0 + ps.n_loc].contiguous()
        es.load_shard_device(xs.data_ptr(), n, ps.row0, ps.n_loc, d, xfull.shape[1], rps, cs, ws)
        for _ in range(max(2, args.warmup)):
            es.fit(*theta0, **opts)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        s_it = 0
        for q in range(args.steps):
            evs[q][0].record(stream)
            fs = es.fit(*theta0, **opts)
            evs[q][1].record(stream)
            s_it += fs.iters
        barrier()
        s_ms = sum(a_.elapsed_time(b_) for a_, b_ in evs)
        lab_s = es.labels()
        t_s = torch.tensor([s_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        es.close()
        capi.comm_destroy(comm_s)
        if rank == 0:
            one = capi.Engine(local)
            one.set_stream(stream.cuda_stream)
            one.load_shard_device(xfull.data_ptr(), n, 0, n, d, xfull.shape[1], rps, cs, ws)
            ms1, it1, _, f1 = timed_fits(torch, one, theta0, opts, min(args.steps, 5), 2, stream, None)
            lab1 = one.labels()
            one.close()
            strong = {"families_total": n, "families_per_gpu": ps.shard_len, "value": n * s_it / (float(t_s[0]) * 1e-3),
                      "unit": "family-iterations/s", "ms_per_step": float(t_s[0]) / args.steps,
                      "em_iterations_per_fit": fs.iters, "cut_edges": sharded.cut_edges(rps, cs, world) if rps is not None else 0,
                      "single_gpu_ms_per_step": ms1 / min(args.steps, 5),
                      "single_gpu_value": n * it1 / (ms1 * 1e-3),
                      "identical_to_single_gpu_fit": bool(fs.iters == f1.iters and np.array_equal(lab_s, lab1)
                                                          and np.array_equal(fs.center, f1.center)
                                                          and np.array_equal(fs.disp, f1.disp)),
                      "what": "BASELINE config 4 as written: ONE 1M-family x 5000-genome pangenome (seed 4242) split "
                              "into contiguous row ranges over the GPUs; rank 0 also fits the whole pangenome on its "
                              "own GPU and compares labels, theta and iteration count (real NCCL run)"}
        del xfull
    depth = eng.dims()["depth"] if mode != "sharded" else 0      # builds the level schedule: untimed
    h2d = x_bytes + (0 if col is None else (n_glob + 1) * 4 + nnz * 8) + (K + 2 * K * d) * 4
    d2h = plan.n_loc + (K + 2 * K * d) * 4 + 256

    # ---- max over ranks
    t_dev = torch.tensor([dev_ms, e_wall * 1e3, wall * 1e3], dtype=torch.float64, device=dev)
    per_rank_units = float(n * iters) if mode != "single" or world == 1 else float(n * iters)
    tot = torch.tensor([per_rank_units, float(n * e_iters), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e_ms_max, wall_ms_max = [float(v) for v in t_dev.tolist()]
    fam_iters, e_fam_iters, launches_all = [float(v) for v in tot.tolist()]

    if rank == 0:
        peak, peak_src = peaks()
        f = fits[-1]
        n_den = sum(x.stage_launches["density"] for x in fits)
        n_ms = sum(x.stage_launches["mstep"] for x in fits)
        ms_den = sum(x.stage_ms["density"] for x in fits) / max(n_den, 1)
        ms_ms = sum(x.stage_ms["mstep"] for x in fits) / max(n_ms, 1)
        ms_sw = sum(x.stage_ms["sweep"] for x in fits) / max(sum(x.stage_launches["sweep"] for x in fits), 1)
        n_dc = sum(x.stage_launches["density_cached"] for x in fits)
        ms_dc = sum(x.stage_ms["density_cached"] for x in fits) / max(n_dc, 1)
        n_md = sum(x.stage_launches["mstep_delta"] for x in fits)
        ms_md = sum(x.stage_ms["mstep_delta"] for x in fits) / max(n_md, 1)
        if args.disp in ("sk_", "s__"):
            traffic, traffic_src = measured_traffic("k_density_tma") if args.workload == "c4" and not args.rows else (None, None)
        else:
            traffic, traffic_src = measured_traffic("k_density_general_tiled") if args.workload == "c3" and not args.rows else (None, None)
        den_bytes = n * wb + n * tk                      # SURVEY 8d "E-step-only bytes (density)"
        achieved = den_bytes / (ms_den * 1e-3) / 1e9 if ms_den > 0 else 0.0
        ms_bytes = n * wb + n * tk                       # M-step: X once + t once
        b_iter = 2 * n * wb + 5 * n * tk + 8 * nnz + 4 * (n_glob + 1) + 2 * 4 * K * d
        sw_bytes = 3 * n * tk + 8 * nnz + 4 * n
        iter_ms = dev_ms / max(iters, 1)
        value = fam_iters / (dev_ms_max * 1e-3)
        cpu = cpu_baseline(args.workload, 1) if (world == 1 and not args.no_cpu) else None
        multi = {"single": "single GPU",
                 "sharded": ("ONE pangenome of %d families row-sharded over %d GPUs (%d per GPU): X sharded, "
                             "graph/theta/labels replicated; " % (n_glob, world, n)) +
                            ("moved labels, re-evaluation requests and the M-step statistics travel through NVLink "
                             "peer memory inside the persistent kernel (cross-rank epoch barriers, no collective call)"
                             if (f.pk and f.pk.get("launches")) else
                             "per iteration one all-gather of the M-step statistics and label all-gathers inside "
                             "the exact sequential sweep (NCCL; the peer-memory kernel serves up to 4 ranks)"),
                 "replicas": "independent replicas, one per GPU, no communication"}[mode]
        line = {
            "metric": "NEM family-iterations/s", "value": value, "unit": "family-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 popcount + f64 log-domain posteriors",
            "data": "synthetic",
            "config": {"workload": (workload_name(args.workload) if n_glob == n else
                                    workload_name(args.workload).replace("%d families" % n, "%d families (%d per GPU, weak scaling)" % (n_glob, n))
                                    ).replace("ncem seq bern pk sk_", "%s %s bern pk %s" % (args.algo, args.update, args.disp)),
                       "families": n_glob,
                       "families_per_gpu": n, "genomes": d, "K": K,
                       "beta": beta, "algo": args.algo, "update": args.update, "dispersion": args.disp, "nnz": nnz,
                       "sweep_dag_depth": depth,
                       "em_iterations_per_fit": f.iters, "converged": f.converged,
                       "fixup_rounds_per_fit": f.fixup_rounds,
                       "site_evaluations_saved_per_fit": f.n_kept,
                       "multi_gpu": multi,
                       "l2": ("inputs larger than L2 (X = %d MB)" % (x_bytes >> 20)) if flush is None
                       else "L2 flushed (512 MB write) between timed steps",
                       "synth_seconds": round(gen_s, 1)},
            "clocks": clocks,
            "gpu_launches": int(launches_all),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": e_fam_iters / (e_ms_max * 1e-3), "unit": "family-iterations/s",
                    "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                    "h2d_bytes_per_rank": int(h2d), "d2h_bytes_per_rank": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e_ms_max / e2e_steps,
                    "load_ms": t_load * 1e3 / e2e_steps, "fit_ms": t_fit * 1e3 / e2e_steps,
                    "labels_equal_resident_fit": e2e_same,
                    "host_numa_binding": numa,
                    "what": "per rank: nemb_load_shard(host pinned X shard + global CSR: H2D, device-side graph validation) + nemb_fit + nemb_get_labels_rows(own families); *_per_step = summed over the ranks"},
            "roofline": {"bound": "hbm", "kernel": "k_density_tma (E-step Bernoulli log-likelihood, popcount path)"
                         if args.disp in ("sk_", "s__") else
                         "k_density_general_tiled (E-step Bernoulli log-likelihood, per-genome dispersions; bound by the "
                         "fp64 pipe -- N*D*K fused multiply-adds per pass -- not by HBM: see fp64_tflops)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "bytes_per_launch": den_bytes, "avg_launch_ms": ms_den,
                         "launches_timed": n_den,
                         "fp64_tflops": (2.0 * n * d * K / (ms_den * 1e-3) / 1e12 if ms_den > 0 else 0.0)
                         if args.disp not in ("sk_", "s__") else None,
                         "note": ("achieved = algorithmic bytes of one X pass (N*Wb + N*K*4) / mean CUDA-event "
                                  "time of the launches that READ X; the iterations whose class centres did not "
                                  "move skip the pass and rebuild logpf from cached Hamming counts "
                                  "(density_cached below), and are not averaged in"),
                         "density_cached": {"launches": n_dc, "avg_ms": ms_dc},
                         "mstep_delta": {"launches": n_md, "avg_ms": ms_md,
                                         "what": "incremental S/n update from the rows that changed class"},
                         "mstep": {"avg_ms": ms_ms, "achieved": ms_bytes / (ms_ms * 1e-3) / 1e9 if ms_ms > 0 else 0.0,
                                   "launches": n_ms,
                                   "what": "full recount: label masks + X^T popcount + closed forms + tables, X^T read once"},
                         "sweep_avg_ms": ms_sw,
                         "sweep": {"bound": "latency (dependent gathers + fix-up rounds of the exact sequential sweep)",
                                   "bytes_per_sweep": sw_bytes, "avg_ms": ms_sw,
                                   "achieved": sw_bytes / (ms_sw * 1e-3) / 1e9 if ms_sw > 0 else 0.0,
                                   "frac": (sw_bytes / (ms_sw * 1e-3) / 1e9 / peak) if ms_sw > 0 else 0.0,
                                   "what": "the other E-step kernel group (dense round + fix-up rounds), reported against "
                                           "ITS algorithmic bytes 3*N*T + 8*nnz + 4*N (SURVEY 8d) although it is latency-bound"},
                         "iteration": {"algorithmic_bytes": b_iter, "avg_ms": iter_ms,
                                       "achieved": b_iter / (iter_ms * 1e-3) / 1e9 if iter_ms > 0 else 0.0,
                                       "note": "whole fit time / EM iterations (includes init sweeps, host syncs)"}},
        }
        line["persistent_kernel"] = {"launches_per_fit": f.pk.get("launches"), "device_barriers_per_fit": f.pk.get("barriers"),
                                     "x_passes_inside": f.pk.get("x_passes"), "what": "csrc/nem_persist.cuh: one cooperative "
                                     "launch runs the EM loop; a problem larger than the L2 leaves it for the TMA "
                                     "density pass and the X^T recount"} if f.pk else None
        if pk_prof:
            sw_us = pk_prof["margin_test"] + pk_prof["eval_list"] + pk_prof["eval_dense"] + pk_prof["fixup"]
            n_sw = f.iters + 2
            line["roofline"]["sweep_avg_ms"] = sw_us / 1e3 / n_sw
            line["roofline"]["sweep"].update({"avg_ms": sw_us / 1e3 / n_sw,
                                              "achieved": sw_bytes / (sw_us / n_sw * 1e-6) / 1e9,
                                              "frac": sw_bytes / (sw_us / n_sw * 1e-6) / 1e9 / peak,
                                              "what": "phases of the persistent kernel that make up a sweep (margin test + "
                                                      "evaluation + fix-up rounds, barrier waits included), averaged over the "
                                                      "sweeps of a fit, against ITS algorithmic bytes 3*N*T + 8*nnz + 4*N"})
            line["persistent_kernel"]["phase_us_per_fit"] = {k: round(v, 1) for k, v in pk_prof.items()}
            line["persistent_kernel"]["per_iteration_us"] = {
                "columns": ["scan", "delta_or_recount", "closed_forms", "margin_test", "evaluation", "fixup", "sites_evaluated", "fixup_rounds"],
                "rows_last_launch": [[round(v, 1) for v in r] for r in pk_trace if sum(r[:6]) > 0]}
        line.update(extras)
        if strong is not None:
            line["strong_scaling"] = strong
            line["scaling_note"] = ("`value`/`scaling` = weak (%d families per GPU, %d in total); strong_scaling = "
                                    "BASELINE config 4 as written (%d families in total)" % (n, n_glob, n))
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if props is not None:
            line["full_size_properties"] = props
        print(json.dumps(line))
    eng.close()
    if comm:
        capi.comm_destroy(comm)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["dropin"])
    ap.add_argument("--rows", type=int, default=0, help="override the number of families (debug)")
    ap.add_argument("--mode", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1: one row-sharded pangenome of N x families (default) or N independent replicas")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip value_no_shortcuts / moving_centres / other_workloads / strong_scaling")
    ap.add_argument("--algo", default="ncem", choices=["ncem", "nem"],
                    help="off-default line: nem = fuzzy posteriors (k_sweep_nem_*, k_mstep_nem_*)")
    ap.add_argument("--update", default="seq", choices=["seq", "para"])
    ap.add_argument("--disp", default="sk_", choices=["s__", "sk_", "s_d", "skd"],
                    help="off-default line: skd / s_d = per-genome dispersions (k_density_general; PPanGGOLiN -fd)")
    ap.add_argument("--runs", type=int, default=0, help="c5: number of independent fits (default 1024)")
    ap.add_argument("--workers", type=int, default=8, help="c5: worker streams per GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        main_reference(args)
    elif args.workload == "dropin":
        main_dropin(args)
    elif args.workload == "c5":
        main_c5(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
