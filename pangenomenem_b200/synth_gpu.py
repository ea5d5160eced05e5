"""Large synthetic pangenomes generated on the GPU with torch (bench plumbing, not the product).

Same recipe as :mod:`pangenomenem_b200.synth` (40 % persistent p=.97 / 20 % shell p~U(.2,.8) /
40 % cloud p=.03, spatially correlated classes, pangenome-like graph, co-presence weights) but
the N x D Bernoulli draws and the bit packing run on the device so that the 1M x 5000
configuration is built in seconds.  The graph is built on the host with numpy."""
from __future__ import annotations

import numpy as np
import torch

from . import synth


def make_packed_on_device(n: int, d: int, seed: int, device: torch.device, shell: str = "default"):
    """Returns (x_packed int32 [n, wpr] on `device`, latent int8 numpy [n]).

    shell="moving": every shell family has presence probability exactly 1/2, so the shell class's
    per-genome majority votes sit on the boundary and its centre keeps flipping genomes from one EM
    iteration to the next (the worst case of the cached-Hamming-count shortcut: an X pass per
    iteration)."""
    rng = np.random.default_rng(seed)
    latent = synth.latent_classes(n, rng)
    q = np.empty(n, dtype=np.float32)
    q[latent == 0] = 0.97
    sh = latent == 1
    q[sh] = rng.uniform(0.2, 0.8, size=int(sh.sum())).astype(np.float32)
    if shell == "moving":
        q[sh] = 0.5
    q[latent == 2] = 0.03
    w = (d + 31) // 32
    wpr = (w + 3) // 4 * 4
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    qd = torch.from_numpy(q).to(device)
    out = torch.zeros((n, wpr), dtype=torch.int32, device=device)
    pow2 = (torch.ones(32, dtype=torch.int64, device=device) << torch.arange(32, device=device))
    chunk = max(1, (1 << 27) // max(d, 1))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        bits = torch.rand((hi - lo, d), generator=g, device=device) < qd[lo:hi, None]
        empty = ~bits.any(dim=1)                      # PPanGGOLiN never emits empty families
        if bool(empty.any()):
            idx = empty.nonzero().squeeze(1)
            colr = torch.randint(0, d, (idx.numel(),), generator=g, device=device)
            bits[idx, colr] = True
        pad = torch.zeros((hi - lo, w * 32), dtype=torch.int64, device=device)
        pad[:, :d] = bits
        words = (pad.view(hi - lo, w, 32) * pow2).sum(dim=2)          # < 2^32
        words = torch.where(words >= (1 << 31), words - (1 << 32), words).to(torch.int32)
        out[lo:hi, :w] = words
        del bits, pad, words
    return out, latent


def make_graph(n: int, x_packed_host: np.ndarray, seed: int, kind: str = "pangenome",
               weighted: bool = True):
    rng = np.random.default_rng(seed + 1)
    edges = synth.pangenome_edges(n, rng, kind)
    if weighted and edges.shape[0]:
        wts = synth.copresence(x_packed_host.view(np.uint32), edges)
    else:
        wts = np.ones(edges.shape[0], dtype=np.float32)
    return synth.edges_to_csr(n, edges, wts)
