"""Seeded synthetic pangenomes and the NEM file contract.

Host-side helpers shared by the tests, the bench and the golden-vector generator:

* :func:`make_pangenome` -- presence/absence matrix X (families x genomes) plus a
  "pangenome-like" chromosomal-neighbour graph: a backbone cycle,
  island chains hanging off it, random chords up to mean degree ~4, symmetric, edge weight =
  co-presence count like PPanGGOLiN's coverage (reference ppanggolin.py:865-880).
* :func:`write_nem_files` -- writes ``<base>.str/.dat/.nei/.m/.index`` byte-for-byte the way
  ``PPanGGOLiN.__write_nem_input_files`` does (reference ppanggolin.py:829-930).
* :func:`read_uf` / :func:`read_mf` / :func:`classify_psc` -- parse the engine's outputs the way
  ``run_partitioning`` does (reference ppanggolin.py:1890-1972).

Nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint8)


@dataclass
class Pangenome:
    x: np.ndarray          # uint8 [N, D] in {0,1}
    row_ptr: np.ndarray    # int32 [N+1]
    col: np.ndarray        # int32 [nnz], 0-based neighbour ids, file order inside a row
    wgt: np.ndarray        # float32 [nnz]
    latent: np.ndarray     # int8 [N] generating class 0=persistent 1=shell 2=cloud

    @property
    def n(self) -> int:
        return int(self.x.shape[0])

    @property
    def d(self) -> int:
        return int(self.x.shape[1])


def pack_rows(x: np.ndarray, words_per_row: int | None = None) -> np.ndarray:
    """CPU bit-packer: genome d of family i -> bit (d % 32) of word d // 32 (LSB first).

    Rows are padded with zero words up to ``words_per_row`` (default: ceil(D/32) rounded up to
    a multiple of 4 so every row is 16-byte aligned -- the engine's HBM layout)."""
    n, d = x.shape
    w = (d + 31) // 32
    if words_per_row is None:
        words_per_row = (w + 3) // 4 * 4
    bits = np.zeros((n, words_per_row * 32), dtype=np.uint8)
    bits[:, :d] = x != 0
    packed = np.packbits(bits, axis=1, bitorder="little")
    return np.ascontiguousarray(packed).view("<u4").reshape(n, words_per_row)


def latent_classes(n: int, rng: np.random.Generator, stay: float = 0.9) -> np.ndarray:
    """Spatially correlated latent classes (40 % persistent / 20 % shell / 40 % cloud)."""
    fresh = rng.choice(3, size=n, p=[0.4, 0.2, 0.4]).astype(np.int8)
    keep = rng.random(n) < stay
    keep[0] = False
    # label of i = label drawn at the last index j <= i where keep[j] is False
    src = np.where(~keep, np.arange(n), 0)
    src = np.maximum.accumulate(src)
    return fresh[src]


def sample_x(latent: np.ndarray, d: int, rng: np.random.Generator) -> np.ndarray:
    n = latent.shape[0]
    q = np.empty(n, dtype=np.float32)
    q[latent == 0] = 0.97
    shell = latent == 1
    q[shell] = rng.uniform(0.2, 0.8, size=int(shell.sum())).astype(np.float32)
    q[latent == 2] = 0.03
    x = np.empty((n, d), dtype=np.uint8)
    step = max(1, (1 << 24) // max(d, 1))
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        x[lo:hi] = rng.random((hi - lo, d), dtype=np.float32) < q[lo:hi, None]
    empty = np.flatnonzero(x.sum(axis=1) == 0)  # PPanGGOLiN never emits empty families
    if empty.size:
        x[empty, rng.integers(0, d, size=empty.size)] = 1
    return x


def pangenome_edges(n: int, rng: np.random.Generator, kind: str = "pangenome",
                    mean_degree: float = 4.0) -> np.ndarray:
    """Undirected edge list [m,2] (i<j, unique, no self loops)."""
    if kind == "none" or n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    if kind == "chain":
        i = np.arange(n - 1)
        return np.stack([i, i + 1], axis=1)
    if kind == "random":
        m = int(n * mean_degree / 2)
        e = rng.integers(0, n, size=(m, 2))
    elif kind == "pangenome":
        b = max(2, min(4000, n // 5))
        parts = [np.stack([np.arange(b), (np.arange(b) + 1) % b], axis=1)]
        rest = n - b
        if rest > 0:
            lens = rng.geometric(0.2, size=rest)          # mean 5; more than enough chains
            ends = np.cumsum(lens)
            nch = int(np.searchsorted(ends, rest)) + 1
            ends = np.minimum(ends[:nch], rest)
            starts = np.concatenate([[0], ends[:-1]])
            ok = ends > starts
            starts, ends = starts[ok] + b, ends[ok] + b
            inner = np.ones(rest, dtype=bool)
            inner[ends - b - 1] = False                   # last of a chain has no successor
            ids = np.arange(b, n)[inner]
            parts.append(np.stack([ids, ids + 1], axis=1))
            anchor = rng.integers(0, b, size=starts.size)
            parts.append(np.stack([anchor, starts], axis=1))
            parts.append(np.stack([(anchor + 1) % b, ends - 1], axis=1))
        e = np.concatenate(parts)
        extra = int(max(0.0, n * mean_degree / 2 - e.shape[0]))
        if extra:
            a = rng.integers(0, n, size=extra)
            near = np.clip(a + rng.integers(-64, 65, size=extra), 0, n - 1)
            far = rng.integers(0, n, size=extra)
            e = np.concatenate([e, np.stack([a, np.where(rng.random(extra) < 0.5, near, far)], 1)])
    else:
        raise ValueError(f"unknown graph kind {kind!r}")
    e = e[e[:, 0] != e[:, 1]]
    e = np.sort(e, axis=1)
    return np.unique(e, axis=0)


def edges_to_csr(n: int, edges: np.ndarray, weights: np.ndarray):
    """Symmetric CSR, neighbours sorted by id inside a row."""
    src = np.concatenate([edges[:, 0], edges[:, 1]])
    dst = np.concatenate([edges[:, 1], edges[:, 0]])
    w = np.concatenate([weights, weights])
    order = np.lexsort((dst, src))
    src, dst, w = src[order], dst[order], w[order]
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return row_ptr.astype(np.int32), dst.astype(np.int32), w.astype(np.float32)


def copresence(xp: np.ndarray, edges: np.ndarray, chunk: int = 1 << 16) -> np.ndarray:
    """Edge weight = number of genomes holding both families (>= 1), from packed rows."""
    out = np.empty(edges.shape[0], dtype=np.float32)
    xb = xp.view(np.uint8)
    for lo in range(0, edges.shape[0], chunk):
        hi = min(edges.shape[0], lo + chunk)
        both = xb[edges[lo:hi, 0]] & xb[edges[lo:hi, 1]]
        out[lo:hi] = _POP8[both].sum(axis=1, dtype=np.int64)
    return np.maximum(out, 1.0)


def make_pangenome(n: int, d: int, seed: int = 42, graph: str = "pangenome",
                   weighted: bool = True, mean_degree: float = 4.0) -> Pangenome:
    rng = np.random.default_rng(seed)
    latent = latent_classes(n, rng)
    x = sample_x(latent, d, rng)
    edges = pangenome_edges(n, rng, graph, mean_degree)
    if weighted and edges.shape[0]:
        w = copresence(pack_rows(x), edges)
    else:
        w = np.ones(edges.shape[0], dtype=np.float32)
    row_ptr, col, wgt = edges_to_csr(n, edges, w)
    return Pangenome(x=x, row_ptr=row_ptr, col=col, wgt=wgt, latent=latent)


def sweep_dag_depth(row_ptr: np.ndarray, col: np.ndarray) -> int:
    """Depth of the Gauss-Seidel dependency DAG of an index-order sweep (SURVEY hard part #1):
    level[i] = 1 + max(level[j] for neighbours j < i)."""
    n = row_ptr.shape[0] - 1
    level = np.zeros(n, dtype=np.int32)
    for i in range(n):
        nb = col[row_ptr[i]:row_ptr[i + 1]]
        nb = nb[nb < i]
        level[i] = 1 + (level[nb].max() if nb.size else 0)
    return int(level.max()) if n else 0


# --------------------------------------------------------------------------- file contract

def default_theta(k: int, d: int, low_disp: float = 0.1):
    """PPanGGOLiN's default .m (ppanggolin.py:893-901) as the reference reads it
    (nem_exe.c:1022-1034: last proportion = 1 - sum of the others, in float)."""
    assert k == 3
    prop = np.array([0.33333, 0.33333, 0.0], dtype=np.float32)
    prop[2] = np.float32(np.float32(1.0) - prop[0]) - prop[1]
    center = np.repeat(np.array([1.0, 0.5, 0.0], dtype=np.float32)[:, None], d, axis=1)
    disp = np.repeat(np.array([low_disp, 0.5, low_disp], dtype=np.float32)[:, None], d, axis=1)
    return prop, center, disp


def default_m_line(d: int, low_disp: float = 0.1) -> str:
    """PPanGGOLiN's default ``.m`` (reference ppanggolin.py:893-901)."""
    return ("1 " + "0.33333 0.33333 " + " ".join(["1"] * d) + " " + " ".join(["0.5"] * d) + " "
            + " ".join(["0"] * d) + " " + " ".join([str(low_disp)] * d) + " "
            + " ".join(["0.5"] * d) + " " + " ".join([str(low_disp)] * d))


def m_line(flag: int, prop: np.ndarray, center: np.ndarray, disp: np.ndarray) -> str:
    """Generic ``.m``: flag, K-1 proportions, K*D centres, K*D dispersions (nem_exe.c:973-1091)."""
    vals = [str(flag)] + [repr(float(p)) for p in prop[:-1]]
    vals += [repr(float(v)) for v in np.asarray(center).ravel()]
    vals += [repr(float(v)) for v in np.asarray(disp).ravel()]
    return " ".join(vals)


def _fmt_w(w: float) -> str:
    r = round(float(w), 4)                      # ppanggolin.py:880 str(round(score, 4))
    return str(int(r)) if r == int(r) else str(r)


def write_nem_files(base: str, pg: Pangenome, m_text: str | None = None,
                    spatial: bool = True, weighted_flag: int = 1) -> None:
    """Write ``base.str/.dat/.nei/.m/.index`` like ppanggolin.py:829-930."""
    os.makedirs(os.path.dirname(os.path.abspath(base)), exist_ok=True)
    n, d = pg.n, pg.d
    with open(base + ".str", "w") as f:
        f.write(("S" if spatial else "N") + "\t" + str(n) + "\t" + str(d) + "\n")
    lut = np.array([ord("0"), ord("1")], dtype=np.uint8)
    line = np.empty((n, 2 * d), dtype=np.uint8)
    line[:, 0::2] = lut[pg.x]
    line[:, 1::2] = ord("\t")
    line[:, -1] = ord("\n")
    with open(base + ".dat", "wb") as f:
        f.write(line.tobytes())
    with open(base + ".index", "w") as f:
        f.write("".join(f"{i + 1}\tfam{i + 1}\n" for i in range(n)))
    if spatial:
        rp, col, wgt = pg.row_ptr, pg.col, pg.wgt
        out = [f"{weighted_flag}\n"]
        for i in range(n):
            lo, hi = int(rp[i]), int(rp[i + 1])
            if hi == lo:
                out.append(f"{i + 1}\t0\n")
                continue
            items = [str(i + 1), str(hi - lo)] + [str(int(c) + 1) for c in col[lo:hi]]
            if weighted_flag:
                items += [_fmt_w(w) for w in wgt[lo:hi]]
            out.append("\t".join(items) + "\n")
        with open(base + ".nei", "w") as f:
            f.write("".join(out))
    if m_text is None:
        m_text = default_m_line(d)
    with open(base + ".m", "w") as f:
        f.write(m_text)


def read_uf(path: str, k: int) -> np.ndarray:
    return np.loadtxt(path, dtype=np.float64).reshape(-1, k)


def read_cf(path: str) -> np.ndarray:
    with open(path) as f:
        return np.array(f.read().split(), dtype=np.int64)


def read_mf(path: str, k: int, d: int) -> dict:
    """Parse ``.mf`` (layout nem_exe.c:1708-1773; consumer ppanggolin.py:1898-1923)."""
    with open(path) as f:
        return read_mf_text(f.read(), k, d)


def read_mf_text(text: str, k: int, d: int) -> dict:
    lines = text.splitlines()
    crit = [float(v) for v in lines[2].split()]
    out = {"U": crit[0], "D": crit[1], "L": crit[2], "M": crit[3], "err": crit[4],
           "beta": float(lines[5].split()[0])}
    mu = np.empty((k, d)); eps = np.empty((k, d)); p = np.empty(k)
    body = [ln for ln in lines if ln.strip()]
    for kk, ln in enumerate(body[-k:]):
        v = ln.split()
        mu[kk] = [float(t) for t in v[:d]]
        p[kk] = float(v[d])
        eps[kk] = [float(t) for t in v[d + 1:]]
    out.update(mu=mu, p=p, eps=eps)
    return out


def classify_psc(uf: np.ndarray, mf: dict) -> list[str]:
    """P/S/C labels with PPanGGOLiN's consistency check and tie rule (ppanggolin.py:1925-1972)."""
    sum_mu = [float((row != 0).sum()) for row in mf["mu"]]
    sum_eps = [float(row.sum()) for row in mf["eps"]]
    pk = sum_mu.index(max(sum_mu))
    sk = sum_eps.index(max(sum_eps))
    rest = list({0, 1, 2} - {pk, sk})
    if len(rest) != 1 or (pk, sk, rest[0]) != (0, 1, 2):
        return ["U"] * uf.shape[0]
    names = "PSC"
    out = []
    for row in uf:
        mx = row.max()
        pos = np.flatnonzero(row == mx)
        out.append("S" if pos.size > 1 else names[int(pos[0])])
    return out
