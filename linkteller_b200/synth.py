"""Seeded synthetic graphs / features / weights of the shapes BASELINE.json names.

There is no dataset in the reference tree (./data is a Google-Drive zip) and no
network, so every parity test and the bench run on synthetic inputs generated here.  The
generators are plain numpy (host) so that the reference, the oracle and the CUDA path can be fed
the very same arrays.  The large scale-free graph of config C5 is generated on the device by
``bench.py`` instead (same Chung-Lu recipe, torch RNG) because 124 M directed entries are too slow
to build with scipy.

Shapes:
  C1 twitch-RU : n=4385,  37 304 undirected edges, f=3170, c=2,   T=500
  C2 PPI       : n=14755, nnz~225K,               f=50,   c=121, T=500 (unbalanced-hi, deg>=30)
  C3 Flickr    : n=89250, nnz~900K,               f=500,  c=7,   T=2000
  C5 large     : n=2.45M, 61.9M undirected edges, f=100,  c=47,  T=8192
"""
from __future__ import annotations

import dataclasses

import numpy as np

SHAPES = {
    # name: n, undirected edges, f, c, T, dataset string handed to the samplers, sample_type
    "twitch": dict(n=4385, m=37304, f=3170, c=2, T=500, dataset="twitch/ES/RU", sample_type="unbalanced"),
    "ppi": dict(n=14755, m=112500, f=50, c=121, T=500, dataset="ppi", sample_type="unbalanced-hi"),
    "flickr": dict(n=89250, m=450000, f=500, c=7, T=2000, dataset="flickr", sample_type="unbalanced"),
    "large": dict(n=2449029, m=61859140, f=100, c=47, T=8192, dataset="large", sample_type="unbalanced"),
    # tiny variants used by the CPU-side and quick GPU parity tests
    "tiny": dict(n=300, m=1200, f=24, c=5, T=40, dataset="twitch/ES/RU", sample_type="unbalanced"),
    "small": dict(n=1500, m=9000, f=70, c=11, T=120, dataset="ppi", sample_type="unbalanced"),
}


def chung_lu_edges(n: int, m: int, rng: np.random.Generator, alpha: float = 0.5) -> np.ndarray:
    """``m`` distinct undirected edges (u<v) of a power-law expected-degree graph.

    Endpoint i is drawn with probability ∝ (i+1)^-alpha (inverse CDF: i = n·u^(1/(1-alpha))), which
    gives a degree exponent 1+1/alpha (3 for alpha=0.5) and hubs of degree ~ m/sqrt(n).  Self loops
    and repeated pairs are discarded and re-drawn, so the graph is simple like the real datasets.
    """
    inv = 1.0 / (1.0 - alpha)
    got = np.empty(0, dtype=np.int64)
    need = m
    while need > 0:
        k = int(need * 1.3) + 16
        a = np.minimum((n * rng.random(k) ** inv).astype(np.int64), n - 1)
        b = np.minimum((n * rng.random(k) ** inv).astype(np.int64), n - 1)
        keep = a != b
        lo = np.minimum(a[keep], b[keep])
        hi = np.maximum(a[keep], b[keep])
        got = np.unique(np.concatenate([got, lo * n + hi]))
        need = m - got.size
    if got.size > m:  # drop a seeded random surplus, keep key order
        got = np.sort(rng.permutation(got)[:m])
    # scatter node ids so that degree does not correlate with the id (real graphs are not sorted)
    perm = rng.permutation(n)
    edges = np.stack([perm[got // n], perm[got % n]], axis=1)
    return edges


@dataclasses.dataclass
class Synth:
    name: str
    n: int
    f: int
    c: int
    T: int
    dataset: str
    sample_type: str
    edges: np.ndarray          # [m, 2] int64, one row per undirected edge (like musae_*_edges.csv)
    features: np.ndarray       # [n, f] float32
    hidden: int = 256


def make(name: str, seed: int = 0, hidden: int = 256, **override) -> Synth:
    """Build the synthetic inputs of a named shape (numpy, deterministic in ``seed``)."""
    cfg = dict(SHAPES[name])
    cfg.update(override)
    rng = np.random.default_rng(seed)
    n, m, f = cfg["n"], cfg["m"], cfg["f"]
    edges = chung_lu_edges(n, m, rng)
    if name == "twitch":
        # binary attributes (~1 % ones) pushed through a StandardScaler, like worker.py:481-484
        x = (rng.random((n, f)) < 0.01).astype(np.float64)
        mu = x.mean(0)
        sd = x.std(0)
        sd[sd == 0] = 1.0
        feats = ((x - mu) / sd).astype(np.float32)
    else:
        feats = rng.standard_normal((n, f), dtype=np.float32)
    return Synth(name=name, n=n, f=f, c=cfg["c"], T=cfg["T"], dataset=cfg["dataset"],
                 sample_type=cfg["sample_type"], edges=edges, features=feats, hidden=hidden)


def gcn_weights(f: int, h: int, c: int, seed: int = 42):
    """Weights with the distribution of GraphConvolution.reset_parameters (gcn/layers.py:24-28):
    U(-1/sqrt(out), 1/sqrt(out)) for weight [in,out] and bias [out].  numpy RNG (not torch) so the
    arrays are identical on every box."""
    rng = np.random.default_rng(seed)
    s1 = 1.0 / np.sqrt(h)
    s2 = 1.0 / np.sqrt(c)
    w1 = rng.uniform(-s1, s1, (f, h)).astype(np.float32)
    b1 = rng.uniform(-s1, s1, h).astype(np.float32)
    w2 = rng.uniform(-s2, s2, (h, c)).astype(np.float32)
    b2 = rng.uniform(-s2, s2, c).astype(np.float32)
    return w1, b1, w2, b2


def gcn3_weights(f: int, h1: int, h2: int, c: int, seed: int = 42):
    """Same distribution for the three GraphConvolution layers of GCN3 (gcn/models.py:28-37)."""
    rng = np.random.default_rng(seed)
    out = []
    for fin, fout in ((f, h1), (h1, h2), (h2, c)):
        s = 1.0 / np.sqrt(fout)
        out.append(rng.uniform(-s, s, (fin, fout)).astype(np.float32))
        out.append(rng.uniform(-s, s, fout).astype(np.float32))
    return tuple(out)


# ---------------------------------------------------------------------------------- shared by the product arm
# (linkteller_b200.workloads) and the reference arm of bench.py (oracle/refgen.py): nothing here touches the CUDA library
def chung_lu_edges_device(n: int, m: int, seed: int, device, alpha: float = 0.5):
    """~m undirected edges [k,2] int64 (self loops removed, repeats possible — the CSR build
    collapses them with binarize=True) of a power-law expected-degree graph; node ids are scattered
    by a random permutation so degree does not correlate with the id."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    inv = 1.0 / (1.0 - alpha)
    k = int(m * 1.02) + 16

    def endpoints():
        u = torch.rand(k, generator=g, device=device, dtype=torch.float64)
        return (u.pow_(inv).mul_(n)).long().clamp_(max=n - 1)

    a, b = endpoints(), endpoints()
    keep = a != b
    a, b = a[keep][:m], b[keep][:m]
    perm = torch.randperm(n, generator=g, device=device)
    return torch.stack([perm[a], perm[b]], dim=1).contiguous()


def sample_test_nodes(n: int, T: int, sample_seed: int = 42) -> np.ndarray:
    """'unbalanced' sampling: np.random.seed(sample_seed); np.random.choice(n, T, replace=False)
    (attacker.py:45, utils/load.py:343-345, 379) — numpy's legacy stream, drawn on the host."""
    rs = np.random.RandomState(sample_seed)
    return rs.choice(np.arange(n), T, replace=False)


def sample_test_nodes_typed(dataset: str, sample_type: str, indptr: np.ndarray, T: int, sample_seed: int = 42):
    """construct_edge_sets_from_random_subgraph's node sampling (utils/load.py:338-379) for
    unbalanced / unbalanced-lo / unbalanced-hi."""
    n = len(indptr) - 1
    if sample_type == "unbalanced":
        cand = np.arange(n)
    elif sample_type in ("unbalanced-lo", "unbalanced-hi"):
        lo, hi = degree_thresholds(dataset)
        deg = np.diff(indptr)
        cand = np.where(deg <= lo)[0] if sample_type == "unbalanced-lo" else np.where(deg >= hi)[0]
    else:
        raise NotImplementedError(f"sample_type = {sample_type} not implemented!")
    return np.random.RandomState(sample_seed).choice(cand, T, replace=False)   # same legacy stream as np.random.seed


def degree_thresholds(dataset: str):
    """utils/load.py:354-370."""
    if dataset.startswith("twitch"):
        return (10 if "PTBR" in dataset else 5), 10
    if dataset in ("flickr", "ppi") or dataset.startswith("deezer"):
        return 15, 30
    if dataset in ("cora"):
        return 3, 4
    if dataset in ("citeseer"):
        return 3, 3
    if dataset in ("pubmed"):
        return 10, 10
    raise NotImplementedError(f"lo and hi for dataset = {dataset} not set!")


