"""Synthetic workloads of the shapes BASELINE.json names, resident on the device.

Small shapes (twitch / ppi / flickr / tiny / small) are generated on the host by
``linkteller_b200.synth`` (plain numpy, so the oracle and the reference see the very same arrays)
and uploaded.  The large scale-free graph of config C5 (2.45 M nodes, ~62 M undirected edges) is
generated on the device with the same Chung-Lu recipe (torch CUDA RNG) because building 124 M
directed entries with scipy takes minutes; its adjacency is then built by the library's own CSR
kernels (lt_csr_from_edges → lt_adj_normalize), which is the reference's load path
(utils/load.py:452-460, 572-578).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

from . import ops, synth


@dataclasses.dataclass
class Workload:
    name: str
    n: int
    f: int
    h: int
    c: int
    T: int
    delta: float
    adj_ori: ops.DeviceCSR          # clean symmetric adjacency (labels come from this one)
    adj_hat: ops.DeviceCSR          # normalised Â the model is queried on
    x: torch.Tensor                 # [n, f] float32
    w1: torch.Tensor
    b1: torch.Tensor
    w2: torch.Tensor
    b2: torch.Tensor
    test_nodes: torch.Tensor        # [T] int32, sample order
    norm: str = "FirstOrderGCN"

    def describe(self):
        return dict(workload=self.name, n=self.n, nnz_adj=self.adj_ori.nnz, nnz_adj_hat=self.adj_hat.nnz, f=self.f,
                    hidden=self.h, c=self.c, n_test=self.T, influence=self.delta, norm=self.norm)


def chung_lu_edges_device(n: int, m: int, seed: int, device, alpha: float = 0.5) -> torch.Tensor:
    """~m undirected edges [k,2] int64 (self loops removed, repeats possible — the CSR build
    collapses them with binarize=True) of a power-law expected-degree graph; node ids are scattered
    by a random permutation so degree does not correlate with the id."""
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


def make(name: str, device="cuda", seed: int = 0, hidden: int = 256, delta: float = 1e-4,
         norm: str = "FirstOrderGCN", n_test: Optional[int] = None, sample_seed: int = 42,
         f: Optional[int] = None) -> Workload:
    cfg = dict(synth.SHAPES[name])
    if f is not None:
        cfg["f"] = f                                  # feature-width override (tests of the width-dependent paths)
    n, f, c = cfg["n"], cfg["f"], cfg["c"]
    T = n_test or cfg["T"]
    if name == "large":
        edges = chung_lu_edges_device(n, cfg["m"], seed, device)
        adj = ops.csr_from_edges(edges, n, binarize=True)
        del edges
        g = torch.Generator(device=device)
        g.manual_seed(seed + 1)
        x = torch.randn((n, f), generator=g, device=device, dtype=torch.float32)
        tn = sample_test_nodes(n, T, sample_seed)
    else:
        s = synth.make(name, seed=seed, hidden=hidden, f=f)
        adj = ops.csr_from_edges(torch.from_numpy(s.edges).to(device), n)
        x = torch.from_numpy(s.features).to(device)
        tn = sample_test_nodes_typed(cfg["dataset"], cfg["sample_type"], adj.indptr.cpu().numpy(), T, sample_seed)
    hat = ops.normalize_adj(adj, norm)
    w1, b1, w2, b2 = [torch.from_numpy(a).to(device) for a in synth.gcn_weights(f, hidden, c, seed=42)]
    return Workload(name=name, n=n, f=f, h=hidden, c=c, T=T, delta=delta, adj_ori=adj, adj_hat=hat, x=x, w1=w1,
                    b1=b1, w2=w2, b2=b2, test_nodes=torch.from_numpy(tn.astype(np.int32)).to(device), norm=norm)
