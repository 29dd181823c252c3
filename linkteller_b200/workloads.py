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
from .synth import (chung_lu_edges_device, degree_thresholds, sample_test_nodes,  # noqa: F401  (re-exported)
                    sample_test_nodes_typed)


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
