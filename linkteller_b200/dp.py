"""DP adjacency perturbation — Worker.perturb_adj (worker.py:206-335), host RNG + device apply.

The random streams stay on the host so that the same seed gives the same flips / noise as the
reference (numpy's legacy MT19937 `np.random.seed` stream: binomial by inversion, Laplace by inverse
CDF); everything after the draw — merging the flips into the adjacency, the dense add and the exact
global top-k, rebuilding the CSR — runs on the device (linkteller_b200/csrc/dp.cu).  Like the
reference this materialises N×N host arrays, so it is bounded to the same graph sizes (≲ 50 K nodes).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def get_noise(noise_type, size, seed, eps=10, delta=1e-5, sensitivity=2):
    """utils/load.py:27-39 — same stream, same call order."""
    rs = np.random.RandomState(seed)
    if noise_type == "laplace":
        return rs.laplace(0, sensitivity / eps, size)
    if noise_type == "gaussian":
        c = np.sqrt(2 * np.log(1.25 / delta))
        return rs.normal(0, c * sensitivity / eps, size)
    raise NotImplementedError("noise {} not implemented!".format(noise_type))


def edgerand_flips(n: int, epsilon: float, noise_seed: int):
    """The host half of perturb_adj_discrete (worker.py:214-231): Bernoulli(s) hits over the N×N
    grid in row-major order and their add/remove coin, restricted to r<c as construct_sparse_mat
    does (worker.py:184-185).  Returns (hits int64 cell ids ascending, coin uint8)."""
    s = 2 / (np.exp(epsilon) + 1)
    rs = np.random.RandomState(noise_seed)
    bern = rs.binomial(1, s, (n, n))
    r, c = np.where(bern)
    dig = rs.binomial(1, 1 / 2, len(r))          # drawn for every hit, before the r<c filter
    up = r < c
    return (r[up].astype(np.int64) * n + c[up].astype(np.int64)), dig[up].astype(np.uint8)


def perturb_adj_discrete(adj: ops.DeviceCSR, epsilon: float, noise_seed: int) -> ops.DeviceCSR:
    """EdgeRand: returns the noisy adjacency (float64 values, explicit zeros where a removal hit a
    non-edge — they are pruned by the normalisation that follows, as upstream)."""
    hits, coin = edgerand_flips(adj.n, epsilon, noise_seed)
    dev = adj.indptr.device
    return ops.dp_edgerand_apply(adj, torch.from_numpy(hits).to(dev), torch.from_numpy(coin).to(dev))


def lapgraph_noise(n: int, n_edges: int, epsilon: float, noise_seed: int, noise_type="laplace", delta=1e-5):
    """The host half of perturb_adj_continuous (worker.py:290-304): the N×N noise matrix for ε₂ and
    the noisy edge count for ε₁ = 0.01ε, drawn from the re-seeded stream."""
    eps_1 = epsilon * 0.01
    eps_2 = epsilon - eps_1
    noise = get_noise(noise_type, (n, n), noise_seed, eps=eps_2, delta=delta, sensitivity=1)
    n_keep = n_edges + int(get_noise(noise_type, 1, noise_seed, eps=eps_1, delta=delta, sensitivity=1)[0])
    return noise, n_keep


def perturb_adj_continuous(adj: ops.DeviceCSR, epsilon: float, noise_seed: int, noise_type="laplace",
                           delta=1e-5) -> ops.DeviceCSR:
    """Lapgraph: symmetric binary adjacency of the top-n_keep noisy lower-triangle cells."""
    noise, n_keep = lapgraph_noise(adj.n, adj.nnz // 2, epsilon, noise_seed, noise_type, delta)
    dev = adj.indptr.device
    return ops.dp_lapgraph_select(adj, torch.from_numpy(noise).to(dev), n_keep)


def perturb_adj(adj: ops.DeviceCSR, perturb_type: str, epsilon: float, noise_seed: int, noise_type="laplace",
                delta=1e-5) -> ops.DeviceCSR:
    """Worker.perturb_adj (worker.py:206-210): 'discrete' → EdgeRand, anything else → Lapgraph."""
    if perturb_type == "discrete":
        return perturb_adj_discrete(adj, epsilon, noise_seed)
    return perturb_adj_continuous(adj, epsilon, noise_seed, noise_type, delta)
