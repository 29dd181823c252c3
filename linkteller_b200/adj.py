"""Â construction with the reference's own function names (utils/load.py), running on the device.

`graph_reader`-style edge lists, `fetch_normalization(type)` and
`sparse_mx_to_torch_sparse_tensor` keep their argument meaning and return types, so
`Worker.prepare_data` (worker.py:643-645) works unchanged when these replace the originals
(linkteller_b200.patch).  Internally the CSR build, the degree vector and the scaling run in
liblinkteller_b200 (csr_build.cu) and are bit-exact with scipy's results (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from . import ops

_DEVICE = "cuda"


def adj_from_edges(edges: np.ndarray, n_nodes: int, device=None) -> ops.DeviceCSR:
    """graph_reader's twitch branch (utils/load.py:452-460) on an in-memory edge list [m,2]:
    csr_matrix((ones,(r,c)), float32) + its transpose — duplicates summed, self loops doubled."""
    e = torch.from_numpy(np.ascontiguousarray(edges, dtype=np.int64)).to(device or _DEVICE)
    return ops.csr_from_edges(e, n_nodes)


def graph_reader_edges(edges: np.ndarray, n_nodes: int):
    """Same as adj_from_edges but returned as the scipy CSR the reference's callers expect."""
    return adj_from_edges(edges, n_nodes).to_scipy()


def normalize(adj, type: str, device=None) -> ops.DeviceCSR:
    """fetch_normalization(type)(adj) followed by the float32 cast (utils/load.py:617-627, 554) as a
    device CSR.  `adj`: scipy sparse matrix or DeviceCSR."""
    if not isinstance(adj, ops.DeviceCSR):
        adj = ops.DeviceCSR.from_scipy(sp.csr_matrix(adj), device or _DEVICE)
    return ops.normalize_adj(adj, type)


def fetch_normalization(type: str):
    """Drop-in for utils.load.fetch_normalization: returns a callable scipy → scipy COO.  Only the
    two normalisations on the hot path are accelerated; other names raise NotImplementedError
    (the reference returns a lambda yielding an error string for unknown names, load.py:626)."""
    if type not in ops.NORM_KINDS:
        raise NotImplementedError(f"normalisation {type!r} is outside the accelerated path "
                                  f"(supported: {sorted(ops.NORM_KINDS)})")

    def normalizer(adj):
        hat = normalize(adj, type)
        indptr = hat.indptr.cpu().numpy()
        rows = np.repeat(np.arange(hat.n, dtype=np.int32), np.diff(indptr))
        # float64 container like scipy's result; the values are already the float32-rounded ones that
        # sparse_mx_to_torch_sparse_tensor's astype(float32) would produce
        return sp.coo_matrix((hat.values.cpu().numpy().astype(np.float64), (rows, hat.indices.cpu().numpy())),
                             shape=(hat.n, hat.n))
    return normalizer


def sparse_mx_to_torch_sparse_tensor(sparse_mx):
    """utils/load.py:552-559 — scipy sparse → torch sparse COO float32 (row-major sorted, unique)."""
    sparse_mx = sparse_mx.tocoo().astype(np.float32)
    indices = torch.from_numpy(np.vstack((sparse_mx.row, sparse_mx.col)).astype(np.int64))
    values = torch.from_numpy(sparse_mx.data)
    return torch.sparse_coo_tensor(indices, values, torch.Size(sparse_mx.shape))


def device_csr_of(adj, device=None) -> ops.DeviceCSR:
    """Whatever the reference hands GraphConvolution / Attacker as `adj` → DeviceCSR (float32):
    a torch sparse COO tensor (CPU or CUDA), a scipy matrix, or a DeviceCSR already."""
    device = device or _DEVICE
    if isinstance(adj, ops.DeviceCSR):
        return adj
    if isinstance(adj, torch.Tensor):
        if adj.layout != torch.sparse_coo:
            raise TypeError("adjacency must be a torch sparse COO tensor")
        idx = adj._indices().to(device)
        val = adj._values().to(device=device, dtype=torch.float32)
        return ops.coo_to_csr(idx[0].contiguous(), idx[1].contiguous(), val, adj.shape[0])
    return ops.DeviceCSR.from_scipy(sp.csr_matrix(adj), device, value_dtype=np.float32)
