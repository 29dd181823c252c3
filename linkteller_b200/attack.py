"""The efficient influence attack as one call — what `Attacker.link_prediction_attack_efficient`
(attacker.py:209-247) computes before pair scoring: the n_test x n_test influence matrix.

`influence_matrix` takes device tensors; `influence_matrix_host` is the reference-facing entry with
HOST buffers (numpy / CPU tensors, the form the reference's Worker holds them in on a CPU run): it
stages them through pinned memory, runs the CUDA path and returns the matrix in host memory.

Multi-GPU: perturbed nodes shard across ranks — rank g owns the contiguous slice
test_nodes[g·T/G : (g+1)·T/G], computes its influence rows against ALL T observers from a replicated
graph, and the rows are all-gathered (NCCL over NVLink) into the full matrix on every rank.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops


def slice_bounds(T: int, world: int, rank: int):
    """Contiguous, near-equal slices of the perturbed nodes: rank r owns [lo, hi)."""
    base, rem = divmod(T, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group):
    """(world, rank) of an EXPLICITLY passed process group; without one the call is single-rank even when
    torch.distributed is initialised (a caller that attacks on one rank only must not hang in a collective)."""
    if group is None:
        return 1, 0
    return torch.distributed.get_world_size(group), torch.distributed.get_rank(group)


def make_node_comm(group, T: int, device, n: int = 0, hidden: int = 0) -> ops.NodeComm:
    """Symmetric influence matrices for the ranks of `group` (one node): CUDA IPC handles are exchanged with
    all_gather_object.  Raises if the peers cannot be mapped (no P2P / IPC): callers then keep the NCCL all-gather.
    With the graph size `n` and the hidden width the buffers also get room for the frontier rows' sign records, and the
    ranks shard the restricted precompute by frontier rows instead of each re-computing the shared hub rows."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def exchange(mine: bytes):
        got = [None] * world
        dist.all_gather_object(got, mine, group=group)
        return got
    return ops.NodeComm(rank, world, T, device, exchange, n=n, hidden=hidden)


def influence_matrix(adj_hat: ops.DeviceCSR, x, w1, b1, w2, test_nodes: torch.Tensor, delta: float,
                     group=None, out: Optional[torch.Tensor] = None, want_info: bool = False, comm=None):
    """[T,T] float32 influence matrix on the device; row i = perturb test_nodes[i], column j =
    ‖Δlogits[test_nodes[j]]‖₂/δ.  With a process group (passed explicitly), rows are sharded: through `comm`
    (make_node_comm) every rank's reduce kernel stores its entries straight into all ranks' matrices over NVLink and the
    result is comm.matrix; without it the dense rows are all-gathered with NCCL into `out`."""
    T = test_nodes.numel()
    world, rank = _world(group)
    if comm is not None:
        lo, hi = slice_bounds(T, comm.world, comm.rank)
        return ops.attack_efficient_comm(adj_hat, x, w1, b1, w2, test_nodes, delta, comm, lo, hi, want_info=want_info)
    if out is None:
        out = torch.empty((T, T), dtype=torch.float32, device=x.device)
    lo, hi = slice_bounds(T, world, rank)
    res = ops.attack_efficient(adj_hat, x, w1, b1, w2, test_nodes, delta, lo=lo, hi=hi, out=out[lo:hi],
                               want_info=want_info)
    if world > 1:
        gather_rows(out, T, world, group)
    if want_info:
        return out, res[1]
    return out


def influence_matrix_model(adj_hat: ops.DeviceCSR, x, weights, test_nodes: torch.Tensor, delta: float,
                           group=None, out: Optional[torch.Tensor] = None):
    """influence_matrix for the model the weights belong to: 4 tensors (w1,b1,w2,b2) → GCN, 6 → GCN3
    (--n-layer 3).  Same sharding: rank g owns a contiguous slice of the perturbed nodes."""
    T = test_nodes.numel()
    world, rank = _world(group)
    if out is None:
        out = torch.empty((T, T), dtype=torch.float32, device=x.device)
    lo, hi = slice_bounds(T, world, rank)
    ops.attack_rows(adj_hat, x, weights, test_nodes, delta, lo=lo, hi=hi, out=out[lo:hi])
    if world > 1:
        gather_rows(out, T, world, group)
    return out


def gather_rows(out: torch.Tensor, T: int, world: int, group=None):
    """All-gather the row slices in place.  Equal slices use one all_gather_into_tensor on the
    output buffer itself; ragged slices (T % world != 0) fall back to per-rank broadcasts."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    if T % world == 0:
        lo, hi = slice_bounds(T, world, rank)
        dist.all_gather_into_tensor(out, out[lo:hi], group=group)
    else:
        for r in range(world):
            lo, hi = slice_bounds(T, world, r)
            dist.broadcast(out[lo:hi], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return out


class HostStage:
    """Pinned host staging buffers + device twins for the host-buffer entry point, allocated once
    and reused across calls (the copies themselves happen on every call)."""

    def __init__(self):
        self._pinned = {}
        self._dev = {}

    def to_device(self, key, arr, device):
        t = torch.from_numpy(np.ascontiguousarray(arr)) if isinstance(arr, np.ndarray) else arr.contiguous()
        p = self._pinned.get(key)
        if p is None or p.shape != t.shape or p.dtype != t.dtype:
            p = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[key] = p
            self._dev[key] = torch.empty(t.shape, dtype=t.dtype, device=device)
        if t.data_ptr() != p.data_ptr():
            p.copy_(t)
        d = self._dev[key]
        d.copy_(p, non_blocking=True)
        return d

    def to_device_sharded(self, key, arr, device, world, rank, group):
        """Upload this rank's 1/world slice of `arr` and all-gather the slices (NCCL over NVLink) into the full array."""
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(arr)) if isinstance(arr, np.ndarray) else arr.contiguous()
        flat = t.reshape(-1)
        chunk = -(-flat.numel() // world)
        full = self._dev.get(key + "/flat")
        if full is None or full.numel() != chunk * world or full.dtype != t.dtype:
            full = torch.empty(chunk * world, dtype=t.dtype, device=device)
            self._dev[key + "/flat"] = full
            self._dev.pop(key, None)                                  # the whole-array twin adopt() made is not needed
        lo, hi = rank * chunk, min((rank + 1) * chunk, flat.numel())
        if hi > lo:
            full[lo:hi].copy_(flat[lo:hi], non_blocking=True)        # pinned when `arr` came from adopt()
        dist.all_gather_into_tensor(full, full[rank * chunk:(rank + 1) * chunk], group=group)
        return full[:flat.numel()].view(t.shape)

    def pinned(self, key, shape, dtype):
        p = self._pinned.get(key)
        if p is None or tuple(p.shape) != tuple(shape) or p.dtype != dtype:
            p = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pinned[key] = p
        return p

    def adopt(self, key, arr, device):
        """Register `arr` (copied once into pinned memory) and return the pinned tensor; later
        to_device(key, that_tensor) calls skip the pageable→pinned copy."""
        t = torch.from_numpy(np.ascontiguousarray(arr)) if isinstance(arr, np.ndarray) else arr.contiguous()
        p = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        p.copy_(t)
        self._pinned[key] = p
        self._dev[key] = torch.empty(t.shape, dtype=t.dtype, device=device)
        return p


def influence_matrix_host(indptr, indices, values, x, w1, b1, w2, test_nodes, delta: float, device="cuda",
                          stage: Optional[HostStage] = None, group=None, comm=None):
    """Host buffers in (Â as CSR int32/int32/float32, X, weights, test nodes), influence matrix out
    in (pinned) host memory.  Every call copies all inputs host→device and the result device→host.

    One rank: everything goes over this GPU's PCIe link.  Several ranks (explicit `group`): every rank holds the same
    host buffers (the reference's Worker is replicated per process), uploads only ITS 1/world slice of the three large
    arrays (indices, values, X) and the slices are all-gathered over NVLink — the host link carries each byte once per
    node instead of once per GPU; the small arrays are uploaded whole.  Only rank 0 reads the matrix back.
    Returns (matrix [T,T] float32 CPU tensor — None on ranks > 0, h2d_bytes of the whole node, d2h_bytes)."""
    stage = stage or HostStage()
    world, rank = _world(group)
    d = {}
    h2d = 0
    for key, arr in (("indptr", indptr), ("indices", indices), ("values", values), ("x", x), ("w1", w1), ("b1", b1),
                     ("w2", w2), ("test_nodes", test_nodes)):
        if world > 1 and key in ("indices", "values", "x"):
            d[key] = stage.to_device_sharded(key, arr, device, world, rank, group)
        else:
            d[key] = stage.to_device(key, arr, device)
        h2d += d[key].numel() * d[key].element_size() * (1 if key in ("indices", "values", "x") else world)
    n = d["indptr"].numel() - 1
    adj = ops.DeviceCSR(d["indptr"], d["indices"], d["values"], n)
    out = influence_matrix(adj, d["x"], d["w1"], d["b1"], d["w2"], d["test_nodes"], delta, group=group, comm=comm)
    host = None
    if rank == 0:
        host = stage.pinned("out", out.shape, out.dtype)
        host.copy_(out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host, h2d, out.numel() * out.element_size()
