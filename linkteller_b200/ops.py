"""Torch-tensor front end of the C ABI.  PyTorch is plumbing here (device memory, streams): every
function validates its tensors, passes raw device pointers + the current CUDA stream to
liblinkteller_b200.so and returns torch tensors that own the outputs."""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional

import numpy as np
import torch

from . import _lib

NORM_KINDS = {"FirstOrderGCN": 0, "AugNormAdj": 1}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, dtype, name: str, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got {t.dim()}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    return t


def _rowmajor(t: torch.Tensor, name: str):
    """2-D float32 CUDA tensor with unit column stride; returns (tensor, leading dimension)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2:
        raise TypeError(f"{name}: expected a 2-D float32 CUDA tensor")
    if t.stride(1) != 1 and t.shape[1] > 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    if ld < t.shape[1]:
        t = t.contiguous()
        ld = t.shape[1]
    return t, ld


def device_info():
    lib = _lib.load()
    sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    _lib.check(lib.lt_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)), "lt_device_info")
    return dict(sm_count=sm.value, cc=(ma.value, mi.value), mem_bytes=mem.value)


def launch_count() -> int:
    """Kernels launched by liblinkteller_b200 in this process so far."""
    return int(_lib.load().lt_launch_count())


def set_option(name: str, value: int) -> None:
    """Process-wide kernel-selection switch (lt_set_option): "overlap", "sign_path", "tc_ws", "spmm_panel", …"""
    _lib.check(_lib.load().lt_set_option(name.encode(), int(value)), "lt_set_option")


def get_option(name: str) -> int:
    v = C.c_int()
    _lib.check(_lib.load().lt_get_option(name.encode(), C.byref(v)), "lt_get_option")
    return v.value


class option:
    """`with ops.option("sign_path", 0): …` — set a switch and restore it afterwards."""

    def __init__(self, name: str, value: int):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = get_option(self.name)
        set_option(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.name, self.old)
        return False


class profile:
    """Context manager around the library's phase timers: `with ops.profile() as p: ...; p.report()`
    → {phase: (launch groups, total ms)} measured with CUDA events on the launching stream."""

    def __enter__(self):
        lib = _lib.load()
        lib.lt_prof_reset()
        lib.lt_prof_enable(1)
        self._rep = None
        return self

    def __exit__(self, *exc):
        lib = _lib.load()
        lib.lt_prof_enable(0)
        buf = C.create_string_buffer(1 << 16)
        _lib.check(lib.lt_prof_report(buf, len(buf)), "lt_prof_report")
        rep = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            rep[name] = (int(cnt), float(ms))
        self._rep = rep
        lib.lt_prof_reset()
        return False

    def report(self):
        return self._rep


@dataclasses.dataclass
class DeviceCSR:
    """CSR matrix resident in HBM: indptr int32[n+1], indices int32[nnz], values float32/float64."""
    indptr: torch.Tensor
    indices: torch.Tensor
    values: torch.Tensor
    n: int

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())

    @classmethod
    def from_scipy(cls, mat, device="cuda", value_dtype=None):
        mat = mat.tocsr()
        if not mat.has_canonical_format:
            mat = mat.copy()
            mat.sum_duplicates()
        vd = value_dtype or (np.float32 if mat.data.dtype == np.float32 else np.float64)
        return cls(torch.from_numpy(mat.indptr.astype(np.int32)).to(device),
                   torch.from_numpy(mat.indices.astype(np.int32)).to(device),
                   torch.from_numpy(mat.data.astype(vd)).to(device), int(mat.shape[0]))

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.values.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()),
                             shape=(self.n, self.n))


# ------------------------------------------------------------------------------------------ a1
def csr_from_edges(edges: torch.Tensor, n: int, binarize: bool = False) -> DeviceCSR:
    """graph_reader, utils/load.py:457-460: csr(ones,(r,c)) + its transpose, duplicates summed."""
    lib = _lib.load()
    _chk(edges, torch.int64, "edges", 2)
    m = edges.shape[0]
    dev = edges.device
    indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    cap = max(2 * m, 1)
    indices = torch.empty(cap, dtype=torch.int32, device=dev)
    values = torch.empty(cap, dtype=torch.float32, device=dev)
    nnz = C.c_int64()
    _lib.check(lib.lt_csr_from_edges(_p(edges), m, n, int(binarize), _p(indptr), _p(indices), _p(values), cap,
                                     C.byref(nnz), _stream()), "lt_csr_from_edges")
    k = nnz.value
    return DeviceCSR(indptr, indices[:k], values[:k], n)


def coo_to_csr(row: torch.Tensor, col: torch.Tensor, values: torch.Tensor, n: int, check: bool = True) -> DeviceCSR:
    """The torch COO `adj` (utils/load.py:552-559) → CSR; COO must be row-major sorted, no duplicates."""
    lib = _lib.load()
    _chk(row, torch.int64, "row", 1)
    _chk(col, torch.int64, "col", 1)
    nnz = row.numel()
    indptr = torch.empty(n + 1, dtype=torch.int32, device=row.device)
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=row.device)
    _lib.check(lib.lt_coo_to_csr(_p(row), _p(col), nnz, n, _p(indptr), _p(indices), _stream()), "lt_coo_to_csr")
    indices = indices[:nnz]
    if check:
        ok = C.c_int()
        _lib.check(lib.lt_csr_check(_p(indptr), _p(indices), n, nnz, C.byref(ok), _stream()), "lt_csr_check")
        if not ok.value:
            raise ValueError("adjacency COO tensor is not row-major sorted / duplicate-free; "
                             "build it with linkteller_b200.adj.sparse_mx_to_torch_sparse_tensor")
    return DeviceCSR(indptr, indices, values.contiguous(), n)


# ------------------------------------------------------------------------------------------ a2
def normalize_adj(adj: DeviceCSR, kind: str = "FirstOrderGCN") -> DeviceCSR:
    """fetch_normalization(kind)(adj) + float32 cast (utils/load.py:562-578, 554), bit-exact.

    Precision of the arithmetic follows the reference: float32 only for FirstOrderGCN on a float32
    adjacency, float64 otherwise.  d^-1/2 is evaluated on the host by numpy exactly as upstream
    (np.power(row_sum, -0.5); inf → 0) so that the values agree bit for bit."""
    lib = _lib.load()
    if kind not in NORM_KINDS:
        raise NotImplementedError(f"normalisation {kind!r} is outside the hot-path scope")
    k = NORM_KINDS[kind]
    f32 = adj.values.dtype == torch.float32 and kind == "FirstOrderGCN"
    vals = adj.values if f32 else adj.values.to(torch.float64)
    vals = vals.contiguous()
    prec = 0 if f32 else 1
    n = adj.n
    rs = torch.empty(n, dtype=vals.dtype, device=vals.device)
    _lib.check(lib.lt_csr_row_sums(_p(adj.indptr), _p(vals), n, prec, _p(rs), _stream()), "lt_csr_row_sums")
    rs_h = rs.cpu().numpy()
    if kind == "AugNormAdj":
        rs_h = rs_h + 1.0                                   # row sums of adj + sp.eye
    with np.errstate(divide="ignore"):
        d = np.power(rs_h, -0.5)                            # utils/load.py:565 / 575
    d[np.isinf(d)] = 0.0
    d_dev = torch.from_numpy(d).to(vals.device)
    cap = adj.nnz + n
    out_indptr = torch.empty(n + 1, dtype=torch.int32, device=vals.device)
    out_indices = torch.empty(cap, dtype=torch.int32, device=vals.device)
    out_values = torch.empty(cap, dtype=torch.float32, device=vals.device)
    nnz = C.c_int64()
    _lib.check(lib.lt_adj_normalize(_p(adj.indptr), _p(adj.indices), _p(vals), n, k, prec, _p(d_dev), _p(out_indptr),
                                    _p(out_indices), _p(out_values), cap, C.byref(nnz), _stream()), "lt_adj_normalize")
    m = nnz.value
    return DeviceCSR(out_indptr, out_indices[:m], out_values[:m], n)


# ------------------------------------------------------------------------------------------ a4/a5
def gemm(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
         a_rows: Optional[torch.Tensor] = None, trans_b: bool = False, algo: int = 0) -> torch.Tensor:
    lib = _lib.load()
    a, lda = _rowmajor(a, "a")
    b, ldb = _rowmajor(b, "b")
    K = a.shape[1]
    M = a.shape[0] if a_rows is None else a_rows.numel()
    N = b.shape[0] if trans_b else b.shape[1]
    if (b.shape[1] if trans_b else b.shape[0]) != K:
        raise ValueError("gemm: inner dimensions differ")
    if a_rows is not None:
        _chk(a_rows, torch.int32, "a_rows", 1)
    if bias is not None:
        _chk(bias, torch.float32, "bias", 1)
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _lib.check(lib.lt_gemm_f32(_p(a), lda, _p(a_rows), _p(b), ldb, int(trans_b), _p(bias), int(relu), _p(out), N, M, N,
                               K, algo, _stream()), "lt_gemm_f32")
    return out


def spmm(adj: DeviceCSR, b: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
         rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    b, ldb = _rowmajor(b, "b")
    _chk(adj.values, torch.float32, "adj.values", 1)
    if b.shape[0] != adj.n:
        raise ValueError("spmm: dense operand must have n rows")
    d = b.shape[1]
    n_rows = adj.n if rows is None else rows.numel()
    if rows is not None:
        _chk(rows, torch.int32, "rows", 1)
    if bias is not None:
        _chk(bias, torch.float32, "bias", 1)
    out = torch.empty((n_rows, d), dtype=torch.float32, device=b.device)
    _lib.check(lib.lt_spmm_csr_f32(_p(adj.indptr), _p(adj.indices), _p(adj.values), adj.n, _p(rows), n_rows, _p(b), ldb, d,
                                   _p(bias), int(relu), _p(out), d, _stream()), "lt_spmm_csr_f32")
    return out


def gcn_forward(adj: DeviceCSR, x, w1, b1, w2, b2) -> torch.Tensor:
    lib = _lib.load()
    x, ldx = _rowmajor(x, "x")
    w1 = _chk(w1, torch.float32, "w1", 2)
    w2 = _chk(w2, torch.float32, "w2", 2)
    n, f = x.shape
    h, c = w2.shape
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    _lib.check(lib.lt_gcn_forward(_p(adj.indptr), _p(adj.indices), _p(adj.values), n, _p(x), ldx, f, _p(w1), _p(b1),
                                  _p(w2), _p(b2), h, c, _p(out), c, _stream()), "lt_gcn_forward")
    return out


def gcn3_forward(adj: DeviceCSR, x, w1, b1, w2, b2, w3, b3) -> torch.Tensor:
    """GCN3.forward in eval mode (gcn/models.py:40-46): raw logits [n,c] in one C call."""
    lib = _lib.load()
    x, ldx = _rowmajor(x, "x")
    w1 = _chk(w1, torch.float32, "w1", 2)
    w2 = _chk(w2, torch.float32, "w2", 2)
    w3 = _chk(w3, torch.float32, "w3", 2)
    n, f = x.shape
    h1, h2 = w2.shape
    c = w3.shape[1]
    if w1.shape != (f, h1) or w3.shape[0] != h2:
        raise ValueError(f"GCN3 weight shapes do not chain: {tuple(w1.shape)}, {tuple(w2.shape)}, {tuple(w3.shape)}")
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    _lib.check(lib.lt_gcn3_forward(_p(adj.indptr), _p(adj.indices), _p(adj.values), n, _p(x), ldx, f, _p(w1), _p(b1),
                                   _p(w2), _p(b2), _p(w3), _p(b3), h1, h2, c, _p(out), c, _stream()), "lt_gcn3_forward")
    return out


def model_forward(adj: DeviceCSR, x, weights) -> torch.Tensor:
    """Raw logits of the model the weights belong to: 4 tensors → GCN, 6 → GCN3 (gcn_trainer.py:81-90)."""
    if len(weights) == 4:
        return gcn_forward(adj, x, *weights)
    if len(weights) == 6:
        return gcn3_forward(adj, x, *weights)
    raise NotImplementedError(f'n_layer = {len(weights) // 2} not implemented!')


# ------------------------------------------------------------------------------------------ a7/a8
INFO_KEYS = ("frontier_rows", "incidences", "observer_entries", "records", "support_pairs", "frontier_bytes",
             "frontier_row_nnz", "exact_path_incidences")


def attack_efficient(adj: DeviceCSR, x, w1, b1, w2, test_nodes: torch.Tensor, delta: float,
                     lo: int = 0, hi: Optional[int] = None, out: Optional[torch.Tensor] = None,
                     want_info: bool = False):
    """Influence rows [hi-lo, T] for the perturbed nodes test_nodes[lo:hi] — the whole hot path of
    one rank (plan + restricted precompute + frontier kernels) in one C call."""
    lib = _lib.load()
    x, ldx = _rowmajor(x, "x")
    w1 = _chk(w1, torch.float32, "w1", 2)
    w2 = _chk(w2, torch.float32, "w2", 2)
    _chk(b1, torch.float32, "b1", 1)
    _chk(test_nodes, torch.int32, "test_nodes", 1)
    _chk(adj.values, torch.float32, "adj.values", 1)
    T = test_nodes.numel()
    hi = T if hi is None else hi
    n, f = x.shape
    h, c = w2.shape
    if out is None:
        out = torch.empty((hi - lo, T), dtype=torch.float32, device=x.device)
    else:
        _chk(out, torch.float32, "out", 2)
    info = (C.c_int64 * 8)()
    if hi == lo:                                    # an empty slice (more ranks than test nodes): nothing to compute
        return (out, dict(zip(INFO_KEYS, [0] * 8))) if want_info else out
    _lib.check(lib.lt_attack_efficient(_p(adj.indptr), _p(adj.indices), _p(adj.values), n, _p(x), ldx, f, _p(w1),
                                       _p(b1), _p(w2), h, c, _p(test_nodes), T, lo, hi, float(delta), _p(out),
                                       out.stride(0), info if want_info else None, _stream()), "lt_attack_efficient")
    if want_info:
        return out, dict(zip(INFO_KEYS, list(info)))
    return out


class NodeComm:
    """The ranks of one node sharing their influence matrices (lt_comm): every rank owns a T x T float32 matrix in a
    symmetric buffer that is mapped into all other ranks by CUDA IPC; lt_attack_efficient_comm stores each entry into all
    of them from inside its reduce kernel.  `exchange(handle_bytes) -> list of every rank's handle bytes` is the host-side
    channel (attack.make_node_comm uses torch.distributed.all_gather_object)."""

    def __init__(self, rank: int, world: int, T: int, device, exchange, n: int = 0, hidden: int = 0):
        lib = _lib.load()
        self.rank, self.world, self.T = rank, world, T
        self.device = torch.device(device)
        self._h = C.c_void_p()
        # n and hidden given: room for the sign records of up to n frontier rows, which lets the ranks shard the
        # restricted precompute by frontier rows (lt_comm_create, aux_bytes)
        aux = n * ((hidden + 31) // 32 + 2) * 4 if n and hidden else 0
        with torch.cuda.device(self.device):
            _lib.check(lib.lt_comm_create(rank, world, T * T * 4, aux, C.byref(self._h)), "lt_comm_create")
            hb = C.create_string_buffer(64)
            _lib.check(lib.lt_comm_get_handle(self._h, hb), "lt_comm_get_handle")
            handles = exchange(hb.raw)
            if len(handles) != world or any(len(x) != 64 for x in handles):
                raise RuntimeError("NodeComm: the handle exchange must return one 64-byte handle per rank")
            _lib.check(lib.lt_comm_open(self._h, b"".join(handles)), "lt_comm_open")
            ptr = lib.lt_comm_buffer(self._h)

        class _Raw:                                         # torch wraps the library-owned buffer without copying
            __cuda_array_interface__ = {"shape": (T, T), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                        "strides": None}
        self._raw = _Raw()
        self.matrix = torch.as_tensor(self._raw, device=self.device)

    def timed_out(self) -> bool:
        v = C.c_int()
        _lib.check(_lib.load().lt_comm_timed_out(self._h, C.byref(v)), "lt_comm_timed_out")
        return bool(v.value)

    def close(self):
        if self._h:
            self.matrix = None
            _lib.load().lt_comm_destroy(self._h)
            self._h = C.c_void_p()


def attack_efficient_comm(adj: DeviceCSR, x, w1, b1, w2, test_nodes: torch.Tensor, delta: float, comm: NodeComm,
                          lo: int, hi: int, want_info: bool = False):
    """lt_attack_efficient_comm: this rank perturbs test_nodes[lo:hi); returns comm.matrix, which holds all T rows once
    the call's work on the current stream is done.  Every rank of the communicator must call."""
    lib = _lib.load()
    x, ldx = _rowmajor(x, "x")
    w1 = _chk(w1, torch.float32, "w1", 2)
    w2 = _chk(w2, torch.float32, "w2", 2)
    _chk(b1, torch.float32, "b1", 1)
    _chk(test_nodes, torch.int32, "test_nodes", 1)
    _chk(adj.values, torch.float32, "adj.values", 1)
    T = test_nodes.numel()
    if T != comm.T:
        raise ValueError(f"communicator was built for T = {comm.T}, attack has T = {T}")
    n, f = x.shape
    h, c = w2.shape
    info = (C.c_int64 * 8)()
    _lib.check(lib.lt_attack_efficient_comm(_p(adj.indptr), _p(adj.indices), _p(adj.values), n, _p(x), ldx, f, _p(w1), _p(b1),
                                            _p(w2), h, c, _p(test_nodes), T, lo, hi, float(delta), comm._h,
                                            info if want_info else None, _stream()), "lt_attack_efficient_comm")
    if want_info:
        return comm.matrix, dict(zip(INFO_KEYS, list(info)))
    return comm.matrix


INFO3_KEYS = ("chunks", "hop1_incidences", "hop2_records", "hop2_incidences", "hop3_records", "support_pairs",
              "reserved0", "reserved1")


def _pad_first_hidden(w1, b1, w2):
    """The frontier kernels read H1 / P rows as float4: pad the first hidden width to a multiple of 4 with
    zero units (zero column of W1, zero bias, zero row of W2 — such a unit never fires and feeds nothing)."""
    h = w1.shape[1]
    pad = (-h) % 4
    if pad == 0:
        return w1, b1, w2
    return (torch.nn.functional.pad(w1, (0, pad)).contiguous(), torch.nn.functional.pad(b1, (0, pad)).contiguous(),
            torch.nn.functional.pad(w2, (0, 0, 0, pad)).contiguous())


def attack_efficient3(adj: DeviceCSR, x, w1, b1, w2, b2, w3, test_nodes: torch.Tensor, delta: float,
                      lo: int = 0, hi: Optional[int] = None, out: Optional[torch.Tensor] = None,
                      want_info: bool = False, max_chunk_records: int = 0):
    """Influence rows [hi-lo, T] of the three-layer model (--n-layer 3, GCN3) for the perturbed nodes
    test_nodes[lo:hi] in one C call (lt_attack_efficient3)."""
    lib = _lib.load()
    x, ldx = _rowmajor(x, "x")
    w1 = _chk(w1, torch.float32, "w1", 2)
    w2 = _chk(w2, torch.float32, "w2", 2)
    w3 = _chk(w3, torch.float32, "w3", 2)
    _chk(b1, torch.float32, "b1", 1)
    _chk(b2, torch.float32, "b2", 1)
    _chk(test_nodes, torch.int32, "test_nodes", 1)
    _chk(adj.values, torch.float32, "adj.values", 1)
    w1, b1, w2 = _pad_first_hidden(w1, b1, w2)
    T = test_nodes.numel()
    hi = T if hi is None else hi
    n, f = x.shape
    h1, h2 = w2.shape
    c = w3.shape[1]
    if w1.shape != (f, h1) or w3.shape[0] != h2 or b1.numel() != h1 or b2.numel() != h2:
        raise ValueError(f"GCN3 weight shapes do not chain: {tuple(w1.shape)}, {tuple(w2.shape)}, {tuple(w3.shape)}")
    if out is None:
        out = torch.empty((hi - lo, T), dtype=torch.float32, device=x.device)
    else:
        _chk(out, torch.float32, "out", 2)
    info = (C.c_int64 * 8)()
    _lib.check(lib.lt_attack_efficient3(_p(adj.indptr), _p(adj.indices), _p(adj.values), n, _p(x), ldx, f, _p(w1),
                                        _p(b1), _p(w2), _p(b2), _p(w3), h1, h2, c, _p(test_nodes), T, lo, hi,
                                        float(delta), _p(out), out.stride(0), int(max_chunk_records),
                                        info if want_info else None, _stream()), "lt_attack_efficient3")
    if want_info:
        return out, dict(zip(INFO3_KEYS, list(info)))
    return out


def attack_rows(adj: DeviceCSR, x, weights, test_nodes: torch.Tensor, delta: float, lo: int = 0,
                hi: Optional[int] = None, out: Optional[torch.Tensor] = None, want_info: bool = False):
    """Influence rows for the model the weights belong to — (w1,b1,w2,b2): GCN, (w1,b1,w2,b2,w3,b3): GCN3.
    The last bias cancels in the finite difference and is not read."""
    if len(weights) == 4:
        w1, b1, w2 = _pad_first_hidden(weights[0], weights[1], weights[2])
        return attack_efficient(adj, x, w1, b1, w2, test_nodes, delta, lo=lo, hi=hi, out=out, want_info=want_info)
    if len(weights) == 6:
        return attack_efficient3(adj, x, *weights[:5], test_nodes, delta, lo=lo, hi=hi, out=out, want_info=want_info)
    raise NotImplementedError(f'n_layer = {len(weights) // 2} not implemented!')


class Plan:
    """Symbolic structure of one attack slice (lt_plan); owns device buffers until closed."""

    def __init__(self, adj: DeviceCSR, test_nodes: torch.Tensor, lo: int = 0, hi: Optional[int] = None):
        lib = _lib.load()
        _chk(test_nodes, torch.int32, "test_nodes", 1)
        _chk(adj.values, torch.float32, "adj.values", 1)
        self.T = test_nodes.numel()
        self.lo, self.hi = lo, self.T if hi is None else hi
        self._h = C.c_void_p()
        self._keep = (adj, test_nodes)
        _lib.check(lib.lt_plan_create(_p(adj.indptr), _p(adj.indices), _p(adj.values), adj.n, _p(test_nodes), self.T,
                                      self.lo, self.hi, _stream(), C.byref(self._h)), "lt_plan_create")

    def info(self, h: int, c: int):
        info = (C.c_int64 * 8)()
        _lib.check(_lib.load().lt_plan_info(self._h, h, c, info, _stream()), "lt_plan_info")
        return dict(zip(INFO_KEYS, list(info)))

    def rows(self) -> torch.Tensor:
        """Frontier rows R (node ids, ascending) as an int32 CUDA tensor (copy)."""
        lib = _lib.load()
        r = self.info(4, 1)["frontier_rows"]
        out = torch.empty(r, dtype=torch.int32, device=self._keep[1].device)
        _lib.check(lib.lt_plan_copy_rows(self._h, _p(out), _stream()), "lt_plan_copy_rows")
        return out

    def influence_rows(self, h_r: torch.Tensor, p: torch.Tensor, w2: torch.Tensor, delta: float,
                       p_index: Optional[torch.Tensor] = None) -> torch.Tensor:
        lib = _lib.load()
        h_r, ldh = _rowmajor(h_r, "h_r")
        p, ldp = _rowmajor(p, "p")
        w2 = _chk(w2, torch.float32, "w2", 2)
        h, c = w2.shape
        out = torch.empty((self.hi - self.lo, self.T), dtype=torch.float32, device=p.device)
        _lib.check(lib.lt_influence_rows(self._h, _p(h_r), ldh, _p(p), ldp, _p(p_index), _p(w2), h, c, float(delta),
                                         _p(out), self.T, _stream()), "lt_influence_rows")
        return out

    def close(self):
        if self._h:
            _lib.load().lt_plan_destroy(self._h, _stream())
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------ a6/a8/a9
def n_pairs(T: int) -> int:
    return T * (T - 1) // 2


def pair_labels(adj_ori: DeviceCSR, test_nodes: torch.Tensor):
    """_get_edge_sets_among_nodes (utils/load.py:304-326): uint8 label of every pair i<j (row-major,
    sample order) = test_nodes[j] stored in row test_nodes[i] of adj_ori.  Returns (labels, n_pos)."""
    lib = _lib.load()
    _chk(test_nodes, torch.int32, "test_nodes", 1)
    T = test_nodes.numel()
    labels = torch.empty(max(n_pairs(T), 1), dtype=torch.uint8, device=test_nodes.device)
    npos = C.c_int64()
    _lib.check(lib.lt_pair_labels(_p(adj_ori.indptr), _p(adj_ori.indices), adj_ori.n, _p(test_nodes), T, _p(labels),
                                  C.byref(npos), _stream()), "lt_pair_labels")
    return labels[:n_pairs(T)], npos.value


def pair_scores(influence: torch.Tensor) -> torch.Tensor:
    """score(i<j) = influence[j][i] (attacker.py:233-245), float32 [T(T-1)/2] in pair order."""
    lib = _lib.load()
    influence, ldi = _rowmajor(influence, "influence")
    T = influence.shape[0]
    if influence.shape[1] != T:
        raise ValueError("influence must be square [T,T]")
    out = torch.empty(max(n_pairs(T), 1), dtype=torch.float32, device=influence.device)
    _lib.check(lib.lt_pair_scores(_p(influence), ldi, T, _p(out), _stream()), "lt_pair_scores")
    return out[:n_pairs(T)]


def auc_ap(scores: torch.Tensor, labels: torch.Tensor):
    """ROC-AUC and average precision with sklearn's tie semantics (attacker.py:382-386)."""
    lib = _lib.load()
    _chk(scores, torch.float32, "scores", 1)
    _chk(labels, torch.uint8, "labels", 1)
    if scores.numel() != labels.numel():
        raise ValueError("scores and labels differ in length")
    auc, ap = C.c_double(), C.c_double()
    _lib.check(lib.lt_auc_ap(_p(scores), _p(labels), scores.numel(), C.byref(auc), C.byref(ap), _stream()),
               "lt_auc_ap")
    return auc.value, ap.value


def topk_tp(scores: torch.Tensor, labels: torch.Tensor, n_pos):
    """True positives among the global top-n_pos scores for every n_pos (attack_stats_all.py:108-116);
    returns (tp_min, tp_max) lists over the possible tie-breaks at the cut."""
    lib = _lib.load()
    _chk(scores, torch.float32, "scores", 1)
    _chk(labels, torch.uint8, "labels", 1)
    q = len(n_pos)
    arr = (C.c_int64 * q)(*[int(v) for v in n_pos])
    lo, hi = (C.c_int64 * q)(), (C.c_int64 * q)()
    _lib.check(lib.lt_topk_tp(_p(scores), _p(labels), scores.numel(), arr, q, lo, hi, _stream()), "lt_topk_tp")
    return list(lo), list(hi)


def rank_curve(scores: torch.Tensor, labels: torch.Tensor):
    """sklearn's _binary_clf_curve on the device: (thresholds float32, tps int64, fps int64), one point
    per distinct score, scores descending."""
    lib = _lib.load()
    _chk(scores, torch.float32, "scores", 1)
    _chk(labels, torch.uint8, "labels", 1)
    n = scores.numel()
    g = C.c_int64()
    _lib.check(lib.lt_rank_curve(_p(scores), _p(labels), n, None, None, None, 0, C.byref(g), _stream()), "lt_rank_curve")
    G = g.value
    thr = torch.empty(G, dtype=torch.float32, device=scores.device)
    tps = torch.empty(G, dtype=torch.int64, device=scores.device)
    fps = torch.empty(G, dtype=torch.int64, device=scores.device)
    _lib.check(lib.lt_rank_curve(_p(scores), _p(labels), n, _p(thr), _p(tps), _p(fps), G, C.byref(g), _stream()),
               "lt_rank_curve")
    return thr, tps, fps


# ------------------------------------------------------------------------------------------ a10
LSA2_TRANSFORMS = {"none": 0, "softmax": 1, "sigmoid": 2}


def lsa2_corr(mat: torch.Tensor, test_nodes: torch.Tensor, transform: str = "none") -> torch.Tensor:
    """baseline_attack (attacker.py:287-334): dist [T,T] float32, strict upper triangle filled."""
    lib = _lib.load()
    mat, ldm = _rowmajor(mat, "mat")
    _chk(test_nodes, torch.int32, "test_nodes", 1)
    T = test_nodes.numel()
    if T and (int(test_nodes.min()) < 0 or int(test_nodes.max()) >= mat.shape[0]):
        raise IndexError("test node outside the matrix")
    dist = torch.empty((T, T), dtype=torch.float32, device=mat.device)
    _lib.check(lib.lt_lsa2_corr(_p(mat), ldm, mat.shape[1], _p(test_nodes), T, LSA2_TRANSFORMS[transform], _p(dist), T,
                                _stream()), "lt_lsa2_corr")
    return dist


def lsa2_pairs(mat: torch.Tensor, pu: torch.Tensor, pv: torch.Tensor, transform: str = "none") -> torch.Tensor:
    """baseline_attack_balanced (attacker.py:337-375): centred cosine similarity of LISTED pairs, the
    mean taken over all rows of `mat`; float32 [K] in list order."""
    lib = _lib.load()
    mat, ldm = _rowmajor(mat, "mat")
    _chk(pu, torch.int32, "pu", 1)
    _chk(pv, torch.int32, "pv", 1)
    K = pu.numel()
    if pv.numel() != K:
        raise ValueError("pu and pv must have the same length")
    n = mat.shape[0]
    if K and (int(torch.minimum(pu.min(), pv.min())) < 0 or int(torch.maximum(pu.max(), pv.max())) >= n):
        raise IndexError("pair node outside the matrix")
    out = torch.empty(max(K, 1), dtype=torch.float32, device=mat.device)
    _lib.check(lib.lt_lsa2_pairs(_p(mat), ldm, n, mat.shape[1], LSA2_TRANSFORMS[transform], _p(pu), _p(pv), K, _p(out),
                                 _stream()), "lt_lsa2_pairs")
    return out[:K]


def pair_gather(mat: torch.Tensor, rows: torch.Tensor, cols: torch.Tensor, row0: int = 0) -> torch.Tensor:
    """out[k] = mat[rows[k] − row0, cols[k]] for the listed entries whose row lies inside `mat`
    (others 0) — lt_pair_gather."""
    lib = _lib.load()
    mat, ld = _rowmajor(mat, "mat")
    _chk(rows, torch.int32, "rows", 1)
    _chk(cols, torch.int32, "cols", 1)
    K = rows.numel()
    if cols.numel() != K:
        raise ValueError("rows and cols must have the same length")
    if K and (int(cols.min()) < 0 or int(cols.max()) >= mat.shape[1]):
        raise IndexError("column outside the matrix")
    out = torch.zeros(max(K, 1), dtype=torch.float32, device=mat.device)
    _lib.check(lib.lt_pair_gather(_p(mat), ld, row0, mat.shape[0], _p(rows), _p(cols), K, _p(out), _stream()),
               "lt_pair_gather")
    return out[:K]


def pair_influence(adj: DeviceCSR, x, w1, b1, w2, perturbed: torch.Tensor, observed: torch.Tensor, delta: float,
                   block_bytes: int = 512 << 20, b2=None, w3=None) -> torch.Tensor:
    """‖((f(X + δ·e_p x_pᵀ) − f(X))/δ)[o]‖₂ for LISTED (perturbed p, observed o) node pairs — what
    link_prediction_attack_efficient_balanced (attacker.py:250-284) and the naive attack
    (attacker.py:143-160) read one entry at a time.  float32 [K].

    Two-layer model: lt_pair_influence — incidences and ΔQ for the distinct perturbed nodes only, then one warp per
    listed pair intersects N̂(p) ∩ N̂(o); nothing of size n x n (or rows x n) is formed, so balanced-full on a
    Flickr-sized graph (every node perturbed, 900 K listed pairs) is a fraction of a second.
    Three-layer model (b2, w3 given): influence rows in blocks of perturbed nodes (lt_attack_efficient3) and a gather."""
    lib = _lib.load()
    _chk(perturbed, torch.int32, "perturbed", 1)
    _chk(observed, torch.int32, "observed", 1)
    K = perturbed.numel()
    if observed.numel() != K:
        raise ValueError("perturbed and observed must have the same length")
    n = adj.n
    out = torch.zeros(max(K, 1), dtype=torch.float32, device=x.device)
    if K == 0:
        return out[:0]
    if int(torch.minimum(perturbed.min(), observed.min())) < 0 or int(torch.maximum(perturbed.max(), observed.max())) >= n:
        raise IndexError("pair node outside the graph")
    w1, b1, w2 = _pad_first_hidden(w1, b1, w2)
    if w3 is None:
        x, ldx = _rowmajor(x, "x")
        w1 = _chk(w1, torch.float32, "w1", 2)
        w2 = _chk(w2, torch.float32, "w2", 2)
        _chk(b1, torch.float32, "b1", 1)
        _chk(adj.values, torch.float32, "adj.values", 1)
        uniq, inv = torch.unique(perturbed, return_inverse=True)          # the distinct perturbed nodes, ascending
        tn = uniq.to(torch.int32).contiguous()
        pj = inv.to(torch.int32).contiguous()
        f = x.shape[1]
        h, c = w2.shape
        _lib.check(lib.lt_pair_influence(_p(adj.indptr), _p(adj.indices), _p(adj.values), n, _p(x), ldx, f, _p(w1), _p(b1), _p(w2),
                                         h, c, _p(tn), tn.numel(), _p(pj), _p(observed), K, float(delta), _p(out), _stream()),
                   "lt_pair_influence")
        return out[:K]
    all_nodes = torch.arange(n, dtype=torch.int32, device=x.device)
    need = torch.zeros(n, dtype=torch.bool, device=x.device)
    need[perturbed.long()] = True
    need = need.cpu().numpy()
    rows_per_block = max(1, min(n, block_bytes // (4 * n)))
    buf = torch.empty((rows_per_block, n), dtype=torch.float32, device=x.device)
    for lo in range(0, n, rows_per_block):
        hi = min(n, lo + rows_per_block)
        if not need[lo:hi].any():
            continue
        attack_efficient3(adj, x, w1, b1, w2, b2, w3, all_nodes, delta, lo=lo, hi=hi, out=buf[: hi - lo])
        _lib.check(lib.lt_pair_gather(_p(buf), buf.stride(0), lo, hi - lo, _p(perturbed), _p(observed), K, _p(out),
                                      _stream()), "lt_pair_gather")
    return out[:K]


# ------------------------------------------------------------------------------------------ a11/a12
def dp_edgerand_apply(adj: DeviceCSR, hits: torch.Tensor, coin: torch.Tensor) -> DeviceCSR:
    """adj + add − minus with −1→0, 2→1 (worker.py:264-267) from host-drawn flips.  hits: int64 cell
    ids r*n+c (r<c, ascending); coin: uint8 (1 = add).  Output values float64 (scipy promotes)."""
    lib = _lib.load()
    _chk(hits, torch.int64, "hits", 1)
    _chk(coin, torch.uint8, "coin", 1)
    vals = adj.values.to(torch.float64).contiguous()
    n, k = adj.n, hits.numel()
    cap = max(adj.nnz + 2 * k, 1)
    dev = vals.device
    oip = torch.empty(n + 1, dtype=torch.int32, device=dev)
    oix = torch.empty(cap, dtype=torch.int32, device=dev)
    ov = torch.empty(cap, dtype=torch.float64, device=dev)
    nnz = C.c_int64()
    _lib.check(lib.lt_dp_edgerand_apply(_p(adj.indptr), _p(adj.indices), _p(vals), n, _p(hits), _p(coin), k, _p(oip),
                                        _p(oix), _p(ov), cap, C.byref(nnz), _stream()), "lt_dp_edgerand_apply")
    return DeviceCSR(oip, oix[:nnz.value], ov[:nnz.value], n)


def dp_lapgraph_select(adj: DeviceCSR, noise: torch.Tensor, n_keep: int) -> DeviceCSR:
    """Top-n_keep cells of tril(adj,-1)+noise → symmetric binary CSR (worker.py:298-335).  `noise`
    ([n,n] float64, device) is consumed.  Values are float64 ones so that the normalisation that
    follows runs in float64 like the reference's int32 matrix does."""
    lib = _lib.load()
    _chk(noise, torch.float64, "noise", 2)
    n = adj.n
    if tuple(noise.shape) != (n, n):
        raise ValueError("noise must be [n,n]")
    vals = adj.values.to(torch.float64).contiguous()
    cap = max(2 * int(n_keep), 1)
    dev = vals.device
    oip = torch.empty(n + 1, dtype=torch.int32, device=dev)
    oix = torch.empty(cap, dtype=torch.int32, device=dev)
    ov = torch.empty(cap, dtype=torch.float32, device=dev)
    nnz = C.c_int64()
    _lib.check(lib.lt_dp_lapgraph_select(_p(adj.indptr), _p(adj.indices), _p(vals), n, _p(noise), int(n_keep), _p(oip),
                                         _p(oix), _p(ov), cap, C.byref(nnz), _stream()), "lt_dp_lapgraph_select")
    return DeviceCSR(oip, oix[:nnz.value], ov[:nnz.value].to(torch.float64), n)


# ------------------------------------------------------------------------------------------ primitives (tests)
def exclusive_scan_i32(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _chk(x, torch.int32, "x", 1)
    out = torch.empty(x.numel() + 1, dtype=torch.int32, device=x.device)
    _lib.check(lib.lt_exclusive_scan_i32(_p(x), _p(out), x.numel(), _stream()), "lt_exclusive_scan_i32")
    return out


def radix_sort_pairs(keys: torch.Tensor, vals: Optional[torch.Tensor], begin_bit=0, end_bit=64):
    """In-place stable sort; keys/vals are int64 tensors reinterpreted as uint64."""
    lib = _lib.load()
    _chk(keys, torch.int64, "keys", 1)
    if vals is not None:
        _chk(vals, torch.int64, "vals", 1)
    _lib.check(lib.lt_radix_sort_pairs_u64(_p(keys), _p(vals), keys.numel(), begin_bit, end_bit, _stream()),
               "lt_radix_sort_pairs_u64")
    return keys, vals
