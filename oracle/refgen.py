"""ORACLE-SIDE INPUTS AND RUNNER FOR THE REFERENCE ARM — test / bench infrastructure only.

`bench.py --impl reference` must time the reference's own implementation of the path on the very
workload the product arm measures WITHOUT loading the product's CUDA library.  This module builds
those inputs with numpy / torch ops only (never `linkteller_b200.ops`, never the `.so`):

  * small shapes (tiny / small / twitch / ppi / flickr): `linkteller_b200.synth` (numpy) → the
    oracle's numpy restatement of graph_reader + fetch_normalization (oracle.lt_oracle);
  * `large` (C5): the same torch-RNG Chung-Lu recipe the product arm uses (synth.chung_lu_edges_device,
    same seeds, same device ⇒ the same stream), then sort/unique + FirstOrderGCN scaling restated
    with torch ops (utils/load.py:452-460, 572-578, 552-559).  tests/test_gpu_large.py checks the
    result bit for bit against the product's lt_csr_from_edges → lt_adj_normalize output.

and drives the UNMODIFIED reference modules from oracle/_ref (oracle/build.py) — gcn.models.GCN,
attacker.Attacker.get_gradient_eps_mat and the `.norm().item()` loop of
link_prediction_attack_efficient (attacker.py:100-108, 220-229) — on CPU tensors (`reference`) or
through the reference's own `.cuda()` placement (`reference-cuda`, worker.py:665-678).
"""
from __future__ import annotations

import contextlib
import dataclasses
import io
import os
import sys
import time
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from linkteller_b200 import synth  # noqa: E402  (numpy/torch only — does not load the CUDA library)
from oracle import lt_oracle as O  # noqa: E402


@dataclasses.dataclass
class HostProblem:
    name: str
    n: int
    f: int
    h: int
    c: int
    T: int
    delta: float
    dataset: str
    nnz_adj: int
    X: np.ndarray                  # [n,f] float32
    coo: tuple                     # Â as (row int64, col int64, val float32), row-major sorted
    ori: tuple                     # clean adjacency CSR (indptr int32, indices int32, data float32)
    weights: tuple                 # (w1, b1, w2, b2) numpy float32
    test_nodes: np.ndarray         # [T] int64, sample order
    norm: str = "FirstOrderGCN"
    generator: str = ""

    def describe(self):
        """Same keys / values as linkteller_b200.workloads.Workload.describe()."""
        return dict(workload=self.name, n=self.n, nnz_adj=self.nnz_adj, nnz_adj_hat=len(self.coo[0]), f=self.f,
                    hidden=self.h, c=self.c, n_test=self.T, influence=self.delta, norm=self.norm)


# ------------------------------------------------------------------------------------------ graph (torch only)
def csr_from_edges_torch(edges: torch.Tensor, n: int):
    """graph_reader (utils/load.py:457-460) for an edge list with repeats, BINARISED like the product
    arm's generator asks (`binarize=True`): symmetrise, sort, collapse duplicates.  Returns
    (row int64, col int64) sorted row-major, on the edges' device."""
    a, b = edges[:, 0], edges[:, 1]
    keys = torch.cat([a * n + b, b * n + a])
    keys = torch.unique(keys)                                  # sorted
    return torch.div(keys, n, rounding_mode="floor"), keys % n


def normalize_first_order_torch(row: torch.Tensor, col: torch.Tensor, n: int):
    """gcn() on a float32 0/1 adjacency without self loops (utils/load.py:572-578) + the float32 COO
    cast (load.py:552-559): I + D^-1/2 A D^-1/2, every product rounded to float32
    (fl32(fl32(d_i·1)·d_j)); d = np.power(row_sum, -0.5) evaluated by numpy on the host exactly as
    upstream so the bits agree; inf → 0."""
    deg = torch.bincount(row, minlength=n).to(torch.float32).cpu().numpy()
    with np.errstate(divide="ignore"):
        d = np.power(deg, np.float32(-0.5)).astype(np.float32)
    d[np.isinf(d)] = 0.0
    dd = torch.from_numpy(d).to(row.device)
    v = (dd[row] * torch.ones((), dtype=torch.float32, device=row.device)) * dd[col]
    keep = v != 0
    row, col, v = row[keep], col[keep], v[keep]
    if bool((row == col).any()):
        raise ValueError("normalize_first_order_torch: adjacency has self loops (the generator removes them)")
    diag = torch.arange(n, dtype=torch.int64, device=row.device)
    keys = torch.cat([row * n + col, diag * (n + 1)])
    vals = torch.cat([v, torch.ones(n, dtype=torch.float32, device=row.device)])
    keys, order = torch.sort(keys)
    vals = vals[order]
    return torch.div(keys, n, rounding_mode="floor"), keys % n, vals


def host_problem(name: str, n_test=None, seed: int = 0, hidden: int = 256, delta: float = 1e-4, sample_seed: int = 42,
                 gen_device=None) -> HostProblem:
    """The workload `linkteller_b200.workloads.make(name)` builds, as host arrays, without the CUDA library."""
    cfg = dict(synth.SHAPES[name])
    n, f, c = cfg["n"], cfg["f"], cfg["c"]
    T = n_test or cfg["T"]
    weights = synth.gcn_weights(f, hidden, c, seed=42)
    if name == "large":
        dev = gen_device or ("cuda" if torch.cuda.is_available() else "cpu")
        edges = synth.chung_lu_edges_device(n, cfg["m"], seed, dev)
        row, col = csr_from_edges_torch(edges, n)
        del edges
        nnz_adj = int(row.numel())
        ip = torch.zeros(n + 1, dtype=torch.int64, device=row.device)
        ip[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
        ori = (ip.to(torch.int32).cpu().numpy(), col.to(torch.int32).cpu().numpy(), None)
        hr, hc, hv = normalize_first_order_torch(row, col, n)
        del row, col
        g = torch.Generator(device=dev)
        g.manual_seed(seed + 1)
        X = torch.randn((n, f), generator=g, device=dev, dtype=torch.float32).cpu().numpy()
        coo = (hr.cpu().numpy(), hc.cpu().numpy(), hv.cpu().numpy())
        del hr, hc, hv
        if dev != "cpu":
            torch.cuda.empty_cache()
        tn = synth.sample_test_nodes(n, T, sample_seed)
        gen = f"torch RNG on {dev} (same recipe and seeds as the product arm; identical graph when the device kind matches)"
    else:
        s = synth.make(name, seed=seed, hidden=hidden)
        ip, ix, dv = O.csr_from_edges(s.edges, n)
        nnz_adj = len(ix)
        ori = (ip, ix, dv)
        coo = O.normalize_adj(ip, ix, dv, n, "FirstOrderGCN")
        X = s.features
        tn = synth.sample_test_nodes_typed(cfg["dataset"], cfg["sample_type"], ip, T, sample_seed)
        gen = "numpy (linkteller_b200.synth) — bit-identical to the product arm's inputs"
    return HostProblem(name=name, n=n, f=f, h=hidden, c=c, T=T, delta=delta, dataset=cfg["dataset"], nnz_adj=nnz_adj,
                       X=np.ascontiguousarray(X), coo=coo, ori=ori, weights=weights,
                       test_nodes=np.asarray(tn, dtype=np.int64), generator=gen)


# ------------------------------------------------------------------------------------------ the reference itself
def have_ref() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "attacker.py"))


def import_ref():
    """Import the unmodified reference modules from oracle/_ref (never from /root/reference: that path does
    not exist on the GPU box).  Returns a namespace with attacker, utils, models (gcn.models)."""
    if not have_ref():
        raise FileNotFoundError("oracle/_ref is missing — run `python oracle/build.py` in the build container")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    if not hasattr(np, "bool"):
        np.bool = bool                     # worker.py:294 uses the alias numpy 2 removed
    with contextlib.redirect_stdout(io.StringIO()):
        import importlib
        att = importlib.import_module("attacker")
        utils = importlib.import_module("utils")
        models = importlib.import_module("gcn.models")
    for m in (att, utils, models):
        if not os.path.abspath(m.__file__).startswith(REF_DIR):
            raise ImportError(f"{m.__name__} resolved to {m.__file__}, not to oracle/_ref")
    return types.SimpleNamespace(attacker=att, utils=utils, models=models)


class ReferenceRunner:
    """The reference's Attacker on a HostProblem, tensors on `device` ('cpu', or 'cuda' for the
    reference's own `.cuda()` path)."""

    def __init__(self, prob: HostProblem, device="cpu", threads=None, dtype=torch.float32):
        import scipy.sparse as sp
        self.prob = prob
        self.device = torch.device(device)
        if threads:
            torch.set_num_threads(threads)
        ref = import_ref()
        row, col, val = prob.coo
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            adj = ref.utils.sparse_mx_to_torch_sparse_tensor(sp.coo_matrix((val, (row, col)), shape=(prob.n, prob.n)))
            with contextlib.redirect_stdout(io.StringIO()):
                model = ref.models.GCN(prob.f, prob.h, prob.c, 0.5)
        w1, b1, w2, b2 = [torch.from_numpy(np.asarray(w)) for w in prob.weights]
        model.load_state_dict({"gc1.weight": w1, "gc1.bias": b1, "gc2.weight": w2, "gc2.bias": b2})
        model.eval()
        feats = torch.from_numpy(prob.X)
        if dtype == torch.float64:
            model, feats, adj = model.double(), feats.double(), adj.double()
        if self.device.type == "cuda":                       # worker.py:665-678, gcn_trainer.py:109-110
            model, feats, adj = model.cuda(), feats.cuda(), adj.cuda()
        args = types.SimpleNamespace(dataset=prob.dataset, sample_type="unbalanced", n_test=prob.T, influence=prob.delta,
                                     sample_seed=42, mode="vanilla-clean", attack_mode="efficient")
        wk = types.SimpleNamespace(n_nodes=prob.n, n_features=prob.f, features=feats, adj_full=adj, features_2=feats,
                                   adj_2=adj, adj_ori=None)
        self.att = ref.attacker.Attacker(args, model, wk)
        self.att.test_nodes = prob.test_nodes
        self.what = "oracle/_ref (unmodified reference): Attacker.get_gradient_eps_mat + the T .norm().item() calls"

    def rows(self, idx):
        """Influence rows `idx` exactly as link_prediction_attack_efficient's loop body computes them
        (attacker.py:220-229).  Returns (values [len(idx), T] float64, seconds)."""
        att, tn, T = self.att, self.prob.test_nodes, self.prob.T
        out = np.zeros((len(idx), T))
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            for k, i in enumerate(idx):
                grad_mat = att.get_gradient_eps_mat(tn[i])
                for j in range(T):
                    out[k][j] = grad_mat[tn[j]].norm().item()
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        return out, time.perf_counter() - t0


class PortRunner:
    """Fallback when oracle/_ref is absent: the oracle's restatement of the same loop (kind "port")."""

    def __init__(self, prob: HostProblem, device="cpu", threads=None, dtype=torch.float32):
        if torch.device(device).type != "cpu":
            raise RuntimeError("the oracle port runs on the CPU only")
        self.prob, self.threads, self.dtype = prob, threads, dtype
        self.what = "oracle port (oracle/lt_oracle.influence_matrix): two torch-CPU forwards + T norms per row"

    def rows(self, idx):
        p = self.prob
        t0 = time.perf_counter()
        out = O.influence_matrix(p.X, p.coo, p.weights, p.n, p.test_nodes, p.delta, dtype=self.dtype, rows=list(idx),
                                 threads=self.threads)
        return out, time.perf_counter() - t0


def host_problem_from_workload(wl, dataset="synthetic") -> HostProblem:
    """Host copy of a device-resident linkteller_b200.workloads.Workload (bench.py's cpu_baseline / parity_check
    legs, where the product arm already holds the inputs): plain D2H copies, no product kernels."""
    hat = wl.adj_hat
    ip = hat.indptr.cpu().numpy()
    row = np.repeat(np.arange(wl.n, dtype=np.int64), np.diff(ip))
    coo = (row, hat.indices.cpu().numpy().astype(np.int64), hat.values.cpu().numpy())
    weights = tuple(t.cpu().numpy() for t in (wl.w1, wl.b1, wl.w2, wl.b2))
    return HostProblem(name=wl.name, n=wl.n, f=wl.f, h=wl.h, c=wl.c, T=wl.T, delta=wl.delta, dataset=dataset,
                       nnz_adj=wl.adj_ori.nnz, X=wl.x.cpu().numpy(), coo=coo, ori=(None, None, None), weights=weights,
                       test_nodes=wl.test_nodes.cpu().numpy().astype(np.int64), norm=wl.norm,
                       generator="copied from the product arm's device tensors")


def make_runner(prob: HostProblem, device="cpu", threads=None, dtype=torch.float32):
    """(runner, kind): the unmodified reference from oracle/_ref when it is there, else the oracle port."""
    if have_ref():
        return ReferenceRunner(prob, device=device, threads=threads, dtype=dtype), "reference"
    return PortRunner(prob, device=device, threads=threads, dtype=dtype), "port"
