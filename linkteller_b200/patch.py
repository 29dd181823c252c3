"""Drop-in installation behind the reference's unchanged main.py.

The reference has no plugin or FFI layer: its seam for this path is a set of Python names that
`main.py` → `GCNTrainer` → `Attacker` / `Worker` resolve at call time.  `install()` rebinds exactly
those names to the B200 implementations, leaving every reference file untouched:

    attacker.Attacker, gcn_trainer.Attacker          → linkteller_b200.attacker.Attacker
    gcn.layers.GraphConvolution.forward              → sm_100a gemm + CSR SpMM (autograd kept)
    gcn.models.GCN.forward (eval, no_grad)           → lt_gcn_forward
    gcn.models.GCN3.forward (eval, no_grad)          → lt_gcn3_forward   (--n-layer 3)
    utils.fetch_normalization, worker.fetch_normalization            → device Â build (bit-exact)
    worker.Worker.perturb_adj                        → host RNG + device apply (EdgeRand / Lapgraph)

Usage (from the reference checkout):   python -c "import linkteller_b200.patch as p; p.install(); \\
    import runpy; runpy.run_path('main.py', run_name='__main__')" --dataset twitch/ES/RU --attack …
`install()` fails loudly if the CUDA library is missing or no GPU is visible — there is no CPU path.
"""
from __future__ import annotations

import importlib
import sys
import types

import scipy.sparse as sp
import torch


def _shim_imports():
    """gcn_trainer.py imports two things this image lacks and never uses on the attack path
    (pandas.core.base.DataError, tensorboardX.SummaryWriter — SURVEY.md §4)."""
    try:
        import pandas.core.base as pcb
        if not hasattr(pcb, "DataError"):
            pcb.DataError = type("DataError", (Exception,), {})
    except Exception:
        pass
    if "tensorboardX" not in sys.modules:
        try:
            importlib.import_module("tensorboardX")
        except Exception:
            m = types.ModuleType("tensorboardX")

            class SummaryWriter:                      # noqa: D401 - minimal stand-in
                def __init__(self, *a, **k):
                    pass

                def __getattr__(self, name):
                    return lambda *a, **k: None
            m.SummaryWriter = SummaryWriter
            sys.modules["tensorboardX"] = m


def install(reference_path=None, patch_layers=True, patch_normalization=True, patch_dp=True):
    """Rebind the reference's hot-path names.  Returns the dict of patched (module, name) pairs."""
    from . import _lib, adj as _adj, attacker as _att, dp as _dp, gcn as _gcn, ops
    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("linkteller_b200.patch.install: no CUDA device visible (the B200 path has no CPU fallback)")
    if reference_path and reference_path not in sys.path:
        sys.path.insert(0, reference_path)
    _shim_imports()
    done = {}

    ref_attacker = importlib.import_module("attacker")
    ref_attacker.Attacker = _att.Attacker
    done[("attacker", "Attacker")] = _att.Attacker
    try:
        gt = importlib.import_module("gcn_trainer")
        gt.Attacker = _att.Attacker
        done[("gcn_trainer", "Attacker")] = _att.Attacker
    except Exception as e:                           # trainer not importable: the attack API still is
        done[("gcn_trainer", "Attacker")] = e

    if patch_layers:
        layers = importlib.import_module("gcn.layers")
        models = importlib.import_module("gcn.models")

        def gc_forward(self, input, adj):
            if not input.is_cuda:
                raise RuntimeError("linkteller_b200: GraphConvolution.forward needs CUDA tensors")
            return _gcn._GraphConvFn.apply(input, self.weight, self.bias, _gcn._csr_pair(adj))

        ref_gcn_forward = models.GCN.forward

        def gcn_forward(self, x, adj):
            if not self.training and not torch.is_grad_enabled() and x.is_cuda:
                return ops.gcn_forward(_gcn._csr_pair(adj)[0], x.to(torch.float32), self.gc1.weight.detach(),
                                       self.gc1.bias.detach(), self.gc2.weight.detach(), self.gc2.bias.detach())
            return ref_gcn_forward(self, x, adj)

        ref_gcn3_forward = getattr(getattr(models, "GCN3", None), "forward", None)

        def gcn3_forward(self, x, adj):
            if not self.training and not torch.is_grad_enabled() and x.is_cuda:
                return ops.gcn3_forward(_gcn._csr_pair(adj)[0], x.to(torch.float32), self.gc1.weight.detach(),
                                        self.gc1.bias.detach(), self.gc2.weight.detach(), self.gc2.bias.detach(),
                                        self.gc3.weight.detach(), self.gc3.bias.detach())
            return ref_gcn3_forward(self, x, adj)

        layers.GraphConvolution.forward = gc_forward
        models.GCN.forward = gcn_forward
        if ref_gcn3_forward is not None:
            models.GCN3.forward = gcn3_forward
            done[("gcn.models", "GCN3.forward")] = gcn3_forward
        done[("gcn.layers", "GraphConvolution.forward")] = gc_forward
        done[("gcn.models", "GCN.forward")] = gcn_forward

    if patch_normalization:
        utils = importlib.import_module("utils")
        ref_fetch = utils.fetch_normalization

        def fetch_normalization(type):
            if type in ops.NORM_KINDS:
                return _adj.fetch_normalization(type)
            return ref_fetch(type)                    # other normalisers are outside the hot path

        utils.fetch_normalization = fetch_normalization
        done[("utils", "fetch_normalization")] = fetch_normalization
        for name in ("worker", "utils.load"):
            try:
                m = importlib.import_module(name)
                if hasattr(m, "fetch_normalization"):
                    m.fetch_normalization = fetch_normalization
                    done[(name, "fetch_normalization")] = fetch_normalization
            except Exception:
                pass

    if patch_dp:
        try:
            worker = importlib.import_module("worker")

            def perturb_adj(self, adj, perturb_type):
                a = self.args
                csr = sp.csr_matrix(adj)
                csr.sort_indices()
                dcsr = ops.DeviceCSR.from_scipy(csr, "cuda")
                out = _dp.perturb_adj(dcsr, perturb_type, a.epsilon, a.noise_seed,
                                      getattr(a, "noise_type", "laplace"), getattr(a, "delta", 1e-5))
                m = out.to_scipy()
                if perturb_type != "discrete":
                    self.n_nodes = adj.shape[0]       # side effect of perturb_adj_continuous (worker.py:282)
                    m = m.astype("int32")             # worker.py:331: np.ones(..., dtype=np.int32)
                return m

            worker.Worker.perturb_adj = perturb_adj
            done[("worker", "Worker.perturb_adj")] = perturb_adj
        except Exception as e:
            done[("worker", "Worker.perturb_adj")] = e
    return done
