"""The reference-facing layer on the GPU: Attacker / GCN / GraphConvolution with the reference's
interface, and the rebinding of the reference's names (patch.install) exercised on stand-in modules
that have the reference's module layout (the real checkout does not exist on the GPU box)."""
import os
import sys
import types

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O

pytestmark = pytest.mark.gpu


def make_problem(g, device="cuda"):
    from linkteller_b200 import adj as A, gcn
    s = synth.make(str(g["shape"]), seed=int(g["seed"]), hidden=int(g["hidden"]))
    w = synth.gcn_weights(s.f, int(g["hidden"]), s.c, seed=42)
    model = gcn.GCN(s.f, int(g["hidden"]), s.c, 0.5)
    model.load_state_dict({"gc1.weight": torch.from_numpy(w[0]), "gc1.bias": torch.from_numpy(w[1]),
                           "gc2.weight": torch.from_numpy(w[2]), "gc2.bias": torch.from_numpy(w[3])})
    model = model.to(device).eval()
    adj_ori = sp.csr_matrix((g["a1_data"], g["a1_indices"], g["a1_indptr"]), shape=(s.n, s.n))
    hat = sp.coo_matrix((g["a2_val"], (g["a2_row"], g["a2_col"])), shape=(s.n, s.n))
    adj_t = A.sparse_mx_to_torch_sparse_tensor(hat).to(device)
    feats = torch.from_numpy(s.features).to(device)
    worker = types.SimpleNamespace(n_nodes=s.n, n_features=s.f, features=feats, adj_full=adj_t, features_2=feats,
                                   adj_2=adj_t, adj_ori=adj_ori)
    args = types.SimpleNamespace(dataset=str(g["dataset"]), sample_type=str(g["sample_type"]), n_test=int(g["T"]),
                                 influence=1e-4, sample_seed=42, mode="vanilla-clean", attack_mode="efficient",
                                 perturb_type=str(g["dp"]) or None, epsilon=float(g["epsilon"]), noise_seed=42)
    if str(g["dp"]):
        args.mode = "vanilla"
    return s, w, model, worker, args


@pytest.mark.parametrize("case", ["tiny_fo", "small_hi", "tiny_lapgraph"])
def test_attacker_efficient_end_to_end(golden, case, tmp_path, monkeypatch, capsys):
    from linkteller_b200.attacker import Attacker
    g = golden(case)
    s, w, model, worker, args = make_problem(g)
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    assert np.array_equal(att.test_nodes, g["a6_test_nodes"])
    assert np.array_equal(np.asarray(att.exist_edges), g["a6_exist"])
    assert np.array_equal(np.asarray(att.nonexist_edges), g["a6_nonexist"])
    att.link_prediction_attack_efficient()
    out = capsys.readouterr().out
    assert "auc =" in out and "attack results saved to:" in out
    # file name and schema exactly as the reference writes them (attacker.py:388-411)
    fn = str(g["a9_filename"])          # the golden run used mode vanilla-clean (attacker.py:391-392)
    if args.mode != "vanilla-clean":    # DP naming, attacker.py:393-394
        fn = (f"eval_{args.dataset}/efficient_{args.sample_type}_{args.perturb_type}_{args.n_test}_42"
              f"_eps-{args.epsilon}_seed-42.pt")
    assert os.path.exists(fn), os.listdir(".")
    saved = torch.load(fn, weights_only=False)
    assert set(saved) == {"auc", "pr", "result"} and set(saved["auc"]) == {"fpr", "tpr", "thresholds"}
    assert set(saved["pr"]) == {"precision", "recall", "thresholds"} and set(saved["result"]) == {"y", "pred"}
    assert np.array_equal(np.asarray(saved["result"]["y"]), g["a9_y"])
    # scores: within 1e-3 relative of the reference run in float64, exact zeros identical
    pred = np.asarray(saved["result"]["pred"])
    ref = np.concatenate([g["a8_exist64"], g["a8_nonexist64"]])
    assert np.array_equal(pred == 0, ref == 0)
    nz = ref != 0
    assert (np.abs(pred[nz] - ref[nz]) / ref[nz]).max() <= 1e-3
    # AUC / AP: equal to the float64 reference's up to near-ties (1e-3 score perturbation moves few ranks)
    auc64, ap64 = O.auc_ap(g["a9_y"], ref)
    assert abs(att.auc - auc64) <= 2e-4 and abs(att.ap - ap64) <= 2e-3
    auc32, _ = O.auc_ap(g["a9_y"], np.concatenate([g["a8_exist32"], g["a8_nonexist32"]]))
    assert abs(att.auc - auc64) <= abs(auc32 - auc64) + 1e-6      # no further from ref64 than the reference's fp32 run


def test_compute_and_save_reproduces_reference_curves(golden, tmp_path, monkeypatch):
    # fed the reference's own fp32 scores, the saved curves equal sklearn's on the reference run
    from linkteller_b200.attacker import Attacker
    g = golden("small_hi")
    s, w, model, worker, args = make_problem(g)
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    fn = att.compute_and_save(list(g["a8_exist32"]), list(g["a8_nonexist32"]))
    saved = torch.load(fn, weights_only=False)
    assert np.array_equal(saved["auc"]["fpr"], g["a9_fpr"]) and np.array_equal(saved["auc"]["tpr"], g["a9_tpr"])
    assert abs(att.auc - float(g["a9_auc"])) < 1e-12 and abs(att.ap - float(g["a9_ap"])) < 1e-12


@pytest.mark.parametrize("case,mode", [("tiny_fo", "baseline"), ("small_hi", "baseline"), ("small_hi", "baseline-feat")])
def test_attacker_baselines(golden, case, mode, tmp_path, monkeypatch):
    from linkteller_b200.attacker import Attacker
    g = golden(case)
    s, w, model, worker, args = make_problem(g)
    args.attack_mode = mode
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    att.baseline_attack()
    fn = f"eval_{args.dataset}/{mode}_{args.sample_type}_{args.n_test}_42.pt"
    saved = torch.load(fn, weights_only=False)
    key = "post" if mode == "baseline" else "attr"
    ref = np.concatenate([g[f"a10_{key}_exist"], g[f"a10_{key}_nonexist"]])
    # float32 cosine similarities in a different summation order: 2e-5 relative, 2e-6 absolute
    np.testing.assert_allclose(np.asarray(saved["result"]["pred"]), ref, rtol=2e-5, atol=2e-6)


def test_attacker_errors_match_reference_behaviour(golden):
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_fo")
    s, w, model, worker, args = make_problem(g)
    args.sample_type = "nonsense"
    with pytest.raises(NotImplementedError):
        Attacker(args, model, worker).prepare_test_data()
    args.sample_type = "unbalanced"
    args.attack_mode = "nonsense"
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    with pytest.raises(NotImplementedError):
        att.baseline_attack()


def test_get_gradient_eps_mat_matches_influence_row(golden):
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_fo")
    s, w, model, worker, args = make_problem(g)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    infl = att.influence_matrix().cpu().numpy()
    tn = att.test_nodes
    grad = att.get_gradient_eps_mat(int(tn[3])).cpu().numpy()
    row = np.linalg.norm(grad[tn], axis=1)
    # the two-forward finite difference in fp32 is the noisy one (errors of 3-10 % are typical,
    # far more on small entries): zeros exact, entries above 10 % of the row maximum within 5 %
    assert np.array_equal(row == 0, infl[3] == 0)
    nz = infl[3] >= 0.1 * infl[3].max()
    assert (np.abs(row[nz] - infl[3][nz]) / infl[3][nz]).max() < 0.05


def test_gcn_modules_forward_backward_match_torch(golden):
    from linkteller_b200 import gcn
    g = golden("tiny_fo")
    s, w, model, worker, args = make_problem(g)
    x = worker.features.clone()
    adj = worker.adj_full
    # forward (eval, no_grad → fused lt_gcn_forward) vs the reference's logits
    with torch.no_grad():
        z = model(x, adj)
    assert np.abs(z.cpu().numpy() - g["a5_logits"]).max() <= 1e-4 * max(1.0, np.abs(g["a5_logits"]).max())
    # training path: autograd through the custom layers vs plain torch ops in float64 on the CPU
    model.train()
    model.dropout = 0.0
    xg = x.clone().requires_grad_(True)
    out = model(xg, adj)
    loss = (out * torch.linspace(-1, 1, out.numel(), device=out.device).reshape(out.shape)).sum()
    loss.backward()
    W1, b1, W2, b2 = [torch.from_numpy(a).double().requires_grad_(True) for a in w]
    A = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([g["a2_row"], g["a2_col"]])),
                                torch.from_numpy(g["a2_val"]).double(), (s.n, s.n))
    x0 = torch.from_numpy(s.features).double().requires_grad_(True)
    h = torch.relu(torch.spmm(A, x0 @ W1) + b1)
    o = torch.spmm(A, h @ W2) + b2
    (o * torch.linspace(-1, 1, o.numel(), dtype=torch.float64).reshape(o.shape)).sum().backward()
    for got, ref in ((model.gc1.weight.grad, W1.grad), (model.gc1.bias.grad, b1.grad), (model.gc2.weight.grad, W2.grad),
                     (model.gc2.bias.grad, b2.grad), (xg.grad, x0.grad)):
        got, ref = got.cpu().double(), ref
        assert (got - ref).abs().max() <= 1e-4 * max(1.0, ref.abs().max())      # fp32 kernels vs fp64 autograd
    # state-dict keys are the reference's
    assert list(model.state_dict()) == ["gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"]


def test_fetch_normalization_drop_in_is_bit_exact(golden):
    from linkteller_b200 import adj as A
    g = golden("tiny_aug")
    s = synth.make(str(g["shape"]), seed=int(g["seed"]))
    ori = sp.csr_matrix((g["a1_data"], g["a1_indices"], g["a1_indptr"]), shape=(s.n, s.n))
    coo = A.fetch_normalization("AugNormAdj")(ori)
    t = A.sparse_mx_to_torch_sparse_tensor(coo)
    assert np.array_equal(t._indices().numpy(), np.vstack([g["a2_row"], g["a2_col"]]))
    assert np.array_equal(t._values().numpy().view(np.uint32), g["a2_val"].view(np.uint32))
    with pytest.raises(NotImplementedError):
        A.fetch_normalization("RWalk")


REF_STUBS = {
    "attacker.py": "class Attacker:\n    pass\n",
    "gcn_trainer.py": "from attacker import Attacker\n",
    "gcn/__init__.py": "",
    "gcn/layers.py": ("import torch\nfrom torch.nn.parameter import Parameter\n"
                      "class GraphConvolution(torch.nn.Module):\n"
                      "    def __init__(self, i, o, bias=True):\n        super().__init__()\n"
                      "        self.weight = Parameter(torch.zeros(i, o)); self.bias = Parameter(torch.zeros(o))\n"
                      "    def forward(self, input, adj):\n        raise RuntimeError('reference path')\n"),
    "gcn/models.py": ("import torch\nfrom .layers import GraphConvolution\n"
                      "class GCN(torch.nn.Module):\n"
                      "    def __init__(self, f, h, c, dropout):\n        super().__init__()\n"
                      "        self.gc1 = GraphConvolution(f, h); self.gc2 = GraphConvolution(h, c); self.dropout = dropout\n"
                      "    def forward(self, x, adj):\n        raise RuntimeError('reference path')\n"
                      "class GCN3(torch.nn.Module):\n"
                      "    def __init__(self, f, h1, h2, c, dropout):\n        super().__init__()\n"
                      "        self.gc1 = GraphConvolution(f, h1); self.gc2 = GraphConvolution(h1, h2)\n"
                      "        self.gc3 = GraphConvolution(h2, c); self.dropout = dropout\n"
                      "    def forward(self, x, adj):\n        raise RuntimeError('reference path')\n"),
    "utils/__init__.py": "from .load import *\n",
    "utils/load.py": "def fetch_normalization(type):\n    return lambda adj: 'reference path'\n",
    "worker.py": ("from utils import fetch_normalization\n"
                  "class Worker:\n    def perturb_adj(self, adj, perturb_type):\n        raise RuntimeError('reference path')\n"),
}


def test_patch_install_rebinds_reference_names(golden, tmp_path, monkeypatch):
    """Stand-in modules with the reference's module/attribute layout (not its code): install() must
    rebind every hot-path name, and the rebound names must produce the golden results."""
    for rel, src in REF_STUBS.items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(src)
    for m in ("attacker", "gcn_trainer", "gcn", "gcn.layers", "gcn.models", "utils", "utils.load", "worker"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    monkeypatch.syspath_prepend(str(tmp_path))
    from linkteller_b200 import attacker as ours, patch
    done = patch.install()
    import attacker, gcn_trainer, utils, worker
    from gcn.models import GCN
    assert attacker.Attacker is ours.Attacker and gcn_trainer.Attacker is ours.Attacker
    g = golden("tiny_fo")
    s, w, _, wk, args = make_problem(g)
    model = GCN(s.f, int(g["hidden"]), s.c, 0.5)
    model.load_state_dict({"gc1.weight": torch.from_numpy(w[0]), "gc1.bias": torch.from_numpy(w[1]),
                           "gc2.weight": torch.from_numpy(w[2]), "gc2.bias": torch.from_numpy(w[3])})
    model = model.cuda().eval()
    with torch.no_grad():
        z = model(wk.features, wk.adj_full)                     # would raise 'reference path' if not rebound
    assert np.abs(z.cpu().numpy() - g["a5_logits"]).max() <= 1e-4 * max(1.0, np.abs(g["a5_logits"]).max())
    from gcn.models import GCN3                                 # --n-layer 3 (gcn_trainer.py:81-86)
    w3 = synth.gcn3_weights(s.f, 16, 8, s.c, seed=1)
    m3 = GCN3(s.f, 16, 8, s.c, 0.5)
    m3.load_state_dict({k: torch.from_numpy(a) for k, a in zip(
        ("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias", "gc3.weight", "gc3.bias"), w3)})
    m3 = m3.cuda().eval()
    with torch.no_grad():
        z3 = m3(wk.features, wk.adj_full)
    coo0 = (g["a2_row"], g["a2_col"], g["a2_val"])
    ref3 = O.gcn_forward_np(s.features, coo0, w3, s.n, torch.float64)
    assert np.abs(z3.cpu().numpy() - ref3).max() <= 1e-4 * max(1.0, np.abs(ref3).max())
    ori = sp.csr_matrix((g["a1_data"], g["a1_indices"], g["a1_indptr"]), shape=(s.n, s.n))
    coo = worker.fetch_normalization("FirstOrderGCN")(ori)
    assert np.array_equal(coo.row, g["a2_row"]) and np.array_equal(coo.data.astype(np.float32), g["a2_val"])
    assert utils.fetch_normalization("RWalk")(ori) == "reference path"   # untouched normalisers fall through
    ge = golden("tiny_edgerand")
    wobj = worker.Worker()
    wobj.args = types.SimpleNamespace(epsilon=float(ge["epsilon"]), noise_seed=42, noise_type="laplace", delta=1e-5)
    se = synth.make(str(ge["shape"]), seed=int(ge["seed"]))
    ori_e = sp.csr_matrix((ge["a1_data"], ge["a1_indices"], ge["a1_indptr"]), shape=(se.n, se.n))
    noisy = wobj.perturb_adj(ori_e, "discrete").tocsr()
    assert np.array_equal(noisy.indptr, ge["dp_indptr"]) and np.array_equal(noisy.data, ge["dp_data"])
    assert ("worker", "Worker.perturb_adj") in done
