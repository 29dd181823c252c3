"""SURVEY.md §8f rows on the GPU, through the reference-facing Attacker: --sample-type balanced-full
(efficient and LSA2 baselines) and the naive per-pair attack, against the reference's own outputs
(tests/golden/tiny_next.npz).  Influence values: within 1e-3 relative of the reference run in float64,
exact-zero pattern identical (the bar of the main path)."""
import os
import types

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O

pytestmark = pytest.mark.gpu


def make(g, sample_type, attack_mode, n_test):
    from linkteller_b200 import adj as A, gcn
    s = synth.make(str(g["shape"]), seed=int(g["seed"]), hidden=int(g["hidden"]))
    w = synth.gcn_weights(s.f, int(g["hidden"]), s.c, seed=42)
    model = gcn.GCN(s.f, int(g["hidden"]), s.c, 0.5)
    model.load_state_dict({"gc1.weight": torch.from_numpy(w[0]), "gc1.bias": torch.from_numpy(w[1]),
                           "gc2.weight": torch.from_numpy(w[2]), "gc2.bias": torch.from_numpy(w[3])})
    model = model.cuda().eval()
    ip, ix, dv = O.csr_from_edges(s.edges, s.n)
    adj_ori = sp.csr_matrix((dv, ix, ip), shape=(s.n, s.n))
    r, c, v = O.normalize_adj(ip, ix, dv, s.n, "FirstOrderGCN")
    adj_t = A.sparse_mx_to_torch_sparse_tensor(sp.coo_matrix((v, (r, c)), shape=(s.n, s.n))).cuda()
    feats = torch.from_numpy(s.features).cuda()
    worker = types.SimpleNamespace(n_nodes=s.n, n_features=s.f, features=feats, adj_full=adj_t, features_2=feats,
                                   adj_2=adj_t, adj_ori=adj_ori)
    args = types.SimpleNamespace(dataset="twitch/ES/RU", sample_type=sample_type, n_test=n_test, influence=1e-4,
                                 sample_seed=42, mode="vanilla-clean", attack_mode=attack_mode, perturb_type=None,
                                 epsilon=5.0, noise_seed=42)
    return s, model, worker, args


def close_to_ref64(got, ref64, ref32):
    got, ref64 = np.asarray(got, dtype=np.float64), np.asarray(ref64)
    assert np.array_equal(got == 0, ref64 == 0), "exact-zero pattern differs from the reference"
    nz = ref64 != 0
    rel = np.abs(got[nz] - ref64[nz]) / np.abs(ref64[nz])
    assert rel.max() <= 1e-3, rel.max()
    rel32 = np.abs(np.asarray(ref32)[nz] - ref64[nz]) / np.abs(ref64[nz])
    assert rel.max() <= max(rel32.max(), 1e-5)         # never further from ref64 than the reference's own fp32 run


def test_balanced_full_efficient(golden, tmp_path, monkeypatch):
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_next")
    s, model, worker, args = make(g, "balanced-full", "efficient", 0)
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    assert args.n_test == s.n                          # attacker.py:19-20
    att.prepare_test_data()
    assert np.array_equal(np.asarray(att.exist_edges), g["bf_exist"])
    assert np.array_equal(np.asarray(att.nonexist_edges), g["bf_nonexist"])
    assert list(att.test_nodes) == list(g["bf_test_nodes"])
    cap = {}
    att.compute_and_save = lambda e, ne, **k: cap.update(e=np.asarray(e), ne=np.asarray(ne))
    att.link_prediction_attack_efficient_balanced()
    close_to_ref64(cap["e"], g["bf_eff_exist64"], g["bf_eff_exist32"])
    close_to_ref64(cap["ne"], g["bf_eff_nonexist64"], g["bf_eff_nonexist32"])


def test_balanced_full_blocks_agree(golden):
    """pair_influence in several row blocks = in one block (block_bytes only sizes the scratch)."""
    from linkteller_b200 import ops
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_next")
    s, model, worker, args = make(g, "balanced-full", "efficient", 0)
    att = Attacker(args, model, worker)
    d = att._staged()
    u, v = att._pair_tensors(g["bf_nonexist"])
    one = ops.pair_influence(d["hat"], d["x"], d["w1"], d["b1"], d["w2"], u, v, 1e-4)
    many = ops.pair_influence(d["hat"], d["x"], d["w1"], d["b1"], d["w2"], u, v, 1e-4, block_bytes=37 * 4 * s.n)
    assert torch.equal(one, many)


@pytest.mark.parametrize("mode,key", [("baseline", "post"), ("baseline-feat", "attr")])
def test_balanced_full_baselines(golden, mode, key, tmp_path, monkeypatch):
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_next")
    s, model, worker, args = make(g, "balanced-full", mode, 0)
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    cap = {}
    att.compute_and_save = lambda e, ne, **k: cap.update(e=np.asarray(e), ne=np.asarray(ne))
    att.baseline_attack_balanced()
    np.testing.assert_allclose(cap["e"], g[f"bf_{key}_exist"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(cap["ne"], g[f"bf_{key}_nonexist"], rtol=2e-4, atol=2e-5)


def test_naive_attack(golden, tmp_path, monkeypatch, capsys):
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_next")
    s, model, worker, args = make(g, "unbalanced", "naive", int(g["T_naive"]))
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    assert np.array_equal(att.test_nodes, g["naive_test_nodes"])
    fn = att.link_prediction_attack()
    assert os.path.normpath(fn) == os.path.normpath(str(g["naive_filename"]))     # no attack_mode prefix (attacker.py:180)
    saved = torch.load(fn, weights_only=False)
    assert np.array_equal(np.asarray(saved["result"]["y"]), g["naive_y"])
    close_to_ref64(saved["result"]["pred"], g["naive_pred64"], g["naive_pred32"])
    out = capsys.readouterr().out
    assert "time for predicting existing edges" in out and "auc =" in out


def test_naive_attack_on_balanced_full_pairs(golden):
    """naive on listed pairs = efficient-balanced with the roles of u and v swapped."""
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_next")
    s, model, worker, args = make(g, "balanced-full", "naive", 0)
    att = Attacker(args, model, worker)
    sub = g["bf_exist"][:200]
    a = att._listed_influence(sub, perturb_first=False)
    b = att._listed_influence(sub[:, ::-1].copy(), perturb_first=True)
    assert np.array_equal(a, b)
