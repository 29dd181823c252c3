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


def test_listed_pairs_equal_the_entries_of_the_influence_matrix(golden):
    """lt_pair_influence (one warp per listed pair intersecting N̂(p) ∩ N̂(o)) against the same entries read out of
    the full n x n influence matrix that lt_attack_efficient produces: same incidences, same ΔQ, same summation order."""
    from linkteller_b200 import ops
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_next")
    s, model, worker, args = make(g, "balanced-full", "efficient", 0)
    att = Attacker(args, model, worker)
    d = att._staged()
    u, v = att._pair_tensors(np.concatenate([g["bf_nonexist"], g["bf_exist"]]))
    listed = ops.pair_influence(d["hat"], d["x"], d["w1"], d["b1"], d["w2"], u, v, 1e-4)
    every = torch.arange(s.n, dtype=torch.int32, device="cuda")
    full = ops.attack_efficient(d["hat"], d["x"], d["w1"], d["b1"], d["w2"], every, 1e-4)
    want = full[u.long(), v.long()]
    assert torch.equal(listed == 0, want == 0)
    assert torch.allclose(listed, want, rtol=2e-6, atol=0)


@pytest.mark.parametrize("sample_type,mode", [("balanced-full", "efficient"), ("balanced-full", "naive")])
def test_listed_attacks_with_hidden_width_not_a_multiple_of_four(sample_type, mode, tmp_path, monkeypatch):
    """--hidden 10: the first hidden width is padded with units that never fire (ops._pad_first_hidden) on every path,
    the listed-pair ones included; checked against the float64 oracle."""
    import types
    from linkteller_b200 import adj as A, gcn
    from linkteller_b200.attacker import Attacker
    s = synth.make("tiny", seed=9, hidden=10)
    w = synth.gcn_weights(s.f, 10, s.c, seed=3)
    model = gcn.GCN(s.f, 10, s.c, 0.5)
    model.load_state_dict({k: torch.from_numpy(a) for k, a in zip(("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"), w)})
    model = model.cuda().eval()
    clean = A.adj_from_edges(s.edges, s.n)
    from linkteller_b200 import ops
    hat = ops.normalize_adj(clean, "FirstOrderGCN")
    feats = torch.from_numpy(s.features).cuda()
    worker = types.SimpleNamespace(n_nodes=s.n, n_features=s.f, features=feats, adj_full=hat, features_2=feats, adj_2=hat,
                                   adj_ori=clean.to_scipy())
    args = types.SimpleNamespace(dataset="cora", sample_type=sample_type, n_test=0, influence=1e-4, sample_seed=42,
                                 mode="vanilla-clean", attack_mode=mode, perturb_type=None, epsilon=0.0, noise_seed=0)
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    cap = {}
    att.compute_and_save = lambda e, ne, **k: cap.update(e=np.asarray(e), ne=np.asarray(ne))
    if mode == "efficient":
        att.link_prediction_attack_efficient_balanced()
    else:
        att.link_prediction_attack()
    ip = hat.indptr.cpu().numpy()
    coo = (np.repeat(np.arange(s.n, dtype=np.int64), np.diff(ip)), hat.indices.cpu().numpy().astype(np.int64), hat.values.cpu().numpy())
    fn = O.efficient_balanced_scores if mode == "efficient" else O.naive_scores
    e64, n64 = fn(s.features, coo, w, s.n, np.asarray(att.exist_edges), np.asarray(att.nonexist_edges), 1e-4)
    for got, ref in ((cap["e"], e64), (cap["ne"], n64)):
        assert np.array_equal(got == 0, ref == 0)
        nz = ref != 0
        assert (np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])).max() <= 1e-3


def test_balanced_full_at_flickr_size_finishes_in_seconds():
    """attacker.py:250-284 on the Flickr-shaped graph (n = 89 250, every node perturbed, 2 x 450 K listed pairs): the
    listed-pair path never forms rows over all observers.  Spot-checked against the float64 sub-problem oracle."""
    import time
    from linkteller_b200 import ops, workloads
    wl = workloads.make("flickr")
    ip = wl.adj_ori.indptr.cpu().numpy().astype(np.int64)
    ix = wl.adj_ori.indices.cpu().numpy().astype(np.int64)
    row = np.repeat(np.arange(wl.n, dtype=np.int64), np.diff(ip))
    up = ix > row
    edges = np.stack([row[up], ix[up]], axis=1)
    rng = np.random.default_rng(0)
    non = rng.integers(0, wl.n, size=(len(edges), 2))
    pairs = np.concatenate([edges, non])
    u = torch.from_numpy(pairs[:, 0].astype(np.int32)).cuda()
    v = torch.from_numpy(pairs[:, 1].astype(np.int32)).cuda()
    torch.cuda.synchronize()
    t0 = time.time()
    sc = ops.pair_influence(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, u, v, wl.delta)
    torch.cuda.synchronize()
    dt = time.time() - t0
    assert dt < 5.0, dt
    assert bool((sc[: len(edges)] > 0).all())             # an edge's endpoints are within one hop of each other
    # float64 check of 40 listed pairs: perturb p, observe o → LocalInfluence row of p read at o
    hat = wl.adj_hat
    pick = rng.choice(len(pairs), 40, replace=False)
    obs = np.unique(pairs[pick, 1])
    loc = O.LocalInfluence(hat.indptr.cpu().numpy(), hat.indices.cpu().numpy(), hat.values.cpu().numpy(), wl.x.cpu().numpy(),
                           [t.cpu().numpy() for t in (wl.w1, wl.b1, wl.w2)], obs)
    got = sc.cpu().numpy()
    for k in pick:
        ref = loc.row(int(pairs[k, 0]), wl.delta)[0][np.searchsorted(obs, pairs[k, 1])]
        assert (got[k] == 0) == (ref == 0)
        if ref != 0:
            assert abs(got[k] - ref) / abs(ref) <= 1e-3


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
