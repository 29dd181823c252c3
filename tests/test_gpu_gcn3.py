"""SURVEY.md §8f row 3 on the GPU: --n-layer 3 (GCN3) through lt_gcn3_forward / lt_attack_efficient3 and the
reference-facing Attacker, against the reference's own outputs (tests/golden/tiny_gcn3.npz, sparse_gcn3.npz).
Bar of the main path: exact-zero pattern identical, non-zeros within 1e-3 relative of the reference run in
float64 and never further from it than the reference's own float32 run."""
import types

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O

pytestmark = pytest.mark.gpu

CASES = ["tiny_gcn3", "sparse_gcn3"]
KEYS = ("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias", "gc3.weight", "gc3.bias")


def problem3(g):
    over = {str(k): int(v) for k, v in zip(g["override_keys"], g["override_vals"])}
    s = synth.make(str(g["shape"]), seed=int(g["seed"]), hidden=int(g["h1"]), **over)
    w = synth.gcn3_weights(s.f, int(g["h1"]), int(g["h2"]), s.c, seed=42)
    ip, ix, dv = O.csr_from_edges(s.edges, s.n)
    coo = O.normalize_adj(ip, ix, dv, s.n, "FirstOrderGCN")
    return s, w, (ip, ix, dv), coo


def device_problem(g):
    from linkteller_b200 import ops
    s, w, (ip, ix, dv), coo = problem3(g)
    hip, hix, hv = O.coo_to_csr(*coo, s.n)
    hat = ops.DeviceCSR.from_scipy(sp.csr_matrix((hv, hix, hip), shape=(s.n, s.n)), "cuda")
    x = torch.from_numpy(s.features).cuda()
    wd = [torch.from_numpy(a).cuda() for a in w]
    tn = torch.from_numpy(g["test_nodes"].astype(np.int32)).cuda()
    return s, w, coo, hat, x, wd, tn


def close_to_ref64(got, ref64, ref32=None, tol=1e-3):
    got, ref64 = np.asarray(got, dtype=np.float64), np.asarray(ref64)
    assert np.array_equal(got == 0, ref64 == 0), "exact-zero pattern differs from the reference"
    nz = ref64 != 0
    rel = np.abs(got[nz] - ref64[nz]) / np.abs(ref64[nz])
    assert rel.max() <= tol, rel.max()
    if ref32 is not None:
        rel32 = np.abs(np.asarray(ref32)[nz] - ref64[nz]) / np.abs(ref64[nz])
        assert rel.max() <= max(rel32.max(), 1e-5)
    return rel.max()


@pytest.mark.parametrize("case", CASES)
def test_gcn3_forward_matches_reference(golden, case):
    from linkteller_b200 import ops
    g = golden(case)
    s, w, coo, hat, x, wd, tn = device_problem(g)
    z = ops.gcn3_forward(hat, x, *wd).cpu().numpy()
    ref = O.gcn_forward_np(s.features, coo, w, s.n, torch.float64)
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(z - ref).max() <= 1e-4 * scale
    assert np.abs(z - g["logits"]).max() <= 2e-4 * scale


@pytest.mark.parametrize("case", CASES)
def test_gcn3_influence_matches_reference(golden, case):
    from linkteller_b200 import ops
    g = golden(case)
    s, w, coo, hat, x, wd, tn = device_problem(g)
    out, info = ops.attack_efficient3(hat, x, *wd[:5], tn, 1e-4, want_info=True)
    got = out.cpu().numpy()
    close_to_ref64(got, g["influence64"], g["influence32"])
    assert info["chunks"] == 1 and info["support_pairs"] == int((g["influence64"] != 0).sum())
    # chunked (a handful of perturbed nodes per chunk) and sliced runs give the same bits
    cap = max(1, info["hop2_records"] // 7)
    out2, info2 = ops.attack_efficient3(hat, x, *wd[:5], tn, 1e-4, want_info=True, max_chunk_records=cap)
    assert info2["chunks"] > 3
    assert torch.equal(out, out2)
    T = tn.numel()
    lo, hi = T // 3, T // 3 + 5
    part = ops.attack_efficient3(hat, x, *wd[:5], tn, 1e-4, lo=lo, hi=hi)
    assert torch.equal(part, out[lo:hi])


def test_gcn3_matches_oracle_on_unpadded_widths():
    """hidden widths that are not multiples of 4 (the float4 paths need padding) against the oracle in float64."""
    from linkteller_b200 import ops
    s = synth.make("tiny", seed=11, hidden=10)
    w = synth.gcn3_weights(s.f, 10, 7, s.c, seed=3)
    ip, ix, dv = O.csr_from_edges(s.edges, s.n)
    coo = O.normalize_adj(ip, ix, dv, s.n, "FirstOrderGCN")
    hip, hix, hv = O.coo_to_csr(*coo, s.n)
    hat = ops.DeviceCSR.from_scipy(sp.csr_matrix((hv, hix, hip), shape=(s.n, s.n)), "cuda")
    x = torch.from_numpy(s.features).cuda()
    wd = [torch.from_numpy(a).cuda() for a in w]
    tn_np = O.sample_test_nodes("twitch/ES/RU", "unbalanced", ip, 30, 42)
    tn = torch.from_numpy(tn_np.astype(np.int32)).cuda()
    got = ops.attack_rows(hat, x, wd, tn, 1e-4).cpu().numpy()
    ref = O.influence_matrix(s.features, coo, w, s.n, tn_np, 1e-4, torch.float64)
    close_to_ref64(got, ref)
    z = ops.model_forward(hat, x, wd).cpu().numpy()
    ref_z = O.gcn_forward_np(s.features, coo, w, s.n, torch.float64)
    assert np.abs(z - ref_z).max() <= 1e-4 * max(1.0, np.abs(ref_z).max())


def test_gcn3_medium_graph_tensor_core_stage():
    """A PPI-sized graph with wide hidden layers: the first-hop product runs on tcgen05 (E·h1·h2 ≥ 2²⁴) and
    several chunks are needed; checked against the oracle in float64 on a sample of rows."""
    from linkteller_b200 import ops
    s = synth.make("ppi", seed=0, hidden=128)
    w = synth.gcn3_weights(s.f, 128, 64, s.c, seed=5)
    ip, ix, dv = O.csr_from_edges(s.edges, s.n)
    coo = O.normalize_adj(ip, ix, dv, s.n, "FirstOrderGCN")
    hip, hix, hv = O.coo_to_csr(*coo, s.n)
    hat = ops.DeviceCSR.from_scipy(sp.csr_matrix((hv, hix, hip), shape=(s.n, s.n)), "cuda")
    x = torch.from_numpy(s.features).cuda()
    wd = [torch.from_numpy(a).cuda() for a in w]
    tn_np = O.sample_test_nodes("ppi", "unbalanced", ip, 300, 42)
    tn = torch.from_numpy(tn_np.astype(np.int32)).cuda()
    out, info = ops.attack_efficient3(hat, x, *wd[:5], tn, 1e-4, want_info=True)
    assert info["chunks"] == 1 and info["hop1_incidences"] * 128 * 64 >= 1 << 24
    rows = [0, 7, 150, 299]
    ref = O.influence_matrix(s.features, coo, w, s.n, tn_np, 1e-4, torch.float64, rows=rows)
    close_to_ref64(out.cpu().numpy()[rows], ref)
    out2, info2 = ops.attack_efficient3(hat, x, *wd[:5], tn, 1e-4, want_info=True,
                                        max_chunk_records=info["hop2_records"] // 3)
    assert info2["chunks"] >= 3 and info2["hop2_records"] == info["hop2_records"]
    close_to_ref64(out2.cpu().numpy()[rows], ref)
    assert torch.equal(out == 0, out2 == 0)


def test_attacker_gcn3_end_to_end(golden, tmp_path, monkeypatch):
    """The reference-facing Attacker with a GCN3 model: same sampled nodes and pair lists, pair scores within
    1e-3 of the reference's float64 run, AUC/AP equal to sklearn on the same scores."""
    from linkteller_b200 import adj as A, gcn
    from linkteller_b200.attacker import Attacker
    g = golden("sparse_gcn3")
    s, w, (ip, ix, dv), coo = problem3(g)
    model = gcn.GCN3(s.f, int(g["h1"]), int(g["h2"]), s.c, 0.5)
    model.load_state_dict({k: torch.from_numpy(a) for k, a in zip(KEYS, w)})
    model = model.cuda().eval()
    adj_ori = sp.csr_matrix((dv, ix, ip), shape=(s.n, s.n))
    r, c, v = coo
    adj_t = A.sparse_mx_to_torch_sparse_tensor(sp.coo_matrix((v, (r, c)), shape=(s.n, s.n))).cuda()
    feats = torch.from_numpy(s.features).cuda()
    worker = types.SimpleNamespace(n_nodes=s.n, n_features=s.f, features=feats, adj_full=adj_t, features_2=feats,
                                   adj_2=adj_t, adj_ori=adj_ori)
    args = types.SimpleNamespace(dataset="twitch/ES/RU", sample_type="unbalanced", n_test=int(g["T"]), influence=1e-4,
                                 sample_seed=42, mode="vanilla-clean", attack_mode="efficient", perturb_type=None,
                                 epsilon=5.0, noise_seed=42, n_layer=3)
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    assert np.array_equal(np.asarray(att.test_nodes), g["test_nodes"])
    assert np.array_equal(np.asarray(att.exist_edges), g["exist"])
    assert np.array_equal(np.asarray(att.nonexist_edges), g["nonexist"])
    cap = {}
    real = att.compute_and_save
    att.compute_and_save = lambda e, ne: (cap.update(e=np.asarray(e), ne=np.asarray(ne)), real(e, ne))[1]
    att.link_prediction_attack_efficient()
    close_to_ref64(cap["e"], g["score_exist64"])
    close_to_ref64(cap["ne"], g["score_nonexist64"])
    y = [1] * len(cap["e"]) + [0] * len(cap["ne"])
    auc0, ap0 = O.auc_ap(y, np.r_[cap["e"], cap["ne"]].astype(np.float32))
    assert abs(att.auc - auc0) < 1e-9 and abs(att.ap - ap0) < 1e-9
    # get_gradient_eps_mat (two forwards, attacker.py:100-108) agrees with the influence row it defines
    row = att.get_gradient_eps_mat(int(g["test_nodes"][3])).double().cpu().numpy()
    norms = np.linalg.norm(row[g["test_nodes"]], axis=1)
    infl = att.influence_val.cpu().numpy()[3]
    nz = infl != 0
    assert np.all(norms[~nz] == 0)
    with torch.no_grad():
        z = model(feats, adj_t).cpu().numpy()
    assert np.abs(z - g["logits"]).max() <= 2e-4 * max(1.0, np.abs(g["logits"]).max())


def test_attacker_gcn3_balanced_full_and_naive(golden, tmp_path, monkeypatch):
    """GCN3 with --sample-type balanced-full (attacker.py:250-284) and the naive per-pair attack (:143-206): listed
    pairs read from blocks of three-hop influence rows, against the oracle in float64."""
    from linkteller_b200 import adj as A, gcn
    from linkteller_b200.attacker import Attacker
    g = golden("tiny_gcn3")
    s, w, (ip, ix, dv), coo = problem3(g)
    model = gcn.GCN3(s.f, int(g["h1"]), int(g["h2"]), s.c, 0.5)
    model.load_state_dict({k: torch.from_numpy(a) for k, a in zip(KEYS, w)})
    model = model.cuda().eval()
    adj_ori = sp.csr_matrix((dv, ix, ip), shape=(s.n, s.n))
    r, c, v = coo
    adj_t = A.sparse_mx_to_torch_sparse_tensor(sp.coo_matrix((v, (r, c)), shape=(s.n, s.n))).cuda()
    feats = torch.from_numpy(s.features).cuda()
    worker = types.SimpleNamespace(n_nodes=s.n, n_features=s.f, features=feats, adj_full=adj_t, features_2=feats,
                                   adj_2=adj_t, adj_ori=adj_ori)
    args = types.SimpleNamespace(dataset="twitch/ES/RU", sample_type="balanced-full", n_test=0, influence=1e-4,
                                 sample_seed=42, mode="vanilla-clean", attack_mode="efficient", perturb_type=None,
                                 epsilon=5.0, noise_seed=42, n_layer=3)
    monkeypatch.chdir(tmp_path)
    att = Attacker(args, model, worker)
    att.prepare_test_data()
    ex, ne = np.asarray(att.exist_edges), np.asarray(att.nonexist_edges)
    cap = {}
    att.compute_and_save = lambda e, n_, **k: cap.update(e=np.asarray(e), ne=np.asarray(n_))
    att.link_prediction_attack_efficient_balanced()
    e64, n64 = O.efficient_balanced_scores(s.features, coo, w, s.n, ex, ne, 1e-4, torch.float64)
    close_to_ref64(cap["e"], e64)
    close_to_ref64(cap["ne"], n64)
    att.link_prediction_attack()                                   # naive: perturb the SECOND node of a pair
    e64, n64 = O.naive_scores(s.features, coo, w, s.n, ex, ne, 1e-4, torch.float64)
    close_to_ref64(cap["e"], e64)
    close_to_ref64(cap["ne"], n64)
