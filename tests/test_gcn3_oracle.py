"""SURVEY.md §8f row 3 on the CPU: the oracle's restatement of the three-layer model (--n-layer 3, GCN3)
under the efficient attack against what the unmodified reference produced
(tests/golden/tiny_gcn3.npz, sparse_gcn3.npz — written by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O

CASES = ["tiny_gcn3", "sparse_gcn3"]


def problem3(g):
    over = {str(k): int(v) for k, v in zip(g["override_keys"], g["override_vals"])}
    s = synth.make(str(g["shape"]), seed=int(g["seed"]), hidden=int(g["h1"]), **over)
    w = synth.gcn3_weights(s.f, int(g["h1"]), int(g["h2"]), s.c, seed=42)
    ip, ix, dv = O.csr_from_edges(s.edges, s.n)
    coo = O.normalize_adj(ip, ix, dv, s.n, "FirstOrderGCN")
    return s, w, (ip, ix), coo


@pytest.mark.parametrize("case", CASES)
def test_gcn3_logits_and_sampling(golden, case):
    g = golden(case)
    s, w, (ip, ix), coo = problem3(g)
    tn = O.sample_test_nodes("twitch/ES/RU", "unbalanced", ip, int(g["T"]), 42)
    assert np.array_equal(tn, g["test_nodes"])
    ex, ne = O.edge_sets_among(ip, ix, tn)
    assert np.array_equal(ex, g["exist"]) and np.array_equal(ne, g["nonexist"])
    logits = O.gcn_forward_np(s.features, coo, w, s.n, torch.float32)
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", CASES)
def test_gcn3_influence_matches_reference(golden, case):
    g = golden(case)
    s, w, (ip, ix), coo = problem3(g)
    tn = g["test_nodes"]
    infl = O.influence_matrix(s.features, coo, w, s.n, tn, 1e-4, torch.float64)
    np.testing.assert_allclose(infl, g["influence64"], rtol=1e-9, atol=1e-12)
    # exact zeros: outside three hops of Â's pattern, in float32 and float64 alike
    hip, hix, _ = O.coo_to_csr(*coo, s.n)
    sup = O.k_hop_support(hip, hix, tn, 3)
    assert not (g["influence64"] != 0)[~sup].any() and not (g["influence32"] != 0)[~sup].any()
    e, n = O.pair_scores(infl, tn, g["exist"], g["nonexist"])
    np.testing.assert_allclose(e, g["score_exist64"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(n, g["score_nonexist64"], rtol=1e-9, atol=1e-12)


def test_model_forward_rejects_other_depths():
    x = torch.zeros(2, 2)
    with pytest.raises(NotImplementedError):
        O.model_forward(x, None, *([x] * 8))
