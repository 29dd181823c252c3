"""SURVEY.md §8f rows on the CPU: the oracle's restatement of --sample-type balanced-full and of the naive
per-pair attack against what the unmodified reference produced (tests/golden/tiny_next.npz)."""
import numpy as np
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O


def problem(g):
    s = synth.make(str(g["shape"]), seed=int(g["seed"]), hidden=int(g["hidden"]))
    w = synth.gcn_weights(s.f, int(g["hidden"]), s.c, seed=42)
    ip, ix, dv = O.csr_from_edges(s.edges, s.n)
    coo = O.normalize_adj(ip, ix, dv, s.n, "FirstOrderGCN")
    return s, w, (ip, ix), coo


def test_balanced_edge_sets_match_reference(golden):
    g = golden("tiny_next")
    s, _, (ip, ix), _ = problem(g)
    ex, ne, tn = O.balanced_edge_sets(ip, ix, 42)
    assert np.array_equal(ex, g["bf_exist"]) and np.array_equal(ne, g["bf_nonexist"])
    assert np.array_equal(tn, g["bf_test_nodes"])
    assert int(g["bf_n_test"]) == s.n                  # Attacker.__init__ overrides n_test (attacker.py:19-20)


def test_efficient_balanced_scores(golden):
    g = golden("tiny_next")
    s, w, _, coo = problem(g)
    e64, n64 = O.efficient_balanced_scores(s.features, coo, w, s.n, g["bf_exist"], g["bf_nonexist"], 1e-4, torch.float64)
    np.testing.assert_allclose(e64, g["bf_eff_exist64"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(n64, g["bf_eff_nonexist64"], rtol=1e-9, atol=1e-12)
    # the reference's own float32 run has the same exact-zero pattern as its float64 run
    assert np.array_equal(g["bf_eff_nonexist32"] == 0, g["bf_eff_nonexist64"] == 0)
    assert np.array_equal(e64 == 0, g["bf_eff_exist32"] == 0)


def test_naive_scores(golden):
    g = golden("tiny_next")
    s, w, (ip, ix), coo = problem(g)
    tn = O.sample_test_nodes("twitch/ES/RU", "unbalanced", ip, int(g["T_naive"]), 42)
    assert np.array_equal(tn, g["naive_test_nodes"])
    ex, ne = O.edge_sets_among(ip, ix, tn)
    assert np.array_equal(ex, g["naive_exist"]) and np.array_equal(ne, g["naive_nonexist"])
    e64, n64 = O.naive_scores(s.features, coo, w, s.n, ex, ne, 1e-4, torch.float64)
    np.testing.assert_allclose(np.r_[e64, n64], g["naive_pred64"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(g["naive_y"], np.r_[np.ones(len(ex)), np.zeros(len(ne))])
    # naive and efficient read the same entries: score(u,v) = influence[idx v][idx u]
    infl = O.influence_matrix(s.features, coo, w, s.n, tn, 1e-4, torch.float64)
    e2, n2 = O.pair_scores(infl, tn, ex, ne)
    np.testing.assert_allclose(np.r_[e2, n2], g["naive_pred64"], rtol=1e-9, atol=1e-12)


def test_lsa2_balanced(golden):
    g = golden("tiny_next")
    s, w, _, coo = problem(g)
    logits = O.gcn_forward_np(s.features, coo, w, s.n, torch.float32)
    pe, pn = O.lsa2_balanced(O.posterior(logits, "twitch/ES/RU"), g["bf_exist"], g["bf_nonexist"])
    np.testing.assert_allclose(pe, g["bf_post_exist"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(pn, g["bf_post_nonexist"], rtol=2e-4, atol=2e-6)
    ae, an = O.lsa2_balanced(s.features, g["bf_exist"], g["bf_nonexist"])
    np.testing.assert_allclose(ae, g["bf_attr_exist"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(an, g["bf_attr_nonexist"], rtol=2e-4, atol=2e-6)
