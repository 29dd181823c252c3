"""Hub-heavy graphs: groups of the frontier reduction above the CTA threshold, rows of the SpMM
split into segments, unbalanced-hi sampling — checked against the float64 oracle."""
import numpy as np
import pytest
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_star_of_hubs_influence_matches_oracle():
    from linkteller_b200 import ops
    # 3 hubs connected to everyone + a sparse random background; test nodes include the hubs, so the
    # (hub,hub) groups have ~n records (> F2_BIG) and hub rows span several SpMM segments
    rng = np.random.default_rng(11)
    n, f, h, c = 1200, 36, 64, 5
    hubs = np.array([7, 400, 901])
    e = [np.stack([np.full(n, hb), np.arange(n)], 1) for hb in hubs]
    e.append(rng.integers(0, n, (2500, 2)))
    e = np.concatenate(e).astype(np.int64)
    e = e[e[:, 0] != e[:, 1]]
    x = rng.standard_normal((n, f), dtype=np.float32)
    w = synth.gcn_weights(f, h, c, seed=3)
    adj = ops.csr_from_edges(dev(e), n, binarize=True)
    hat = ops.normalize_adj(adj, "FirstOrderGCN")
    tn = np.concatenate([hubs, rng.permutation(np.setdiff1d(np.arange(n), hubs))[:45]]).astype(np.int64)
    out, info = ops.attack_efficient(hat, dev(x), dev(w[0]), dev(w[1]), dev(w[2]), dev(tn.astype(np.int32)), 1e-4,
                                     want_info=True)
    out = out.cpu().numpy()
    A = hat.to_scipy().tocoo()
    ref = O.influence_matrix(x, (A.row.astype(np.int64), A.col.astype(np.int64), A.data), w, n, tn, 1e-4,
                             dtype=torch.float64)
    sup = ref != 0
    assert np.array_equal(out != 0, sup)
    rel = np.abs(out[sup] - ref[sup]) / np.abs(ref[sup])
    assert rel.max() <= 1e-3, rel.max()          # north_star tolerance vs the float64 reference
    assert info["records"] > 3 * 1000            # the hub groups really are big
