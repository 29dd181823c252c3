"""BASELINE.json configurations at full size: the CUDA path vs the float64 oracle on a bounded number
of influence rows (the oracle needs two full forwards per row), plus size-independent properties on
the whole matrix (exact-zero support = 2-hop reachability, slices reproduce the full run)."""
import numpy as np
import pytest
import torch

from oracle import lt_oracle as O

pytestmark = pytest.mark.gpu


def check_rows(wl, rows, tol=1e-3):
    from linkteller_b200 import attack
    infl = attack.influence_matrix(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta).cpu().numpy()
    hat = wl.adj_hat
    ip, ix = hat.indptr.cpu().numpy(), hat.indices.cpu().numpy()
    row = np.repeat(np.arange(wl.n, dtype=np.int64), np.diff(ip))
    coo = (row, ix.astype(np.int64), hat.values.cpu().numpy())
    weights = [t.cpu().numpy() for t in (wl.w1, wl.b1, wl.w2, wl.b2)]
    tn = wl.test_nodes.cpu().numpy().astype(np.int64)
    ref = O.influence_matrix(wl.x.cpu().numpy(), coo, weights, wl.n, tn, wl.delta, dtype=torch.float64, rows=rows)
    got = infl[list(rows)]
    sup = ref != 0
    assert np.array_equal(got != 0, sup), "exact-zero support differs from the reference"
    rel = np.abs(got[sup] - ref[sup]) / np.abs(ref[sup])
    assert rel.max() <= tol, rel.max()              # north_star: within 1e-3 relative (vs the float64 reference run)
    # whole matrix: support = pairs within two hops in Â's pattern, nothing else
    assert np.array_equal(infl != 0, O.two_hop_support(ip, ix, tn) & (infl != 0))
    return infl


def test_config_ppi_tensor_core_frontier():
    # C2: PPI-shaped, unbalanced-hi, c=121, h=256 — the ΔR·W2 product runs on tcgen05 (3xTF32)
    from linkteller_b200 import ops, workloads
    wl = workloads.make("ppi")
    with ops.profile() as prof:
        infl = check_rows(wl, range(0, 500, 21))
    assert "frontier_dq" in prof.report()
    # slices (the multi-GPU sharding) reproduce the same rows bit for bit
    part = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, lo=100, hi=164)
    assert np.array_equal(part.cpu().numpy(), infl[100:164])


def test_config_twitch():
    # C1: twitch-RU-shaped, f=3170 > h (P-first precompute, X·W1 on the tensor cores), c=2
    from linkteller_b200 import workloads
    wl = workloads.make("twitch")
    check_rows(wl, range(0, 500, 63))


def test_config_flickr():
    # C3: Flickr-shaped, T=2000, f=500, c=7
    from linkteller_b200 import workloads
    wl = workloads.make("flickr")
    check_rows(wl, [0, 777, 1999])


def test_sign_mask_path_matches_value_path_and_oracle(monkeypatch):
    """lt_attack_efficient keeps only the signs of H_R (one bit per element + the row's smallest magnitude) when
    the products run on the tensor cores.  Rows where no ReLU can switch give the bits of the value path; rows
    where one can are recomputed exactly.  A large δ forces many such rows."""
    from linkteller_b200 import ops, workloads
    for delta, min_exact in ((1e-4, 0), (3e-2, 100)):
        wl = workloads.make("ppi", delta=delta, f=52)       # f % 4 == 0 and f < h: aggregate-first precompute
        with ops.option("sign_path", 1), ops.profile() as prof:
            a, info = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, want_info=True)
        assert "gemm_h" in prof.report() and "frontier_dq" in prof.report()
        assert info["exact_path_incidences"] >= min_exact
        with ops.option("sign_path", 0):
            b, info_b = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, want_info=True)
        assert info_b["exact_path_incidences"] >= min_exact        # the value path queues the same switching units
        a, b = a.cpu().numpy(), b.cpu().numpy()
        assert np.array_equal(a == 0, b == 0)
        nz = b != 0
        rel = np.abs(a[nz] - b[nz]) / np.abs(b[nz])
        assert rel.max() <= 1e-3, rel.max()   # a switching ReLU contributes −H/δ: two fp32 evaluations of H differ by ~1e-7/δ there
        if info["exact_path_incidences"] == 0 and info_b["exact_path_incidences"] == 0:
            assert np.array_equal(a, b)                 # same A operand, same GEMM: same bits
        check_rows(wl, range(3, 500, 45))               # and the float64 oracle, at this δ


def test_config_twitch_three_layers():
    # C1 graph under --n-layer 3 with the reference's default widths (--hidden1 16 --hidden2 16, main.py:32-35):
    # three hops cover most of a 4.4 K-node graph, so nearly every pair is in the support
    from linkteller_b200 import ops, synth, workloads
    wl = workloads.make("twitch")
    w = synth.gcn3_weights(wl.f, 16, 16, wl.c, seed=42)
    wd = [torch.from_numpy(a).to(wl.x.device) for a in w]
    out, info = ops.attack_efficient3(wl.adj_hat, wl.x, *wd[:5], wl.test_nodes, wl.delta, want_info=True)
    infl = out.cpu().numpy()
    hat = wl.adj_hat
    ip, ix = hat.indptr.cpu().numpy(), hat.indices.cpu().numpy()
    row = np.repeat(np.arange(wl.n, dtype=np.int64), np.diff(ip))
    coo = (row, ix.astype(np.int64), hat.values.cpu().numpy())
    tn = wl.test_nodes.cpu().numpy().astype(np.int64)
    rows = [0, 123, 499]
    ref = O.influence_matrix(wl.x.cpu().numpy(), coo, w, wl.n, tn, wl.delta, dtype=torch.float64, rows=rows)
    sup = ref != 0
    assert np.array_equal(infl[rows] != 0, sup)
    rel = np.abs(infl[rows][sup] - ref[sup]) / np.abs(ref[sup])
    # Two ReLU layers, f = 3170: where a unit switches under the perturbation ΔR = −H/δ, and H carries the fp32
    # rounding of a 3170-term product, so a few of the 1.5 K checked entries sit at ~1e-2 while the bulk is at ~3e-5.
    # The bar here: 99 % of the entries within 1e-3 of the reference's float64 run, none beyond 5e-2, and never
    # further from it than the reference's own float32 run (which is what a user of the reference sees).
    assert np.quantile(rel, 0.99) <= 1e-3, np.quantile(rel, 0.99)
    assert rel.max() <= 5e-2, rel.max()
    ref32 = O.influence_matrix(wl.x.cpu().numpy(), coo, w, wl.n, tn, wl.delta, dtype=torch.float32, rows=rows)
    rel32 = np.abs(ref32[sup] - ref[sup]) / np.abs(ref[sup])
    assert rel.max() <= rel32.max() and np.median(rel) <= np.median(rel32), (rel.max(), rel32.max())
    assert np.array_equal(infl != 0, O.k_hop_support(ip, ix, tn, 3))       # whole matrix: support = three-hop reachability
    assert info["support_pairs"] == int((infl != 0).sum())
    # the same rows through the sharding entry point, as two ranks would compute them
    part = ops.attack_rows(wl.adj_hat, wl.x, wd, wl.test_nodes, wl.delta, lo=250, hi=500)
    assert np.array_equal(part.cpu().numpy(), infl[250:500])


def test_more_test_nodes_than_one_tile_of_the_plan_head():
    """The plan head scans the test-node lengths in tiles of 8192 with a running carry (k_test_lens_scan).  An entry
    influence[i][j] depends on v_i, v_j and the graph only, and its sum runs over the shared neighbours in node order,
    so the matrix of the first 8192 test nodes must reappear BIT FOR BIT inside the matrix of 10 000."""
    from linkteller_b200 import ops, workloads
    wl = workloads.make("flickr", n_test=10000)
    big = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta)
    small = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes[:8192].contiguous(), wl.delta)
    assert bool(torch.equal(big[:8192, :8192], small))
    tail = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes[8192:].contiguous(), wl.delta)
    assert bool(torch.equal(big[8192:, 8192:], tail))
    assert int((big != 0).sum()) > int((small != 0).sum()) + int((tail != 0).sum())      # and the cross blocks are populated
