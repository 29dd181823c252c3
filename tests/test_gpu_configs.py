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
