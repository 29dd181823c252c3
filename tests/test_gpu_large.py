"""Config C5 (the benchmarked workload: 2.45 M nodes, 126 M stored entries of Â, T = 8192) against the float64
oracle.  Two full forwards per row cost minutes at this size, so the rows are checked with the 2-hop sub-problem
oracle (oracle.LocalInfluence — pinned to the reference's float64 run by tests/test_oracle_golden.py): the 4
highest-degree test nodes (hub rows: hundreds of SpMM segments through k_spmm_fixup, F2 groups through
k_frontier_reduce_big), 12 seeded random rows, and at least two rows whose incidences go through the exact
fix-up (k_flip_fix).  Plus size-independent properties of the whole 8192 x 8192 matrix."""
import numpy as np
import pytest
import torch

from oracle import lt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def large():
    from linkteller_b200 import attack, ops, workloads
    wl = workloads.make("large")
    infl, info = attack.influence_matrix(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, want_info=True)
    torch.cuda.synchronize()
    return wl, infl, info


@pytest.fixture(scope="module")
def local_oracle(large):
    wl, _, _ = large
    hat = wl.adj_hat
    weights = [t.cpu().numpy() for t in (wl.w1, wl.b1, wl.w2)]
    return O.LocalInfluence(hat.indptr.cpu().numpy(), hat.indices.cpu().numpy(), hat.values.cpu().numpy(),
                            wl.x.cpu().numpy(), weights, wl.test_nodes.cpu().numpy())


def test_large_rows_match_the_float64_oracle(large, local_oracle):
    wl, infl, info = large
    assert info["exact_path_incidences"] > 0           # the sign-mask path with its fix-up queue is the one that ran
    rows = O.pick_parity_rows(wl.adj_hat.indptr.cpu().numpy(), wl.test_nodes.cpu().numpy(), n_hub=4, n_rand=12)
    assert len(rows) >= 16
    ref, mask, hits = local_oracle.rows(rows, wl.delta)
    got = infl[torch.from_numpy(rows).to(infl.device)].cpu().numpy().astype(np.float64)
    same, rel = O.compare_rows(got, ref)
    assert same, "exact-zero support differs from the float64 oracle"
    assert rel <= 1e-3, rel                            # north_star: influence within 1e-3 relative
    assert (mask | (got == 0)).all()                   # nothing outside two hops
    assert int((hits > 0).sum()) >= 2, hits            # ≥ 2 of the rows exercise the exact fix-up (k_flip_fix)


def test_large_whole_matrix_properties(large):
    wl, infl, info = large
    T = wl.T
    nz = infl != 0
    assert int(nz.sum()) == info["support_pairs"]
    assert bool(torch.equal(nz, nz.t()))               # two-hop reachability is symmetric
    assert bool(nz.diagonal().all())                   # v is within two hops of itself through the diagonal of Â
    assert bool(torch.isfinite(infl).all())
    # slices — the multi-GPU sharding — reproduce the full run bit for bit
    from linkteller_b200 import ops
    lo, hi = T // 8 * 3, T // 8 * 4
    part = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, lo=lo, hi=hi)
    assert bool(torch.equal(part, infl[lo:hi]))


def test_large_spmm_variants_give_the_same_bits(large):
    """The restricted SpMM has opt-in variants kept for the record: hub rows of X gathered from an
    L2-resident panel under a persisting access-policy window ("spmm_hub"), rows of X staged through shared memory with
    cp.async ("spmm_stage"), and the work-item table that replaces the binary search ("spmm_segmeta", default on).
    Every variant sums a row in the same order: switching them must not change a bit of the influence matrix."""
    from linkteller_b200 import ops
    wl, infl, info = large

    def run():
        return ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta)
    with ops.option("spmm_hub", 1), ops.profile() as prof:
        hub = run()
    assert "spmm_hub_prepare" in prof.report()
    with ops.option("spmm_stage", 1):
        staged = run()
    with ops.option("spmm_stage", 1), ops.option("spmm_hub", 1):
        both = run()
    with ops.option("spmm_segmeta", 0), ops.profile() as prof0:
        searched = run()
    assert "spmm_hub_prepare" not in prof0.report()
    # L2 eviction-priority hints derived from the entry values (default on, threshold measured from the degree
    # histogram): they steer caching only
    with ops.option("spmm_hint", 0):
        unhinted = run()
    with ops.option("spmm_hint", 1), ops.option("spmm_hint_deg", 37):
        hinted = run()
    for name, t in (("hub", hub), ("staged", staged), ("hub+staged", both), ("binary search", searched),
                    ("no hints", unhinted), ("other hints", hinted)):
        assert bool(torch.equal(t, infl)), name


def test_reference_arm_inputs_equal_the_product_arm_inputs(large):
    """bench.py --impl reference builds its C5 inputs with torch ops only (oracle/refgen.py) so that it never loads
    the product library; they must be the very arrays the product arm measures."""
    from oracle import refgen
    wl, _, _ = large
    p = refgen.host_problem("large")
    hat = wl.adj_hat
    ip = hat.indptr.cpu().numpy().astype(np.int64)
    assert p.describe() == wl.describe()
    assert np.array_equal(np.repeat(np.arange(wl.n, dtype=np.int64), np.diff(ip)), p.coo[0])
    assert np.array_equal(hat.indices.cpu().numpy().astype(np.int64), p.coo[1])
    assert np.array_equal(hat.values.cpu().numpy().view(np.uint32), p.coo[2].view(np.uint32))
    assert np.array_equal(wl.x.cpu().numpy(), p.X)
    assert np.array_equal(wl.test_nodes.cpu().numpy().astype(np.int64), p.test_nodes)
    for a, b in zip((wl.w1, wl.b1, wl.w2, wl.b2), p.weights):
        assert np.array_equal(a.cpu().numpy(), b)
