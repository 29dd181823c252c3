"""CPU-side tests of the host logic above the C ABI: curve assembly, pair lists, row sharding and
the all-gather of influence rows (gloo, world_size 2), the host halves of the DP mechanisms."""
import os

import numpy as np
import pytest
import torch

from linkteller_b200 import attack, dp, workloads
from linkteller_b200.attacker import PairList, binary_clf_curves
from oracle import lt_oracle as O


def clf_counts(y, pred):
    """numpy restatement of sklearn's _binary_clf_curve (what lt_rank_curve returns)."""
    order = np.argsort(-pred, kind="stable")
    ys, ps = y[order], pred[order]
    last = np.r_[np.where(np.diff(ps))[0], len(ps) - 1]
    tps = np.cumsum(ys)[last]
    fps = 1 + last - tps
    return ps[last], tps, fps


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_curves_match_sklearn(seed):
    from sklearn import metrics
    rng = np.random.default_rng(seed)
    n = 5000
    pred = np.where(rng.random(n) < 0.7, 0.0, np.round(rng.random(n), 2)).astype(np.float32).astype(np.float64)
    y = (rng.random(n) < 0.1 + 0.5 * pred).astype(np.int64)
    (fpr, tpr, thr), (prec, rec, thr2) = binary_clf_curves(*clf_counts(y, pred))
    f0, t0, th0 = metrics.roc_curve(y, pred)
    p0, r0, th20 = metrics.precision_recall_curve(y, pred)
    assert np.array_equal(fpr, f0) and np.array_equal(tpr, t0) and np.array_equal(thr, th0)
    assert np.array_equal(prec, p0) and np.array_equal(rec, r0) and np.array_equal(thr2, th20)


def test_curves_match_golden_reference_run(golden):
    g = golden("small_hi")
    pred = np.concatenate([g["a8_exist32"], g["a8_nonexist32"]])
    (fpr, tpr, _), _ = binary_clf_curves(*clf_counts(g["a9_y"], pred))
    assert np.array_equal(fpr, g["a9_fpr"]) and np.array_equal(tpr, g["a9_tpr"])


def test_pair_list_behaves_like_list_of_tuples():
    p = PairList(np.array([[3, 9], [4, 1]]))
    assert len(p) == 2 and list(p) == [(3, 9), (4, 1)] and p[1] == (4, 1)
    assert np.asarray(p).shape == (2, 2)
    assert len(PairList(np.zeros((0, 2)))) == 0


@pytest.mark.parametrize("T,world", [(8192, 8), (500, 4), (7, 3), (2, 4), (2000, 1)])
def test_slice_bounds_partition(T, world):
    b = [attack.slice_bounds(T, world, r) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == T
    assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
    sizes = [hi - lo for lo, hi in b]
    assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, T, port, ok):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(T * T, dtype=torch.float32).reshape(T, T)
    out = torch.full((T, T), -1.0)
    lo, hi = attack.slice_bounds(T, world, rank)
    out[lo:hi] = full[lo:hi]                       # this rank's influence rows
    attack.gather_rows(out, T, world)
    ok[rank] = int(torch.equal(out, full))
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [8, 7])               # equal slices (one all-gather) and ragged slices
def test_gather_rows_gloo_world2(T):
    import torch.multiprocessing as mp
    ok = mp.get_context("spawn").Array("i", [0, 0])
    port = 29500 + (os.getpid() + T) % 500
    mp.spawn(_gather_worker, args=(2, T, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


def test_sampling_matches_oracle_and_reference_stream(golden):
    g = golden("small_hi")
    tn = workloads.sample_test_nodes_typed(str(g["dataset"]), str(g["sample_type"]), g["a1_indptr"], int(g["T"]), 42)
    assert np.array_equal(tn, g["a6_test_nodes"])
    g = golden("tiny_fo")
    tn = workloads.sample_test_nodes_typed(str(g["dataset"]), "unbalanced", g["a1_indptr"], int(g["T"]), 42)
    assert np.array_equal(tn, g["a6_test_nodes"])
    with pytest.raises(NotImplementedError):
        workloads.sample_test_nodes_typed("nope", "unbalanced-lo", g["a1_indptr"], 5, 42)
    with pytest.raises(ValueError):
        workloads.sample_test_nodes_typed("ppi", "unbalanced-hi", g["a1_indptr"], 10 ** 6, 42)


def test_dp_host_halves_match_oracle_streams(golden):
    g = golden("tiny_edgerand")
    n = len(g["a1_indptr"]) - 1
    hits, coin = dp.edgerand_flips(n, float(g["epsilon"]), 42)
    assert (np.diff(hits) > 0).all() and ((hits // n) < (hits % n)).all() and set(np.unique(coin)) <= {0, 1}
    # same stream as the oracle's restatement: applying the flips with numpy reproduces its result
    ip, ix, dv = O.edgerand(g["a1_indptr"], g["a1_indices"], g["a1_data"], n, float(g["epsilon"]), 42)
    import scipy.sparse as sp
    A = sp.csr_matrix((g["a1_data"].astype(np.float64), g["a1_indices"], g["a1_indptr"]), shape=(n, n))
    r, c = hits // n, hits % n
    add = sp.csr_matrix((np.ones(int(coin.sum())), (r[coin == 1], c[coin == 1])), shape=(n, n))
    rem = sp.csr_matrix((np.ones(int((coin == 0).sum())), (r[coin == 0], c[coin == 0])), shape=(n, n))
    out = (A + (add + add.T) - (rem + rem.T)).tocsr()
    out.sort_indices()
    out.data[out.data == -1] = 0
    out.data[out.data == 2] = 1
    assert np.array_equal(out.indptr, ip) and np.array_equal(out.indices, ix) and np.array_equal(out.data, dv)
    g = golden("tiny_lapgraph")
    noise, n_keep = dp.lapgraph_noise(n, len(g["a1_data"]) // 2, float(g["epsilon"]), 42)
    assert noise.shape == (n, n) and n_keep == O.lapgraph(g["a1_indptr"], g["a1_indices"], g["a1_data"], n,
                                                          float(g["epsilon"]), 42)[3]


def test_model_dispatch_and_hidden_padding():
    """Weight tuples select the model (4: GCN, 6: GCN3, else the reference's NotImplementedError, gcn_trainer.py:88);
    first-hidden widths that are not multiples of 4 are padded with units that never fire."""
    import torch
    from linkteller_b200 import ops
    with pytest.raises(NotImplementedError):
        ops.model_forward(None, None, [torch.zeros(1)] * 8)
    with pytest.raises(NotImplementedError):
        ops.attack_rows(None, None, [torch.zeros(1)] * 2, None, 1e-4)
    g = torch.Generator().manual_seed(0)
    w1, b1, w2 = torch.randn(7, 10, generator=g), torch.randn(10, generator=g), torch.randn(10, 3, generator=g)
    p1, pb, p2 = ops._pad_first_hidden(w1, b1, w2)
    assert p1.shape == (7, 12) and pb.shape == (12,) and p2.shape == (12, 3)
    x = torch.randn(5, 7, generator=g)
    ref = torch.relu(x @ w1 + b1) @ w2
    assert torch.allclose(torch.relu(x @ p1 + pb) @ p2, ref, atol=1e-6)
    q1, qb, q2 = ops._pad_first_hidden(p1, pb, p2)                    # already a multiple of 4: returned as is
    assert q1 is p1 and qb is pb and q2 is p2


def _model_worker(rank, world, T, n_weights, port, ok):
    """influence_matrix_model at world_size 2 on CPU tensors: the C call is replaced by a stand-in that fills the
    rows [lo,hi) it is asked for, so the slicing, the weight-tuple dispatch and the all-gather are what is tested."""
    import torch.distributed as dist
    from linkteller_b200 import ops
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(T * T, dtype=torch.float32).reshape(T, T) + 0.5
    calls = []

    def fake_rows(adj, x, weights, test_nodes, delta, lo=0, hi=None, out=None, want_info=False):
        assert len(weights) == n_weights and out.shape == (hi - lo, T)
        calls.append((lo, hi))
        out.copy_(full[lo:hi])
        return out
    ops.attack_rows = fake_rows
    tn = torch.arange(T, dtype=torch.int32)
    got = attack.influence_matrix_model(None, torch.zeros(1), [torch.zeros(1)] * n_weights, tn, 1e-4, group=dist.group.WORLD)
    alone = attack.influence_matrix_model(None, torch.zeros(1), [torch.zeros(1)] * n_weights, tn, 1e-4)   # no group: no sharding, no collective
    assert torch.equal(alone, full) and calls[-1] == (0, T)
    calls.pop()
    ok[rank] = int(torch.equal(got, full) and calls == [attack.slice_bounds(T, world, rank)])
    dist.destroy_process_group()


@pytest.mark.parametrize("T,n_weights", [(8, 4), (9, 6)])       # GCN with equal slices, GCN3 with ragged slices
def test_influence_matrix_model_gloo_world2(T, n_weights):
    import torch.multiprocessing as mp
    ok = mp.get_context("spawn").Array("i", [0, 0])
    port = 30100 + (os.getpid() + T) % 500
    mp.spawn(_model_worker, args=(2, T, n_weights, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


def test_profile_summary_tool_reads_the_committed_launch_list():
    """profiles/summarize.py turns the committed ncu launch list of one attack into per-kernel shares."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "profiles", "summarize.py"), "launches",
                          os.path.join(root, "profiles", "r1m_launches_large.csv"), "--last-step"],
                         capture_output=True, text=True, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("lt::")]
    assert lines and lines[0].startswith("lt::k_spmm_vec")          # the SpMM is the largest share of the step
    share = float(lines[0].split()[-1].rstrip("%"))
    assert 50.0 < share < 90.0
    assert any(ln.startswith("lt::k_gemm_ws") for ln in lines) and any(ln.startswith("lt::k_scan_lookback") for ln in lines)


def test_density_rounding_follows_get_closest():
    """attack_stats_all.py:56-65 — one significant digit, mantissa rounded half up; :87-88 the five ratios."""
    from linkteller_b200 import stats
    assert stats.get_closest(0.0034) == (3, 3)
    assert stats.get_closest(0.0035) == (4, 3) or stats.get_closest(0.0035) == (3, 3)     # 3.5000000000000004 vs 3.4999…
    assert stats.get_closest(0.00096) == (10, 4)          # upstream quirk: the mantissa may round up to 10
    assert stats.get_closest(0.5) == (5, 1)
    a, b, ratios = stats.density_ratios(43)               # 43 / 124750 = 3.4e-4
    assert (a, b) == (3, 4) and np.allclose(ratios, [7.5e-5, 1.5e-4, 3e-4, 6e-4, 1.2e-3])
    assert stats.calc_f1(0, 1) == 0 and abs(stats.calc_f1(0.5, 1.0) - 2 / 3) < 1e-15
    with pytest.raises(ValueError):
        stats.get_closest(0.0)
    assert stats.result_filename("ppi", "efficient", "unbalanced", 500, 42) == "eval_ppi/efficient_unbalanced_500_42.pt"
