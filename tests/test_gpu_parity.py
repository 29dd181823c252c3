"""GPU parity tests proper: the CUDA path, called through the C ABI (linkteller_b200.ops → ctypes →
liblinkteller_b200.so), against the golden vectors the unmodified reference produced and against
the oracle on the same seeded inputs.  Integer / index work must be bit-exact; floating point
tolerances are written next to each assertion.  Nothing here reads /root/reference."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O

pytestmark = pytest.mark.gpu

CASES = ["tiny_fo", "tiny_aug", "small_hi", "tiny_edgerand", "tiny_lapgraph"]
DELTA = 1e-4


@pytest.fixture(scope="module")
def ops():
    from linkteller_b200 import ops as _ops
    _ops.device_info()
    return _ops


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def case_inputs(g):
    s = synth.make(str(g["shape"]), seed=int(g["seed"]), hidden=int(g["hidden"]))
    w = synth.gcn_weights(s.f, int(g["hidden"]), s.c, seed=42)
    return s, w


def golden_adj_hat(ops, g, n):
    ip, ix, v = O.coo_to_csr(g["a2_row"], g["a2_col"], g["a2_val"], n)
    return ops.DeviceCSR(dev(ip), dev(ix), dev(v), n)


def model_adj(ops, g, n):
    """The adjacency the model sees before normalisation (clean, or the DP output)."""
    if str(g["dp"]):
        vals = g["dp_data"]
        vd = torch.float64
        return ops.DeviceCSR(dev(g["dp_indptr"].astype(np.int32)), dev(g["dp_indices"].astype(np.int32)),
                             dev(vals.astype(np.float64), vd), n)
    return ops.DeviceCSR(dev(g["a1_indptr"]), dev(g["a1_indices"]), dev(g["a1_data"]), n)


# ------------------------------------------------------------------------------------ primitives
@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 1_000_003])
def test_exclusive_scan(ops, n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 7, n, dtype=np.int32)
    out = ops.exclusive_scan_i32(dev(x)).cpu().numpy()
    ref = np.concatenate([[0], np.cumsum(x, dtype=np.int64)]).astype(np.int32)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("n,bits", [(2, (0, 64)), (5000, (0, 64)), (300_000, (8, 40)), (1_000_000, (32, 64))])
def test_radix_sort_stable(ops, n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 63 - 1, n, dtype=np.int64)
    keys[: n // 3] = keys[n // 3: 2 * (n // 3)]            # force many duplicates
    vals = np.arange(n, dtype=np.int64)
    k, v = ops.radix_sort_pairs(dev(keys), dev(vals), *bits)
    mask = ((1 << (bits[1] - bits[0])) - 1) if bits[1] - bits[0] < 64 else -1
    sub = (keys.view(np.uint64) >> np.uint64(bits[0])) & np.uint64(mask & 0xFFFFFFFFFFFFFFFF)
    order = np.argsort(sub, kind="stable")
    assert np.array_equal(v.cpu().numpy(), vals[order])
    assert np.array_equal(k.cpu().numpy(), keys[order])


# ------------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("case", CASES)
def test_a1_csr_from_edges_bit_exact(ops, golden, case):
    g = golden(case)
    s, _ = case_inputs(g)
    csr = ops.csr_from_edges(dev(s.edges), s.n)
    assert np.array_equal(csr.indptr.cpu().numpy(), g["a1_indptr"])
    assert np.array_equal(csr.indices.cpu().numpy(), g["a1_indices"])
    assert np.array_equal(csr.values.cpu().numpy(), g["a1_data"])


def test_a1_duplicates_selfloops_empty(ops):
    # duplicates are summed, both orientations merge, self loops are doubled (scipy semantics)
    e = np.array([[0, 1], [0, 1], [1, 0], [2, 2], [3, 0], [4, 4], [4, 4]], dtype=np.int64)
    n = 6
    ref = sp.csr_matrix((np.ones(len(e), np.float32), (e[:, 0], e[:, 1])), shape=(n, n))
    ref = (ref + ref.T).tocsr()
    ref.sort_indices()
    csr = ops.csr_from_edges(dev(e), n)
    assert np.array_equal(csr.indptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(csr.indices.cpu().numpy(), ref.indices)
    assert np.array_equal(csr.values.cpu().numpy(), ref.data)
    oip, oix, odv = O.csr_from_edges(e, n)
    assert np.array_equal(oip, ref.indptr) and np.array_equal(oix, ref.indices) and np.array_equal(odv, ref.data)
    b = ops.csr_from_edges(dev(e), n, binarize=True)
    assert np.array_equal(b.values.cpu().numpy(), np.ones(ref.nnz, np.float32))
    empty = ops.csr_from_edges(dev(np.zeros((0, 2), np.int64)), n)
    assert empty.nnz == 0 and np.array_equal(empty.indptr.cpu().numpy(), np.zeros(n + 1, np.int32))
    with pytest.raises(RuntimeError):
        ops.csr_from_edges(dev(np.array([[0, 9]], np.int64)), n)


# ------------------------------------------------------------------------------------ a2/a3
@pytest.mark.parametrize("case", CASES)
def test_a2_normalize_bit_exact(ops, golden, case):
    g = golden(case)
    s, _ = case_inputs(g)
    hat = ops.normalize_adj(model_adj(ops, g, s.n), str(g["norm"]))
    ip, ix, v = O.coo_to_csr(g["a2_row"], g["a2_col"], g["a2_val"], s.n)
    assert np.array_equal(hat.indptr.cpu().numpy(), ip)
    assert np.array_equal(hat.indices.cpu().numpy(), ix)
    assert np.array_equal(hat.values.cpu().numpy().view(np.uint32), v.view(np.uint32))


@pytest.mark.parametrize("kind", ["FirstOrderGCN", "AugNormAdj"])
def test_a2_isolated_nodes_and_weights(ops, kind):
    # isolated nodes (d^-1/2 = inf → 0), weighted entries and a stored diagonal
    rng = np.random.default_rng(7)
    n = 200
    e = rng.integers(0, n - 20, (600, 2)).astype(np.int64)      # nodes n-20.. stay isolated
    e = np.concatenate([e, e[:50], [[3, 3], [5, 5]]])
    csr = ops.csr_from_edges(dev(e), n)
    hat = ops.normalize_adj(csr, kind)
    r, c, v = O.normalize_adj(csr.indptr.cpu().numpy(), csr.indices.cpu().numpy(), csr.values.cpu().numpy(), n, kind)
    ip, ix, vv = O.coo_to_csr(r, c, v, n)
    assert np.array_equal(hat.indptr.cpu().numpy(), ip) and np.array_equal(hat.indices.cpu().numpy(), ix)
    assert np.array_equal(hat.values.cpu().numpy().view(np.uint32), vv.view(np.uint32))


def test_a3_coo_to_csr_and_check(ops, golden):
    g = golden("tiny_fo")
    s, _ = case_inputs(g)
    csr = ops.coo_to_csr(dev(g["a2_row"]), dev(g["a2_col"]), dev(g["a2_val"]), s.n)
    ip, ix, _ = O.coo_to_csr(g["a2_row"], g["a2_col"], g["a2_val"], s.n)
    assert np.array_equal(csr.indptr.cpu().numpy(), ip) and np.array_equal(csr.indices.cpu().numpy(), ix)
    perm = np.random.default_rng(0).permutation(len(g["a2_row"]))
    with pytest.raises(ValueError):
        ops.coo_to_csr(dev(g["a2_row"][perm]), dev(g["a2_col"][perm]), dev(g["a2_val"][perm]), s.n)


# ------------------------------------------------------------------------------------ a4/a5
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 5, 19), (300, 256, 70), (513, 121, 256), (1000, 64, 3170)])
def test_gemm_f32(ops, M, N, K):
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K), dtype=np.float32)
    b = rng.standard_normal((K, N), dtype=np.float32)
    bias = rng.standard_normal(N, dtype=np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    out = ops.gemm(dev(a), dev(b)).cpu().numpy()
    # fp32 accumulation over K terms: error ~ sqrt(K)·eps·|a||b|
    tol = 4e-6 * np.sqrt(K) * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)).max()
    assert np.abs(out - ref).max() <= tol
    out2 = ops.gemm(dev(a), dev(b), bias=dev(bias), relu=True).cpu().numpy()
    assert np.abs(out2 - np.maximum(ref + bias, 0)).max() <= tol
    rows = rng.permutation(M)[: max(1, M // 2)].astype(np.int32)
    out3 = ops.gemm(dev(a), dev(b), a_rows=dev(rows)).cpu().numpy()
    assert np.abs(out3 - ref[rows]).max() <= tol
    out4 = ops.gemm(dev(a), dev(np.ascontiguousarray(b.T)), trans_b=True).cpu().numpy()
    assert np.abs(out4 - ref).max() <= tol


@pytest.mark.parametrize("d", [1, 2, 5, 7, 32, 47, 100, 121, 256, 500])
def test_spmm_csr(ops, golden, d):
    g = golden("small_hi")
    s, _ = case_inputs(g)
    hat = golden_adj_hat(ops, g, s.n)
    rng = np.random.default_rng(d)
    b = rng.standard_normal((s.n, d), dtype=np.float32)
    bias = rng.standard_normal(d, dtype=np.float32)
    A = sp.csr_matrix((g["a2_val"].astype(np.float64), (g["a2_row"], g["a2_col"])), shape=(s.n, s.n))
    ref = A @ b.astype(np.float64)
    out = ops.spmm(hat, dev(b)).cpu().numpy()
    tol = 2e-6 * (abs(A) @ np.abs(b).astype(np.float64)).max()
    assert np.abs(out - ref).max() <= tol
    out2 = ops.spmm(hat, dev(b), bias=dev(bias), relu=True).cpu().numpy()
    assert np.abs(out2 - np.maximum(ref + bias, 0)).max() <= tol
    rows = rng.permutation(s.n)[:333].astype(np.int32)
    out3 = ops.spmm(hat, dev(b), rows=dev(rows)).cpu().numpy()
    assert np.abs(out3 - ref[rows]).max() <= tol


def test_spmm_hub_rows_split_into_segments(ops):
    # a star graph: the hub row has n-1 non-zeros (> one 256-entry segment) — exercises the fix-up
    n = 3000
    e = np.stack([np.zeros(n - 1, np.int64), np.arange(1, n, dtype=np.int64)], axis=1)
    csr = ops.csr_from_edges(dev(e), n)
    hat = ops.normalize_adj(csr, "FirstOrderGCN")
    b = np.random.default_rng(0).standard_normal((n, 64), dtype=np.float32)
    A = hat.to_scipy().astype(np.float64)
    ref = A @ b.astype(np.float64)
    out = ops.spmm(hat, dev(b)).cpu().numpy()
    assert np.abs(out - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("case", CASES)
def test_a5_gcn_forward_logits(ops, golden, case):
    g = golden(case)
    s, w = case_inputs(g)
    hat = golden_adj_hat(ops, g, s.n)
    z = ops.gcn_forward(hat, dev(s.features), *[dev(a) for a in w]).cpu().numpy()
    # fp32 forward vs the reference's fp32 forward (different summation order): 1e-4 of the logit scale
    assert np.abs(z - g["a5_logits"]).max() <= 1e-4 * max(1.0, np.abs(g["a5_logits"]).max())


# ------------------------------------------------------------------------------------ a7/a8
def influence_errors(out, ref):
    """(support equal?, max relative error on the support)"""
    sup = ref != 0
    same = np.array_equal(out != 0, sup)
    rel = np.abs(out[sup] - ref[sup]) / np.abs(ref[sup])
    return same, (rel.max() if rel.size else 0.0)


@pytest.mark.parametrize("case", CASES)
def test_a8_influence_matrix_vs_reference_float64(ops, golden, case):
    g = golden(case)
    s, w = case_inputs(g)
    hat = golden_adj_hat(ops, g, s.n)
    tn = dev(g["a6_test_nodes"].astype(np.int32))
    w1, b1, w2, b2 = [dev(a) for a in w]
    out, info = ops.attack_efficient(hat, dev(s.features), w1, b1, w2, tn, DELTA, want_info=True)
    out = out.cpu().numpy()
    ref64 = g["a8_influence64"]
    same, rel = influence_errors(out, ref64)
    assert same, "exact-zero support differs from the reference"
    # north_star tolerance: influence matrix within 1e-3 relative of the reference (float64 run)
    assert rel <= 1e-3, rel
    # … and the reference's own fp32 run is further from its fp64 run than we are
    _, rel32 = influence_errors(g["a8_influence32"], ref64)
    assert rel <= rel32
    assert info["support_pairs"] == int((ref64 != 0).sum())
    # slices (multi-GPU sharding) reproduce the same rows bit for bit
    T = len(g["a6_test_nodes"])
    lo, hi = T // 3, T // 3 + 7
    part = ops.attack_efficient(hat, dev(s.features), w1, b1, w2, tn, DELTA, lo=lo, hi=hi).cpu().numpy()
    assert np.array_equal(part, out[lo:hi])


def test_a8_plan_api_matches_one_call(ops, golden):
    g = golden("small_hi")
    s, w = case_inputs(g)
    hat = golden_adj_hat(ops, g, s.n)
    tn = dev(g["a6_test_nodes"].astype(np.int32))
    w1, b1, w2, b2 = [dev(a) for a in w]
    x = dev(s.features)
    ref = ops.attack_efficient(hat, x, w1, b1, w2, tn, DELTA)
    plan = ops.Plan(hat, tn)
    rows = plan.rows()
    P = ops.gemm(x, w1)
    H_R = ops.spmm(hat, P, bias=b1, rows=rows)
    out = plan.influence_rows(H_R, P, w2, DELTA, p_index=tn)
    plan.close()
    # same math, different precompute order (Â·(X·W1) vs (Â·X)·W1): 1e-4 relative
    a, b = out.cpu().numpy(), ref.cpu().numpy()
    assert np.array_equal(a != 0, b != 0)
    nz = b != 0
    assert (np.abs(a[nz] - b[nz]) / np.abs(b[nz])).max() <= 1e-4


def test_a8_errors(ops, golden):
    g = golden("tiny_fo")
    s, w = case_inputs(g)
    hat = golden_adj_hat(ops, g, s.n)
    w1, b1, w2, b2 = [dev(a) for a in w]
    x = dev(s.features)
    with pytest.raises(RuntimeError, match="duplicate"):
        ops.attack_efficient(hat, x, w1, b1, w2, dev(np.array([1, 2, 1], np.int32)), DELTA)
    with pytest.raises(RuntimeError, match="outside"):
        ops.attack_efficient(hat, x, w1, b1, w2, dev(np.array([1, s.n], np.int32)), DELTA)
    # δ = 0 would make every entry NaN, a row stride below T would overrun the caller's buffer: both refused up front
    tn = dev(g["a6_test_nodes"].astype(np.int32))
    with pytest.raises(RuntimeError, match="delta must not be 0"):
        ops.attack_efficient(hat, x, w1, b1, w2, tn, 0.0)
    with pytest.raises(RuntimeError, match="ldo"):
        ops.attack_efficient(hat, x, w1, b1, w2, tn, DELTA,
                             out=torch.empty((tn.numel(), tn.numel() - 1), dtype=torch.float32, device="cuda"))
    # asymmetric pattern is refused (the rank-1 column = row shortcut needs symmetry)
    ip, ix, v = O.coo_to_csr(g["a2_row"], g["a2_col"], g["a2_val"], s.n)
    A = sp.csr_matrix((v, ix, ip), shape=(s.n, s.n)).tolil()
    r = int(g["a6_test_nodes"][0])
    c = [k for k in A.rows[r] if k != r][0]
    A[c, r] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    bad = ops.DeviceCSR(dev(A.indptr.astype(np.int32)), dev(A.indices.astype(np.int32)), dev(A.data.astype(np.float32)), s.n)
    with pytest.raises(RuntimeError, match="symmetric"):
        ops.attack_efficient(bad, x, w1, b1, w2, dev(g["a6_test_nodes"].astype(np.int32)), DELTA)


# ------------------------------------------------------------------------------------ a6/a8/a9
@pytest.mark.parametrize("case", CASES)
def test_a6_pair_labels_and_a8_scores(ops, golden, case):
    g = golden(case)
    s, _ = case_inputs(g)
    tn = g["a6_test_nodes"]
    T = len(tn)
    ori = ops.DeviceCSR(dev(g["a1_indptr"]), dev(g["a1_indices"]), dev(g["a1_data"]), s.n)
    labels, npos = ops.pair_labels(ori, dev(tn.astype(np.int32)))
    labels = labels.cpu().numpy().astype(bool)
    ii, jj = np.triu_indices(T, k=1)
    pairs = np.stack([tn[ii], tn[jj]], axis=1)
    assert npos == len(g["a6_exist"]) == labels.sum()
    assert np.array_equal(pairs[labels], g["a6_exist"]) and np.array_equal(pairs[~labels], g["a6_nonexist"])
    # pair scores from the reference's own float32 influence matrix: bit-exact gather
    infl = g["a8_influence32"].astype(np.float32)
    sc = ops.pair_scores(dev(infl)).cpu().numpy()
    assert np.array_equal(sc[labels], g["a8_exist32"].astype(np.float32))
    assert np.array_equal(sc[~labels], g["a8_nonexist32"].astype(np.float32))


@pytest.mark.parametrize("case", CASES)
def test_a9_auc_ap_topk(ops, golden, case):
    g = golden(case)
    y = g["a9_y"].astype(np.uint8)
    pred = np.concatenate([g["a8_exist32"], g["a8_nonexist32"]]).astype(np.float32)
    auc, ap = ops.auc_ap(dev(pred), dev(y))
    # sklearn on the reference's run (attacker.py:382-386); float64 summation order differs: 1e-12
    assert abs(auc - float(g["a9_auc"])) < 1e-12 and abs(ap - float(g["a9_ap"])) < 1e-12
    st = O.topk_stats(y, pred)
    lo, hi = ops.topk_tp(dev(pred), dev(y), [q["n_pos"] for q in st])
    assert lo == [q["tp_min"] for q in st] and hi == [q["tp_max"] for q in st]


def test_a9_heavy_ties_large(ops):
    # 2M scores, 90 % exact zeros (the regime of the influence matrix), many repeated values
    rng = np.random.default_rng(5)
    n = 2_000_000
    pred = np.where(rng.random(n) < 0.9, 0.0, np.round(rng.random(n), 3)).astype(np.float32)
    y = (rng.random(n) < 0.05 + 0.3 * pred).astype(np.uint8)
    auc, ap = ops.auc_ap(dev(pred), dev(y))
    a0, p0 = O.auc_ap(y, pred)
    assert abs(auc - a0) < 1e-10 and abs(ap - p0) < 1e-10
    st = O.topk_stats(y, pred)
    lo, hi = ops.topk_tp(dev(pred), dev(y), [q["n_pos"] for q in st])
    assert lo == [q["tp_min"] for q in st] and hi == [q["tp_max"] for q in st]


# ------------------------------------------------------------------------------------ a10
@pytest.mark.parametrize("case", ["tiny_fo", "small_hi"])
def test_a10_lsa2(ops, golden, case):
    g = golden(case)
    s, _ = case_inputs(g)
    tn = g["a6_test_nodes"]
    pos = {int(nd): i for i, nd in enumerate(tn)}
    tnd = dev(tn.astype(np.int32))
    tr = "sigmoid" if str(g["dataset"]) == "ppi" else "softmax"
    for key, mat, t in (("post", g["a5_logits"], tr), ("attr", s.features, "none")):
        dist = ops.lsa2_corr(dev(mat), tnd, t).cpu().numpy()
        assert not np.tril(dist).any()
        pick = lambda prs: np.array([dist[min(pos[u], pos[v]), max(pos[u], pos[v])] for u, v in prs])
        # float32 dot/norm in a different summation order than torch's: 2e-5 relative, 2e-6 absolute
        np.testing.assert_allclose(pick(g["a6_exist"]), g[f"a10_{key}_exist"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(pick(g["a6_nonexist"]), g[f"a10_{key}_nonexist"], rtol=2e-5, atol=2e-6)


# ------------------------------------------------------------------------------------ a11/a12
def test_a11_edgerand_bit_exact(ops, golden):
    from linkteller_b200 import dp
    g = golden("tiny_edgerand")
    s, _ = case_inputs(g)
    ori = ops.DeviceCSR(dev(g["a1_indptr"]), dev(g["a1_indices"]), dev(g["a1_data"]), s.n)
    out = dp.perturb_adj(ori, "discrete", float(g["epsilon"]), int(g["noise_seed"]))
    assert np.array_equal(out.indptr.cpu().numpy(), g["dp_indptr"])
    assert np.array_equal(out.indices.cpu().numpy(), g["dp_indices"])
    assert np.array_equal(out.values.cpu().numpy(), g["dp_data"]) and (g["dp_data"] == 0).any()
    # … and the normalisation that follows prunes the explicit zeros exactly like the reference
    hat = ops.normalize_adj(out, "FirstOrderGCN")
    ip, ix, v = O.coo_to_csr(g["a2_row"], g["a2_col"], g["a2_val"], s.n)
    assert np.array_equal(hat.indptr.cpu().numpy(), ip) and np.array_equal(hat.indices.cpu().numpy(), ix)
    assert np.array_equal(hat.values.cpu().numpy().view(np.uint32), v.view(np.uint32))


def test_a12_lapgraph_bit_exact(ops, golden):
    from linkteller_b200 import dp
    g = golden("tiny_lapgraph")
    s, _ = case_inputs(g)
    ori = ops.DeviceCSR(dev(g["a1_indptr"]), dev(g["a1_indices"]), dev(g["a1_data"]), s.n)
    out = dp.perturb_adj(ori, "continuous", float(g["epsilon"]), int(g["noise_seed"]))
    assert np.array_equal(out.indptr.cpu().numpy(), g["dp_indptr"])
    assert np.array_equal(out.indices.cpu().numpy(), g["dp_indices"])
    assert np.array_equal(out.values.cpu().numpy(), g["dp_data"].astype(np.float64))
    hat = ops.normalize_adj(out, "FirstOrderGCN")
    ip, ix, v = O.coo_to_csr(g["a2_row"], g["a2_col"], g["a2_val"], s.n)
    assert np.array_equal(hat.indptr.cpu().numpy(), ip) and np.array_equal(hat.indices.cpu().numpy(), ix)
    assert np.array_equal(hat.values.cpu().numpy().view(np.uint32), v.view(np.uint32))


@pytest.mark.parametrize("kind", ["discrete", "continuous"])
def test_a11_a12_twitch_size_vs_oracle(ops, kind):
    # full C4 size (n=4385): device apply vs the oracle's numpy restatement, same seed
    from linkteller_b200 import dp
    s = synth.make("twitch", seed=0)
    ori = ops.csr_from_edges(dev(s.edges), s.n)
    out = dp.perturb_adj(ori, kind, 5.0, 42)
    ip, ix, dv = ori.indptr.cpu().numpy(), ori.indices.cpu().numpy(), ori.values.cpu().numpy()
    if kind == "discrete":
        rip, rix, rdv = O.edgerand(ip, ix, dv, s.n, 5.0, 42)
    else:
        rip, rix, rdv, _ = O.lapgraph(ip, ix, dv, s.n, 5.0, 42)
    assert np.array_equal(out.indptr.cpu().numpy(), rip) and np.array_equal(out.indices.cpu().numpy(), rix)
    assert np.array_equal(out.values.cpu().numpy(), rdv.astype(np.float64))
