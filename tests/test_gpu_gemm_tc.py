"""3xTF32 tensor-core GEMM (tcgen05, gemm_tc.cu) against float64: the fp32-parity claim of the split."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 48, 256), (300, 256, 100), (1000, 64, 3170), (4385, 256, 3170),
                                   (513, 121, 256), (77, 16, 40), (2048, 512, 96), (129, 32, 8)])
def test_gemm_3xtf32_matches_float64(M, N, K):
    from linkteller_b200 import ops
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = rng.standard_normal((M, K), dtype=np.float32)
    b = rng.standard_normal((K, N), dtype=np.float32)
    bias = rng.standard_normal(N, dtype=np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    scale = (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)).max()
    # fp32-level accuracy: same bound as the SIMT fp32 kernel (4e-6·sqrt(K)·|a||b|)
    tol = 4e-6 * np.sqrt(K) * scale
    out = ops.gemm(dev(a), dev(b), algo=2).cpu().numpy()
    assert np.abs(out - ref).max() <= tol, (np.abs(out - ref).max(), tol)
    # a plain single-pass TF32 product would be ~1e-3 relative: make sure we are far below that
    assert np.abs(out - ref).max() <= 2e-5 * scale
    out2 = ops.gemm(dev(a), dev(b), bias=dev(bias), relu=True, algo=2).cpu().numpy()
    assert np.abs(out2 - np.maximum(ref + bias, 0)).max() <= tol
    rows = rng.permutation(M)[: max(1, M // 2)].astype(np.int32)
    out3 = ops.gemm(dev(a), dev(b), a_rows=dev(rows), algo=2).cpu().numpy()
    assert np.abs(out3 - ref[rows]).max() <= tol
    out4 = ops.gemm(dev(a), dev(np.ascontiguousarray(b.T)), trans_b=True, algo=2).cpu().numpy()
    assert np.abs(out4 - ref).max() <= tol
    # SIMT and tensor-core paths agree to fp32 rounding
    out5 = ops.gemm(dev(a), dev(b), algo=1).cpu().numpy()
    assert np.abs(out5 - out).max() <= tol


@pytest.mark.parametrize("M,N,K", [(300, 256, 100), (20000, 48, 256), (513, 121, 40)])
def test_lockstep_and_warp_specialised_kernels_agree(M, N, K, monkeypatch):
    """k_gemm_ws (default) and k_gemm_tc (option tc_ws = 0) stage the same operands and issue the same MMAs in the same
    order per accumulator: the results are bit-identical."""
    from linkteller_b200 import ops
    rng = np.random.default_rng(M + N + K)
    a = dev(rng.standard_normal((M, K), dtype=np.float32))
    b = dev(rng.standard_normal((K, N), dtype=np.float32))
    bias = dev(rng.standard_normal(N, dtype=np.float32))
    with ops.option("tc_ws", 1):
        ws = ops.gemm(a, b, bias=bias, relu=True, algo=2)
    with ops.option("tc_ws", 0):
        ls = ops.gemm(a, b, bias=bias, relu=True, algo=2)
    assert torch.equal(ws, ls)


def test_attack_paths_agree_without_overlap(monkeypatch):
    """option overlap = 0 (everything on the caller's stream) and the default (plan tail and big-group reduce on the side
    stream) run the same kernels on the same data: identical influence rows."""
    from linkteller_b200 import ops, workloads
    wl = workloads.make("ppi", f=52)
    with ops.option("overlap", 1):
        a = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta)
    with ops.option("overlap", 0):
        b = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta)
    assert torch.equal(a, b)
    for _ in range(3):                                  # and it is repeatable (no race between the streams)
        assert torch.equal(a, ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta))
