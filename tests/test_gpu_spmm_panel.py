"""Panel-blocked SpMM (k_spmm_pan, one launch per column panel) vs the plain kernel and a float64
product, on a graph whose dense operand exceeds the L2 (the regime it is selected for), including hub
rows split into segments, a row subset, and several panel widths.  Both kernels add the terms of a
row in ascending column order, so they must agree bit for bit."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_spmm_panel_matches_plain_and_float64():
    from linkteller_b200 import ops, workloads
    n, m, d = 700_000, 3_500_000, 100            # B = 280 MB > L2
    edges = workloads.chung_lu_edges_device(n, m, 5, "cuda")
    adj = ops.csr_from_edges(edges, n, binarize=True)
    hat = ops.normalize_adj(adj, "FirstOrderGCN")
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    b = torch.randn((n, d), generator=g, device="cuda")
    bias = torch.randn(d, generator=g, device="cuda")
    rows = torch.randperm(n, generator=g, device="cuda")[:200_000].sort().values.int()
    deg = (hat.indptr[1:] - hat.indptr[:-1]).max().item()
    assert deg > 256                             # hub rows span several segments
    out0 = ops.spmm(hat, b, bias=bias, relu=True)     # plain kernel (default)
    sub0 = ops.spmm(hat, b, rows=rows)
    with ops.option("spmm_panel", 1):                # the panel path is opt-in
        out = ops.spmm(hat, b, bias=bias, relu=True)
        sub = ops.spmm(hat, b, rows=rows)
        for mb, lmin in ((16, 1), (100, 64), (48, 300)):
            with ops.option("spmm_panel_mb", mb), ops.option("spmm_panel_lmin", lmin):
                assert torch.equal(ops.spmm(hat, b, rows=rows), sub0)
    # float64 check through torch's CSR product
    A = torch.sparse_csr_tensor(hat.indptr.long(), hat.indices.long(), hat.values.double(), (n, n))
    ref = torch.relu(A @ b.double() + bias.double())
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() <= 2e-6 * max(scale, 1.0) * 4     # fp32 sums of ≤ ~1e4 terms
    assert (out0.double() - ref).abs().max().item() <= 2e-6 * max(scale, 1.0) * 4
    assert torch.equal(out, out0)
    assert torch.equal(sub, sub0)
    assert torch.equal(sub0, ops.spmm(hat, b)[rows.long()])       # a row subset reproduces the full run exactly
