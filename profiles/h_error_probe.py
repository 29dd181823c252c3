"""How accurate is the float32 evaluation of H_R = (Â[R,:]·X)·W1 + b1 that the sign-mask path thresholds?
Compares it with float64 on sampled frontier rows of the C5 workload; the ratio |H32 − H64| / max|H[r,:]| calibrates
the `fix_eps` option (rows with min|H| below 2^-fix_eps·max|H| are re-evaluated in float64).  Run on the GPU box:
    python profiles/h_error_probe.py > gpurun_out/h_error.txt
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from linkteller_b200 import ops, workloads

for name in ("large", "twitch", "ppi", "flickr"):
    wl = workloads.make(name)
    plan = ops.Plan(wl.adj_hat, wl.test_nodes)
    R = plan.rows()
    plan.close()
    g = torch.Generator(device="cuda").manual_seed(0)
    deg = (wl.adj_hat.indptr[1:] - wl.adj_hat.indptr[:-1])[R.long()]
    pick = torch.cat([torch.randperm(R.numel(), generator=g, device="cuda")[:3000], torch.argsort(deg, descending=True)[:300]])
    rows = R[pick].contiguous()
    if wl.f < wl.h:
        ax = ops.spmm(wl.adj_hat, wl.x, rows=rows)                       # the attack's aggregate-first order
        h32 = ops.gemm(ax, wl.w1, bias=wl.b1, algo=2 if ax.shape[0] * wl.h * wl.f >= 1 << 24 else 0)
    else:
        p = ops.gemm(wl.x, wl.w1)
        h32 = ops.spmm(wl.adj_hat, p, bias=wl.b1, rows=rows)
    A = torch.sparse_csr_tensor(wl.adj_hat.indptr.long(), wl.adj_hat.indices.long(), wl.adj_hat.values.double(), (wl.n, wl.n))
    sel = torch.sparse_coo_tensor(torch.stack([torch.arange(rows.numel(), device="cuda"), rows.long()]),
                                  torch.ones(rows.numel(), dtype=torch.float64, device="cuda"), (rows.numel(), wl.n))
    h64 = (torch.sparse.mm(torch.sparse.mm(sel, A.to_sparse_coo()), wl.x.double())) @ wl.w1.double() + wl.b1.double()
    err = (h32.double() - h64).abs()
    scale = h64.abs().amax(dim=1, keepdim=True)
    rel = (err / scale).amax(dim=1)
    print(f"{name}: rows {rows.numel()}  max|H| median {scale.median().item():.3f}  err/max|H|: median {rel.median().item():.3e} "
          f"p99 {rel.quantile(0.99).item():.3e} max {rel.max().item():.3e} -> log2 {np.log2(rel.max().item()):.1f}; "
          f"abs err max {err.max().item():.3e}; hub rows max {rel[-300:].max().item():.3e}")
