"""ORACLE — test infrastructure only.  NOT part of the product path.

A CPU restatement of the reference's algorithm for the LinkTeller influence-analysis hot path
(SURVEY.md §8a), written from the reference's behaviour, each function citing the reference
file:line it follows (paths relative to the upstream tree).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / the reported CPU baseline.  The product
(`linkteller_b200`) never imports it and fails loudly when its CUDA library is missing.

Pinning.  The reference ships no tests, fixtures or golden vectors, so the oracle
is pinned against the reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference from /root/reference (in the build container, where it exists), runs it on seeded
synthetic inputs and commits the inputs' seeds and the reference's outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below against
those vectors.  Third-party arithmetic the reference delegates to (torch.mm / torch.spmm,
scipy.sparse algebra, numpy's legacy MT19937 RandomState, sklearn.metrics) is restated here with
numpy where it is integer/byte work and called through the same public API where it is a random
stream (np.random.RandomState) or a BLAS contraction (torch CPU mm/spmm) — the reference pins no
versions; the container's numpy 2.3 / scipy 1.18 / torch 2.11 / sklearn 1.9 are the de-facto pin.

Floating point.  ``dtype=torch.float64`` gives the numerical ground truth ("ref64": the reference
code run with model.double()); ``torch.float32`` reproduces the reference as shipped ("ref32"),
whose finite difference at δ=1e-4 is itself 3-10 % away from ref64.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# a1  graph_reader (utils/load.py:452-460): csr_matrix((ones,(r,c)), float32) ; adj + adj.T
# --------------------------------------------------------------------------------------------

def csr_from_edges(edges: np.ndarray, n: int):
    """Symmetrised adjacency of an edge list, duplicates SUMMED, not binarised, self loops doubled
    (utils/load.py:457-460).  Returns canonical CSR (indptr int32, indices int32, data float32)."""
    r = np.asarray(edges[:, 0], dtype=np.int64)
    c = np.asarray(edges[:, 1], dtype=np.int64)
    keys = np.concatenate([r * n + c, c * n + r])
    uniq, cnt = np.unique(keys, return_counts=True)
    rows = uniq // n
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr.astype(np.int32), (uniq % n).astype(np.int32), cnt.astype(np.float32)


# --------------------------------------------------------------------------------------------
# a2/a3  fetch_normalization → gcn() / aug_normalized_adjacency() (utils/load.py:562-578,617-627)
#        followed by sparse_mx_to_torch_sparse_tensor (utils/load.py:552-559)
# --------------------------------------------------------------------------------------------

def _row_sums(indptr, data, n, dtype):
    out = np.zeros(n, dtype=dtype)
    rows = np.repeat(np.arange(n), np.diff(indptr))
    # sequential accumulation per row in the working dtype (scipy: coo @ ones); np.add.at is
    # unbuffered and in-order, so rounding matches for non-integer weights too
    np.add.at(out, rows, data.astype(dtype))
    return out


def normalize_adj(indptr, indices, data, n: int, kind: str = "FirstOrderGCN"):
    """Â as the torch COO tensor the reference feeds GraphConvolution: (row int64, col int64,
    val float32), row-major sorted, no explicit zeros.

    FirstOrderGCN = I + D^-1/2 A D^-1/2 with D from A (no self loops added) — utils/load.py:572-578.
      Arithmetic dtype follows A.dtype: float32 A ⇒ every product is rounded to float32
      (fl32(fl32(d_i·a_ij)·d_j)); any other dtype ⇒ float64 then one final cast (load.py:554).
    AugNormAdj = D̃^-1/2 (A+I) D̃^-1/2, always float64 (sp.eye promotes) — utils/load.py:562-569.
    Isolated nodes: d^-1/2 = inf → 0 (load.py:566,576).  Zero products are dropped (scipy's
    csr_matmat/binop emit only non-zero sums), which is what prunes EdgeRand's explicit zeros.
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    data = np.asarray(data)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    if kind == "FirstOrderGCN":
        wd = np.float32 if data.dtype == np.float32 else np.float64
        if np.issubdtype(data.dtype, np.integer):
            rs = _row_sums(indptr, data, n, np.int64)
        else:
            rs = _row_sums(indptr, data, n, wd)
        with np.errstate(divide="ignore"):
            d = np.power(rs, -0.5)
        d[np.isinf(d)] = 0.0
        d = d.astype(wd)
        v = (d[rows] * data.astype(wd)).astype(wd)
        v = (v * d[indices]).astype(wd)
        keep = v != 0
        r2, c2, v2 = rows[keep], indices[keep], v[keep].astype(np.float64)
        # + sp.eye (float64): union with the diagonal
        key = np.concatenate([r2 * n + c2, np.arange(n, dtype=np.int64) * (n + 1)])
        val = np.concatenate([v2, np.ones(n)])
    elif kind == "AugNormAdj":
        key = np.concatenate([rows * n + indices, np.arange(n, dtype=np.int64) * (n + 1)])
        val = np.concatenate([data.astype(np.float64), np.ones(n)])
    else:
        raise NotImplementedError(kind)
    order = np.argsort(key, kind="stable")
    key, val = key[order], val[order]
    uk, start = np.unique(key, return_index=True)
    sv = np.add.reduceat(val, start)
    if kind == "AugNormAdj":
        rr = uk // n
        rs = np.zeros(n)
        np.add.at(rs, rr, sv)
        with np.errstate(divide="ignore"):
            d = np.power(rs, -0.5)
        d[np.isinf(d)] = 0.0
        sv = (d[rr] * sv) * d[uk % n]
    nz = sv != 0
    uk, sv = uk[nz], sv[nz]
    return uk // n, uk % n, sv.astype(np.float32)


def coo_to_csr(row, col, val, n):
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, np.asarray(row) + 1, 1)
    return np.cumsum(indptr).astype(np.int32), np.asarray(col, dtype=np.int32), np.asarray(val)


# --------------------------------------------------------------------------------------------
# a4/a5  GraphConvolution.forward (gcn/layers.py:30-36) and GCN.forward (gcn/models.py:19-24)
# --------------------------------------------------------------------------------------------

def _sparse(row, col, val, n, dtype):
    idx = torch.from_numpy(np.vstack([row, col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(np.asarray(val)).to(dtype), (n, n))


def graph_conv(x, adj, w, b):
    """mm → spmm → + bias (gcn/layers.py:31-34)."""
    return torch.spmm(adj, torch.mm(x, w)) + b


def gcn_forward(x, adj, w1, b1, w2, b2):
    """gc1 → ReLU → (dropout is the identity in eval) → gc2, raw logits (gcn/models.py:19-24)."""
    return graph_conv(torch.relu(graph_conv(x, adj, w1, b1)), adj, w2, b2)


def gcn3_forward(x, adj, w1, b1, w2, b2, w3, b3):
    """GCN3 (--n-layer 3): gc1 → ReLU → gc2 → ReLU → gc3, raw logits (gcn/models.py:40-46; dropout is
    the identity in eval)."""
    x = torch.relu(graph_conv(x, adj, w1, b1))
    x = torch.relu(graph_conv(x, adj, w2, b2))
    return graph_conv(x, adj, w3, b3)


def model_forward(x, adj, *weights):
    """GCN for 4 weight arrays, GCN3 for 6 — the two models gcn_trainer.py:81-90 builds for the path."""
    if len(weights) == 4:
        return gcn_forward(x, adj, *weights)
    if len(weights) == 6:
        return gcn3_forward(x, adj, *weights)
    raise NotImplementedError(f"n_layer = {len(weights) // 2} not implemented!")   # gcn_trainer.py:88


def gcn_forward_np(X, coo, weights, n, dtype=torch.float64):
    row, col, val = coo
    adj = _sparse(row, col, val, n, dtype)
    t = [torch.from_numpy(np.asarray(a)).to(dtype) for a in (X, *weights)]
    return model_forward(t[0], adj, *t[1:]).numpy()


# --------------------------------------------------------------------------------------------
# a7/a8  get_gradient_eps_mat (attacker.py:100-108) + link_prediction_attack_efficient
#        (attacker.py:209-247)
# --------------------------------------------------------------------------------------------

def influence_matrix(X, coo, weights, n, test_nodes, delta, dtype=torch.float64, rows=None,
                     threads=None):
    """influence_val[i][j] = || ((f(X + δ·e_v x_vᵀ) − f(X)) / δ)[test_nodes[j]] ||₂ with
    v = test_nodes[i]: two FULL forwards per row, exactly the reference's loop
    (attacker.py:101-106, 220-229).  ``rows`` restricts the outer loop (used to time a bounded
    sample); the result always has len(test_nodes) columns.  float64 storage like np.zeros((T,T)).
    """
    if threads:
        torch.set_num_threads(threads)
    row, col, val = coo
    adj = _sparse(row, col, val, n, dtype)
    x = torch.from_numpy(np.asarray(X)).to(dtype)
    w = [torch.from_numpy(np.asarray(a)).to(dtype) for a in weights]   # 4 arrays: GCN, 6: GCN3
    tn = np.asarray(test_nodes)
    rows = range(len(tn)) if rows is None else rows
    out = np.zeros((len(rows), len(tn)))
    sel = torch.from_numpy(tn.astype(np.int64))
    with torch.no_grad():
        for k, i in enumerate(rows):
            v = int(tn[i])
            pert = torch.zeros_like(x)
            pert[v] = x[v] * delta
            grad = (model_forward(x + pert, adj, *w) - model_forward(x, adj, *w)) / delta
            out[k] = grad[sel].norm(dim=1).double().numpy()
    return out


class LocalInfluence:
    """Float64 influence rows on graphs where two full forwards per row take minutes (config C5).

    The reference's row for a perturbed node v is (f(X') − f(X))/δ read at the test nodes
    (attacker.py:100-108, 220-229).  Every row of both forwards outside N̂(v) = {r : Â[r,v] ≠ 0} goes
    through identical operations on identical inputs, so relu(H')−relu(H) is exactly 0 there and the
    difference of the logits at u is Σ_{r ∈ N̂(u)∩N̂(v)} Â[u,r]·(relu(H'[r]) − relu(H[r]))·W2 (b2 cancels).
    This class evaluates exactly that sub-problem in float64: H[r] for r ∈ N̂(v) from the FULL row r of Â
    and the rows of X it touches, H'[r] = H[r] + Â[r,v]·(δ·x_v)·W1 (layers.py:31-34 on the perturbed row),
    the ReLU difference taken by actual subtraction as the finite difference does (no case analysis),
    then the layer-2 products over N̂(u)∩N̂(v).  It differs from two full float64 forwards only by float64
    rounding of the cancelled terms (~1e-16·|logit|/δ ≈ 1e-11 absolute); tests/test_oracle_golden.py pins it
    to `influence_matrix(dtype=float64)` and to the reference's own float64 run on every golden case.

    `hits` (per row): hidden units (r,k) whose ReLU is within reach of the perturbation,
    |H[r,k]| ≤ |δ·Â[r,v]|·max_k|P[v,k]| — the incidences the CUDA path sends through its exact fix-up."""

    def __init__(self, indptr, indices, values, X, weights, test_nodes):
        import scipy.sparse as sp
        self.n = len(indptr) - 1
        self.A = sp.csr_matrix((np.asarray(values, dtype=np.float64), np.asarray(indices), np.asarray(indptr)),
                               shape=(self.n, self.n))
        self.X = np.asarray(X)
        self.w1, self.b1, self.w2 = [np.asarray(a, dtype=np.float64) for a in weights[:3]]
        self.tn = np.asarray(test_nodes, dtype=np.int64)
        self.AT = self.A[self.tn].tocsc()              # rows of Â at the test nodes, column-addressable

    def row(self, v, delta, chunk_nnz=500_000):
        """(influence [T] float64, structural 2-hop mask [T] bool, n_hits) for the perturbed node v."""
        import scipy.sparse as sp
        A = self.A
        v = int(v)
        R = A.indices[A.indptr[v]:A.indptr[v + 1]].astype(np.int64)          # pattern symmetric: N̂(v) = row v
        sub = A[R]                                                          # full rows of N̂(v)
        a_rv = np.asarray(sub[:, v].todense()).ravel()                      # Â[r,v], looked up in row r
        H = np.empty((len(R), self.w1.shape[1]))
        lens = np.diff(sub.indptr)
        start = 0
        while start < len(R):                                               # bounded temporaries on hub rows
            stop = start + 1
            acc = lens[start]
            while stop < len(R) and acc + lens[stop] <= chunk_nnz:
                acc += lens[stop]
                stop += 1
            part = sub[start:stop]
            cols = np.unique(part.indices)
            loc = sp.csr_matrix((part.data, np.searchsorted(cols, part.indices), part.indptr), shape=(stop - start, len(cols)))
            H[start:stop] = (loc @ self.X[cols].astype(np.float64)) @ self.w1 + self.b1
            start = stop
        pv = self.X[v].astype(np.float64) @ self.w1
        Hp = H + a_rv[:, None] * (delta * pv)[None, :]
        dR = np.maximum(Hp, 0.0) - np.maximum(H, 0.0)
        dQ = dR @ self.w2
        B = self.AT[:, R]                                                   # [T, |R|]: Â[u,r] for test nodes u
        dZ = B @ dQ
        infl = np.sqrt((dZ * dZ).sum(axis=1)) / delta
        mask = np.diff(B.tocsr().indptr) > 0
        hits = int((np.abs(H) <= (np.abs(delta * a_rv) * np.abs(pv).max())[:, None]).any(axis=1).sum())
        return infl, mask, hits

    def rows(self, idx, delta):
        out = np.zeros((len(idx), len(self.tn)))
        mask = np.zeros((len(idx), len(self.tn)), dtype=bool)
        hits = np.zeros(len(idx), dtype=np.int64)
        for k, i in enumerate(idx):
            out[k], mask[k], hits[k] = self.row(self.tn[i], delta)
        return out, mask, hits


def pick_parity_rows(indptr, test_nodes, n_hub=4, n_rand=12, seed=7):
    """Rows for a bounded parity check: the `n_hub` highest-degree test nodes (longest incidence lists, multi-segment
    SpMM rows, big F2 groups) plus `n_rand` seeded random ones.  Returns sorted unique indices into test_nodes."""
    tn = np.asarray(test_nodes, dtype=np.int64)
    deg = np.asarray(indptr)[tn + 1] - np.asarray(indptr)[tn]
    hub = np.argsort(-deg, kind="stable")[:n_hub]
    rnd = np.random.default_rng(seed).choice(len(tn), size=min(n_rand, len(tn)), replace=False)
    return np.unique(np.concatenate([hub, rnd]))


def compare_rows(got, ref):
    """(support_equal, max relative error on the support) of influence rows against the float64 oracle."""
    sup = ref != 0
    same = bool(np.array_equal(got != 0, sup))
    both = sup & (got != 0)
    rel = float((np.abs(got[both] - ref[both]) / np.abs(ref[both])).max()) if both.any() else 0.0
    return same, rel


def influence_rows_local(indptr, indices, values, X, weights, test_nodes, rows, delta):
    """Float64 influence rows `rows` (indices into test_nodes) by the 2-hop sub-problem — see LocalInfluence."""
    return LocalInfluence(indptr, indices, values, X, weights, test_nodes).rows(list(rows), delta)[0]


def k_hop_support(indptr, indices, test_nodes, hops):
    """Boolean [T,T]: test_nodes[j] within ``hops`` hops of test_nodes[i] in Â's pattern (a k-layer GCN's
    influence is bit-exactly 0 outside it: every forward touches the same zeros)."""
    import scipy.sparse as sp
    n = len(indptr) - 1
    tn = np.asarray(test_nodes, dtype=np.int64)
    A = sp.csr_matrix((np.ones(len(indices), dtype=np.float64), indices, indptr), shape=(n, n))
    reach = sp.csr_matrix((np.ones(len(tn)), (np.arange(len(tn)), tn)), shape=(len(tn), n))
    for _ in range(hops):
        reach = reach @ A
        reach.data[:] = 1.0
    return np.asarray(reach[:, tn].todense()) != 0


def two_hop_support(indptr, indices, test_nodes):
    """Boolean [T,T]: entry (i,j) may be non-zero iff test_nodes[j] is within two hops of
    test_nodes[i] in Â's pattern (everything else is bit-exactly 0)."""
    tn = np.asarray(test_nodes)
    n = len(indptr) - 1
    pos = -np.ones(n, dtype=np.int64)
    pos[tn] = np.arange(len(tn))
    sup = np.zeros((len(tn), len(tn)), dtype=bool)
    for i, v in enumerate(tn):
        nb = indices[indptr[v]:indptr[v + 1]]
        for r in nb:
            u = indices[indptr[r]:indptr[r + 1]]
            hit = pos[u]
            sup[i, hit[hit >= 0]] = True
    return sup


# --------------------------------------------------------------------------------------------
# a6  prepare_test_data (attacker.py:33-48) → construct_edge_sets_from_random_subgraph
#     (utils/load.py:338-381), _get_degree (:329-335), _get_edge_sets_among_nodes (:304-326)
# --------------------------------------------------------------------------------------------

def degree_thresholds(dataset: str):
    """lo / hi degree cut-offs keyed by dataset name (utils/load.py:354-370)."""
    if dataset.startswith("twitch"):
        return (5 if "PTBR" not in dataset else 10), 10
    if dataset in ("flickr", "ppi") or dataset.startswith("deezer"):
        return 15, 30
    if dataset in ("cora"):          # substring test, as written upstream (load.py:360)
        return 3, 4
    if dataset in ("citeseer"):
        return 3, 3
    if dataset in ("pubmed"):
        return 10, 10
    raise NotImplementedError(f"lo and hi for dataset = {dataset} not set!")


def sample_test_nodes(dataset, sample_type, indptr, n_samples, sample_seed):
    """np.random.seed(sample_seed) (attacker.py:45) then np.random.choice(candidates, T,
    replace=False) (load.py:379); candidates = all nodes / deg<=lo / deg>=hi on adj_ori's row
    lengths (load.py:343-375)."""
    indptr = np.asarray(indptr)
    n = len(indptr) - 1
    if sample_type == "unbalanced":
        cand = np.arange(n)
    else:
        deg = (indptr[1:] - indptr[:-1]).astype(np.int32)
        lo, hi = degree_thresholds(dataset)
        cand = np.where(deg <= lo)[0] if sample_type == "unbalanced-lo" else np.where(deg >= hi)[0]
    rs = np.random.RandomState(sample_seed)
    if n_samples > len(cand):
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    return cand[rs.permutation(len(cand))[:n_samples]]


def edge_sets_among(indptr, indices, nodes):
    """All pairs i<j in sample order, split by membership of nodes[j] in row nodes[i] of adj_ori
    (load.py:315-321).  Returns (exist [k,2], nonexist [m,2]) in the reference's list order."""
    nodes = np.asarray(nodes)
    T = len(nodes)
    ii, jj = np.triu_indices(T, k=1)          # row-major i<j, same order as the nested loops
    u, v = nodes[ii], nodes[jj]
    n = len(indptr) - 1
    keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr)) * n + np.asarray(indices, dtype=np.int64)
    lab = np.isin(u.astype(np.int64) * n + v, keys)
    pairs = np.stack([u, v], axis=1)
    return pairs[lab], pairs[~lab]


def pair_scores(influence_val, test_nodes, exist, nonexist):
    """score(u,v) = influence_val[idx(v)][idx(u)] — perturb v, observe u (attacker.py:233-245)."""
    pos = {int(nd): i for i, nd in enumerate(test_nodes)}
    f = lambda prs: np.array([influence_val[pos[int(v)], pos[int(u)]] for u, v in prs])
    return f(exist), f(nonexist)


# --------------------------------------------------------------------------------------------
# a9  compute_and_save (attacker.py:378-386) metrics and attack_stats_all.py:56-65,86-116
# --------------------------------------------------------------------------------------------

def auc_ap(y, pred):
    """ROC-AUC and average precision as sklearn computes them for tied scores: one threshold per
    distinct score; AUC = trapezoid over (fpr,tpr) = tie-aware Mann-Whitney; AP = Σ ΔR·P."""
    y = np.asarray(y, dtype=np.int64)
    pred = np.asarray(pred, dtype=np.float64)
    order = np.argsort(-pred, kind="stable")
    ys, ps = y[order], pred[order]
    last = np.r_[np.where(np.diff(ps))[0], len(ps) - 1]   # last index of each tie group
    tps = np.cumsum(ys)[last].astype(np.float64)
    fps = (last + 1) - tps
    P, N = tps[-1], fps[-1]
    tpr = np.r_[0.0, tps / P]
    fpr = np.r_[0.0, fps / N]
    auc = float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0))
    prec = tps / (tps + fps)
    rec = tps / P
    ap = float(np.sum(np.diff(np.r_[0.0, rec]) * prec))
    return auc, ap


def density_round(value):
    """get_closest (attack_stats_all.py:56-65): mantissa rounded half-up to one digit + exponent."""
    cnt = 0
    while value < 1:
        value *= 10
        cnt += 1
    base = int(value)
    value -= int(value)
    if value >= 0.5:
        base += 1
    return base, cnt


def topk_stats(y, pred):
    """precision / recall / F1 at n_pos = ceil(ratio·n_pairs) for ratio ∈ k̂·{¼,½,1,2,4}
    (attack_stats_all.py:86-116).  Ties at the cut are broken by np.argpartition upstream, so the
    oracle returns for every ratio the (min, max) true-positive count any tie-break can give."""
    y = np.asarray(y, dtype=np.int64)
    pred = np.asarray(pred, dtype=np.float64)
    n_total = len(pred)
    base, e = density_round(y.sum() / n_total)
    ratio = base / 10 ** e
    out = []
    for rr in (ratio / 4, ratio / 2, ratio, ratio * 2, ratio * 4):
        n_pos = int(math.ceil(rr * n_total))
        thr = np.partition(pred, n_total - n_pos)[n_total - n_pos]
        above = pred > thr
        tie = pred == thr
        need = n_pos - int(above.sum())
        tp_above = int(y[above].sum())
        tie_pos = int(y[tie].sum())
        tie_neg = int(tie.sum()) - tie_pos
        tp_min = tp_above + max(0, need - tie_neg)
        tp_max = tp_above + min(need, tie_pos)
        out.append(dict(n_pos=n_pos, tp_min=tp_min, tp_max=tp_max, threshold=float(thr)))
    return out


# --------------------------------------------------------------------------------------------
# a10  baseline_attack — LSA2-post / LSA2-attr (attacker.py:287-334)
# --------------------------------------------------------------------------------------------

def lsa2(posterior_or_feats, test_nodes):
    """dist[i][j] (i<j) = <p_u−m, p_v−m> / ‖p_u−m‖ / ‖p_v−m‖, m = mean over the test rows
    (attacker.py:306-316); float32 arithmetic stored into a float64 [T,T] upper triangle."""
    p = torch.as_tensor(np.asarray(posterior_or_feats))
    tn = np.asarray(test_nodes)
    sub = p[torch.from_numpy(tn.astype(np.int64))]
    mean = torch.mean(sub, dim=0)
    cen = sub - mean
    nrm = torch.norm(cen, dim=1)
    T = len(tn)
    dist = np.zeros((T, T))
    for i in range(T):
        if i + 1 < T:
            d = (cen[i + 1:] @ cen[i]) / nrm[i] / nrm[i + 1:]
            dist[i, i + 1:] = d.double().numpy()
    return dist


def posterior(logits, dataset):
    """softmax, or sigmoid iff dataset == 'ppi' (attacker.py:296-299)."""
    z = torch.as_tensor(np.asarray(logits))
    return (torch.sigmoid(z) if dataset == "ppi" else torch.softmax(z, dim=1)).numpy()


# --------------------------------------------------------------------------------------------
# f2 / f4  the "next" rows of SURVEY.md §8f: --sample-type balanced-full
#          (construct_balanced_edge_sets, utils/load.py:219-249;
#          link_prediction_attack_efficient_balanced, attacker.py:250-284;
#          baseline_attack_balanced, attacker.py:337-375) and the naive per-pair attack
#          (link_prediction_attack, attacker.py:143-160 → get_gradient_eps, :87-97)
# --------------------------------------------------------------------------------------------

def balanced_edge_sets(indptr, indices, sample_seed):
    """Every edge u<v of adj_ori in row order (load.py:233-237) and as many non-edges drawn by
    rejection from numpy's legacy stream, one np.random.choice(n) per endpoint (load.py:241-246;
    the stream is seeded by attacker.py:45).  Returns (edge [m,2], nonedge [m,2], test_nodes)."""
    indptr, indices = np.asarray(indptr), np.asarray(indices)
    n = len(indptr) - 1
    rows = [indices[indptr[u]:indptr[u + 1]] for u in range(n)]
    edge = [(u, int(v)) for u in range(n) for v in rows[u] if v > u]
    rs = np.random.RandomState(sample_seed)
    non = []
    while len(non) < len(edge):
        u = rs.choice(n)
        v = rs.choice(n)
        if v not in rows[u] and u not in rows[v]:
            non.append((int(u), int(v)))
    return (np.asarray(edge, dtype=np.int64).reshape(-1, 2), np.asarray(non, dtype=np.int64).reshape(-1, 2),
            np.arange(n))


def _grad_mat(x, adj, w, v, delta):
    pert = torch.zeros_like(x)
    pert[v] = x[v] * delta
    return (model_forward(x + pert, adj, *w) - model_forward(x, adj, *w)) / delta


def _tensors(X, coo, weights, n, dtype):
    row, col, val = coo
    adj = _sparse(row, col, val, n, dtype)
    x = torch.from_numpy(np.asarray(X)).to(dtype)
    w = [torch.from_numpy(np.asarray(a)).to(dtype) for a in weights]
    return x, adj, w


def efficient_balanced_scores(X, coo, weights, n, exist, nonexist, delta, dtype=torch.float64):
    """attacker.py:250-284: for u = 0..n-1 with listed pairs, grad_mat = get_gradient_eps_mat(u) and
    the score of (u,v) is ‖grad_mat[v]‖ — PERTURB THE FIRST node, observe the second.  The two
    output lists are ordered by u, then by position in the input list (the dict-of-lists order)."""
    x, adj, w = _tensors(X, coo, weights, n, dtype)
    outs = []
    with torch.no_grad():
        for prs in (np.asarray(exist).reshape(-1, 2), np.asarray(nonexist).reshape(-1, 2)):
            order = np.argsort(prs[:, 0], kind="stable")
            res = np.zeros(len(prs))
            cur, g = -1, None
            for k in order:
                u, v = int(prs[k, 0]), int(prs[k, 1])
                if u != cur:
                    cur, g = u, _grad_mat(x, adj, w, u, delta)
                res[k] = g[v].norm().item()
            outs.append(res[order])
    return outs[0], outs[1]


def naive_scores(X, coo, weights, n, exist, nonexist, delta, dtype=torch.float64):
    """attacker.py:143-160: score(u,v) = ‖get_gradient_eps(u, v)‖ = ‖grad_mat(v)[u]‖ — perturb the
    SECOND node, observe the first (attacker.py:87-97); list order kept."""
    x, adj, w = _tensors(X, coo, weights, n, dtype)
    outs = []
    with torch.no_grad():
        for prs in (np.asarray(exist).reshape(-1, 2), np.asarray(nonexist).reshape(-1, 2)):
            res = np.zeros(len(prs))
            cache = {}
            for k, (u, v) in enumerate(prs):
                v = int(v)
                if v not in cache:
                    if len(cache) > 8:
                        cache.clear()
                    cache[v] = _grad_mat(x, adj, w, v, delta)
                res[k] = cache[v][int(u)].norm().item()
            outs.append(res)
    return outs[0], outs[1]


def lsa2_balanced(posterior_or_feats, exist, nonexist):
    """attacker.py:337-375: mean over ALL rows (:361), then per listed pair
    <p_u−m, p_v−m> / ‖p_u−m‖ / ‖p_v−m‖ in list order (:364-368); float32 arithmetic."""
    p = torch.as_tensor(np.asarray(posterior_or_feats))
    cen = p - torch.mean(p, dim=0)
    nrm = torch.norm(cen, dim=1)
    outs = []
    for prs in (np.asarray(exist).reshape(-1, 2), np.asarray(nonexist).reshape(-1, 2)):
        u = torch.from_numpy(prs[:, 0].astype(np.int64))
        v = torch.from_numpy(prs[:, 1].astype(np.int64))
        outs.append(((cen[u] * cen[v]).sum(dim=1) / nrm[u] / nrm[v]).double().numpy())
    return outs[0], outs[1]


# --------------------------------------------------------------------------------------------
# a11  perturb_adj_discrete — EdgeRand (worker.py:213-278) + construct_sparse_mat (:178-203)
# --------------------------------------------------------------------------------------------

def edgerand(indptr, indices, data, n, epsilon, noise_seed):
    """Returns (indptr, indices, data float64) of adj + add − minus with −1→0 and 2→1, explicit
    zeros KEPT where a 'minus' hit a non-edge (worker.py:264-267); canonical CSR order."""
    s = 2.0 / (np.exp(epsilon) + 1.0)
    rs = np.random.RandomState(noise_seed)
    hit_r, hit_c = np.where(rs.binomial(1, s, (n, n)))       # worker.py:222-226, row-major order
    dig = rs.binomial(1, 0.5, len(hit_r))                    # same stream, worker.py:228
    up = hit_r < hit_c                                       # construct_sparse_mat keeps i<j
    add = up & (dig == 1)
    rem = up & (dig == 0)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    a_key = rows * n + np.asarray(indices, dtype=np.int64)
    a_val = np.asarray(data, dtype=np.float64)

    def sym(mask):
        r, c = hit_r[mask].astype(np.int64), hit_c[mask].astype(np.int64)
        return np.concatenate([r * n + c, c * n + r])

    k_add, k_rem = sym(add), sym(rem)
    # (adj + add): union, zero sums dropped (none: all positive)
    key = np.concatenate([a_key, k_add])
    val = np.concatenate([a_val, np.ones(len(k_add))])
    uk, inv = np.unique(key, return_inverse=True)
    sv = np.zeros(len(uk))
    np.add.at(sv, inv, val)
    nz = sv != 0
    uk, sv = uk[nz], sv[nz]
    # (… − minus): union, zero results dropped by scipy's binop
    key = np.concatenate([uk, k_rem])
    val = np.concatenate([sv, -np.ones(len(k_rem))])
    uk, inv = np.unique(key, return_inverse=True)
    sv = np.zeros(len(uk))
    np.add.at(sv, inv, val)
    nz = sv != 0
    uk, sv = uk[nz], sv[nz]
    sv[sv == -1] = 0.0
    sv[sv == 2] = 1.0
    ip, ix, dv = coo_to_csr(uk // n, uk % n, sv, n)
    return ip, ix, dv


# --------------------------------------------------------------------------------------------
# a12  perturb_adj_continuous — Lapgraph (worker.py:281-335) + get_noise (utils/load.py:27-39)
# --------------------------------------------------------------------------------------------

def lapgraph(indptr, indices, data, n, epsilon, noise_seed):
    """Top-n_keep cells of tril(A,-1) + Lap(0, 1/ε₂) over the strict lower triangle, rebuilt as a
    symmetric binary int32 CSR.  ε₁ = 0.01ε, ε₂ = ε − ε₁; n_keep = |E| + int(Lap(0,1/ε₁)) drawn
    after RE-SEEDING with the same seed (worker.py:302-304)."""
    n_edges = len(data) // 2
    eps1 = epsilon * 0.01
    eps2 = epsilon - eps1
    noise = np.random.RandomState(noise_seed).laplace(0, 1.0 / eps2, (n, n))
    noise *= np.tri(n, n, k=-1, dtype=bool)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    cols = np.asarray(indices, dtype=np.int64)
    low = cols < rows
    dense = np.zeros((n, n))
    np.add.at(dense, (rows[low], cols[low]), np.asarray(data, dtype=np.float64)[low])
    dense += noise
    n_keep = n_edges + int(np.random.RandomState(noise_seed).laplace(0, 1.0 / eps1, 1)[0])
    flat = dense.ravel()
    ind = np.argpartition(flat, -n_keep)[-n_keep:]
    r, c = ind // n, ind % n
    key = np.unique(np.concatenate([r * n + c, c * n + r]))
    ip, ix, dv = coo_to_csr(key // n, key % n, np.ones(len(key), dtype=np.int32), n)
    return ip, ix, dv, n_keep


# --------------------------------------------------------------------------------------------
# CPU timing of the reference's per-row work (bench.py cpu_baseline / --impl reference only)
# --------------------------------------------------------------------------------------------

def time_reference_rows(X, coo, weights, n, test_nodes, delta, n_rows, slice_div=1, threads=None):
    """Seconds the reference's loop body (attacker.py:220-229 → 100-108 → gcn/models.py:19-24 →
    gcn/layers.py:30-36) takes for `n_rows` influence rows, float32, torch CPU ops exactly as the
    reference issues them.  Returns (seconds_per_row, description).

    slice_div == 1: the rows are evaluated in full (influence_matrix above) and timed.
    slice_div  > 1: for graphs where ONE row costs minutes on a CPU (config C5: ~1e8 non-zeros x
      256 columns per sparse product), every dense op of the row is still executed and timed in
      full, but the two torch.spmm calls of a forward run on the first 1/slice_div of Â's rows (a
      contiguous COO slice) and their time is scaled by nnz(Â)/nnz(slice) — torch's COO spmm is a
      single pass over the non-zeros, so the cost is linear in nnz.  The result is an extrapolated
      timing, labelled as such by the caller; no values are produced in this mode.
    """
    import time
    if threads:
        torch.set_num_threads(threads)
    row, col, val = coo
    tn = np.asarray(test_nodes)
    if slice_div <= 1:
        t0 = time.perf_counter()
        influence_matrix(X, coo, weights, n, tn, delta, dtype=torch.float32, rows=range(n_rows))
        dt = (time.perf_counter() - t0) / n_rows
        return dt, f"{n_rows} full rows (2 GCN forwards + {len(tn)} norms each), reference ops on torch CPU"
    x = torch.from_numpy(np.asarray(X))
    w1, b1, w2, b2 = [torch.from_numpy(np.asarray(a)) for a in weights]
    n_s = max(1, n // slice_div)
    cut = int(np.searchsorted(np.asarray(row), n_s))          # COO is row-major sorted
    adj_s = _sparse(np.asarray(row[:cut]), np.asarray(col[:cut]), np.asarray(val[:cut]), n, torch.float32)
    scale = len(row) / max(cut, 1)
    sel = torch.from_numpy(tn.astype(np.int64))
    tm = time.perf_counter
    t_dense = t_sparse = 0.0
    with torch.no_grad():
        for i in range(n_rows):
            v = int(tn[i])
            t0 = tm()
            pert = torch.zeros_like(x)                         # attacker.py:101-103
            pert[v] = x[v] * delta
            xp = x + pert
            t_dense += tm() - t0
            outs = []
            for feats in (xp, x):                              # attacker.py:105-106: two forwards
                t0 = tm()
                sup = torch.mm(feats, w1)                      # layers.py:31
                t_dense += tm() - t0
                t0 = tm()
                hs = torch.spmm(adj_s, sup)                    # layers.py:32 (sliced, extrapolated)
                t_sparse += tm() - t0
                t0 = tm()
                hfull = torch.relu(sup + b1)                   # same-size stand-in for (Â·sup + b1), relu
                sup2 = torch.mm(hfull, w2)
                t_dense += tm() - t0
                t0 = tm()
                zs = torch.spmm(adj_s, sup2)
                t_sparse += tm() - t0
                t0 = tm()
                outs.append(sup2 + b2)                         # [n,c] output of the forward
                t_dense += tm() - t0
                del hs, zs
            t0 = tm()
            grad = (outs[0] - outs[1]) / delta
            for j in range(len(tn)):                           # attacker.py:224-229 (T .item() calls)
                grad[sel[j]].norm().item()
            t_dense += tm() - t0
    per_row = (t_dense + t_sparse * scale) / n_rows
    return per_row, (f"{n_rows} rows; dense ops in full, torch.spmm on the first 1/{slice_div} of Â's rows "
                     f"({cut} of {len(row)} non-zeros) scaled by {scale:.2f} (extrapolated)")
