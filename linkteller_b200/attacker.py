"""`Attacker` with the reference's interface (attacker.py:14-412) on the B200 path.

Same constructor `Attacker(args, model, worker)`, same public methods and attributes
(`prepare_test_data`, `get_gradient_eps_mat`, `link_prediction_attack_efficient`, `baseline_attack`,
`compute_and_save`; `test_nodes`, `exist_edges`, `nonexist_edges`), same printed lines, same
result file name and dict schema.  What changes is where the work happens:

  prepare_test_data   node sampling stays on numpy's legacy stream (same seed ⇒ same nodes); the
                      O(T²·deg) Python pair enumeration (utils/load.py:304-326) becomes lt_pair_labels
  efficient attack    the T x 2 full GCN forwards + T² `.item()` calls (attacker.py:220-229) become
                      one lt_attack_efficient call (frontier propagation, csrc/frontier.cu);
                      lt_attack_efficient3 when the model is GCN3 (--n-layer 3)
  baseline attacks    the T² Python loop of torch expressions (attacker.py:311-316) becomes lt_lsa2_corr
  compute_and_save    sklearn's sort-based curves become lt_rank_curve / lt_auc_ap on the device

`args` needs the fields the reference reads: dataset, sample_type, n_test, influence, sample_seed,
mode, attack_mode (+ perturb_type, epsilon, noise_seed when mode != 'vanilla-clean').  `worker` is
duck-typed exactly like upstream: n_nodes, features / adj_full (or features_2 / adj_2 for
twitch·deezer), adj_ori (scipy CSR).  Tensors may live on the CPU or the GPU; they are staged to the
device once per Attacker.
"""
from __future__ import annotations

import os
import os.path as osp
import time

import numpy as np
import torch

from . import adj as _adj
from . import attack, ops, workloads

SAMPLE_TYPES = ("balanced", "balanced-full", "unbalanced", "unbalanced-lo", "unbalanced-hi", "bfs")


class PairList:
    """exist_edges / nonexist_edges: behaves like the reference's list of (u, v) tuples (len, iter,
    index) but is backed by one int64 [k,2] array — T=8192 gives 33.5 M pairs."""

    def __init__(self, arr: np.ndarray):
        self.arr = np.ascontiguousarray(arr, dtype=np.int64).reshape(-1, 2)

    def __len__(self):
        return len(self.arr)

    def __iter__(self):
        return ((int(u), int(v)) for u, v in self.arr)

    def __getitem__(self, k):
        u, v = self.arr[k]
        return int(u), int(v)

    def __array__(self, dtype=None, copy=None):
        return self.arr if dtype is None else self.arr.astype(dtype)


def binary_clf_curves(thr, tps, fps):
    """sklearn.metrics.roc_curve (drop_intermediate=True) and precision_recall_curve assembled from
    the per-distinct-threshold counts (sklearn/metrics/_ranking.py), as compute_and_save stores them
    (attacker.py:382-385)."""
    thr = np.asarray(thr, dtype=np.float64)
    tps = np.asarray(tps, dtype=np.float64)
    fps = np.asarray(fps, dtype=np.float64)
    # --- roc_curve
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        r_fps, r_tps, r_thr = fps[keep], tps[keep], thr[keep]
    else:
        r_fps, r_tps, r_thr = fps, tps, thr
    r_tps = np.r_[0, r_tps]
    r_fps = np.r_[0, r_fps]
    r_thr = np.r_[np.inf, r_thr]
    fpr = r_fps / r_fps[-1] if r_fps[-1] > 0 else np.repeat(np.nan, r_fps.shape)
    tpr = r_tps / r_tps[-1] if r_tps[-1] > 0 else np.repeat(np.nan, r_tps.shape)
    # --- precision_recall_curve
    ps = tps + fps
    precision = np.zeros_like(tps)
    np.divide(tps, ps, out=precision, where=(ps != 0))
    recall = tps / tps[-1] if tps[-1] != 0 else np.ones_like(tps)
    sl = slice(None, None, -1)
    precision = np.hstack((precision[sl], 1))
    recall = np.hstack((recall[sl], 0))
    return (fpr, tpr, r_thr), (precision, recall, thr[sl])


class Attacker:
    def __init__(self, args, model, worker, device=None, group=None):
        self.args = args
        self.dataset = args.dataset
        self.model = model
        self.worker = worker
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.group = group

        if args.sample_type == 'balanced-full':
            self.args.n_test = self.worker.n_nodes

        if self.dataset.startswith('twitch') or self.dataset.startswith('deezer'):   # attacker.py:24-26
            self.features = self.worker.features_2
            self.adj = self.worker.adj_2
        else:
            self.features = self.worker.features
            self.adj = self.worker.adj_full
        self._dev = None

    # ------------------------------------------------------------------ device staging
    def _staged(self):
        if self._dev is None:
            x = self.features
            if not isinstance(x, torch.Tensor):
                x = torch.from_numpy(np.asarray(x))
            x = x.to(device=self.device, dtype=torch.float32).contiguous()
            hat = _adj.device_csr_of(self.adj, self.device)
            sd = {k: v.detach().to(device=self.device, dtype=torch.float32).contiguous()
                  for k, v in self.model.state_dict().items()}
            need = ["gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"]
            if "gc3.weight" in sd:                   # --n-layer 3: GCN3 (gcn/models.py:28-46)
                need += ["gc3.weight", "gc3.bias"]
            if any(k not in sd for k in need) or len(sd) != len(need):
                raise NotImplementedError("the accelerated attack covers GCN and GCN3 (gcn/models.py:8-46); "
                                          f"model has parameters {sorted(sd)}")
            self._dev = dict(x=x, hat=hat, w1=sd["gc1.weight"], b1=sd["gc1.bias"], w2=sd["gc2.weight"],
                             b2=sd["gc2.bias"], weights=tuple(sd[k] for k in need))
        return self._dev

    def _adj_ori_device(self):
        ori = self.worker.adj_ori
        if isinstance(ori, ops.DeviceCSR):
            return ori
        import scipy.sparse as sp
        m = sp.csr_matrix(ori)
        m.sort_indices()
        return ops.DeviceCSR.from_scipy(m, self.device)

    # ------------------------------------------------------------------ a6
    def prepare_test_data(self):
        st = self.args.sample_type
        if st not in SAMPLE_TYPES:
            raise NotImplementedError(f'sample_type = {st} not implemented!')
        if st in ('balanced', 'bfs'):
            # both crash upstream (ValueError / TypeError)
            raise NotImplementedError(f'sample_type = {st} is outside the accelerated path '
                                      '(unbalanced, unbalanced-lo, unbalanced-hi, balanced-full)')
        ori = self._adj_ori_device()
        if st == 'balanced-full':
            return self._prepare_balanced_full(ori)
        indptr = ori.indptr.cpu().numpy()
        # np.random.seed(sample_seed) + np.random.choice(candidates, T, replace=False)  (attacker.py:45, load.py:379)
        self.test_nodes = workloads.sample_test_nodes_typed(self.dataset, st, indptr, self.args.n_test,
                                                            self.args.sample_seed)
        print('#indice =', ori.n if st == 'unbalanced' else int(self._n_candidates(indptr, st)))
        tn = torch.from_numpy(self.test_nodes.astype(np.int32)).to(self.device)
        labels, n_pos = ops.pair_labels(ori, tn)
        self._labels = labels
        lab = labels.cpu().numpy().astype(bool)
        T = len(self.test_nodes)
        ii, jj = np.triu_indices(T, k=1)
        pairs = np.stack([self.test_nodes[ii], self.test_nodes[jj]], axis=1)
        self.exist_edges = PairList(pairs[lab])
        self.nonexist_edges = PairList(pairs[~lab])
        print('#nodes =', T)
        print('#edges_set =', len(self.exist_edges))
        print('#nonedge_set =', len(self.nonexist_edges))
        print(f'generating testing (non-)edge set done!')

    def _prepare_balanced_full(self, ori):
        """construct_balanced_edge_sets (utils/load.py:219-249): every edge u<v of adj_ori in row order
        and as many non-edges drawn by rejection from numpy's legacy stream.  The stream is consumed
        exactly as upstream does (np.random.choice(n) per endpoint ≡ one randint(0, n) draw), in
        batches; the O(n·deg) Python membership tests become sorted-key lookups."""
        n = ori.n
        indptr = ori.indptr.cpu().numpy().astype(np.int64)
        indices = ori.indices.cpu().numpy().astype(np.int64)
        row = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
        up = indices > row
        edges = np.stack([row[up], indices[up]], axis=1)
        keys = row * n + indices                      # ascending: CSR rows are sorted
        rs = np.random.RandomState(self.args.sample_seed)            # attacker.py:45
        need, got = len(edges), []
        if need == 0:
            raise ValueError("balanced-full needs at least one edge (upstream loops forever here)")
        while need > 0:
            draw = rs.randint(0, n, size=2 * max(1024, int(need * 1.2)))
            u, v = draw[0::2].astype(np.int64), draw[1::2].astype(np.int64)
            def stored(a, b):
                k = a * n + b
                p = np.searchsorted(keys, k)
                return (p < len(keys)) & (keys[np.minimum(p, len(keys) - 1)] == k)
            ok = ~stored(u, v) & ~stored(v, u)        # load.py:244
            take = np.flatnonzero(ok)[:need]
            got.append(np.stack([u[take], v[take]], axis=1))
            need -= len(take)
        nonedges = np.concatenate(got)
        print(f'sampling done! len(edge_set) = {len(edges)}, len(nonedge_set) = {len(nonedges)}')
        self.test_nodes = list(range(n))
        self.exist_edges = PairList(edges)
        self.nonexist_edges = PairList(nonedges)
        self._labels = None
        print(f'generating testing (non-)edge set done!')

    def _n_candidates(self, indptr, st):
        lo, hi = workloads.degree_thresholds(self.dataset)
        deg = np.diff(indptr)
        return (deg <= lo).sum() if st == 'unbalanced-lo' else (deg >= hi).sum()

    # ------------------------------------------------------------------ a7
    def get_gradient_eps_mat(self, v):
        """(f(X + δ·e_v x_vᵀ) − f(X)) / δ for ONE node, [n,c] — attacker.py:100-108, evaluated as the
        reference does (two forwards) but on the sm_100a GCN kernels.  The efficient attack below
        does not call this; it is kept for API parity."""
        d = self._staged()
        delta = float(self.args.influence)
        xp = d["x"].clone()
        xp[v] += d["x"][v] * delta
        f1 = ops.model_forward(d["hat"], xp, d["weights"])
        f0 = ops.model_forward(d["hat"], d["x"], d["weights"])
        return (f1 - f0) / delta

    # ------------------------------------------------------------------ a8
    def influence_matrix(self) -> torch.Tensor:
        """influence_val of attacker.py:216-229 as a float32 [T,T] device tensor."""
        d = self._staged()
        tn = torch.from_numpy(np.asarray(self.test_nodes).astype(np.int32)).to(self.device)
        return attack.influence_matrix_model(d["hat"], d["x"], d["weights"], tn, float(self.args.influence),
                                             group=self.group)

    def link_prediction_attack_efficient(self):
        t = time.time()
        infl = self.influence_matrix()
        scores = ops.pair_scores(infl)             # score(u,v) = influence_val[idx v][idx u], attacker.py:233-245
        torch.cuda.synchronize(self.device)
        print(f'time for predicting edges: {time.time() - t}')
        self.influence_val = infl
        self._finish(scores)

    # ------------------------------------------------------------------ f2 / f4
    def _pair_tensors(self, prs):
        a = np.asarray(prs).reshape(-1, 2)
        return (torch.from_numpy(a[:, 0].astype(np.int32)).to(self.device),
                torch.from_numpy(a[:, 1].astype(np.int32)).to(self.device))

    def _listed_influence(self, prs, perturb_first: bool) -> np.ndarray:
        d = self._staged()
        u, v = self._pair_tensors(prs)
        p, o = (u, v) if perturb_first else (v, u)
        w = d["weights"]
        sc = ops.pair_influence(d["hat"], d["x"], d["w1"], d["b1"], d["w2"], p, o, float(self.args.influence),
                                b2=d["b2"] if len(w) == 6 else None, w3=w[4] if len(w) == 6 else None)
        return sc.cpu().numpy().astype(np.float64)

    def link_prediction_attack_efficient_balanced(self):
        """attacker.py:250-284 (--sample-type balanced-full): score(u,v) = ‖get_gradient_eps_mat(u)[v]‖ —
        perturb the FIRST node of the pair; each list comes out ordered by u, then by list position."""
        t = time.time()
        outs = []
        for prs in (self.exist_edges, self.nonexist_edges):
            sc = self._listed_influence(prs, perturb_first=True)
            order = np.argsort(np.asarray(prs).reshape(-1, 2)[:, 0], kind="stable")
            outs.append(sc[order])
        torch.cuda.synchronize(self.device)
        print(f'time for predicting edges: {time.time() - t}')
        self.compute_and_save(outs[0], outs[1])

    def link_prediction_attack(self):
        """The naive per-pair attack (attacker.py:143-206): score(u,v) = ‖get_gradient_eps(u, v)‖ =
        ‖grad_mat(v)[u]‖ (:87-97).  Same numbers as the efficient attack, one entry at a time upstream;
        here the listed entries are gathered from blocks of influence rows.  Keeps upstream's file
        name (no attack_mode prefix, :180-183)."""
        t = time.time()
        if self.args.sample_type == 'balanced-full':
            listed = lambda prs: self._listed_influence(prs, perturb_first=False)
        else:
            # pairs among the T sampled nodes: their entries all live in the T x T influence matrix
            infl = self.influence_matrix()
            tn = np.asarray(self.test_nodes)
            order = np.argsort(tn, kind="stable")                # node id → position in the sample, vectorised

            def listed(prs):
                a = np.asarray(prs).reshape(-1, 2)
                rows = torch.from_numpy(order[np.searchsorted(tn[order], a[:, 1])].astype(np.int32)).to(self.device)
                cols = torch.from_numpy(order[np.searchsorted(tn[order], a[:, 0])].astype(np.int32)).to(self.device)
                return ops.pair_gather(infl, rows, cols).cpu().numpy().astype(np.float64)
        ex = listed(self.exist_edges)
        torch.cuda.synchronize(self.device)
        print(f'time for predicting existing edges: {time.time() - t}')
        t = time.time()
        ne = listed(self.nonexist_edges)
        torch.cuda.synchronize(self.device)
        print(f'time for predicting non-existing edges: {time.time() - t}')
        return self.compute_and_save(ex, ne, naive_name=True)

    def baseline_attack_balanced(self):
        """attacker.py:337-375: LSA2 on listed pairs, mean over ALL nodes (:361), list order kept."""
        t = time.time()
        d = self._staged()
        mode = self.args.attack_mode
        if mode == 'baseline':
            mat = ops.model_forward(d["hat"], d["x"], d["weights"])
            tr = "sigmoid" if self.dataset == 'ppi' else "softmax"
        elif mode == 'baseline-feat':
            mat, tr = d["x"], "none"
        else:
            raise NotImplementedError(f'attack_mode={mode} not implemented!')
        outs = []
        for prs in (self.exist_edges, self.nonexist_edges):
            u, v = self._pair_tensors(prs)
            outs.append(ops.lsa2_pairs(mat, u, v, tr).cpu().numpy().astype(np.float64))
        torch.cuda.synchronize(self.device)
        print(f'time for computing correlation value: {time.time() - t}')
        self.compute_and_save(outs[0], outs[1])

    # ------------------------------------------------------------------ a10
    def baseline_attack(self):
        t = time.time()
        d = self._staged()
        mode = self.args.attack_mode
        tn = torch.from_numpy(np.asarray(self.test_nodes).astype(np.int32)).to(self.device)
        if mode == 'baseline':
            logits = ops.model_forward(d["hat"], d["x"], d["weights"])
            dist = ops.lsa2_corr(logits, tn, "sigmoid" if self.dataset == 'ppi' else "softmax")   # attacker.py:296-299
        elif mode == 'baseline-feat':
            dist = ops.lsa2_corr(d["x"], tn, "none")
        else:
            raise NotImplementedError(f'attack_mode={mode} not implemented!')
        # pairs read dist[i][j] with i<j (attacker.py:322-332): the transposed gather of distᵀ
        scores = ops.pair_scores(dist.t().contiguous())
        torch.cuda.synchronize(self.device)
        print(f'time for computing correlation value: {time.time() - t}')
        self.dist = dist
        self._finish(scores)

    # ------------------------------------------------------------------ a9
    def _finish(self, scores: torch.Tensor):
        labels = getattr(self, "_labels", None)
        if labels is None:
            raise RuntimeError("call prepare_test_data() first")
        sc = scores.cpu().numpy()
        lab = labels.cpu().numpy().astype(bool)
        self.compute_and_save(sc[lab].astype(np.float64), sc[~lab].astype(np.float64))

    def compute_and_save(self, norm_exist, norm_nonexist, naive_name: bool = False):
        norm_exist = np.asarray(norm_exist, dtype=np.float64)
        norm_nonexist = np.asarray(norm_nonexist, dtype=np.float64)
        y = np.r_[np.ones(len(norm_exist), dtype=np.int64), np.zeros(len(norm_nonexist), dtype=np.int64)]
        pred = np.r_[norm_exist, norm_nonexist]
        sc = torch.from_numpy(pred.astype(np.float32)).to(self.device)
        lb = torch.from_numpy(y.astype(np.uint8)).to(self.device)
        auc, ap = ops.auc_ap(sc, lb)
        thr, tps, fps = ops.rank_curve(sc, lb)
        (fpr, tpr, thresholds), (precision, recall, thresholds_2) = binary_clf_curves(
            thr.cpu().numpy(), tps.cpu().numpy(), fps.cpu().numpy())
        self.auc, self.ap = auc, ap
        if self.group is not None and torch.distributed.get_rank(self.group) != 0:
            return None                                  # sharded run: every rank holds the result, rank 0 reports and saves
        print('auc =', auc)
        print('ap =', ap)

        folder_name = f'eval_{self.dataset}'
        if not osp.exists(folder_name):
            os.makedirs(folder_name)
        prefix = '' if naive_name else f'{self.args.attack_mode}_'          # attacker.py:180-183 vs :391-394
        if self.args.mode == 'vanilla-clean':
            filename = osp.join(folder_name, f'{prefix}{self.args.sample_type}_{self.args.n_test}_{self.args.sample_seed}.pt')
        else:
            filename = osp.join(folder_name, f'{prefix}{self.args.sample_type}_{self.args.perturb_type}_{self.args.n_test}_{self.args.sample_seed}_eps-{self.args.epsilon}_seed-{self.args.noise_seed}.pt')
        big = len(pred) > 5_000_000      # the reference stores Python lists; keep arrays past 5 M pairs
        torch.save({
            'auc': {'fpr': fpr, 'tpr': tpr, 'thresholds': thresholds},
            'pr': {'precision': precision, 'recall': recall, 'thresholds': thresholds_2},
            'result': {'y': y if big else y.tolist(), 'pred': pred if big else pred.tolist()},
        }, filename)
        print(f'attack results saved to: {filename}')
        return filename
