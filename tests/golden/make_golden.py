"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on seeded synthetic inputs.  Run in the build container only (the reference does not exist on
the GPU box):

    python tests/golden/make_golden.py

Each .npz stores the generator arguments (shape name, seeds) and the reference's outputs for
every row of SURVEY.md §8a: symmetrised adjacency (a1), normalised Â as torch COO (a2/a3), clean
logits (a4/a5), sampled nodes and pair lists (a6), the influence matrix in float32 as shipped
("ref32") and with model.double() ("ref64") (a7/a8), the pair scores and sklearn metrics (a9),
LSA2 scores (a10), EdgeRand / Lapgraph outputs (a11/a12).  Inputs are regenerated from the seeds
by linkteller_b200.synth, so only outputs are stored.
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402

from linkteller_b200 import synth  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import attacker as ref_attacker  # noqa: E402
    import utils as ref_utils  # noqa: E402
    import worker as ref_worker  # noqa: E402
    from gcn.models import GCN, GCN3  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def ref_graph_reader(edges, n, dataset):
    """Call the reference's graph_reader twitch branch (load.py:452-460) on a temp CSV."""
    import pandas as pd
    ident = dataset[dataset.find("/") + 1:]
    with tempfile.TemporaryDirectory() as td:
        csv = os.path.join(td, "data", dataset, f"musae_{ident}_edges.csv")
        os.makedirs(os.path.dirname(csv))
        pd.DataFrame(edges, columns=["from", "to"]).to_csv(csv, index=False)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            return ref_utils.graph_reader(None, dataset=dataset, n_nodes=n)
        finally:
            os.chdir(cwd)


def build_model(f, h, c, weights, double=False):
    model = quiet(GCN, f, h, c, 0.5)
    w1, b1, w2, b2 = [torch.from_numpy(w) for w in weights]
    model.load_state_dict({"gc1.weight": w1, "gc1.bias": b1, "gc2.weight": w2, "gc2.bias": b2})
    model.eval()
    return model.double() if double else model


def run_case(name, shape, seed, hidden, norm, sample_type, dataset, T, dp=None, noise_seed=42,
             epsilon=5.0, with_lsa2=True):
    s = synth.make(shape, seed=seed, hidden=hidden)
    n, f, c = s.n, s.f, s.c
    weights = synth.gcn_weights(f, hidden, c, seed=42)
    out = dict(shape=shape, seed=seed, hidden=hidden, norm=norm, sample_type=sample_type,
               dataset=dataset, T=T, delta=1e-4, sample_seed=42, noise_seed=noise_seed,
               epsilon=epsilon, dp=dp or "")

    # a1
    adj_ori = ref_graph_reader(s.edges, n, "twitch/ES/RU").tocsr()
    adj_ori.sort_indices()
    out.update(a1_indptr=adj_ori.indptr, a1_indices=adj_ori.indices, a1_data=adj_ori.data)

    # a11 / a12 (run on the clean graph, before normalisation: worker.py:631-635)
    adj_model = adj_ori
    if dp:
        w = ref_worker.Worker.__new__(ref_worker.Worker)
        w.args = types.SimpleNamespace(epsilon=epsilon, noise_seed=noise_seed, noise_type="laplace",
                                       delta=1e-5)
        if not hasattr(np, "bool"):
            np.bool = bool  # worker.py:294 uses the alias numpy 2 removed
        noisy = quiet(w.perturb_adj, adj_ori.copy(), dp).tocsr()
        noisy.sort_indices()
        out.update(dp_indptr=noisy.indptr, dp_indices=noisy.indices, dp_data=noisy.data)
        adj_model = noisy

    # a2 / a3
    coo = ref_utils.fetch_normalization(norm)(adj_model)
    adj_t = quiet(ref_utils.sparse_mx_to_torch_sparse_tensor, coo)
    ind = adj_t._indices().numpy()
    out.update(a2_row=ind[0], a2_col=ind[1], a2_val=adj_t._values().numpy())

    X = torch.from_numpy(s.features)
    args = types.SimpleNamespace(dataset=dataset, sample_type=sample_type, n_test=T, influence=1e-4,
                                 sample_seed=42, mode="vanilla-clean", attack_mode="efficient",
                                 perturb_type=dp, epsilon=epsilon, noise_seed=noise_seed)
    for tag, dbl in (("32", False), ("64", True)):
        model = build_model(f, hidden, c, weights, double=dbl)
        feats = X.double() if dbl else X
        adj = adj_t.double() if dbl else adj_t
        wk = types.SimpleNamespace(n_nodes=n, n_features=f, features=feats, adj_full=adj,
                                   features_2=feats, adj_2=adj, adj_ori=adj_ori)
        att = ref_attacker.Attacker(args, model, wk)
        quiet(att.prepare_test_data)                                  # a6
        tn = np.asarray(att.test_nodes)
        if tag == "32":
            out.update(a6_test_nodes=tn,
                       a6_exist=np.asarray(att.exist_edges).reshape(-1, 2),
                       a6_nonexist=np.asarray(att.nonexist_edges).reshape(-1, 2))
            with torch.no_grad():
                out["a5_logits"] = model(feats, adj).numpy()          # a4/a5
        infl = np.zeros((T, T))
        with torch.no_grad():
            for i in range(T):                                        # a7/a8 (attacker.py:220-229)
                g = att.get_gradient_eps_mat(tn[i])
                for j in range(T):
                    infl[i][j] = g[tn[j]].norm().item()
        out[f"a8_influence{tag}"] = infl
        cap = {}
        att.compute_and_save = lambda e, ne: cap.update(e=np.asarray(e), ne=np.asarray(ne))
        quiet(att.link_prediction_attack_efficient)
        out[f"a8_exist{tag}"] = cap["e"]
        out[f"a8_nonexist{tag}"] = cap["ne"]
        if tag == "32":
            # a9: the real compute_and_save in a scratch dir, then read the .pt back
            att2 = ref_attacker.Attacker(args, model, wk)
            with tempfile.TemporaryDirectory() as td:
                cwd = os.getcwd()
                os.chdir(td)
                try:
                    buf = io.StringIO()
                    with contextlib.redirect_stdout(buf):
                        att2.compute_and_save(list(cap["e"]), list(cap["ne"]))
                    fn = [os.path.join(dp_, f_) for dp_, _, fs in os.walk(td) for f_ in fs][0]
                    saved = torch.load(fn, weights_only=False)
                    out["a9_filename"] = os.path.relpath(fn, td)
                finally:
                    os.chdir(cwd)
            lines = buf.getvalue().splitlines()
            out["a9_auc"] = float([l for l in lines if l.startswith("auc =")][0].split("=")[1])
            out["a9_ap"] = float([l for l in lines if l.startswith("ap =")][0].split("=")[1])
            out["a9_fpr"] = saved["auc"]["fpr"]
            out["a9_tpr"] = saved["auc"]["tpr"]
            out["a9_y"] = np.asarray(saved["result"]["y"])
            if with_lsa2:                                             # a10
                for mode in ("baseline", "baseline-feat"):
                    a3 = types.SimpleNamespace(**vars(args))
                    a3.attack_mode = mode
                    att3 = ref_attacker.Attacker(a3, model, wk)
                    att3.test_nodes, att3.exist_edges, att3.nonexist_edges = \
                        att.test_nodes, att.exist_edges, att.nonexist_edges
                    cap3 = {}
                    att3.compute_and_save = lambda e, ne: cap3.update(e=np.asarray(e), ne=np.asarray(ne))
                    quiet(att3.baseline_attack)
                    key = "post" if mode == "baseline" else "attr"
                    out[f"a10_{key}_exist"] = cap3["e"]
                    out[f"a10_{key}_nonexist"] = cap3["ne"]
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} nnz(Â)={len(out['a2_val'])} T={T} pairs+={len(out['a6_exist'])} "
          f"auc={out['a9_auc']:.4f} → {os.path.getsize(path) / 1024:.0f} KiB")


def run_case_next(name, shape, seed, hidden, T_naive=12):
    """--sample-type balanced-full (construct_balanced_edge_sets,
    link_prediction_attack_efficient_balanced, baseline_attack_balanced) and the naive per-pair attack
    (link_prediction_attack) — the reference's own outputs, float32 as shipped and float64."""
    s = synth.make(shape, seed=seed, hidden=hidden)
    n, f, c = s.n, s.f, s.c
    weights = synth.gcn_weights(f, hidden, c, seed=42)
    out = dict(shape=shape, seed=seed, hidden=hidden, norm="FirstOrderGCN", delta=1e-4, sample_seed=42,
               T_naive=T_naive)
    adj_ori = ref_graph_reader(s.edges, n, "twitch/ES/RU").tocsr()
    adj_ori.sort_indices()
    coo = ref_utils.fetch_normalization("FirstOrderGCN")(adj_ori)
    adj_t = quiet(ref_utils.sparse_mx_to_torch_sparse_tensor, coo)
    X = torch.from_numpy(s.features)
    base = dict(dataset="twitch/ES/RU", influence=1e-4, sample_seed=42, mode="vanilla-clean", perturb_type=None,
                epsilon=5.0, noise_seed=42)
    for tag, dbl in (("32", False), ("64", True)):
        model = build_model(f, hidden, c, weights, double=dbl)
        feats = X.double() if dbl else X
        adj = adj_t.double() if dbl else adj_t
        wk = types.SimpleNamespace(n_nodes=n, n_features=f, features=feats, adj_full=adj, features_2=feats,
                                   adj_2=adj, adj_ori=adj_ori)
        # ---- balanced-full, efficient
        args = types.SimpleNamespace(sample_type="balanced-full", n_test=0, attack_mode="efficient", **base)
        att = ref_attacker.Attacker(args, model, wk)
        quiet(att.prepare_test_data)
        if tag == "32":
            out.update(bf_exist=np.asarray(att.exist_edges).reshape(-1, 2),
                       bf_nonexist=np.asarray(att.nonexist_edges).reshape(-1, 2),
                       bf_test_nodes=np.asarray(att.test_nodes), bf_n_test=args.n_test)
        cap = {}
        att.compute_and_save = lambda e, ne: cap.update(e=np.asarray(e), ne=np.asarray(ne))
        quiet(att.link_prediction_attack_efficient_balanced)
        out[f"bf_eff_exist{tag}"], out[f"bf_eff_nonexist{tag}"] = cap["e"], cap["ne"]
        if tag == "32":
            for mode in ("baseline", "baseline-feat"):
                a3 = types.SimpleNamespace(**vars(args))
                a3.attack_mode = mode
                att3 = ref_attacker.Attacker(a3, model, wk)
                att3.test_nodes, att3.exist_edges, att3.nonexist_edges = att.test_nodes, att.exist_edges, att.nonexist_edges
                cap3 = {}
                att3.compute_and_save = lambda e, ne: cap3.update(e=np.asarray(e), ne=np.asarray(ne))
                quiet(att3.baseline_attack_balanced)
                key = "post" if mode == "baseline" else "attr"
                out[f"bf_{key}_exist"], out[f"bf_{key}_nonexist"] = cap3["e"], cap3["ne"]
        # ---- naive on a small unbalanced sample (the method saves by itself: read the file back)
        a4 = types.SimpleNamespace(sample_type="unbalanced", n_test=T_naive, attack_mode="naive", **base)
        att4 = ref_attacker.Attacker(a4, model, wk)
        quiet(att4.prepare_test_data)
        with tempfile.TemporaryDirectory() as td:
            cwd = os.getcwd()
            os.chdir(td)
            try:
                quiet(att4.link_prediction_attack)
                fn = [os.path.join(dp_, f_) for dp_, _, fs in os.walk(td) for f_ in fs][0]
                saved = torch.load(fn, weights_only=False)
                if tag == "32":
                    out["naive_filename"] = os.path.relpath(fn, td)
            finally:
                os.chdir(cwd)
        out[f"naive_pred{tag}"] = np.asarray(saved["result"]["pred"])
        if tag == "32":
            out.update(naive_test_nodes=np.asarray(att4.test_nodes), naive_y=np.asarray(saved["result"]["y"]),
                       naive_exist=np.asarray(att4.exist_edges).reshape(-1, 2),
                       naive_nonexist=np.asarray(att4.nonexist_edges).reshape(-1, 2))
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} edges={len(out['bf_exist'])} nonedges={len(out['bf_nonexist'])} "
          f"naive pairs={len(out['naive_pred32'])} → {os.path.getsize(path) / 1024:.0f} KiB")


def run_case_gcn3(name, shape, seed, h1, h2, T, **override):
    """--n-layer 3 (GCN3, gcn/models.py:28-46; built by gcn_trainer.py:81-86) under the
    same efficient attack — the reference's logits, influence matrix (float32 as shipped and float64),
    pair scores and sklearn metrics."""
    s = synth.make(shape, seed=seed, hidden=h1, **override)
    n, f, c = s.n, s.f, s.c
    weights = synth.gcn3_weights(f, h1, h2, c, seed=42)
    out = dict(shape=shape, seed=seed, h1=h1, h2=h2, T=T, delta=1e-4, sample_seed=42, norm="FirstOrderGCN",
               override_keys=np.asarray(sorted(override)), override_vals=np.asarray([override[k] for k in sorted(override)]))
    adj_ori = ref_graph_reader(s.edges, n, "twitch/ES/RU").tocsr()
    adj_ori.sort_indices()
    coo = ref_utils.fetch_normalization("FirstOrderGCN")(adj_ori)
    adj_t = quiet(ref_utils.sparse_mx_to_torch_sparse_tensor, coo)
    X = torch.from_numpy(s.features)
    args = types.SimpleNamespace(dataset="twitch/ES/RU", sample_type="unbalanced", n_test=T, influence=1e-4,
                                 sample_seed=42, mode="vanilla-clean", attack_mode="efficient", perturb_type=None,
                                 epsilon=5.0, noise_seed=42, n_layer=3, hidden1=h1, hidden2=h2)
    keys = ("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias", "gc3.weight", "gc3.bias")
    for tag, dbl in (("32", False), ("64", True)):
        model = quiet(GCN3, f, h1, h2, c, 0.5)
        model.load_state_dict({k: torch.from_numpy(w) for k, w in zip(keys, weights)})
        model.eval()
        if dbl:
            model = model.double()
        feats = X.double() if dbl else X
        adj = adj_t.double() if dbl else adj_t
        wk = types.SimpleNamespace(n_nodes=n, n_features=f, features=feats, adj_full=adj, features_2=feats,
                                   adj_2=adj, adj_ori=adj_ori)
        att = ref_attacker.Attacker(args, model, wk)
        quiet(att.prepare_test_data)
        tn = np.asarray(att.test_nodes)
        if tag == "32":
            out.update(test_nodes=tn, exist=np.asarray(att.exist_edges).reshape(-1, 2),
                       nonexist=np.asarray(att.nonexist_edges).reshape(-1, 2))
            with torch.no_grad():
                out["logits"] = model(feats, adj).numpy()
        infl = np.zeros((T, T))
        with torch.no_grad():
            for i in range(T):
                g = att.get_gradient_eps_mat(tn[i])
                for j in range(T):
                    infl[i][j] = g[tn[j]].norm().item()
        out[f"influence{tag}"] = infl
        cap = {}
        att.compute_and_save = lambda e, ne: cap.update(e=np.asarray(e), ne=np.asarray(ne))
        quiet(att.link_prediction_attack_efficient)
        out[f"score_exist{tag}"], out[f"score_nonexist{tag}"] = cap["e"], cap["ne"]
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    nz = int((out["influence64"] != 0).sum())
    print(f"{name}: n={n} T={T} h1={h1} h2={h2} non-zero influence entries {nz}/{T * T} → "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    torch.set_num_threads(4)
    if "--gcn3-only" in sys.argv:
        run_case_gcn3("tiny_gcn3", "tiny", seed=6, h1=32, h2=16, T=40)
        run_case_gcn3("sparse_gcn3", "small", seed=7, h1=32, h2=24, T=120, m=1800)
        sys.exit(0)
    if "--next-only" in sys.argv:
        run_case_next("tiny_next", "tiny", seed=5, hidden=32)
        sys.exit(0)
    run_case("tiny_fo", "tiny", seed=0, hidden=32, norm="FirstOrderGCN", sample_type="unbalanced",
             dataset="twitch/ES/RU", T=40)
    run_case("tiny_aug", "tiny", seed=1, hidden=32, norm="AugNormAdj", sample_type="unbalanced-lo",
             dataset="twitch/ES/RU", T=40, with_lsa2=False)
    run_case("small_hi", "small", seed=2, hidden=64, norm="FirstOrderGCN", sample_type="unbalanced-hi",
             dataset="ppi", T=60)
    run_case("tiny_edgerand", "tiny", seed=3, hidden=32, norm="FirstOrderGCN", sample_type="unbalanced",
             dataset="twitch/ES/RU", T=40, dp="discrete", with_lsa2=False)
    run_case("tiny_lapgraph", "tiny", seed=4, hidden=32, norm="FirstOrderGCN", sample_type="unbalanced",
             dataset="twitch/ES/RU", T=40, dp="continuous", with_lsa2=False)
    run_case_next("tiny_next", "tiny", seed=5, hidden=32)
    run_case_gcn3("tiny_gcn3", "tiny", seed=6, h1=32, h2=16, T=40)
    run_case_gcn3("sparse_gcn3", "small", seed=7, h1=32, h2=24, T=120, m=1800)
