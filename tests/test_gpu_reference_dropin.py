"""The drop-in under the REAL reference: the unmodified `main.py` (oracle/_ref, copied from the reference checkout by
oracle/build.py) is run as a user runs it (README.md:43-49: `main.py --test --model-path M --attack … --attack-mode
efficient`) on synthetic files in the twitch layout, once as shipped on the CPU and once with
`linkteller_b200.patch.install()` on the GPU.  Same CLI, same files, same result file name and schema, identical
labels and exact-zero support, scores within 1e-3 of the float64 oracle, AUC equal to the oracle's.

This exercises the whole seam of SURVEY.md §8 a13: Worker.load_data → prepare_data (perturb_adj → fetch_normalization
→ sparse_mx_to_torch_sparse_tensor → .cuda()) → GCNTrainer.init_model / test → eval_output → Attacker."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from linkteller_b200 import synth
from oracle import lt_oracle as O
from oracle import refgen

import twitch_files

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refgen.have_ref(), reason="oracle/_ref not built")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HIDDEN, T = 32, 40


def run_main(td, patched, extra):
    code = (twitch_files.RUN_MAIN if patched else twitch_files.RUN_MAIN_UNPATCHED).format(root=ROOT, ref=refgen.REF_DIR)
    argv = ["--dataset", "twitch/ES/RU", "--hidden", str(HIDDEN), "--norm", "FirstOrderGCN", "--test", "--model-path",
            "model.pt", "--attack", "--approx", "--n-test", str(T), "--influence", "0.0001", "--sample-seed", "42"] + extra
    env = dict(os.environ)
    if not patched:
        env["CUDA_VISIBLE_DEVICES"] = ""             # the reference moves to the GPU whenever it sees one (worker.py:665)
    sub = os.path.join(td, "patched" if patched else "reference")
    os.makedirs(sub, exist_ok=True)
    for name in ("data", "model.pt"):
        if not os.path.exists(os.path.join(sub, name)):
            os.symlink(os.path.join(td, name), os.path.join(sub, name))
    res = subprocess.run([sys.executable, "-c", code] + argv, cwd=sub, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("attack results saved to")]
    assert line, res.stdout[-1500:]
    rel = line[0].split(": ")[1]
    auc = float([ln for ln in res.stdout.splitlines() if ln.startswith("auc =")][0].split("=")[1])
    return rel, torch.load(os.path.join(sub, rel), weights_only=False), auc, res.stdout


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    td = str(tmp_path_factory.mktemp("twitch"))
    ds = twitch_files.make_dataset(td)
    w = synth.gcn_weights(3170, HIDDEN, 2, seed=42)
    torch.save({k: torch.from_numpy(a) for k, a in zip(("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"), w)},
               os.path.join(td, "model.pt"))
    return td, ds, w


def oracle_scores(ds, w, adj_model):
    """float64 pair scores of graph 2 by the oracle (two full forwards per row), pairs labelled from the clean graph."""
    n2, edges2 = ds["RU"][0], ds["RU"][1]
    _, x2 = twitch_files.scaled_features(ds)
    ip, ix, dv = O.csr_from_edges(edges2, n2)
    tn = O.sample_test_nodes("twitch/ES/RU", "unbalanced", ip, T, 42)
    ex, ne = O.edge_sets_among(ip, ix, tn)
    coo = O.normalize_adj(*adj_model(ip, ix, dv, n2), n2, "FirstOrderGCN")
    infl = O.influence_matrix(x2, coo, w, n2, tn, 1e-4, dtype=torch.float64)
    e_s, n_s = O.pair_scores(infl, tn, ex, ne)
    return np.r_[np.ones(len(e_s)), np.zeros(len(n_s))].astype(np.int64), np.r_[e_s, n_s]


def compare(ref, ours, y64, p64):
    (rel_r, d_r, auc_r, _), (rel_o, d_o, auc_o, out_o) = ref, ours
    assert rel_o == rel_r                                                     # same file name (attacker.py:388-394)
    assert set(d_o) == set(d_r) and all(set(d_o[k]) == set(d_r[k]) for k in d_r)   # same dict schema (:396-411)
    y_r, y_o = np.asarray(d_r["result"]["y"]), np.asarray(d_o["result"]["y"])
    p_r, p_o = np.asarray(d_r["result"]["pred"]), np.asarray(d_o["result"]["pred"])
    assert np.array_equal(y_o, y_r) and np.array_equal(y_o, y64)              # same sampled nodes, same pair labels
    assert np.array_equal(p_o == 0, p64 == 0)                                 # same exact zeros as the float64 oracle
    assert not ((p_r != 0) & (p_o == 0)).any()                                # the reference's float32 run can only lose tiny entries
    nz = p64 != 0
    rel = np.abs(p_o[nz] - p64[nz]) / np.abs(p64[nz])
    assert rel.max() <= 1e-3, rel.max()
    rel_ref = np.abs(p_r[nz] - p64[nz]) / np.abs(p64[nz])
    assert rel.max() <= max(rel_ref.max(), 1e-6)                              # never further from float64 than the reference is
    auc64, _ = O.auc_ap(y64, p64)
    assert abs(auc_o - auc64) <= 1e-6, (auc_o, auc64)
    assert abs(auc_o - auc_r) <= 0.02, (auc_o, auc_r)                         # the reference's own float32 noise
    assert "attacks done using" in out_o and "Test set results" in out_o      # the rest of eval_output still ran


def test_reference_main_clean_model_through_the_drop_in(dataset):
    td, ds, w = dataset
    extra = ["--mode", "vanilla-clean", "--sample-type", "unbalanced", "--attack-mode", "efficient"]
    ref = run_main(td, False, extra)
    ours = run_main(td, True, extra)
    y64, p64 = oracle_scores(ds, w, lambda ip, ix, dv, n: (ip, ix, dv))
    compare(ref, ours, y64, p64)


@pytest.mark.parametrize("perturb,fn", [("continuous", "lapgraph"), ("discrete", "edgerand")])
def test_reference_main_dp_model_through_the_drop_in(dataset, perturb, fn):
    """README.md:68-98: `--mode vanilla --perturb-type … --epsilon 5` — Worker.perturb_adj runs on both graphs before
    the normalisation, the attack queries the noisy model and labels pairs from the clean adj_ori."""
    td, ds, w = dataset
    extra = ["--mode", "vanilla", "--perturb-type", perturb, "--epsilon", "5", "--noise-seed", "42", "--sample-type",
             "unbalanced", "--attack-mode", "efficient"]
    ref = run_main(td, False, extra)
    ours = run_main(td, True, extra)

    def noisy(ip, ix, dv, n):
        if fn == "lapgraph":
            return O.lapgraph(ip, ix, dv, n, 5.0, 42)[:3]
        return O.edgerand(ip, ix, dv, n, 5.0, 42)
    y64, p64 = oracle_scores(ds, w, noisy)
    compare(ref, ours, y64, p64)
    assert f"_{perturb}_" in ours[0] and "eps-5.0_seed-42" in ours[0]


@pytest.mark.parametrize("mode", ["baseline", "baseline-feat"])
def test_reference_main_lsa2_baselines_through_the_drop_in(dataset, mode):
    td, _, _ = dataset
    extra = ["--mode", "vanilla-clean", "--sample-type", "unbalanced", "--attack-mode", mode]
    (rel_r, d_r, auc_r, _), (rel_o, d_o, auc_o, _) = run_main(td, False, extra), run_main(td, True, extra)
    assert rel_o == rel_r
    p_r, p_o = np.asarray(d_r["result"]["pred"]), np.asarray(d_o["result"]["pred"])
    y = np.asarray(d_o["result"]["y"])
    assert np.array_equal(y, np.asarray(d_r["result"]["y"]))
    assert np.abs(p_o - p_r).max() <= 2e-5 * max(1.0, np.abs(p_r).max())
    assert abs(auc_o - O.auc_ap(y, p_o)[0]) <= 1e-9                 # the printed AUC is the AUC of the saved scores
    if mode == "baseline-feat":
        # raw features: well conditioned.  (The posterior variant on this random, saturated model yields correlations
        # within 1e-6 of ±1, so its ranking — and AUC — is decided by float32 noise in either implementation.)
        assert abs(auc_o - auc_r) <= 1e-3
