"""Offline edge ranking of saved attack results — the hot-path part of `attack_stats_all.py`
(:30-40 load, :47-74 density estimate, :86-116 global top-k precision / recall / F1, AUC) on the device.

The reference script loads nine result files (3 sample types x 3 sample seeds) that `Attacker.compute_and_save`
wrote, estimates the edge density k̂ of every run as the true density rounded to ONE significant digit
(`get_closest`, :56-65), and for ratio ∈ k̂·{¼, ½, 1, 2, 4} predicts the `n_pos = ceil(ratio·n_pairs)` highest-scored
pairs as edges (`np.argpartition`, :108-110): precision = TP/n_pos, recall = TP/Σy, F1.  Its xlsx / csv export is
outside the hot path.

Here the selection runs on the GPU (`lt_topk_tp`: radix select over the scores, rank.cu) and AUC through
`lt_auc_ap`.  `np.argpartition` breaks ties at the cut in an implementation-defined way; the influence scores tie
massively at exactly 0, so every count is returned as the interval [tp_min, tp_max] any tie-break can produce
(equal whenever the cut does not fall inside a tie group) and `cnt` / `prec` / `rec` / `f1` use tp_max.

Quirks kept from upstream, on purpose: the density denominator is the hard-coded `n_nodes = 500` pair count
(:44-45) whatever --n-attack says; `get_closest(0)` does not terminate upstream — here it raises ValueError.
"""
from __future__ import annotations

import os.path as osp
from math import ceil

import numpy as np
import torch

from . import ops

SAMPLE_SEEDS = (42, 2, 82)                                        # attack_stats_all.py:11
SAMPLE_TYPES = ("unbalanced-lo", "unbalanced", "unbalanced-hi")   # :13
N_NODES_DENSITY = 500                                             # :44


def get_closest(value: float):
    """attack_stats_all.py:56-65 — (mantissa rounded half-up to one digit, decimal exponent)."""
    if not value > 0:
        raise ValueError("get_closest: density must be positive (a run without a single edge makes the "
                         "reference script loop forever, attack_stats_all.py:58)")
    cnt = 0
    while value < 1:
        value *= 10
        cnt += 1
    base = int(value)
    value -= int(value)
    if value >= 0.5:
        base += 1
    return base, cnt


def calc_f1(x, y):
    """attack_stats_all.py:76-78."""
    if x == 0 or y == 0:
        return 0
    return 2 * x * y / (x + y)


def density_ratios(n_edges: int, n_nodes: int = N_NODES_DENSITY):
    """The five ratios of :87-88 from the number of true edges of a run: k̂/4 … 4k̂."""
    n_total = (n_nodes - 1) * n_nodes // 2
    a, b = get_closest(n_edges / n_total)
    ratio = a / 10 ** b
    return a, b, [ratio / 4, ratio / 2, ratio, ratio * 2, ratio * 4]


def rank_stats(y, pred, device=None):
    """Precision / recall / F1 at the five density-derived cuts and the AUC of one saved result
    (`data['result']['y']`, `['pred']`).  Returns a dict of python / numpy values."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    y = np.asarray(y)
    pred = np.asarray(pred, dtype=np.float64)
    n_total = len(pred)
    n_true = int(y.sum())
    a, b, ratios = density_ratios(n_true)
    n_pos = [int(ceil(r * n_total)) for r in ratios]                 # :109
    if max(n_pos) > n_total:
        raise ValueError(f"n_pos = {max(n_pos)} exceeds the {n_total} scored pairs (np.argpartition raises upstream)")
    sc = torch.from_numpy(pred.astype(np.float32)).to(device)         # the scores are float32-computed norms
    if not np.array_equal(sc.cpu().numpy().astype(np.float64), pred):
        raise ValueError("scores are not exactly representable in float32; ranking would differ from upstream")
    lb = torch.from_numpy(y.astype(np.uint8)).to(device)
    tp_min, tp_max = ops.topk_tp(sc, lb, n_pos)
    auc, ap = ops.auc_ap(sc, lb)
    cnt = np.asarray(tp_max, dtype=np.int64)
    prec = cnt / np.asarray(n_pos, dtype=np.float64)                  # :112
    rec = cnt / float(n_true)                                         # :113
    f1 = np.asarray([calc_f1(p, r) for p, r in zip(prec, rec)], dtype=np.float64)
    return dict(approx=a, expn=b, ratios=ratios, n_pos=n_pos, tp_min=list(tp_min), tp_max=list(tp_max), cnt=cnt,
                prec=prec, rec=rec, f1=f1, auc=auc, ap=ap, n_edges=n_true, n_total=n_total)


def result_filename(dataset, method_type, sample_type, n_attack, seed, perturb_type=None, eps=None, noise_seed=42):
    """The file compute_and_save wrote (attacker.py:388-394); the DP pattern is the one attack_stats_all.py:34 reads."""
    folder = f"eval_{dataset}"
    if perturb_type is None:
        return osp.join(folder, f"{method_type}_{sample_type}_{n_attack}_{seed}.pt")
    return osp.join(folder, f"{method_type}_{sample_type}_{perturb_type}_{n_attack}_{seed}_eps-{eps}_seed-{noise_seed}.pt")


def attack_stats_all(dataset, method_type="efficient", perturb_type="discrete", n_attack=500, eps=6.0, root=".",
                     sample_types=SAMPLE_TYPES, sample_seeds=SAMPLE_SEEDS, device=None):
    """The arrays attack_stats_all.py fills (:80-116): result_prec / rec / cnt / f1 [types, seeds, 5], result_auc
    [types, seeds], n_edges, approx, expn, plus cnt_min (the other end of the tie interval)."""
    J, K = len(sample_types), len(sample_seeds)
    out = dict(result_prec=np.zeros((J, K, 5)), result_rec=np.zeros((J, K, 5)), result_cnt=np.zeros((J, K, 5), dtype=np.int32),
               result_cnt_min=np.zeros((J, K, 5), dtype=np.int32), result_f1=np.zeros((J, K, 5)), result_auc=np.zeros((J, K)),
               n_edges=np.zeros((J, K), dtype=np.int32), approx=np.zeros((J, K), dtype=np.int32),
               expn=np.zeros((J, K), dtype=np.int32))
    for j, st in enumerate(sample_types):
        for k, seed in enumerate(sample_seeds):
            fn = osp.join(root, result_filename(dataset, method_type, st, n_attack, seed, perturb_type, eps))
            data = torch.load(fn, weights_only=False)
            r = rank_stats(data["result"]["y"], data["result"]["pred"], device=device)
            out["result_prec"][j, k], out["result_rec"][j, k], out["result_f1"][j, k] = r["prec"], r["rec"], r["f1"]
            out["result_cnt"][j, k], out["result_cnt_min"][j, k] = r["tp_max"], r["tp_min"]
            out["result_auc"][j, k] = r["auc"]
            out["n_edges"][j, k], out["approx"][j, k], out["expn"][j, k] = r["n_edges"], r["approx"], r["expn"]
    return out
