"""Synthetic data files in the layout the reference's loaders read for `--dataset twitch/<A>/<B>`
(utils/load.py:42-71 feature_reader, :452-460 graph_reader; worker.py:470-496): per country
data/twitch/<C>/musae_<C>_{edges.csv, features.json, target.csv}.  Test infrastructure."""
import json
import os

import numpy as np


def write_country(root, country, n, edges, feat_ids, mature, rng):
    d = os.path.join(root, "data", "twitch", country)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f"musae_{country}_edges.csv"), "w") as fh:
        fh.write("from,to\n")
        for u, v in edges:
            fh.write(f"{int(u)},{int(v)}\n")
    with open(os.path.join(d, f"musae_{country}_features.json"), "w") as fh:
        json.dump({str(i): [int(k) for k in feat_ids[i]] for i in range(n)}, fh)
    order = rng.permutation(n)                           # the target file is not sorted by new_id upstream either
    with open(os.path.join(d, f"musae_{country}_target.csv"), "w") as fh:
        fh.write("id,days,mature,views,partner,new_id\n")
        for i in order:
            fh.write(f"{1000 + int(i)},100,{bool(mature[i])},10,False,{int(i)}\n")


def make_dataset(root, n1=260, n2=300, m1=1000, m2=1200, seed=11, n_features=3170, per_node=30):
    """Two countries (train graph, test graph).  Returns dict country → (n, edges, feat_ids, mature)."""
    from linkteller_b200 import synth
    rng = np.random.default_rng(seed)
    out = {}
    for country, n, m in (("ES", n1, m1), ("RU", n2, m2)):
        edges = synth.chung_lu_edges(n, m, rng)
        feat_ids = [np.sort(rng.choice(n_features, size=per_node, replace=False)) for _ in range(n)]
        mature = rng.random(n) < 0.4
        write_country(root, country, n, edges, feat_ids, mature, rng)
        out[country] = (n, edges, feat_ids, mature)
    return out


def scaled_features(ds, n_features=3170):
    """worker.py:476-484: 0/1 matrices, StandardScaler fitted on graph 1, applied to both, cast to float32."""
    from sklearn.preprocessing import StandardScaler

    def dense(n, feat_ids):
        x = np.zeros((n, n_features))
        for i in range(n):
            x[i, feat_ids[i]] = 1
        return x
    x1 = dense(ds["ES"][0], ds["ES"][2])
    x2 = dense(ds["RU"][0], ds["RU"][2])
    sc = StandardScaler().fit(x1)
    return sc.transform(x1).astype(np.float32), sc.transform(x2).astype(np.float32)


# what a user of the reference runs, with the drop-in installed first; argv follows
RUN_MAIN = ("import sys\n"
            "sys.path.insert(0, {root!r})\n"
            "import linkteller_b200.patch as p\n"
            "p.install(reference_path={ref!r})\n"
            "import runpy\n"
            "sys.argv = ['main.py'] + sys.argv[1:]\n"
            "runpy.run_path({ref!r} + '/main.py', run_name='__main__')\n")

# the same main.py WITHOUT the drop-in (CPU only; import shims for the two modules this image lacks)
RUN_MAIN_UNPATCHED = ("import sys, types\n"
                      "sys.path.insert(0, {ref!r})\n"
                      "import pandas.core.base as pcb\n"
                      "pcb.DataError = type('DataError', (Exception,), {{}})\n"
                      "m = types.ModuleType('tensorboardX')\n"
                      "class SW:\n"
                      "    def __init__(self, *a, **k): pass\n"
                      "    def __getattr__(self, n): return lambda *a, **k: None\n"
                      "m.SummaryWriter = SW\n"
                      "sys.modules['tensorboardX'] = m\n"
                      "import runpy\n"
                      "sys.argv = ['main.py'] + sys.argv[1:]\n"
                      "runpy.run_path({ref!r} + '/main.py', run_name='__main__')\n")
