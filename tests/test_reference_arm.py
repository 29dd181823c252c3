"""The reference arm of bench.py: inputs built without the product library (oracle/refgen.py) and the UNMODIFIED
reference from oracle/_ref (oracle/build.py) driving them.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import lt_oracle as O
from oracle import refgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not refgen.have_ref(), reason="oracle/_ref not built (python oracle/build.py)")


def test_oracle_ref_copy_matches_its_manifest():
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_build", os.path.join(ROOT, "oracle", "build.py"))
    ob = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ob)
    if not os.path.isdir(ob.DST):
        pytest.skip("oracle/_ref not built")
    assert ob.verify()
    if os.path.isdir(ob.SRC):                      # build container: byte-identical to /root/reference
        for f in ob._files(ob.SRC):
            assert open(os.path.join(ob.SRC, f), "rb").read() == open(os.path.join(ob.DST, f), "rb").read(), f


@pytest.mark.parametrize("case", ["tiny_fo", "small_hi"])
def test_refgen_inputs_equal_the_golden_inputs(golden, case):
    g = golden(case)
    p = refgen.host_problem(str(g["shape"]), n_test=int(g["T"]), seed=int(g["seed"]), hidden=int(g["hidden"]))
    assert np.array_equal(p.coo[0], g["a2_row"]) and np.array_equal(p.coo[1], g["a2_col"])
    assert np.array_equal(p.coo[2].view(np.uint32), g["a2_val"].view(np.uint32))
    if str(g["sample_type"]) == "unbalanced":
        assert np.array_equal(p.test_nodes, g["a6_test_nodes"])


def test_torch_only_graph_builder_matches_the_oracle():
    """The C5 inputs of the reference arm come from torch ops (sort/unique + float32 scaling); on a small graph with
    repeated edges they must equal the numpy oracle's binarised FirstOrderGCN Â bit for bit."""
    from linkteller_b200 import synth
    n, m = 3000, 20000
    edges = synth.chung_lu_edges_device(n, m, 3, "cpu")
    row, col = refgen.csr_from_edges_torch(edges, n)
    ip, ix, dv = O.csr_from_edges(edges.numpy(), n)
    assert np.array_equal(ix, col.numpy().astype(np.int32))
    r2, c2, v2 = refgen.normalize_first_order_torch(row, col, n)
    rr, cc, vv = O.normalize_adj(ip, ix, np.ones_like(dv), n, "FirstOrderGCN")      # binarised, float32 arithmetic
    assert np.array_equal(rr, r2.numpy()) and np.array_equal(cc, c2.numpy())
    assert np.array_equal(vv.view(np.uint32), v2.numpy().view(np.uint32))


@needs_ref
def test_reference_runner_reproduces_the_golden_rows(golden):
    g = golden("tiny_fo")
    p = refgen.host_problem(str(g["shape"]), n_test=int(g["T"]), seed=int(g["seed"]), hidden=int(g["hidden"]))
    r32, kind = refgen.make_runner(p, "cpu")
    assert kind == "reference"
    out, _ = r32.rows(range(6))
    assert np.array_equal(out, g["a8_influence32"][:6])                 # the reference as shipped, bit for bit
    out64, _ = refgen.ReferenceRunner(p, "cpu", dtype=torch.float64).rows(range(6))
    assert np.array_equal(out64, g["a8_influence64"][:6])
    port, _ = refgen.PortRunner(p, "cpu").rows(range(6))
    np.testing.assert_allclose(port, out, rtol=1e-6, atol=0)


@needs_ref
def test_bench_reference_arm_prints_the_contract_line_and_never_loads_the_product_library():
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py','--impl','reference','--workload','tiny','--steps','3','--warmup','1','--ref-budget','20']\n"
            "runpy.run_path('bench.py', run_name='__main__')\n"
            "print('PRODUCT_SO', 'liblinkteller' in open('/proc/self/maps').read())\n")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 3 and line["value"] > 0 and line["config"]["workload"] == "tiny"
    assert "PRODUCT_SO False" in res.stdout
    # under torchrun only rank 0 works
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "tiny"], cwd=ROOT,
                         env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True, text=True,
                         timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""
