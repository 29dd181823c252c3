#!/usr/bin/env python
"""bench.py — influence rows/s of the efficient LinkTeller attack on synthetic graphs (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload large|flickr|ppi|twitch] [--impl reference]

One "step" = one whole efficient attack over the workload: for every node v of the sampled set
(T = n_test), the influence row ‖Δlogits[u]‖₂/δ over all T observers — i.e. T influence rows.
Default workload = config C5 of BASELINE.json (2.45 M nodes, ~62 M undirected edges, f=100, h=256,
c=47, T=8192), the configuration the metric's scaling target is quoted on; it fits one B200.

  value   : rows/s with Â, X, weights and the test nodes already resident in HBM (CUDA events on the
            launching stream, barrier + synchronize on both sides, max over ranks)
  e2e     : rows/s through the host-buffer entry point (linkteller_b200.attack.influence_matrix_host):
            every step copies Â (CSR), X, weights, test nodes from pinned host memory and reads the
            T x T influence matrix back
  roofline: dominant kernel family (largest share of the step, timed live with CUDA events inside
            the library's phase timers), algorithmic bytes per launch from the attack plan
  cpu_baseline (N=1, rank 0): the reference's per-row work on the host cores (oracle port)

N>1 is launched by torchrun (one rank per GPU); perturbed nodes shard across ranks, the graph is
replicated, rows are all-gathered (NCCL).  Total work is fixed ⇒ "scaling": "strong".
Inputs (X 0.98 GB, Â 1.5 GB, H_R ≥ 0.4 GB) exceed the 126 MB L2, so no explicit flush between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "influence rows/sec (perturbed GCN forwards/s)"
UNIT = "rows/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="large", choices=["large", "flickr", "ppi", "twitch", "small", "tiny"])
    ap.add_argument("--n-test", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="wall-clock cap of the reference arm (s)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel, from the
# ncu --set full captures committed under profiles/ (never measured inside a bench run: a run under ncu
# is not a bench value).  Keyed by (workload, n_gpus, kernel family).
NCU_TRAFFIC = {
    ("large", 1, "spmm_ax"): (13.754050e9 + 0.160299e9, "profiles/r1k_ncu_top_summary.txt"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------- reference arm
def host_problem(wl):
    """Host copies of the workload in the form the reference holds them (torch COO pieces, numpy)."""
    import torch
    hat = wl.adj_hat
    indptr = hat.indptr.cpu().numpy()
    row = np.repeat(np.arange(wl.n, dtype=np.int64), np.diff(indptr))
    coo = (row, hat.indices.cpu().numpy().astype(np.int64), hat.values.cpu().numpy())
    weights = [t.cpu().numpy() for t in (wl.w1, wl.b1, wl.w2, wl.b2)]
    return wl.x.cpu().numpy(), coo, weights, wl.test_nodes.cpu().numpy().astype(np.int64)


def cpu_rows_per_s(wl, target_seconds, max_rows=None):
    """The reference's per-row work on the host cores (oracle port; the one place bench.py runs oracle/)."""
    import torch
    from oracle import lt_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    X, coo, weights, tn = host_problem(wl)
    nnz = len(coo[0])
    # rough per-row cost model (s): two forwards, ~0.5 µs per non-zero for the 256-wide COO spmm
    est = 2.0 * (nnz * 0.6e-6 + wl.n * wl.f * wl.h * 2 / 2e11)
    if est > 1.5 * target_seconds:
        slice_div = int(min(64, max(2, round(est / (0.6 * target_seconds)))))
        n_rows = 1
    else:
        slice_div = 1
        n_rows = int(max(1, min(wl.T, target_seconds / max(est, 1e-3))))
    if max_rows:
        n_rows = min(n_rows, max_rows)
    per_row, what = O.time_reference_rows(X, coo, weights, wl.n, tn, wl.delta, n_rows, slice_div=slice_div,
                                          threads=cores)
    return 1.0 / per_row, cores, what


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port — the
    reference is pure Python over torch CPU ops; its modules cannot travel to the GPU box) on the
    host cores, same workload / metric / unit.  Rank 0 only."""
    if rank != 0:
        return
    import torch
    from linkteller_b200 import workloads
    wl = workloads.make(args.workload, device="cuda" if torch.cuda.is_available() else "cpu", n_test=args.n_test)
    t_start = time.time()
    vals, what, cores = [], "", os.cpu_count() or 1
    budget = args.ref_budget
    per_step = max(2.0, min(args.cpu_seconds, budget / max(1, args.steps + min(args.warmup, 1))))
    done_w = 0
    for _ in range(min(args.warmup, 1)):                  # CPU path: one warm-up pass is enough to page in
        cpu_rows_per_s(wl, per_step, max_rows=4)
        done_w += 1
    for _ in range(args.steps):
        v, cores, what = cpu_rows_per_s(wl, per_step)
        vals.append(v)
        if time.time() - t_start > budget:
            break
    value = float(len(vals) / sum(1.0 / v for v in vals))   # total rows / total time
    cfg = wl.describe()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": done_w, "ms_per_step": 1e3 * wl.T / value, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": what},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from linkteller_b200 import attack, ops, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.device_info()

    wl = workloads.make(args.workload, device=dev, n_test=args.n_test)
    T = wl.T
    out = torch.empty((T, T), dtype=torch.float32, device=dev)

    def step():
        return attack.influence_matrix(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-resident timed region (phase timers on: CUDA events on the launching stream)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ops.profile() as prof:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ops.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps * T / (ms / 1e3)
    phases = prof.report()

    # ---- plan statistics → algorithmic bytes
    lo, hi = attack.slice_bounds(T, world, rank)
    _, info = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, lo=lo, hi=hi,
                                   want_info=True)
    torch.cuda.synchronize()

    # ---- end to end through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        stage = attack.HostStage()
        hat = wl.adj_hat
        host = {k: stage.adopt(k, v.cpu(), dev) for k, v in
                (("indptr", hat.indptr), ("indices", hat.indices), ("values", hat.values), ("x", wl.x), ("w1", wl.w1),
                 ("b1", wl.b1), ("w2", wl.w2), ("test_nodes", wl.test_nodes))}
        group = dist.group.WORLD if world > 1 else None

        def host_step():
            return attack.influence_matrix_host(host["indptr"], host["indices"], host["values"], host["x"], host["w1"],
                                                host["b1"], host["w2"], host["test_nodes"], wl.delta, device=dev,
                                                stage=stage, group=group)
        for _ in range(2):
            host_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_e2e = max(3, min(args.steps, 5))
        e0.record()
        for _ in range(k_e2e):
            _, h2d, d2h = host_step()
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": k_e2e * T / (ems / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": k_e2e, "ms_per_step": ems / k_e2e}

    if rank == 0:
        peak, peak_src = peaks()
        h, c, f = wl.h, wl.c, wl.f
        R, E, M, nnzR = info["frontier_rows"], info["incidences"], info["records"], info["frontier_row_nnz"]
        # algorithmic bytes per launch of each kernel family
        alg = {
            "spmm_ax": 8 * nnzR + 8 * R + 4 * f * min(wl.n, nnzR) + 4 * f * R,
            "spmm_h": 8 * nnzR + 8 * R + 4 * h * min(wl.n, nnzR) + 4 * h * R,
            "gemm_h": 4 * (R * f + f * h + R * h),
            "gemm_p": 4 * (wl.n * f + f * h + wl.n * h),
            "gemm_pv": 4 * ((hi - lo) * f + f * h + (hi - lo) * h),
            "frontier_dq": E * (4 * h + 12 + 4 * c) + (hi - lo) * 4 * h + 4 * h * c,
            "frontier_reduce": M * 16 + E * 4 * c + 4 * info["support_pairs"],
        }
        per_step = {k: v[1] / args.steps for k, v in phases.items()}
        kern = {k: v for k, v in per_step.items() if k in alg}
        dom = max(kern, key=kern.get) if kern else None
        roofline = None
        if dom:
            n_launch = phases[dom][0] / args.steps
            dur_ms = per_step[dom] / max(n_launch, 1)
            achieved = alg[dom] / (dur_ms / 1e3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": int(alg[dom]), "launch_ms": dur_ms,
                        "share_of_step": per_step[dom] / (ms / args.steps)}
            if dom in ("spmm_ax", "spmm_h"):
                # SURVEY.md §8d asks for both figures: `achieved` uses the compulsory bytes (every gathered row once),
                # the gather figure counts every gathered row as moved (no reuse) — the bytes the kernel really requests
                d_feat = f if dom == "spmm_ax" else h
                gather = nnzR * (8 + 4 * d_feat) + 8 * R + 4 * d_feat * R
                roofline["gather_bytes_per_launch"] = int(gather)
                roofline["achieved_gather"] = gather / (dur_ms / 1e3) / 1e9
                roofline["frac_gather"] = roofline["achieved_gather"] / peak
            hit = NCU_TRAFFIC.get((args.workload, world, dom)) if args.n_test is None else None
            if hit:
                roofline["traffic"], roofline["traffic_source"] = hit
                # the same launch against the bytes it really moves (gathered rows are re-fetched: L2 hit rate 10 %)
                roofline["dram_frac_of_peak"] = hit[0] / (dur_ms / 1e3) / 1e9 / peak
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(wl.describe(), parallelism=f"rows sharded over {world} GPU(s), graph replicated",
                           l2="inputs (X, Â, H_R) exceed the 126 MB L2; no explicit flush between steps"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "phases_ms_per_step": {k: round(v, 4) for k, v in per_step.items()},
            "plan": {k: int(v) for k, v in info.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, what = cpu_rows_per_s(wl, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": what}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
