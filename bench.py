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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="large", choices=["large", "flickr", "ppi", "twitch", "small", "tiny"])
    ap.add_argument("--n-test", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=330.0, help="wall-clock cap of the reference arm's probe + warm-up + timed steps (s)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_check / cpu_baseline legs")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="library switch for experiments (lt_set_option), e.g. --opt spmm_hub=0; recorded on the line")
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
    ("large", 1, "spmm_ax"): (12.955522e9 + 0.153716e9, "profiles/r3g_ncu_spmm_summary.txt"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------- reference arm
def bench_config(desc: dict, world: int) -> dict:
    """The `config` object of the JSON line — the same for the product arm and the reference arm."""
    ws = working_set_bytes(desc)
    l2 = (f"inputs (X + Â = {ws / 2**20:.0f} MB) exceed the 126 MB L2; no explicit flush between steps" if not needs_flush(desc)
          else f"inputs (X + Â = {ws / 2**20:.0f} MB) could stay in the 126 MB L2: a 512 MB buffer is written between steps, outside the per-step timers")
    return dict(desc, parallelism=f"rows sharded over {world} GPU(s), graph replicated", l2=l2)


def working_set_bytes(desc: dict) -> int:
    return desc["n"] * desc["f"] * 4 + desc["nnz_adj_hat"] * 8


def needs_flush(desc: dict) -> bool:
    return working_set_bytes(desc) < (512 << 20)


def time_reference(runner, T, steps, warmup, budget_s, step_seconds):
    """Time the reference's loop body (attacker.py:220-229) in steps of `rows_per_step` influence rows.

    One row is probed first (it doubles as warm-up: the CPU path has no clocks to ramp, only pages to fault in);
    rows_per_step is then sized so that a step costs about `step_seconds`, at least one full row and at most one
    whole attack (T rows).  Warm-up and timed steps are cut short only by the wall-clock budget — the counts
    actually run are the ones reported."""
    t_begin = time.time()
    _, t_probe = runner.rows([0])
    rows_per_step = int(max(1, min(T, step_seconds / max(t_probe, 1e-6))))
    est = t_probe * rows_per_step
    nxt = 1

    def take():
        nonlocal nxt
        idx = [(nxt + k) % T for k in range(rows_per_step)]
        nxt = (nxt + rows_per_step) % T
        return idx
    done_w = 1 if rows_per_step == 1 else 0          # the probe row was a whole step
    while done_w < warmup and (time.time() - t_begin) + est * 2 <= budget_s * 0.35:
        runner.rows(take())
        done_w += 1
    times = []
    while len(times) < steps and (not times or (time.time() - t_begin) + est <= budget_s):
        _, dt = runner.rows(take())
        times.append(dt)
        est = sum(times) / len(times)
    return rows_per_step, times, done_w, t_probe


def run_reference(args, rank, world, cuda=False):
    """--impl reference: the UNMODIFIED reference (oracle/_ref: gcn.models.GCN, Attacker.get_gradient_eps_mat and
    the T `.norm().item()` calls per row, attacker.py:100-108, 220-229) on the host cores, on inputs built with
    numpy / torch only (oracle/refgen.py) — the product's CUDA library is never loaded in this process.
    --impl reference-cuda: the same code through the reference's own `.cuda()` placement (what an un-accelerated
    user of the reference gets on this GPU: cuBLAS sgemm + cuSPARSE).  Rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import refgen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    prob = refgen.host_problem(args.workload, n_test=args.n_test)
    runner, kind = refgen.make_runner(prob, device="cuda" if cuda else "cpu", threads=cores)
    step_seconds = min(args.cpu_seconds, args.ref_budget / max(1, args.steps + args.warmup))
    rows_per_step, times, done_w, t_probe = time_reference(runner, prob.T, args.steps, args.warmup, args.ref_budget,
                                                           step_seconds)
    total_rows = rows_per_step * len(times)
    value = total_rows / sum(times)
    sample = (f"{len(times)} timed step(s) of {rows_per_step} full influence row(s) each (2 GCN forwards + {prob.T} norms per row, "
              f"no slicing) after {done_w} warm-up step(s) + 1 probe row; requested steps={args.steps} warmup={args.warmup}"
              + (f", cut by the {args.ref_budget:.0f} s wall-clock budget" if len(times) < args.steps or done_w < args.warmup else "")
              + f"; {runner.what}; inputs: {prob.generator}")
    line = {"impl": "reference-cuda" if cuda else "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": len(times), "warmup": done_w, "ms_per_step": 1e3 * sum(times) / len(times),
            "rows_per_step": rows_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(prob.describe(), args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if cuda:
        line["cpu_baseline"]["device"] = torch.cuda.get_device_name(0)
    print(json.dumps(line), flush=True)


def cpu_legs(wl, info, infl, args):
    """cpu_baseline + parity_check of the product arm's line (rank 0, N = 1): one host copy of the inputs feeds both.
    cpu_baseline: the reference's loop body from oracle/_ref on the host cores, a bounded sample of full rows.
    parity_check: ≥ 16 influence rows of the matrix just timed (hub rows + seeded random rows) against the
    float64 sub-problem oracle (oracle.LocalInfluence) — the oracle is the checker here, never the thing measured."""
    import torch
    from oracle import lt_oracle as O
    from oracle import refgen
    cores = os.cpu_count() or 1
    prob = refgen.host_problem_from_workload(wl)
    out = {}
    # ---- parity_check
    ip = wl.adj_hat.indptr.cpu().numpy()
    rows = O.pick_parity_rows(ip, prob.test_nodes, n_hub=4, n_rand=12)
    t0 = time.time()
    loc = O.LocalInfluence(ip, prob.coo[1], prob.coo[2], prob.X, prob.weights, prob.test_nodes)
    ref, mask, hits = loc.rows(rows, wl.delta)
    got = infl[torch.from_numpy(rows).to(infl.device)].cpu().numpy().astype(np.float64)
    same, rel = O.compare_rows(got, ref)
    out["parity_check"] = {"rows": int(len(rows)), "support_equal": bool(same), "max_rel": rel, "tolerance": 1e-3,
                           "ok": bool(same and rel <= 1e-3), "rows_through_exact_fixup": int((hits > 0).sum()),
                           "oracle": "oracle.LocalInfluence float64 (2-hop sub-problem, pinned to the reference's float64 run)",
                           "seconds": round(time.time() - t0, 1)}
    del loc
    # ---- cpu_baseline
    if not args.no_cpu_baseline:
        runner, kind = refgen.make_runner(prob, device="cpu", threads=cores)
        _, t_probe = runner.rows([0])
        k = int(max(0, min(wl.T - 1, (args.cpu_seconds - t_probe) / max(t_probe, 1e-6))))
        if k >= 1:
            _, dt = runner.rows(list(range(1, 1 + k)))
            value, sample = k / dt, f"{k} full rows after 1 warm-up row"
        else:
            value, sample = 1.0 / t_probe, "1 full row, no warm-up (one row exceeds the sample budget)"
        out["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": f"{sample}; {runner.what}; 2 GCN forwards + {wl.T} norms per row, no slicing"}
    return out


# ----------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl in ("reference", "reference-cuda"):
        return run_reference(args, rank, world, cuda=args.impl == "reference-cuda")

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
    for kv in args.opt:
        name, val = kv.split("=")
        ops.set_option(name, int(val))

    wl = workloads.make(args.workload, device=dev, n_test=args.n_test)
    T = wl.T
    out = torch.empty((T, T), dtype=torch.float32, device=dev)
    group = dist.group.WORLD if world > 1 else None
    comm, exchange = None, "single GPU"
    if world > 1:
        # every rank must take the same branch: agree on whether all of them mapped their peers
        ok = torch.ones(1, device=dev)
        try:
            comm = attack.make_node_comm(group, T, dev, n=wl.n, hidden=wl.h)
        except Exception as e:                                   # no P2P / IPC on this box: NCCL all-gather of the dense rows
            ok.zero_()
            if rank == 0:
                print(f"[bench] peer mapping failed ({e}); falling back to the NCCL all-gather", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            comm = None
        exchange = ("reduce kernel stores into every rank's matrix over NVLink (CUDA IPC peer memory, no collective)" if comm
                    else "NCCL all_gather_into_tensor of the dense rows")

    def step():
        return attack.influence_matrix(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, out=out, group=group,
                                       comm=comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Timing rules: inputs larger than the L2 need no flush (the default workload: X 0.98 GB + Â 1.5 GB against 126 MB);
    # smaller workloads get a 512 MB write between steps, outside the per-step event pairs.
    flush = needs_flush(wl.describe())
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if flush else None

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-resident timed region (phase timers on: CUDA events on the launching stream)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    with ops.profile() as prof:
        barrier()
        if flush:
            evs = []
            for _ in range(args.steps):
                flush_buf.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                evs.append((e0, e1))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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

        def host_step():
            return attack.influence_matrix_host(host["indptr"], host["indices"], host["values"], host["x"], host["w1"],
                                                host["b1"], host["w2"], host["test_nodes"], wl.delta, device=dev,
                                                stage=stage, group=group, comm=comm)
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
        del stage, host

    # several GPUs: the matrix the ranks assembled together must equal, bit for bit, what ONE GPU computes for all T rows
    # (that single-GPU path is the one checked against the float64 oracle at N = 1 and in tests/test_gpu_large.py)
    multi_parity = None
    if world > 1 and not args.no_parity:
        together = step().clone()
        alone = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta)
        torch.cuda.synchronize()
        same = torch.tensor([int(torch.equal(together, alone))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        multi_parity = {"equal_to_single_gpu_bitwise": bool(same.item()), "ranks": world, "rows": T,
                        "note": "every rank compared the jointly assembled T x T matrix with its own single-GPU run of all rows"}
        del together, alone

    if rank == 0:
        peak, peak_src = peaks()
        h, c, f = wl.h, wl.c, wl.f
        Ts = hi - lo
        R, E, M, nnzR = info["frontier_rows"], info["incidences"], info["records"], info["frontier_row_nnz"]
        pairs = info["support_pairs"]
        sign_path = info["exact_path_incidences"] >= 0 and "gemm_h" in phases and f < h and E * h * c >= (1 << 24) and R * h * f >= (1 << 24)
        wpr = (h + 31) // 32
        # algorithmic bytes per launch of each kernel family
        dq_in = E * (4 * wpr + 4 + 12) if sign_path else E * (4 * h + 12)      # sign words + min|H| + (j,ρ,a), or one H row
        alg = {
            "spmm_ax": 8 * nnzR + 8 * R + 4 * f * min(wl.n, nnzR) + 4 * f * R,
            "spmm_h": 8 * nnzR + 8 * R + 4 * h * min(wl.n, nnzR) + 4 * h * R,
            "gemm_h": 4 * (R * f + f * h) + (R * (4 * wpr + 4) if sign_path else 4 * R * h),
            "gemm_p": 4 * (wl.n * f + f * h + wl.n * h),
            "gemm_pv": 4 * (Ts * f + f * h + Ts * h),
            "frontier_dq": dq_in + E * 4 * c + Ts * 4 * h + 4 * h * c,
            "frontier_reduce": M * 16 + E * 4 * c + 4 * pairs,
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
                # the same launch against the bytes it really moves
                roofline["dram_frac_of_peak"] = hit[0] / (dur_ms / 1e3) / 1e9 / peak
        # ---- the frontier-propagation kernels (north_star's "≥ 50 % of the HBM roofline" target): F1 + F2 together
        roofline_frontier = None
        t_front = per_step.get("frontier_dq", 0.0) + per_step.get("frontier_reduce", 0.0)
        if t_front > 0:
            ip = wl.adj_hat.indptr.long()
            deg = ip[1:] - ip[:-1]
            v = wl.test_nodes[lo:hi].long()
            nhat = deg[v]
            starts = ip[v]
            offs = torch.cumsum(nhat, 0) - nhat
            idx = torch.repeat_interleave(starts - offs, nhat) + torch.arange(int(nhat.sum()), device=dev)
            hop2 = int(deg[wl.adj_hat.indices[idx].long()].sum())
            # B_row(v) = 4h(1+|N̂(v)|) + 8|N̂(v)| + 12·Σ_{r∈N̂(v)}|N̂(r)| + 4T, summed exactly over the slice
            survey = int(4 * h * (Ts + int(nhat.sum())) + 8 * int(nhat.sum()) + 12 * hop2 + 4 * T * Ts)
            needed = int(alg["frontier_dq"] + alg["frontier_reduce"])
            roofline_frontier = {
                "bound": "hbm", "kernels": "frontier_dq (F1) + frontier_reduce (F2)", "ms": t_front, "peak": peak, "unit": "GB/s",
                "survey_bytes": survey, "achieved_survey": survey / (t_front / 1e3) / 1e9,
                "frac_survey": survey / (t_front / 1e3) / 1e9 / peak,
                "needed_bytes": needed, "achieved": needed / (t_front / 1e3) / 1e9, "frac": needed / (t_front / 1e3) / 1e9 / peak,
                "note": "survey_bytes = Σ_v B_row(v) of SURVEY.md §8(d), which charges one 4h-byte H row per incidence and 12 B per "
                        "second-hop entry; needed_bytes = what this implementation must move (sign words instead of H rows, "
                        "records only for test-node observers)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(wl.describe(), world),
            "l2_flush_between_steps": bool(flush), "options": list(args.opt), "exchange": exchange,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "library_options": {k: ops.get_option(k) for k in ("overlap", "sign_path", "tc_ws", "spmm_hint", "spmm_hint_deg",
                                                               "spmm_hub", "spmm_segmeta", "shard_rows", "fix_eps")},
            "roofline": roofline, "roofline_frontier": roofline_frontier,
            "phases_ms_per_step": {k: round(v, 4) for k, v in per_step.items()},
            "plan": {k: int(v) for k, v in info.items()},
        }
        if world == 1 and not args.no_parity:
            step()
            torch.cuda.synchronize()
            line.update(cpu_legs(wl, info, out, args))
        if world > 1:
            line["parity_check"] = multi_parity           # computed by every rank above (collective-free), rank 0 reports
        print(json.dumps(line), flush=True)
    if world > 1:
        if comm is not None:
            if comm.timed_out() and rank == 0:
                print("[bench] a peer wait timed out during the run", file=sys.stderr)
            dist.barrier()
            comm.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
