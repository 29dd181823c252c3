#!/usr/bin/env python
"""Turn raw ncu output (gpurun_out/) into the text summaries committed under profiles/.

    python profiles/summarize.py launches <launches.csv> [--last-step]      # per-kernel time shares
    python profiles/summarize.py full <report.ncu-rep> [kernel-regex]       # key counters per launch
    python profiles/summarize.py stalls <report.ncu-rep> [n]                # stall reasons + hottest SASS lines

`launches` reads the CSV of  ncu --metrics gpu__time_duration.sum --clock-control none --csv ;
with --last-step only the launches after the last lt::k_test_lens (= the last attack in the capture)
are counted.  Only kernels of this library (namespace lt::) are listed individually.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name


def launches(path, last_step):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        rows.append((short(r["Kernel Name"]), v))
    if last_step:
        idx = [i for i, (k, _) in enumerate(rows) if k.startswith("lt::k_test_lens")]
        if idx:
            rows = rows[idx[-1]:]
    agg = OrderedDict()
    for k, v in rows:
        if not k.startswith("lt::"):
            k = "(torch / other)"
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# launches: {len(rows)}   sum of kernel time: {tot:.1f} us")
    print(f"{'kernel':<52}{'launches':>9}{'us':>11}{'share':>8}")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<52}{n:>9}{v:>11.1f}{100 * v / tot:>7.1f}%")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, pattern):
    r = ncu_csv(rep, "raw")
    hdr, units = r[0], r[1]
    for row in r[2:]:
        d = dict(zip(hdr, row))
        if pattern and not re.search(pattern, d.get("Kernel Name", "")):
            continue
        print(f"kernel: {d.get('Kernel Name', '')[:100]}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:<68}{d[k]:>16} {units[hdr.index(k)]}")
        print()


def stalls(rep, n):
    r = ncu_csv(rep, "source")
    # one table per kernel: a "Kernel Name" line, a header line, then instructions
    i = 0
    while i < len(r):
        if r[i] and r[i][0] == "Kernel Name":
            name, hdr = r[i][1], r[i + 1]
            j = i + 2
            body = []
            while j < len(r) and not (r[j] and r[j][0] == "Kernel Name"):
                if r[j]:
                    body.append(r[j])
                j += 1
            ix = {h: k for k, h in enumerate(hdr)}
            st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

            def num(x):
                try:
                    return int(x)
                except ValueError:
                    return 0
            tot = {s: sum(num(b[ix[s]]) for b in body) for s in st}
            T = max(1, sum(tot.values()))
            print(f"kernel: {name[:100]}")
            print("  stall reasons (all samples): " +
                  ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
            for b in sorted(body, key=lambda b: -num(b[ix["# Samples"]]))[:n]:
                top = sorted(st, key=lambda s: -num(b[ix[s]]))[:2]
                print(f"  {num(b[ix['# Samples']]):>8}  {b[ix['Source']][:84]:<84} {top[0][6:]},{top[1][6:]}")
            print()
            i = j
        else:
            i += 1


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], "--last-step" in sys.argv)
    elif mode == "full":
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    elif mode == "stalls":
        stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 12)
