#!/usr/bin/env python
"""Condense gpurun_out/ ncu exports into small tracked files under profiles/ (per round).

    python scripts/summarise_profiles.py <suffix used on the GPU box> [<prefix under profiles/>]
"""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
]


def read_csv(path):
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    return list(csv.reader(io.StringIO("".join(lines))))


def launch_list(path, tag, rnd):
    rows = read_csv(path)
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = r[idx["Kernel Name"]].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    out = os.path.join(DST, f"{rnd}_launches_{tag}.md")
    with open(out, "w") as f:
        f.write(f"# ncu launch list, `bench.py --workload {tag} --steps 2 --warmup 3 --no-graph` ({rnd})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over the whole process "
                "(3 warm-up + 2 timed + 2 e2e + 2 profiled steps).  Times are cold-cache and serialised: "
                "compare shares, not absolutes.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:90]}` | {n} | {v:.1f} | {100*v/total:.2f}% |\n")
    print("wrote", out)


def full(path, tag, rnd):
    rows = read_csv(path)
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        item = {"kernel": d.get("Kernel Name", "")[:80], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k in KEEP:
            if k in d:
                item[k] = f"{d[k]} {u[k]}".strip()
        res.append(item)
    out = os.path.join(DST, f"{rnd}_ncu_full_{tag}.json")
    json.dump(res, open(out, "w"), indent=1)
    print("wrote", out)


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "r01"   # suffix used by scripts/gpu_profile.sh
    rnd = sys.argv[2] if len(sys.argv) > 2 else src     # prefix of the files written under profiles/
    os.makedirs(DST, exist_ok=True)
    for tag in ("train", "render"):
        p = os.path.join(OUT, f"launches_{tag}_{src}.csv")
        if os.path.exists(p):
            launch_list(p, tag, rnd)
    for tag in ("chain_render", "chain_train", "dw", "render_ops"):
        p = os.path.join(OUT, f"prof_{tag}_{src}_raw.csv")
        if os.path.exists(p):
            full(p, tag, rnd)


if __name__ == "__main__":
    main()
