#!/usr/bin/env python
"""SASS census of the built library -> profiles/<round>_sass_summary.md (runs without a GPU: cuobjdump only).

    python scripts/sass_summary.py r02

Per kernel: registers, tensor-core (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2), tensor-memory (LDTM / STTM =
tcgen05.ld / st), bulk-copy (UBLKCP = cp.async.bulk; UTMALDG would be tensor-map TMA), commit (UTCBAR = tcgen05.commit)
and mbarrier (SYNCS) instruction counts -- the mnemonics /opt/skills/guides/B200_PROFILING.md names as proof of the
tcgen05 / TMA path.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "moco_flow_b200", "csrc", "libmoco_flow_b200.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP.S.G", "UBLKCP.G.S", "UTMALDG", "UTCBAR", "SYNCS", "FADD2",
        "F2FP", "USETMAXREG", "LDL+STL"]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    demangled = dict(zip(regs, subprocess.run(["c++filt"], input="\n".join(regs), capture_output=True,
                                              text=True).stdout.splitlines()))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            c = counts[cur]
            base = op.split(".")[0]
            c[base] += 1
            if op.startswith("UTCHMMA.2CTA"):
                c["UTCHMMA.2CTA"] += 1
            if op.startswith("UBLKCP.G.S") or op.startswith("UBLKCP.S.G"):
                c[op[:10]] += 1
            if base in ("LDL", "STL"):
                c["LDL+STL"] += 1
    out = os.path.join(ROOT, "profiles", f"{rnd}_sass_summary.md")
    total = collections.Counter()
    with open(out, "w") as f:
        f.write(f"# SASS census of `libmoco_flow_b200.so` ({rnd}, `scripts/sass_summary.py`, sm_100a)\n\n"
                "Static instruction counts per kernel (`cuobjdump -sass`).  UTCHMMA = `tcgen05.mma` (`.2CTA` = "
                "`cta_group::2`), LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UBLKCP = `cp.async.bulk` (S.G = into shared "
                "from global: weights; G.S = into global from shared: training saves), UTMALDG = tensor-map TMA (not used: the operand images are pre-swizzled "
                "and contiguous, DESIGN.md 4.4), UTCBAR = `tcgen05.commit`, SYNCS = mbarrier operations, USETMAXREG = "
                "`setmaxnreg`.  LDL+STL are local-memory accesses: the slow path of `sincosf`, the bounded-wait trap "
                "paths and a 52-byte spill in the training-forward k_chain; executed share in the r02g capture: 0.39 % "
                "of the warp instructions of k_chain<256,2,fwd,save>, 0.02 % of the backward (3.5 % before the mask "
                "words were selected from registers), 0 in k_nof forward.\n\n")
        f.write("| kernel | regs | " + " | ".join(COLS) + " |\n|---|---:|" + "---:|" * len(COLS) + "\n")
        for fn, c in counts.items():
            if not any(c[k] for k in ("UTCHMMA", "LDTM", "UBLKCP")):
                continue
            name = demangled.get(fn, fn).replace("mcf::", "").replace("(mcf_chain_params_t)", "")
            name = name.replace("(bool)", "").replace("(int)", "").replace("void ", "")
            f.write(f"| `{name[:70]}` | {regs.get(fn, '')} | " + " | ".join(str(c[k]) for k in COLS) + " |\n")
            for k in COLS:
                total[k] += c[k]
        f.write("| **all tensor-core kernels** | | " + " | ".join(f"**{total[k]}**" for k in COLS) + " |\n")
        others = [demangled.get(fn, fn) for fn, c in counts.items() if not any(c[k] for k in ("UTCHMMA", "LDTM", "UBLKCP"))]
        f.write(f"\n{len(others)} further kernels (fp32 HBM-bound glue, no tensor-core or bulk-copy instructions): "
                + ", ".join(sorted({re.sub(r'<.*', '', o.replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
                                           .replace('void ', '').replace('mcf::', '').split('(')[0]) for o in others})) + ".\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
