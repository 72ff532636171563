"""Aggregate the `ncu --page source --csv` exports of the chain kernels into profiles/<round>_stalls_chain.json.

usage: python scripts/summarise_stalls.py <capture-suffix[,older-suffix...]> <round>     e.g.  r02g,r02f r02

Reads gpurun_out/prof_chain_{render,train}_<suffix>_source.csv (written by scripts/gpu_profile_r02.sh).  ncu prints
every launch twice; the repeated block is dropped.  Per kernel: share of warp-stall samples per reason, and the instruction mix as the share of
executed warp instructions per SASS mnemonic.
"""
import csv
import json
import os
import re
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def blocks(path):
    name, header, rows = None, None, []
    with open(path, newline="") as f:
        for rec in csv.reader(f):
            if not rec:
                continue
            if rec[0] == "Kernel Name":
                if name is not None:
                    yield name, header, rows
                name, header, rows = rec[1], None, []
            elif header is None:
                header = rec
            else:
                rows.append(rec)
    if name is not None:
        yield name, header, rows


def mnemonic(sass):
    toks = [t for t in sass.strip().split() if not t.startswith("@")]
    return re.split(r"[.\s]", toks[0])[0] if toks else "?"


def summarise(path):
    out, seen = [], Counter()
    for name, header, rows in blocks(path):
        seen[name] += 1
        col = {h: i for i, h in enumerate(header)}
        if "Instructions Executed" not in col:
            continue
        src = col["Source"]
        # the SASS view comes first for each launch; the second block of the same launch is the CUDA-C view
        is_sass = any(re.match(r"\s*(@!?U?P\d+\s+)?[A-Z][A-Z0-9_]*(\.|\s|$)", r[src]) for r in rows[:20])
        if not is_sass:
            continue
        stall_cols = [h for h in header if h.startswith("stall_") and "Not Issued" not in h]
        stalls, mix, insts, samples = Counter(), Counter(), 0, 0
        for r in rows:
            def num(h):
                try:
                    return float(r[col[h]])
                except (ValueError, IndexError):
                    return 0.0
            n = num("Instructions Executed")
            insts += n
            mix[mnemonic(r[src])] += n
            samples += num("# Samples")
            for h in stall_cols:
                stalls[h] += num(h)
        tot = sum(stalls.values()) or 1.0
        if out and out[-1]["kernel"] == name[:72] and out[-1]["warp_instructions"] == int(insts) \
                and out[-1]["samples"] == int(samples):
            continue  # the same launch printed a second time
        out.append({
            "kernel": name[:72],
            "warp_instructions": int(insts),
            "samples": int(samples),
            "stall_share": {k: round(v / tot, 3) for k, v in stalls.most_common(8)},
            "instruction_mix": {k: round(v / (insts or 1.0), 3) for k, v in mix.most_common(12)},
        })
    return out


def main():
    suffix, rnd = sys.argv[1], sys.argv[2]
    doc = {"what": "ncu --set full --import-source on, source page aggregated per kernel "
                   "(scripts/gpu_profile_r02.sh, scripts/summarise_stalls.py): share of warp-stall samples per reason "
                   "and instruction mix (share of executed warp instructions); BRA/ISETP/SYNCS are mostly the "
                   "bounded mbarrier poll loops"}
    for leg in ("render", "train"):
        for sfx in suffix.split(","):     # several captures: the first one that has this leg
            p = os.path.join(ROOT, "gpurun_out", "prof_chain_%s_%s_source.csv" % (leg, sfx))
            if os.path.exists(p):
                doc[leg] = summarise(p)
                doc[leg + "_capture"] = sfx
                break
    dst = os.path.join(ROOT, "profiles", "%s_stalls_chain.json" % rnd)
    with open(dst, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
