#!/usr/bin/env python
"""SASS evidence without a GPU: `cuobjdump -sass` of the shipped library, per kernel the instruction count, the most frequent
mnemonics and the Blackwell/Hopper-specific ones the design relies on (TMA bulk copies UBLKCP + SYNCS mbarrier ops, packed
3-input min/max VIMNMX3, byte absolute difference VABSDIFF4, warp reduction REDUX, population count POPC, cp.async LDGSTS).
    python profiles/sass_mix.py > profiles/r02_sass_mix.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mofreak_b200", "libmofreak_b200.so")
MARK = ["UBLKCP", "SYNCS", "UTMALDG", "VIMNMX3", "VIMNMX", "VABSDIFF4", "REDUX", "POPC", "LDGSTS", "LDS", "STS", "LDG", "ATOMS", "SHFL", "VOTE", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            k = re.search(r"(k_\w+(?:<[^>]*>)?)", dem)
            name = k.group(1) if k else dem.split("(")[0]
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    print("# SASS instruction mix of mofreak_b200/libmofreak_b200.so (sm_100a), `cuobjdump -sass`\n")
    print("| kernel | instructions | " + " | ".join(MARK) + " | most frequent |")
    print("|---|---|" + "---|" * len(MARK) + "---|")
    for k, c in kernels.items():
        short = k.replace("mfk::", "").replace("(anonymous namespace)::", "")
        print("| `%s` | %d | " % (short, sum(c.values())) + " | ".join(str(c.get(m, 0)) for m in MARK) +
              " | " + ", ".join("%s %d" % kv for kv in c.most_common(6)) + " |")
    print("\nTMA evidence: `UTMALDG.2D` (cp.async.bulk.tensor.2d, one 38 x 96-byte box per tile from a tensor map in the kernel "
          "parameters) in `k_scores`; `UBLKCP.S.G` (cp.async.bulk global->shared) in `k_front<true>` (two frame bands per block) and "
          "`k_bow_assign` (the codebook); their completion is awaited on mbarriers (`SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK`). "
          "No kernel uses tensor cores (no UTC*MMA / HMMA): no stage of the path is a dense contraction.")
    excerpt = [l for l in sass.splitlines() if re.search(r"UBLKCP|SYNCS\.|VIMNMX3|REDUX", l)]
    print("\n## Excerpt (first occurrences)\n```")
    seen = collections.Counter()
    for l in excerpt:
        op = re.search(r"(UBLKCP[.\w]*|SYNCS[.\w]*|VIMNMX3[.\w]*|REDUX[.\w]*)", l).group(1)
        if seen[op] < 2:
            print(l.strip()[:150]); seen[op] += 1
    print("```")


if __name__ == "__main__":
    main()
