#!/usr/bin/env python
"""SASS excerpts of the hot kernels of libb200join.so (sm_100a): instruction mix and the first occurrence, in program
order, of every memory / synchronisation instruction kind.   python scripts/sass_excerpt.py > profiles/r2_sass_scatter.txt"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "sigmod-2018_b200" / "lib" / "libb200join.so"
WANT = [
    ("scatter (probe side, histogram-free, carried SUM column)", r"radix_scatter_kernelILi512ELi32ELi1EjLb1ELb1ELb0E"),
    ("scatter (probe side, histogram-free, fused predicates)", r"radix_scatter_kernelILi512ELi32ELi1EjLb1ELb0ELb1E"),
    ("join (32-bit keys, fused SUM, 2 projections)", r"tag_join_kernelILi1024ELi1ELi4ELi2ELi2ELb0ELb0E"),
    ("join (32-bit keys, segmented build side of the multi-GPU broadcast: waits on arrival flags)", r"tag_join_kernelILi1024ELi1ELi4ELi2ELi2ELb1ELb0E"),
    ("join (64-bit keys, fused SUM)", r"tag_join_kernelILi768ELi1ELi2ELi2ELi2ELb0ELb1E"),
    ("fetch kernel of the pull broadcast (TMA bulk copies over NVLink, mbarrier ring)", r"pull_regions_kernel"),
    ("exchange kernel (stores into the owners' receive buffers over NVLink)", r"segment_exchange2_kernel"),
]
MEM = re.compile(r"\b(LDG|STG|LDS|STS|ATOMS|ATOMG|RED|CCTL|BAR|UBLKCP|UTMA\w*|SYNCS|LDGSTS|MATCH|ELECT|NANOSLEEP|MEMBAR|FENCE|LD\.|ST\.|LD |ST )")
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
funcs = {}
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
    elif name and re.search(r"/\*[0-9a-f]{4}\*/", line):
        funcs[name].append(line)
print("# cuobjdump -sass sigmod-2018_b200/lib/libb200join.so (sm_100a), scripts/sass_excerpt.py: instruction mix of the hot kernels and")
print("# the first occurrences of their memory / synchronisation instructions in program order.  TMA (UBLKCP + SYNCS = mbarrier)")
print("# appears in the fetch kernel only; DESIGN.md §4 says why the scatter keeps register prefetches (LDG.E.128 issued a")
print("# copy-out phase ahead) and 64-bit stores, and why MATCH.ANY would only add instructions with 2^10 bins and 32 lanes.")
for title, pat in WANT:
    hits = [n for n in funcs if re.search(pat, n)]
    if not hits:
        print(f"\n## {title}\n#  (no kernel matching {pat})")
        continue
    n = hits[0]
    body = funcs[n]
    ops = collections.Counter()
    first = []
    seen = set()
    for line in body:
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(2).strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
        if MEM.search(op + " "):
            ops[op] += 1
            if op not in seen and len(first) < 40:
                seen.add(op)
                first.append(f"    /*{m.group(1)}*/ {ins}")
    print(f"\n## {title}\n#  {n[:70]}...  {len(body)} SASS instructions")
    print("#  memory / sync instructions (static counts): " + ", ".join(f"{k} x{v}" for k, v in ops.most_common(16)))
    print("#  first occurrence of each kind in program order:")
    print("\n".join(first))
