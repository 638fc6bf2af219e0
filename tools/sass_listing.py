#!/usr/bin/env python
"""Regenerates profiles/rNN_sass_listing.md from the built library: per kernel that uses Blackwell-specific machinery, the
tcgen05 / TMEM / TMA / mbarrier / cluster mnemonics with counts and first occurrences (no GPU needed).

    python tools/sass_listing.py [ee_semantic_segmentation_b200/libeeseg_b200.so] > profiles/r02_sass_listing.md
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "ee_semantic_segmentation_b200/libeeseg_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = re.compile(r"^(UTC|UTMA|LDTM|STTM|SYNCS|UCGABAR|ELECT|ACQBULK|FENCE\.VIEW\.ASYNC|MATCH|REDUX|UBLKCP|CCTL)")
INTERESTING = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UBLKCP")

print(f"# SASS evidence per kernel (round 2, final library): `cuobjdump -sass {lib}`\n")
print("For every kernel that uses Blackwell-specific machinery: the distinct tcgen05 / TMEM / TMA / mbarrier / cluster mnemonics with their\n"
      "occurrence counts, followed by the first occurrence of each in the listing (address + instruction) so the claim can be checked against a\n"
      "fresh `nvcc -gencode arch=compute_100a,code=sm_100a` build of the same source (`python tools/sass_listing.py`). Mnemonic key: `UTCHMMA` =\n"
      "tcgen05.mma (`.2CTA` = cta_group::2), `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTCATOMSWS` = tcgen05.alloc/dealloc, `UTMALDG` /\n"
      "`UTMASTG` = TMA tensor load / store, `SYNCS` = mbarrier ops, `UCGABAR_*` = barrier.cluster, `MATCH` / `REDUX` = warp match / reduce.\n")
funcs = re.split(r"\n\s*Function : ", sass)[1:]
for f in funcs:
    name, _, body = f.partition("\n")
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem)
    counts, first = collections.OrderedDict(), {}
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(2).strip()
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        if PAT.match(op):
            counts[op] = counts.get(op, 0) + 1
            first.setdefault(op, f"/*{m.group(1)}*/ {ins}")
    if not any(k.startswith(INTERESTING) for k in counts):
        continue
    print(f"## `{dem}`\n\n| mnemonic | count | first occurrence |\n|---|---|---|")
    for op, n in counts.items():
        print(f"| `{op}` | {n} | `{first[op][:110]}` |")
    print()
