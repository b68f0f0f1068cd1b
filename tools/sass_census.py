#!/usr/bin/env python
"""Per-kernel census of the SASS mnemonics that matter for this engine (bulk TMA, mbarrier, byte-lane
integer ops, packed FP32 adds, shared-memory look-ups), from the built library.
usage: python tools/sass_census.py > profiles/sass_census_rNN.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "lightmotif_b200" / "liblightmotif_b200.so"
WATCH = ["UBLKCP", "UBLKPF", "SYNCS", "NANOSLEEP", "PRMT", "IDP", "IMAD", "LOP3", "LDS", "STS", "FADD", "FADD2", "FFMA", "ATOM",
         "ATOMS", "RED", "LDG", "STG", "VOTE", "SHFL", "UTMALDG", "UTCMMA", "HMMA", "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
kern, counts, totals = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        totals[kern] = 0
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
        totals[kern] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS census of {LIB.name} (sm_100a): instructions per kernel, selected mnemonics")
print("# no UTMALDG / UTC*MMA / HMMA anywhere: contiguous tiles are staged with 1-D bulk TMA (UBLKCP), the work is byte/integer")
for (k, c), name in zip(counts.items(), demangle):
    name = re.sub(r"\(lm::.*", "", name).replace("void lm::(anonymous namespace)::", "")
    sel = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{name:55s} total={totals[k]:5d}  {sel}")
