#!/usr/bin/env python
"""Print the hottest SASS instructions (warp-stall samples) of an .ncu-rep.
usage: tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[hi + 1:]):
    try:
        data.append((int(r[col["# Samples"]]), idx, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("kernel:", rows[0][1][:100]); print("total samples", tot, " instructions", len(data))
agg = {s: 0 for s in stalls}
for smp, idx, r in data:
    for s in stalls:
        try: agg[s] += int(r[col[s]])
        except ValueError: pass
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for smp, idx, r in sorted(data, reverse=True)[:n]:
    top = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{100*smp/tot:5.2f}%  #{idx:5d} exec={r[col['Instructions Executed']]:>9}  {r[col['Source']].strip()[:70]:70s} {top}")
