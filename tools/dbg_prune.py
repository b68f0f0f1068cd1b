import ctypes as C, sys, os
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lightmotif_b200 as lm
from lightmotif_b200 import capi
from oracle import oracle as O
z = np.load(Path(__file__).resolve().parent.parent / "tests/golden/jaspar2024.npz")
pssm = O.pssm_from_counts(O.DNA, z["rest"], 0.1)
pli = lm.Pipeline.cuda(lm.DNA)
bp = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000_000
seq = pli.sequence_synthetic(bp, 19)
for rep in range(2):
    res = seq.score_argmax(pssm)
    print(pli.last_kernel_ms, pli.last_stats)
n = 444
buf = (C.c_uint64 * (4 * n))()
capi.lib().lm_debug_dump(buf, n)
a = np.frombuffer(buf, dtype=np.uint64).reshape(n, 4)
cyc = a[:, 0].astype(np.float64); cand = a[:, 1]; smid = (a[:, 2] & 0xffffffff).astype(int); its = (a[:, 2] >> 32).astype(int)
start = a[:, 3].astype(np.float64); start -= start.min()
print("cycles: min %.0f med %.0f max %.0f" % (cyc.min(), np.median(cyc), cyc.max()))
print("iters: ", np.bincount(its))
print("start ns: med %.0f max %.0f" % (np.median(start), start.max()))
order = np.argsort(cyc)
print("slowest CTAs (bid, smid, cyc, cand, its, start):")
for i in order[-12:]: print(i, smid[i], int(cyc[i]), int(cand[i]), its[i], int(start[i]))
print("fastest:")
for i in order[:6]: print(i, smid[i], int(cyc[i]), int(cand[i]), its[i], int(start[i]))
print("corr(cyc,cand)=%.3f" % np.corrcoef(cyc, cand.astype(float))[0, 1])
per_sm = {}
for i in range(n): per_sm.setdefault(smid[i], []).append(cyc[i])
tot = np.array([max(v) for v in per_sm.values()]); cnt = np.bincount([len(v) for v in per_sm.values()])
print("CTAs per SM histogram:", cnt, " per-SM max cycles: min %.0f med %.0f max %.0f" % (tot.min(), np.median(tot), tot.max()))
