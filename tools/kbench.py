#!/usr/bin/env python
"""Kernel micro-bench used during development: fused score+argmax on synthetic
DNA, device-resident, timed with the library's own CUDA events.

  LM_B200_LIB=path/to/variant.so python tools/kbench.py [--bp N] [--reps R] [--m M]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lightmotif_b200 as lm  # noqa: E402
from lightmotif_b200 import pwm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--m", type=int, default=20)
    ap.add_argument("--mode", default="argmax")
    ap.add_argument("--tag", default="")
    ap.add_argument("--prune", type=int, default=1)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--fmode", type=int, default=0)
    ap.add_argument("--boot", type=int, default=0)
    ap.add_argument("--thr", type=float, default=18.0)
    a = ap.parse_args()
    z = np.load(Path(__file__).resolve().parent.parent / "tests/golden/jaspar2024.npz")
    pssm = pwm.pssm_from_counts(pwm.DNA, z["rest"], 0.1)
    if a.m != 20:
        rng = np.random.default_rng(a.m)
        pssm = np.concatenate([pssm] * (a.m // 20 + 1))[: a.m].copy()
    pli = lm.Pipeline.cuda(lm.DNA)
    pli.set_prune(a.prune)
    pli.set_filter_rows(a.rows)
    pli.set_filter_mode(a.fmode)
    if a.boot:
        pli.set_boot_rows(a.boot)
    seq = pli.sequence_synthetic(a.bp, a.m - 1)
    times = []
    res = None
    for _ in range(a.reps):
        if a.mode == "argmax":
            res = seq.score_argmax(pssm)
        elif a.mode == "threshold":
            res = len(seq.score_threshold(pssm, a.thr))
        times.append(pli.last_kernel_ms)
    best = min(times)
    print(f"{a.tag:24s} mode={a.mode} m={a.m} bp={a.bp:.3g}  best {best:.3f} ms  median {np.median(times):.3f} ms  "
          f"-> {a.bp / best / 1e6:.1f} Gbp/s  result={res} stats={pli.last_stats}")


if __name__ == "__main__":
    main()
