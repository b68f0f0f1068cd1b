#!/usr/bin/env python
"""Measure the BASELINE.json configs that are not the headline bench line
(configs[0], [2], [3]) plus the stage kernels, device-resident, with the
library's own CUDA events.  Prints one JSON object (profiles/configs_rNN.json).

  python tools/configs_bench.py [--quick]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import lightmotif_b200 as lm  # noqa: E402
from lightmotif_b200 import pwm  # noqa: E402


def best_ms(fn, pli, reps):
    ts = []
    for _ in range(reps):
        r = fn()
        ts.append(pli.last_kernel_ms)
    return min(ts), float(np.median(ts)), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    scale = 10 if a.quick else 1
    out = {}
    z = np.load(REPO / "tests/golden/jaspar2024.npz")
    widths = z["widths"]
    offs = np.concatenate([[0], np.cumsum(widths)])
    motifs = [z["counts"][offs[i]:offs[i + 1]] for i in range(len(widths))]
    dna = lm.Pipeline.cuda(lm.DNA)
    prot = lm.Pipeline.cuda(lm.PROTEIN)

    # ---- configs[0]: MX000001 (width 15) over the 4.6 Mbp E. coli genome
    from conftest import load_ecoli_codes
    codes = load_ecoli_codes()
    kat = json.loads((REPO / "tests/golden/kat.json").read_text())["dna"]
    pssm = pwm.pssm_from_sequences(pwm.DNA, kat["patterns"], kat["pseudocount"])
    seq = dna.sequence_from_symbols(codes, len(pssm) - 1)
    b, med, res = best_ms(lambda: seq.score_argmax(pssm), dna, 20)
    t0 = time.perf_counter()
    for _ in range(20):
        seq.score_argmax(pssm)
    wall = (time.perf_counter() - t0) / 20
    out["config0_ecoli_mx000001"] = {"bp": int(codes.size), "kernel_ms": b, "kernel_ms_median": med,
                                     "gbp_s_kernel": codes.size / b / 1e6, "call_ms_wall": wall * 1e3,
                                     "gbp_s_call": codes.size / wall / 1e9, "result": res,
                                     "note": "fused score+argmax, resident sequence; full-genome max 18.986465 @ 1112931 (SURVEY 8c)"}
    b, med, sc = best_ms(lambda: seq.score(pssm), dna, 5)
    out["config0_ecoli_mx000001"]["materialised_scores_kernel_ms"] = b
    out["config0_ecoli_mx000001"]["materialised_gbp_s_kernel"] = codes.size / b / 1e6
    seq.free()

    # ---- configs[2]: 1024 JASPAR motifs (widths 6..30) over 100 Mbp, per-motif argmax
    sel = [m for m in motifs if 6 <= len(m) <= 30][:1024]
    pssms = [pwm.pssm_from_counts(pwm.DNA, c, 0.1) for c in sel]
    n = 100_000_000 // scale
    maxw = max(len(p) for p in pssms)
    seq = dna.sequence_synthetic(n, maxw - 1)
    b, med, res = best_ms(lambda: seq.score_argmax_batch(pssms), dna, 2)
    cells = float(n) * len(pssms)
    adds = float(n) * sum(len(p) for p in pssms)
    out["config2_many_motifs"] = {"bp": n, "motifs": len(pssms), "mean_width": adds / cells, "kernel_ms": b,
                                  "gbp_s_sequence": n / b / 1e6, "gcells_s": cells / b / 1e6,
                                  "gadds_s": adds / b / 1e6,
                                  "fp32_add_ceiling_gadds_s": 148 * 128 * 1.965,
                                  "lds_ceiling_glookups_s": 148 * 32 * 1.965,
                                  "note": "per-motif planning: filter + recheck kernel where profitable, exact per-cell f32 "
                                          "sums otherwise (bound by the shared-memory look-up pipe, 1 LDS per cell update)"}
    out["config2_many_motifs"]["motifs_through_filter"] = int(dna.last_stats["pruned"])
    t0 = time.perf_counter()
    seq.score_argmax_batch(pssms)
    out["config2_many_motifs"]["call_s_wall"] = time.perf_counter() - t0
    dna.set_prune(0)
    b0, _, res0 = best_ms(lambda: seq.score_argmax_batch(pssms), dna, 1)
    dna.set_prune(1)
    out["config2_many_motifs"]["exact_only_kernel_ms"] = b0
    assert [(r["position"], r["score"]) for r in res0] == [(r["position"], r["score"]) for r in res]
    # per-motif thresholds 3 bits below each motif's maximum (the CLI's p-value cut-offs, roughly)
    thr = [float(np.float32(r["score"] - 3.0)) for r in res]
    seq.score_threshold_batch(pssms, thr, cap=1 << 14, counts_only=True)
    t0 = time.perf_counter()
    counts = seq.score_threshold_batch(pssms, thr, cap=1 << 14, counts_only=True)
    out["config2_many_motifs"]["threshold_batch"] = {"kernel_ms": dna.last_kernel_ms, "call_s_wall": time.perf_counter() - t0,
                                                     "hits": int(sum(counts)), "thresholds": "per-motif maximum - 3.0",
                                                     "motifs_through_filter": int(dna.last_stats["pruned"])}
    # spot check 3 motifs against single calls
    for i in (0, len(pssms) // 2, len(pssms) - 1):
        single = seq.score_argmax(pssms[i])
        assert (single["position"], single["score"]) == (res[i]["position"], res[i]["score"]), (i, single, res[i])
    seq.free()

    # ---- configs[3]: protein width-15 PSSM over 500 M residues
    pk = json.loads((REPO / "tests/golden/kat.json").read_text())["protein"]
    pp = pwm.pssm_from_sequences(pwm.PROTEIN, pk["patterns"], pk["pseudocount"])[:15]
    n = 500_000_000 // scale
    seq = prot.sequence_synthetic(n, 14)
    prot.set_prune(0)
    b, med, res = best_ms(lambda: seq.score_argmax(pp), prot, 3)
    out["config3_protein"] = {"residues": n, "exact_kernel_ms": b, "exact_gres_s": n / b / 1e6, "result": res,
                              "note": "exact = f32 sum at every cell (K = 21 table in shared memory); filter = 8 residue "
                                      "classes chosen per matrix, class-coded copy of the sequence (one extra pass, cached "
                                      "per sequence + matrix), PRMT filter on the class codes, exact recheck from the residues"}
    prot.set_prune(1)
    r1 = seq.score_argmax(pp)               # allocates the class-copy buffer
    prot.set_class_cache(0)
    first_ms, _, r1 = best_ms(lambda: seq.score_argmax(pp), prot, 3)   # class-map pass + filter on every call
    prot.set_class_cache(1)
    b, med, r2 = best_ms(lambda: seq.score_argmax(pp), prot, 3)   # class copy cached
    assert (r1["position"], r1["score"]) == (res["position"], res["score"]) == (r2["position"], r2["score"])
    out["config3_protein"].update({"filter_uncached_kernel_ms": first_ms, "filter_uncached_gres_s": n / first_ms / 1e6,
                                   "filter_cached_kernel_ms": b, "filter_cached_gres_s": n / b / 1e6,
                                   "filter_stats": prot.last_stats,
                                   "hbm_frac_of_measured_cached": n / b / 1e6 / 6555.5})
    seq.free()

    # ---- stage kernels: ASCII -> encode + stripe + wrap in one pass
    n = 1_000_000_000 // scale
    rng = np.random.default_rng(1)
    ascii_ = np.frombuffer(b"ACTG", np.uint8)[rng.integers(0, 4, min(n, 64_000_000))]
    ascii_ = np.tile(ascii_, n // ascii_.size + 1)[:n].copy()
    t0 = time.perf_counter()
    s = dna.sequence_from_ascii(ascii_, 19)
    wall = time.perf_counter() - t0
    out["encode_stripe_wrap_from_ascii"] = {"bp": n, "kernel_ms": dna.last_kernel_ms,
                                            "gbp_s_kernel": n / dna.last_kernel_ms / 1e6,
                                            "algorithmic_gb_s": 2 * n / dna.last_kernel_ms / 1e6,
                                            "call_s_wall_incl_h2d_pageable": wall}
    s.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
