"""CPU property tests of the filter planner (lm_plan_filter, no GPU): for every sampled window the
integer upper bound the kernel computes must dominate the score,

    c0 + delta * sum_q(window) >= score(window)                (real arithmetic)
    sum_q(window) >= ceil((s32 - eps - c0) / delta - 1e-6)        (what the kernel tests, with T = s32)

where s32 is the reference's sequential f32 sum -- i.e. a position can never be filtered out at a
threshold its own exact score reaches.  This is what makes the fused results bit-identical to scoring
every cell; the GPU tests check the outcome, this checks the bound itself (both look-up flavours,
-inf cells, finite N weights, N in the window, windows hanging over the sequence end)."""
import math

from pathlib import Path

import numpy as np
import pytest

import lightmotif_b200 as lm
from oracle import oracle as O


REPO = Path(__file__).resolve().parent.parent


def make_pssm(rng, m, kind):
    if kind == "logodds":
        counts = rng.integers(0, 50, size=(m, 5)).astype(np.uint32)
        counts[:, 4] = 0
        counts[rng.integers(0, m, m // 2), rng.integers(0, 4, m // 2)] += 400
        return O.pssm_from_counts(O.DNA, counts, 0.1)
    if kind == "neginf":
        counts = rng.integers(0, 3, size=(m, 5)).astype(np.uint32)
        counts[:, 4] = 0
        counts[np.arange(m), rng.integers(0, 4, m)] += 30
        return O.pssm_from_counts(O.DNA, counts, 0.0)
    if kind == "ints":
        p = rng.integers(-3, 3, size=(m, 5)).astype(np.float32)
        p[:, 4] = -np.inf
        return p
    if kind == "finite_n":
        return rng.normal(0, 2, size=(m, 5)).astype(np.float32)
    if kind == "rest":
        return O.pssm_from_counts(O.DNA, np.load(REPO / "tests/golden/jaspar2024.npz")["rest"], 0.1)[:m]
    raise ValueError(kind)


def sum_q(plan, win):
    """sum of the table entries the kernel looks up for the window `win` (symbols of motif rows 0..m-1)."""
    tot = np.zeros(win.shape[0], dtype=np.int64)
    m = win.shape[1]
    for k in range(plan["r"]):
        tab = [(plan["tab_lo"][k] >> (8 * b)) & 0xFF for b in range(4)] + [(plan["tab_hi"][k] >> (8 * b)) & 0xFF for b in range(4)]
        tab = np.array(tab, dtype=np.int64)
        if not plan["pair"]:
            entry = win[:, plan["j0"] + k].astype(np.int64)  # 0..4 (N = 4 -> tab_hi byte 0)
        else:
            s1 = win[:, plan["j0"] + 2 * k].astype(np.int64)
            j2 = plan["j0"] + 2 * k + 1
            s2 = win[:, j2].astype(np.int64) if j2 < m else np.full(win.shape[0], 4, dtype=np.int64)
            cls = ((s2 * plan["pair_mul"]) >> 2) & 1
            entry = s1 | (cls << 2)  # N (4) as first symbol: 4 | 4 * cls = 4
        tot += tab[entry]
    return tot


@pytest.mark.parametrize("m,kind", [(4, "logodds"), (6, "ints"), (8, "logodds"), (11, "neginf"), (15, "logodds"),
                                    (20, "rest"), (20, "finite_n"), (21, "ints"), (33, "logodds"), (12, "finite_n")])
@pytest.mark.parametrize("mode", [1, 2])
def test_bound_dominates_score(m, kind, mode):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{m}-{kind}-{mode}".encode()))
    pssm = make_pssm(rng, m, kind)
    checked = 0
    for lookups in (0, 3, 4, 5, 6, 8, 10, 12, 16):
        for thr in (None, float(np.nanmax(np.where(np.isfinite(pssm[:, :4]), pssm[:, :4], np.nan), axis=1).sum()) - 3.0):
            plan = lm.plan_filter(pssm, threshold=thr, n_positions=1e9, lookups=lookups, mode=mode)
            if plan is None:
                continue
            assert plan["pair"] == (mode == 2) and 0 < plan["smax"] <= 127
            n = 4000
            win = rng.integers(0, 4, size=(n, m)).astype(np.uint8)
            win[: n // 8][rng.random((n // 8, m)) < 0.15] = 4            # windows with N
            win[n // 8: n // 4] = np.argmax(pssm[:, :4], axis=1)[None, :]  # consensus ...
            flip = rng.random((n // 8, m)) < 0.1                           # ... with a few mismatches
            win[n // 8: n // 4][flip] = rng.integers(0, 4, size=int(flip.sum()))
            w64 = pssm.astype(np.float64)[np.arange(m)[None, :], win]      # [n, m]
            real = w64.sum(axis=1)
            s32 = np.zeros(n, dtype=np.float32)
            for j in range(m):                                             # the reference's sequential f32 sum
                s32 = (s32 + pssm[j, win[:, j]]).astype(np.float32)
            sq = sum_q(plan, win)
            assert sq.max() <= plan["smax"]
            ub = plan["c0"] + plan["delta"] * sq
            assert np.all(ub >= real - 1e-9), (lookups, thr, float((real - ub).max()))
            finite = np.isfinite(s32)
            assert np.all(np.abs(s32[finite].astype(np.float64) - real[finite]) <= plan["eps"])
            need = np.ceil((s32[finite].astype(np.float64) - plan["eps"] - plan["c0"]) / plan["delta"] - 1e-6)
            assert np.all(sq[finite] >= need), (lookups, thr)
            checked += 1
    assert checked >= 2 or (mode == 2 and m < 5)  # paired look-ups need at least 5 motif rows


def test_planner_prefers_the_filter_for_informative_motifs():
    pssm = O.pssm_from_counts(O.DNA, np.load(REPO / "tests/golden/jaspar2024.npz")["rest"], 0.1)
    plan = lm.plan_filter(pssm, n_positions=5e9)
    assert plan["ok"] and plan["est_pass_rate"] < 1e-3 and plan["est_cost"] < 40
    flat = np.zeros((8, 5), np.float32)  # no information: every cell ties, nothing can be filtered
    flat[:, 4] = -np.inf
    p2 = lm.plan_filter(flat, n_positions=5e9)
    assert p2 is None or not p2["ok"]


@pytest.mark.parametrize("m,kind", [(15, "acp"), (23, "acp"), (8, "random"), (12, "neginf"), (16, "finite_x")])
def test_protein_bound_dominates_score(m, kind):
    """Same property for the protein flavour: 8 residue classes, entry = best weight in the class."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"prot-{m}-{kind}".encode()))
    if kind == "acp":
        pats = ["SFKELGFDSLTAVELRNRLAAAT", "AFKELGFDSLAAIQLRNRLLADV", "PSRRLGFDSLTAVELRNQLAAST",
                "AFREIGFDSLTAVELRNRLGAAA", "SLMEEGLDSLAAVELGGTLQRDT", "GFFDLGMDSLMAVELRRRIEQGV"]
        pssm = O.pssm_from_sequences(O.PROTEIN, pats, 0.1)[:m]
    elif kind == "finite_x":
        pssm = rng.normal(0, 2, size=(m, 21)).astype(np.float32)
    else:
        counts = rng.integers(0, 6 if kind == "random" else 2, size=(m, 21)).astype(np.uint32)
        counts[:, 20] = 0
        counts[np.arange(m), rng.integers(0, 20, m)] += 40
        pssm = O.pssm_from_counts(O.PROTEIN, counts, 0.1 if kind == "random" else 0.0)
    checked = 0
    for lookups in (0, 4, 6, 8, 10, 12, 16):
        for thr in (None, 5.0):
            plan = lm.plan_filter_protein(pssm, threshold=thr, lookups=lookups)
            if plan is None:
                continue
            cls = plan["class_of"]
            assert cls.max() <= 7 and 0 < plan["smax"] <= 127
            n = 4000
            win = rng.integers(0, 20, size=(n, m)).astype(np.uint8)
            win[: n // 8][rng.random((n // 8, m)) < 0.1] = 20                 # X
            win[n // 8: n // 4] = np.argmax(pssm[:, :20], axis=1)[None, :]     # consensus with mismatches
            flip = rng.random((n // 8, m)) < 0.1
            win[n // 8: n // 4][flip] = rng.integers(0, 20, size=int(flip.sum()))
            real = pssm.astype(np.float64)[np.arange(m)[None, :], win].sum(axis=1)
            s32 = np.zeros(n, dtype=np.float32)
            for j in range(m):
                s32 = (s32 + pssm[j, win[:, j]]).astype(np.float32)
            sq = np.zeros(n, dtype=np.int64)
            for k in range(plan["r"]):
                tab = np.array([(plan["tab_lo"][k] >> (8 * b)) & 0xFF for b in range(4)] +
                               [(plan["tab_hi"][k] >> (8 * b)) & 0xFF for b in range(4)], dtype=np.int64)
                sq += tab[cls[win[:, plan["j0"] + k]]]
            assert np.all(plan["c0"] + plan["delta"] * sq >= real - 1e-9)
            finite = np.isfinite(s32)
            need = np.ceil((s32[finite].astype(np.float64) - plan["eps"] - plan["c0"]) / plan["delta"] - 1e-6)
            assert np.all(sq[finite] >= need)
            checked += 1
    assert checked >= 2
