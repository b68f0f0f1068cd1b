"""The filter + recheck ("pruned") fused scans must return exactly what the
score-every-cell path returns: same cell, bit-identical score, same hit list --
including adversarial inputs (low complexity, N runs, ties, -inf weights)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def lm():
    import lightmotif_b200 as lm
    return lm


@pytest.fixture()
def dna(lm):
    p = lm.Pipeline.cuda(lm.DNA)
    yield p
    p.close()


def random_pssm(rng, m, kind):
    if kind == "logodds":
        counts = rng.integers(0, 50, size=(m, 5)).astype(np.uint32)
        counts[:, 4] = 0
        counts[rng.integers(0, m, m // 2), rng.integers(0, 4, m // 2)] += 400  # informative rows
        return O.pssm_from_counts(O.DNA, counts, 0.1)
    if kind == "neginf":
        counts = rng.integers(0, 3, size=(m, 5)).astype(np.uint32)
        counts[:, 4] = 0
        counts[np.arange(m), rng.integers(0, 4, m)] += 30
        return O.pssm_from_counts(O.DNA, counts, 0.0)  # zero counts -> -inf cells
    if kind == "ints":  # many exact ties
        p = rng.integers(-3, 3, size=(m, 5)).astype(np.float32)
        p[:, 4] = -np.inf
        return p
    if kind == "finite_n":  # N column finite
        p = rng.normal(0, 2, size=(m, 5)).astype(np.float32)
        return p
    raise ValueError(kind)


def make_sequence(rng, n, kind):
    if kind == "uniform":
        return rng.integers(0, 4, n).astype(np.uint8)
    if kind == "with_n":
        s = rng.integers(0, 4, n).astype(np.uint8)
        for _ in range(20):
            a = rng.integers(0, n)
            s[a:a + rng.integers(1, 3000)] = 4
        return s
    if kind == "lowcomplexity":
        s = rng.integers(0, 4, n).astype(np.uint8)
        a = n // 3
        s[a:a + 200000] = 0          # poly-A
        s[2 * a:2 * a + 100000] = np.tile(np.array([1, 3], np.uint8), 50000)  # CGCG...
        return s
    raise ValueError(kind)


def compare(pli, seq, pssm, thresholds):
    """exact-everywhere (prune 0) vs the filter in both look-up flavours (single rows, paired rows)."""
    res = {}
    for key, mode, fmode in (("exact", 0, 0), ("single", 2, 1), ("paired", 2, 2)):
        pli.set_prune(mode)
        pli.set_filter_mode(fmode)
        res[key] = {"argmax": {r: seq.score_argmax(pssm, rule=r) for r in (0, 1, 2)},
                    "stats": pli.last_stats,
                    "thr": {t: seq.score_threshold(pssm, t) for t in thresholds}}
    pli.set_filter_mode(0)
    for key in ("single", "paired"):
        for r in (0, 1, 2):
            assert res["exact"]["argmax"][r] == res[key]["argmax"][r], (key, r, res["exact"]["argmax"][r], res[key]["argmax"][r])
        for t in thresholds:
            assert res["exact"]["thr"][t] == res[key]["thr"][t], (key, t)
    if len(pssm) >= 8:  # both flavours have a plan for motifs this long
        assert res["single"]["stats"]["pruned"] and not res["single"]["stats"]["paired"]
        assert res["paired"]["stats"]["pruned"] and res["paired"]["stats"]["paired"]
    return res


@pytest.mark.parametrize("m,pk", [(4, "logodds"), (6, "ints"), (8, "logodds"), (11, "neginf"), (15, "logodds"),
                                  (20, "logodds"), (20, "finite_n"), (21, "ints"), (33, "logodds")])
@pytest.mark.parametrize("sk", ["uniform", "with_n", "lowcomplexity"])
def test_pruned_equals_exact(dna, m, pk, sk):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{m}-{pk}-{sk}".encode()))
    n = 1_500_003
    pssm = random_pssm(rng, m, pk)
    sym = make_sequence(rng, n, sk)
    seq = dna.sequence_from_symbols(sym, m - 1)
    dna.set_boot_rows(4096)
    # thresholds: high (few hits), medium, and one so low that tiles overflow to the exact fallback
    dna.set_prune(0)
    ref = seq.score(pssm)
    finite = ref[np.isfinite(ref)]
    ts = [float(np.quantile(finite, q)) for q in (0.99999, 0.999, 0.5)]
    res = compare(dna, seq, pssm, ts)
    # and against the CPU oracle on the materialised scores
    for r in (0, 1, 2):
        rr, cc = O.argmax(ref, r)
        got = res["paired"]["argmax"][r]
        assert (got["row"], got["col"]) == (rr, cc)
        assert np.float32(got["score"]).view(np.uint32) == ref[rr, cc].view(np.uint32)
    ro, co = O.threshold(ref, ts[0])
    for key in ("single", "paired"):
        assert [(h["row"], h["col"]) for h in res[key]["thr"][ts[0]]] == list(zip(ro.tolist(), co.tolist()))


def test_pruned_actually_prunes_and_matches_on_rest(dna, jaspar):
    pssm = O.pssm_from_counts(O.DNA, jaspar["rest"], 0.1)
    n = 64_000_000
    seq = dna.sequence_synthetic(n, len(pssm) - 1)
    dna.set_prune(0)
    exact = seq.score_argmax(pssm, with_columns=True)
    assert not dna.last_stats["pruned"]
    dna.set_prune(1)  # auto: the cost model must pick the filter for this motif
    pr = seq.score_argmax(pssm, with_columns=True)
    st = dna.last_stats
    assert st["pruned"] and st["overflow_tiles"] == 0 and st["rechecked"] < n // 100
    assert {k: pr[k] for k in ("row", "col", "position", "score")} == {k: exact[k] for k in ("row", "col", "position", "score")}
    assert pr["first_cell"] == exact["first_cell"]
    # per-column winners: exact wherever they attain the maximum, lower bounds elsewhere
    for (r0, s0, v0), (r1, s1, v1) in zip(exact["columns"], pr["columns"]):
        assert v1 and s1 <= s0
        if s0 == exact["score"]:
            assert (r1, s1) == (r0, s0)
    for t in (12.0, 18.0):
        dna.set_prune(0)
        a = seq.score_threshold(pssm, t)
        dna.set_prune(1)
        b = seq.score_threshold(pssm, t)
        assert dna.last_stats["pruned"] and a == b and len(a) > 0
    # sub-range scans
    dna.set_boot_rows(8192)
    lo, hi = 123456, 1_900_000
    dna.set_prune(0)
    a = seq.score_argmax(pssm, rows=(lo, hi))
    dna.set_prune(2)
    b = seq.score_argmax(pssm, rows=(lo, hi))
    assert a == b and dna.last_stats["pruned"]


def test_all_equal_sequence_ties(dna):
    """Every cell ties: the filter passes everything, tiles fall back to exact
    scoring, and the tie rules still pick the reference's cell."""
    n = 400_000
    sym = np.zeros(n, np.uint8)
    rng = np.random.default_rng(5)
    pssm = random_pssm(rng, 12, "logodds")
    seq = dna.sequence_from_symbols(sym, 11)
    dna.set_boot_rows(1024)
    dna.set_prune(0)
    ref = seq.score(pssm)
    for fmode in (1, 2):
        dna.set_filter_mode(fmode)
        for r in (0, 1, 2):
            dna.set_prune(2)
            got = seq.score_argmax(pssm, rule=r)
            rr, cc = O.argmax(ref, r)
            assert (got["row"], got["col"]) == (rr, cc), (fmode, r)


# ------------------------------------------------------------------ protein
@pytest.fixture()
def prot(lm):
    p = lm.Pipeline.cuda(lm.PROTEIN)
    yield p
    p.close()


def protein_pssm(rng, m, kind):
    if kind == "acp":  # the reference's bench motif (lightmotif/benches/score.rs:179-186)
        pats = ["SFKELGFDSLTAVELRNRLAAAT", "AFKELGFDSLAAIQLRNRLLADV", "PSRRLGFDSLTAVELRNQLAAST",
                "AFREIGFDSLTAVELRNRLGAAA", "SLMEEGLDSLAAVELGGTLQRDT", "GFFDLGMDSLMAVELRRRIEQGV"]
        return O.pssm_from_sequences(O.PROTEIN, pats, 0.1)[:m]
    if kind == "random":
        counts = rng.integers(0, 6, size=(m, 21)).astype(np.uint32)
        counts[:, 20] = 0
        counts[np.arange(m), rng.integers(0, 20, m)] += 60
        return O.pssm_from_counts(O.PROTEIN, counts, 0.1)
    if kind == "neginf":
        counts = rng.integers(0, 2, size=(m, 21)).astype(np.uint32)
        counts[:, 20] = 0
        counts[np.arange(m), rng.integers(0, 20, m)] += 25
        return O.pssm_from_counts(O.PROTEIN, counts, 0.0)
    if kind == "finite_x":
        return rng.normal(0, 2, size=(m, 21)).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("m,kind", [(15, "acp"), (23, "acp"), (8, "random"), (12, "neginf"), (16, "finite_x"), (30, "random")])
def test_protein_filter_equals_exact(prot, m, kind):
    """Protein fused scans through the class-coded filter (8 residue classes, exact recheck from the
    residues) == scoring every cell: same cell, bit-identical score, same hit list."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"prot-{m}-{kind}".encode()))
    n = 1_200_007
    pssm = protein_pssm(rng, m, kind)
    sym = O.synth_protein(n)
    sym[::997] = 20                      # X
    sym[5000:5600] = 3                   # low complexity
    # plant near-consensus windows so that thresholds have hits
    cons = np.argmax(pssm[:, :20], axis=1).astype(np.uint8)
    for at in (1234, 400_000, n - m - 3):
        sym[at:at + m] = cons
        sym[at + 2] = (cons[2] + 1) % 20
    seq = prot.sequence_from_symbols(sym, m - 1)
    prot.set_boot_rows(2048)
    prot.set_prune(0)
    ref = seq.score(pssm)
    want = O.score_f32(pssm, O.stripe(O.PROTEIN, sym, 32, m - 1), n)
    assert np.array_equal(ref.view(np.uint32), want.view(np.uint32))
    finite = ref[np.isfinite(ref)]
    ts = [float(np.quantile(finite, q)) for q in (0.999999, 0.9999, 0.5)]
    exact = {"argmax": {r: seq.score_argmax(pssm, rule=r) for r in (0, 1, 2)},
             "thr": {t: seq.score_threshold(pssm, t) for t in ts}}
    assert not prot.last_stats["pruned"]
    prot.set_prune(2)
    for r in (0, 1, 2):
        got = seq.score_argmax(pssm, rule=r)
        assert prot.last_stats["pruned"]
        assert got == exact["argmax"][r], (r, got, exact["argmax"][r])
        rr, cc = O.argmax(want, r)
        assert (got["row"], got["col"]) == (rr, cc)
    for t in ts:
        assert seq.score_threshold(pssm, t) == exact["thr"][t], t
    # a second sequence at (possibly) the same address must not see the first one's class copy
    seq.free()
    sym2 = O.synth_protein(n, offset=12345)
    seq2 = prot.sequence_from_symbols(sym2, m - 1)
    prot.set_prune(0)
    e2 = seq2.score_argmax(pssm)
    prot.set_prune(2)
    assert seq2.score_argmax(pssm) == e2


# ------------------------------------------------------------------ edge cases
def test_long_motifs_and_limits(dna):
    """Motifs at and beyond the filter's width limit (256 rows) and sub-range scans: the filter is used
    where a plan exists, the exact kernel otherwise; results identical either way."""
    rng = np.random.default_rng(11)
    n = 2_000_003
    sym = rng.integers(0, 4, n).astype(np.uint8)
    sym[70_000:70_050] = 4
    dna.set_boot_rows(2048)
    for m in (64, 200, 256, 300):
        counts = rng.integers(0, 30, size=(m, 5)).astype(np.uint32)
        counts[:, 4] = 0
        counts[np.arange(m), rng.integers(0, 4, m)] += 60
        pssm = O.pssm_from_counts(O.DNA, counts, 0.1)
        sym[500_000:500_000 + m] = np.argmax(pssm[:, :4], axis=1)  # one perfect site
        seq = dna.sequence_from_symbols(sym, m - 1)
        dna.set_prune(0)
        exact = seq.score_argmax(pssm)
        sub_exact = seq.score_argmax(pssm, rows=(1000, 40_000))
        t = exact["score"] - 40.0
        hits_exact = seq.score_threshold(pssm, t)
        dna.set_prune(2)
        assert seq.score_argmax(pssm) == exact
        assert dna.last_stats["pruned"] == (1 if m <= 256 else 0)
        assert seq.score_argmax(pssm, rows=(1000, 40_000)) == sub_exact
        assert seq.score_threshold(pssm, t) == hits_exact and len(hits_exact) >= 1
        dna.set_prune(0)
        sub_hits = seq.score_threshold(pssm, t, rows=(15_000, 16_000))
        dna.set_prune(2)
        assert seq.score_threshold(pssm, t, rows=(15_000, 16_000)) == sub_hits
        seq.free()
    dna.set_prune(1)


def test_all_default_symbols_and_minus_infinity(dna):
    """A sequence of N only: every score is -inf; the fused calls still return the reference's cell."""
    n = 700_001
    sym = np.full(n, 4, np.uint8)
    rng = np.random.default_rng(2)
    pssm = random_pssm(rng, 12, "logodds")
    seq = dna.sequence_from_symbols(sym, 11)
    dna.set_boot_rows(1024)
    dna.set_prune(0)
    ref = seq.score(pssm)
    assert np.all(np.isneginf(ref))
    for r in (0, 1, 2):
        dna.set_prune(0)
        exact = seq.score_argmax(pssm, rule=r)
        dna.set_prune(2)
        got = seq.score_argmax(pssm, rule=r)
        rr, cc = O.argmax(ref, r)
        assert (got["row"], got["col"]) == (exact["row"], exact["col"]) == (rr, cc), r
    dna.set_prune(2)
    assert seq.score_threshold(pssm, 0.0) == []


def test_dense_candidates_and_replanning_stay_exact(lm, dna, jaspar, ecoli_codes):
    """Sequences without a strong site for the motif: the filter planned for the maximum of a
    uniform-random sequence floods with candidates.  The scan that observes it switches the next
    scans of the same matrix to the dense-candidate kernel instance and to a plan for the maximum
    actually found -- every scan must still return exactly what scoring every cell returns, for
    argmax (all rules) and for thresholds with thousands of hits."""
    from lightmotif_b200 import pwm
    pssm = pwm.pssm_from_counts(pwm.DNA, jaspar["rest"], 0.1)
    rng = np.random.default_rng(5)
    inputs = {"ecoli x4": np.tile(ecoli_codes, 4),
              "two-letter": rng.integers(0, 2, 12_000_000).astype(np.uint8)}  # A / C only: REST never scores high
    for name, codes in inputs.items():
        seq = dna.sequence_from_symbols(codes, len(pssm) - 1)
        # a deliberately weak filter (4 single-row look-ups) on the second input: certain flooding
        dna.set_filter_rows(4 if name == "two-letter" else 0)
        dna.set_filter_mode(1 if name == "two-letter" else 0)
        dna.set_prune(0)
        want = {rule: seq.score_argmax(pssm, rule=rule) for rule in (0, 1, 2)}
        t = want[1]["score"] - 8.0
        want_hits = seq.score_threshold(pssm, t)
        assert len(want_hits) > 200
        dna.set_boot_rows(4096)
        dna.set_prune(2)
        seen = []
        for rep in range(4):
            for rule in (0, 1, 2):
                assert seq.score_argmax(pssm, rule=rule) == want[rule], (name, rep, rule)
            st = dna.last_stats
            seen.append((st["filter_rows"], st["paired"], st["rechecked"], st["overflow_tiles"]))
        work = [c + 128 * o for _, _, c, o in seen]
        assert work[-1] <= work[0], seen  # re-planning never makes the filter weaker
        if name == "two-letter":
            assert work[0] > 8 * codes.size / 3968, seen  # it did flood
        for rep in range(3):  # first the sparse, then (if dense) the enumerating instance
            assert seq.score_threshold(pssm, t) == want_hits, (name, rep)
        # another matrix in between: observations are per matrix
        other = pwm.pssm_from_counts(pwm.DNA, jaspar["motifs"][5], 0.1)
        dna.set_prune(0)
        w2 = seq.score_argmax(other)
        dna.set_prune(2)
        assert seq.score_argmax(other) == w2 and seq.score_argmax(pssm) == want[1]
        seq.free()
    dna.set_filter_rows(0)
    dna.set_filter_mode(0)
