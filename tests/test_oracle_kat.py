"""Pin the CPU oracle against every golden vector the reference's own tests
hold for the hot path.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

LETTERS = "ACTGN"


def _dna_pssm(kat):
    d = kat["dna"]
    return O.pssm_from_sequences(O.DNA, d["patterns"], d["pseudocount"])


@pytest.mark.parametrize("c", [1, 2, 4, 16, 32])
def test_score_expected(kat, c):
    """lightmotif/tests/dna.rs:65-91 (test_score) and :40-63 (test_score_rows)."""
    d = kat["dna"]
    pssm = _dna_pssm(kat)
    enc = O.encode(O.DNA, d["sequence"])
    st = O.stripe(O.DNA, enc, c, wrap=pssm.shape[0] - 1)
    sc = O.score_f32(pssm, st, len(d["sequence"]))
    u = O.unstripe(sc, len(d["sequence"]) + 1 - pssm.shape[0])
    exp = np.array(d["expected"], dtype=np.float32)
    assert u.shape == exp.shape
    assert np.abs(u - exp).max() < d["tolerance"]
    for p in d["exact_positions"]:
        assert u[p] == exp[p]
    assert u[0] == np.float32(d["readme_score0"])
    # score_rows_into(0..2) then (1..2): tests/dna.rs:56-62
    s02 = O.score_f32(pssm, st, len(d["sequence"]), 0, min(2, sc.shape[0]))
    assert s02[0, 0] == exp[0]
    if c == 32:
        assert s02.shape[0] == 2 and s02[1, 0] == exp[1]
        s12 = O.score_f32(pssm, st, len(d["sequence"]), 1, 2)
        assert s12.shape[0] == 1 and s12[0, 0] == exp[1]


@pytest.mark.parametrize("c", [1, 16, 32])
@pytest.mark.parametrize("rule", [O.RULE_GENERIC, O.RULE_AVX2, O.RULE_SSE2])
def test_argmax_threshold(kat, c, rule):
    """tests/dna.rs:122-173"""
    d = kat["dna"]
    pssm = _dna_pssm(kat)
    st = O.stripe(O.DNA, O.encode(O.DNA, d["sequence"]), c, wrap=14)
    sc = O.score_f32(pssm, st, len(d["sequence"]))
    r, col = O.argmax(sc, rule)
    assert O.offset(r, col, sc.shape[0]) == d["argmax"]
    for t, hits in d["threshold"].items():
        ro, co = O.threshold(sc, float(t))
        assert sorted((co.astype(np.int64) * sc.shape[0] + ro.astype(np.int64)).tolist()) == hits
    ro, co = O.threshold(sc, 10.0)  # README.md:90
    assert ro.size == 0


def test_score_discrete_upper_bound(kat):
    """tests/dna.rs:93-120: unscale(u8 score) >= f32 score at every position."""
    d = kat["dna"]
    pssm = _dna_pssm(kat)
    dm = O.Discrete(O.DNA, pssm)
    st = O.stripe(O.DNA, O.encode(O.DNA, d["sequence"]), 32, wrap=14)
    ds = O.unstripe(O.score_u8(dm.data, st, len(d["sequence"])), 50)
    exp = np.array(d["expected"], dtype=np.float32)
    for i in range(50):
        assert dm.unscale(ds[i]) >= exp[i]


def test_score_position(kat):
    """tests/dna.rs:175-199"""
    d = kat["dna"]
    pssm = _dna_pssm(kat)
    st = O.stripe(O.DNA, O.encode(O.DNA, d["sequence"]), 32, wrap=14)
    exp = np.array(d["expected"], dtype=np.float32)
    for i in range(50):
        assert abs(O.score_position(pssm, st, 2, i) - exp[i]) < 1e-5


def test_encode(kat):
    """tests/encode.rs:10-26"""
    e = kat["encode"]
    enc = O.encode(O.DNA, e["sequence"])
    assert enc.tolist() == [e["codes"][ch] for ch in e["expected_letters"]]
    with pytest.raises(O.InvalidSymbol) as ei:
        O.encode(O.DNA, e["unknowns"])
    assert ei.value.ch == e["unknown_char"]
    lossy = O.encode_lossy(O.DNA, e["unknowns"])
    assert "".join(LETTERS[x] for x in lossy) == e["unknowns"].replace(".", "N")
    # lower case is invalid (exact upper-case match, abc.rs:161-171)
    with pytest.raises(O.InvalidSymbol):
        O.encode(O.DNA, "acgt")
    assert O.encode(O.PROTEIN, "ACDEFGHIKLMNPQRSTVWYX").tolist() == list(range(21))


@pytest.mark.parametrize("which", ["s1", "s2"])
@pytest.mark.parametrize("c", [16, 32])
def test_stripe(kat, which, c):
    """tests/stripe.rs:17-45"""
    s = kat["stripe"]
    seq = s["s2"] if which == "s2" else s["s1_unit"] * s["s1_repeats"] + s["s1_tail"]
    enc = O.encode(O.DNA, seq)
    m = O.stripe(O.DNA, enc, c)
    rows = m.shape[0]
    assert rows == -(-len(seq) // c)
    for i, ch in enumerate("ATGT"[:rows]):
        assert LETTERS[m[i, 0]] == ch
    idx = np.arange(len(seq))
    assert np.array_equal(m[idx % rows, idx // rows], enc)
    idx = np.arange(len(seq), rows * c)
    assert np.all(m[idx % rows, idx // rows] == 4)


def test_stripe_literals(kat):
    """seq.rs:497-541 (empty, U4 / U2 literal matrices, configure_wrap rows)."""
    assert O.stripe(O.DNA, np.zeros(0, np.uint8), 2).shape[0] == 0
    u4 = kat["stripe"]["u4"]
    enc = O.encode(O.DNA, u4["sequence"])
    m = O.stripe(O.DNA, enc, 4)
    assert ["".join(LETTERS[x] for x in r) for r in m] == u4["matrix"]
    m2 = O.stripe(O.DNA, enc, 2)
    assert ["".join(LETTERS[x] for x in r) for r in m2] == ["AC", "TA", "GN"]
    w = O.stripe(O.DNA, enc, 4, wrap=2)
    assert ["".join(LETTERS[x] for x in r) for r in w] == u4["matrix"] + u4["wrap2"]


def test_wrap_closed_form():
    """configure_wrap == 'cell (r, c) holds symbol c*rows + r, N past the end',
    including the cascading case wrap > rows."""
    rng = np.random.default_rng(1)
    for n, c, wrap in [(5, 4, 2), (100, 32, 14), (40, 32, 9), (1031, 32, 40), (64, 16, 7)]:
        sym = rng.integers(0, 4, n).astype(np.uint8)
        st = O.stripe(O.DNA, sym, c, wrap=wrap)
        rows = st.shape[0] - wrap
        ext = np.concatenate([sym, np.full(c * (rows + wrap) + wrap, 4, np.uint8)])
        r = np.arange(rows + wrap)[:, None]
        col = np.arange(c)[None, :]
        assert np.array_equal(st, ext[col * rows + r])


def test_ecoli_tenth_argmax(kat, ecoli_codes):
    """tests/argmax.rs:23-53 pins the LAST maximum (401649); the strict-'>'
    scan of lightmotif-bench/dna.rs:247,282 pins the first (391677)."""
    k = kat["ecoli_tenth"]
    pssm = O.pssm_from_sequences(O.DNA, k["patterns"], 0.1)
    enc = ecoli_codes[: k["n"]]
    for c in (16, 32):
        st = O.stripe(O.DNA, enc, c, wrap=14)
        sc = O.score_f32(pssm, st, k["n"])
        u = O.unstripe(sc, k["n"] + 1 - 15)
        best = np.flatnonzero(u == u.max())
        assert best[0] == k["argmax_strict_first"] and best[-1] == k["argmax_pipeline"]
        for rule in (O.RULE_GENERIC, O.RULE_AVX2, O.RULE_SSE2):
            if c == 16 and rule == O.RULE_AVX2:
                continue
            r, col = O.argmax(sc, rule)
            assert O.offset(r, col, sc.shape[0]) == k["argmax_pipeline"]
            assert sc[r, col] == u.max()
    st = O.stripe(O.DNA, enc, 32, wrap=14)
    pos, score = O.scanner_max(O.DNA, pssm, st, 14, k["n"])
    assert pos == k["argmax_pipeline"] and score == u.max()


def test_avx2_restatement_bit_exact(kat, ecoli_codes):
    """The AVX2 restatement used for CPU timing equals the generic oracle."""
    k = kat["ecoli_tenth"]
    pssm = O.pssm_from_sequences(O.DNA, k["patterns"], 0.1)
    st = O.stripe(O.DNA, ecoli_codes[: k["n"]], 32, wrap=14)
    sc = O.score_f32(pssm, st, k["n"])
    a = O.avx2_score_f32_dna(pssm, st, 0, sc.shape[0])
    assert np.array_equal(a.view(np.uint32), sc.view(np.uint32))
    assert O.avx2_argmax_f32(a) == O.argmax(sc, O.RULE_AVX2)
    best, r, c = O.avx2_score_argmax_mt(pssm, st, sc.shape[0], 3)
    assert (r, c) == O.argmax(sc, O.RULE_AVX2) and best == sc[r, c]


@pytest.mark.parametrize("pc", [0.1, 0.0])
def test_scanner_consistency(kat, ecoli_codes, ecoli_kat, pc):
    """tests/scan.rs:9-83: Scanner hits == brute-force filter, exact f32."""
    s = kat["scan"]
    counts = np.zeros((15, 5), np.uint32)
    for p in s["patterns"]:
        for i, x in enumerate(O.encode_lossy(O.DNA, p)):
            counts[i, x] += 1
    pssm = O.pssm_from_counts(O.DNA, counts, pc)
    n = ecoli_codes.size
    st = O.stripe(O.DNA, ecoli_codes, 32, wrap=s["wrap"])
    u = O.unstripe(O.score_f32(pssm, st, n, wrap=s["wrap"]), n + 1 - 15)
    want = np.flatnonzero(u >= s["threshold"])
    if pc == 0.1:
        assert want.size == ecoli_kat["scan_hits_pc0.1_thr5"]
    for bs in (32, 256):
        pos, sc = O.scanner_collect(O.DNA, pssm, st, s["wrap"], n, s["threshold"], bs)
        assert np.array_equal(pos.astype(np.int64), want)
        assert np.array_equal(sc.view(np.uint32), u[want].view(np.uint32))


def test_full_genome_anchors(kat, ecoli_codes, ecoli_kat):
    pssm = _dna_pssm(kat)
    n = ecoli_codes.size
    st = O.stripe(O.DNA, ecoli_codes, 32, wrap=14)
    u = O.unstripe(O.score_f32(pssm, st, n), n + 1 - 15)
    assert float(u.max()) == ecoli_kat["full_genome_max"]
    assert int(np.flatnonzero(u == u.max())[-1]) == ecoli_kat["full_genome_argmax_last"]
    assert int((u >= 10.0).sum()) == ecoli_kat["count_ge_10"]


def test_synth_generators():
    a = O.synth_dna(1000, 0)
    b = O.synth_dna(500, 500)
    assert np.array_equal(a[500:], b) and a.max() == 3
    counts = np.bincount(O.synth_dna(200000), minlength=4)
    assert counts.min() > 49000
    p = O.synth_protein(1000, 7)
    assert p.max() <= 19 and np.array_equal(p[10:20], O.synth_protein(10, 17))
