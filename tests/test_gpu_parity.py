"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs -- bit-exact for symbols, u8 scores, indices AND f32
scores (the kernels keep the reference's sequential summation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)

LETTERS = np.frombuffer(b"ACTGN", dtype=np.uint8)
AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)


@pytest.fixture(scope="module")
def lm():
    import lightmotif_b200 as lm
    return lm


@pytest.fixture(scope="module")
def dna(lm):
    p = lm.Pipeline.cuda(lm.DNA)
    yield p
    p.close()


@pytest.fixture(scope="module")
def prot(lm):
    p = lm.Pipeline.cuda(lm.PROTEIN)
    yield p
    p.close()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def random_pssm(rng, m, k, neg_inf_cells=0):
    p = rng.normal(0, 2, size=(m, k)).astype(np.float32)
    p[:, k - 1] = -np.inf
    for _ in range(neg_inf_cells):
        p[rng.integers(0, m), rng.integers(0, k - 1)] = -np.inf
    return p


# ------------------------------------------------------------------ encode

@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 64, 1000, 65537])
def test_encode_matches_oracle(dna, prot, n):
    rng = np.random.default_rng(n)
    a = LETTERS[rng.integers(0, 5, n)]
    assert np.array_equal(dna.encode(a), O.encode(O.DNA, a))
    b = AA[rng.integers(0, 21, n)]
    assert np.array_equal(prot.encode(b), O.encode(O.PROTEIN, b))


def test_encode_kat_and_errors(lm, dna, kat):
    e = kat["encode"]
    assert dna.encode(e["sequence"]).tolist() == [e["codes"][c] for c in e["expected_letters"]]
    with pytest.raises(lm.InvalidSymbol) as ei:
        dna.encode(e["unknowns"])
    assert ei.value.ch == "." and ei.value.index == e["unknowns"].index(".")
    assert np.array_equal(dna.encode_lossy(e["unknowns"]), O.encode_lossy(O.DNA, e["unknowns"]))
    # first invalid byte wins, wherever it is; lower case is invalid
    rng = np.random.default_rng(7)
    a = LETTERS[rng.integers(0, 4, 100000)].copy()
    a[[77777, 99999, 12345]] = ord("x")
    with pytest.raises(lm.InvalidSymbol) as ei:
        dna.encode(a)
    assert ei.value.index == 12345 and ei.value.ch == "x"
    with pytest.raises(lm.InvalidSymbol):
        dna.encode("acgt")
    # unaligned host views
    buf = LETTERS[rng.integers(0, 5, 4099)]
    assert np.array_equal(dna.encode(buf[3:]), O.encode(O.DNA, buf[3:]))


# ------------------------------------------------------------ stripe / wrap

@pytest.mark.parametrize("n,wrap", [(0, 0), (0, 3), (1, 0), (5, 2), (31, 4), (32, 1), (33, 14), (64, 14),
                                    (40, 9), (1031, 40), (100003, 29), (4096, 0)])
def test_stripe_wrap_matches_oracle(dna, n, wrap):
    rng = np.random.default_rng(n + wrap)
    sym = rng.integers(0, 5, n).astype(np.uint8)
    want = O.stripe(O.DNA, sym, 32, wrap)
    got = dna.stripe(sym, wrap)
    assert got.shape == want.shape and np.array_equal(got, want)
    seq = dna.sequence_from_symbols(sym, wrap)
    assert (seq.rows, seq.wrap, seq.length) == (want.shape[0] - wrap, wrap, n)
    assert np.array_equal(seq.download(), want)
    ascii_ = LETTERS[sym]
    seq2 = dna.sequence_from_ascii(ascii_, wrap)
    assert np.array_equal(seq2.download(), want)
    seq.configure_wrap(wrap + 7)  # grow on the device
    assert np.array_equal(seq.download(), O.stripe(O.DNA, sym, 32, wrap + 7))
    up = dna.sequence_from_striped(want, n, wrap)
    up.configure_wrap(wrap + 5)
    assert np.array_equal(up.download(), O.stripe(O.DNA, sym, 32, wrap + 5))


def test_stripe_kat(dna, kat):
    s = kat["stripe"]
    for seq in (s["s2"], s["s1_unit"] * s["s1_repeats"] + s["s1_tail"]):
        enc = dna.encode(seq)
        m = dna.stripe(enc)
        rows = m.shape[0]
        idx = np.arange(len(seq))
        assert np.array_equal(m[idx % rows, idx // rows], enc)
        idx = np.arange(len(seq), rows * 32)
        assert np.all(m[idx % rows, idx // rows] == 4)


def test_sequence_from_ascii_errors(lm, dna):
    with pytest.raises(lm.InvalidSymbol) as ei:
        dna.sequence_from_ascii("ACGT" * 100 + "Z" + "ACGT" * 50, 3)
    assert ei.value.index == 400
    seq = dna.sequence_from_ascii("ACGT" * 100 + "Z" + "ACGT" * 50, 3, lossy=True)
    want = O.stripe(O.DNA, O.encode_lossy(O.DNA, "ACGT" * 100 + "Z" + "ACGT" * 50), 32, 3)
    assert np.array_equal(seq.download(), want)


# ------------------------------------------------------------------- score

def test_score_kat(dna, kat):
    """lightmotif/tests/dna.rs: test_score, test_score_rows, test_argmax, test_threshold."""
    d = kat["dna"]
    pssm = O.pssm_from_sequences(O.DNA, d["patterns"], d["pseudocount"])
    enc = dna.encode(d["sequence"])
    striped = dna.stripe(enc, wrap=len(pssm) - 1)
    n = len(d["sequence"])
    sc = dna.score_rows(pssm, striped, n, len(pssm) - 1)
    want = O.score_f32(pssm, O.stripe(O.DNA, enc, 32, 14), n)
    assert np.array_equal(bits(sc), bits(want))
    u = O.unstripe(sc, n + 1 - len(pssm))
    exp = np.array(d["expected"], dtype=np.float32)
    assert np.abs(u - exp).max() < d["tolerance"]
    assert u[0] == exp[0] and u[1] == exp[1]
    s02 = dna.score_rows(pssm, striped, n, 14, rows=(0, 2))
    assert s02.shape[0] == 2 and s02[0, 0] == exp[0] and s02[1, 0] == exp[1]
    s12 = dna.score_rows(pssm, striped, n, 14, rows=(1, 2))
    assert s12.shape[0] == 1 and s12[0, 0] == exp[1]
    assert dna.score_rows(pssm, striped, n, 14, rows=(1, 1)).shape[0] == 0  # pli/mod.rs:603-623
    for rule in (0, 1, 2):
        mc = dna.argmax(sc, rule)
        assert mc.col * sc.shape[0] + mc.row == d["argmax"]
    for t, hits in d["threshold"].items():
        got = dna.threshold(sc, float(t))
        assert sorted(h.col * sc.shape[0] + h.row for h in got) == hits
        ro, co = O.threshold(sc, float(t))
        assert [(h.row, h.col) for h in got] == list(zip(ro.tolist(), co.tolist()))
    assert dna.threshold(sc, 10.0) == []
    seq = dna.sequence_from_symbols(enc, 14)
    best = seq.score_argmax(pssm)
    assert best["position"] == d["argmax"] and np.float32(best["score"]) == u[18]
    hits = seq.score_threshold(pssm, -15.0)
    assert sorted(h["position"] for h in hits) == d["threshold"]["-15.0"]


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 7, 8, 15, 16, 20, 21, 33, 64])
@pytest.mark.parametrize("n", [50, 4099])
def test_score_random_dna(dna, m, n):
    rng = np.random.default_rng(1000 * m + n)
    pssm = random_pssm(rng, m, 5, neg_inf_cells=m // 8)
    sym = rng.choice(5, size=n, p=[0.24, 0.25, 0.25, 0.24, 0.02]).astype(np.uint8)
    st = O.stripe(O.DNA, sym, 32, m - 1)
    want = O.score_f32(pssm, st, n)
    seq = dna.sequence_from_symbols(sym, m - 1)
    if n < m:
        assert seq.score(pssm).shape[0] == 0
        return
    got = seq.score(pssm)
    assert np.array_equal(bits(got), bits(want))
    rows = st.shape[0] - (m - 1)
    lo, hi = rows // 3, max(rows // 3 + 1, 2 * rows // 3)
    assert np.array_equal(bits(seq.score(pssm, rows=(lo, hi))), bits(want[lo:hi]))
    assert np.array_equal(bits(dna.score_rows(pssm, st, n, m - 1)), bits(want))
    for rule in (0, 1, 2):
        r, c = O.argmax(want, rule)
        best = seq.score_argmax(pssm, rule=rule)
        assert (best["row"], best["col"]) == (r, c), rule
        assert bits(np.float32(best["score"])) == bits(want[r, c])
        mc = dna.argmax(want, rule)
        assert (mc.row, mc.col) == (r, c)
    finite = want[np.isfinite(want)]
    t = float(np.quantile(finite, 0.9)) if finite.size else 0.0
    ro, co = O.threshold(want, t)
    hits = seq.score_threshold(pssm, t)
    assert [(h["row"], h["col"]) for h in hits] == list(zip(ro.tolist(), co.tolist()))
    assert np.array_equal(bits(np.array([h["score"] for h in hits], np.float32)), bits(want[ro, co]))
    assert dna.max(want) == O.max_(want)


def test_wrap_and_length_errors(lm, dna):
    rng = np.random.default_rng(3)
    sym = rng.integers(0, 4, 300).astype(np.uint8)
    pssm = random_pssm(rng, 10, 5)
    seq = dna.sequence_from_symbols(sym, 5)
    with pytest.raises(lm.LightmotifError) as ei:
        seq.score(pssm)
    assert ei.value.status == lm.capi.ERR_WRAP
    short = dna.sequence_from_symbols(sym[:6], 9)
    assert short.score(pssm).shape[0] == 0 and short.score_argmax(pssm) is None
    assert short.score_threshold(pssm, -100.0) == []
    assert dna.argmax(np.zeros((0, 32), np.float32)) is None


def test_score_ecoli(dna, kat, ecoli_codes, ecoli_kat):
    """tests/argmax.rs (first tenth: tie 391677 / 401649) and the full genome."""
    k = kat["ecoli_tenth"]
    pssm = O.pssm_from_sequences(O.DNA, k["patterns"], 0.1)
    enc = ecoli_codes[: k["n"]]
    seq = dna.sequence_from_symbols(enc, 14)
    want = O.score_f32(pssm, O.stripe(O.DNA, enc, 32, 14), k["n"])
    assert np.array_equal(bits(seq.score(pssm)), bits(want))
    for rule in (0, 1, 2):
        best = seq.score_argmax(pssm, rule=rule)
        assert best["position"] == k["argmax_pipeline"]
        assert np.float32(best["score"]) == want.max()
    n = ecoli_codes.size
    full = dna.sequence_from_symbols(ecoli_codes, 14)
    wantf = O.score_f32(pssm, O.stripe(O.DNA, ecoli_codes, 32, 14), n)
    assert np.array_equal(bits(full.score(pssm)), bits(wantf))
    best = full.score_argmax(pssm)
    assert best["position"] == ecoli_kat["full_genome_argmax_last"]
    assert best["score"] == ecoli_kat["full_genome_max"]
    hits = full.score_threshold(pssm, 10.0)
    assert len(hits) == ecoli_kat["count_ge_10"]
    ro, co = O.threshold(wantf, 10.0)
    assert [(h["row"], h["col"]) for h in hits] == list(zip(ro.tolist(), co.tolist()))


@pytest.mark.parametrize("pc", [0.1, 0.0])
def test_scanner_consistency(lm, dna, kat, ecoli_codes, pc):
    """tests/scan.rs:9-83 on the device: Scanner hits over the full E. coli genome (3 patterns,
    threshold 5.0, pseudocount 0.1 and 0.0 = -inf cells) == the oracle's Scanner == brute force,
    exact f32 scores; scan.rs:235 tie rule for max()."""
    s = kat["scan"]
    pssm = O.pssm_from_sequences(O.DNA, s["patterns"], pc)
    n = ecoli_codes.size
    st = O.stripe(O.DNA, ecoli_codes, 32, s["wrap"])
    wpos, wsc = O.scanner_collect(O.DNA, pssm, st, s["wrap"], n, s["threshold"], 256)
    order = np.argsort(wpos)
    seq = dna.sequence_from_symbols(ecoli_codes, s["wrap"])
    for prune in (1, 0):
        dna.set_prune(prune)
        hits = lm.Scanner(pssm, seq).threshold(s["threshold"]).block_size(64).collect()
        assert [h["position"] for h in hits] == wpos[order].tolist()
        assert np.array_equal(bits(np.array([h["score"] for h in hits], np.float32)), bits(wsc[order]))
    dna.set_prune(1)
    best = lm.Scanner(pssm, seq).threshold(s["threshold"]).max()
    full = O.score_f32(pssm, st, n, wrap=s["wrap"])
    flat = O.unstripe(full, n - len(pssm) + 1)
    top = flat.max()
    assert best["score"] == float(top) and best["position"] == int(np.flatnonzero(flat == top)[-1])
    assert lm.Scanner(pssm, seq).threshold(float(top) + 1.0).max() is None


def test_scanner_max_ecoli_tenth(lm, dna, kat, ecoli_codes):
    """tests/argmax.rs:76-107: Scanner::max on the first tenth of E. coli == last maximum (401649)."""
    k = kat["ecoli_tenth"]
    pssm = O.pssm_from_sequences(O.DNA, k["patterns"], 0.1)
    seq = dna.sequence_from_symbols(ecoli_codes[: k["n"]], 14)
    best = lm.Scanner(pssm, seq).max()
    assert best["position"] == k["argmax_pipeline"]


def test_score_u8(dna, kat, ecoli_codes):
    d = kat["dna"]
    pssm = O.pssm_from_sequences(O.DNA, d["patterns"], d["pseudocount"])
    dm = O.Discrete(O.DNA, pssm)
    enc = ecoli_codes[:200003]
    st = O.stripe(O.DNA, enc, 32, 14)
    want = O.score_u8(dm.data, st, enc.size)
    seq = dna.sequence_from_symbols(enc, 14)
    got = seq.score(dm.data)
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    assert np.array_equal(dna.score_rows(dm.data, st, enc.size, 14), want)
    # saturation: a matrix whose sums exceed 255
    big = np.full((20, 5), 40, np.uint8)
    seq.configure_wrap(19)
    st19 = O.stripe(O.DNA, enc, 32, 19)
    assert np.array_equal(seq.score(big), O.score_u8(big, st19, enc.size))
    assert int(seq.score(big).max()) == 255
    for rule in (0, 1):
        r, c = O.argmax(want, rule)
        mc = dna.argmax(want, rule)
        assert (mc.row, mc.col) == (r, c)
    assert dna.max(want) == O.max_(want)
    t = int(want.max()) - 3
    ro, co = O.threshold(want, t)
    assert [(h.row, h.col) for h in dna.threshold(want, t)] == list(zip(ro.tolist(), co.tolist()))


def test_score_protein(prot, kat, abyb1):
    p = kat["protein"]
    pssm = O.pssm_from_sequences(O.PROTEIN, p["patterns"], p["pseudocount"])
    for m in (15, 23):
        ps = pssm[:m]
        enc = prot.encode(abyb1)
        assert np.array_equal(enc, O.encode(O.PROTEIN, abyb1))
        st = O.stripe(O.PROTEIN, enc, 32, m - 1)
        want = O.score_f32(ps, st, enc.size)
        seq = prot.sequence_from_ascii(abyb1, m - 1)
        assert np.array_equal(seq.download(), st)
        assert np.array_equal(bits(seq.score(ps)), bits(want))
        r, c = O.argmax(want, O.RULE_AVX2)
        best = seq.score_argmax(ps)
        assert (best["row"], best["col"]) == (r, c)
    n = 300007
    sym = O.synth_protein(n)
    sym[::1000] = 20  # some X
    ps = pssm[:15]
    want = O.score_f32(ps, O.stripe(O.PROTEIN, sym, 32, 14), n)
    seq = prot.sequence_from_symbols(sym, 14)
    assert np.array_equal(bits(seq.score(ps)), bits(want))
    for rule in (0, 1, 2):
        r, c = O.argmax(want, rule)
        best = seq.score_argmax(ps, rule=rule)
        assert (best["row"], best["col"]) == (r, c)


def test_synthetic_matches_oracle(dna, prot):
    n = 100003
    for pli, alpha, gen in ((dna, O.DNA, O.synth_dna), (prot, O.PROTEIN, O.synth_protein)):
        want = O.stripe(alpha, gen(n), 32, 19)
        seq = pli.sequence_synthetic(n, 19)
        assert np.array_equal(seq.download(), want)
        rows = want.shape[0] - 19
        lo, cnt = 1000, 777  # a row slab generated in place, halo taken from the following rows
        slab = pli.sequence_synthetic(n, 19, row_begin=lo, n_rows=cnt)
        assert np.array_equal(slab.download(), want[lo:lo + cnt + 19])
        assert rows > lo + cnt


def test_batch_argmax(dna, jaspar):
    n = 200000
    sym = O.synth_dna(n)
    motifs = jaspar["motifs"][:48] + [jaspar["rest"]]
    pssms = [O.pssm_from_counts(O.DNA, c, 0.1) for c in motifs]
    maxw = max(len(p) for p in pssms)
    seq = dna.sequence_from_symbols(sym, maxw - 1)
    got = seq.score_argmax_batch(pssms)
    st = O.stripe(O.DNA, sym, 32, maxw - 1)
    for p, g in zip(pssms, got):
        want = O.score_f32(p, st, n, wrap=maxw - 1)
        r, c = O.argmax(want, O.RULE_AVX2)
        assert (g["row"], g["col"]) == (r, c)
        assert bits(np.float32(g["score"])) == bits(want[r, c])
        single = seq.score_argmax(p)
        assert (single["row"], single["col"]) == (r, c)


def test_batch_through_filter(dna, jaspar):
    """The many-motif job loop with per-motif planning: motifs with a profitable filter run the
    filter + recheck kernel, the rest the exact kernel; results == one exact call per motif, for
    argmax and for per-motif thresholds (hits in row-major order, exact scores)."""
    n = 6_000_000
    motifs = [m for m in jaspar["motifs"] if 6 <= len(m) <= 30][:40] + [jaspar["rest"]]
    pssms = [O.pssm_from_counts(O.DNA, c, 0.1) for c in motifs]
    maxw = max(len(p) for p in pssms)
    seq = dna.sequence_synthetic(n, maxw - 1)
    dna.set_prune(0)
    want_best = [seq.score_argmax(p) for p in pssms]
    thr = [float(np.float32(b["score"] - 2.5)) for b in want_best]
    want_hits = [seq.score_threshold(p, t) for p, t in zip(pssms, thr)]
    for mode in (1, 2):
        dna.set_prune(mode)
        got = seq.score_argmax_batch(pssms)
        assert dna.last_stats["pruned"] >= 1  # stats[2] = number of motifs that ran the filter
        for g, w in zip(got, want_best):
            assert (g["row"], g["col"], g["position"]) == (w["row"], w["col"], w["position"])
            assert bits(np.float32(g["score"])) == bits(np.float32(w["score"]))
        hits = seq.score_threshold_batch(pssms, thr, cap=64)  # small cap: exercises the overflow retry
        assert hits == want_hits
    dna.set_prune(1)


def test_batch_both_strands_cache_and_empty(lm, dna, jaspar):
    """Strand selection of the batch calls (reverse complement = rows reversed, A<->T, C<->G,
    pwm/mod.rs:566-577), the per-motif-set plan cache (same set twice, then a different set, then the
    first again) and the per-entry status (a motif longer than the sequence is None)."""
    n = 3_000_000
    motifs = [m for m in jaspar["motifs"] if 8 <= len(m) <= 24][:24] + [jaspar["rest"]]
    pssms = [O.pssm_from_counts(O.DNA, c, 0.1) for c in motifs]
    # the reverse complement against its definition
    p0 = pssms[0]
    rc = lm.reverse_complement(p0)
    assert np.array_equal(rc, p0[::-1][:, [2, 3, 0, 1, 4]])
    maxw = max(len(p) for p in pssms)
    seq = dna.sequence_synthetic(n, maxw - 1)
    dna.set_prune(0)
    want = []
    for p in pssms:
        want += [seq.score_argmax(p), seq.score_argmax(lm.reverse_complement(p))]
    dna.set_prune(2)
    for rep in range(2):
        got = seq.score_argmax_batch(pssms, strands=lm.STRAND_BOTH)
        assert got == want
    assert seq.score_argmax_batch(pssms[:5], strands=lm.STRAND_REVERSE) == want[1:10:2]
    assert seq.score_argmax_batch(pssms, strands=lm.STRAND_DIRECT) == want[0::2]
    thr = [float(np.float32(w["score"] - 3.0)) for w in want[0::2]]
    dna.set_prune(0)
    want_hits = []
    for p, t in zip(pssms, thr):
        want_hits += [seq.score_threshold(p, t), seq.score_threshold(lm.reverse_complement(p), t)]
    dna.set_prune(2)
    assert seq.score_threshold_batch(pssms, thr, cap=4096, strands=lm.STRAND_BOTH) == want_hits
    dna.set_prune(1)
    # a motif longer than the sequence: None for that entry only
    tiny = dna.sequence_from_symbols(O.synth_dna(12), maxw - 1)
    res = tiny.score_argmax_batch([pssms[0][:4], pssms[-1]])
    assert res[0] is not None and res[1] is None


def test_sharded_merge_equals_single_call(lm, dna, jaspar):
    """Row slabs scored independently + lm_argmax_merge == one call (the
    multi-GPU reduction, exercised on one device)."""
    n = 1000003
    pssm = O.pssm_from_counts(O.DNA, jaspar["rest"], 0.1)
    m = len(pssm)
    whole = dna.sequence_synthetic(n, m - 1)
    rows = whole.rows
    for rule in (0, 1, 2):
        ref = whole.score_argmax(pssm, rule=rule)
        cuts = [0, rows // 5, rows // 2, rows - 7, rows]
        parts, first = [], None
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            slab = dna.sequence_synthetic(n, m - 1, row_begin=lo, n_rows=hi - lo)
            res = slab.score_argmax(pssm, rule=rule, with_columns=True)
            parts.append(res["columns"])
            if first is None:
                first = res["first_cell"]
        merged = lm.argmax_merge(parts, cuts, first, rows, rule)
        assert (merged["row"], merged["col"], merged["position"]) == (ref["row"], ref["col"], ref["position"])
        assert merged["score"] == ref["score"]
    st = O.stripe(O.DNA, O.synth_dna(n), 32, m - 1)
    want = O.score_f32(pssm, st, n)
    r, c = O.argmax(want, O.RULE_AVX2)
    ref = whole.score_argmax(pssm)
    assert (ref["row"], ref["col"]) == (r, c)


def test_kernels_actually_launched(dna):
    before = dna.kernel_launches
    dna.encode("ACGT")
    assert dna.kernel_launches == before + 1


def test_launch_finish_pipeline(lm, dna, jaspar):
    """lm_seq_score_argmax_f32_launch / _finish: two scans in flight return what
    the one-shot call returns."""
    n = 30_000_003
    seq = dna.sequence_synthetic(n, 29)
    pssms = [O.pssm_from_counts(O.DNA, c, 0.1) for c in jaspar["motifs"][:6]] + [O.pssm_from_counts(O.DNA, jaspar["rest"], 0.1)]
    want = [seq.score_argmax(p, with_columns=True) for p in pssms]
    dna.set_reserve_sms(1)
    got = []
    seq.score_argmax_launch(pssms[0], slot=0)
    for k in range(len(pssms)):
        if k + 1 < len(pssms):
            seq.score_argmax_launch(pssms[k + 1], slot=(k + 1) & 1)
        got.append(seq.score_argmax_finish(slot=k & 1, with_columns=True))
    dna.set_reserve_sms(0)
    for w, g in zip(want, got):
        assert {k: w[k] for k in ("row", "col", "position", "score", "first_cell")} == \
               {k: g[k] for k in ("row", "col", "position", "score", "first_cell")}
    with pytest.raises(lm.LightmotifError):
        seq.score_argmax_finish(slot=1)  # nothing in flight


# ------------------------------------------------------------------ packed upload (upload.cu)

def _random_striped(rng, rows, wrap, n_frac=0.0):
    a = rng.integers(0, 4, size=(rows + wrap, 32), dtype=np.uint8)
    if n_frac:
        a[rng.random(a.shape) < n_frac] = 4
    return a


@pytest.mark.parametrize("mode", [4, 2])
@pytest.mark.parametrize("pinned", [True, False])
def test_packed_upload_is_byte_identical(lm, dna, mode, pinned):
    """lm_seq_from_striped of a large DNA matrix (plain chunks from the front, 2-bit packed chunks from the
    back, chunks with N as they are) leaves exactly the caller's bytes on the device."""
    import torch
    rng = np.random.default_rng(7 + mode)
    rows, wrap = (41 << 20) // 32 + 77, 19          # 11 chunks of 4 MiB, the last one ragged
    a = _random_striped(rng, rows, wrap)
    a[(3 << 20) // 32 * 8 + 5, 7] = 4               # one N in the seventh chunk: that chunk goes plain
    a[-1, 31] = 4                                   # and one in the very last byte
    if pinned:
        host = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
        host.numpy()[:] = a
        src = host.numpy()
    else:
        src = a
    dna.set_upload(mode, threads=3, min_bytes=0)
    try:
        seq = dna.sequence_from_striped(src, rows * 32, wrap)
        up = dna.last_upload
        got = seq.download()
        seq.free()
    finally:
        dna.set_upload(1, 0, 64 << 20)
    assert np.array_equal(got[: rows + wrap], a)
    assert up["packed_chunks"] + up["plain_chunks"] == -(-a.nbytes // (4 << 20)) and up["threads"] == 3
    assert up["plain_chunks"] >= 2                   # the two chunks holding N
    if mode == 2 or not pinned:
        assert up["plain_chunks"] == 2 and up["pcie_bytes"] < a.nbytes // 2


def test_packed_upload_scan_equals_plain(lm, dna, jaspar):
    """Same fused scan results whichever way the matrix travelled."""
    rng = np.random.default_rng(11)
    pssm = O.pssm_from_counts(O.DNA, jaspar["rest"], 0.1)
    m = pssm.shape[0]
    rows, wrap = (20 << 20) // 32, m - 1
    a = _random_striped(rng, rows, wrap, n_frac=1e-7)
    a[rows:] = np.roll(a[:wrap], -1, axis=1)          # wrap rows as configure_wrap leaves them (seq.rs:384-400)
    a[rows:, 31] = 4
    res = []
    for mode in (0, 4, 2, 3):
        dna.set_upload(mode, threads=2, min_bytes=0)
        try:
            seq = dna.sequence_from_striped(a, rows * 32, wrap)
            res.append((seq.score_argmax(pssm), seq.score_threshold(pssm, 8.0)))
            seq.free()
        finally:
            dna.set_upload(1, 0, 64 << 20)
    assert res[0] == res[1] == res[2] == res[3]
    assert res[0][0] is not None


def test_packed_upload_of_flat_symbols(lm, dna):
    """lm_seq_from_symbols of a long encoded DNA sequence travels through the same two-ended upload (ragged
    length, N inside, unaligned tail) and is striped on the device exactly as before."""
    rng = np.random.default_rng(21)
    n = (13 << 20) + 37
    sym = rng.integers(0, 4, n, dtype=np.uint8)
    sym[5_000_000:5_000_300] = 4
    sym[-1] = 4
    want = O.stripe(O.DNA, sym, 32, 9)
    for mode in (0, 4, 2):
        dna.set_upload(mode, threads=3, min_bytes=0)
        try:
            seq = dna.sequence_from_symbols(sym, 9)
            got = seq.download()
            up = dna.last_upload
            seq.free()
        finally:
            dna.set_upload(1, 0, 64 << 20)
        assert np.array_equal(got, want), mode
        if mode:
            assert up["packed_chunks"] >= 1 and up["pcie_bytes"] < n


def test_adaptive_upload_measures_both_ways_then_keeps_one(lm, dna):
    """LM_OPT_UPLOAD = 1: the first large upload is two-ended, the second plain, later ones whichever was
    faster -- and the device matrix is the caller's bytes every time."""
    import torch
    rng = np.random.default_rng(31)
    rows, wrap = (24 << 20) // 32, 19
    a = _random_striped(rng, rows, wrap)
    host = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
    host.numpy()[:] = a
    dna.set_upload(1, threads=3, min_bytes=0)
    ways = []
    try:
        dna.set_upload(4, threads=3, min_bytes=0)          # (the ring exists before the measurements start)
        dna.sequence_from_striped(host.numpy(), rows * 32, wrap).free()
        dna.set_upload(1, threads=3, min_bytes=0)
        for _ in range(4):
            seq = dna.sequence_from_striped(host.numpy(), rows * 32, wrap)
            ways.append(dna.last_upload["threads"] > 0)
            assert np.array_equal(seq.download(), a)
            seq.free()
    finally:
        dna.set_upload(1, 0, 64 << 20)
    assert ways[0] is True and ways[1] is False and ways[2] == ways[3]


@pytest.mark.parametrize("m", [500, 560])
def test_score_very_long_motifs(dna, m):
    """Motifs so long that their tables crowd the exact kernel's shared memory: at m = 500 the 4-row prefix
    table still fits, at m = 560 the kernel must do without it -- bit-identical scores either way."""
    rng = np.random.default_rng(m)
    pssm = random_pssm(rng, m, 5)
    n = 60_013
    sym = rng.integers(0, 4, n).astype(np.uint8)   # no N: the paired-position path
    seq = dna.sequence_from_symbols(sym, m - 1)
    got = seq.score(pssm)
    want = O.score_f32(pssm, O.stripe(O.DNA, sym, 32, m - 1), n)
    assert np.array_equal(bits(got), bits(want))
    seq.free()
