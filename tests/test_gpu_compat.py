"""The reference's own Python tests for the pipeline (lightmotif-py/lightmotif/tests/test_pipeline.py,
test_scanner.py, test_sequence.py), replayed against `lightmotif_b200.compat` -- the lightmotif-compatible
surface on the CUDA back-end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lightmotif():
    import lightmotif_b200.compat as compat
    return compat


def test_calculate(lightmotif, kat):  # test_pipeline.py TestDNAMotif.test_calculate
    d = kat["dna"]
    motif = lightmotif.create(d["patterns"])
    pssm = motif.counts.normalize(0.1).log_odds()
    seq = lightmotif.stripe(d["sequence"])
    scores = pssm.calculate(seq)
    assert len(scores) == len(d["expected"])
    for x, y in zip(scores, d["expected"]):
        assert abs(x - y) < 0.5e-5
    assert scores.argmax() == d["argmax"]
    assert abs(scores.max() - (-5.50167)) < 1e-5
    assert sorted(scores.threshold(-10.0)) == d["threshold"]["-10.0"]
    assert sorted(scores.threshold(-15.0)) == d["threshold"]["-15.0"]
    assert scores.threshold(10.0) == []
    assert scores[0] == pytest.approx(-23.07094, abs=1e-6)
    with pytest.raises(IndexError):
        scores[len(scores)]


def test_scan(lightmotif, kat):  # test_scanner.py TestScanner.test_scan
    d = kat["dna"]
    pssm = lightmotif.create(d["patterns"]).counts.normalize(0.1).log_odds()
    seq = lightmotif.stripe(d["sequence"])
    assert len(list(lightmotif.scan(pssm, seq))) == 0
    hits = list(lightmotif.scan(pssm, seq, threshold=-10.0))
    assert len(hits) == 3
    hits.sort(key=lambda h: h.position)
    assert [h.position for h in hits] == [18, 27, 32]
    assert abs(hits[0].score - (-5.50167)) < 1e-5
    assert abs(hits[1].score - (-6.4345555)) < 1e-5
    assert abs(hits[2].score - (-8.961102)) < 1e-5


def test_sequences(lightmotif):  # test_sequence.py
    A, C, T, G, N = range(5)
    s1 = lightmotif.EncodedSequence("ATGC")
    s2 = lightmotif.EncodedSequence("ATGCTTAGATAC")
    assert (len(s1), len(s2)) == (4, 12)
    assert list(s1) == [A, T, G, C] and [s2[i] for i in range(7)] == [A, T, G, C, T, T, A]
    with pytest.raises(IndexError):
        s1[10]
    assert str(s1) == "ATGC" and np.asarray(s1).tolist() == [A, T, G, C]
    st = lightmotif.stripe("ATGC")
    mem = memoryview(st)  # lib.rs: Fortran-ordered [32 columns][rows]; mem[c, r] = symbol c * rows + r
    assert mem.shape == (32, 1) and [mem[c, 0] for c in range(4)] == [A, T, G, C] and mem[4, 0] == N
    assert memoryview(s1).shape == (4,) and memoryview(s1)[1] == T
    with pytest.raises(ValueError):
        lightmotif.EncodedSequence("ATG.")
    with pytest.raises(ValueError):
        lightmotif.create(["PILFFRLK", "KDMLKEYL", "PFRLTHKL"])
    motif = lightmotif.create(["PILFFRLK", "KDMLKEYL", "PFRLTHKL"], protein=True)
    pssm = motif.counts.normalize(0.1).log_odds()
    scores = pssm.calculate(lightmotif.stripe("MKDMLKEYLPILFFRLKAA", protein=True))
    assert len(scores) == 19 - 8 + 1 and scores.argmax() in (1, 9)
    with pytest.raises(ValueError):
        pssm.calculate(lightmotif.stripe("ATGC"))
