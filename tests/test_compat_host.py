"""Host-only parts of the lightmotif-compatible surface (no GPU): matrix constructors against the
oracle, reverse complement, out-of-scope entry points fail loudly."""
import numpy as np
import pytest

from oracle import oracle as O


def test_matrices_match_oracle(kat):
    import lightmotif_b200.compat as lightmotif
    d = kat["dna"]
    motif = lightmotif.create(d["patterns"], name="MX000001")
    assert motif.name == "MX000001" and not motif.protein and len(motif.counts) == 15
    assert motif.counts[0] == [0, 0, 0, 2, 0]  # both patterns start with G
    pssm = motif.counts.normalize(0.1).log_odds()
    want = O.pssm_from_sequences(O.DNA, d["patterns"], 0.1)
    assert np.array_equal(np.asarray(pssm).view(np.uint32), want.view(np.uint32))
    cm = lightmotif.CountMatrix({"A": [1, 0], "C": [0, 2], "T": [3, 0], "G": [0, 0]})
    assert cm == lightmotif.CountMatrix(np.array([[1, 0, 3, 0, 0], [0, 2, 0, 0, 0]]))
    rc = pssm.reverse_complement()
    assert np.array_equal(np.asarray(rc)[0, [2, 3, 0, 1]], np.asarray(pssm)[-1, [0, 1, 2, 3]])
    for call in (lambda: pssm.pvalue(1.0), lambda: lightmotif.load("x.jaspar"), lambda: motif.counts.normalize({"A": 0.1})):
        with pytest.raises(NotImplementedError):
            call()
    with pytest.raises(ValueError):
        lightmotif.create(["ATG."])


def test_scoring_matrix_from_dict_zero_fills_missing_symbols():
    """lib.rs:729-745: the matrix starts zero-filled, so a symbol missing from the dict (typically N)
    weighs 0.0, not -inf; columns of unequal length are rejected."""
    import lightmotif_b200.compat as lightmotif
    sm = lightmotif.ScoringMatrix({"A": [1.0, -2.0], "C": [0.5, 0.25], "T": [-1.0, 3.0], "G": [0.0, -0.5]})
    a = np.asarray(sm)
    assert a.shape == (2, 5) and a[:, 4].tolist() == [0.0, 0.0] and a[1].tolist() == [-2.0, 0.25, 3.0, -0.5, 0.0]
    with pytest.raises(ValueError):
        lightmotif.ScoringMatrix({"A": [1.0, 2.0], "C": [1.0]})
    with pytest.raises(ValueError):
        lightmotif.ScoringMatrix({})
    assert memoryview(sm).shape == (2, 5)
