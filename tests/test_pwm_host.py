"""The harness's host-side matrix constructors (lightmotif_b200/pwm.py) give the oracle's pinned
matrices bit for bit (JASPAR count matrices, the reference's test patterns, protein)."""
import numpy as np

from lightmotif_b200 import pwm
from oracle import oracle as O


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_pssm_from_counts_matches_oracle(jaspar):
    for counts in jaspar["motifs"][:200] + [jaspar["rest"]]:
        for pc in (0.1, 0.0, 1.0):
            assert np.array_equal(bits(pwm.pssm_from_counts(pwm.DNA, counts, pc)), bits(O.pssm_from_counts(O.DNA, counts, pc)))


def test_pssm_from_sequences_matches_oracle(kat):
    d = kat["dna"]
    assert np.array_equal(bits(pwm.pssm_from_sequences(pwm.DNA, d["patterns"], d["pseudocount"])),
                          bits(O.pssm_from_sequences(O.DNA, d["patterns"], d["pseudocount"])))
    p = kat["protein"]
    assert np.array_equal(bits(pwm.pssm_from_sequences(pwm.PROTEIN, p["patterns"], p["pseudocount"])),
                          bits(O.pssm_from_sequences(O.PROTEIN, p["patterns"], p["pseudocount"])))
