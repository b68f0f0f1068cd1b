"""Host-side matrix constructors of the Python harness -- the O(M x K) part of lightmotif's
`pwm` module that feeds the scoring pipeline (host code in the reference too):

    CountMatrix::from_sequences   lightmotif/src/pwm/mod.rs:209-237
    CountMatrix::to_freq          pwm/mod.rs:240-258  (pseudocount 0 for the default symbol, abc.rs:558-574;
                                                       rows normalised by their sequential f32 sum)
    FrequencyMatrix::to_scoring   pwm/mod.rs:415-430  (log2(x / f), f == 0 -> -inf; uniform background
                                                       abc.rs:465-487)

`f32::log2` lowers to the platform libm's `log2f`; the same function is called here through ctypes so
that the weights are the bits the reference produces on this platform.  Mirrors the C++ host layer
(include/lightmotif_b200.hpp: CountMatrix / FrequencyMatrix / ScoringMatrix).
"""
from __future__ import annotations

import ctypes as C
import ctypes.util

import numpy as np

from .capi import DNA, PROTEIN

_SYMBOLS = {DNA: "ACTGN", PROTEIN: "ACDEFGHIKLMNPQRSTVWYX"}

_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.log2f.argtypes = [C.c_float]
_libm.log2f.restype = C.c_float


def counts_from_sequences(alphabet, patterns):
    """CountMatrix::from_sequences: patterns = equal-length upper-case strings."""
    sym = _SYMBOLS[alphabet]
    m = len(patterns[0])
    counts = np.zeros((m, len(sym)), dtype=np.uint32)
    for p in patterns:
        if len(p) != m:
            raise ValueError("invalid data: sequences differ in length")  # err.rs InvalidData
        for j, ch in enumerate(p):
            counts[j, sym.index(ch)] += 1
    return counts


def to_freq(alphabet, counts, pseudo):
    counts = np.asarray(counts, dtype=np.uint32)
    k = len(_SYMBOLS[alphabet])
    pc = np.full(k, np.float32(pseudo), dtype=np.float32)
    pc[k - 1] = 0.0
    freq = counts.astype(np.float32) + pc
    s = np.zeros(freq.shape[0], dtype=np.float32)
    for a in range(k):  # Iterator::sum over the row: sequential f32 adds
        s = (s + freq[:, a]).astype(np.float32)
    return (freq / s[:, None]).astype(np.float32)


def to_scoring(alphabet, freq, background=None):
    freq = np.asarray(freq, dtype=np.float32)
    k = len(_SYMBOLS[alphabet])
    if background is None:
        background = np.full(k, np.float32(1.0) / np.float32(k - 1), dtype=np.float32)
        background[k - 1] = 0.0
    bg = np.asarray(background, dtype=np.float32)
    out = np.empty(freq.shape, dtype=np.float32)
    for j in range(freq.shape[0]):
        for a in range(k):
            out[j, a] = -np.inf if bg[a] == 0.0 else _libm.log2f(np.float32(freq[j, a] / bg[a]))
    return out


def pssm_from_counts(alphabet, counts, pseudo=0.1, background=None):
    """CountMatrix(counts).to_freq(pseudo).to_scoring(background)."""
    return to_scoring(alphabet, to_freq(alphabet, counts, pseudo), background)


def pssm_from_sequences(alphabet, patterns, pseudo=0.1, background=None):
    return pssm_from_counts(alphabet, counts_from_sequences(alphabet, patterns), pseudo, background)
