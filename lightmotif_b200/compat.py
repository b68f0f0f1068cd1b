"""`lightmotif`-compatible Python surface (the hot-path subset of the PyO3 module
``lightmotif.lib``, lightmotif-py/lightmotif/lib.rs and lib.pyi) on the CUDA back-end, so that the
reference's own Python tests for the pipeline (lightmotif-py/lightmotif/tests/test_pipeline.py,
test_scanner.py, test_sequence.py) read the same against this engine:

    import lightmotif_b200.compat as lightmotif
    motif  = lightmotif.create(["GTTGACCTTATCAAC", "GTTGATCCAGTCAAC"])
    pssm   = motif.counts.normalize(0.1).log_odds()
    seq    = lightmotif.stripe("ATGTCCCAACAACGATACC...")
    scores = pssm.calculate(seq)            # StripedScores: len, [], iteration, argmax, max, threshold
    hits   = list(lightmotif.scan(pssm, seq, threshold=-10.0))

Covered: EncodedSequence, StripedSequence, CountMatrix.normalize, WeightMatrix.log_odds,
ScoringMatrix.calculate, StripedScores.{argmax,max,threshold}, Scanner / scan / Hit, Motif / create /
stripe.  Out of scope: motif file parsers (`load`), p-values
(`ScoringMatrix.pvalue/score`, `ScoreDistribution`), custom backgrounds: they raise NotImplementedError.
Everything that touches sequence data runs on the GPU; there is no CPU fall-back.
"""
from __future__ import annotations

import numpy as np

from . import DNA, PROTEIN, RULE_AVX2, InvalidSymbol, Pipeline, pwm
from . import Scanner as _Scanner

_pipelines = {}


def _pli(protein):
    key = bool(protein)
    if key not in _pipelines:
        _pipelines[key] = Pipeline.cuda(PROTEIN if key else DNA)
    return _pipelines[key]


def _alphabet(protein):
    return PROTEIN if protein else DNA


class EncodedSequence:
    """lib.rs EncodedSequence: symbols as indices (A, C, T, G, N = 0..4)."""

    def __init__(self, sequence, *, protein=False):
        self.protein = bool(protein)
        try:
            self._data = _pli(protein).encode(sequence)
        except InvalidSymbol as e:  # lib.rs: ValueError("Invalid symbol in sequence")
            raise ValueError(f"Invalid symbol in sequence: {e}") from None

    def __len__(self):
        return int(self._data.size)

    def __getitem__(self, index):
        if not -len(self) <= index < len(self):
            raise IndexError("sequence index out of range")
        return int(self._data[index])

    def __iter__(self):
        return iter(int(x) for x in self._data)

    def __str__(self):
        return "".join(pwm._SYMBOLS[_alphabet(self.protein)][x] for x in self._data)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __buffer__(self, flags):  # lib.rs:220-249: read-only 1-D buffer of symbol codes
        v = self._data.view()
        v.flags.writeable = False
        return memoryview(v)

    def copy(self):
        c = object.__new__(EncodedSequence)
        c.protein, c._data = self.protein, self._data.copy()
        return c

    __copy__ = copy

    def stripe(self):
        return StripedSequence._from_symbols(self._data, self.protein)


class StripedSequence:
    """lib.rs StripedSequence: [rows][32] column-major striping, resident on the GPU."""

    @classmethod
    def _from_symbols(cls, sym, protein):
        s = object.__new__(cls)
        s.protein, s._symbols = bool(protein), np.ascontiguousarray(sym, dtype=np.uint8)
        s._dev = _pli(protein).sequence_from_symbols(s._symbols, 0)
        return s

    def __len__(self):
        return int(self._symbols.size)

    def _view(self):
        # lib.rs:304-317, 334-364: the matrix (wrap rows included) as a read-only Fortran-ordered buffer
        # of shape [32 columns][rows] -- mem[c, r] is the symbol at sequence index c * rows + r
        v = self._dev.download().T
        v.flags.writeable = False
        return v

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._view(), dtype=dtype)

    def __buffer__(self, flags):
        return memoryview(self._view())

    def copy(self):
        return StripedSequence._from_symbols(self._symbols, self.protein)

    __copy__ = copy


def stripe(sequence, *, protein=False):
    return EncodedSequence(sequence, protein=protein).stripe()


class CountMatrix:
    def __init__(self, values, *, protein=False):
        self.protein = bool(protein)
        sym = pwm._SYMBOLS[_alphabet(protein)]
        if isinstance(values, dict):
            m = len(next(iter(values.values())))
            self._counts = np.zeros((m, len(sym)), dtype=np.uint32)
            for ch, col in values.items():
                if ch not in sym or len(col) != m:
                    raise ValueError("Invalid count matrix")
                self._counts[:, sym.index(ch)] = np.asarray(list(col), dtype=np.uint32)
        else:
            self._counts = np.asarray(values, dtype=np.uint32)

    def __len__(self):
        return int(self._counts.shape[0])

    def __getitem__(self, index):
        return [int(x) for x in self._counts[index]]

    def __eq__(self, other):
        return isinstance(other, CountMatrix) and self.protein == other.protein and np.array_equal(self._counts, other._counts)

    def normalize(self, pseudocount=None):
        """to_freq(pseudocount).to_weight(None) (lib.rs:501-526)."""
        if isinstance(pseudocount, dict):
            raise NotImplementedError("per-symbol pseudocounts are outside the hot-path scope")
        freq = pwm.to_freq(_alphabet(self.protein), self._counts, 0.0 if pseudocount is None else float(pseudocount))
        return WeightMatrix(freq, self.protein)


class WeightMatrix:
    def __init__(self, freq, protein):
        self.protein, self._freq = bool(protein), np.asarray(freq, dtype=np.float32)

    def __len__(self):
        return int(self._freq.shape[0])

    def __eq__(self, other):
        return isinstance(other, WeightMatrix) and self.protein == other.protein and np.array_equal(self._freq, other._freq)

    def log_odds(self, background=None, base=2.0):
        if background is not None or base != 2.0:
            raise NotImplementedError("custom backgrounds / bases are outside the hot-path scope")
        return ScoringMatrix._from_array(pwm.to_scoring(_alphabet(self.protein), self._freq), self.protein)


class ScoringMatrix:
    def __init__(self, values, background=None, *, protein=False):
        if background is not None:
            raise NotImplementedError("custom backgrounds are outside the hot-path scope")
        # lib.rs:729-745: the matrix starts as DenseMatrix::new (zero-filled), so a symbol missing from
        # the dict (typically N) weighs 0.0; columns are taken in alphabet order and must agree in length
        sym = pwm._SYMBOLS[_alphabet(protein)]
        data = None
        for j, ch in enumerate(sym):
            if ch not in values:
                continue
            col = np.asarray(list(values[ch]), dtype=np.float32)
            if data is None:
                data = np.zeros((len(col), len(sym)), dtype=np.float32)
            if data.shape[0] != len(col):
                raise ValueError("Invalid number of rows")
            data[:, j] = col
        if data is None:
            raise ValueError("Invalid count matrix")
        self.protein, self._data = bool(protein), data

    @classmethod
    def _from_array(cls, data, protein):
        s = object.__new__(cls)
        s.protein, s._data = bool(protein), np.ascontiguousarray(data, dtype=np.float32)
        return s

    def __len__(self):
        return int(self._data.shape[0])

    def __eq__(self, other):
        return isinstance(other, ScoringMatrix) and self.protein == other.protein and np.array_equal(self._data, other._data)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __buffer__(self, flags):  # lib.rs:676-697, 789-819: [M rows][K] f32, read-only
        v = self._data.view()
        v.flags.writeable = False
        return memoryview(v)

    def calculate(self, sequence):
        """lib.rs:855-874: configure the wrap rows for this matrix, then score every position."""
        if sequence.protein != self.protein:
            raise ValueError("alphabet mismatch")
        sequence._dev.configure(self._data)
        return StripedScores(sequence._dev.score(self._data), len(sequence) + 1 - len(self) if len(sequence) >= len(self) else 0,
                             self.protein)

    def pvalue(self, score, method="meme"):
        raise NotImplementedError("p-values (pwm/dist.rs, lightmotif-tfmpvalue) are outside the hot-path scope")

    score = pvalue

    def reverse_complement(self):
        if self.protein:
            raise ValueError("cannot complement a protein matrix")
        # pwm/mod.rs:566-577: rows reversed, A<->T and C<->G swapped (codes A, C, T, G, N)
        return ScoringMatrix._from_array(self._data[::-1][:, [2, 3, 0, 1, 4]], False)


class StripedScores:
    """lib.rs StripedScores: positions are `col * rows + row`; len() = max_index."""

    def __init__(self, matrix, max_index, protein):
        self._m, self._n, self.protein = matrix, int(max_index), protein

    def __len__(self):
        return min(self._n, self._m.shape[0] * 32)

    def __getitem__(self, index):
        if not 0 <= index < len(self):
            raise IndexError("list index out of range")
        rows = self._m.shape[0]
        return float(self._m[index % rows, index // rows])

    def __iter__(self):
        rows = self._m.shape[0]
        return iter(float(self._m[i % rows, i // rows]) for i in range(len(self)))

    def _view(self):
        # lib.rs:1128-1155: Fortran-ordered buffer of shape [32 columns][rows]
        v = self._m.T
        v.flags.writeable = False
        return v

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._view(), dtype=dtype)

    def __buffer__(self, flags):
        return memoryview(self._view())

    def threshold(self, threshold):
        rows = self._m.shape[0]
        return [mc.col * rows + mc.row for mc in _pli(self.protein).threshold(self._m, float(threshold))]

    def max(self):
        return _pli(self.protein).max(self._m)

    def argmax(self):
        mc = _pli(self.protein).argmax(self._m, RULE_AVX2)
        return None if mc is None else mc.col * self._m.shape[0] + mc.row


class Hit:
    def __init__(self, position, score):
        self.position, self.score = int(position), float(score)

    def __repr__(self):
        return f"Hit(position={self.position}, score={self.score})"


class Scanner:
    """lib.rs Scanner (scan.rs:96-250) over the fused filter + recheck kernel; DNA only, like the reference."""

    def __init__(self, pssm, sequence, threshold=0.0, block_size=256):
        if pssm.protein or sequence.protein:
            raise ValueError("Scanner is only available for DNA motifs")
        sequence._dev.configure(pssm._data)
        self._hits = _Scanner(pssm._data, sequence._dev).threshold(threshold).block_size(block_size).collect()
        self._hits.reverse()  # the reference pops hits off a stack; order is unspecified, tests sort

    def __iter__(self):
        return self

    def __next__(self):
        if not self._hits:
            raise StopIteration
        h = self._hits.pop()
        return Hit(h["position"], h["score"])


def scan(pssm, sequence, *, threshold=0.0, block_size=256):
    return Scanner(pssm, sequence, threshold, block_size)


class Motif:
    def __init__(self, counts, name=None):
        self.counts, self.name, self.protein = counts, name, counts.protein

    @property
    def pwm(self):
        return self.counts.normalize(0.0)

    @property
    def pssm(self):
        return self.counts.normalize(0.0).log_odds()


def create(sequences, *, protein=False, name=None):
    """lib.rs create(): count matrix from equal-length sequences; ValueError on an invalid symbol."""
    sequences = list(sequences)
    sym = pwm._SYMBOLS[_alphabet(protein)]
    for s in sequences:
        if any(ch not in sym for ch in s):
            raise ValueError("Invalid symbol in sequence")
    return Motif(CountMatrix(pwm.counts_from_sequences(_alphabet(protein), sequences), protein=protein), name)


LIMITED_API = False  # lightmotif.lib.LIMITED_API: the buffer protocol is available (PEP 688 __buffer__)
AVX2_SUPPORTED = False  # lightmotif.lib: the CPU back-ends are not what runs here


def load(file, format="jaspar", *, protein=False):
    raise NotImplementedError("motif file parsers (lightmotif-io) are outside the hot-path scope")
