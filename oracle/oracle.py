"""ctypes front-end of the CPU oracle (``oracle/lm_oracle.c``).

TEST INFRASTRUCTURE ONLY.  Import this from ``tests/``, from ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg and from
``__graft_entry__.smoke()`` -- never from ``lightmotif_b200`` (the product).

Each helper is a thin numpy wrapper over one C function; the C function cites
the reference file:line it restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liblm_oracle.so"

DNA, PROTEIN = 0, 1
RULE_GENERIC, RULE_AVX2, RULE_SSE2 = 0, 1, 2
LANES = 32


class InvalidSymbol(ValueError):
    """Mirror of lightmotif::err::InvalidSymbol(char) (err.rs:10)."""

    def __init__(self, ch: str, index: int):
        super().__init__(f"invalid symbol {ch!r} at index {index}")
        self.ch = ch
        self.index = index


def build(force: bool = False) -> Path:
    srcs = [_HERE / "lm_oracle.c", _HERE / "lm_avx2_ref.c", _HERE / "lm_oracle.h"]
    stale = (not _LIB_PATH.exists()) or any(
        s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True,
                       env={**os.environ, "CC": os.environ.get("CC", "gcc")})
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _declare(_lib)
    return _lib


_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)


def _declare(L):
    sz = C.c_size_t
    L.lmo_alphabet_k.restype = C.c_int
    L.lmo_alphabet_str.restype = C.c_char_p
    L.lmo_default_symbol.restype = C.c_uint8
    L.lmo_encode.argtypes = [C.c_int, _u8p, sz, _u8p, _szp]
    L.lmo_encode.restype = C.c_int
    L.lmo_encode_lossy.argtypes = [C.c_int, _u8p, sz, _u8p]
    L.lmo_encode_lossy.restype = None
    L.lmo_stripe_rows.argtypes = [sz, sz]
    L.lmo_stripe_rows.restype = sz
    L.lmo_stripe.argtypes = [_u8p, sz, sz, C.c_uint8, _u8p, sz]
    L.lmo_stripe.restype = None
    L.lmo_configure_wrap.argtypes = [_u8p, sz, sz, sz, sz, C.c_uint8]
    L.lmo_configure_wrap.restype = None
    L.lmo_score_f32_rows.argtypes = [_f32p, sz, sz, _u8p, sz, sz, sz, sz, sz, _f32p, sz]
    L.lmo_score_f32_rows.restype = sz
    L.lmo_score_u8_rows.argtypes = [_u8p, sz, sz, _u8p, sz, sz, sz, sz, sz, _u8p, sz, C.c_int]
    L.lmo_score_u8_rows.restype = sz
    L.lmo_argmax_f32.argtypes = [_f32p, sz, sz, sz, C.c_int, _szp, _szp]
    L.lmo_argmax_f32.restype = C.c_int
    L.lmo_argmax_u8.argtypes = [_u8p, sz, sz, sz, C.c_int, _szp, _szp]
    L.lmo_argmax_u8.restype = C.c_int
    L.lmo_max_f32.argtypes = [_f32p, sz, sz, sz, C.c_int, _f32p]
    L.lmo_max_f32.restype = C.c_int
    L.lmo_max_u8.argtypes = [_u8p, sz, sz, sz, _u8p]
    L.lmo_max_u8.restype = C.c_int
    L.lmo_threshold_f32.argtypes = [_f32p, sz, sz, sz, C.c_float, _u64p, _u32p, sz]
    L.lmo_threshold_f32.restype = sz
    L.lmo_threshold_u8.argtypes = [_u8p, sz, sz, sz, C.c_uint8, _u64p, _u32p, sz]
    L.lmo_threshold_u8.restype = sz
    L.lmo_counts_from_sequences.argtypes = [C.c_int, C.POINTER(_u8p), sz, sz, _u32p]
    L.lmo_counts_from_sequences.restype = None
    L.lmo_to_freq.argtypes = [C.c_int, _u32p, sz, C.c_float, _f32p]
    L.lmo_to_freq.restype = None
    L.lmo_to_scoring.argtypes = [C.c_int, _f32p, sz, _f32p, _f32p]
    L.lmo_to_scoring.restype = None
    L.lmo_min_score.argtypes = [C.c_int, _f32p, sz]
    L.lmo_min_score.restype = C.c_float
    L.lmo_max_score.argtypes = [C.c_int, _f32p, sz]
    L.lmo_max_score.restype = C.c_float
    L.lmo_to_discrete.argtypes = [C.c_int, _f32p, sz, _u8p, _f32p, _f32p, _f32p]
    L.lmo_to_discrete.restype = None
    L.lmo_scale.argtypes = [C.c_float, C.c_float, C.c_float]
    L.lmo_scale.restype = C.c_uint8
    L.lmo_unscale.argtypes = [C.c_uint8, C.c_float, C.c_float]
    L.lmo_unscale.restype = C.c_float
    L.lmo_score_position.argtypes = [_f32p, sz, sz, _u8p, sz, sz, sz]
    L.lmo_score_position.restype = C.c_float
    L.lmo_scanner_collect.argtypes = [C.c_int, _f32p, sz, _u8p, sz, sz, sz, sz, sz,
                                      C.c_float, sz, _u64p, _f32p, sz]
    L.lmo_scanner_collect.restype = sz
    L.lmo_scanner_max.argtypes = [C.c_int, _f32p, sz, _u8p, sz, sz, sz, sz, sz,
                                  C.c_float, sz, _u64p, _f32p]
    L.lmo_scanner_max.restype = C.c_int
    L.lmo_synth_dna.argtypes = [C.c_uint64, C.c_uint64, sz, _u8p]
    L.lmo_synth_dna.restype = None
    L.lmo_synth_protein.argtypes = [C.c_uint64, C.c_uint64, sz, _u8p]
    L.lmo_synth_protein.restype = None
    L.lmo_avx2_score_f32_dna.argtypes = [_f32p, sz, _u8p, sz, sz, _f32p]
    L.lmo_avx2_score_f32_dna.restype = None
    L.lmo_avx2_argmax_f32.argtypes = [_f32p, sz, _szp, _szp]
    L.lmo_avx2_argmax_f32.restype = C.c_int
    L.lmo_avx2_score_argmax_mt.argtypes = [_f32p, sz, _u8p, sz, C.c_int, _f32p,
                                           _f32p, _szp, _szp]
    L.lmo_avx2_score_argmax_mt.restype = C.c_int


def _p(a, t):
    return a.ctypes.data_as(t)


def aligned_empty(shape, dtype, align=64):
    """numpy array whose data pointer is `align`-byte aligned."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    raw = np.empty(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


# ------------------------------------------------------------------ alphabet

def alphabet_k(alphabet):
    return lib().lmo_alphabet_k(alphabet)


def alphabet_str(alphabet):
    return lib().lmo_alphabet_str(alphabet).decode()


def default_symbol(alphabet):
    return alphabet_k(alphabet) - 1


def _as_bytes(seq):
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    return np.frombuffer(bytes(seq), dtype=np.uint8) if not isinstance(seq, np.ndarray) else seq


def encode(alphabet, seq):
    a = np.ascontiguousarray(_as_bytes(seq))
    dst = np.empty(a.size, dtype=np.uint8)
    bad = C.c_size_t(0)
    rc = lib().lmo_encode(alphabet, _p(a, _u8p), a.size, _p(dst, _u8p), C.byref(bad))
    if rc != 0:
        raise InvalidSymbol(chr(int(a[bad.value])), bad.value)
    return dst


def encode_lossy(alphabet, seq):
    a = np.ascontiguousarray(_as_bytes(seq))
    dst = np.empty(a.size, dtype=np.uint8)
    lib().lmo_encode_lossy(alphabet, _p(a, _u8p), a.size, _p(dst, _u8p))
    return dst


# -------------------------------------------------------------------- stripe

def stripe(alphabet, sym, c=LANES, wrap=0):
    """Stripe::stripe followed by configure_wrap(wrap). Returns [rows+wrap, c]."""
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    rows = lib().lmo_stripe_rows(sym.size, c)
    pad = default_symbol(alphabet)
    data = aligned_empty((rows + wrap, c), np.uint8)
    data[rows:, :] = pad  # DenseMatrix::resize default-fills new rows (dense.rs)
    lib().lmo_stripe(_p(sym, _u8p), sym.size, c, pad, _p(data, _u8p), c)
    if wrap:
        lib().lmo_configure_wrap(_p(data, _u8p), rows, c, c, wrap, pad)
    return data


# --------------------------------------------------------------------- score

def score_f32(pssm, striped, length, row_begin=0, row_end=None, wrap=None):
    """Score::score_rows_into (generic). pssm [m,k] f32; striped [rows+wrap,c]."""
    pssm = np.ascontiguousarray(pssm, dtype=np.float32)
    m, k = pssm.shape
    c = striped.shape[1]
    if wrap is None:
        wrap = m - 1
    rows = striped.shape[0] - wrap
    if row_end is None:
        row_end = rows
    n = max(row_end - row_begin, 0)
    out = aligned_empty((n, c), np.float32)
    got = lib().lmo_score_f32_rows(_p(pssm, _f32p), m, k, _p(striped, _u8p), c, c,
                                   length, row_begin, row_end, _p(out, _f32p), c)
    return out[:got]


def score_u8(dm, striped, length, row_begin=0, row_end=None, wrap=None, saturate=True):
    dm = np.ascontiguousarray(dm, dtype=np.uint8)
    m, k = dm.shape
    c = striped.shape[1]
    if wrap is None:
        wrap = m - 1
    rows = striped.shape[0] - wrap
    if row_end is None:
        row_end = rows
    n = max(row_end - row_begin, 0)
    out = aligned_empty((n, c), np.uint8)
    got = lib().lmo_score_u8_rows(_p(dm, _u8p), m, k, _p(striped, _u8p), c, c, length,
                                  row_begin, row_end, _p(out, _u8p), c, int(saturate))
    return out[:got]


def argmax(scores, rule=RULE_GENERIC):
    """Maximum::argmax -> (row, col) or None."""
    rows, c = scores.shape
    r, col = C.c_size_t(0), C.c_size_t(0)
    if scores.dtype == np.float32:
        rc = lib().lmo_argmax_f32(_p(scores, _f32p), rows, c, scores.strides[0] // 4, rule,
                                  C.byref(r), C.byref(col))
    else:
        rc = lib().lmo_argmax_u8(_p(scores, _u8p), rows, c, scores.strides[0], rule,
                                 C.byref(r), C.byref(col))
    return None if rc != 0 else (r.value, col.value)


def max_(scores, rule=RULE_GENERIC):
    rows, c = scores.shape
    if scores.dtype == np.float32:
        out = C.c_float(0)
        rc = lib().lmo_max_f32(_p(scores, _f32p), rows, c, scores.strides[0] // 4, rule, C.byref(out))
    else:
        out = C.c_uint8(0)
        rc = lib().lmo_max_u8(_p(scores, _u8p), rows, c, scores.strides[0], C.byref(out))
    return None if rc != 0 else out.value


def threshold(scores, t):
    """Threshold::threshold -> (rows[u64], cols[u32]) in row-major order."""
    rows, c = scores.shape
    cap = rows * c
    ro = np.empty(cap, dtype=np.uint64)
    co = np.empty(cap, dtype=np.uint32)
    if scores.dtype == np.float32:
        n = lib().lmo_threshold_f32(_p(scores, _f32p), rows, c, scores.strides[0] // 4, t,
                                    _p(ro, _u64p), _p(co, _u32p), cap)
    else:
        n = lib().lmo_threshold_u8(_p(scores, _u8p), rows, c, scores.strides[0], int(t),
                                   _p(ro, _u64p), _p(co, _u32p), cap)
    return ro[:n].copy(), co[:n].copy()


def offset(row, col, rows):
    """StripedScores::offset (scores.rs:155-157)."""
    return col * rows + row


def unstripe(scores, max_index):
    """StripedScores::unstripe (scores.rs:167,272-279)."""
    rows, c = scores.shape
    end = min(max_index, rows * c)
    return np.ascontiguousarray(scores.T).reshape(-1)[:end].copy()


# ---------------------------------------------------------------------- pssm

def counts_from_sequences(alphabet, patterns):
    enc = [encode(alphabet, p) for p in patterns]
    m = enc[0].size
    k = alphabet_k(alphabet)
    ptrs = (_u8p * len(enc))(*[_p(e, _u8p) for e in enc])
    counts = np.zeros((m, k), dtype=np.uint32)
    lib().lmo_counts_from_sequences(alphabet, ptrs, len(enc), m, _p(counts, _u32p))
    return counts


def to_freq(alphabet, counts, pseudo):
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    freq = np.empty(counts.shape, dtype=np.float32)
    lib().lmo_to_freq(alphabet, _p(counts, _u32p), counts.shape[0], pseudo, _p(freq, _f32p))
    return freq


def to_scoring(alphabet, freq, bg=None):
    freq = np.ascontiguousarray(freq, dtype=np.float32)
    pssm = np.empty(freq.shape, dtype=np.float32)
    bgp = None
    if bg is not None:
        bg = np.ascontiguousarray(bg, dtype=np.float32)
        bgp = _p(bg, _f32p)
    lib().lmo_to_scoring(alphabet, _p(freq, _f32p), freq.shape[0], bgp, _p(pssm, _f32p))
    return pssm


def pssm_from_counts(alphabet, counts, pseudo=0.1, bg=None):
    """CountMatrix::to_freq(pseudo).to_scoring(bg)."""
    return to_scoring(alphabet, to_freq(alphabet, counts, pseudo), bg)


def pssm_from_sequences(alphabet, patterns, pseudo=0.1, bg=None):
    return pssm_from_counts(alphabet, counts_from_sequences(alphabet, patterns), pseudo, bg)


def max_score(alphabet, pssm):
    pssm = np.ascontiguousarray(pssm, dtype=np.float32)
    return lib().lmo_max_score(alphabet, _p(pssm, _f32p), pssm.shape[0])


def min_score(alphabet, pssm):
    pssm = np.ascontiguousarray(pssm, dtype=np.float32)
    return lib().lmo_min_score(alphabet, _p(pssm, _f32p), pssm.shape[0])


class Discrete:
    """DiscreteMatrix (pwm/mod.rs:748-790)."""

    def __init__(self, alphabet, pssm):
        pssm = np.ascontiguousarray(pssm, dtype=np.float32)
        m, k = pssm.shape
        self.data = np.empty((m, k), dtype=np.uint8)
        self.offsets = np.empty(m, dtype=np.float32)
        f, o = C.c_float(0), C.c_float(0)
        lib().lmo_to_discrete(alphabet, _p(pssm, _f32p), m, _p(self.data, _u8p),
                              C.byref(f), _p(self.offsets, _f32p), C.byref(o))
        self.factor, self.offset = f.value, o.value

    def scale(self, score):
        return lib().lmo_scale(score, self.factor, self.offset)

    def unscale(self, score):
        return lib().lmo_unscale(int(score), self.factor, self.offset)


def score_position(pssm, striped, rows, pos):
    pssm = np.ascontiguousarray(pssm, dtype=np.float32)
    m, k = pssm.shape
    return lib().lmo_score_position(_p(pssm, _f32p), m, k, _p(striped, _u8p),
                                    striped.shape[1], rows, pos)


def scanner_collect(alphabet, pssm, striped, wrap, length, threshold_, block_size=256):
    pssm = np.ascontiguousarray(pssm, dtype=np.float32)
    m, _ = pssm.shape
    c = striped.shape[1]
    rows = striped.shape[0] - wrap
    cap = 1 << 20
    pos = np.empty(cap, dtype=np.uint64)
    sc = np.empty(cap, dtype=np.float32)
    n = lib().lmo_scanner_collect(alphabet, _p(pssm, _f32p), m, _p(striped, _u8p), c, c, rows,
                                  wrap, length, threshold_, block_size, _p(pos, _u64p),
                                  _p(sc, _f32p), cap)
    assert n <= cap
    return pos[:n].copy(), sc[:n].copy()


def scanner_max(alphabet, pssm, striped, wrap, length, threshold_=0.0, block_size=256):
    pssm = np.ascontiguousarray(pssm, dtype=np.float32)
    m, _ = pssm.shape
    c = striped.shape[1]
    rows = striped.shape[0] - wrap
    pos, sc = C.c_uint64(0), C.c_float(0)
    rc = lib().lmo_scanner_max(alphabet, _p(pssm, _f32p), m, _p(striped, _u8p), c, c, rows, wrap,
                               length, threshold_, block_size, C.byref(pos), C.byref(sc))
    return None if rc != 0 else (pos.value, sc.value)


# ----------------------------------------------------------------- synthetic

SEED = 0x5EED


def synth_dna(n, offset=0, seed=SEED):
    dst = np.empty(n, dtype=np.uint8)
    lib().lmo_synth_dna(seed, offset, n, _p(dst, _u8p))
    return dst


def synth_protein(n, offset=0, seed=SEED):
    dst = np.empty(n, dtype=np.uint8)
    lib().lmo_synth_protein(seed, offset, n, _p(dst, _u8p))
    return dst


# ------------------------------------------------------- AVX2 CPU baseline

def pad_pssm_dna(pssm):
    """DenseMatrix<f32, U5> layout: rows padded to 8 floats (dense.rs:125-128)."""
    m, k = pssm.shape
    out = aligned_empty((m, 8), np.float32)
    out[:] = 0
    out[:, :k] = pssm
    return out


def avx2_score_f32_dna(pssm, striped, row_begin, row_end):
    p8 = pad_pssm_dna(np.asarray(pssm, dtype=np.float32))
    assert striped.shape[1] == 32 and striped.ctypes.data % 32 == 0
    out = aligned_empty((row_end - row_begin, 32), np.float32)
    lib().lmo_avx2_score_f32_dna(_p(p8, _f32p), p8.shape[0], _p(striped, _u8p), row_begin,
                                 row_end, _p(out, _f32p))
    return out


def avx2_argmax_f32(scores):
    r, c = C.c_size_t(0), C.c_size_t(0)
    rc = lib().lmo_avx2_argmax_f32(_p(scores, _f32p), scores.shape[0], C.byref(r), C.byref(c))
    return None if rc != 0 else (r.value, c.value)


def avx2_score_argmax_mt(pssm, striped, rows, threads, scratch=None):
    p8 = pad_pssm_dna(np.asarray(pssm, dtype=np.float32))
    if scratch is None:
        scratch = aligned_empty((rows, 32), np.float32)
    best, r, c = C.c_float(0), C.c_size_t(0), C.c_size_t(0)
    rc = lib().lmo_avx2_score_argmax_mt(_p(p8, _f32p), p8.shape[0], _p(striped, _u8p), rows,
                                        threads, _p(scratch, _f32p), C.byref(best),
                                        C.byref(r), C.byref(c))
    return None if rc != 0 else (best.value, r.value, c.value)
