"""lightmotif_b200 -- B200-native PWM scanning engine (drop-in for the
``lightmotif::pli`` Score pipeline of althonos/lightmotif).

The product is the CUDA library behind ``include/lightmotif_b200.h``; this
package is the thin Python harness over that C ABI used by the tests and by
``bench.py``.  The C++ host mirror of the reference's trait surface lives in
``include/lightmotif_b200.hpp``.

Nothing here computes scores on the CPU: every method forwards to the C ABI
and raises if the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import (DNA, PROTEIN, RULE_AVX2, RULE_GENERIC, RULE_SSE2, STRAND_DIRECT, STRAND_REVERSE, STRAND_BOTH,
                   LANES, ColWin, Hit, LightmotifError)

__all__ = ["Pipeline", "DeviceSequence", "Comm", "InvalidSymbol", "UnsupportedBackend", "MatrixCoordinates",
           "DNA", "PROTEIN", "RULE_GENERIC", "RULE_AVX2", "RULE_SSE2", "STRAND_DIRECT", "STRAND_REVERSE", "STRAND_BOTH",
           "reverse_complement", "LANES", "alphabet_k",
           "pad_pssm", "aligned_empty", "argmax_merge", "COLWIN_DTYPE", "HIT_DTYPE"]


class InvalidSymbol(ValueError):
    """lightmotif::err::InvalidSymbol(char), err.rs:10."""

    def __init__(self, ch, index):
        super().__init__(f"invalid symbol {ch!r} at index {index}")
        self.ch, self.index = ch, index


class UnsupportedBackend(RuntimeError):
    """lightmotif::err::UnsupportedBackend, err.rs:34."""


@dataclass(frozen=True)
class MatrixCoordinates:
    """dense.rs:28-39"""
    row: int
    col: int


def alphabet_k(alphabet):
    return 5 if alphabet == DNA else 21


def aligned_empty(shape, dtype, align=64):
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    raw = np.empty(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


def pad_pssm(pssm, alphabet, dtype=np.float32):
    """Lay a [m, K] matrix out like DenseMatrix<T, K> (dense.rs:125-128): rows
    padded to 32 bytes (stride 8 / 24 floats, 32 bytes for u8)."""
    pssm = np.asarray(pssm, dtype=dtype)
    m, k = pssm.shape
    row_bytes = -(-(k * pssm.itemsize) // 32) * 32
    stride = row_bytes // pssm.itemsize
    out = np.zeros((m, stride), dtype=dtype)
    out[:, :k] = pssm
    return out, stride


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _hit(h):
    return {"row": int(h.row), "col": int(h.col), "position": int(h.position), "score": float(np.float32(h.score))}


class DeviceSequence:
    """A striped sequence resident in HBM (lm_seq handle)."""

    def __init__(self, pli, handle):
        self._pli, self._h = pli, handle
        a, r, w, n = C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        capi.lib().lm_seq_info(handle, C.byref(a), C.byref(r), C.byref(w), C.byref(n))
        self.alphabet, self.rows, self.wrap, self.length = a.value, r.value, w.value, n.value

    def _refresh(self):
        self.__init__(self._pli, self._h)

    def configure_wrap(self, wrap):
        self._pli._check(capi.lib().lm_seq_configure_wrap(self._pli._ctx, self._h, wrap))
        self._refresh()

    def configure(self, pssm):
        """StripedSequence::configure (seq.rs:362-366)."""
        if len(pssm):
            self.configure_wrap(len(pssm) - 1)

    def download(self):
        out = aligned_empty((self.rows + self.wrap, LANES), np.uint8)
        self._pli._check(capi.lib().lm_seq_download(self._pli._ctx, self._h, _ptr(out), LANES))
        return out

    @property
    def device_ptr(self):
        return capi.lib().lm_seq_device_ptr(self._h)

    def free(self):
        if self._h:
            capi.lib().lm_seq_free(self._pli._ctx, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- scoring ------------------------------------------------------
    def _rng(self, rows):
        return (0, self.rows) if rows is None else (rows[0], rows[1])

    def score(self, pssm, rows=None, out=None):
        """Score::score_rows_into -> [n, 32] f32 (or u8 for a u8 matrix)."""
        pssm = np.asarray(pssm)
        is_u8 = pssm.dtype == np.uint8
        p, stride = pad_pssm(pssm, self.alphabet, np.uint8 if is_u8 else np.float32)
        b, e = self._rng(rows)
        n = max(e - b, 0)
        if out is None:
            out = aligned_empty((n, LANES), np.uint8 if is_u8 else np.float32)
        got = C.c_size_t(0)
        fn = capi.lib().lm_seq_score_u8 if is_u8 else capi.lib().lm_seq_score_f32
        self._pli._check(fn(self._pli._ctx, self._h, _ptr(p), p.shape[0], stride, b, e, _ptr(out), LANES, 0,
                            C.byref(got)))
        return out[:got.value]

    def score_into_device(self, pssm, device_ptr, out_stride=LANES, rows=None):
        pssm = np.asarray(pssm)
        is_u8 = pssm.dtype == np.uint8
        p, stride = pad_pssm(pssm, self.alphabet, np.uint8 if is_u8 else np.float32)
        b, e = self._rng(rows)
        got = C.c_size_t(0)
        fn = capi.lib().lm_seq_score_u8 if is_u8 else capi.lib().lm_seq_score_f32
        self._pli._check(fn(self._pli._ctx, self._h, _ptr(p), p.shape[0], stride, b, e, C.c_void_p(device_ptr),
                            out_stride, 1, C.byref(got)))
        return got.value

    def score_argmax(self, pssm, rows=None, rule=RULE_AVX2, with_columns=False):
        """Fused ScoringMatrix::score + StripedScores::argmax."""
        p, stride = pad_pssm(pssm, self.alphabet)
        b, e = self._rng(rows)
        best = Hit()
        cols = (ColWin * LANES)()
        fc = C.c_float(0)
        st = capi.lib().lm_seq_score_argmax_f32(self._pli._ctx, self._h, _ptr(p), p.shape[0], stride, b, e, rule,
                                                C.byref(best), cols, C.byref(fc))
        self._pli._check(st)
        if st == capi.EMPTY:
            return None
        res = _hit(best)
        if with_columns == "numpy":  # zero-copy view of the lm_colwin[32] array (fast path for dist)
            res["columns"] = np.frombuffer(cols, dtype=COLWIN_DTYPE).copy()
            res["first_cell"] = float(np.float32(fc.value))
        elif with_columns:
            res["columns"] = [(int(c.row), float(np.float32(c.score)), int(c.valid)) for c in cols]
            res["first_cell"] = float(np.float32(fc.value))
        return res

    def score_argmax_launch(self, pssm, slot=0, rows=None):
        """Enqueue a fused argmax scan (lm_seq_score_argmax_f32_launch); pair with score_argmax_finish."""
        p, stride = pad_pssm(pssm, self.alphabet)
        b, e = self._rng(rows)
        self._pli._check(capi.lib().lm_seq_score_argmax_f32_launch(self._pli._ctx, self._h, _ptr(p), p.shape[0],
                                                                   stride, b, e, slot))

    def score_argmax_finish(self, slot=0, rule=RULE_AVX2, with_columns=False):
        best = Hit()
        cols = (ColWin * LANES)()
        fc = C.c_float(0)
        st = capi.lib().lm_seq_score_argmax_f32_finish(self._pli._ctx, slot, rule, C.byref(best), cols, C.byref(fc))
        self._pli._check(st)
        if st == capi.EMPTY:
            return None
        res = _hit(best)
        if with_columns == "numpy":
            res["columns"] = np.frombuffer(cols, dtype=COLWIN_DTYPE).copy()
            res["first_cell"] = float(np.float32(fc.value))
        elif with_columns:
            res["columns"] = [(int(c.row), float(np.float32(c.score)), int(c.valid)) for c in cols]
            res["first_cell"] = float(np.float32(fc.value))
        return res

    def score_threshold(self, pssm, t, rows=None, cap=1 << 20, as_array=False):
        """Fused score + Threshold::threshold -> list of hits in row-major order (as_array: a numpy
        array of HIT_DTYPE records instead of a list of dicts)."""
        p, stride = pad_pssm(pssm, self.alphabet)
        b, e = self._rng(rows)
        hits = np.zeros(cap, dtype=HIT_DTYPE)
        count = C.c_size_t(0)
        st = capi.lib().lm_seq_score_threshold_f32(self._pli._ctx, self._h, _ptr(p), p.shape[0], stride, b, e,
                                                   float(t), hits.ctypes.data_as(C.POINTER(Hit)), cap, C.byref(count))
        if st == capi.ERR_OVERFLOW:
            return self.score_threshold(pssm, t, rows, cap=count.value, as_array=as_array)
        self._pli._check(st)
        hits = hits[:count.value]
        if as_array:
            return hits
        return [{"row": int(h["row"]), "col": int(h["col"]), "position": int(h["position"]),
                 "score": float(h["score"])} for h in hits]

    def _flatten(self, pssms):
        k = alphabet_k(self.alphabet)
        stride = pad_pssm(np.zeros((1, k), np.float32), self.alphabet)[1]
        widths = np.array([len(p) for p in pssms], dtype=np.uint32)
        offs = np.zeros(len(pssms), dtype=np.uint64)
        flat = np.zeros((int(widths.sum()), stride), dtype=np.float32)
        at = 0
        for i, p in enumerate(pssms):
            offs[i] = at * stride
            flat[at:at + len(p), :k] = np.asarray(p, dtype=np.float32)
            at += len(p)
        return flat, widths, offs, stride

    def score_threshold_batch(self, pssms, thresholds, cap=4096, counts_only=False, strands=capi.STRAND_DIRECT):
        """Many motifs, one threshold each (lm_seq_score_threshold_f32_batch_rc) -> list of hit lists
        (or, with counts_only, just the number of hits per entry).  strands = STRAND_BOTH scans every
        motif and its reverse complement: entries 2i (direct) and 2i + 1 (reverse complement)."""
        flat, widths, offs, stride = self._flatten(pssms)
        thr = np.ascontiguousarray(thresholds, dtype=np.float32)
        assert thr.size == len(pssms)
        n = len(pssms) * (2 if strands == capi.STRAND_BOTH else 1)
        hits = np.zeros(n * cap, dtype=HIT_DTYPE)
        counts = (C.c_size_t * n)()
        st = capi.lib().lm_seq_score_threshold_f32_batch_rc(self._pli._ctx, self._h, _ptr(flat), _ptr(widths), _ptr(offs),
                                                            len(pssms), stride, strands, _ptr(thr),
                                                            hits.ctypes.data_as(C.POINTER(Hit)), cap, counts)
        if counts_only and st in (capi.OK, capi.ERR_OVERFLOW):
            return [int(c) for c in counts]
        if st == capi.ERR_OVERFLOW:
            return self.score_threshold_batch(pssms, thresholds, cap=max(counts), strands=strands)
        self._pli._check(st)
        return [[{"row": int(h["row"]), "col": int(h["col"]), "position": int(h["position"]), "score": float(h["score"])}
                 for h in hits[q * cap:q * cap + counts[q]]] for q in range(n)]

    def score_argmax_batch(self, pssms, rule=RULE_AVX2, strands=capi.STRAND_DIRECT):
        """Many motifs over this sequence (the lightmotif-cli job loop) -> best hit per entry (None for
        a motif longer than the sequence)."""
        flat, widths, offs, stride = self._flatten(pssms)
        n = len(pssms) * (2 if strands == capi.STRAND_BOTH else 1)
        best = (Hit * n)()
        status = (C.c_int * n)()
        self._pli._check(capi.lib().lm_seq_score_argmax_f32_batch_rc(
            self._pli._ctx, self._h, _ptr(flat), _ptr(widths), _ptr(offs), len(pssms), stride, strands, rule, best, status))
        return [_hit(b) if status[i] == capi.OK else None for i, b in enumerate(best)]


def reverse_complement(pssm):
    """lm_pssm_reverse_complement: ScoringMatrix::reverse_complement (pwm/mod.rs:566-577), DNA."""
    p, stride = pad_pssm(pssm, DNA)
    out = np.zeros_like(p)
    st = capi.lib().lm_pssm_reverse_complement(_ptr(p), p.shape[0], stride, _ptr(out))
    if st < 0:
        raise LightmotifError(st)
    return out[:, :5].copy()


class Scanner:
    """Mirror of ``lightmotif::scan::Scanner`` (scan.rs:96-250) on a resident sequence.

    The reference scans blocks of rows with the u8 DiscreteMatrix, thresholds the u8 scores and
    rescores the survivors exactly in f32; here the same filter + exact recheck runs inside ONE fused
    kernel over the whole sequence (prune.cu), so `block_size` is accepted and ignored.  Hits are the
    positions p with p + M <= len(seq) whose exact f32 score is >= threshold (scan.rs:186-191), each
    with that exact score.  `max()` returns the best hit, ties resolved towards the larger position
    (scan.rs:235); it searches every position exactly, i.e. it does not reproduce the reference's
    dependence on the u8 rounding of near-ties."""

    def __init__(self, pssm, seq):
        self.pssm, self.seq, self._threshold, self._block_size = np.asarray(pssm, dtype=np.float32), seq, 0.0, 256

    def threshold(self, t):
        self._threshold = float(t)
        return self

    def block_size(self, b):
        self._block_size = int(b)
        return self

    def collect(self):
        m = len(self.pssm)
        hits = self.seq.score_threshold(self.pssm, self._threshold)
        out = [{"position": h["position"], "score": h["score"]} for h in hits if h["position"] + m <= self.seq.length]
        out.sort(key=lambda h: h["position"])
        return out

    def __iter__(self):
        return iter(self.collect())

    def max(self):
        # the SSE2 rule keeps the LAST column among equal maxima and the last row within it, i.e.
        # the largest position col * rows + row -- the Scanner's tie-break
        best = self.seq.score_argmax(self.pssm, rule=RULE_SSE2)
        if best is None or not (best["score"] >= self._threshold):
            return None
        return {"position": best["position"], "score": best["score"]}


class Pipeline:
    """``Pipeline::<A, Cuda>::cuda()``: the CUDA back-end of the pipeline traits
    (Encode / Stripe / Score / Maximum / Threshold, pli/mod.rs:34-222)."""

    def __init__(self, alphabet=DNA, device=0, stream=None):
        self.alphabet = alphabet
        ctx = C.c_void_p()
        if stream is None:
            st = capi.lib().lm_ctx_create(device, C.byref(ctx))
        else:
            st = capi.lib().lm_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(ctx))
        if st == capi.ERR_UNSUPPORTED:
            raise UnsupportedBackend("no sm_100 CUDA device")
        if st != capi.OK:
            raise LightmotifError(st)
        self._ctx = ctx

    @classmethod
    def cuda(cls, alphabet=DNA, device=0, stream=None):
        return cls(alphabet, device, stream)

    @classmethod
    def _borrow(cls, ctx, alphabet):
        """Wrap a context owned by someone else (the ranks of an lm_comm)."""
        self = cls.__new__(cls)
        self.alphabet, self._ctx, self._borrowed = alphabet, C.c_void_p(ctx), True
        return self

    def close(self):
        if getattr(self, "_ctx", None):
            if not getattr(self, "_borrowed", False):
                capi.lib().lm_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        if status < 0:
            raise LightmotifError(status, capi.lib().lm_ctx_last_error(self._ctx).decode())
        return status

    @property
    def kernel_launches(self):
        return capi.lib().lm_ctx_kernel_launches(self._ctx)

    @property
    def last_kernel_ms(self):
        return capi.lib().lm_ctx_last_kernel_ms(self._ctx)

    @property
    def last_sm_clock_mhz(self):
        """SM clock the last filtered scan ran at, measured inside the kernel (lm_ctx_last_sm_clock_mhz)."""
        return float(capi.lib().lm_ctx_last_sm_clock_mhz(self._ctx))

    @property
    def stream(self):
        return capi.lib().lm_ctx_stream(self._ctx)

    def synchronize(self):
        self._check(capi.lib().lm_ctx_synchronize(self._ctx))

    def result_block_ptr(self, slot=0):
        """Device address of the 272-byte result block of launch slot `slot` (lm_ctx_result_block)."""
        return int(capi.lib().lm_ctx_result_block(self._ctx, int(slot)))

    def set_prune(self, mode):
        """LM_OPT_PRUNE: 0 = score every cell, 1 = auto (default), 2 = filter whenever possible."""
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 1, int(mode)))

    def set_filter_rows(self, r):
        """LM_OPT_FILTER_ROWS: force the filter depth (0 = planner decides)."""
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 3, int(r)))

    def set_filter_mode(self, mode):
        """LM_OPT_FILTER_MODE: 0 = planner decides, 1 = single-row look-ups, 2 = paired look-ups."""
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 5, int(mode)))

    def set_class_cache(self, on):
        """LM_OPT_CLASS_CACHE: reuse (1) or rebuild (0) the class-coded copy of a protein sequence."""
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 6, int(on)))

    def set_reserve_sms(self, n):
        """LM_OPT_RESERVE_SMS: SMs left free for concurrent kernels (NCCL)."""
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 4, int(n)))

    def set_boot_rows(self, rows):
        """LM_OPT_BOOT_ROWS: rows scored exactly before the filtered argmax scan."""
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 2, int(rows)))

    def set_upload(self, mode=1, threads=0, min_bytes=None):
        """LM_OPT_UPLOAD (0 plain copy, 1 adaptive: times both ways and keeps the faster, 2 packed chunks only,
        3 / 4 two-ended with plain chunks only in link gaps / always), LM_OPT_UPLOAD_THREADS
        (0 = all hardware threads but one), LM_OPT_UPLOAD_MIN_BYTES: how sequence_from_striped moves a large
        DNA matrix to the device."""
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 7, int(mode)))
        self._check(capi.lib().lm_ctx_set_option(self._ctx, 8, int(threads)))
        if min_bytes is not None:
            self._check(capi.lib().lm_ctx_set_option(self._ctx, 9, int(min_bytes)))

    @property
    def last_upload(self):
        out = (C.c_uint64 * 4)()
        self._check(capi.lib().lm_ctx_last_upload(self._ctx, C.byref(out)))
        return {"pcie_bytes": int(out[0]), "packed_chunks": int(out[1]), "plain_chunks": int(out[2]), "threads": int(out[3])}

    @property
    def last_stats(self):
        out = (C.c_uint64 * 4)()
        self._check(capi.lib().lm_ctx_last_stats(self._ctx, C.byref(out)))
        return {"rechecked": int(out[0]), "overflow_tiles": int(out[1]), "pruned": int(out[2]),  # 1 if the filter ran (batch calls: number of motifs that ran it)
                "filter_rows": int(out[3]) & 0xff, "paired": bool(int(out[3]) >> 8)}

    # ---- Encode ---------------------------------------------------------
    @staticmethod
    def _bytes(seq):
        if isinstance(seq, str):
            seq = seq.encode("ascii")
        if isinstance(seq, (bytes, bytearray)):
            return np.frombuffer(bytes(seq), dtype=np.uint8)
        return np.ascontiguousarray(seq, dtype=np.uint8)

    def encode(self, seq):
        a = self._bytes(seq)
        dst = np.empty(a.size, dtype=np.uint8)
        bad = C.c_size_t(0)
        st = capi.lib().lm_encode(self._ctx, self.alphabet, _ptr(a), a.size, _ptr(dst), C.byref(bad))
        if st == capi.ERR_INVALID_SYMBOL:
            raise InvalidSymbol(chr(int(a[bad.value])), bad.value)
        self._check(st)
        return dst

    def encode_lossy(self, seq):
        a = self._bytes(seq)
        dst = np.empty(a.size, dtype=np.uint8)
        self._check(capi.lib().lm_encode_lossy(self._ctx, self.alphabet, _ptr(a), a.size, _ptr(dst)))
        return dst

    # ---- Stripe ---------------------------------------------------------
    def stripe(self, sym, wrap=0):
        """Stripe::stripe (+ configure_wrap) on host buffers -> [rows + wrap, 32]."""
        sym = np.ascontiguousarray(sym, dtype=np.uint8)
        rows = -(-sym.size // LANES)
        data = aligned_empty((rows + wrap, LANES), np.uint8)
        data[rows:] = 255
        self._check(capi.lib().lm_stripe(self._ctx, self.alphabet, _ptr(sym), sym.size, _ptr(data), LANES))
        if wrap:
            self._check(capi.lib().lm_configure_wrap(self._ctx, self.alphabet, _ptr(data), rows, LANES, wrap))
        return data

    # ---- Score (host buffers) --------------------------------------------
    def score_rows(self, pssm, striped, length, wrap, rows=None):
        pssm = np.asarray(pssm)
        is_u8 = pssm.dtype == np.uint8
        p, stride = pad_pssm(pssm, self.alphabet, np.uint8 if is_u8 else np.float32)
        nrows = striped.shape[0] - wrap
        b, e = (0, nrows) if rows is None else rows
        out = aligned_empty((max(e - b, 0), LANES), np.uint8 if is_u8 else np.float32)
        got = C.c_size_t(0)
        fn = capi.lib().lm_score_u8_rows if is_u8 else capi.lib().lm_score_f32_rows
        self._check(fn(self._ctx, self.alphabet, _ptr(p), p.shape[0], stride, _ptr(striped), nrows, wrap,
                       striped.strides[0], length, b, e, _ptr(out), LANES, C.byref(got)))
        return out[:got.value]

    # ---- Maximum / Threshold (host score matrices) -------------------------
    def argmax(self, scores, rule=RULE_AVX2):
        best = Hit()
        is_u8 = scores.dtype == np.uint8
        fn = capi.lib().lm_argmax_u8 if is_u8 else capi.lib().lm_argmax_f32
        st = self._check(fn(self._ctx, _ptr(scores), scores.shape[0], scores.strides[0] // scores.itemsize, rule,
                            C.byref(best)))
        return None if st == capi.EMPTY else MatrixCoordinates(int(best.row), int(best.col))

    def max(self, scores, rule=RULE_GENERIC):
        """Maximum::max; rule=RULE_AVX2 reproduces that back-end's floor at 0.0 (lm_max_f32_rule)."""
        is_u8 = scores.dtype == np.uint8
        out = C.c_uint8(0) if is_u8 else C.c_float(0)
        if is_u8:
            st = self._check(capi.lib().lm_max_u8(self._ctx, _ptr(scores), scores.shape[0], scores.strides[0], C.byref(out)))
        else:
            st = self._check(capi.lib().lm_max_f32_rule(self._ctx, _ptr(scores), scores.shape[0], scores.strides[0] // 4,
                                                        rule, C.byref(out)))
        return None if st == capi.EMPTY else (int(out.value) if is_u8 else np.float32(out.value))

    def threshold(self, scores, t, cap=None):
        is_u8 = scores.dtype == np.uint8
        cap = cap or max(scores.size, 1)
        hits = (Hit * cap)()
        count = C.c_size_t(0)
        if is_u8:
            st = capi.lib().lm_threshold_u8(self._ctx, _ptr(scores), scores.shape[0], scores.strides[0], int(t),
                                            hits, cap, C.byref(count))
        else:
            st = capi.lib().lm_threshold_f32(self._ctx, _ptr(scores), scores.shape[0], scores.strides[0] // 4,
                                             float(t), hits, cap, C.byref(count))
        self._check(st)
        return [MatrixCoordinates(int(hits[i].row), int(hits[i].col)) for i in range(count.value)]

    # ---- device-resident sequences ----------------------------------------
    def _seq(self, handle):
        return DeviceSequence(self, handle)

    def sequence_from_ascii(self, seq, wrap=0, lossy=False):
        a = self._bytes(seq)
        h = C.c_void_p()
        bad = C.c_size_t(0)
        st = capi.lib().lm_seq_from_ascii(self._ctx, self.alphabet, _ptr(a), a.size, int(lossy), wrap, C.byref(h),
                                          C.byref(bad))
        if st == capi.ERR_INVALID_SYMBOL:
            raise InvalidSymbol(chr(int(a[bad.value])), bad.value)
        self._check(st)
        return self._seq(h)

    def sequence_from_symbols(self, sym, wrap=0):
        sym = np.ascontiguousarray(sym, dtype=np.uint8)
        h = C.c_void_p()
        self._check(capi.lib().lm_seq_from_symbols(self._ctx, self.alphabet, _ptr(sym), sym.size, wrap, C.byref(h)))
        return self._seq(h)

    def sequence_from_striped(self, data, length, wrap):
        h = C.c_void_p()
        rows = data.shape[0] - wrap
        self._check(capi.lib().lm_seq_from_striped(self._ctx, self.alphabet, _ptr(data), rows, wrap,
                                                   data.strides[0], length, C.byref(h)))
        return self._seq(h)

    def sequence_from_striped_slab(self, data, length, halo, row_offset):
        """Row slab [row_offset, row_offset + len(data) - halo) of the striped matrix of a
        `length`-symbol sequence, followed by `halo` rows (lm_seq_from_striped_slab)."""
        h = C.c_void_p()
        rows = data.shape[0] - halo
        self._check(capi.lib().lm_seq_from_striped_slab(self._ctx, self.alphabet, _ptr(data), rows, halo,
                                                        data.strides[0], length, row_offset, -(-length // LANES),
                                                        C.byref(h)))
        return self._seq(h)

    def sequence_from_device(self, device_ptr, rows, wrap, length):
        h = C.c_void_p()
        self._check(capi.lib().lm_seq_from_device(self._ctx, self.alphabet, C.c_void_p(device_ptr), rows, wrap,
                                                  length, C.byref(h)))
        return self._seq(h)

    def sequence_synthetic(self, total_length, wrap, seed=0x5EED, row_begin=0, n_rows=None):
        rows_global = -(-total_length // LANES)
        if n_rows is None:
            n_rows = rows_global - row_begin
        h = C.c_void_p()
        self._check(capi.lib().lm_seq_synthetic(self._ctx, self.alphabet, seed, total_length, row_begin, n_rows,
                                                wrap, C.byref(h)))
        return self._seq(h)


COLWIN_DTYPE = np.dtype([("row", np.uint64), ("score", np.float32), ("valid", np.uint32)])  # == lm_colwin
HIT_DTYPE = np.dtype([("row", np.uint64), ("position", np.uint64), ("col", np.uint32), ("score", np.float32)])  # == lm_hit


def f32_key(x):
    """Order-preserving u32 key of f32 values (common.cuh f32_to_key), vectorised."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def plan_filter(pssm, threshold=None, n_positions=5e9, lookups=0, mode=0):
    """lm_plan_filter: the filter the fused DNA calls would use (host only).  threshold=None plans
    for an argmax scan over n_positions.  Returns a dict, or None when no filter exists."""
    p, stride = pad_pssm(pssm, DNA)
    out = capi.FilterPlan()
    st = capi.lib().lm_plan_filter(_ptr(p), p.shape[0], stride, 0.0 if threshold is None else float(threshold),
                                   1 if threshold is None else 0, float(n_positions), int(lookups), int(mode), C.byref(out))
    if st != capi.OK:
        return None
    return {"ok": bool(out.ok), "r": out.r, "pair": bool(out.pair), "pair_mul": out.pair_mul, "j0": out.j0,
            "smax": out.smax, "tab_lo": list(out.tab_lo), "tab_hi": list(out.tab_hi), "c0": out.c0,
            "delta": out.delta, "eps": out.eps, "est_pass_rate": out.est_pass_rate, "est_cost": out.est_cost}


def plan_filter_protein(pssm, threshold=None, n_positions=5e8, lookups=0):
    """lm_plan_filter_protein: class partition + look-up tables the fused protein calls would use."""
    p, stride = pad_pssm(pssm, PROTEIN)
    out = capi.FilterPlan()
    cls = np.zeros(21, dtype=np.uint8)
    st = capi.lib().lm_plan_filter_protein(_ptr(p), p.shape[0], stride, 0.0 if threshold is None else float(threshold),
                                           1 if threshold is None else 0, float(n_positions), int(lookups), C.byref(out),
                                           _ptr(cls))
    if st != capi.OK:
        return None
    return {"ok": bool(out.ok), "r": out.r, "pair": False, "j0": out.j0, "smax": out.smax, "tab_lo": list(out.tab_lo),
            "tab_hi": list(out.tab_hi), "c0": out.c0, "delta": out.delta, "eps": out.eps,
            "est_pass_rate": out.est_pass_rate, "est_cost": out.est_cost, "class_of": cls}


RESULT_BLOCK_WORDS = 34  # uint64 words of a result block: 32 column keys, first_cell, padding


def decode_keys(keys):
    """uint64[32] column keys of a result block -> numpy array of COLWIN_DTYPE (lm_decode_keys)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(LANES, dtype=COLWIN_DTYPE)
    capi.lib().lm_decode_keys(_ptr(keys), out.ctypes.data_as(C.POINTER(ColWin)))
    return out


def argmax_merge(parts, row_offsets, first_cell, total_rows, rule=RULE_AVX2):
    """lm_argmax_merge.  parts: [n_slabs, 32] array of COLWIN_DTYPE, or a list
    (one per slab) of 32 (row, score, valid) tuples."""
    if isinstance(parts, np.ndarray) and parts.dtype == COLWIN_DTYPE:
        arr = np.ascontiguousarray(parts)
    else:
        arr = np.zeros((len(parts), LANES), dtype=COLWIN_DTYPE)
        for p, cols in enumerate(parts):
            for c, (row, score, valid) in enumerate(cols):
                arr[p, c] = (int(row), float(score), int(valid))
    offs = np.ascontiguousarray(row_offsets, dtype=np.uint64)
    best = Hit()
    st = capi.lib().lm_argmax_merge(arr.ctypes.data_as(C.POINTER(ColWin)), arr.shape[0], _ptr(offs),
                                    float(first_cell), int(total_rows), rule, C.byref(best))
    if st < 0:
        raise LightmotifError(st)
    return None if st == capi.EMPTY else _hit(best)


class Comm:
    """``lm_comm``: row-sharded fused scans across the GPUs of one box (include/lightmotif_b200.h).

    ``Comm.local(devices)`` drives N GPUs from this process; ``Comm.from_process_group(pli)`` joins
    one-process-per-GPU ranks (torchrun): the cudaIpc handles of the exchange buffers travel through
    ``torch.distributed`` once, after that the GPUs talk through peer-mapped HBM only."""

    def __init__(self, handle, alphabet):
        self._h, self.alphabet = handle, alphabet
        L = capi.lib()
        self.world, self.n_local = L.lm_comm_world(handle), L.lm_comm_n_local(handle)
        self.ranks = [L.lm_comm_rank(handle, i) for i in range(self.n_local)]
        self.pipelines = [Pipeline._borrow(L.lm_comm_ctx(handle, i), alphabet) for i in range(self.n_local)]

    @classmethod
    def local(cls, devices, alphabet=DNA, hit_capacity=0):
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        st = capi.lib().lm_comm_init(len(devices), devs, hit_capacity, C.byref(h))
        if st == capi.ERR_UNSUPPORTED:
            raise UnsupportedBackend("the devices cannot address each other's memory (or no sm_100 device)")
        if st != capi.OK:
            raise LightmotifError(st)
        return cls(h, alphabet)

    @classmethod
    def from_process_group(cls, pli, group=None, hit_capacity=0):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        mine = np.zeros(capi.COMM_HANDLE_BYTES, dtype=np.uint8)
        h = C.c_void_p()
        pli._check(capi.lib().lm_comm_init_rank(pli._ctx, rank, world, hit_capacity, _ptr(mine), C.byref(h)))
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.from_numpy(mine)
        t = t.cuda() if on_gpu else t
        allh = torch.empty((world, capi.COMM_HANDLE_BYTES), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(allh, t.reshape(1, -1), group=group)
        handles = np.ascontiguousarray(allh.cpu().numpy())
        st = capi.lib().lm_comm_connect(h, _ptr(handles))
        if st != capi.OK:
            msg = capi.lib().lm_comm_last_error(h).decode()
            capi.lib().lm_comm_destroy(h)
            raise LightmotifError(st, msg)
        self = cls(h, pli.alphabet)
        self.pipelines = [pli]  # the caller's own wrapper (same context)
        dist.barrier(group)     # every rank has mapped every buffer before anyone writes
        return self

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().lm_comm_destroy(self._h)
            self._h = None
            for p in self.pipelines:  # contexts owned by the communicator are gone: slabs still alive
                if getattr(p, "_borrowed", False):  # free their memory without one (lm_seq_free(NULL, .))
                    p._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st < 0:
            raise LightmotifError(st, capi.lib().lm_comm_last_error(self._h).decode())
        return st

    def _slabs(self, handles):
        return [DeviceSequence(self.pipelines[i], C.c_void_p(handles[i])) for i in range(self.n_local)]

    def sequence_synthetic(self, total_length, wrap, seed=0x5EED):
        hs = (C.c_void_p * self.n_local)()
        self._check(capi.lib().lm_comm_seq_synthetic(self._h, self.alphabet, seed, total_length, wrap, hs))
        return self._slabs(hs)

    def shard_from_striped(self, data, length, wrap):
        hs = (C.c_void_p * self.n_local)()
        rows = data.shape[0] - wrap
        self._check(capi.lib().lm_comm_shard_from_striped(self._h, self.alphabet, _ptr(data), rows, wrap,
                                                          data.strides[0], length, hs))
        return self._slabs(hs)

    def _harr(self, slabs):
        return (C.c_void_p * self.n_local)(*[s._h for s in slabs])

    def score_argmax_launch(self, slabs, pssm, slot=0):
        p, stride = pad_pssm(pssm, self.alphabet)
        self._check(capi.lib().lm_comm_score_argmax_f32_launch(self._h, self._harr(slabs), _ptr(p), p.shape[0], stride, slot))

    def score_argmax_finish(self, slot=0, rule=RULE_AVX2):
        best = Hit()
        st = self._check(capi.lib().lm_comm_score_argmax_f32_finish(self._h, slot, rule, C.byref(best)))
        return None if st == capi.EMPTY else _hit(best)

    def score_argmax(self, slabs, pssm, rule=RULE_AVX2):
        self.score_argmax_launch(slabs, pssm, 0)
        return self.score_argmax_finish(0, rule)

    def score_threshold(self, slabs, pssm, t, cap=1 << 20, with_counts=False):
        p, stride = pad_pssm(pssm, self.alphabet)
        hits = (Hit * cap)()
        count = C.c_size_t(0)
        local = (C.c_size_t * self.world)()
        st = capi.lib().lm_seq_score_threshold_f32_sharded(self._h, self._harr(slabs), _ptr(p), p.shape[0], stride,
                                                           float(t), hits, cap, C.byref(count), local)
        if st == capi.ERR_OVERFLOW and count.value > cap:
            return self.score_threshold(slabs, pssm, t, cap=count.value, with_counts=with_counts)
        self._check(st)
        out = [_hit(hits[i]) for i in range(count.value)]
        return (out, [int(x) for x in local]) if with_counts else out
