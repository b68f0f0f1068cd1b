//! `lightmotif::pli::platform::cuda` -- the back-end a lightmotif maintainer
//! adds to bind liblightmotif_b200.so.  SOURCE ONLY: there is no Rust
//! toolchain in the build image, so this file has not been compiled; the C ABI
//! it binds is exercised from C++ (tests/cpp) and Python (tests/test_gpu_*.py).
//!
//! Drop next to `avx2.rs` (lightmotif/src/pli/platform/) and add `pub mod cuda;`
//! + `pub use cuda::Cuda;` to platform/mod.rs.  Nothing else in the crate changes:
//! `Cuda` is a zero-sized marker like `Avx2` / `Sse2` / `Neon` (platform/mod.rs:17-21),
//! so `Pipeline<A, Cuda>` stays `Default + Clone + Debug` (pli/mod.rs:227-231) and is
//! built by `Pipeline::cuda()` exactly as `Pipeline::avx2()` is (pli/mod.rs:401-407).
//! The device context the C ABI needs (one per host thread: stream + scratch) lives
//! in a thread-local that is created on first use and destroyed with the thread.

use std::ops::Range;
use std::os::raw::{c_int, c_void};

use crate::abc::{Alphabet, Dna, Symbol};
use crate::dense::{DenseMatrix, MatrixCoordinates};
use crate::err::{InvalidSymbol, UnsupportedBackend};
use crate::num::{Unsigned, U32};
use crate::pli::{Encode, Maximum, Pipeline, Score, Stripe, Threshold};
use crate::scores::StripedScores;
use crate::seq::StripedSequence;

use super::Backend;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct LmHit {
    pub row: u64,
    pub position: u64,
    pub col: u32,
    pub score: f32,
}

#[allow(non_camel_case_types)]
type lm_ctx = c_void;

#[link(name = "lightmotif_b200")]
extern "C" {
    fn lm_ctx_create(device: c_int, out: *mut *mut lm_ctx) -> c_int;
    fn lm_ctx_destroy(ctx: *mut lm_ctx);
    fn lm_encode(ctx: *mut lm_ctx, alphabet: c_int, ascii: *const u8, n: usize, dst: *mut u8,
                 bad_index: *mut usize) -> c_int;
    fn lm_stripe(ctx: *mut lm_ctx, alphabet: c_int, sym: *const u8, n: usize, dst: *mut u8,
                 dst_stride: usize) -> c_int;
    fn lm_score_f32_rows(ctx: *mut lm_ctx, alphabet: c_int, pssm: *const f32, m: usize,
                         pssm_stride: usize, seq: *const u8, rows: usize, wrap: usize,
                         seq_stride: usize, length: usize, row_begin: usize, row_end: usize,
                         out: *mut f32, out_stride: usize, out_rows: *mut usize) -> c_int;
    fn lm_score_u8_rows(ctx: *mut lm_ctx, alphabet: c_int, dm: *const u8, m: usize,
                        dm_stride: usize, seq: *const u8, rows: usize, wrap: usize,
                        seq_stride: usize, length: usize, row_begin: usize, row_end: usize,
                        out: *mut u8, out_stride: usize, out_rows: *mut usize) -> c_int;
    fn lm_argmax_f32(ctx: *mut lm_ctx, scores: *const f32, rows: usize, stride: usize,
                     rule: c_int, best: *mut LmHit) -> c_int;
    fn lm_max_f32_rule(ctx: *mut lm_ctx, scores: *const f32, rows: usize, stride: usize,
                       rule: c_int, best: *mut f32) -> c_int;
    fn lm_threshold_f32(ctx: *mut lm_ctx, scores: *const f32, rows: usize, stride: usize, t: f32,
                        hits: *mut LmHit, cap: usize, count: *mut usize) -> c_int;
}

const LM_OK: c_int = 0;
const LM_EMPTY: c_int = 1;
const LM_ERR_INVALID_SYMBOL: c_int = -2;
const LM_ERR_WRAP: c_int = -3;
const LM_ERR_OVERFLOW: c_int = -7;
const LM_RULE_GENERIC: c_int = 0;
const LM_RULE_AVX2: c_int = 1;
/// Which reference back-end `Maximum::max` imitates.  `Pipeline<A, Cuda>` is a back-end of its own,
/// so it returns the TRUE maximum (the generic definition, pli/mod.rs:158-160); set this to
/// `LM_RULE_AVX2` to reproduce what `Pipeline::dispatch()` returns on an AVX2 host, whose
/// accumulators start at 0.0 (avx2.rs:438-441) and therefore never report less.
const MAX_RULE: c_int = LM_RULE_GENERIC;

/// A marker type for the CUDA (B200, sm_100a) implementation of the pipeline: zero-sized, like
/// the SIMD markers.
#[derive(Debug, Default, Clone, Copy)]
pub struct Cuda;

impl Backend for Cuda {
    type Lanes = U32; // the device reads DenseMatrix<_, U32> rows as-is (one warp per row)
}

/// This thread's `lm_ctx` (an `lm_ctx` is not thread-safe; the reference's `Pipeline` is shared
/// freely between threads, lightmotif-cli/src/main.rs:264-369, so every thread gets its own).
struct ThreadCtx(std::cell::Cell<*mut lm_ctx>);

impl Drop for ThreadCtx {
    fn drop(&mut self) {
        let p = self.0.get();
        if !p.is_null() {
            unsafe { lm_ctx_destroy(p) }
        }
    }
}

thread_local! {
    static CTX: ThreadCtx = ThreadCtx(std::cell::Cell::new(std::ptr::null_mut()));
}

/// The calling thread's context, created on first use; `None` when no sm_100 device is usable.
fn try_ctx() -> Option<*mut lm_ctx> {
    CTX.with(|c| {
        if c.0.get().is_null() {
            let device = std::env::var("LIGHTMOTIF_CUDA_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
            let mut p = std::ptr::null_mut();
            if unsafe { lm_ctx_create(device, &mut p) } != LM_OK {
                return None;
            }
            c.0.set(p);
        }
        Some(c.0.get())
    })
}

fn ctx() -> *mut lm_ctx {
    try_ctx().expect("lightmotif_b200: no usable CUDA device (Pipeline::cuda() would have returned UnsupportedBackend)")
}

/// Alphabet tag passed over the ABI (LM_DNA = 0, LM_PROTEIN = 1).
fn alphabet_tag<A: Alphabet>() -> c_int {
    if A::K::USIZE == 5 { 0 } else { 1 }
}

impl<A: Alphabet> Pipeline<A, Cuda> {
    /// Attempt to create a new CUDA-accelerated pipeline (cf. `Pipeline::avx2`,
    /// pli/mod.rs:401-407).
    pub fn cuda() -> Result<Self, UnsupportedBackend> {
        match try_ctx() {
            Some(_) => Ok(Self::default()),
            None => Err(UnsupportedBackend),
        }
    }
}

impl<A: Alphabet> Encode<A> for Pipeline<A, Cuda> {
    fn encode_into<S: AsRef<[u8]>>(&self, seq: S, dst: &mut [A::Symbol]) -> Result<(), InvalidSymbol> {
        let s = seq.as_ref();
        assert_eq!(s.len(), dst.len());
        let mut bad = 0usize;
        let st = unsafe {
            lm_encode(ctx(), alphabet_tag::<A>(), s.as_ptr(), s.len(),
                      dst.as_mut_ptr() as *mut u8, &mut bad)
        };
        match st {
            LM_OK => Ok(()),
            LM_ERR_INVALID_SYMBOL => Err(InvalidSymbol(s[bad] as char)),
            e => panic!("lightmotif_b200: lm_encode failed with status {}", e),
        }
    }
}

impl<A: Alphabet> Stripe<A, U32> for Pipeline<A, Cuda> {
    fn stripe_into<S: AsRef<[A::Symbol]>>(&self, seq: S, striped: &mut StripedSequence<A, U32>) {
        let s = seq.as_ref();
        let rows = s.len().div_ceil(32);
        let mut data = std::mem::take(striped).into_matrix();
        data.reserve(rows + crate::seq::DEFAULT_EXTRA_ROWS);
        data.resize(rows);
        if rows > 0 {
            let st = unsafe {
                lm_stripe(ctx(), alphabet_tag::<A>(), s.as_ptr() as *const u8, s.len(),
                          data[0].as_mut_ptr() as *mut u8, data.stride())
            };
            assert_eq!(st, LM_OK, "lightmotif_b200: lm_stripe failed");
        }
        *striped = StripedSequence::new(data, s.len()).unwrap();
    }
}

impl<A: Alphabet> Score<f32, A, U32> for Pipeline<A, Cuda> {
    fn score_rows_into<S, M>(&self, pssm: M, seq: S, rows: Range<usize>, scores: &mut StripedScores<f32, U32>)
    where
        S: AsRef<StripedSequence<A, U32>>,
        M: AsRef<DenseMatrix<f32, A::K>>,
    {
        let seq = seq.as_ref();
        let pssm = pssm.as_ref();
        if seq.len() < pssm.rows() || rows.is_empty() {
            scores.resize(0, 0);
            return;
        }
        // the Rust side keeps owning (and resizing) the output, avx2.rs:844
        scores.resize(rows.len(), (seq.len() + 1).saturating_sub(pssm.rows()));
        let m = seq.matrix();
        let mut out_rows = 0usize;
        let st = unsafe {
            lm_score_f32_rows(
                ctx(), alphabet_tag::<A>(), pssm[0].as_ptr(), pssm.rows(), pssm.stride(),
                m[0].as_ptr() as *const u8, m.rows() - seq.wrap(), seq.wrap(), m.stride(), seq.len(),
                rows.start, rows.end, scores.matrix_mut()[0].as_mut_ptr(),
                scores.matrix().stride(), &mut out_rows)
        };
        match st {
            LM_OK => {}
            LM_ERR_WRAP => panic!("not enough wrapping rows for motif of length {}", pssm.rows()),
            e => panic!("lightmotif_b200: lm_score_f32_rows failed with status {}", e),
        }
    }
}

impl Score<u8, Dna, U32> for Pipeline<Dna, Cuda> {
    fn score_rows_into<S, M>(&self, pssm: M, seq: S, rows: Range<usize>, scores: &mut StripedScores<u8, U32>)
    where
        S: AsRef<StripedSequence<Dna, U32>>,
        M: AsRef<DenseMatrix<u8, <Dna as Alphabet>::K>>,
    {
        let seq = seq.as_ref();
        let pssm = pssm.as_ref();
        if seq.len() < pssm.rows() || rows.is_empty() {
            scores.resize(0, 0);
            return;
        }
        scores.resize(rows.len(), (seq.len() + 1).saturating_sub(pssm.rows()));
        let m = seq.matrix();
        let mut out_rows = 0usize;
        let st = unsafe {
            lm_score_u8_rows(
                ctx(), 0, pssm[0].as_ptr(), pssm.rows(), pssm.stride(),
                m[0].as_ptr() as *const u8, m.rows() - seq.wrap(), seq.wrap(), m.stride(), seq.len(),
                rows.start, rows.end, scores.matrix_mut()[0].as_mut_ptr(),
                scores.matrix().stride(), &mut out_rows)
        };
        if st == LM_ERR_WRAP {
            panic!("not enough wrapping rows for motif of length {}", pssm.rows());
        }
        assert_eq!(st, LM_OK);
    }
}

impl<A: Alphabet> Maximum<f32, U32> for Pipeline<A, Cuda> {
    fn argmax(&self, scores: &StripedScores<f32, U32>) -> Option<MatrixCoordinates> {
        let m = scores.matrix();
        let mut best = LmHit::default();
        let ptr = if m.rows() > 0 { m[0].as_ptr() } else { std::ptr::null() };
        match unsafe { lm_argmax_f32(ctx(), ptr, m.rows(), m.stride(), LM_RULE_AVX2, &mut best) } {
            LM_OK => Some(MatrixCoordinates::new(best.row as usize, best.col as usize)),
            LM_EMPTY => None,
            e => panic!("lightmotif_b200: lm_argmax_f32 failed with status {}", e),
        }
    }

    fn max(&self, scores: &StripedScores<f32, U32>) -> Option<f32> {
        let m = scores.matrix();
        let mut best = 0f32;
        let ptr = if m.rows() > 0 { m[0].as_ptr() } else { std::ptr::null() };
        match unsafe { lm_max_f32_rule(ctx(), ptr, m.rows(), m.stride(), MAX_RULE, &mut best) } {
            LM_OK => Some(best),
            LM_EMPTY => None,
            e => panic!("lightmotif_b200: lm_max_f32_rule failed with status {}", e),
        }
    }
}

impl<A: Alphabet> Threshold<f32, U32> for Pipeline<A, Cuda> {
    fn threshold(&self, scores: &StripedScores<f32, U32>, threshold: f32) -> Vec<MatrixCoordinates> {
        let m = scores.matrix();
        if m.rows() == 0 {
            return Vec::new();
        }
        let mut cap = 1024usize;
        loop {
            let mut hits = vec![LmHit::default(); cap];
            let mut count = 0usize;
            let st = unsafe {
                lm_threshold_f32(ctx(), m[0].as_ptr(), m.rows(), m.stride(), threshold,
                                 hits.as_mut_ptr(), cap, &mut count)
            };
            match st {
                LM_OK => {
                    hits.truncate(count);
                    return hits.iter().map(|h| MatrixCoordinates::new(h.row as usize, h.col as usize)).collect();
                }
                LM_ERR_OVERFLOW => cap = count,
                e => panic!("lightmotif_b200: lm_threshold_f32 failed with status {}", e),
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Resident sequences and fused scans (what bulk callers -- the CLI job loop,
// `Scanner` -- should use: the sequence is uploaded once, scores never leave
// the GPU).  Not part of the reference's trait surface; a thin inherent API.

#[allow(non_camel_case_types)]
type lm_seq = c_void;

extern "C" {
    fn lm_seq_from_ascii(ctx: *mut lm_ctx, alphabet: c_int, ascii: *const u8, n: usize, lossy: c_int,
                         wrap: usize, out: *mut *mut lm_seq, bad_index: *mut usize) -> c_int;
    fn lm_seq_from_striped(ctx: *mut lm_ctx, alphabet: c_int, data: *const u8, rows: usize, wrap: usize,
                           stride: usize, length: usize, out: *mut *mut lm_seq) -> c_int;
    fn lm_seq_info(seq: *const lm_seq, alphabet: *mut c_int, rows: *mut usize, wrap: *mut usize,
                   length: *mut usize) -> c_int;
    fn lm_seq_free(ctx: *mut lm_ctx, seq: *mut lm_seq);
    fn lm_seq_score_argmax_f32(ctx: *mut lm_ctx, seq: *const lm_seq, pssm: *const f32, m: usize,
                               pssm_stride: usize, row_begin: usize, row_end: usize, rule: c_int,
                               best: *mut LmHit, per_column: *mut c_void, first_cell: *mut f32) -> c_int;
    fn lm_seq_score_threshold_f32(ctx: *mut lm_ctx, seq: *const lm_seq, pssm: *const f32, m: usize,
                                  pssm_stride: usize, row_begin: usize, row_end: usize, t: f32,
                                  hits: *mut LmHit, cap: usize, count: *mut usize) -> c_int;
}

const LM_RULE_SSE2: c_int = 2;

/// A `StripedSequence` resident in GPU memory (cf. the `Arc<StripedSequence>` the CLI shares
/// between its workers, lightmotif-cli/src/main.rs:540-562).
pub struct CudaSequence<'p, A: Alphabet> {
    pli: &'p Pipeline<A, Cuda>,
    seq: *mut lm_seq,
    rows: usize,
    length: usize,
}

impl<A: Alphabet> Drop for CudaSequence<'_, A> {
    fn drop(&mut self) {
        unsafe { lm_seq_free(ctx(), self.seq) }
    }
}

impl<A: Alphabet> Pipeline<A, Cuda> {
    /// `EncodedSequence::encode(text)?.to_striped()` + `configure_wrap(wrap)` in one device pass.
    pub fn upload_ascii<S: AsRef<[u8]>>(&self, text: S, wrap: usize) -> Result<CudaSequence<'_, A>, InvalidSymbol> {
        let s = text.as_ref();
        let (mut seq, mut bad) = (std::ptr::null_mut(), 0usize);
        match unsafe { lm_seq_from_ascii(ctx(), alphabet_tag::<A>(), s.as_ptr(), s.len(), 0, wrap, &mut seq, &mut bad) } {
            LM_OK => Ok(self.adopt(seq)),
            LM_ERR_INVALID_SYMBOL => Err(InvalidSymbol(s[bad] as char)),
            e => panic!("lightmotif_b200: lm_seq_from_ascii failed with status {}", e),
        }
    }

    /// Upload an existing (configured) `StripedSequence`.
    pub fn upload(&self, striped: &StripedSequence<A, U32>) -> CudaSequence<'_, A> {
        let m = striped.matrix();
        let mut seq = std::ptr::null_mut();
        let st = unsafe {
            lm_seq_from_striped(ctx(), alphabet_tag::<A>(), m[0].as_ptr() as *const u8,
                                m.rows() - striped.wrap(), striped.wrap(), m.stride(), striped.len(), &mut seq)
        };
        assert_eq!(st, LM_OK, "lightmotif_b200: lm_seq_from_striped failed");
        self.adopt(seq)
    }

    fn adopt(&self, seq: *mut lm_seq) -> CudaSequence<'_, A> {
        let (mut a, mut rows, mut wrap, mut length) = (0, 0usize, 0usize, 0usize);
        unsafe { lm_seq_info(seq, &mut a, &mut rows, &mut wrap, &mut length) };
        CudaSequence { pli: self, seq, rows, length }
    }
}

impl<A: Alphabet> CudaSequence<'_, A> {
    /// `pssm.score(&seq).argmax()` fused: `(position, score)` of the cell `Pipeline::dispatch()`
    /// would return (AVX2 tie rule), bit-identical score.
    pub fn score_argmax<M: AsRef<DenseMatrix<f32, A::K>>>(&self, pssm: M) -> Option<(usize, f32)> {
        self.argmax_with_rule(pssm, LM_RULE_AVX2)
    }

    fn argmax_with_rule<M: AsRef<DenseMatrix<f32, A::K>>>(&self, pssm: M, rule: c_int) -> Option<(usize, f32)> {
        let p = pssm.as_ref();
        let mut best = LmHit::default();
        let st = unsafe {
            lm_seq_score_argmax_f32(ctx(), self.seq, p[0].as_ptr(), p.rows(), p.stride(), 0, self.rows,
                                    rule, &mut best, std::ptr::null_mut(), std::ptr::null_mut())
        };
        match st {
            LM_OK => Some((best.position as usize, best.score)),
            LM_EMPTY => None,
            LM_ERR_WRAP => panic!("not enough wrapping rows for motif of length {}", p.rows()),
            e => panic!("lightmotif_b200: lm_seq_score_argmax_f32 failed with status {}", e),
        }
    }

    /// What `Scanner::new(pssm, seq).threshold(t).collect()` yields (scan.rs:169-198): every position
    /// p with p + M <= len whose exact f32 score is >= t, sorted by position.
    pub fn scan<M: AsRef<DenseMatrix<f32, A::K>>>(&self, pssm: M, t: f32) -> Vec<(usize, f32)> {
        let p = pssm.as_ref();
        let mut cap = 1024usize;
        loop {
            let mut hits = vec![LmHit::default(); cap];
            let mut count = 0usize;
            let st = unsafe {
                lm_seq_score_threshold_f32(ctx(), self.seq, p[0].as_ptr(), p.rows(), p.stride(), 0,
                                           self.rows, t, hits.as_mut_ptr(), cap, &mut count)
            };
            match st {
                LM_OK => {
                    hits.truncate(count);
                    let mut out: Vec<(usize, f32)> = hits.iter()
                        .filter(|h| h.position as usize + p.rows() <= self.length)
                        .map(|h| (h.position as usize, h.score)).collect();
                    out.sort_by_key(|h| h.0);
                    return out;
                }
                LM_ERR_OVERFLOW => cap = count,
                e => panic!("lightmotif_b200: lm_seq_score_threshold_f32 failed with status {}", e),
            }
        }
    }

    /// `Scanner::max()` (scan.rs:200-249): best hit, ties towards the larger position.
    pub fn scan_max<M: AsRef<DenseMatrix<f32, A::K>>>(&self, pssm: M) -> Option<(usize, f32)> {
        self.argmax_with_rule(pssm, LM_RULE_SSE2)
    }
}

// ---------------------------------------------------------------------------
// Row-sharded scans over all GPUs of the box (one process, lm_comm_init): the natural
// sharding axis is the row range of `score_rows_into` (pli/mod.rs:72-77).

#[allow(non_camel_case_types)]
type lm_comm = c_void;

extern "C" {
    fn lm_comm_init(n_dev: c_int, devices: *const c_int, hit_capacity: usize, out: *mut *mut lm_comm) -> c_int;
    fn lm_comm_destroy(comm: *mut lm_comm);
    fn lm_comm_ctx(comm: *mut lm_comm, local_index: c_int) -> *mut lm_ctx;
    fn lm_comm_shard_from_striped(comm: *mut lm_comm, alphabet: c_int, data: *const u8, rows: usize, wrap: usize,
                                  stride: usize, length: usize, slabs: *mut *mut lm_seq) -> c_int;
    fn lm_seq_score_argmax_f32_sharded(comm: *mut lm_comm, slabs: *const *const lm_seq, pssm: *const f32, m: usize,
                                       pssm_stride: usize, rule: c_int, best: *mut LmHit) -> c_int;
    fn lm_seq_score_threshold_f32_sharded(comm: *mut lm_comm, slabs: *const *const lm_seq, pssm: *const f32, m: usize,
                                          pssm_stride: usize, t: f32, hits: *mut LmHit, cap: usize,
                                          count: *mut usize, local_counts: *mut usize) -> c_int;
}

/// A `StripedSequence` cut into row slabs resident on `n` GPUs.
pub struct ShardedSequence<A: Alphabet> {
    comm: *mut lm_comm,
    slabs: Vec<*mut lm_seq>,
    length: usize,
    _a: std::marker::PhantomData<A>,
}

impl<A: Alphabet> Drop for ShardedSequence<A> {
    fn drop(&mut self) {
        for (i, s) in self.slabs.iter().enumerate() {
            unsafe { lm_seq_free(lm_comm_ctx(self.comm, i as c_int), *s) }
        }
        unsafe { lm_comm_destroy(self.comm) }
    }
}

impl<A: Alphabet> ShardedSequence<A> {
    /// Shard a configured `StripedSequence` over GPUs `0..n`.
    pub fn new(striped: &StripedSequence<A, U32>, n: usize) -> Result<Self, UnsupportedBackend> {
        let mut comm = std::ptr::null_mut();
        if unsafe { lm_comm_init(n as c_int, std::ptr::null(), 0, &mut comm) } != LM_OK {
            return Err(UnsupportedBackend);
        }
        let m = striped.matrix();
        let mut slabs = vec![std::ptr::null_mut(); n];
        let st = unsafe {
            lm_comm_shard_from_striped(comm, alphabet_tag::<A>(), m[0].as_ptr() as *const u8, m.rows() - striped.wrap(),
                                       striped.wrap(), m.stride(), striped.len(), slabs.as_mut_ptr())
        };
        assert_eq!(st, LM_OK, "lightmotif_b200: lm_comm_shard_from_striped failed");
        Ok(Self { comm, slabs, length: striped.len(), _a: std::marker::PhantomData })
    }

    /// `pssm.score(&seq).argmax()` over the whole sequence: `(position, score)`.
    pub fn score_argmax<M: AsRef<DenseMatrix<f32, A::K>>>(&self, pssm: M) -> Option<(usize, f32)> {
        let p = pssm.as_ref();
        let mut best = LmHit::default();
        let st = unsafe {
            lm_seq_score_argmax_f32_sharded(self.comm, self.slabs.as_ptr() as *const *const lm_seq, p[0].as_ptr(), p.rows(),
                                            p.stride(), LM_RULE_AVX2, &mut best)
        };
        match st {
            LM_OK => Some((best.position as usize, best.score)),
            LM_EMPTY => None,
            e => panic!("lightmotif_b200: lm_seq_score_argmax_f32_sharded failed with status {}", e),
        }
    }

    /// `Scanner::new(pssm, seq).threshold(t).collect()` over the whole sequence.
    pub fn scan<M: AsRef<DenseMatrix<f32, A::K>>>(&self, pssm: M, t: f32) -> Vec<(usize, f32)> {
        let p = pssm.as_ref();
        let mut cap = 4096usize;
        loop {
            let mut hits = vec![LmHit::default(); cap];
            let mut count = 0usize;
            let st = unsafe {
                lm_seq_score_threshold_f32_sharded(self.comm, self.slabs.as_ptr() as *const *const lm_seq, p[0].as_ptr(),
                                                   p.rows(), p.stride(), t, hits.as_mut_ptr(), cap, &mut count,
                                                   std::ptr::null_mut())
            };
            match st {
                LM_OK => {
                    hits.truncate(count);
                    let mut out: Vec<(usize, f32)> = hits.iter()
                        .filter(|h| h.position as usize + p.rows() <= self.length)
                        .map(|h| (h.position as usize, h.score)).collect();
                    out.sort_by_key(|h| h.0);
                    return out;
                }
                LM_ERR_OVERFLOW if count > cap => cap = count,
                e => panic!("lightmotif_b200: lm_seq_score_threshold_f32_sharded failed with status {}", e),
            }
        }
    }
}
