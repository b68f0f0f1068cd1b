// lightmotif_b200.hpp -- C++17 host-side mirror of lightmotif's pipeline surface
// over the C ABI (lightmotif_b200.h).
//
// The reference's host language is Rust; no Rust toolchain exists in the build
// image, so this header is the compiled-language host layer above the C ABI
// (the Rust shim a maintainer would add is bindings/rust/cuda.rs).  Names,
// argument meaning and error behaviour follow the reference:
//
//   lightmotif::Pipeline<A>::cuda()       Pipeline::avx2()           pli/mod.rs:399-408
//   Pipeline::encode / encode_into        trait Encode               pli/mod.rs:34-67
//   Pipeline::stripe / stripe_into        trait Stripe               pli/mod.rs:164-201
//   Pipeline::score / score_into /
//             score_rows_into             trait Score<T, A, C>       pli/mod.rs:70-130
//   Pipeline::argmax / max                trait Maximum<T, C>        pli/mod.rs:133-161
//   Pipeline::threshold                   trait Threshold<T, C>      pli/mod.rs:204-222
//   EncodedSequence / StripedSequence     seq.rs:100-175, 288-381
//   StripedScores                         scores.rs:101-213
//   Scanner                               scan.rs:96-250
//   CountMatrix -> FrequencyMatrix -> WeightMatrix -> ScoringMatrix -> DiscreteMatrix
//                                         pwm/mod.rs:150-258, 376-430, 513-526, 640-696, 763-790
//
// Everything that touches sequence data runs on the GPU through the C ABI; the
// O(M x K) matrix constructors are plain host code, as in the reference.  There
// is no CPU fall-back: Pipeline::cuda() throws UnsupportedBackend without an
// sm_100 device.  Header-only; link with -llightmotif_b200.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "lightmotif_b200.h"

namespace lightmotif {

// ---------------------------------------------------------------- alphabets
// abc.rs:88-135 (Dna: A, C, T, G, N) and abc.rs:190-256 (Protein, X last)
struct Dna {
  static constexpr int ID = LM_DNA;
  static constexpr int K = 5;
  static constexpr const char* SYMBOLS = "ACTGN";
  static constexpr uint8_t DEFAULT = 4;  // N
};
struct Protein {
  static constexpr int ID = LM_PROTEIN;
  static constexpr int K = 21;
  static constexpr const char* SYMBOLS = "ACDEFGHIKLMNPQRSTVWYX";
  static constexpr uint8_t DEFAULT = 20;  // X
};

// ------------------------------------------------------------------- errors
// err.rs:10 InvalidSymbol(char), err.rs:22 InvalidData, err.rs:34 UnsupportedBackend
struct InvalidSymbol : std::runtime_error {
  char symbol;
  explicit InvalidSymbol(char c) : std::runtime_error(std::string("invalid symbol: '") + c + "'"), symbol(c) {}
};
struct InvalidData : std::runtime_error {
  InvalidData() : std::runtime_error("invalid data") {}
};
struct UnsupportedBackend : std::runtime_error {
  UnsupportedBackend() : std::runtime_error("unsupported backend: no sm_100 CUDA device") {}
};
struct BackendError : std::runtime_error {
  int status;
  BackendError(int st, const std::string& detail)
      : std::runtime_error(std::string(lm_strerror(st)) + (detail.empty() ? "" : ": " + detail)), status(st) {}
};

constexpr size_t LANES = LM_LANES;  // C = U32, the AVX2 lane count (pli/dispatch.rs:45)

// ------------------------------------------------------------- DenseMatrix
// dense.rs:43-55,125-128: row-major, every row 32-byte aligned, so the row
// stride is the row size rounded up to 32 bytes.  Padding cells are never read.
template <typename T, size_t C>
class DenseMatrix {
 public:
  static constexpr size_t STRIDE = ((C * sizeof(T) + 31) / 32) * 32 / sizeof(T);
  DenseMatrix() = default;
  explicit DenseMatrix(size_t rows) { resize(rows); }
  DenseMatrix(const DenseMatrix& o) { *this = o; }
  DenseMatrix(DenseMatrix&& o) noexcept { *this = std::move(o); }
  DenseMatrix& operator=(const DenseMatrix& o) {
    if (this != &o) {
      resize(o.rows_);
      if (rows_) std::memcpy(data_, o.data_, rows_ * STRIDE * sizeof(T));
    }
    return *this;
  }
  DenseMatrix& operator=(DenseMatrix&& o) noexcept {
    if (this != &o) {
      release();
      data_ = o.data_; rows_ = o.rows_; cap_ = o.cap_;
      o.data_ = nullptr; o.rows_ = o.cap_ = 0;
    }
    return *this;
  }
  ~DenseMatrix() { release(); }
  static constexpr size_t columns() { return C; }
  static constexpr size_t stride() { return STRIDE; }
  size_t rows() const { return rows_; }
  // keeps the leading rows, zero-fills new ones (dense.rs resize default-fills)
  void resize(size_t rows) {
    if (rows > cap_) {
      const size_t bytes = std::max<size_t>(rows * STRIDE * sizeof(T), 32);
      T* fresh = static_cast<T*>(::operator new(bytes, std::align_val_t(32)));
      std::memset(static_cast<void*>(fresh), 0, bytes);
      if (rows_) std::memcpy(static_cast<void*>(fresh), data_, rows_ * STRIDE * sizeof(T));
      release();
      data_ = fresh;
      cap_ = rows;
    } else if (rows > rows_) {
      std::memset(static_cast<void*>(data_ + rows_ * STRIDE), 0, (rows - rows_) * STRIDE * sizeof(T));
    }
    rows_ = rows;
  }
  T* data() { return data_; }
  const T* data() const { return data_; }
  T* operator[](size_t r) { return data_ + r * STRIDE; }
  const T* operator[](size_t r) const { return data_ + r * STRIDE; }

 private:
  void release() {
    if (data_) ::operator delete(static_cast<void*>(data_), std::align_val_t(32));
    data_ = nullptr;
  }
  T* data_ = nullptr;  // 32-byte aligned, rows_ x STRIDE
  size_t rows_ = 0, cap_ = 0;
};

struct MatrixCoordinates {  // dense.rs:28-39
  size_t row, col;
  bool operator==(const MatrixCoordinates& o) const { return row == o.row && col == o.col; }
};

template <typename A> class Pipeline;
template <typename A> class StripedSequence;
template <typename A> class ScoringMatrix;
template <typename A> class DiscreteMatrix;

// -------------------------------------------------------- EncodedSequence
template <typename A>
class EncodedSequence {  // seq.rs:100-175
 public:
  EncodedSequence() = default;
  explicit EncodedSequence(std::vector<uint8_t> symbols) : data_(std::move(symbols)) {}
  // EncodedSequence::encode (seq.rs:111-114): Err(InvalidSymbol) on the first bad character
  static EncodedSequence encode(const Pipeline<A>& pli, std::string_view text) { return pli.encode(text); }
  // seq.rs:122-129: unknown characters become the default symbol
  static EncodedSequence encode_lossy(const Pipeline<A>& pli, std::string_view text) { return pli.encode_lossy(text); }
  size_t len() const { return data_.size(); }
  const std::vector<uint8_t>& symbols() const { return data_; }
  uint8_t operator[](size_t i) const { return data_[i]; }
  // seq.rs:168-175
  StripedSequence<A> to_striped(const Pipeline<A>& pli) const { return pli.stripe(*this); }
  std::string to_string() const {
    std::string s(data_.size(), '?');
    for (size_t i = 0; i < s.size(); ++i) s[i] = A::SYMBOLS[data_[i]];
    return s;
  }

 private:
  std::vector<uint8_t> data_;
};

// -------------------------------------------------------- StripedSequence
template <typename A>
class StripedSequence {  // seq.rs:288-381
 public:
  StripedSequence() = default;
  size_t len() const { return length_; }
  size_t wrap() const { return wrap_; }
  size_t rows() const { return matrix_.rows() - wrap_; }  // rows of the sequence proper
  const DenseMatrix<uint8_t, LANES>& matrix() const { return matrix_; }
  DenseMatrix<uint8_t, LANES>& matrix_mut() { return matrix_; }
  // seq.rs:433-442: element i lives at [i % rows][i / rows]
  uint8_t operator[](size_t i) const { return matrix_[i % rows()][i / rows()]; }
  // seq.rs:362-366: make sure the motif can slide over the lane boundaries
  void configure(const Pipeline<A>& pli, const ScoringMatrix<A>& pssm);
  // seq.rs:369-381 (only ever grows)
  void configure_wrap(const Pipeline<A>& pli, size_t m);

 private:
  friend class Pipeline<A>;
  DenseMatrix<uint8_t, LANES> matrix_;
  size_t length_ = 0, wrap_ = 0;
};

// ------------------------------------------------------------ StripedScores
template <typename T>
class StripedScores {  // scores.rs:101-213
 public:
  size_t max_index() const { return max_index_; }
  bool is_empty() const { return matrix_.rows() == 0; }
  const DenseMatrix<T, LANES>& matrix() const { return matrix_; }
  DenseMatrix<T, LANES>& matrix_mut() { return matrix_; }
  // scores.rs:155-157
  size_t offset(MatrixCoordinates mc) const { return mc.col * matrix_.rows() + mc.row; }
  // scores.rs:246-254
  T operator[](size_t i) const { return matrix_[i % matrix_.rows()][i / matrix_.rows()]; }
  // scores.rs:167-169, 272-279: truncated to min(max_index, rows * C)
  std::vector<T> unstripe() const {
    const size_t n = std::min(max_index_, matrix_.rows() * LANES);
    std::vector<T> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = (*this)[i];
    return out;
  }
  // scores.rs:181-213: positions, through the pipeline
  template <typename A> std::optional<size_t> argmax(const Pipeline<A>& pli) const {
    auto mc = pli.argmax(*this);
    return mc ? std::optional<size_t>(offset(*mc)) : std::nullopt;
  }
  template <typename A> std::optional<T> max(const Pipeline<A>& pli) const { return pli.max(*this); }
  template <typename A> std::vector<size_t> threshold(const Pipeline<A>& pli, T t) const {
    std::vector<size_t> out;
    for (auto mc : pli.threshold(*this, t)) out.push_back(offset(mc));
    return out;
  }
  void resize(size_t rows, size_t max_index) {  // scores.rs:131-134
    matrix_.resize(rows);
    max_index_ = max_index;
  }

 private:
  DenseMatrix<T, LANES> matrix_;
  size_t max_index_ = 0;
};

// ----------------------------------------------------------------- matrices
// abc.rs:465-487 uniform background (0 for the default symbol), abc.rs:558-574 pseudocounts
template <typename A>
struct Background {
  float frequencies[A::K];
  static Background uniform() {
    Background b;
    for (int i = 0; i < A::K; ++i) b.frequencies[i] = i != A::DEFAULT ? 1.0f / float(A::K - 1) : 0.0f;
    return b;
  }
};

template <typename A>
class DiscreteMatrix {  // pwm/mod.rs:700-790
 public:
  const DenseMatrix<uint8_t, A::K>& matrix() const { return data_; }
  size_t len() const { return data_.rows(); }
  // pwm/mod.rs:763-790
  uint8_t scale(float score) const {
    float x = std::floor((score - offset_) / factor_);
    return !(x > 0.f) ? 0 : (x >= 255.f ? 255 : (uint8_t)x);  // saturating `as u8`
  }
  float unscale(uint8_t score) const { return float(score) * factor_ + offset_; }
  StripedScores<uint8_t> score(const Pipeline<A>& pli, const StripedSequence<A>& seq) const { return pli.score(*this, seq); }

 private:
  friend class ScoringMatrix<A>;
  DenseMatrix<uint8_t, A::K> data_;
  float factor_ = 1.f, offset_ = 0.f;
  std::vector<float> offsets_;
};

template <typename A>
class ScoringMatrix {  // pwm/mod.rs:560-696
 public:
  ScoringMatrix() = default;
  ScoringMatrix(Background<A> bg, DenseMatrix<float, A::K> data) : bg_(bg), data_(std::move(data)) {}
  const DenseMatrix<float, A::K>& matrix() const { return data_; }
  size_t len() const { return data_.rows(); }
  // pwm/mod.rs:640-648
  StripedScores<float> score(const Pipeline<A>& pli, const StripedSequence<A>& seq) const { return pli.score(*this, seq); }
  // pwm/mod.rs:651-662: exact sequential f32 sum at one position
  float score_position(const StripedSequence<A>& seq, size_t pos) const {
    float s = 0.0f;
    for (size_t j = 0; j < len(); ++j) s += data_[j][seq[pos + j]];
    return s;
  }
  // pwm/mod.rs:600-612: sum of the row maxima (default symbol excluded)
  float max_score() const {
    float s = 0.f;
    for (size_t j = 0; j < len(); ++j) s += *std::max_element(data_[j], data_[j] + A::K - 1);
    return s;
  }
  // pwm/mod.rs:665-696
  DiscreteMatrix<A> to_discrete() const {
    DiscreteMatrix<A> dm;
    const float mx = max_score();
    dm.offsets_.resize(len());
    float offset = 0.f;
    for (size_t j = 0; j < len(); ++j) {
      float lo = std::numeric_limits<float>::infinity();
      for (int a = 0; a < A::K - 1; ++a) lo = std::min(lo, std::isinf(data_[j][a]) ? -mx : data_[j][a]);
      dm.offsets_[j] = lo;
      offset += lo;
    }
    dm.offset_ = offset;
    dm.factor_ = (mx - offset) / 255.0f;
    dm.data_.resize(len());
    for (size_t j = 0; j < len(); ++j)
      for (int a = 0; a < A::K; ++a) {
        float x = std::ceil((data_[j][a] - dm.offsets_[j]) / dm.factor_);
        dm.data_[j][a] = !(x > 0.f) ? 0 : (x >= 255.f ? 255 : (uint8_t)x);  // Rust `as u8`: NaN/-inf -> 0
      }
    return dm;
  }

 private:
  Background<A> bg_ = Background<A>::uniform();
  DenseMatrix<float, A::K> data_;
};

template <typename A>
class WeightMatrix {  // pwm/mod.rs:440-526
 public:
  WeightMatrix(Background<A> bg, DenseMatrix<float, A::K> data) : bg_(bg), data_(std::move(data)) {}
  const DenseMatrix<float, A::K>& matrix() const { return data_; }
  ScoringMatrix<A> to_scoring() const {  // pwm/mod.rs:513-526 (base 2)
    DenseMatrix<float, A::K> d = data_;
    for (size_t j = 0; j < d.rows(); ++j)
      for (int a = 0; a < A::K; ++a) d[j][a] = std::log2(d[j][a]);
    return ScoringMatrix<A>(bg_, std::move(d));
  }

 private:
  Background<A> bg_;
  DenseMatrix<float, A::K> data_;
};

template <typename A>
class FrequencyMatrix {  // pwm/mod.rs:300-430
 public:
  explicit FrequencyMatrix(DenseMatrix<float, A::K> data) : data_(std::move(data)) {}
  const DenseMatrix<float, A::K>& matrix() const { return data_; }
  WeightMatrix<A> to_weight(std::optional<Background<A>> background = std::nullopt) const {  // :376-392
    Background<A> bg = background.value_or(Background<A>::uniform());
    DenseMatrix<float, A::K> w(data_.rows());
    for (size_t j = 0; j < data_.rows(); ++j)
      for (int a = 0; a < A::K; ++a) w[j][a] = bg.frequencies[a] == 0.0f ? 0.0f : data_[j][a] / bg.frequencies[a];
    return WeightMatrix<A>(bg, std::move(w));
  }
  ScoringMatrix<A> to_scoring(std::optional<Background<A>> background = std::nullopt) const {  // :415-430
    Background<A> bg = background.value_or(Background<A>::uniform());
    DenseMatrix<float, A::K> s(data_.rows());
    for (size_t j = 0; j < data_.rows(); ++j)
      for (int a = 0; a < A::K; ++a)
        s[j][a] = bg.frequencies[a] == 0.0f ? -std::numeric_limits<float>::infinity()
                                            : std::log2(data_[j][a] / bg.frequencies[a]);
    return ScoringMatrix<A>(bg, std::move(s));
  }

 private:
  DenseMatrix<float, A::K> data_;
};

template <typename A>
class CountMatrix {  // pwm/mod.rs:150-258
 public:
  explicit CountMatrix(DenseMatrix<uint32_t, A::K> data) : data_(std::move(data)) {}
  // pwm/mod.rs:209-237: Err(InvalidData) when the sequences differ in length
  static CountMatrix from_sequences(const std::vector<EncodedSequence<A>>& seqs) {
    DenseMatrix<uint32_t, A::K> d(seqs.empty() ? 0 : seqs[0].len());
    for (size_t j = 0; j < d.rows(); ++j) std::fill(d[j], d[j] + A::K, 0u);
    for (const auto& s : seqs) {
      if (s.len() != d.rows()) throw InvalidData();
      for (size_t j = 0; j < s.len(); ++j) d[j][s[j]] += 1;
    }
    return CountMatrix(std::move(d));
  }
  const DenseMatrix<uint32_t, A::K>& matrix() const { return data_; }
  // pwm/mod.rs:240-258: counts + pseudocount (0 for the default symbol), rows normalised
  // by their sequential f32 sum
  FrequencyMatrix<A> to_freq(float pseudo) const {
    DenseMatrix<float, A::K> p(data_.rows());
    for (size_t j = 0; j < data_.rows(); ++j) {
      float s = 0.0f;
      for (int a = 0; a < A::K; ++a) {
        p[j][a] = float(data_[j][a]) + (a != A::DEFAULT ? pseudo : 0.0f);
        s += p[j][a];
      }
      for (int a = 0; a < A::K; ++a) p[j][a] /= s;
    }
    return FrequencyMatrix<A>(std::move(p));
  }

 private:
  DenseMatrix<uint32_t, A::K> data_;
};

// ----------------------------------------------------------------- Pipeline
// A sequence kept resident in HBM (what the CLI holds as Arc<StripedSequence>,
// lightmotif-cli/src/main.rs:540-562): upload once, scan with many motifs.
template <typename A>
class DeviceSequence {
 public:
  size_t rows() const { return rows_; }
  size_t wrap() const { return wrap_; }
  size_t len() const { return length_; }
  lm_seq* handle() const { return seq_.get(); }

 private:
  friend class Pipeline<A>;
  std::shared_ptr<lm_seq> seq_;
  size_t rows_ = 0, wrap_ = 0, length_ = 0;
};

struct Hit {  // scan.rs:20-40
  size_t position;
  float score;
};

template <typename A>
class Pipeline {
 public:
  // Pipeline::avx2()-style constructor (pli/mod.rs:399-408): Err(UnsupportedBackend)
  static Pipeline cuda(int device = 0) {
    lm_ctx* raw = nullptr;
    int st = lm_ctx_create(device, &raw);
    if (st == LM_ERR_UNSUPPORTED) throw UnsupportedBackend();
    if (st < 0) throw BackendError(st, "lm_ctx_create");
    Pipeline p;
    p.ctx_ = std::shared_ptr<lm_ctx>(raw, lm_ctx_destroy);
    return p;
  }
  lm_ctx* context() const { return ctx_.get(); }
  // tie rule applied by argmax: default = what Pipeline::dispatch() does on x86-64
  void set_argmax_rule(int rule) { rule_ = rule; }

  // ---- Encode (pli/mod.rs:34-67)
  void encode_into(std::string_view text, uint8_t* dst) const {
    size_t bad = 0;
    int st = lm_encode(ctx(), A::ID, reinterpret_cast<const uint8_t*>(text.data()), text.size(), dst, &bad);
    if (st == LM_ERR_INVALID_SYMBOL) throw InvalidSymbol(text[bad]);
    check(st, "lm_encode");
  }
  EncodedSequence<A> encode(std::string_view text) const {
    std::vector<uint8_t> sym(text.size());
    encode_into(text, sym.data());
    return EncodedSequence<A>(std::move(sym));
  }
  EncodedSequence<A> encode_lossy(std::string_view text) const {
    std::vector<uint8_t> sym(text.size());
    check(lm_encode_lossy(ctx(), A::ID, reinterpret_cast<const uint8_t*>(text.data()), text.size(), sym.data()),
          "lm_encode_lossy");
    return EncodedSequence<A>(std::move(sym));
  }

  // ---- Stripe (pli/mod.rs:164-201): result has wrap = 0
  void stripe_into(const EncodedSequence<A>& seq, StripedSequence<A>& out) const {
    const size_t rows = (seq.len() + LANES - 1) / LANES;
    out.matrix_.resize(rows);
    out.length_ = seq.len();
    out.wrap_ = 0;
    check(lm_stripe(ctx(), A::ID, seq.symbols().data(), seq.len(), out.matrix_.data(), out.matrix_.stride()), "lm_stripe");
  }
  StripedSequence<A> stripe(const EncodedSequence<A>& seq) const {
    StripedSequence<A> s;
    stripe_into(seq, s);
    return s;
  }
  void configure_wrap(StripedSequence<A>& seq, size_t m) const {  // seq.rs:369-381
    if (m <= seq.wrap_) return;
    const size_t rows = seq.rows();
    DenseMatrix<uint8_t, LANES> grown(rows + m);
    std::memcpy(grown.data(), seq.matrix_.data(), rows * grown.stride());
    check(lm_configure_wrap(ctx(), A::ID, grown.data(), rows, grown.stride(), m), "lm_configure_wrap");
    seq.matrix_ = std::move(grown);
    seq.wrap_ = m;
  }

  // ---- Score (pli/mod.rs:70-130).  Panics of the reference (wrap < M - 1,
  // avx2.rs:832-837) surface as BackendError(LM_ERR_WRAP).
  void score_rows_into(const ScoringMatrix<A>& pssm, const StripedSequence<A>& seq, size_t row_begin, size_t row_end,
                       StripedScores<float>& scores) const {
    score_rows_impl<float>(pssm.matrix().data(), pssm.len(), pssm.matrix().stride(), seq, row_begin, row_end, scores);
  }
  void score_rows_into(const DiscreteMatrix<A>& dm, const StripedSequence<A>& seq, size_t row_begin, size_t row_end,
                       StripedScores<uint8_t>& scores) const {
    score_rows_impl<uint8_t>(dm.matrix().data(), dm.len(), dm.matrix().stride(), seq, row_begin, row_end, scores);
  }
  template <typename M, typename T>
  void score_into(const M& pssm, const StripedSequence<A>& seq, StripedScores<T>& scores) const {  // :109-117
    score_rows_into(pssm, seq, 0, seq.rows(), scores);
  }
  StripedScores<float> score(const ScoringMatrix<A>& pssm, const StripedSequence<A>& seq) const {  // :120-129
    StripedScores<float> s;
    score_into(pssm, seq, s);
    return s;
  }
  StripedScores<uint8_t> score(const DiscreteMatrix<A>& dm, const StripedSequence<A>& seq) const {
    StripedScores<uint8_t> s;
    score_into(dm, seq, s);
    return s;
  }

  // ---- Maximum (pli/mod.rs:133-161)
  std::optional<MatrixCoordinates> argmax(const StripedScores<float>& s) const {
    lm_hit h;
    int st = lm_argmax_f32(ctx(), s.matrix().data(), s.matrix().rows(), s.matrix().stride(), rule_, &h);
    if (st == LM_EMPTY) return std::nullopt;
    check(st, "lm_argmax_f32");
    return MatrixCoordinates{(size_t)h.row, (size_t)h.col};
  }
  std::optional<MatrixCoordinates> argmax(const StripedScores<uint8_t>& s) const {
    lm_hit h;
    int st = lm_argmax_u8(ctx(), s.matrix().data(), s.matrix().rows(), s.matrix().stride(), rule_, &h);
    if (st == LM_EMPTY) return std::nullopt;
    check(st, "lm_argmax_u8");
    return MatrixCoordinates{(size_t)h.row, (size_t)h.col};
  }
  std::optional<float> max(const StripedScores<float>& s) const {
    float v = 0.f;
    int st = lm_max_f32(ctx(), s.matrix().data(), s.matrix().rows(), s.matrix().stride(), &v);
    if (st == LM_EMPTY) return std::nullopt;
    check(st, "lm_max_f32");
    return v;
  }
  std::optional<uint8_t> max(const StripedScores<uint8_t>& s) const {
    uint8_t v = 0;
    int st = lm_max_u8(ctx(), s.matrix().data(), s.matrix().rows(), s.matrix().stride(), &v);
    if (st == LM_EMPTY) return std::nullopt;
    check(st, "lm_max_u8");
    return v;
  }

  // ---- Threshold (pli/mod.rs:204-222): every cell >= t, row-major order
  std::vector<MatrixCoordinates> threshold(const StripedScores<float>& s, float t) const {
    return threshold_impl(s, [&](lm_hit* h, size_t cap, size_t* n) {
      return lm_threshold_f32(ctx(), s.matrix().data(), s.matrix().rows(), s.matrix().stride(), t, h, cap, n);
    });
  }
  std::vector<MatrixCoordinates> threshold(const StripedScores<uint8_t>& s, uint8_t t) const {
    return threshold_impl(s, [&](lm_hit* h, size_t cap, size_t* n) {
      return lm_threshold_u8(ctx(), s.matrix().data(), s.matrix().rows(), s.matrix().stride(), t, h, cap, n);
    });
  }

  // ---- resident sequences + fused scans (scores never leave the GPU)
  // encode[_lossy] -> to_striped -> configure_wrap in one device pass (seq.rs:111-129,168-175,369-381)
  DeviceSequence<A> upload_ascii(std::string_view text, size_t wrap, bool lossy = false) const {
    lm_seq* raw = nullptr;
    size_t bad = 0;
    int st = lm_seq_from_ascii(ctx(), A::ID, reinterpret_cast<const uint8_t*>(text.data()), text.size(), lossy ? 1 : 0,
                               wrap, &raw, &bad);
    if (st == LM_ERR_INVALID_SYMBOL) throw InvalidSymbol(text[bad]);
    check(st, "lm_seq_from_ascii");
    return adopt(raw);
  }
  DeviceSequence<A> upload(const StripedSequence<A>& seq) const {
    lm_seq* raw = nullptr;
    check(lm_seq_from_striped(ctx(), A::ID, seq.matrix().data(), seq.rows(), seq.wrap(), seq.matrix().stride(), seq.len(),
                              &raw), "lm_seq_from_striped");
    return adopt(raw);
  }
  // ScoringMatrix::score + StripedScores::argmax (pwm/mod.rs:640-648, scores.rs:190-192)
  std::optional<Hit> score_argmax(const ScoringMatrix<A>& pssm, const DeviceSequence<A>& seq) const {
    lm_hit h;
    int st = lm_seq_score_argmax_f32(ctx(), seq.handle(), pssm.matrix().data(), pssm.len(), pssm.matrix().stride(), 0,
                                     seq.rows(), rule_, &h, nullptr, nullptr);
    if (st == LM_EMPTY) return std::nullopt;
    check(st, "lm_seq_score_argmax_f32");
    return Hit{(size_t)h.position, h.score};
  }
  // ScoringMatrix::score + StripedScores::threshold (scores.rs:207-213); the device analogue of
  // Scanner (scan.rs:160-250): integer prefilter, exact f32 recheck, hits carry exact scores
  std::vector<Hit> score_threshold(const ScoringMatrix<A>& pssm, const DeviceSequence<A>& seq, float t) const {
    std::vector<lm_hit> buf(1024);
    size_t n = 0;
    for (;;) {
      int st = lm_seq_score_threshold_f32(ctx(), seq.handle(), pssm.matrix().data(), pssm.len(), pssm.matrix().stride(), 0,
                                          seq.rows(), t, buf.data(), buf.size(), &n);
      if (st == LM_ERR_OVERFLOW) { buf.resize(n); continue; }
      check(st, "lm_seq_score_threshold_f32");
      break;
    }
    std::vector<Hit> out;
    for (size_t i = 0; i < n; ++i)
      if (buf[i].position < seq.len()) out.push_back(Hit{(size_t)buf[i].position, buf[i].score});
    return out;
  }

 private:
  lm_ctx* ctx() const { return ctx_.get(); }
  void check(int st, const char* what) const {
    if (st < 0) throw BackendError(st, std::string(what) + ": " + lm_ctx_last_error(ctx()));
  }
  DeviceSequence<A> adopt(lm_seq* raw) const {
    DeviceSequence<A> d;
    std::shared_ptr<lm_ctx> keep = ctx_;
    d.seq_ = std::shared_ptr<lm_seq>(raw, [keep](lm_seq* s) { lm_seq_free(keep.get(), s); });
    int alphabet = 0;
    lm_seq_info(raw, &alphabet, &d.rows_, &d.wrap_, &d.length_);
    return d;
  }
  template <typename T, typename M>
  void score_rows_impl(const M* matrix, size_t m, size_t mstride, const StripedSequence<A>& seq, size_t row_begin,
                       size_t row_end, StripedScores<T>& scores) const {
    const size_t rows = seq.rows();
    // pli/mod.rs:85-88 / avx2.rs:838-844: nothing to score -> resize(0, 0)
    if (seq.len() < m || row_end <= row_begin) {
      if (seq.wrap() + 1 < m) throw BackendError(LM_ERR_WRAP, "not enough wrapping rows for motif");
      scores.resize(0, 0);
      return;
    }
    scores.resize(row_end - row_begin, seq.len() + 1 - m);
    size_t out_rows = 0;
    int st;
    if constexpr (std::is_same<T, float>::value)
      st = lm_score_f32_rows(ctx(), A::ID, matrix, m, mstride, seq.matrix().data(), rows, seq.wrap(), seq.matrix().stride(),
                             seq.len(), row_begin, row_end, scores.matrix_mut().data(), scores.matrix().stride(), &out_rows);
    else
      st = lm_score_u8_rows(ctx(), A::ID, matrix, m, mstride, seq.matrix().data(), rows, seq.wrap(), seq.matrix().stride(),
                            seq.len(), row_begin, row_end, scores.matrix_mut().data(), scores.matrix().stride(), &out_rows);
    check(st, "lm_score_rows");
  }
  template <typename S, typename F>
  std::vector<MatrixCoordinates> threshold_impl(const S& s, F call) const {
    std::vector<lm_hit> buf(256);
    size_t n = 0;
    for (;;) {
      int st = call(buf.data(), buf.size(), &n);
      if (st == LM_ERR_OVERFLOW) { buf.resize(n); continue; }
      check(st, "lm_threshold");
      break;
    }
    std::vector<MatrixCoordinates> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = MatrixCoordinates{(size_t)buf[i].row, (size_t)buf[i].col};
    (void)s;
    return out;
  }

  std::shared_ptr<lm_ctx> ctx_;
  int rule_ = LM_RULE_AVX2;
};

// ------------------------------------------------------------------ Scanner
// Mirror of lightmotif::scan::Scanner (scan.rs:96-250) on a resident sequence.  The reference scans
// blocks of rows with the u8 DiscreteMatrix, thresholds the u8 scores and rescores the survivors in
// f32; here the same filter + exact recheck runs inside one fused kernel over the whole sequence, so
// block_size() is accepted and ignored.  Hits: positions p with p + M <= len whose exact f32 score
// is >= threshold (scan.rs:186-191).  max(): best hit, ties towards the larger position (scan.rs:235).
template <typename A>
class Scanner {
 public:
  Scanner(const Pipeline<A>& pli, const ScoringMatrix<A>& pssm, const DeviceSequence<A>& seq)
      : pli_(pli), pssm_(pssm), seq_(seq) {}
  Scanner& threshold(float t) { threshold_ = t; return *this; }
  Scanner& block_size(size_t b) { block_size_ = b; return *this; }
  std::vector<Hit> collect() const {
    std::vector<Hit> hits = pli_.score_threshold(pssm_, seq_, threshold_);
    hits.erase(std::remove_if(hits.begin(), hits.end(),
                              [&](const Hit& h) { return h.position + pssm_.len() > seq_.len(); }), hits.end());
    std::sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) { return a.position < b.position; });
    return hits;
  }
  std::optional<Hit> max() const {
    Pipeline<A> p = pli_;
    p.set_argmax_rule(LM_RULE_SSE2);  // last column, last row among equal maxima = largest position
    auto best = p.score_argmax(pssm_, seq_);
    if (!best || !(best->score >= threshold_)) return std::nullopt;
    return best;
  }

 private:
  Pipeline<A> pli_;
  const ScoringMatrix<A>& pssm_;
  const DeviceSequence<A>& seq_;
  float threshold_ = 0.0f;
  size_t block_size_ = 256;
};

// ---------------------------------------------------------- ShardedSequence
// A configured StripedSequence cut into row slabs on N GPUs of the box (one process drives them:
// lm_comm_init).  The sharding axis is the row range of Score::score_rows_into (pli/mod.rs:72-77);
// score_argmax / score_threshold return what ONE call over the whole matrix returns (same tie rule,
// positions of the whole matrix).  The same device may be listed more than once (tests).
template <typename A>
class ShardedSequence {
 public:
  ShardedSequence(const StripedSequence<A>& striped, const std::vector<int>& devices) : length_(striped.len()) {
    lm_comm* raw = nullptr;
    int st = lm_comm_init((int)devices.size(), devices.data(), 0, &raw);
    if (st == LM_ERR_UNSUPPORTED) throw UnsupportedBackend();
    if (st < 0) throw BackendError(st, "lm_comm_init");
    comm_ = std::shared_ptr<lm_comm>(raw, lm_comm_destroy);
    slabs_.assign(devices.size(), nullptr);
    st = lm_comm_shard_from_striped(raw, A::ID, striped.matrix().data(), striped.rows(), striped.wrap(),
                                    striped.matrix().stride(), striped.len(), slabs_.data());
    if (st < 0) throw BackendError(st, std::string("lm_comm_shard_from_striped: ") + lm_comm_last_error(raw));
  }
  ~ShardedSequence() {
    for (size_t i = 0; i < slabs_.size(); ++i)
      if (slabs_[i]) lm_seq_free(lm_comm_ctx(comm_.get(), (int)i), slabs_[i]);
  }
  ShardedSequence(const ShardedSequence&) = delete;
  ShardedSequence& operator=(const ShardedSequence&) = delete;

  std::optional<Hit> score_argmax(const ScoringMatrix<A>& pssm, int rule = LM_RULE_AVX2) const {
    lm_hit h;
    int st = lm_seq_score_argmax_f32_sharded(comm_.get(), slabs_.data(), pssm.matrix().data(), pssm.len(),
                                             pssm.matrix().stride(), rule, &h);
    if (st == LM_EMPTY) return std::nullopt;
    if (st < 0) throw BackendError(st, std::string("lm_seq_score_argmax_f32_sharded: ") + lm_comm_last_error(comm_.get()));
    return Hit{(size_t)h.position, h.score};
  }
  // every position p with p + M <= len whose exact f32 score is >= t (Scanner semantics), sorted
  std::vector<Hit> scan(const ScoringMatrix<A>& pssm, float t) const {
    std::vector<lm_hit> buf(4096);
    size_t n = 0;
    for (;;) {
      int st = lm_seq_score_threshold_f32_sharded(comm_.get(), slabs_.data(), pssm.matrix().data(), pssm.len(),
                                                  pssm.matrix().stride(), t, buf.data(), buf.size(), &n, nullptr);
      if (st == LM_ERR_OVERFLOW && n > buf.size()) { buf.resize(n); continue; }
      if (st < 0) throw BackendError(st, std::string("lm_seq_score_threshold_f32_sharded: ") + lm_comm_last_error(comm_.get()));
      break;
    }
    std::vector<Hit> out;
    for (size_t i = 0; i < n; ++i)
      if (buf[i].position + pssm.len() <= length_) out.push_back(Hit{(size_t)buf[i].position, buf[i].score});
    std::sort(out.begin(), out.end(), [](const Hit& a, const Hit& b) { return a.position < b.position; });
    return out;
  }

 private:
  std::shared_ptr<lm_comm> comm_;
  std::vector<lm_seq*> slabs_;
  size_t length_;
};

template <typename A>
void StripedSequence<A>::configure(const Pipeline<A>& pli, const ScoringMatrix<A>& pssm) {
  if (pssm.len() > 0) configure_wrap(pli, pssm.len() - 1);
}
template <typename A>
void StripedSequence<A>::configure_wrap(const Pipeline<A>& pli, size_t m) {
  pli.configure_wrap(*this, m);
}

}  // namespace lightmotif
