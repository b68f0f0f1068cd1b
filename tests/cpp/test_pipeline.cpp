// test_pipeline.cpp -- the reference's own pipeline tests replayed through the
// C++ host mirror (include/lightmotif_b200.hpp) on the CUDA back-end:
//   lightmotif/tests/dna.rs     test_score_rows, test_score, test_score_discrete, test_argmax,
//                               test_threshold, test_score_position
//   lightmotif/tests/encode.rs  exact codes, InvalidSymbol('.')
//   lightmotif/tests/stripe.rs  layout + padding;  seq.rs:509-541 literal matrices and wrap rows
//   README.md:55-91             scores[0] == -23.07094, argmax 18, threshold(10.0) == []
// Needs a B200 (sm_100) device: there is no CPU fall-back.  Exit code 0 = all passed.
#include <cstdio>
#include <cmath>
#include <string>
#include <vector>

#include "lightmotif_b200.hpp"

using namespace lightmotif;

static int failures = 0, checks = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    ++checks;                                                                \
    if (!(cond)) { ++failures; std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

// lightmotif/tests/dna.rs:19-38
static const char* SEQUENCE = "ATGTCCCAACAACGATACCCCGAGCCCATCGCCGTCATCGGCTCGGCATGCAGATTCCCAGGCG";
static const char* PATTERNS[] = {"GTTGACCTTATCAAC", "GTTGATCCAGTCAAC"};
static const float EXPECTED[] = {
    -23.07094,  -18.678621, -15.219191, -17.745737, -18.678621, -23.07094,  -17.745737, -19.611507, -27.463257, -29.989803,
    -14.286304, -26.53037,  -15.219191, -10.826873, -10.826873, -22.138054, -38.774437, -30.922688, -5.50167,   -24.003826,
    -18.678621, -15.219191, -35.315006, -17.745737, -10.826873, -30.922688, -23.07094,  -6.4345555, -31.855574, -23.07094,
    -15.219191, -31.855574, -8.961102,  -26.53037,  -27.463257, -14.286304, -15.219191, -26.53037,  -23.07094,  -18.678621,
    -14.286304, -18.678621, -26.53037,  -16.152077, -17.745737, -18.678621, -17.745737, -14.286304, -30.922688, -18.678621};
static const size_t N_EXPECTED = sizeof(EXPECTED) / sizeof(EXPECTED[0]);

static ScoringMatrix<Dna> build_pssm(const Pipeline<Dna>& pli) {
  std::vector<EncodedSequence<Dna>> pats;
  for (const char* p : PATTERNS) pats.push_back(EncodedSequence<Dna>::encode(pli, p));
  auto cm = CountMatrix<Dna>::from_sequences(pats);
  return cm.to_freq(0.1f).to_weight().to_scoring();  // dna.rs:48-51
}

static void test_encode(const Pipeline<Dna>& pli) {  // tests/encode.rs:10-26
  auto enc = pli.encode(SEQUENCE);
  CHECK(enc.to_string() == SEQUENCE);
  const std::string codes = "ACTGN";
  for (size_t i = 0; i < enc.len(); ++i) CHECK(enc[i] == codes.find(SEQUENCE[i]));
  bool thrown = false;
  try {
    pli.encode("ATGTCCCAACAACGATACCNN..................NNNNNNNNATGCAGATTCCCAGGCG");
  } catch (const InvalidSymbol& e) {
    thrown = true;
    CHECK(e.symbol == '.');
  }
  CHECK(thrown);
  auto lossy = pli.encode_lossy("ATGX.acgt");  // seq.rs:122-129 (lower case is not in the alphabet)
  CHECK(lossy.to_string() == "ATGNNNNNN");
}

static void test_stripe(const Pipeline<Dna>& pli) {  // tests/stripe.rs:17-45
  std::string s1;
  for (int i = 0; i < 16; ++i) s1 += SEQUENCE;
  s1 += "TTATTAT";  // 1031 symbols
  for (const std::string& text : {std::string(SEQUENCE), s1}) {
    auto enc = pli.encode(text);
    auto striped = pli.stripe(enc);
    const size_t rows = striped.rows();
    CHECK(rows == (text.size() + 31) / 32);
    for (size_t i = 0; i < enc.len(); ++i) CHECK(striped.matrix()[i % rows][i / rows] == enc[i]);
    for (size_t i = enc.len(); i < rows * 32; ++i) CHECK(striped.matrix()[i % rows][i / rows] == Dna::DEFAULT);
  }
  // seq.rs:509-541 restated for C = 32: configure_wrap appends row i shifted left by one column
  auto striped = pli.stripe(pli.encode(s1));
  const size_t rows = striped.rows();
  striped.configure_wrap(pli, 3);
  CHECK(striped.wrap() == 3 && striped.rows() == rows);
  for (size_t i = 0; i < 3; ++i) {
    for (size_t c = 0; c + 1 < 32; ++c) CHECK(striped.matrix()[rows + i][c] == striped.matrix()[i][c + 1]);
    CHECK(striped.matrix()[rows + i][31] == Dna::DEFAULT);
  }
}

static void test_score_rows(const Pipeline<Dna>& pli) {  // dna.rs:40-63
  auto striped = pli.stripe(pli.encode(SEQUENCE));
  auto pssm = build_pssm(pli);
  striped.configure(pli, pssm);
  StripedScores<float> scores;
  pli.score_rows_into(pssm, striped, 0, 2, scores);
  CHECK(scores.matrix().rows() == 2);
  CHECK(scores.matrix()[0][0] == EXPECTED[0]);
  CHECK(scores.matrix()[1][0] == EXPECTED[1]);
  pli.score_rows_into(pssm, striped, 1, 2, scores);
  CHECK(scores.matrix().rows() == 1);
  CHECK(scores.matrix()[0][0] == EXPECTED[1]);
  pli.score_rows_into(pssm, striped, 1, 1, scores);  // pli/mod.rs:603-623 empty range
  CHECK(scores.matrix().rows() == 0 && scores.max_index() == 0);
}

static void test_score(const Pipeline<Dna>& pli) {  // dna.rs:65-91 + README.md:81
  auto striped = pli.stripe(pli.encode(SEQUENCE));
  auto pssm = build_pssm(pli);
  striped.configure(pli, pssm);
  auto scores = pssm.score(pli, striped).unstripe();
  CHECK(scores.size() == N_EXPECTED);
  for (size_t i = 0; i < scores.size() && i < N_EXPECTED; ++i) CHECK(std::fabs(scores[i] - EXPECTED[i]) < 1e-5f);
  CHECK(scores[0] == -23.07094f);
}

static void test_score_discrete(const Pipeline<Dna>& pli) {  // dna.rs:93-120
  auto striped = pli.stripe(pli.encode(SEQUENCE));
  auto pssm = build_pssm(pli);
  auto dm = pssm.to_discrete();
  striped.configure(pli, pssm);
  auto scores = pli.score(dm, striped).unstripe();
  CHECK(scores.size() == N_EXPECTED);
  for (size_t i = 0; i < scores.size() && i < N_EXPECTED; ++i) CHECK(dm.unscale(scores[i]) >= EXPECTED[i]);
}

static void test_argmax_threshold(const Pipeline<Dna>& pli) {  // dna.rs:122-173, README.md:84-90
  auto striped = pli.stripe(pli.encode(SEQUENCE));
  auto pssm = build_pssm(pli);
  striped.configure(pli, pssm);
  auto result = pli.score(pssm, striped);
  auto best = pli.argmax(result);
  CHECK(best.has_value() && result.offset(*best) == 18);
  CHECK(result.argmax(pli) == std::optional<size_t>(18));
  auto mx = pli.max(result);
  CHECK(mx.has_value() && std::fabs(*mx - (-5.50167f)) < 1e-5f);
  auto sorted = [](std::vector<size_t> v) { std::sort(v.begin(), v.end()); return v; };
  CHECK(sorted(result.threshold(pli, -10.0f)) == (std::vector<size_t>{18, 27, 32}));
  CHECK(sorted(result.threshold(pli, -15.0f)) == (std::vector<size_t>{10, 13, 14, 18, 24, 27, 32, 35, 40, 47}));
  CHECK(result.threshold(pli, 10.0f).empty());
  StripedScores<float> empty;
  CHECK(!pli.argmax(empty).has_value());  // pli/mod.rs:136-138
}

static void test_score_position(const Pipeline<Dna>& pli) {  // dna.rs:175-199
  auto striped = pli.stripe(pli.encode(SEQUENCE));
  auto pssm = build_pssm(pli);
  striped.configure(pli, pssm);
  auto scores = pli.score(pssm, striped).unstripe();
  for (size_t i = 0; i < N_EXPECTED; ++i) CHECK(pssm.score_position(striped, i) == scores[i]);
}

static void test_resident(const Pipeline<Dna>& pli) {  // fused calls on a resident sequence
  auto pssm = build_pssm(pli);
  auto dev = pli.upload_ascii(SEQUENCE, pssm.len() - 1);
  CHECK(dev.len() == 64 && dev.rows() == 2 && dev.wrap() == 14);
  auto best = pli.score_argmax(pssm, dev);
  CHECK(best.has_value() && best->position == 18 && std::fabs(best->score + 5.50167f) < 1e-5f);  // scan.rs:317-333
  auto hits = pli.score_threshold(pssm, dev, -10.0f);
  std::vector<size_t> pos;
  for (auto& h : hits) pos.push_back(h.position);
  std::sort(pos.begin(), pos.end());
  CHECK(pos == (std::vector<size_t>{18, 27, 32}));  // scan.rs:279-315
  // scan.rs:279-333: Scanner hits 18 / 27 / 32 at threshold -10, max -5.50167 @ 18
  Scanner<Dna> scanner(pli, pssm, dev);
  auto sh = scanner.threshold(-10.0f).block_size(16).collect();
  CHECK(sh.size() == 3 && sh[0].position == 18 && sh[1].position == 27 && sh[2].position == 32);
  auto smax = scanner.max();
  CHECK(smax.has_value() && smax->position == 18 && std::fabs(smax->score + 5.50167f) < 1e-5f);
  CHECK(!Scanner<Dna>(pli, pssm, dev).threshold(0.0f).max().has_value());
  bool thrown = false;
  try {
    pli.upload_ascii("ATG.", 3);
  } catch (const InvalidSymbol& e) {
    thrown = e.symbol == '.';
  }
  CHECK(thrown);
  // not enough wrap rows: the reference panics (avx2.rs:832-837)
  auto striped = pli.stripe(pli.encode(SEQUENCE));
  thrown = false;
  try {
    pli.score(pssm, striped);
  } catch (const BackendError& e) {
    thrown = e.status == LM_ERR_WRAP;
  }
  CHECK(thrown);
}

// the same sequence cut into row slabs on a communicator (one process, the device listed twice so
// that one GPU suffices): whole-matrix positions, same winner, same hits (pli/mod.rs:72-77 row axis)
static void test_sharded(const Pipeline<Dna>& pli) {
  auto pssm = build_pssm(pli);
  std::string text;
  for (int i = 0; i < 300; ++i) text += SEQUENCE;  // 19 200 symbols = 600 rows
  auto striped = pli.stripe(pli.encode(text));
  striped.configure(pli, pssm);
  auto dev = pli.upload_ascii(text, pssm.len() - 1);
  auto one = pli.score_argmax(pssm, dev);
  ShardedSequence<Dna> sharded(striped, {0, 0, 0});
  auto best = sharded.score_argmax(pssm);
  CHECK(one.has_value() && best.has_value() && best->position == one->position && best->score == one->score);
  Scanner<Dna> scanner(pli, pssm, dev);
  auto want = scanner.threshold(-10.0f).collect();
  auto got = sharded.scan(pssm, -10.0f);
  CHECK(got.size() == want.size() && got.size() >= 900);
  bool same = got.size() == want.size();
  for (size_t i = 0; same && i < got.size(); ++i) same = got[i].position == want[i].position && got[i].score == want[i].score;
  CHECK(same);
}

// --host: no GPU needed.  Exercises the host-side value types (matrix copies, alignment,
// CountMatrix -> ScoringMatrix -> DiscreteMatrix) and prints the PSSM bit patterns so that the
// pytest wrapper can compare them with the oracle's restatement of pwm/mod.rs.
static int host_only() {
  auto enc = [](const char* t) {
    std::vector<uint8_t> v;
    for (const char* p = t; *p; ++p) v.push_back((uint8_t)std::string("ACTGN").find(*p));
    return EncodedSequence<Dna>(v);
  };
  auto cm = CountMatrix<Dna>::from_sequences({enc(PATTERNS[0]), enc(PATTERNS[1])});
  auto pssm = cm.to_freq(0.1f).to_weight().to_scoring();
  auto direct = cm.to_freq(0.1f).to_scoring();
  CHECK(pssm.len() == 15 && pssm.matrix().stride() == 8);
  CHECK(reinterpret_cast<uintptr_t>(pssm.matrix().data()) % 32 == 0);
  ScoringMatrix<Dna> copy = pssm;  // deep copy keeps alignment and contents
  CHECK(reinterpret_cast<uintptr_t>(copy.matrix().data()) % 32 == 0 && copy.matrix().data() != pssm.matrix().data());
  for (size_t j = 0; j < pssm.len(); ++j)
    for (int a = 0; a < 5; ++a) {
      uint32_t x, y;
      std::memcpy(&x, &pssm.matrix()[j][a], 4);
      std::memcpy(&y, &copy.matrix()[j][a], 4);
      CHECK(x == y);
      std::printf("%08x%c", x, a == 4 ? '\n' : ' ');
    }
  auto dm = pssm.to_discrete();
  for (size_t j = 0; j < dm.len(); ++j)
    for (int a = 0; a < 5; ++a) std::printf("%u%c", (unsigned)dm.matrix()[j][a], a == 4 ? '\n' : ' ');
  CHECK((DenseMatrix<float, 5>::stride() == 8) && (DenseMatrix<float, 21>::stride() == 24));  // dense.rs:366-391
  CHECK((DenseMatrix<uint8_t, 32>::stride() == 32) && (DenseMatrix<float, 32>::stride() == 32));
  bool thrown = false;
  try {
    CountMatrix<Dna>::from_sequences({enc("TTAT"), enc("TCA")});  // pwm/mod.rs:195-208
  } catch (const InvalidData&) {
    thrown = true;
  }
  CHECK(thrown);
  (void)direct;
  std::printf("%d checks, %d failures\n", checks, failures);
  return failures ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--host") return host_only();
  try {
    auto pli = Pipeline<Dna>::cuda();
    test_encode(pli);
    test_stripe(pli);
    test_score_rows(pli);
    test_score(pli);
    test_score_discrete(pli);
    test_argmax_threshold(pli);
    test_score_position(pli);
    test_resident(pli);
    test_sharded(pli);
  } catch (const UnsupportedBackend& e) {
    std::printf("UNSUPPORTED: %s\n", e.what());
    return 77;
  } catch (const std::exception& e) {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("%d checks, %d failures\n", checks, failures);
  return failures ? 1 : 0;
}
