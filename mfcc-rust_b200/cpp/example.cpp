// The reference's own unit tests (speechsauce/src/lib.rs:41-134) re-expressed against the C++
// mirror.  Build: make -C mfcc-rust_b200 cpp_example ; needs a CUDA device to run.
#include <cstdio>
#include <random>

#include "speechsauce.hpp"

using namespace speechsauce;

static int fails = 0;
#define EXPECT(c)                                                  \
  do {                                                             \
    if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } \
  } while (0)

int main() {
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 0.1f);  // create_signal(), lib.rs:18-22
  std::vector<float> sig(1000000);
  for (auto& v : sig) v = nd(rng);
  View1 s{sig.data(), sig.size()};
  try {
    auto cfg = config::SpeechConfigBuilder(16000).build();  // default_config, lib.rs:28-30
    // test_preemphasis
    EXPECT(processing::preemphasis(s, 1, 0.98f).len() == sig.size());
    // test_stack_frames: zero padding, ceil((N - L) / S) rows
    auto fr = processing::stack_frames(s, 16000, 0.02f, 0.02f, true);
    EXPECT(fr.rows == (size_t)std::ceil((1000000.f - 320.f) / 320.f) && fr.cols == 320);
    // test_cmvn
    std::uniform_real_distribution<float> ud(0.f, 1.f);
    Array2 fv(50, 100);
    for (auto& v : fv.v) v = ud(rng);
    auto nv = processing::cmvn(fv.view(), true);
    for (size_t c = 0; c < 100; ++c) {
      double m = 0, q = 0;
      for (size_t r = 0; r < 50; ++r) m += nv(r, c);
      m /= 50;
      for (size_t r = 0; r < 50; ++r) q += (nv(r, c) - m) * (nv(r, c) - m);
      EXPECT(std::fabs(m) < 1e-5 && std::fabs(std::sqrt(q / 50) - 1.0) < 1e-4);
    }
    // test_mfcc / test_mfe
    auto m = feature::mfcc(s, cfg);
    EXPECT(m.rows == 6248 && m.cols == 13);
    for (float v : m.v) EXPECT(!std::isnan(v));
    auto fe = feature::mfe(s, cfg);
    EXPECT(fe.first.rows == 6248 && fe.first.cols == 40 && fe.second.len() == 6248);
    // errors are exceptions (the reference panics): default config has H = 320 > W/2
    bool threw = false;
    try { functions::stft2(View2{sig.data(), 1, 4000}, cfg); } catch (const Error& e) { threw = e.code == SS_ERR_INVALID_ARG; }
    EXPECT(threw);
    auto c4 = config::SpeechConfig(16000, 400, 0.01f, 0.01f, 13, 80, 0.f, 8000.f, true);
    auto mel = feature::mel_spectrogram2(View2{sig.data(), 4, 16000}, c4);
    EXPECT(mel.d0 == 4 && mel.d1 == 80 && mel.d2 == 100);
  } catch (const Error& e) {
    std::printf("speechsauce error: %s\n", e.what());
    return 2;
  }
  std::printf(fails ? "cpp example: %d failures\n" : "cpp example: all reference-shaped tests passed\n", fails);
  return fails ? 1 : 0;
}
