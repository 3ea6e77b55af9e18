// speechsauce.hpp — header-only C++17 mirror of the `speechsauce` crate's hot-path API
// (secretsauceai/mfcc-rust) over the C ABI of include/speechsauce_cuda.h.
//
// The reference is a compiled Rust library; no Rust toolchain exists in this image, so this is
// the host-side mirror that IS compiled and run here.  Module / function names, argument order
// and error behaviour follow the reference (every error there is a panic; here: ss::Error):
//   speechsauce::config::{SpeechConfig, SpeechConfigBuilder}          config.rs:10-190
//   speechsauce::feature::{mfcc, mfe, lmfe, ssc, mel_spectrogram1/2}  feature.rs:99-245
//   speechsauce::processing::{preemphasis, stack_frames, power_spectrum, cmvn}  processing.rs:31-300
//   speechsauce::functions::{stft1, stft2}                            functions.rs:86-233
// There is no CPU fallback.
#pragma once
#include <cmath>
#include <complex>
#include <cstring>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "speechsauce_cuda.h"

namespace speechsauce {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error("[ss_status " + std::to_string(c) + "] " + m), code(c) {}
};
inline void check(int code) {
  if (code != 0) throw Error(code, ss_last_error());
}

// minimal stand-ins for ndarray's views / owned arrays (row-major, contiguous)
struct View1 { const float* data; size_t len; };
struct View2 { const float* data; size_t rows, cols; };
struct Array1 { std::vector<float> v; size_t len() const { return v.size(); } float& operator[](size_t i) { return v[i]; } };
struct Array2 {
  std::vector<float> v; size_t rows = 0, cols = 0;
  Array2() = default;
  Array2(size_t r, size_t c) : v(r * c), rows(r), cols(c) {}
  float& operator()(size_t i, size_t j) { return v[i * cols + j]; }
  View2 view() const { return View2{v.data(), rows, cols}; }
};
struct Array3 {
  std::vector<float> v; size_t d0 = 0, d1 = 0, d2 = 0;
  Array3() = default;
  Array3(size_t a, size_t b, size_t c) : v(a * b * c), d0(a), d1(b), d2(c) {}
  float& operator()(size_t i, size_t j, size_t k) { return v[(i * d1 + j) * d2 + k]; }
};
struct CArray3 {
  std::vector<std::complex<float>> v; size_t d0 = 0, d1 = 0, d2 = 0;
  CArray3(size_t a, size_t b, size_t c) : v(a * b * c), d0(a), d1(b), d2(c) {}
  std::complex<float>& operator()(size_t i, size_t j, size_t k) { return v[(i * d1 + j) * d2 + k]; }
};

namespace config {

// config.rs:99-185.  Move-only: like the reference type it is Send but not Sync.
class SpeechConfig {
 public:
  size_t sample_rate, window_size, window_size_half;
  float frame_length, frame_stride;
  size_t num_cepstral, num_filters;
  float low_frequency, high_frequency;
  size_t freq_size, frame_size;
  bool dc_elimination;
  float wnorm;
  size_t frame_len, frame_step;  // stack_frames' own rounding (processing.rs:77-78)

  // config.rs:140-150 — the same nine arguments in the same order
  SpeechConfig(size_t sample_rate_, size_t fft_points, float frame_length_, float frame_stride_, size_t num_cepstral_,
               size_t num_filters_, float low_frequency_, float high_frequency_, bool dc_elimination_, uint32_t flags = 0)
      : sample_rate(sample_rate_), window_size(fft_points), window_size_half(fft_points / 2), frame_length(frame_length_),
        frame_stride(frame_stride_), num_cepstral(num_cepstral_), num_filters(num_filters_), low_frequency(low_frequency_),
        high_frequency(high_frequency_), dc_elimination(dc_elimination_) {
    ss_config c{(int64_t)sample_rate, (int64_t)fft_points, frame_length, frame_stride, (int64_t)num_cepstral,
                (int64_t)num_filters, low_frequency, high_frequency, dc_elimination ? 1 : 0, flags, 0.0f};
    check(ss_plan_create(&c, nullptr, -1, &plan_));  // every visible GPU
    ss_plan_info info;
    check(ss_plan_get_info(plan_, &info));
    freq_size = (size_t)info.freq_size;
    frame_size = (size_t)info.hop;
    wnorm = info.wnorm;
    frame_len = (size_t)info.frame_len;
    frame_step = (size_t)info.frame_step;
  }
  SpeechConfig(SpeechConfig&& o) noexcept { *this = std::move(o); }
  SpeechConfig& operator=(SpeechConfig&& o) noexcept {
    if (this != &o) {
      if (plan_) ss_plan_destroy(plan_);
      std::memcpy((void*)this, (void*)&o, sizeof(SpeechConfig));
      o.plan_ = nullptr;
    }
    return *this;
  }
  SpeechConfig(const SpeechConfig&) = delete;
  SpeechConfig& operator=(const SpeechConfig&) = delete;
  ~SpeechConfig() { if (plan_) ss_plan_destroy(plan_); }
  ss_plan* plan() const { return plan_; }

 private:
  ss_plan* plan_ = nullptr;
};

// config.rs:34-97
class SpeechConfigBuilder {
 public:
  explicit SpeechConfigBuilder(size_t sample_rate)
      : sr_(sample_rate), fft_(512), fl_(0.02f), fs_(0.01f), nc_(13), nf_(40), lo_(0.f), hi_((float)sample_rate / 2.0f), dc_(true) {}
  SpeechConfigBuilder& high_freq(float v) { hi_ = v; return *this; }
  SpeechConfigBuilder& low_freq(float v) { lo_ = v; return *this; }
  SpeechConfigBuilder& dc_elimination(bool v) { dc_ = v; return *this; }
  SpeechConfigBuilder& num_cepstral(size_t v) { nc_ = v; return *this; }
  SpeechConfigBuilder& num_filters(size_t v) { nf_ = v; return *this; }  // missing in the reference (Q15)
  SpeechConfigBuilder& frame_stride(float v) { fs_ = v; return *this; }
  SpeechConfigBuilder& frame_length(float v) { fl_ = v; return *this; }
  SpeechConfigBuilder& fft_points(size_t v) { fft_ = v; return *this; }
  SpeechConfig build() const { return SpeechConfig(sr_, fft_, fl_, fs_, nc_, nf_, lo_, hi_, dc_); }

 private:
  size_t sr_, fft_;
  float fl_, fs_;
  size_t nc_, nf_;
  float lo_, hi_;
  bool dc_;
};

}  // namespace config

namespace feature {
using config::SpeechConfig;

inline size_t num_frames(const SpeechConfig& cfg, size_t n) {
  const int64_t t = ss_num_frames((int64_t)n, (int64_t)cfg.frame_len, (int64_t)cfg.frame_step, 0);
  if (t < 1) throw Error(SS_ERR_INVALID_ARG, "signal yields no frame (usize underflow in the reference, processing.rs:101-105)");
  return (size_t)t;
}
// feature.rs:99
inline Array2 mfcc(View1 signal, const SpeechConfig& cfg) {
  Array2 out(num_frames(cfg, signal.len), cfg.num_cepstral);
  check(ss_mfcc(cfg.plan(), signal.data, 1, (int64_t)signal.len, (int64_t)signal.len, out.v.data(), SS_CMVN_NONE));
  return out;
}
// batched: one clip per row, sharded over the GPUs; cmvn_mode applies processing::cmvn per clip
inline Array3 mfcc_batch(View2 signals, const SpeechConfig& cfg, int cmvn_mode = SS_CMVN_NONE) {
  Array3 out(signals.rows, num_frames(cfg, signals.cols), cfg.num_cepstral);
  check(ss_mfcc(cfg.plan(), signals.data, (int64_t)signals.rows, (int64_t)signals.cols, (int64_t)signals.cols, out.v.data(), cmvn_mode));
  return out;
}
// feature.rs:200
inline std::pair<Array2, Array1> mfe(View1 signal, const SpeechConfig& cfg) {
  const size_t t = num_frames(cfg, signal.len);
  Array2 feat(t, cfg.num_filters);
  Array1 energy{std::vector<float>(t)};
  check(ss_mfe(cfg.plan(), signal.data, 1, (int64_t)signal.len, (int64_t)signal.len, feat.v.data(), energy.v.data()));
  return {std::move(feat), std::move(energy)};
}
// feature.rs:242 (private in the reference)
inline Array2 lmfe(View1 signal, const SpeechConfig& cfg) {
  Array2 out(num_frames(cfg, signal.len), cfg.num_filters);
  check(ss_lmfe(cfg.plan(), signal.data, 1, (int64_t)signal.len, (int64_t)signal.len, out.v.data()));
  return out;
}
// README.md:15
inline Array2 ssc(View1 signal, const SpeechConfig& cfg) {
  Array2 out(num_frames(cfg, signal.len), cfg.num_filters);
  check(ss_ssc(cfg.plan(), signal.data, 1, (int64_t)signal.len, (int64_t)signal.len, out.v.data()));
  return out;
}
// feature.rs:163 -> [ch][nf][T']
inline Array3 mel_spectrogram2(View2 signal, const SpeechConfig& cfg) {
  const int64_t tp = ss_stft_num_frames((int64_t)signal.cols, (int64_t)cfg.frame_size);
  Array3 out(signal.rows, cfg.num_filters, (size_t)(tp < 0 ? 0 : tp));
  check(ss_mel_spectrogram(cfg.plan(), signal.data, (int64_t)signal.rows, (int64_t)signal.cols, (int64_t)signal.cols, out.v.data(), nullptr, 0));
  return out;
}
// librosa.feature.melspectrogram (Hann, centred reflect-padded frames, Slaney mel bank, power 2) for a config
// created with flags = SS_FLAG_LIBROSA and frame_length = (hop + 0.5) / sample_rate -> [ch][n_mels][1 + ttd / hop]
inline Array3 melspectrogram_librosa(View2 signal, const SpeechConfig& cfg, bool db = false) {
  const int64_t tp = ss_librosa_num_frames((int64_t)signal.cols, (int64_t)cfg.frame_size);
  Array3 out(signal.rows, cfg.num_filters, (size_t)(tp < 0 ? 0 : tp));
  check(ss_librosa_melspectrogram(cfg.plan(), signal.data, (int64_t)signal.rows, (int64_t)signal.cols, (int64_t)signal.cols, out.v.data(), db ? 1 : 0));
  return out;
}
// feature.rs:151 (intended semantics) -> [nf][T']
inline Array2 mel_spectrogram1(View1 signal, const SpeechConfig& cfg) {
  const int64_t tp = ss_stft_num_frames((int64_t)signal.len, (int64_t)cfg.frame_size);
  Array2 out(cfg.num_filters, (size_t)(tp < 0 ? 0 : tp));
  check(ss_mel_spectrogram(cfg.plan(), signal.data, 1, (int64_t)signal.len, (int64_t)signal.len, out.v.data(), nullptr, 0));
  return out;
}
// feature.rs:253 -> [T][C][3] = static, delta, delta-delta
inline Array3 extract_derivative_feature(View2 feat) {
  Array3 out(feat.rows, feat.cols, 3);
  check(ss_extract_derivative_feature(feat.data, (int64_t)feat.rows, (int64_t)feat.cols, out.v.data()));
  return out;
}
}  // namespace feature

namespace processing {
// processing.rs:31
inline Array1 preemphasis(View1 signal, int64_t shift = 1, float cof = 0.98f) {
  Array1 out{std::vector<float>(signal.len)};
  check(ss_preemphasis(signal.data, out.v.data(), (int64_t)signal.len, shift, cof));
  return out;
}
// processing.rs:65 (filter = None)
inline Array2 stack_frames(View1 signal, size_t sample_rate, float frame_length, float frame_stride, bool zero_padding) {
  auto rr = [](float x) { return (int64_t)std::floor(std::fabs(x) + 0.5f); };
  const int64_t L = rr((float)sample_rate * frame_length), S = rr((float)sample_rate * frame_stride);
  int64_t t = 0;
  check(ss_stack_frames(signal.data, (int64_t)signal.len, L, S, zero_padding, 0, nullptr, &t));
  Array2 out((size_t)t, (size_t)L);
  check(ss_stack_frames(signal.data, (int64_t)signal.len, L, S, zero_padding, 0, out.v.data(), &t));
  return out;
}
// processing.rs:179
inline Array2 power_spectrum(View2 frames, size_t fft_points) {
  Array2 out(frames.rows, fft_points / 2 + 1);
  check(ss_power_spectrum(frames.data, (int64_t)frames.rows, (int64_t)frames.cols, (int64_t)fft_points, out.v.data()));
  return out;
}
// processing.rs:194 (private in the reference)
inline Array2 log_power_spectrum(View2 frames, size_t fft_points, bool normalize) {
  Array2 out(frames.rows, fft_points / 2 + 1);
  check(ss_log_power_spectrum(frames.data, (int64_t)frames.rows, (int64_t)frames.cols, (int64_t)fft_points, normalize ? 1 : 0,
                              out.v.data()));
  return out;
}
// processing.rs:265
inline Array2 cmvn(View2 vec, bool variance_normalization) {
  Array2 out(vec.rows, vec.cols);
  check(ss_cmvn(vec.data, (int64_t)vec.rows, (int64_t)vec.cols, variance_normalization, out.v.data()));
  return out;
}
// processing.rs:222 (numpy 'edge' padding along the feature axis)
inline Array2 derivative_extraction(View2 feat, int delta_windows = 2) {
  Array2 out(feat.rows, feat.cols);
  check(ss_derivative_extraction(feat.data, (int64_t)feat.rows, (int64_t)feat.cols, delta_windows, out.v.data()));
  return out;
}
// processing.rs:315 (numpy 'symmetric' padding along time)
inline Array2 cmvnw(View2 vec, size_t win_size = 301, bool variance_normalization = false) {
  Array2 out(vec.rows, vec.cols);
  check(ss_cmvnw(vec.data, (int64_t)vec.rows, (int64_t)vec.cols, (int64_t)win_size, variance_normalization, out.v.data()));
  return out;
}
}  // namespace processing

namespace functions {
using config::SpeechConfig;
// functions.rs:86 -> [ch][T'][F]
inline CArray3 stft2(View2 input, const SpeechConfig& cfg) {
  const int64_t tp = ss_stft_num_frames((int64_t)input.cols, (int64_t)cfg.frame_size);
  CArray3 out(input.rows, (size_t)(tp < 0 ? 0 : tp), cfg.freq_size);
  check(ss_stft(cfg.plan(), input.data, (int64_t)input.rows, (int64_t)input.cols, (int64_t)input.cols,
                reinterpret_cast<float*>(out.v.data()), nullptr));
  return out;
}
// functions.rs:199
inline CArray3 stft1(View1 input, const SpeechConfig& cfg) { return stft2(View2{input.data, 1, input.len}, cfg); }
}  // namespace functions

}  // namespace speechsauce
