// Host-side tables and derived sizes.  Pure C++ (no CUDA calls) so that the host logic is
// testable on a machine without a GPU.  Each routine cites the reference lines it follows.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>

#include "ss_internal.h"

namespace ss {

int64_t rust_round_f32(float x) {
  // f32::round: half away from zero (processing.rs:77-78)
  float a = std::floor(std::fabs(x) + 0.5f);
  return (int64_t)(x < 0 ? -a : a);
}

int64_t num_frames(int64_t n, int64_t L, int64_t S, int zero_padding) {
  // processing.rs:91-92 / 101-102.  `length_signal - frame_sample_length` is a usize
  // subtraction (underflow = panic), the quotient is computed in f32.
  if (L <= 0 || S <= 0 || n < L) return -1;
  float q = (float)(n - L) / (float)S;
  int64_t t = zero_padding ? (int64_t)std::ceil(q) : (int64_t)std::floor(q);
  if (!zero_padding && t < 1) return -1;  // (numframes - 1) underflows, processing.rs:105
  return t;
}

int64_t stft_num_frames(int64_t ttd, int64_t H) {
  if (H <= 0 || ttd < 0) return -1;
  return (int64_t)std::ceil((float)ttd / (float)H);  // functions.rs:97
}

std::string derive(const ss_config& c, Derived* d) {
  d->cfg = c;
  if (c.sample_rate <= 0) return "sample_rate must be positive";
  if (c.fft_points < 2) return "fft_points must be >= 2";
  if (c.num_filters < 1 || c.num_filters > 1024) return "num_filters out of range [1, 1024]";
  if (c.num_cepstral < 0 || c.num_cepstral > c.num_filters)
    return "num_cepstral must be in [0, num_filters] (slice ..num_cepstral panics otherwise, feature.rs:133)";
  const float sr = (float)c.sample_rate;
  float hi = c.high_frequency > 0.f ? c.high_frequency : sr / 2.0f;  // config.rs:44
  d->cfg.high_frequency = hi;
  if (!(hi <= sr / 2.0f))
    return "High frequency cannot be greater than half of the sampling frequency! (feature.rs:47-50)";
  if (!(c.low_frequency >= 0.f)) return "low frequency cannot be less than zero! (feature.rs:51)";
  d->W = c.fft_points;
  d->F = c.fft_points / 2 + 1;                              // config.rs:179
  d->L = rust_round_f32(sr * c.frame_length);               // processing.rs:77
  d->S = rust_round_f32(sr * c.frame_stride);               // processing.rs:78
  float hs = c.frame_length * sr;                           // config.rs:154
  d->H = hs > 0.f ? (int64_t)hs : 0;                        // `as usize` saturates at 0
  d->nf = (int)c.num_filters;
  d->nc = (int)c.num_cepstral;
  d->dct2_scale = c.dct2_scale != 0.f ? c.dct2_scale : 2.0f;
  d->stft_ok = d->H > 0 && 2 * d->H <= d->W;                // functions.rs:136 underflows otherwise
  d->speechpy_ok = d->L > 0 && d->L <= d->W && d->S > 0;    // ndrustfft panics when L > n_fft
  d->n_pad = d->stft_ok ? d->W / d->H - 1 : -1;             // functions.rs:96
  // config.rs:178: 1.0 / (fft_points.pow(2) as f32 / (2 * frame_size) as f32)
  d->wnorm = d->H > 0 ? 1.0f / ((float)(d->W * d->W) / (float)(2 * d->H)) : 0.f;
  d->librosa = (c.flags & SS_FLAG_LIBROSA) != 0;
  if (d->librosa) {  // librosa.stft: any hop <= n_fft, no dropped / zero frames, no normalisation
    if (c.fft_points % 2) return "SS_FLAG_LIBROSA needs an even fft_points";
    d->stft_ok = d->H > 0 && d->H <= d->W;
    d->speechpy_ok = false;
    d->n_pad = 0;
    d->wnorm = 1.0f;
  }
  return "";
}

static inline float hz_to_mel(float f) { return 1127.0f * std::log(1.0f + f / 700.0f); }      // functions.rs:19-21
static inline float mel_to_hz(float m) { return 700.0f * (std::exp(m / 1127.0f) - 1.0f); }    // functions.rs:31-33

// Slaney mel scale (librosa.hz_to_mel / mel_to_hz with htk=False): linear below 1 kHz, log above
static inline double slaney_hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static inline double slaney_mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm="slaney"), computed in f64 and
// stored as f32 like librosa does, in the banded form the kernels use
static void build_melbank_slaney(const Derived& d, MelBank* mb) {
  const int nf = d.nf;
  const double sr = (double)d.cfg.sample_rate;
  const double m0 = slaney_hz_to_mel(d.cfg.low_frequency), m1 = slaney_hz_to_mel(d.cfg.high_frequency);
  std::vector<double> fpt(nf + 2);
  for (int i = 0; i < nf + 2; ++i) fpt[i] = slaney_mel_to_hz(m0 + (m1 - m0) * (double)i / (double)(nf + 1));
  mb->idx.assign(nf + 2, 0);
  mb->start.assign(nf, 0);
  mb->len.assign(nf, 0);
  mb->wofs.assign(nf, 0);
  mb->w.clear();
  mb->max_len = 0;
  mb->max_bin = -1;
  for (int i = 0; i < nf; ++i) {
    const double enorm = 2.0 / (fpt[i + 2] - fpt[i]);
    int lo = -1, hi = -1;
    std::vector<float> w((size_t)d.F, 0.f);
    for (int64_t k = 0; k < d.F; ++k) {
      const double f = (double)k * sr / (double)d.W;
      const double lower = (f - fpt[i]) / (fpt[i + 1] - fpt[i]), upper = (fpt[i + 2] - f) / (fpt[i + 2] - fpt[i + 1]);
      const double v = std::max(0.0, std::min(lower, upper)) * enorm;
      w[k] = (float)v;
      if (w[k] != 0.f) {
        if (lo < 0) lo = (int)k;
        hi = (int)k;
      }
    }
    mb->wofs[i] = (int)mb->w.size();
    if (lo >= 0) {
      mb->start[i] = lo;
      mb->len[i] = hi - lo + 1;
      for (int k = lo; k <= hi; ++k) mb->w.push_back(w[k]);
      mb->max_len = std::max(mb->max_len, mb->len[i]);
      mb->max_bin = std::max(mb->max_bin, hi);
    }
  }
  if (mb->w.empty()) mb->w.push_back(0.f);
}

void build_melbank(const Derived& d, MelBank* mb) {
  if (d.librosa) return build_melbank_slaney(d, mb);
  // feature.rs:43-88 with functions::triangle (functions.rs:43-60)
  const int nf = d.nf;
  const float sr = (float)d.cfg.sample_rate;
  const float m0 = hz_to_mel(d.cfg.low_frequency), m1 = hz_to_mel(d.cfg.high_frequency);
  const int n = nf + 2;
  const float step = n > 1 ? (m1 - m0) / (float)(n - 1) : 0.f;  // ndarray linspace: start + step*i
  mb->idx.resize(n);
  for (int i = 0; i < n; ++i) {
    float mel = m0 + step * (float)i;
    float hz = mel_to_hz(mel);
    float v = (float)(d.F + 1) * hz / sr;  // feature.rs:70
    int64_t b = v > 0.f ? (int64_t)v : 0;
    mb->idx[i] = b;
  }
  mb->start.assign(nf, 0);
  mb->len.assign(nf, 0);
  mb->wofs.assign(nf, 0);
  mb->w.clear();
  mb->max_len = 0;
  mb->max_bin = -1;
  for (int i = 0; i < nf; ++i) {
    const float left = (float)mb->idx[i], middle = (float)mb->idx[i + 1], right = (float)mb->idx[i + 2];
    int lo = -1, hi = -1;
    std::vector<float> w;
    for (int64_t b = mb->idx[i]; b <= mb->idx[i + 2] && b < d.F; ++b) {
      const float x = (float)b;
      float v = 0.f;
      if (left <= x && x < right) {
        if (x <= middle) v = (x - left) / (middle - left);
        if (middle <= x) v = (right - x) / (right - middle);
      }
      w.push_back(v);
    }
    for (int k = 0; k < (int)w.size(); ++k)
      if (w[k] != 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    mb->wofs[i] = (int)mb->w.size();
    if (lo >= 0) {
      mb->start[i] = (int)mb->idx[i] + lo;
      mb->len[i] = hi - lo + 1;
      for (int k = lo; k <= hi; ++k) mb->w.push_back(w[k]);
      mb->max_len = std::max(mb->max_len, mb->len[i]);
      mb->max_bin = std::max(mb->max_bin, mb->start[i] + mb->len[i] - 1);
    }
  }
  if (mb->w.empty()) mb->w.push_back(0.f);
}

void build_window(const Derived& d, std::vector<float>* out) {
  // config.rs:151-160: sin(pi/2 * sin^2(pi/2 * (i + 0.5) / (W/2))) in f64, stored as f32
  const double pi = 3.14159265358979323846264338327950288;
  const int64_t W = d.W, Wh = d.W / 2;
  out->resize(W);
  if (d.librosa) {  // scipy.signal.get_window("hann", W, fftbins=True): periodic Hann
    for (int64_t i = 0; i < W; ++i) (*out)[i] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * (double)i / (double)W));
    return;
  }
  for (int64_t i = 0; i < W; ++i) {
    double s = std::sin(0.5 * pi * ((double)i + 0.5) / (double)Wh);
    (*out)[i] = (float)std::sin(0.5 * pi * s * s);
  }
}

void build_dct(const Derived& d, std::vector<float>* out) {
  // nddct2 (feature.rs:120-123): y[k] = dct2_scale * sum_n x[n] cos(pi k (2n+1) / (2 nf))
  const double pi = 3.14159265358979323846264338327950288;
  out->resize((size_t)d.nc * d.nf);
  for (int k = 0; k < d.nc; ++k)
    for (int n = 0; n < d.nf; ++n)
      (*out)[(size_t)k * d.nf + n] =
          (float)((double)d.dct2_scale * std::cos(pi * (double)k * (2.0 * n + 1.0) / (2.0 * d.nf)));
}

bool build_fft_plan(int n, FftPlanHost* p) {
  // real FFT of even length n = complex FFT of length n/2 + post-processing; odd n: full length.
  // The stage list factorises that complex length; the twiddle table is W_n^q for q < n.
  p->n = n;
  p->n_stages = 0;
  int m = (n % 2 == 0) ? n / 2 : n;
  if (m == 1) { p->radix[p->n_stages++] = 1; }
  auto push = [&](int r) {
    if (p->n_stages >= kMaxStages) return false;
    p->radix[p->n_stages++] = r;
    return true;
  };
  while (m % 4 == 0) { if (!push(4)) return false; m /= 4; }
  while (m % 2 == 0) { if (!push(2)) return false; m /= 2; }
  for (int r = 3; r <= 61 && m > 1; r += 2)
    while (m % r == 0) { if (!push(r)) return false; m /= r; }
  if (m != 1) return false;  // a prime factor above 61
  const double pi = 3.14159265358979323846264338327950288;
  p->tw.resize((size_t)2 * n);
  for (int q = 0; q < n; ++q) {
    double a = -2.0 * pi * (double)q / (double)n;
    p->tw[2 * q] = (float)std::cos(a);
    p->tw[2 * q + 1] = (float)std::sin(a);
  }
  return true;
}

void balance_filters(const MelBank& mb, int lanes, std::vector<std::vector<int>>* per_lane) {
  per_lane->assign(lanes, {});
  std::vector<int> order(mb.len.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return mb.len[a] > mb.len[b]; });
  std::vector<int> load(lanes, 0);
  for (int f : order) {
    int best = 0;
    for (int l = 1; l < lanes; ++l)
      if (load[l] + 2 * (int)(*per_lane)[l].size() < load[best] + 2 * (int)(*per_lane)[best].size()) best = l;
    (*per_lane)[best].push_back(f);
    load[best] += mb.len[f];
  }
}

}  // namespace ss
