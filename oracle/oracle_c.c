/* C restatement of speechsauce's CPU hot path — the timed CPU baseline ("port").
 *
 * TEST / BENCH INFRASTRUCTURE ONLY: nothing under mfcc-rust_b200/ links or calls this file.
 * PARITY UNPINNED by the reference (no golden vectors, Rust toolchain absent): this port is
 * checked against oracle/speechsauce_oracle.py (tests/test_oracle_c.py), which carries the
 * line-by-line citations.  It keeps the reference's ALGORITHMIC structure so that its timing is
 * a fair stand-in for the crate's CPU path:
 *   stack_frames gather (processing.rs:65-129, intended semantics) -> zero pad -> R2C FFT
 *   (processing.rs:143-169) -> |X|/N (:179-181) -> row sums (feature.rs:216) -> DENSE
 *   P x fb^T GEMM (feature.rs:229) -> ln (:105) -> full-length DCT-II then slice (:120-133)
 *   -> cmvn (processing.rs:265-300);  stft2 + |X|^2 + dense einsum (functions.rs:86-170,
 *   feature.rs:163-174).
 * One plan per thread (SpeechConfig is !Sync, config.rs:126,130); OpenMP `parallel for` over
 * clips stands in for "rayon over clips".  Deliberate deviations, both in the reference's
 * favour: FFT/DCT plans are built once per thread rather than once per call (Q5), and the 8
 * full-size temporaries of the reference are not re-allocated per call.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EPS 1.1920929e-07f
#define ORC_CMVN_EPS 9.31322574615478515625e-10f

typedef struct { float re, im; } cpx;

/* ---------------------------------------------------------------- mixed-radix complex FFT */
typedef struct {
  int n;
  int nfac, fac[32];
  cpx* tw;      /* e^{-2 pi i q / n} */
  cpx* scratch; /* n */
} fft_plan;

static int fft_plan_init(fft_plan* p, int n) {
  p->n = n;
  p->nfac = 0;
  int m = n;
  while (m % 4 == 0) { p->fac[p->nfac++] = 4; m /= 4; }
  while (m % 2 == 0) { p->fac[p->nfac++] = 2; m /= 2; }
  for (int r = 3; m > 1; r += 2)
    while (m % r == 0) { if (p->nfac >= 32) return -1; p->fac[p->nfac++] = r; m /= r; }
  p->tw = (cpx*)malloc(sizeof(cpx) * n);
  p->scratch = (cpx*)malloc(sizeof(cpx) * n);
  const double pi = 3.14159265358979323846;
  for (int q = 0; q < n; ++q) {
    p->tw[q].re = (float)cos(-2.0 * pi * q / n);
    p->tw[q].im = (float)sin(-2.0 * pi * q / n);
  }
  return 0;
}
static void fft_plan_free(fft_plan* p) { free(p->tw); free(p->scratch); }

/* Stockham autosort, out-of-place ping-pong between x and scratch; result returned in x */
static void fft_exec(const fft_plan* p, cpx* x) {
  const int n = p->n;
  cpx* src = x;
  cpx* dst = p->scratch;
  int ns = 1;
  for (int s = 0; s < p->nfac; ++s) {
    const int R = p->fac[s], nb = n / R, tstride = n / (ns * R), rstride = n / R;
    /* butterfly j = jb * ns + k (k < ns) reads src[j + r * nb] and writes dst[jb * ns * R + k + q * ns] */
    for (int jb = 0; jb < nb / ns; ++jb) {
      if (R == 4 || R == 2) {
        const cpx* s0 = src + jb * ns;
        cpx* d0 = dst + jb * ns * R;
        for (int k = 0; k < ns; ++k) {
          cpx v[4];
          for (int r = 0; r < R; ++r) {
            cpx a = s0[k + r * nb];
            if (r > 0 && k > 0) {
              const cpx w = p->tw[k * r * tstride];
              const float re = a.re * w.re - a.im * w.im, im = a.re * w.im + a.im * w.re;
              a.re = re; a.im = im;
            }
            v[r] = a;
          }
          if (R == 2) {
            d0[k].re = v[0].re + v[1].re; d0[k].im = v[0].im + v[1].im;
            d0[k + ns].re = v[0].re - v[1].re; d0[k + ns].im = v[0].im - v[1].im;
          } else {
            const float ar = v[0].re + v[2].re, ai = v[0].im + v[2].im, br = v[0].re - v[2].re, bi = v[0].im - v[2].im;
            const float cr = v[1].re + v[3].re, ci = v[1].im + v[3].im, dr = v[1].re - v[3].re, di = v[1].im - v[3].im;
            d0[k].re = ar + cr; d0[k].im = ai + ci;
            d0[k + ns].re = br + di; d0[k + ns].im = bi - dr;
            d0[k + 2 * ns].re = ar - cr; d0[k + 2 * ns].im = ai - ci;
            d0[k + 3 * ns].re = br - di; d0[k + 3 * ns].im = bi + dr;
          }
        }
        continue;
      }
      for (int k = 0; k < ns; ++k) {
        const int j = jb * ns + k, j0 = jb * ns * R + k;
        cpx v[64];
        for (int r = 0; r < R; ++r) {
          cpx a = src[j + r * nb];
          if (r > 0 && k > 0) {
            const cpx w = p->tw[k * r * tstride];
            const float re = a.re * w.re - a.im * w.im, im = a.re * w.im + a.im * w.re;
            a.re = re; a.im = im;
          }
          v[r] = a;
        }
        for (int q = 0; q < R; ++q) {
          float re = v[0].re, im = v[0].im;
          for (int r = 1; r < R; ++r) {
            const cpx w = p->tw[((r * q) % R) * rstride];
            re += v[r].re * w.re - v[r].im * w.im;
            im += v[r].re * w.im + v[r].im * w.re;
          }
          dst[j0 + q * ns].re = re; dst[j0 + q * ns].im = im;
        }
      }
    }
    cpx* t = src; src = dst; dst = t;
    ns *= R;
  }
  if (src != x) memcpy(x, src, sizeof(cpx) * n);
}

/* real-input FFT of length n (any n): even n through an n/2 complex FFT, odd n directly */
typedef struct { int n; fft_plan half; fft_plan full; cpx* buf; cpx* twr; } rfft_plan;
static int rfft_init(rfft_plan* p, int n) {
  p->n = n;
  const double pi = 3.14159265358979323846;
  if (n % 2 == 0) {
    if (fft_plan_init(&p->half, n / 2)) return -1;
    p->buf = (cpx*)malloc(sizeof(cpx) * (n / 2 + 1));
    p->twr = (cpx*)malloc(sizeof(cpx) * (n / 2 + 1));
    for (int k = 0; k <= n / 2; ++k) { p->twr[k].re = (float)cos(-2.0 * pi * k / n); p->twr[k].im = (float)sin(-2.0 * pi * k / n); }
  } else {
    if (fft_plan_init(&p->full, n)) return -1;
    p->buf = (cpx*)malloc(sizeof(cpx) * n);
    p->twr = NULL;
  }
  return 0;
}
static void rfft_free(rfft_plan* p) {
  if (p->n % 2 == 0) fft_plan_free(&p->half); else fft_plan_free(&p->full);
  free(p->buf); free(p->twr);
}
/* out[0..n/2] = sum_j in[j] e^{-2 pi i jk/n} */
static void rfft_exec(rfft_plan* p, const float* in, cpx* out) {
  const int n = p->n;
  if (n % 2) {
    for (int j = 0; j < n; ++j) { p->buf[j].re = in[j]; p->buf[j].im = 0.f; }
    fft_exec(&p->full, p->buf);
    for (int k = 0; k <= n / 2; ++k) out[k] = p->buf[k];
    return;
  }
  const int h = n / 2;
  cpx* z = p->buf;
  for (int j = 0; j < h; ++j) { z[j].re = in[2 * j]; z[j].im = in[2 * j + 1]; }
  fft_exec(&p->half, z);
  z[h] = z[0];
  for (int k = 0; k <= h; ++k) {
    const cpx a = z[k], b = z[h - k];
    const float er = 0.5f * (a.re + b.re), ei = 0.5f * (a.im - b.im);   /* even part */
    const float or_ = 0.5f * (a.im + b.im), oi = -0.5f * (a.re - b.re); /* odd part */
    const cpx w = p->twr[k];
    out[k].re = er + or_ * w.re - oi * w.im;
    out[k].im = ei + or_ * w.im + oi * w.re;
  }
}

/* ------------------------------------------------------------------------------ config */
typedef struct {
  int64_t sample_rate, fft_points;
  float frame_length, frame_stride;
  int64_t num_cepstral, num_filters;
  float low_frequency, high_frequency;
  int32_t dc_elimination;
  float dct2_scale;
} orc_config;

static int64_t rust_round(float x) { float a = floorf(fabsf(x) + 0.5f); return (int64_t)(x < 0 ? -a : a); }

int64_t orc_num_frames(int64_t n, int64_t L, int64_t S, int zero_padding) {
  if (L <= 0 || S <= 0 || n < L) return -1;
  const float q = (float)(n - L) / (float)S;
  const int64_t t = zero_padding ? (int64_t)ceilf(q) : (int64_t)floorf(q);
  return (!zero_padding && t < 1) ? -1 : t;
}

/* feature.rs:36-90 -> dense [nf][F] */
static void filterbanks(const orc_config* c, int F, float* fb) {
  const int nf = (int)c->num_filters;
  const float sr = (float)c->sample_rate;
  const float m0 = 1127.0f * logf(1.0f + c->low_frequency / 700.0f), m1 = 1127.0f * logf(1.0f + c->high_frequency / 700.0f);
  const float step = (m1 - m0) / (float)(nf + 1);
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (nf + 2));
  for (int i = 0; i < nf + 2; ++i) {
    const float hz = 700.0f * (expf((m0 + step * (float)i) / 1127.0f) - 1.0f);
    idx[i] = (int64_t)((float)(F + 1) * hz / sr);
  }
  memset(fb, 0, sizeof(float) * nf * F);
  for (int i = 0; i < nf; ++i) {
    const float l = (float)idx[i], m = (float)idx[i + 1], r = (float)idx[i + 2];
    for (int64_t b = idx[i]; b <= idx[i + 2] && b < F; ++b) {
      const float x = (float)b;
      float v = 0.f;
      if (l <= x && x < r) {
        if (x <= m) v = (x - l) / (m - l);
        if (m <= x) v = (r - x) / (r - m);
      }
      fb[i * F + b] = v;
    }
  }
  free(idx);
}

/* per-thread plan: everything SpeechConfig + the per-call handlers hold */
typedef struct {
  orc_config c;
  int L, S, W, F, H, nf, nc;
  rfft_plan rf;
  float* fb;    /* [nf][F] */
  float* fbT;   /* [F][nf]: the same matrix transposed, so the dense product vectorises over filters while
                   every filter still sums its bins in index order (the sequential order of a plain dot) */
  float* dct;   /* [nf][nf] full DCT-II (nddct2 transforms the whole axis, then the slice) */
  float* dctT;  /* [nf][nf] transposed, same reason */
  float* window;
  float* frame; /* [W] */
  cpx* spec;    /* [F] */
  float* P;     /* [F] */
  float* feat;  /* [nf] */
  float* full;  /* [nf] */
} orc_plan;

static int plan_init(orc_plan* p, const orc_config* c) {
  p->c = *c;
  p->W = (int)c->fft_points; p->F = p->W / 2 + 1;
  p->L = (int)rust_round((float)c->sample_rate * c->frame_length);
  p->S = (int)rust_round((float)c->sample_rate * c->frame_stride);
  p->H = (int)(c->frame_length * (float)c->sample_rate);
  p->nf = (int)c->num_filters; p->nc = (int)c->num_cepstral;
  if (rfft_init(&p->rf, p->W)) return -1;
  p->fb = (float*)malloc(sizeof(float) * p->nf * p->F);
  filterbanks(c, p->F, p->fb);
  p->fbT = (float*)malloc(sizeof(float) * p->nf * p->F);
  for (int i = 0; i < p->nf; ++i)
    for (int k = 0; k < p->F; ++k) p->fbT[(size_t)k * p->nf + i] = p->fb[(size_t)i * p->F + k];
  p->dct = (float*)malloc(sizeof(float) * p->nf * p->nf);
  p->dctT = (float*)malloc(sizeof(float) * p->nf * p->nf);
  const double pi = 3.14159265358979323846;
  const double sc = c->dct2_scale != 0.f ? c->dct2_scale : 2.0;
  for (int k = 0; k < p->nf; ++k)
    for (int n = 0; n < p->nf; ++n)
      p->dctT[n * p->nf + k] = p->dct[k * p->nf + n] = (float)(sc * cos(pi * k * (2.0 * n + 1.0) / (2.0 * p->nf)));
  p->window = (float*)malloc(sizeof(float) * p->W);
  for (int i = 0; i < p->W; ++i) {
    const double s = sin(0.5 * pi * (i + 0.5) / (double)(p->W / 2));
    p->window[i] = (float)sin(0.5 * pi * s * s);
  }
  p->frame = (float*)calloc(p->W, sizeof(float));
  p->spec = (cpx*)malloc(sizeof(cpx) * p->F);
  p->P = (float*)malloc(sizeof(float) * p->F);
  p->feat = (float*)malloc(sizeof(float) * p->nf);
  p->full = (float*)malloc(sizeof(float) * p->nf);
  return 0;
}
static void plan_free(orc_plan* p) {
  rfft_free(&p->rf);
  free(p->fb); free(p->fbT); free(p->dct); free(p->dctT); free(p->window); free(p->frame); free(p->spec); free(p->P); free(p->feat); free(p->full);
}

/* one clip: mfe (+ energy), optional lmfe, mfcc, cmvn.  Outputs may be NULL. */
static void clip_features(orc_plan* p, const float* x, int64_t T, float* mfe, float* energy, float* lmfe, float* mfcc,
                          float* ssc, int cmvn_mode) {
  const int L = p->L, S = p->S, W = p->W, F = p->F, nf = p->nf, nc = p->nc;
  const float inv = 1.0f / (float)W;
  const float n = (float)(T * nf);
  const float s00 = 1.0f / sqrtf(4.0f * n), srest = 1.0f / sqrtf(2.0f * n);
  for (int64_t t = 0; t < T; ++t) {
    memcpy(p->frame, x + t * S, sizeof(float) * L);          /* stack_frames gather */
    memset(p->frame + L, 0, sizeof(float) * (W - L));        /* util::pad constant */
    rfft_exec(&p->rf, p->frame, p->spec);
    float e = 0.f;
    for (int k = 0; k < F; ++k) {
      p->P[k] = inv * sqrtf(p->spec[k].re * p->spec[k].re + p->spec[k].im * p->spec[k].im);
      e += p->P[k];
    }
    if (e == 0.f) e = ORC_EPS;
    for (int i = 0; i < nf; ++i) p->feat[i] = 0.f;          /* dense GEMM row, feature.rs:229 */
    for (int k = 0; k < F; ++k) {
      const float pk = p->P[k];
      const float* w = p->fbT + (size_t)k * nf;
      for (int i = 0; i < nf; ++i) p->feat[i] += pk * w[i];
    }
    for (int i = 0; i < nf; ++i) p->feat[i] = p->feat[i] == 0.f ? ORC_EPS : p->feat[i];
    if (ssc) {  /* (P o R) fb^T / mfe with R = linspace(1, sr/2, F) */
      const float step = F > 1 ? ((float)p->c.sample_rate / 2.0f - 1.0f) / (float)(F - 1) : 0.f;
      for (int i = 0; i < nf; ++i) p->full[i] = 0.f;
      for (int k = 0; k < F; ++k) {
        const float pr = p->P[k] * (1.0f + (float)k * step);
        const float* w = p->fbT + (size_t)k * nf;
        for (int i = 0; i < nf; ++i) p->full[i] += pr * w[i];
      }
      for (int i = 0; i < nf; ++i) ssc[t * nf + i] = p->full[i] / p->feat[i];
    }
    if (mfe) memcpy(mfe + t * nf, p->feat, sizeof(float) * nf);
    if (energy) energy[t] = e;
    if (lmfe || mfcc)
      for (int i = 0; i < nf; ++i) p->feat[i] = logf(p->feat[i]);
    if (lmfe) memcpy(lmfe + t * nf, p->feat, sizeof(float) * nf);
    if (mfcc) {
      for (int k = 0; k < nf; ++k) p->full[k] = 0.f;          /* nddct2 over the whole axis */
      for (int i = 0; i < nf; ++i) {
        const float fi = p->feat[i];
        const float* d = p->dctT + (size_t)i * nf;
        for (int k = 0; k < nf; ++k) p->full[k] += fi * d[k];
      }
      float* o = mfcc + t * nc;
      for (int k = 0; k < nc; ++k) o[k] = k == 0 ? (t == 0 ? p->full[0] * s00 : p->full[0]) : p->full[k] * srest;
      if (p->c.dc_elimination && nc > 0) o[0] = logf(e);
    }
  }
  if (mfcc && cmvn_mode) {                                    /* processing.rs:265-300 */
    for (int c = 0; c < nc; ++c) {
      float s = 0.f;
      for (int64_t t = 0; t < T; ++t) s += mfcc[t * nc + c];
      const float mean = s / (float)T;
      float m = 0.f, ssq = 0.f;
      for (int64_t t = 0; t < T; ++t) {
        const float y = mfcc[t * nc + c] - mean;
        mfcc[t * nc + c] = y;
        const float d = y - m;
        m += d / (float)(t + 1);
        ssq += d * (y - m);
      }
      if (cmvn_mode == 2) {
        const float sd = sqrtf(ssq / (float)T) + ORC_CMVN_EPS;
        for (int64_t t = 0; t < T; ++t) mfcc[t * nc + c] /= sd;
      }
    }
  }
}

/* pcm [n_clips][clip_len] -> any of mfe [n][T][nf], energy [n][T], lmfe [n][T][nf], mfcc [n][T][nc] */
int orc_features(const orc_config* c, const float* pcm, int64_t n_clips, int64_t clip_len, float* mfe, float* energy,
                 float* lmfe, float* mfcc, float* ssc, int cmvn_mode, int n_threads) {
  orc_plan probe;
  if (plan_init(&probe, c)) return -1;
  const int64_t T = orc_num_frames(clip_len, probe.L, probe.S, 0);
  const int nf = probe.nf, nc = probe.nc;
  const int ok = probe.L <= probe.W && T >= 1;
  plan_free(&probe);
  if (!ok) return -1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    orc_plan p;  /* one plan per worker, like one SpeechConfig clone per rayon worker */
    plan_init(&p, c);
#pragma omp for schedule(dynamic, 8)
    for (int64_t i = 0; i < n_clips; ++i)
      clip_features(&p, pcm + i * clip_len, T, mfe ? mfe + i * T * nf : NULL, energy ? energy + i * T : NULL,
                    lmfe ? lmfe + i * T * nf : NULL, mfcc ? mfcc + i * T * nc : NULL, ssc ? ssc + i * T * nf : NULL,
                    cmvn_mode);
    plan_free(&p);
  }
  return 0;
}

/* mel_spectrogram2 (feature.rs:163-174) over stft2 (functions.rs:86-170).  The reference threads
 * ONE overlap memory through all rows sequentially; for the parallel baseline each row's initial
 * memory is restated from the previous row's tail (identical values, no serial dependency).
 * out [ch][nf][Tp]; mem_inout [W-H] (zeros when NULL, not written back). */
int orc_mel_spectrogram2(const orc_config* c, const float* pcm, int64_t ch, int64_t ttd, float* out, const float* mem0,
                         int log_mode, int n_threads) {
  orc_plan probe;
  if (plan_init(&probe, c)) return -1;
  const int W = probe.W, H = probe.H, F = probe.F, nf = probe.nf;
  plan_free(&probe);
  if (H <= 0 || 2 * H > W) return -1;
  const int n_pad = W / H - 1;
  const int64_t Tp = (int64_t)ceilf((float)ttd / (float)H);
  const int memlen = W - H;
  const float wnorm = 1.0f / ((float)((int64_t)W * W) / (float)(2 * H));
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    orc_plan p;
    plan_init(&p, c);
    float* stream = (float*)malloc(sizeof(float) * (memlen + Tp * H));
    float* pw = (float*)malloc(sizeof(float) * Tp * F);
#pragma omp for schedule(dynamic, 1)
    for (int64_t r = 0; r < ch; ++r) {
      /* memory in front of row r: last W-H samples of mem0 ++ pad(row 0) ++ ... ++ pad(row r-1) */
      for (int i = 0; i < memlen; ++i) {
        int64_t q = (int64_t)memlen + r * Tp * H - memlen + i;
        float v;
        if (q < memlen) v = mem0 ? mem0[q] : 0.f;
        else { q -= memlen; const int64_t rr = q / (Tp * H), o = q % (Tp * H); v = o < ttd ? pcm[rr * ttd + o] : 0.f; }
        stream[i] = v;
      }
      memcpy(stream + memlen, pcm + r * ttd, sizeof(float) * ttd);
      memset(stream + memlen + ttd, 0, sizeof(float) * (Tp * H - ttd));
      memset(pw, 0, sizeof(float) * Tp * F);
      for (int64_t j = n_pad; j < Tp; ++j) {                 /* frame j -> output row j - n_pad */
        for (int i = 0; i < W; ++i) p.frame[i] = stream[j * H + i] * p.window[i];
        rfft_exec(&p.rf, p.frame, p.spec);
        float* dst = pw + (j - n_pad) * F;
        for (int k = 0; k < F; ++k) {
          const float re = p.spec[k].re * wnorm, im = p.spec[k].im * wnorm;
          const float a = hypotf(re, im);
          dst[k] = a * a;
        }
      }
      for (int64_t t = 0; t < Tp; ++t) {                      /* dense einsum ntf,mf->nmt: every (m, t) sums */
        const float* pr = pw + t * F;                         /* its bins in index order; vectorised over m   */
        for (int m = 0; m < nf; ++m) p.feat[m] = 0.f;
        for (int k = 0; k < F; ++k) {
          const float pk = pr[k];
          const float* w = p.fbT + (size_t)k * nf;
          for (int m = 0; m < nf; ++m) p.feat[m] += pk * w[m];
        }
        for (int m = 0; m < nf; ++m) {
          const float acc = p.feat[m];
          out[(r * nf + m) * Tp + t] = log_mode ? logf(acc == 0.f ? ORC_EPS : acc) : acc;
        }
      }
    }
    free(stream); free(pw);
    plan_free(&p);
  }
  return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
