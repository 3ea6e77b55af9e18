// Generic frame pipeline for any FFT length that factors into primes <= 61.
//
// One CTA processes a tile of 8 (or fewer, for long FFTs) consecutive frames of one row
// (clip / channel), one thread group per frame.  Each real frame goes through a half-length
// complex FFT (z[m] = x[2m] + i x[2m+1]) plus the standard real-FFT post-processing step, so a
// frame's rounding error is relative to that frame alone (two frames are deliberately NOT packed
// into one complex FFT: the ~1e-7 crosstalk would turn an all-zero frame next to a loud one into
// tiny non-zeros, and the reference's zero_handling -> EPS would no longer trigger).  The
// transform is a Stockham autosort FFT that ping-pongs between two shared-memory buffers, radix
// 4/2 and odd radices through a small O(R^2) DFT with table twiddles.  Everything after the FFT (separation of
// the two spectra, magnitude, frame energy, banded mel reduction, ln, DCT-II, SSC) happens
// on-chip; only final outputs are written.
//
// Serves: feature::{mfe,lmfe,mfcc,ssc} (feature.rs:99-245), processing::power_spectrum
// (processing.rs:143-181), functions::{stft1,stft2} (functions.rs:86-233) and
// feature::mel_spectrogram{1,2} (feature.rs:151-174) for sizes that have no specialised
// kernel, and is the in-library cross-check of the specialised kernels.
#include <cuda_runtime.h>

#include "ss_internal.h"

namespace ss {

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }  // functions.rs:66-71

// intended gather frames[f][n] = x[f*S + n] (processing.rs:65-129, policy D1), or the literal
// behaviour of processing.rs:110-120 when SS_FLAG_FRAMING_LITERAL is set.
__device__ __forceinline__ float fetch_speechpy(const FrameParams& p, const float* row, int64_t f, int n) {
  if (f >= p.T || n >= p.L) return 0.0f;
  if (p.flags & SS_FLAG_FRAMING_LITERAL) {
    if (p.T > 2 || n >= 2 * (p.L / 2)) return 0.0f;
    return row[n];
  }
  return row[f * (int64_t)p.S + n];
}

// STFT stream model (functions.rs:86-170): frame j of row r starts at index r*Tp*H + j*H of
// the stream  [mem_in (W-H)] ++ [row 0 zero-padded to Tp*H] ++ [row 1 ...] ...
__device__ __forceinline__ float fetch_stream(const FrameParams& p, int64_t q) {
  const int64_t memlen = p.N - p.S;
  if (q < memlen) return p.mem_in ? p.mem_in[q] : 0.0f;
  q -= memlen;
  const int64_t span = p.Tp * (int64_t)p.S;
  const int64_t r = q / span;
  const int64_t o = q - r * span;
  if (r >= p.n_rows || o >= p.row_len) return 0.0f;
  return p.pcm[r * p.row_stride + o];
}

template <int R>
__device__ __forceinline__ void stockham_butterfly(const float2* __restrict__ src, float2* __restrict__ dst,
                                                   const float2* __restrict__ tw, int N, int Ns, int j,
                                                   int twmul) {
  const int nb = N / R;
  const int k = j % Ns;
  const int tstride = twmul * (N / (Ns * R));
  float2 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    v[r] = src[j + r * nb];
    if (r > 0 && k > 0) v[r] = cmul(v[r], __ldg(&tw[k * r * tstride]));
  }
  float2 o[R];
  if (R == 2) {
    o[0] = cadd(v[0], v[1]);
    o[1] = csub(v[0], v[1]);
  } else if (R == 4) {
    float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    float2 dmi = make_float2(d.y, -d.x);  // -i * d
    o[0] = cadd(a, c);
    o[1] = cadd(b, dmi);
    o[2] = csub(a, c);
    o[3] = csub(b, dmi);
  } else {
    const int rstride = twmul * (N / R);
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float2 acc = v[0];
#pragma unroll
      for (int r = 1; r < R; ++r) acc = cadd(acc, cmul(v[r], __ldg(&tw[((r * q) % R) * rstride])));
      o[q] = acc;
    }
  }
  const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) dst[j0 + r * Ns] = o[r];
}

// odd radices without a template instantiation (11..61): values live in local memory
__device__ void stockham_butterfly_any(const float2* __restrict__ src, float2* __restrict__ dst,
                                       const float2* __restrict__ tw, int N, int Ns, int R, int j, int twmul) {
  const int nb = N / R;
  const int k = j % Ns;
  const int tstride = twmul * (N / (Ns * R));
  const int rstride = twmul * (N / R);
  float2 v[61];
  for (int r = 0; r < R; ++r) {
    v[r] = src[j + r * nb];
    if (r > 0 && k > 0) v[r] = cmul(v[r], __ldg(&tw[k * r * tstride]));
  }
  const int j0 = (j / Ns) * Ns * R + k;
  for (int q = 0; q < R; ++q) {
    float2 acc = v[0];
    for (int r = 1; r < R; ++r) acc = cadd(acc, cmul(v[r], __ldg(&tw[((r * q) % R) * rstride])));
    dst[j0 + q * Ns] = acc;
  }
}

__global__ void __launch_bounds__(256) k_frames_generic(const FrameParams p, const int fpc, const int tpg,
                                                        const int64_t tiles_per_row) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N, F = p.F, nf = p.nf, nc = p.nc;
  const bool even = (N & 1) == 0;
  const int M = even ? N / 2 : N;  // complex FFT length: real FFT through a half-length complex FFT
  const int twmul = even ? 2 : 1;  // the table holds W_N^q; the M-point FFT needs W_M = W_N^2
  const int TF = fpc;
  float2* fftbuf = reinterpret_cast<float2*>(smem_raw);                   // [fpc][2][M+1]
  float* mel_s = reinterpret_cast<float*>(fftbuf + (size_t)fpc * 2 * (M + 1));  // [TF][nf]
  float* ssc_s = mel_s + TF * nf;                                         // [TF][nf]
  float* log_s = ssc_s + TF * nf;                                         // [TF][nf]
  float* energy_s = log_s + TF * nf;                                      // [TF]
  float* red_s = energy_s + TF;                                           // [8 warps]

  const int tid = threadIdx.x;
  const int g = tid / tpg, lt = tid % tpg;
  const int64_t row = blockIdx.x / tiles_per_row;
  const int64_t t0 = (blockIdx.x % tiles_per_row) * TF;
  const bool librosa = p.mode == kModeLibrosa;
  const bool stft = p.mode == kModeStft || librosa;  // both: window from the table, |X|^2, time-minor mel output
  const int64_t fr = t0 + g;  // this thread group's frame
  float2* bufA = fftbuf + (size_t)g * 2 * (M + 1);
  float2* bufB = bufA + (M + 1);

  // ---- 1. framing (+ window) straight into the FFT buffer: z[m] = x[2m] + i x[2m+1] ---------
  {
    const float* rowp = p.pcm + row * p.row_stride;
    const int64_t span = p.Tp * (int64_t)p.S;
    const int64_t memlen = p.N - p.S;
    const int64_t rbase = memlen + row * span;  // stream index of this row's sample 0 (STFT)
    const int64_t q0 = row * span + (fr + p.n_pad) * (int64_t)p.S;
    auto sample = [&](int n) -> float {
      if (n >= N) return 0.f;
      if (!stft) return fetch_speechpy(p, rowp, fr, n);
      if (fr >= p.T) return 0.f;
      if (librosa) {  // librosa.stft(center=True, pad_mode="reflect"): frame fr covers y_padded[fr H .. fr H + N)
        int64_t i = fr * (int64_t)p.S + n - N / 2;
        if (i < 0) i = -i;
        if (i >= p.row_len) i = 2 * (p.row_len - 1) - i;
        return rowp[i] * __ldg(&p.tab.window[n]);
      }
      const int64_t q = q0 + n;
      float v;
      if (q >= rbase) {
        const int64_t o = q - rbase;
        v = o < p.row_len ? rowp[o] : 0.f;
      } else {
        v = fetch_stream(p, q);
      }
      return v * __ldg(&p.tab.window[n]);
    };
    if (even)
      for (int m = lt; m < M; m += tpg) bufA[m] = make_float2(sample(2 * m), sample(2 * m + 1));
    else
      for (int m = lt; m < M; m += tpg) bufA[m] = make_float2(sample(m), 0.f);
  }
  __syncthreads();

  // ---- 2. Stockham FFT of length M ------------------------------------------------------------
  float2* src = bufA;
  float2* dst = bufB;
  int Ns = 1;
  for (int s = 0; s < p.tab.n_stages; ++s) {
    const int R = p.tab.radix[s];
    const int nb = M / R;
    for (int j = lt; j < nb; j += tpg) {
      switch (R) {
        case 2: stockham_butterfly<2>(src, dst, p.tab.tw, M, Ns, j, twmul); break;
        case 4: stockham_butterfly<4>(src, dst, p.tab.tw, M, Ns, j, twmul); break;
        case 3: stockham_butterfly<3>(src, dst, p.tab.tw, M, Ns, j, twmul); break;
        case 5: stockham_butterfly<5>(src, dst, p.tab.tw, M, Ns, j, twmul); break;
        case 7: stockham_butterfly<7>(src, dst, p.tab.tw, M, Ns, j, twmul); break;
        default: stockham_butterfly_any(src, dst, p.tab.tw, M, Ns, R, j, twmul); break;
      }
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
    Ns *= R;
  }
  const float2* Z = src;                      // Z[k], k < M
  float* P = reinterpret_cast<float*>(dst);   // the other buffer is free now: F <= 2(M+1) floats

  // ---- 3. real-FFT post-processing, magnitude / power ---------------------------------------
  //   X[k] = E + W_N^k O,  X[M-k] = conj(E - W_N^k O),  E = (Z[k]+conj Z[M-k])/2, O = (Z[k]-conj Z[M-k])/(2i)
  float esum = 0.f;
  const bool fvalid = fr < p.T;
  auto emit = [&](int k, float2 x) {
    if (!stft) {
      const float a = p.mag_scale * sqrtf(x.x * x.x + x.y * x.y);  // processing.rs:168,180
      P[k] = a;
      esum += a;
      if (p.pspec_out && fvalid) p.pspec_out[(row * p.T + fr) * F + k] = a;
    } else {
      x.x *= p.wnorm;  // functions.rs:166-169
      x.y *= p.wnorm;
      if (p.stft_out && fr < p.T_out) p.stft_out[(row * p.T_out + fr) * F + k] = fvalid ? x : make_float2(0.f, 0.f);
      P[k] = x.x * x.x + x.y * x.y;  // feature.rs:164: |X|^2
    }
  };
  if (even) {
    for (int k = lt; k <= M / 2; k += tpg) {
      const float2 a = Z[k], b = Z[(M - k) % M];
      const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      const float2 O = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
      const float2 Tw = cmul(O, __ldg(&p.tab.tw[k]));
      emit(k, make_float2(E.x + Tw.x, E.y + Tw.y));
      if (M - k != k) emit(M - k, make_float2(E.x - Tw.x, -(E.y - Tw.y)));
    }
  } else {
    for (int k = lt; k < F; k += tpg) emit(k, Z[k]);
  }
  // ---- 4. frame energy (feature.rs:216-219) ----------------------------------------------
  if (!stft) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    if ((tid & 31) == 0) red_s[tid >> 5] = esum;
  }
  __syncthreads();
  if (!stft && lt == 0) {
    float a = 0.f;
    for (int w = 0; w < tpg / 32; ++w) a += red_s[(tid >> 5) + w];
    energy_s[g] = zero_handling(a);
  }

  // ---- 5. banded mel reduction (feature.rs:229 / :173), SSC numerator --------------------
  const bool want_mel = stft ? (p.mel_out != nullptr) : (p.mfe || p.lmfe || p.ssc || p.mfcc);
  if (want_mel) {
    for (int i = lt; i < nf; i += tpg) {
      const int b0 = __ldg(&p.tab.fstart[i]), len = __ldg(&p.tab.flen[i]);
      const float* w = p.tab.fw + __ldg(&p.tab.fwofs[i]);
      float acc = 0.f, acc2 = 0.f;
      for (int q = 0; q < len; ++q) {
        const float pw = P[b0 + q] * __ldg(&w[q]);
        acc += pw;
        acc2 += pw * (p.ssc_r0 + p.ssc_step * (float)(b0 + q));
      }
      const int o = g * nf + i;
      if (!stft) {
        acc = zero_handling(acc);  // feature.rs:230
        mel_s[o] = acc;
        ssc_s[o] = acc2 / acc;
        log_s[o] = logf(acc);
      } else {
        if (p.log_mode == 2) mel_s[o] = 10.0f * log10f(fmaxf(acc, 1e-10f));  // librosa.power_to_db(ref=1, amin=1e-10, top_db=None)
        else mel_s[o] = p.log_mode ? logf(zero_handling(acc)) : acc;
      }
    }
  }
  __syncthreads();

  // ---- 6. outputs ------------------------------------------------------------------------
  const int nthreads = blockDim.x;
  if (!stft) {
    const int64_t obase = (row * p.T + t0);
    if (p.mfe || p.lmfe || p.ssc) {
      for (int idx = tid; idx < TF * nf; idx += nthreads) {
        const int f = idx / nf;
        if (t0 + f >= p.T) break;
        if (p.mfe) p.mfe[obase * nf + idx] = mel_s[idx];
        if (p.lmfe) p.lmfe[obase * nf + idx] = log_s[idx];
        if (p.ssc) p.ssc[obase * nf + idx] = ssc_s[idx];
      }
    }
    if (p.energy && tid < TF && t0 + tid < p.T) p.energy[obase + tid] = energy_s[tid];
    if (p.mfcc) {
      for (int idx = tid; idx < TF * nc; idx += nthreads) {
        const int f = idx / nc, c = idx - f * nc;
        const int64_t t = t0 + f;
        if (t >= p.T) break;
        const float* lg = log_s + f * nf;
        const float* dm = p.tab.dct + c * nf;
        float acc = 0.f;
        for (int n = 0; n < nf; ++n) acc += lg[n] * __ldg(&dm[n]);
        // feature.rs:126-131: [0,0] *= 1/sqrt(4n); columns 1.. *= 1/sqrt(2n); rest of column 0 untouched
        if (c == 0) acc = (t == 0) ? acc * p.scale_00 : acc;
        else acc *= p.scale_rest;
        if (c == 0 && p.dc_elim) acc = logf(energy_s[f]);  // feature.rs:137-146
        p.mfcc[obase * nc + idx] = acc;
      }
    }
  } else if (p.mel_out) {
    // [row][m][t], time-minor (feature.rs:173); rows t >= T are the reference's zero tail
    for (int idx = tid; idx < nf * TF; idx += nthreads) {
      const int m = idx / TF, f = idx - m * TF;
      const int64_t t = t0 + f;
      if (t >= p.T_out) continue;
      float v = 0.f;
      if (t < p.T) v = mel_s[f * nf + m];
      else if (p.log_mode) v = logf(kEps);
      p.mel_out[(row * nf + m) * p.T_out + t] = v;
    }
  }
}

__global__ void k_stft_mem_update(const FrameParams p) {
  // overlap memory after the last row = last W-H samples of the whole stream (functions.rs:153-160)
  const int64_t memlen = p.N - p.S;
  const int64_t total = memlen + p.n_rows * p.Tp * (int64_t)p.S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < memlen; i += (int64_t)gridDim.x * blockDim.x)
    p.mem_out[i] = fetch_stream(p, total - memlen + i);
}

}  // namespace

int launch_frames_generic(const FrameParams& p, void* stream) {
  const int N = p.N;
  if (N > 8192) return (int)cudaErrorInvalidValue;
  const int M = (N & 1) ? N : N / 2;
  int fpc = 8;  // frames per CTA
  while (fpc > 1 && (size_t)fpc * 2 * (M + 1) * sizeof(float2) > 128 * 1024) fpc >>= 1;
  const int tpg = 256 / fpc;
  const int TF = fpc;
  const size_t smem = (size_t)fpc * 2 * (M + 1) * sizeof(float2) + (size_t)(3 * TF * p.nf + TF + 16) * sizeof(float);
  if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
  const int64_t frames = (p.mode == kModeStft || p.mode == kModeLibrosa) ? p.T_out : p.T;
  const int64_t tiles_per_row = (frames + TF - 1) / TF;
  const int64_t nblocks = tiles_per_row * p.n_rows;
  if (nblocks <= 0) return 0;
  if (nblocks > 2147483647LL) return (int)cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_frames_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_frames_generic<<<(unsigned)nblocks, 256, smem, (cudaStream_t)stream>>>(p, fpc, tpg, tiles_per_row);
  count_launch();
  note_kernel("k_frames_generic");
  return (int)cudaGetLastError();
}

int launch_stft_mem_update(const FrameParams& p, void* stream) {
  if (!p.mem_out || p.N - p.S <= 0) return 0;
  const int64_t memlen = p.N - p.S;
  k_stft_mem_update<<<(unsigned)((memlen + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace ss
