// fft_points = 512 and 1024 on the STFT / librosa paths: functions::stft1/2 (functions.rs:86-233),
// feature::mel_spectrogram1/2 (feature.rs:151-174) and the librosa-compatible mel spectrogram for those n_fft.
// (Kernel template k_stft512<R, ..>: R = 16 points per lane -> fft_points 512, R = 32 -> 1024, R = 8 -> 256.)
//
// Two kernels instead of one fused one (the fused MMA epilogue of k_fast512c is tied to speechpy framing):
//   k_stft512        one half-warp per frame: window, the 16 x 16 four-step register FFT of k_fast512c / k_pspec512,
//                    real-FFT split; writes X wnorm (complex STFT output) and / or |X wnorm|^2 rows into a scratch.
//                    Frames inside their row are loaded as float2; frames that touch the overlap memory, the previous
//                    row, the zero padding (stream model, functions.rs:99-121) or the reflected edges (librosa mode)
//                    fetch per sample through the same index formulas as k_frames_generic.
//   k_mel_from_power banded mel reduction of 16 (F = 257) or 8 (wider spectra) frames x all filters per block from those rows, time-minor store
//                    [row][m][t] (feature.rs:173), ln(zero_handling) / power_to_db on request.
// The |X|^2 scratch costs one extra write and read of 1 KB per frame; measured 4 x the rate of the any-length kernel
// (8,192 x 10 s rows, hop 256: 5.9 ms against 23.0 ms; librosa mode, hop 160: 9.5 ms against 39.7 ms).  Each filter sums its bins in index order, as the reference's einsum does.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

constexpr int kWarps = 8;
constexpr int kXs = 18;                       // transpose row stride in float2 (144 B: conflict-free LDS.128)
// R complex points per lane, 16 lanes per frame: a 16 R-point complex FFT = a real FFT of W = 32 R samples
template <int R>
struct Geo {
  static constexpr int kW = 32 * R, kNc = 16 * R, kF = kNc + 1;
  static constexpr int kNS = R >= 16 ? R / 16 : 1;     // 16-point sub-FFTs per lane in the second stage
  static constexpr int kAct = R >= 16 ? 16 : R;        // lanes of a frame that own rows in the second stage (R = 8: half of them)
  static constexpr int kXchFrame = R * kXs;            // float2 per frame
  static constexpr int kWarpFloats = 2 * kXchFrame * 2;
  static constexpr int kTwRow = 2 * R + 4;             // floats per lane row of the W_Nc table
  static constexpr int kTabTw = 16 * kTwRow;
  static constexpr size_t kSmem = (size_t)(kTabTw + kW + kWarps * kWarpFloats) * sizeof(float);
};

__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }

// STFT stream model (functions.rs:86-170), as in k_frames_generic
__device__ __forceinline__ float stream_at(const FrameParams& p, int64_t q) {
  const int64_t memlen = p.N - p.S;
  if (q < 0) return 0.f;
  if (q < memlen) return p.mem_in ? p.mem_in[q] : 0.0f;
  q -= memlen;
  const int64_t span = p.Tp * (int64_t)p.S;
  const int64_t r = q / span;
  const int64_t o = q - r * span;
  if (r >= p.n_rows || o >= p.row_len) return 0.0f;
  return p.pcm[r * p.row_stride + o];
}

template <int R, bool LIBROSA>
__global__ void __launch_bounds__(kWarps * 32) k_stft512(const FrameParams p, const float* __restrict__ tw_tab, float* __restrict__ power,
                                                          int64_t row0, int64_t n_rows) {
  using G = Geo<R>;
  constexpr int kW = G::kW, kNc = G::kNc, kF = G::kF, kNS = G::kNS, kAct = G::kAct;
  extern __shared__ __align__(16) float smem[];
  float* tw_s = smem;
  float* win_s = smem + G::kTabTw;
  float* xch_s = win_s + kW;
  for (int i = threadIdx.x; i < G::kTabTw; i += kWarps * 32) tw_s[i] = __ldg(tw_tab + i);
  for (int i = threadIdx.x; i < kW; i += kWarps * 32) win_s[i] = __ldg(p.tab.window + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = lane >> 4, u = lane & 15;
  // second stage: lane u owns the rows k1 = ua + 16 s.  With R = 8 there are only 8 rows: lanes 8..15 shadow lanes 0..7
  // (same loads, same shuffles, no stores).
  const int ua = u & (kAct - 1);
  const bool owner = u < kAct;
  const bool isz = (ua == 0);
  const int partner = (lane & 16) | ((kAct - ua) & (kAct - 1));
  const float4* twrow = reinterpret_cast<const float4*>(tw_s + u * G::kTwRow);
  float2 wu[kNS];  // W_{2 Nc}^{ua + 16 s}
#pragma unroll
  for (int s = 0; s < kNS; ++s) sincospif(-(float)(ua + 16 * s) / (float)kNc, &wu[s].y, &wu[s].x);
  float2* xch = reinterpret_cast<float2*>(xch_s + warp * G::kWarpFloats) + fi * G::kXchFrame;
  const float2* win2 = reinterpret_cast<const float2*>(win_s) + u;  // window pairs (2m, 2m+1), m = u + 16 j

  const int64_t T = p.T, T_out = p.T_out;
  const int H = p.S;
  const int64_t span = p.Tp * (int64_t)H;
  const int64_t memlen = kW - H;
  const int c0 = LIBROSA ? -kW / 2 : (p.n_pad + 1) * H - kW;  // first sample of frame 0 relative to its row
  const float scale = 0.5f * p.wnorm;                           // the 1/2 of E', O' and functions.rs:166-169
  const bool vec_ok = (H % 2 == 0) && (c0 % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  // frames of rows [row0, row0 + n_rows); the frames [T, T_out) of a row are the reference's zero tail
  const int64_t pairs_per_row = (T_out + 1) >> 1;
  const int64_t n_pairs = n_rows * pairs_per_row;
  for (int64_t pr = (int64_t)blockIdx.x * kWarps + warp; pr < n_pairs; pr += (int64_t)gridDim.x * kWarps) {
    const int64_t rl = pr / pairs_per_row;               // row within this launch
    const int64_t row = row0 + rl;
    const int64_t t = 2 * (pr - rl * pairs_per_row) + fi;
    const bool stored = t < T_out;                        // an odd T_out leaves the last half-warp idle
    const bool computed = t < T;
    const float* rowp = p.pcm + row * p.row_stride;
    const int64_t o0 = t * H + c0;
    // ---- framing + window: v[j] = (x w)[2m] + i (x w)[2m+1], m = u + 16 j
    float2 v[R];
    if (!computed) {
      sfor<0, R>([&](auto jj) { v[decltype(jj)::value] = make_float2(0.f, 0.f); });
    } else if (o0 >= 0 && o0 + kW <= p.row_len && vec_ok) {
      const float* px = rowp + o0 + 2 * u;
      sfor<0, R>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        v[j] = __fmul2_rn(__ldg(reinterpret_cast<const float2*>(px + 32 * j)), win2[16 * j]);
      });
    } else {
      sfor<0, R>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float2 w = win2[16 * j];
        float xs[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          int64_t o = o0 + 2 * u + 32 * j + e;
          if (LIBROSA) {  // librosa.stft(center=True, pad_mode="reflect")
            if (o < 0) o = -o;
            if (o >= p.row_len) o = 2 * (p.row_len - 1) - o;
            xs[e] = rowp[o];
          } else {
            xs[e] = o >= 0 ? (o < p.row_len ? rowp[o] : 0.f) : stream_at(p, memlen + row * span + o);
          }
        }
        v[j] = make_float2(xs[0] * w.x, xs[1] * w.y);
      });
    }
    // ---- Nc-point complex FFT: R-point FFT over j, twiddle W_Nc^{u k1}, transpose, 16-point FFTs over the lanes
    ssfft::fft_pow2<R>(v);
    sfor<0, R / 2>([&](auto kk) {
      constexpr int k1 = 2 * decltype(kk)::value;
      const float4 w = twrow[k1 / 2];
      float2 y0 = v[k1], y1 = v[k1 + 1];
      if constexpr (k1 > 0) y0 = ssfft::cmulf(y0, make_float2(w.x, w.y));
      y1 = ssfft::cmulf(y1, make_float2(w.z, w.w));
      xch[k1 * kXs + u] = y0;
      xch[(k1 + 1) * kXs + u] = y1;
    });
    __syncwarp();
    // lane u takes the rows k1 = ua + 16 s: gg[s][k2] = Z[ua + 16 s + R k2]
    float2 gg[kNS][16];
#pragma unroll
    for (int s = 0; s < kNS; ++s) {
      const float4* r4 = reinterpret_cast<const float4*>(xch + (ua + 16 * s) * kXs);
      sfor<0, 8>([&](auto tt) {
        constexpr int i = decltype(tt)::value;
        const float4 x = r4[i];
        gg[s][2 * i] = make_float2(x.x, x.y);
        gg[s][2 * i + 1] = make_float2(x.z, x.w);
      });
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kNS; ++s) ssfft::fft_pow2<16>(gg[s]);
    // ---- real-FFT split: X[k] = xa / 2, X[Nc - k] = conj(xb) / 2 for k = ua + 16 s + R k2 < Nc / 2 (k2 < 8).
    // Nc - k = (R - ua - 16 s) + R (15 - k2): row R - ua - 16 s belongs to lane kAct - ua, sub-FFT kNS - 1 - s (ua > 0);
    // lane 0 mirrors onto itself: sub-FFT 0 onto its own entry 16 - k2, sub-FFT s >= 1 onto sub-FFT kNS - s.
    float2* srow = (p.stft_out && stored && owner) ? p.stft_out + (row * T_out + t) * kF : nullptr;
    float* prow = (power && computed && owner) ? power + (rl * T + t) * kF : nullptr;
#pragma unroll
    for (int s = 0; s < kNS; ++s) {
      sfor<0, 8>([&](auto kk) {
        constexpr int k2 = decltype(kk)::value;
        float2 zb;
        zb.x = __shfl_sync(0xffffffffu, gg[kNS - 1 - s][15 - k2].x, partner);
        zb.y = __shfl_sync(0xffffffffu, gg[kNS - 1 - s][15 - k2].y, partner);
        if (isz) zb = s == 0 ? gg[0][(16 - k2) & 15] : gg[kNS - s][15 - k2];
        const float2 za = gg[s][k2];
        const float2 e = __fadd2_rn(za, make_float2(zb.x, -zb.y));
        const float2 od = __fadd2_rn(make_float2(za.y, -za.x), make_float2(zb.y, zb.x));
        const float2 tt = ssfft::cmulf(ssfft::twmul<32, k2>(od), wu[s]);  // W_{2 Nc}^{R k2} = W32^{k2}
        const float2 xa = ssfft::cadd(e, tt), xb = ssfft::csub(e, tt);
        const float2 Xa = make_float2(scale * xa.x, scale * xa.y), Xb = make_float2(scale * xb.x, -scale * xb.y);
        const int ka = ua + 16 * s + R * k2, kb = kNc - ka;
        if (srow) {
          srow[ka] = Xa;
          srow[kb] = Xb;
        }
        if (prow) {
          prow[ka] = fmaf(Xa.x, Xa.x, Xa.y * Xa.y);  // feature.rs:164: |X|^2
          prow[kb] = fmaf(Xb.x, Xb.x, Xb.y * Xb.y);
        }
      });
    }
    if (isz) {  // X[Nc / 2] = conj Z[Nc / 2], entry 8 of sub-FFT 0
      const float2 X = make_float2(2.0f * scale * gg[0][8].x, -2.0f * scale * gg[0][8].y);
      if (srow) srow[kNc / 2] = X;
      if (prow) prow[kNc / 2] = fmaf(X.x, X.x, X.y * X.y);
    }
  }
}

// mel[row][m][t] = sum_k P[row][t][k] fb[m][k] for TF frames x all filters per block.  TF = 16 for F = 257 (16 KB tile),
// TF = 8 for wider spectra (F = 1025: 33 KB, six blocks per SM); the 32 / TF lane groups of a warp take alternate filters
constexpr int kMelFrames = 32;
template <int TF>
__global__ void __launch_bounds__(256) k_mel_from_power(const float* __restrict__ power, const FrameParams p, int64_t row0, int F) {
  extern __shared__ float tile[];  // [TF][F | 1]: an odd row stride keeps the column walks conflict-free
  constexpr int kSub = 32 / TF;    // filters in flight per warp
  const int Fs = F | 1;
  const int64_t rl = blockIdx.y;                           // row within this launch
  const int64_t row = row0 + rl;
  const int64_t t0 = (int64_t)blockIdx.x * TF;
  const int nfr = (int)min((int64_t)TF, p.T - t0);  // computed frames in this tile (may be <= 0)
  const float* src = power + (rl * p.T + t0) * F;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int f = warp; f < TF; f += 8) {  // a warp per frame row: coalesced, no index division
    const float* s = src + (int64_t)f * F;
    float* d = tile + f * Fs;
    if (f < nfr)
      for (int k = lane; k < F; k += 32) d[k] = __ldg(s + k);
    else
      for (int k = lane; k < F; k += 32) d[k] = 0.f;
  }
  __syncthreads();
  const int fl = lane % TF;
  const int64_t t = t0 + fl;
  const float* trow = tile + fl * Fs;
  for (int m = warp * kSub + lane / TF; m < p.nf; m += 8 * kSub) {
    const int b0 = __ldg(&p.tab.fstart[m]), len = __ldg(&p.tab.flen[m]);
    const float* w = p.tab.fw + __ldg(&p.tab.fwofs[m]);
    float acc = 0.f;
    for (int q = 0; q < len; ++q) acc += trow[b0 + q] * __ldg(&w[q]);
    float v;
    if (t < p.T) {
      if (p.log_mode == 2) v = 10.0f * log10f(fmaxf(acc, 1e-10f));  // librosa.power_to_db(ref=1, amin=1e-10, top_db=None)
      else v = p.log_mode ? logf(zero_handling(acc)) : acc;
    } else {
      v = p.log_mode ? logf(kEps) : 0.f;  // the reference's zero tail
    }
    if (t < p.T_out) p.mel_out[(row * p.nf + m) * p.T_out + t] = v;
  }
}

std::mutex g_mu;
std::map<std::pair<int, int>, float*> g_tw;  // W_Nc table per (device, R)

}  // namespace

template <int TF>
static int launch_mel_tf(const float* power, const FrameParams& p, int64_t row0, int64_t n_rows, int F, void* stream) {
  const size_t smem = (size_t)TF * (F | 1) * sizeof(float);
  if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_mel_from_power<TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 g((unsigned)((p.T_out + TF - 1) / TF), (unsigned)n_rows);
  k_mel_from_power<TF><<<g, 256, smem, (cudaStream_t)stream>>>(power, p, row0, F);
  count_launch();
  return (int)cudaGetLastError();
}

// banded mel reduction of `n_rows` rows of |X|^2 ([n_rows][p.T][F], rows row0.. of the call) into p.mel_out
int launch_mel_from_power(const float* power, const FrameParams& p, int64_t row0, int64_t n_rows, int F, void* stream) {
  if (n_rows <= 0 || p.T_out <= 0) return 0;
  static const int tf_env = getenv("SS_MEL_TF") ? atoi(getenv("SS_MEL_TF")) : 0;  // A/B switch for measurements
  // measured (2,048 rows): F = 1025: 2.00 / 2.30 / 3.46 ms with 8 / 16 / 32 frames per tile, F = 513: 1.89 / 2.02 / 2.47,
  // F = 257: 2.19 / 2.14 / 2.26 -- small tiles keep several blocks per SM in flight while one loads
  const int tf = tf_env ? tf_env : (F > 512 ? 8 : 16);
  if (tf == 8) return launch_mel_tf<8>(power, p, row0, n_rows, F, stream);
  return tf == 16 ? launch_mel_tf<16>(power, p, row0, n_rows, F, stream) : launch_mel_tf<kMelFrames>(power, p, row0, n_rows, F, stream);
}

int64_t power_slab_bytes() {
  const char* e = getenv("SS_POWER_SLAB_KB");  // read per call: the tests toggle it
  const int64_t kb = e ? atoll(e) : 0;
  return kb > 0 ? kb << 10 : 256LL << 20;
}

bool stft512_supported(const Derived& d) { return (d.W == 256 || d.W == 512 || d.W == 1024) && d.stft_ok; }

// scratch: power rows of the rows of one launch; sized by stft512_scratch_bytes for `rows_per_launch` rows
size_t stft512_scratch_bytes(const FrameParams& p, int64_t* rows_per_launch) {
  const int64_t per_row = std::max<int64_t>(1, p.T) * (p.N / 2 + 1) * (int64_t)sizeof(float);
  int64_t rpl = std::max<int64_t>(1, power_slab_bytes() / per_row);
  rpl = std::min<int64_t>(rpl, std::max<int64_t>(1, p.n_rows));
  rpl = std::min<int64_t>(rpl, 65535);  // gridDim.y of the mel kernel
  if (rows_per_launch) *rows_per_launch = rpl;
  return (size_t)(rpl * per_row);
}

template <int R>
static int launch_stft_r(const FrameParams& p, float* scratch, int sm_count, void* stream) {
  using G = Geo<R>;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  const float* tw = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    float*& t = g_tw[{dev, R}];
    if (!t) {
      std::vector<float> h(G::kTabTw, 0.f);
      const double pi = 3.14159265358979323846264338327950288;
      for (int u = 0; u < 16; ++u)
        for (int k1 = 0; k1 < R; ++k1) {
          const double ang = -2.0 * pi * (double)(u * k1) / (double)G::kNc;
          h[u * G::kTwRow + 2 * k1] = (float)std::cos(ang);
          h[u * G::kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
        }
      if ((e = cudaMalloc(&t, G::kTabTw * sizeof(float))) != cudaSuccess) return (int)e;
      if ((e = cudaMemcpy(t, h.data(), G::kTabTw * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    }
    tw = t;
  }
  const bool librosa = p.mode == kModeLibrosa;
  if (G::kSmem > 48 * 1024) {
    e = librosa ? cudaFuncSetAttribute(k_stft512<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem)
                : cudaFuncSetAttribute(k_stft512<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem);
    if (e != cudaSuccess) return (int)e;
  }
  int64_t rpl = p.n_rows;
  if (p.mel_out) stft512_scratch_bytes(p, &rpl);
  cudaStream_t s = (cudaStream_t)stream;
  for (int64_t r0 = 0; r0 < p.n_rows; r0 += rpl) {
    const int64_t nr = std::min(rpl, p.n_rows - r0);
    const int64_t pairs = nr * ((p.T_out + 1) / 2);
    const int64_t grid = std::min<int64_t>((pairs + kWarps - 1) / kWarps, 16LL * sm_count);
    float* pw = p.mel_out ? scratch : nullptr;
    if (librosa) k_stft512<R, true><<<(unsigned)grid, kWarps * 32, G::kSmem, s>>>(p, tw, pw, r0, nr);
    else k_stft512<R, false><<<(unsigned)grid, kWarps * 32, G::kSmem, s>>>(p, tw, pw, r0, nr);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    if (p.mel_out) {
      const int me = launch_mel_from_power(scratch, p, r0, nr, G::kF, stream);
      if (me) return me;
    }
  }
  note_kernel(R == 16 ? "k_stft512" : R == 32 ? "k_stft1024" : "k_stft256");
  return 0;
}

int launch_stft512(const FrameParams& p, float* scratch, int sm_count, void* stream) {
  if (p.N != 256 && p.N != 512 && p.N != 1024) return -1000;
  if (p.n_rows <= 0 || p.T_out <= 0) return 0;
  if (p.mel_out && !scratch) return -1000;
  if (p.N == 256) return launch_stft_r<8>(p, scratch, sm_count, stream);
  return p.N == 512 ? launch_stft_r<16>(p, scratch, sm_count, stream) : launch_stft_r<32>(p, scratch, sm_count, stream);
}

}  // namespace ss
