// Fused mel-spectrogram kernel for the STFT path with fft_points = 2048 (config c3: 22.05 kHz,
// hop 512, 128 mels):  stream framing (functions.rs:86-160) -> Vorbis window -> real FFT(2048) ->
// |X wnorm|^2 -> banded mel (feature.rs:163-174) -> optional ln(zero_handling) -> time-minor store.
//
// Mapping (DESIGN.md "stft2048"):
//   * one real frame = one 1024-point complex FFT of z[m] = (x w)[2m] + i (x w)[2m+1] computed by
//     ONE warp, 32 complex points per lane (coalesced 256-byte float2 loads per step);
//   * 1024 = 32 x 32 (four-step): 32-point register FFT, W1024^{u k1} twiddle from a shared table,
//     one transpose through a warp-private padded shared-memory tile, second 32-point register FFT;
//   * lane u then holds Z[u + 32 k2]; its mirror Z[1024 - k] lives in lane 32 - u, which hands it
//     over with warp shuffles (16 complex values per lane), so the real-FFT split needs no second
//     shared-memory pass.  Only bins <= 512 are produced (no filter reaches higher, quirk Q8);
//   * a CTA of 16 warps covers 32 consecutive frames per round (2 per warp) and stages them in a
//     double-buffered [nf][32] tile: the time-minor output is written in 128-byte runs with one
//     __syncthreads per round.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

constexpr int kW = 2048, kM = 1024, kR = 32;
constexpr int kWarps = 16, kThreads = kWarps * 32, kFpw = 2, kFpr = kFpw * kWarps;  // 32 frames per round
constexpr int kXs = 34;                  // transpose row stride in float2 (272 B: 16-byte rows, conflict-free LDS.128)
constexpr int kWarpFloats = kR * kXs * 2;  // 2176 floats per warp
constexpr int kPs = 516;                 // P row: bins 0..512 + zero pad to 515
constexpr int kStashRow = kFpr + 1;
constexpr int kNfMax = 128, kMaxIt = 48;
constexpr int kTwRow = 68;               // floats per lane row of the W1024 table (32 float2 + pad)
constexpr int kTabTw = 0;                          // [32][kTwRow]
constexpr int kTabTw2048 = kTabTw + kR * kTwRow;   // [512] float2 W2048^k
constexpr int kTabWin = kTabTw2048 + 2 * 512;      // [1024] float2 window pairs
constexpr int kTabFixed = kTabWin + 2 * kM;        // then int2 desc[nit][32], float4 w[nit][32]
static_assert(kTabTw2048 % 4 == 0 && kTabWin % 4 == 0 && kTabFixed % 4 == 0, "table alignment");
static_assert(kPs <= kWarpFloats, "P overlays the transpose tile");

__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }

struct Stft2048Args {
  FrameParams p;
  const float* tab;
  int64_t n_items;
  int rounds_per_row;
  int nit, tab_floats;
  int vec2;
  int64_t row0 = 0;        // first row of this launch (MODE 2 runs in slabs of rows)
  float* power = nullptr;  // MODE 2: [rows of the slab][T][1025]
};

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

// MODE 0: fused banded mel (first generation; serves shapes k_stft2048m declines).  MODE 1: complex output of
// functions::stft1/2.  MODE 2: |X wnorm|^2 rows of all 1025 bins into a scratch for k_mel_from_power (the librosa-
// compatible mode, whose Slaney bank spans the whole spectrum); framing by p.mode: stream model or centred / reflected.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) k_stft2048(const Stft2048Args a) {
  constexpr bool STFT = MODE != 0;
  extern __shared__ __align__(16) float smem[];
  float* tab_s = smem;
  float* warp_s = smem + a.tab_floats + (threadIdx.x >> 5) * kWarpFloats;
  float* stash0 = smem + a.tab_floats + kWarps * kWarpFloats;  // 2 x [nf][kStashRow]
  const FrameParams& p = a.p;
  const int nf = p.nf;
  for (int i = threadIdx.x; i < a.tab_floats; i += kThreads) tab_s[i] = __ldg(a.tab + i);
  __syncthreads();

  const int u = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool isz = (u == 0);
  const int partner = (32 - u) & 31;
  const float4* twrow = reinterpret_cast<const float4*>(tab_s + kTabTw + u * kTwRow);
  const float2* twk = reinterpret_cast<const float2*>(tab_s + kTabTw2048) + u;  // W2048^{u + 32 k2}
  const float2* win2 = reinterpret_cast<const float2*>(tab_s + kTabWin) + u;
  const int2* desc = reinterpret_cast<const int2*>(tab_s + kTabFixed) + u;
  const float4* pw = reinterpret_cast<const float4*>(tab_s + kTabFixed + a.nit * 2 * 32) + u;
  float2* xch = reinterpret_cast<float2*>(warp_s);
  float* Pf = warp_s;  // overlays the transpose tile once it is consumed
  float* pA = Pf + u;

  const int64_t span = p.Tp * (int64_t)p.S;
  const int64_t memlen = p.N - p.S;
  const bool librosa = p.mode == kModeLibrosa;
  const int c0 = librosa ? -kW / 2 : (p.n_pad + 1) * p.S - p.N;
  float* stash = stash0;

  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int64_t row = a.row0 + item / a.rounds_per_row;
    const int64_t t_round = (item - (row - a.row0) * a.rounds_per_row) * kFpr;
    const float* rowp = p.pcm + row * p.row_stride;

#pragma unroll 1
    for (int f = 0; f < kFpw; ++f) {
      const int64_t t = t_round + warp * kFpw + f;
      const bool computed = t < p.T;

      // ---- framing + window: v[j] = (x w)[2m] + i (x w)[2m+1], m = u + 32 j
      float2 v[kR];
      {
        const int64_t o0 = t * p.S + c0;
        if (!computed) {
          sfor<0, kR>([&](auto jj) { v[decltype(jj)::value] = make_float2(0.f, 0.f); });
        } else if (o0 >= 0 && o0 + kW <= p.row_len && a.vec2) {
          const float* px = rowp + o0 + 2 * u;
          sfor<0, kR>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const float2 x = __ldg(reinterpret_cast<const float2*>(px + 64 * j));
            const float2 w = win2[32 * j];
            v[j] = __fmul2_rn(x, w);
          });
        } else {
          const int64_t q0 = memlen + row * span + o0 + 2 * u;
          sfor<0, kR>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const float2 w = win2[32 * j];
            const int64_t o = o0 + 2 * u + 64 * j;
            float xa, xb;
            if (librosa) {  // librosa.stft(center=True, pad_mode="reflect")
              int64_t oa = o, ob = o + 1;
              if (oa < 0) oa = -oa;
              if (ob < 0) ob = -ob;
              if (oa >= p.row_len) oa = 2 * (p.row_len - 1) - oa;
              if (ob >= p.row_len) ob = 2 * (p.row_len - 1) - ob;
              xa = rowp[oa];
              xb = rowp[ob];
            } else {
              xa = o >= 0 ? (o < p.row_len ? rowp[o] : 0.f) : stream_at(p, q0 + 64 * j);
              xb = o + 1 >= 0 ? (o + 1 < p.row_len ? rowp[o + 1] : 0.f) : stream_at(p, q0 + 64 * j + 1);
            }
            v[j] = make_float2(xa * w.x, xb * w.y);
          });
        }
      }
      // ---- stage 1: 32-point FFT over j
      ssfft::fft_pow2<kR>(v);
      // ---- twiddle W1024^{u k1} and transpose
      __syncwarp();
      sfor<0, kR / 2>([&](auto kk) {
        constexpr int k1 = 2 * decltype(kk)::value;
        const float4 w = twrow[k1 / 2];
        float2 y0 = v[k1], y1 = v[k1 + 1];
        if constexpr (k1 > 0) y0 = ssfft::cmulf(y0, make_float2(w.x, w.y));
        y1 = ssfft::cmulf(y1, make_float2(w.z, w.w));
        xch[k1 * kXs + u] = y0;
        xch[(k1 + 1) * kXs + u] = y1;
      });
      __syncwarp();
      float2 g[kR];
      {
        const float4* r4 = reinterpret_cast<const float4*>(xch + u * kXs);
        sfor<0, kR / 2>([&](auto tt) {
          constexpr int i = decltype(tt)::value;
          const float4 x = r4[i];
          g[2 * i] = make_float2(x.x, x.y);
          g[2 * i + 1] = make_float2(x.z, x.w);
        });
      }
      __syncwarp();
      // ---- stage 2: 32-point FFT over the lane index.  g[k2] = Z[u + 32 k2]
      ssfft::fft_pow2<kR>(g);

      // ---- real-FFT post-processing for k = u + 32 k2 <= 511 (k2 < 16); Z[1024 - k] is
      // g[31 - k2] of lane 32 - u (lane 0: its own g[32 - k2]); |X|^2 x 4 goes to P[k]
      sfor<0, 16>([&](auto kk) {
        constexpr int k2 = decltype(kk)::value;
        float2 zb;
        zb.x = __shfl_sync(0xffffffffu, g[31 - k2].x, partner);
        zb.y = __shfl_sync(0xffffffffu, g[31 - k2].y, partner);
        if (isz) zb = g[(32 - k2) & 31];
        const float2 za = g[k2];
        const float2 w = twk[32 * k2];
        const float2 e = __fadd2_rn(za, make_float2(zb.x, -zb.y));
        const float2 od = __fadd2_rn(make_float2(za.y, -za.x), make_float2(zb.y, zb.x));
        const float2 tt = ssfft::cmulf(od, w);
        const float2 xa = ssfft::cadd(e, tt);  // 2 X[k]
        if constexpr (MODE == 2) {
          if (computed) {
            float* po = a.power + ((row - a.row0) * p.T + t) * (int64_t)(kM + 1);
            const float2 xb = ssfft::csub(e, tt);
            const float hs2 = 0.25f * p.wnorm * p.wnorm;
            po[u + 32 * k2] = hs2 * fmaf(xa.x, xa.x, xa.y * xa.y);
            po[kM - (u + 32 * k2)] = hs2 * fmaf(xb.x, xb.x, xb.y * xb.y);
          }
        } else if constexpr (STFT) {
          // [row][T_out][1025] (re, im): X wnorm (functions.rs:166-169); zeros in the reference's zero tail (v = 0)
          if (t < p.T_out) {
            float2* so = p.stft_out + (row * p.T_out + t) * (int64_t)(kM + 1);
            const float2 xb = ssfft::csub(e, tt);  // 2 conj X[1024 - k]
            const float hs = 0.5f * p.wnorm;
            so[u + 32 * k2] = make_float2(xa.x * hs, xa.y * hs);
            so[kM - (u + 32 * k2)] = make_float2(xb.x * hs, -xb.y * hs);
          }
        } else {
          pA[32 * k2] = fmaf(xa.x, xa.x, xa.y * xa.y);
        }
      });
      if constexpr (MODE == 2) {
        if (isz && computed)  // X[512] = conj Z[512]
          a.power[((row - a.row0) * p.T + t) * (int64_t)(kM + 1) + 512] = p.wnorm * p.wnorm * fmaf(g[16].x, g[16].x, g[16].y * g[16].y);
        continue;
      }
      if constexpr (MODE == 1) {
        if (isz && t < p.T_out)  // X[512] = conj Z[512]
          p.stft_out[(row * p.T_out + t) * (int64_t)(kM + 1) + 512] = make_float2(g[16].x * p.wnorm, -g[16].y * p.wnorm);
        continue;
      }
      if (isz) {
        Pf[512] = 4.0f * fmaf(g[16].x, g[16].x, g[16].y * g[16].y);  // X[512] = conj Z[512]
        Pf[513] = 0.f; Pf[514] = 0.f; Pf[515] = 0.f;
      }
      __syncwarp();

      // ---- banded mel program over 32 lanes -> output tile [nf][32]
      {
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float* out_col = stash + warp * kFpw + f;
#pragma unroll 1
        for (int it = 0; it < a.nit; ++it) {
          const int2 d = desc[it * 32];
          const float4 w = pw[it * 32];
          const float4 pv = *reinterpret_cast<const float4*>(Pf + d.x);
          acc4.x = fmaf(w.x, pv.x, acc4.x); acc4.y = fmaf(w.y, pv.y, acc4.y);
          acc4.z = fmaf(w.z, pv.z, acc4.z); acc4.w = fmaf(w.w, pv.w, acc4.w);
          if (d.y != -2) {
            if (d.y >= 0) {
              float acc = (acc4.x + acc4.y) + (acc4.z + acc4.w);
              if (!computed) acc = 0.f;
              out_col[d.y * kStashRow] = p.log_mode ? __logf(zero_handling(acc)) : acc;
            }
            acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      __syncwarp();  // P is overwritten by the next frame's transpose
    }
    if constexpr (STFT) continue;
    __syncthreads();  // tile complete; everyone is done reading the other tile
    {
      const int64_t nvalid = min((int64_t)kFpr, p.T_out - t_round);
      float* dst = p.mel_out + row * nf * p.T_out + t_round;
      for (int idx = threadIdx.x; idx < nf * kFpr; idx += kThreads) {
        const int m = idx >> 5, ff = idx & 31;
        if (ff < nvalid) dst[(int64_t)m * p.T_out + ff] = stash[m * kStashRow + ff];
      }
    }
    stash = (stash == stash0) ? stash0 + nf * kStashRow : stash0;
  }
}

struct MelProgram2048 {
  int nit = 0;
  std::vector<int> desc;
  std::vector<float> w;
};

static bool build_program(const Derived& d, const MelBank& mb, float pscale, MelProgram2048* mp) {
  const int nf = d.nf, NL = 32;
  if (mb.max_bin > 512) return false;
  std::vector<int> order(nf);
  for (int i = 0; i < nf; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return mb.len[x] > mb.len[y]; });
  mp->nit = 0;
  const int ns = (nf + NL - 1) / NL;
  for (int s = 0; s < ns; ++s) {
    int its = 1, b0[32];
    for (int l = 0; l < NL; ++l) {
      b0[l] = 0;
      if (s * NL + l >= nf) continue;
      const int f = order[s * NL + l];
      if (mb.len[f] == 0) continue;
      b0[l] = mb.start[f] & ~3;
      its = std::max(its, (mb.start[f] + mb.len[f] - b0[l] + 3) / 4);
    }
    if (4 * its > kPs) return false;
    for (int l = 0; l < NL; ++l) b0[l] = std::max(0, std::min(b0[l], kPs - 4 * its));
    for (int it = 0; it < its; ++it)
      for (int l = 0; l < NL; ++l) {
        const bool used = s * NL + l < nf;
        const int f = used ? order[s * NL + l] : -1;
        mp->desc.push_back(b0[l] + 4 * it);
        mp->desc.push_back(it == its - 1 ? (used ? f : -1) : -2);
        for (int k = 0; k < 4; ++k) {
          const int bin = b0[l] + 4 * it + k;
          float w = 0.f;
          if (used && bin >= mb.start[f] && bin < mb.start[f] + mb.len[f]) w = mb.w[mb.wofs[f] + bin - mb.start[f]];
          mp->w.push_back(w * pscale);
        }
      }
    mp->nit += its;
  }
  return mp->nit <= kMaxIt;
}

}  // namespace

struct Stft2048Tables {
  float* d_tab = nullptr;
  int nit = 0, tab_floats = 0;
};

bool stft2048_supported(const Derived& d, const MelBank& mb) {
  if (d.W != kW || !d.stft_ok) return false;
  if (d.librosa) return true;  // |X|^2 rows + k_mel_from_power: no mel program needed
  if (d.nf > kNfMax) return false;
  MelProgram2048 mp;
  return build_program(d, mb, 1.0f, &mp);
}

int stft2048_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft2048Tables** out) {
  MelProgram2048 mp;
  const float pscale = 0.25f * d.wnorm * d.wnorm;  // (1/2)^2 of E, O and wnorm^2 (functions.rs:166-169)
  if (d.librosa) {
    mp = MelProgram2048();
  } else if (!build_program(d, mb, pscale, &mp)) {
    return (int)cudaErrorInvalidValue;
  }
  std::vector<float> h(kTabFixed + (size_t)mp.nit * 32 * 6, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  for (int u = 0; u < kR; ++u)
    for (int k1 = 0; k1 < kR; ++k1) {
      const double ang = -2.0 * pi * (double)(u * k1) / (double)kM;
      h[kTabTw + u * kTwRow + 2 * k1] = (float)std::cos(ang);
      h[kTabTw + u * kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
    }
  for (int k = 0; k < 512; ++k) {
    const double ang = -2.0 * pi * (double)k / (double)kW;
    h[kTabTw2048 + 2 * k] = (float)std::cos(ang);
    h[kTabTw2048 + 2 * k + 1] = (float)std::sin(ang);
  }
  for (int n = 0; n < kW; ++n) h[kTabWin + n] = window[n];
  int* desc = reinterpret_cast<int*>(h.data() + kTabFixed);
  for (size_t i = 0; i < mp.desc.size(); ++i) desc[i] = mp.desc[i];
  float* w = h.data() + kTabFixed + (size_t)mp.nit * 32 * 2;
  for (size_t i = 0; i < mp.w.size(); ++i) w[i] = mp.w[i];
  Stft2048Tables* t = new Stft2048Tables;
  t->nit = mp.nit;
  t->tab_floats = (int)h.size();
  cudaError_t e = cudaMalloc(&t->d_tab, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return (int)e; }
  *out = t;
  return 0;
}

void stft2048_free_tables(Stft2048Tables* t) {
  if (!t) return;
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

int launch_stft2048(const FrameParams& p, const Stft2048Tables* t, int sm_count, void* stream) {
  if (p.mode == kModeLibrosa) return -1000;                                     // launch_stft2048_power serves that mode
  if (p.N != kW || (!p.mel_out == !p.stft_out) || p.T_out < 1) return -1000;  // exactly one of the two outputs
  Stft2048Args a;
  a.p = p;
  a.tab = t->d_tab;
  a.nit = t->nit;
  a.tab_floats = t->tab_floats;
  a.rounds_per_row = (int)((p.T_out + kFpr - 1) / kFpr);
  a.n_items = p.n_rows * (int64_t)a.rounds_per_row;
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  const size_t smem = (size_t)(a.tab_floats + kWarps * kWarpFloats + 2 * p.nf * kStashRow) * sizeof(float);
  if (smem > 227 * 1024) return -1000;
  const int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sm_count);
  if (grid <= 0) return 0;
  cudaError_t e;
  if (p.stft_out) {
    e = cudaFuncSetAttribute(k_stft2048<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_stft2048<1><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);
  } else {
    e = cudaFuncSetAttribute(k_stft2048<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_stft2048<0><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);
  }
  count_launch();
  note_kernel("k_stft2048");
  return (int)cudaGetLastError();
}

// the librosa-compatible mel spectrogram for n_fft = 2048 (librosa's default): |X|^2 rows of a slab of rows into
// `scratch` (stft2048_power_scratch_bytes), then k_mel_from_power (ss_kernels_stft512.cu) over the Slaney bank
size_t stft2048_power_scratch_bytes(const FrameParams& p, int64_t* rows_per_launch) {
  const int64_t per_row = std::max<int64_t>(1, p.T) * (kM + 1) * (int64_t)sizeof(float);
  int64_t rpl = std::max<int64_t>(1, power_slab_bytes() / per_row);
  rpl = std::min<int64_t>(rpl, std::max<int64_t>(1, p.n_rows));
  rpl = std::min<int64_t>(rpl, 65535);
  if (rows_per_launch) *rows_per_launch = rpl;
  return (size_t)(rpl * per_row);
}

int launch_stft2048_power(const FrameParams& p, const Stft2048Tables* t, float* scratch, int sm_count, void* stream) {
  if (p.N != kW || !p.mel_out || !scratch || p.T_out < 1 || p.n_rows <= 0) return -1000;
  int64_t rpl = 1;
  stft2048_power_scratch_bytes(p, &rpl);
  Stft2048Args a;
  a.p = p;
  a.tab = t->d_tab;
  a.nit = 0;
  a.tab_floats = kTabFixed;  // no mel program behind the fixed tables
  a.rounds_per_row = (int)((p.T_out + kFpr - 1) / kFpr);
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0) &&
           (p.mode != kModeLibrosa || (kW / 2) % 2 == 0);
  a.power = scratch;
  const size_t smem = (size_t)(a.tab_floats + kWarps * kWarpFloats) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k_stft2048<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  for (int64_t r0 = 0; r0 < p.n_rows; r0 += rpl) {
    const int64_t nr = std::min(rpl, p.n_rows - r0);
    a.row0 = r0;
    a.n_items = nr * (int64_t)a.rounds_per_row;
    const int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sm_count);
    k_stft2048<2><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    const int me = launch_mel_from_power(scratch, p, r0, nr, kM + 1, stream);
    if (me) return me;
  }
  note_kernel("k_stft2048");
  return 0;
}

}  // namespace ss
