// Fused mel-spectrogram kernel for the STFT path with fft_points = 400 (config c4, the
// Whisper-shaped log-mel: 16 kHz, hop 160, 80 mels):
//   stream framing (functions.rs:86-160) -> Vorbis window -> real FFT(400) -> |X wnorm|^2 ->
//   banded mel (feature.rs:163-174) -> optional ln(zero_handling) -> time-minor store.
// Replaces feature::mel_spectrogram{1,2} for this FFT size; functions::stft{1,2} (complex output)
// stay on the generic kernel.
//
// Mapping (DESIGN.md "stft400"):
//   * one real frame = one 200-point complex FFT of z[m] = (x w)[2m] + i (x w)[2m+1] plus the
//     real-FFT post-processing, computed by 10 lanes holding 20 complex points each; a warp owns 3
//     consecutive frames (lanes 30, 31 shadow lanes 20, 21 and never store);
//   * 200 = 20 x 10 (four-step): stage 1 is a 20-point FFT in registers (radix 5 then 2 x radix 2,
//     compile-time twiddles), then the W200^{u k1} twiddle from a shared table, one transpose
//     through a warp-private conflict-free shared-memory tile, stage 2 is two 10-point register
//     FFTs per lane;
//   * lane u owns output groups k1 = u and 20 - u (u = 0: 0 and 10), so Z[k] and Z[200-k] meet in
//     one thread and X[k] = E + W400^k O needs no second exchange;
//   * the CTA (8 warps = 24 consecutive frames per round) stages log-mel values in a
//     double-buffered [nf][24] tile so the time-minor output is written in 96-byte runs with one
//     __syncthreads per round.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

constexpr int kW = 400, kM = 200, kT = 10, kR1 = 20;
constexpr int kWarps = 8, kThreads = kWarps * 32, kFpw = 3, kFpr = kFpw * kWarps;  // 24 frames per round
constexpr int kXchFrame = 202;           // float2 per frame: 20 rows x 10 + 2 (frame stride = 10 mod 16 banks)
constexpr int kWarpFloats = kFpw * kXchFrame * 2;
constexpr int kPs = 104;                 // P row stride: bins 0..100 + zero pad to 103
constexpr int kStashRow = kFpr + 1;      // odd row stride of the output tile
constexpr int kNfMax = 128, kMaxIt = 64;
constexpr int kTwRow = 44;               // floats per lane row of the W200 table (20 float2 + pad)
// table blob (floats)
constexpr int kTabTw = 0;                         // [10][kTwRow]
constexpr int kTabTw400 = kTabTw + kT * kTwRow;   // [200] float2 W400^k
constexpr int kTabWin = kTabTw400 + 2 * kM;       // [200] float2 window pairs
constexpr int kTabFixed = kTabWin + 2 * kM;       // then int2 desc[nit][10], float4 w[nit][10]
static_assert(kTabTw400 % 4 == 0 && kTabWin % 4 == 0 && kTabFixed % 4 == 0, "table alignment");

__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }

struct Stft400Args {
  FrameParams p;
  const float* tab;
  int64_t n_items;      // rows * rounds_per_row
  int rounds_per_row;
  int nit, tab_floats;
  int vec2;             // frames start 8-byte aligned
  float pscale;         // (wnorm / 2)^2: the 1/2 of E, O and the STFT normalisation (functions.rs:166-169)
};

// stream sample of the virtual concatenation  mem_in ++ pad(row 0) ++ pad(row 1) ...
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

__device__ __forceinline__ float2 pair_power(float2 za, float2 zb, float2 w) {
  // E' = za + conj zb, O' = (za - conj zb)/i, T' = W400^k O'; returns |E'+T'|^2, |E'-T'|^2
  const float2 e = __fadd2_rn(za, make_float2(zb.x, -zb.y));
  const float2 od = __fadd2_rn(make_float2(za.y, -za.x), make_float2(zb.y, zb.x));
  const float2 tt = ssfft::cmulf(od, w);
  const float2 xa = ssfft::cadd(e, tt), xb = ssfft::csub(e, tt);
  return make_float2(fmaf(xa.x, xa.x, xa.y * xa.y), fmaf(xb.x, xb.x, xb.y * xb.y));
}

#ifndef SS_STFT400_MINB
#define SS_STFT400_MINB 3  // 80 registers, 24 warps/SM: measured 8 % faster than 2 CTAs at 126 registers
#endif
__global__ void __launch_bounds__(kThreads, SS_STFT400_MINB) k_stft400(const Stft400Args a) {
  extern __shared__ __align__(16) float smem[];
  float* tab_s = smem;
  float* warp_s = smem + a.tab_floats + (threadIdx.x >> 5) * kWarpFloats;
  float* stash0 = smem + a.tab_floats + kWarps * kWarpFloats;  // 2 x [nf][kStashRow]
  const FrameParams& p = a.p;
  const int nf = p.nf;
  for (int i = threadIdx.x; i < a.tab_floats; i += kThreads) tab_s[i] = __ldg(a.tab + i);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sl = lane < 30 ? lane : lane - 10;  // lanes 30, 31 shadow lanes 20, 21
  const int fi = sl / 10, u = sl - fi * 10;
  const bool real_lane = lane < 30;
  const bool isz = (u == 0);
  const int k1B = isz ? 10 : 20 - u;
  const int kb = isz ? -90 : u;  // k of slots 5..9 is kb + 20 e (u == 0: 10 + 20 (e - 5))

  const float4* twrow = reinterpret_cast<const float4*>(tab_s + kTabTw + u * kTwRow);
  const float2* tw400 = reinterpret_cast<const float2*>(tab_s + kTabTw400);
  const float2* win2 = reinterpret_cast<const float2*>(tab_s + kTabWin) + u;
  const int2* desc = reinterpret_cast<const int2*>(tab_s + kTabFixed) + u;
  const float4* pw = reinterpret_cast<const float4*>(tab_s + kTabFixed + a.nit * 2 * kT) + u;

  float2* xch = reinterpret_cast<float2*>(warp_s) + fi * kXchFrame;
  float* Pf = warp_s + fi * (2 * kXchFrame);  // overlays this frame's transpose tile once it is consumed
  const float2* twA = tw400 + u;
  const float2* twB = tw400 + kb;
  float* pA = Pf + u;
  float* plo[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) plo[i] = isz ? Pf + 10 + 20 * i : Pf + 200 - u - 20 * (i + 5);

  const int64_t span = p.Tp * (int64_t)p.S;
  const int64_t memlen = p.N - p.S;
  const int c0 = (p.n_pad + 1) * p.S - p.N;  // row-relative position of frame 0's first sample (<= 0)
  float* stash = stash0;

  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int64_t row = item / a.rounds_per_row;
    const int64_t t_round = (item - row * a.rounds_per_row) * kFpr;
    const int64_t t = t_round + warp * kFpw + fi;  // this lane group's output frame
    const bool computed = t < p.T;                 // frames >= T are the reference's zero tail
    const float* rowp = p.pcm + row * p.row_stride;

    // ---- framing + window: v[j] = (x w)[2m] + i (x w)[2m+1], m = u + 10 j
    float2 v[kR1];
    {
      const int64_t o0 = t * p.S + c0;  // first sample of the frame, relative to the row start
      if (!computed) {
        sfor<0, kR1>([&](auto jj) { v[decltype(jj)::value] = make_float2(0.f, 0.f); });
      } else if (o0 >= 0 && o0 + kW <= p.row_len) {
        const float* px = rowp + o0 + 2 * u;
        if (a.vec2) {
          sfor<0, kR1>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const float2 x = __ldg(reinterpret_cast<const float2*>(px + 20 * j));
            const float2 w = win2[10 * j];
            v[j] = __fmul2_rn(x, w);
          });
        } else {
          sfor<0, kR1>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const float2 w = win2[10 * j];
            v[j] = make_float2(__ldg(px + 20 * j) * w.x, __ldg(px + 20 * j + 1) * w.y);
          });
        }
      } else {  // frame touches the overlap memory / previous row / zero padding
        const int64_t q0 = memlen + row * span + o0 + 2 * u;
        sfor<0, kR1>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          const float2 w = win2[10 * j];
          const int64_t q = q0 + 20 * j;
          // samples of this row beyond row_len are zero padding; earlier positions come from the stream
          const int64_t o = o0 + 2 * u + 20 * j;
          const float xa = o >= 0 ? (o < p.row_len ? rowp[o] : 0.f) : stream_at(p, q);
          const float xb = o + 1 >= 0 ? (o + 1 < p.row_len ? rowp[o + 1] : 0.f) : stream_at(p, q + 1);
          v[j] = make_float2(xa * w.x, xb * w.y);
        });
      }
    }
    // ---- stage 1: 20-point FFT over j, in registers
    ssfft::fft_5x2<kR1>(v);
    // ---- twiddle W200^{u k1} and transpose
    __syncwarp();
    sfor<0, kR1 / 2>([&](auto kk) {
      constexpr int k1 = 2 * decltype(kk)::value;
      const float4 w = twrow[k1 / 2];
      float2 y0 = v[k1], y1 = v[k1 + 1];
      if constexpr (k1 > 0) y0 = ssfft::cmulf(y0, make_float2(w.x, w.y));
      y1 = ssfft::cmulf(y1, make_float2(w.z, w.w));
      xch[k1 * kT + u] = y0;
      xch[(k1 + 1) * kT + u] = y1;
    });
    __syncwarp();
    float2 gA[kT], gB[kT];
    {
      const float4* rA = reinterpret_cast<const float4*>(xch + u * kT);
      const float4* rB = reinterpret_cast<const float4*>(xch + k1B * kT);
      sfor<0, kT / 2>([&](auto tt) {
        constexpr int i = decltype(tt)::value;
        const float4 xa = rA[i], xb = rB[i];
        gA[2 * i] = make_float2(xa.x, xa.y); gA[2 * i + 1] = make_float2(xa.z, xa.w);
        gB[2 * i] = make_float2(xb.x, xb.y); gB[2 * i + 1] = make_float2(xb.z, xb.w);
      });
    }
    __syncwarp();
    // ---- stage 2: 10-point FFTs over the lane index.  g?[k2] = Z[k1? + 20 k2]
    ssfft::fft_5x2<kT>(gA);
    ssfft::fft_5x2<kT>(gB);

    // ---- real-FFT post-processing, |X|^2 of the bins a mel filter can reach (<= 100).
    // u != 0: slot e pairs Z[u+20e] (gA[e]) with Z[200-(u+20e)] = gB[9-e].
    // u == 0: groups 0 and 10 mirror onto themselves: e=0 -> k=0; e=1..4 -> (gA[e], gA[10-e]);
    //         e=5..9 -> (gB[e-5], gB[14-e]) with k = 10 + 20 (e-5); bin 100 = |gA[5]|^2 below.
    sfor<0, 10>([&](auto ee) {
      constexpr int e = decltype(ee)::value;
      float2 za, zb;
      if constexpr (e == 0) {
        za = gA[0];
        zb = isz ? gA[0] : gB[9];
      } else if constexpr (e < 5) {
        za = gA[e];
        zb = isz ? gA[10 - e] : gB[9 - e];
      } else {
        za = isz ? gB[e - 5] : gA[e];
        zb = isz ? gB[14 - e] : gB[9 - e];
      }
      const float2 m = pair_power(za, zb, (e < 5 ? twA : twB)[20 * e]);
      if constexpr (e < 5) pA[20 * e] = m.x;        // bin u + 20 e <= 89
      else *plo[e - 5] = isz ? m.x : m.y;           // the bin of the pair that is <= 100
    });
    if (isz) {
      Pf[100] = 4.0f * fmaf(gA[5].x, gA[5].x, gA[5].y * gA[5].y);  // X[100] = conj Z[100]; x4 to match the pairs
      Pf[101] = 0.f; Pf[102] = 0.f; Pf[103] = 0.f;                // tail of the last aligned group of 4
    }
    __syncwarp();

    // ---- banded mel program (see ss_kernels_fast512b.cu) -> output tile [nf][24]
    {
      float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float* out_col = stash + warp * kFpw + fi;
#pragma unroll 1
      for (int it = 0; it < a.nit; ++it) {
        const int2 d = desc[it * kT];
        const float4 w = pw[it * kT];
        const float4 pv = *reinterpret_cast<const float4*>(Pf + d.x);
        acc4.x = fmaf(w.x, pv.x, acc4.x); acc4.y = fmaf(w.y, pv.y, acc4.y);
        acc4.z = fmaf(w.z, pv.z, acc4.z); acc4.w = fmaf(w.w, pv.w, acc4.w);
        if (d.y != -2) {
          if (d.y >= 0 && real_lane) {
            float acc = (acc4.x + acc4.y) + (acc4.z + acc4.w);
            if (!computed) acc = 0.f;  // exact zeros in the reference's zero tail
            out_col[d.y * kStashRow] = p.log_mode ? __logf(zero_handling(acc)) : acc;
          }
          acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    __syncthreads();  // tile complete; also: everyone is done reading the other tile
    // ---- time-minor store (feature.rs:173): [row][m][t], runs of 24 frames
    {
      const int64_t nvalid = min((int64_t)kFpr, p.T_out - t_round);
      float* dst = p.mel_out + row * nf * p.T_out + t_round;
      for (int idx = threadIdx.x; idx < nf * kFpr; idx += kThreads) {
        const int m = idx / kFpr, f = idx - m * kFpr;
        if (f < nvalid) dst[(int64_t)m * p.T_out + f] = stash[m * kStashRow + f];
      }
    }
    stash = (stash == stash0) ? stash0 + nf * kStashRow : stash0;
  }
}

struct MelProgram400 {
  int nit = 0;
  std::vector<int> desc;
  std::vector<float> w;
};

// same construction as the mel program of ss_kernels_fast512b.cu, for 10 lanes and bins [0, 104)
static bool build_program(const Derived& d, const MelBank& mb, float pscale, MelProgram400* mp) {
  const int nf = d.nf;
  if (mb.max_bin > 100) return false;
  std::vector<int> order(nf);
  for (int i = 0; i < nf; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return mb.len[x] > mb.len[y]; });
  mp->nit = 0;
  const int ns = (nf + kT - 1) / kT;
  for (int s = 0; s < ns; ++s) {
    int its = 1, b0[kT];
    for (int l = 0; l < kT; ++l) {
      b0[l] = 0;
      if (s * kT + l >= nf) continue;
      const int f = order[s * kT + l];
      if (mb.len[f] == 0) continue;
      b0[l] = mb.start[f] & ~3;
      its = std::max(its, (mb.start[f] + mb.len[f] - b0[l] + 3) / 4);
    }
    if (4 * its > kPs) return false;
    for (int l = 0; l < kT; ++l) b0[l] = std::max(0, std::min(b0[l], kPs - 4 * its));
    for (int it = 0; it < its; ++it)
      for (int l = 0; l < kT; ++l) {
        const bool used = s * kT + l < nf;
        const int f = used ? order[s * kT + l] : -1;
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

struct Stft400Tables {
  float* d_tab = nullptr;
  int nit = 0, tab_floats = 0;
  float pscale = 0.f;
};

bool stft400_supported(const Derived& d, const MelBank& mb) {
  if (d.librosa || d.W != kW || !d.stft_ok || d.nf > kNfMax) return false;
  MelProgram400 mp;
  return build_program(d, mb, 1.0f, &mp);
}

int stft400_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft400Tables** out) {
  MelProgram400 mp;
  const float pscale = 0.25f * d.wnorm * d.wnorm;
  if (!build_program(d, mb, pscale, &mp)) return (int)cudaErrorInvalidValue;
  std::vector<float> h(kTabFixed + (size_t)mp.nit * kT * 6, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  for (int u = 0; u < kT; ++u)
    for (int k1 = 0; k1 < kR1; ++k1) {
      const double ang = -2.0 * pi * (double)(u * k1) / (double)kM;
      h[kTabTw + u * kTwRow + 2 * k1] = (float)std::cos(ang);
      h[kTabTw + u * kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
    }
  for (int k = 0; k < kM; ++k) {
    const double ang = -2.0 * pi * (double)k / (double)kW;
    h[kTabTw400 + 2 * k] = (float)std::cos(ang);
    h[kTabTw400 + 2 * k + 1] = (float)std::sin(ang);
  }
  for (int n = 0; n < kW; ++n) h[kTabWin + n] = window[n];
  int* desc = reinterpret_cast<int*>(h.data() + kTabFixed);
  for (size_t i = 0; i < mp.desc.size(); ++i) desc[i] = mp.desc[i];
  float* w = h.data() + kTabFixed + (size_t)mp.nit * kT * 2;
  for (size_t i = 0; i < mp.w.size(); ++i) w[i] = mp.w[i];
  Stft400Tables* t = new Stft400Tables;
  t->nit = mp.nit;
  t->tab_floats = (int)h.size();
  t->pscale = pscale;
  cudaError_t e = cudaMalloc(&t->d_tab, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return (int)e; }
  *out = t;
  return 0;
}

void stft400_free_tables(Stft400Tables* t) {
  if (!t) return;
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

// mel_spectrogram only (p.mel_out set, p.stft_out null); returns -1000 if the shape is declined
int launch_stft400(const FrameParams& p, const Stft400Tables* t, int sm_count, void* stream) {
  if (p.N != kW || !p.mel_out || p.stft_out || p.T_out < 1) return -1000;
  Stft400Args a;
  a.p = p;
  a.tab = t->d_tab;
  a.nit = t->nit;
  a.tab_floats = t->tab_floats;
  a.pscale = t->pscale;
  a.rounds_per_row = (int)((p.T_out + kFpr - 1) / kFpr);
  a.n_items = p.n_rows * (int64_t)a.rounds_per_row;
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  const size_t smem = (size_t)(a.tab_floats + kWarps * kWarpFloats + 2 * p.nf * kStashRow) * sizeof(float);
  if (smem > 110 * 1024) return -1000;
  const int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sm_count * SS_STFT400_MINB);
  if (grid <= 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(k_stft400, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_stft400<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);
  count_launch();
  note_kernel("k_stft400");
  return (int)cudaGetLastError();
}

}  // namespace ss
