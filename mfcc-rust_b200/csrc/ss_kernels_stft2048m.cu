// Fused mel-spectrogram kernel for the STFT path with fft_points = 2048, second generation
// (config c3: 22.05 kHz, hop 512, 128 mels):  stream framing (functions.rs:86-160) -> Vorbis window
// -> real FFT(2048) -> |X wnorm|^2 -> mel filterbank (feature.rs:163-174) -> optional
// ln(zero_handling) -> time-minor store.
//
// FFT stage as in the first generation (ss_kernels_stft2048.cu): one warp per frame, 32 complex
// points per lane, 1024 = 32 x 32 four-step FFT in registers (packed FFMA2 butterflies), mirror bin
// by warp shuffle.  New here: the filterbank runs on the tensor cores.
//   * The CTA's 16 warps form two independent groups of 8.  A group transforms 8 consecutive frames
//     of one row per round; every warp writes the |X|^2 row of its frame into the group's [8][520]
//     tile (double-buffered, so one named barrier of 256 threads per round is enough; the two groups
//     drift apart in phase and hide each other's load and barrier latency).
//   * The tile then feeds mel^T [128 filters][8 frames] = W [128][520 bins] x P^T on
//     mma.sync.m16n8k8 TF32 split 3-way (hi x hi + lo x hi + hi x lo): 8 m-tiles of 16 filters, one
//     per warp of the group; a job visits only the 8-bin k-steps its 16 filters' band touches (75
//     of 520 blocks for the 128-filter bank).  The pre-split weight
//     fragments (77 KB) are read from global memory: they are L2-resident and every warp needs
//     only its own m-tile's share per round.  (A per-warp |X|^2 tile as in k_fast512c would be
//     16.6 KB for 8 frames of 513 bins; sixteen of them do not fit next to the transpose tiles.)
//   * The accumulator fragment holds (filter g [+ 8], frames 2t, 2t + 1): neighbours along the
//     time-minor output [row][m][t], stored straight from registers as 8-byte pairs, 32 bytes per
//     filter and job; the first generation's [nf][32] staging tile and its copy loop are gone.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

constexpr int kW = 2048, kM = 1024, kR = 32;
constexpr int kWarps = 16, kThreads = kWarps * 32;
constexpr int kGroup = 8, kGroups = kWarps / kGroup, kFpr = kGroup;  // 8 frames per group and round
constexpr int kXs = 34;                    // transpose row stride in float2 (272 B: 16-byte rows, conflict-free LDS.128)
constexpr int kWarpFloats = kR * kXs * 2;  // 2176 floats per warp
constexpr int kPs = 520;                   // P row: bins 0..512 + zero pad; 65 k-steps of 8; 520 % 32 == 8 -> conflict-free LDS.64
constexpr int kKs = kPs / 8;
constexpr int kKsTab = kKs + 1;            // bands are padded to an even number of k-steps
constexpr int kTileFloats = kFpr * kPs + 8;  // + slack for the padded k-step of the last row
constexpr int kMtMax = 8;                  // nf <= 128
constexpr int kTwRow = 68;                 // floats per lane row of the W1024 table (32 float2 + pad)
constexpr int kTabTw = 0;                          // [32][kTwRow]
constexpr int kTabTw2048 = kTabTw + kR * kTwRow;   // [512] float2 W2048^k
constexpr int kTabWin = kTabTw2048 + 2 * 512;      // [1024] float2 window pairs
constexpr int kTabFixed = kTabWin + 2 * kM;        // 5248 floats
static_assert(kTabTw2048 % 4 == 0 && kTabWin % 4 == 0 && kTabFixed % 4 == 0, "table alignment");

__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }
__device__ __forceinline__ float fast_ln(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.693147180559945309f;
}
// x = hi + lo with hi = x truncated to TF32; lo is exact in f32 and the tensor core drops its low 13 bits
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// D += A(16x8, row) * B(8x8, col), TF32 inputs, f32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
        "r"(b0), "r"(b1));
}

struct Stft2048mArgs {
  FrameParams p;
  const float* tab;        // fixed tables (twiddles, window), copied to shared memory once per CTA
  const float4* mel_frag;  // [blocks][hi, lo][32]: A fragments of the filterbank (16 filters x 8 bins), global
  int64_t n_items;         // rows * rounds_per_row
  int rounds_per_row;
  int vec2;                // frames start 8-byte aligned
  int pair_ok;             // T_out is even and the output is 8-byte aligned: float2 stores
  int n_mt;                // m-tiles of 16 filters
  int ks_lo[kMtMax], ks_n2[kMtMax], mel_base[kMtMax];
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

__global__ void __launch_bounds__(kThreads, 1) k_stft2048m(const Stft2048mArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* tab_s = smem;
  float* warp_s = smem + kTabFixed + (threadIdx.x >> 5) * kWarpFloats;
  float* tile0 = smem + kTabFixed + kWarps * kWarpFloats + (threadIdx.x >> 8) * 2 * kTileFloats;  // this group's 2 x [8][kPs] (+ slack)
  const FrameParams& p = a.p;
  const int nf = p.nf;
  for (int i = threadIdx.x; i < kTabFixed; i += kThreads) tab_s[i] = __ldg(a.tab + i);
  // the pad bins of the tile rows are never written again and must be finite
  for (int i = threadIdx.x & 255; i < 2 * kTileFloats; i += 256) tile0[i] = 0.0f;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int u = lane;
  const int g = lane >> 2, t4 = lane & 3;  // MMA stage: fragment row group, thread in group
  const bool isz = (u == 0);
  const int partner = (32 - u) & 31;
  const float4* twrow = reinterpret_cast<const float4*>(tab_s + kTabTw + u * kTwRow);
  const float2* twk = reinterpret_cast<const float2*>(tab_s + kTabTw2048) + u;  // W2048^{u + 32 k2}
  const float2* win2 = reinterpret_cast<const float2*>(tab_s + kTabWin) + u;
  float2* xch = reinterpret_cast<float2*>(warp_s);

  const int64_t span = p.Tp * (int64_t)p.S;
  const int64_t memlen = p.N - p.S;
  const int c0 = (p.n_pad + 1) * p.S - p.N;
  // this warp's job in the filterbank product: one m-tile (16 filters) of its group's 8 frames
  const int group = warp >> 3, wg = warp & 7, mt = wg;
  int buf = 0;

  for (int64_t item = (int64_t)blockIdx.x * kGroups + group; item < a.n_items; item += (int64_t)gridDim.x * kGroups, buf ^= 1) {
    const int64_t row = item / a.rounds_per_row;
    const int64_t t_round = (item - row * a.rounds_per_row) * kFpr;
    const float* rowp = p.pcm + row * p.row_stride;
    float* tile = tile0 + buf * kTileFloats;
    {
      const int64_t t = t_round + wg;
      const bool computed = t < p.T;  // frames in [T, T_out) are the reference's zero tail

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
            const float xa = o >= 0 ? (o < p.row_len ? rowp[o] : 0.f) : stream_at(p, q0 + 64 * j);
            const float xb = o + 1 >= 0 ? (o + 1 < p.row_len ? rowp[o + 1] : 0.f) : stream_at(p, q0 + 64 * j + 1);
            v[j] = make_float2(xa * w.x, xb * w.y);
          });
        }
      }
      // ---- stage 1: 32-point FFT over j
      ssfft::fft_pow2<kR>(v);
      // ---- twiddle W1024^{u k1} and transpose
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
      float2 gg[kR];
      {
        const float4* r4 = reinterpret_cast<const float4*>(xch + u * kXs);
        sfor<0, kR / 2>([&](auto tt) {
          constexpr int i = decltype(tt)::value;
          const float4 x = r4[i];
          gg[2 * i] = make_float2(x.x, x.y);
          gg[2 * i + 1] = make_float2(x.z, x.w);
        });
      }
      __syncwarp();  // the transpose tile is rewritten in the next round
      // ---- stage 2: 32-point FFT over the lane index.  gg[k2] = Z[u + 32 k2]
      ssfft::fft_pow2<kR>(gg);

      // ---- real-FFT post-processing for k = u + 32 k2 <= 511 (k2 < 16); Z[1024 - k] is
      // gg[31 - k2] of lane 32 - u (lane 0: its own gg[32 - k2]); |X|^2 x 4 goes to the tile row
      float* Pf = tile + wg * kPs;
      float* pA = Pf + u;
      sfor<0, 16>([&](auto kk) {
        constexpr int k2 = decltype(kk)::value;
        float2 zb;
        zb.x = __shfl_sync(0xffffffffu, gg[31 - k2].x, partner);
        zb.y = __shfl_sync(0xffffffffu, gg[31 - k2].y, partner);
        if (isz) zb = gg[(32 - k2) & 31];
        const float2 za = gg[k2];
        const float2 w = twk[32 * k2];
        const float2 e = __fadd2_rn(za, make_float2(zb.x, -zb.y));
        const float2 od = __fadd2_rn(make_float2(za.y, -za.x), make_float2(zb.y, zb.x));
        const float2 xa = ssfft::cadd(e, ssfft::cmulf(od, w));
        pA[32 * k2] = fmaf(xa.x, xa.x, xa.y * xa.y);
      });
      if (isz)  // X[512] = conj Z[512]; bins 513..519 stay zero
        *reinterpret_cast<float4*>(Pf + 512) = make_float4(4.0f * fmaf(gg[16].x, gg[16].x, gg[16].y * gg[16].y), 0.f, 0.f, 0.f);
    }
    // the group's tile of this round is complete (the other buffer belongs to the next round)
    asm volatile("bar.sync %0, 256;" ::"r"(1 + group) : "memory");

    // ---- mel filterbank on the tensor cores (feature.rs:163-174), transposed: this warp's job is
    //   mel^T [16 filters of m-tile mt][8 frames] = W [16][band] x P^T,
    // contraction slot t <-> bin 8 ks + 2t, slot t + 4 <-> bin 8 ks + 2t + 1 (one 8-byte load).
    if (mt < a.n_mt) {
      float accB[4] = {0.f, 0.f, 0.f, 0.f}, accS[4] = {0.f, 0.f, 0.f, 0.f}, accT[4] = {0.f, 0.f, 0.f, 0.f};
      const float* pp = tile + g * kPs + 2 * t4 + 8 * a.ks_lo[mt];
      const float4* ap = a.mel_frag + (int64_t)a.mel_base[mt] * 64 + lane;
      int n2 = a.ks_n2[mt];
      // two k-steps per iteration (the band is padded to an even count with zero-weight blocks), the
      // weight fragments of the next iteration are requested before the MMAs of this one
      float4 h0, l0, h1, l1;
      if (n2 > 0) { h0 = __ldg(ap); l0 = __ldg(ap + 32); h1 = __ldg(ap + 64); l1 = __ldg(ap + 96); }
#pragma unroll 1
      for (; n2 > 0; --n2) {
        const float2 pv0 = *reinterpret_cast<const float2*>(pp);
        const float2 pv1 = *reinterpret_cast<const float2*>(pp + 8);
        const float4 ch0 = h0, cl0 = l0, ch1 = h1, cl1 = l1;
        pp += 16;
        ap += 128;
        if (n2 > 1) { h0 = __ldg(ap); l0 = __ldg(ap + 32); h1 = __ldg(ap + 64); l1 = __ldg(ap + 96); }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 pv = h ? pv1 : pv0;
          const float4 wh = h ? ch1 : ch0, wl = h ? cl1 : cl0;
          uint32_t bh0, bl0, bh1, bl1;
          split_tf32(pv.x, bh0, bl0);
          split_tf32(pv.y, bh1, bl1);
          mma_tf32(accB, wh, bh0, bh1);  // hi x hi
          mma_tf32(accS, wl, bh0, bh1);  // lo x hi
          mma_tf32(accT, wh, bl0, bl1);  // hi x lo
        }
      }
      // accumulator: (filter 16 mt + g [+ 8], frames t_round + 2 t4 [+ 1])
      const int64_t fa = t_round + 2 * t4;
      const bool va = fa < p.T_out, vb = fa + 1 < p.T_out;
      float* outp = p.mel_out + row * nf * p.T_out + fa;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int f = 16 * mt + g + 8 * hh;
        float m0 = accB[2 * hh] + (accS[2 * hh] + accT[2 * hh]);
        float m1 = accB[2 * hh + 1] + (accS[2 * hh + 1] + accT[2 * hh + 1]);
        if (p.log_mode) {
          m0 = fast_ln(zero_handling(m0));
          m1 = fast_ln(zero_handling(m1));
        }
        if (f < nf) {
          float* dst = outp + (int64_t)f * p.T_out;
          if (a.pair_ok && vb) {
            *reinterpret_cast<float2*>(dst) = make_float2(m0, m1);
          } else {
            if (va) dst[0] = m0;
            if (vb) dst[1] = m1;
          }
        }
      }
    }
  }
}

// round-to-nearest (ties away) conversion f32 -> TF32, as cvt.rna.tf32.f32
static float tf32_rna(float x) {
  uint32_t b;
  std::memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xffffe000u;  // the weights are finite and far from overflow
  float r;
  std::memcpy(&r, &b, 4);
  return r;
}

struct FragTables2048 {
  int n_mt = 0;
  int ks_lo[kMtMax] = {0}, ks_n2[kMtMax] = {0}, mel_base[kMtMax] = {0};
  std::vector<float> mel;  // [blocks][hi, lo][32][4]
};

// A fragments of mma.m16n8k8 (row-major 16 filters x 8 bins) for lane (g, t): rows g, g + 8;
// contraction slot s <-> bin 8 ks + 2 (s % 4) + s / 4.  Only the k-steps inside each 16-filter
// group's band are stored, padded to an even count.
static bool build_frags(const Derived& d, const MelBank& mb, float pscale, FragTables2048* ft) {
  const int nf = d.nf;
  if (mb.max_bin > 512 || nf > 16 * kMtMax || nf < 1) return false;
  const int MT = (nf + 15) / 16, ncol = 8 * kKsTab;
  ft->n_mt = MT;
  std::vector<float> dense((size_t)16 * MT * ncol, 0.f);  // [filter][bin]
  for (int f = 0; f < nf; ++f)
    for (int i = 0; i < mb.len[f]; ++i) dense[(size_t)f * ncol + mb.start[f] + i] = mb.w[mb.wofs[f] + i] * pscale;
  int n_blocks = 0;
  ft->mel.clear();
  for (int mt = 0; mt < MT; ++mt) {
    int lo = kKs, hi = 0;
    for (int ks = 0; ks < kKs; ++ks) {
      bool any = false;
      for (int r = 0; r < 16; ++r)
        for (int c = 0; c < 8; ++c) any = any || dense[(size_t)(16 * mt + r) * ncol + 8 * ks + c] != 0.f;
      if (any) { lo = std::min(lo, ks); hi = std::max(hi, ks + 1); }
    }
    ft->ks_lo[mt] = 0;
    ft->ks_n2[mt] = 0;
    ft->mel_base[mt] = n_blocks;
    if (hi <= lo) continue;  // 16 empty filters: no k-steps, the outputs are exact zeros
    const int n2 = (hi - lo + 1) / 2;
    ft->ks_lo[mt] = lo;
    ft->ks_n2[mt] = n2;
    for (int ks = lo; ks < lo + 2 * n2; ++ks, ++n_blocks) {
      const size_t o = ft->mel.size();
      ft->mel.resize(o + 2 * 32 * 4, 0.f);
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        for (int i = 0; i < 4; ++i) {
          const int r = g + 8 * (i & 1), s = t + 4 * (i >> 1);
          const float v = dense[(size_t)(16 * mt + r) * ncol + 8 * ks + 2 * (s & 3) + (s >> 2)];
          const float h = tf32_rna(v);
          ft->mel[o + lane * 4 + i] = h;
          ft->mel[o + (32 + lane) * 4 + i] = v - h;
        }
      }
    }
  }
  return true;
}

}  // namespace

struct Stft2048mTables {
  float* d_tab = nullptr;
  size_t o_mel = 0;
  int n_mt = 0;
  int ks_lo[kMtMax] = {0}, ks_n2[kMtMax] = {0}, mel_base[kMtMax] = {0};
};

bool stft2048m_supported(const Derived& d, const MelBank& mb) {
  if (d.librosa || d.W != kW || !d.stft_ok) return false;
  FragTables2048 ft;
  return build_frags(d, mb, 1.0f, &ft);
}

int stft2048m_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft2048mTables** out) {
  FragTables2048 ft;
  const float pscale = 0.25f * d.wnorm * d.wnorm;  // (1/2)^2 of E, O and wnorm^2 (functions.rs:166-169)
  if (!build_frags(d, mb, pscale, &ft)) return (int)cudaErrorInvalidValue;
  std::vector<float> h(kTabFixed, 0.f);
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
  Stft2048mTables* t = new Stft2048mTables;
  t->o_mel = h.size();
  h.insert(h.end(), ft.mel.begin(), ft.mel.end());
  t->n_mt = ft.n_mt;
  for (int i = 0; i < kMtMax; ++i) { t->ks_lo[i] = ft.ks_lo[i]; t->ks_n2[i] = ft.ks_n2[i]; t->mel_base[i] = ft.mel_base[i]; }
  cudaError_t e = cudaMalloc(&t->d_tab, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return (int)e; }
  *out = t;
  return 0;
}

void stft2048m_free_tables(Stft2048mTables* t) {
  if (!t) return;
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

// mel_spectrogram only (p.mel_out set, p.stft_out null); returns -1000 if the shape is declined
int launch_stft2048m(const FrameParams& p, const Stft2048mTables* t, int sm_count, void* stream) {
  if (p.N != kW || !p.mel_out || p.stft_out || p.T_out < 1) return -1000;
  Stft2048mArgs a;
  a.p = p;
  a.tab = t->d_tab;
  a.mel_frag = reinterpret_cast<const float4*>(t->d_tab + t->o_mel);
  a.n_mt = t->n_mt;
  for (int i = 0; i < kMtMax; ++i) { a.ks_lo[i] = t->ks_lo[i]; a.ks_n2[i] = t->ks_n2[i]; a.mel_base[i] = t->mel_base[i]; }
  a.rounds_per_row = (int)((p.T_out + kFpr - 1) / kFpr);
  a.n_items = p.n_rows * (int64_t)a.rounds_per_row;
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  a.pair_ok = (p.T_out % 2 == 0) && (reinterpret_cast<uintptr_t>(p.mel_out) % 8 == 0);
  const size_t smem = (size_t)(kTabFixed + kWarps * kWarpFloats + kGroups * 2 * kTileFloats) * sizeof(float);
  const int64_t grid = std::min<int64_t>((a.n_items + kGroups - 1) / kGroups, (int64_t)sm_count);
  if (grid <= 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(k_stft2048m, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_stft2048m<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);
  count_launch();
  note_kernel("k_stft2048m");
  return (int)cudaGetLastError();
}

}  // namespace ss
