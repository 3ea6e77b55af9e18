// Fused speechpy-framing feature kernel for n_fft = 512 (configs c1 / c2 / c5):
//   framing -> zero-pad -> real FFT -> |X|/N -> frame energy -> banded mel -> ln -> DCT-II
//   (+ SSC, + per-clip CMVN), one pass, every intermediate in registers / shared memory.
// Replaces feature::{mfe, lmfe, mfcc} (feature.rs:99-245) + processing::{stack_frames,
// power_spectrum, cmvn} (processing.rs:65-181,265-300) for this FFT size.
//
// Mapping (see DESIGN.md "fast512"):
//   * two consecutive real frames A, B are packed as z = A + iB into ONE complex 512-point FFT,
//     computed by 16 threads (a half-warp) holding 32 complex points each;
//   * 512 = 32 x 16 (four-step): stage 1 is a 32-point DIT FFT entirely in registers with
//     compile-time twiddles (6-FFMA butterflies), then the W512^{u*k1} twiddle from a shared
//     table, ONE transpose through shared memory (padded, conflict-free, warp-local: only
//     __syncwarp), stage 2 is two 16-point register FFTs per thread;
//   * thread u owns output groups k1 = u and k1 = 32-u (u = 0: 0 and 16), so Z[k] and Z[512-k]
//     sit in the same thread and the A/B spectra separate without another exchange;
//   * a warp = 2 pairs = 4 consecutive frames; warps never synchronise with each other until the
//     optional per-clip CMVN at the end of a clip.
#include <cuda_runtime.h>

#include <type_traits>
#include <vector>

#include "ss_internal.h"

namespace ss {

namespace {

constexpr int kWarps = 5;
constexpr int kThreads = kWarps * 32;
constexpr int kXs = 17;                      // exchange row stride in float2 (16 + 1 pad)
constexpr int kXchPair = 32 * kXs;           // float2 per pair
constexpr int kWarpFloats = 2 * kXchPair * 2;  // floats per warp (2 pairs) = 2176
constexpr int kPs = 264;                     // P row stride (floats); 264 % 32 == 8
constexpr int kNfMax = 64, kNcMax = 16, kNwMax = 640, kSlots = 4;
constexpr int kRow = 68;                     // row stride of log_s and the DCT rows (68/4 odd)
// per-warp float offsets after the exchange buffer has been consumed
constexpr int kOffP = 0;
constexpr int kOffMel = 4 * kPs;             // 1056
constexpr int kOffSsc = kOffMel + 4 * kNfMax;  // 1312
constexpr int kOffLog = kOffSsc + 4 * kNfMax;  // 1568
static_assert(kOffLog + 4 * kRow <= kWarpFloats, "per-warp scratch overflows the exchange buffer");

// CTA-shared table blob (float units)
constexpr int kTabTw = 0;                       // [32][16] float2
constexpr int kTabFw = kTabTw + 32 * 16 * 2;    // [kNwMax]
constexpr int kTabFw2 = kTabFw + kNwMax;        // [kNwMax]  w * R[bin] (SSC)
constexpr int kTabDct = kTabFw2 + kNwMax;       // [kNcMax][kRow]
constexpr int kTabSched = kTabDct + kNcMax * kRow;  // int [16][kSlots][4] = filt, b0, len, wofs
constexpr int kTabNslot = kTabSched + 16 * kSlots * 4;  // int [16]
constexpr int kTabFloats = kTabNslot + 16;

template <int I, int E, class F>
__device__ __forceinline__ void sfor(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    sfor<I + 1, E>(static_cast<F&&>(f));
  }
}

// cos(2 pi m / 32), m = 0..8
constexpr float kC32[9] = {1.0f,          0.98078528040323043f, 0.92387953251128674f,
                           0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                           0.38268343236508978f, 0.19509032201612825f, 0.0f};
constexpr float cos32(int m) {
  m = ((m % 32) + 32) % 32;
  return m <= 8 ? kC32[m] : m <= 16 ? -kC32[16 - m] : m <= 24 ? -kC32[m - 16] : kC32[32 - m];
}
constexpr float sin32(int m) { return cos32(m - 8); }
template <int BITS>
constexpr int brev(int x) {
  int r = 0;
  for (int i = 0; i < BITS; ++i) r |= ((x >> i) & 1) << (BITS - 1 - i);
  return r;
}

// DIT butterfly with w = e^{-2 pi i M/32}:  a' = a + w b,  b' = a - w b = 2a - a'
template <int M>
__device__ __forceinline__ void bfly(float2& a, float2& b) {
  if constexpr (M == 0) {
    const float2 t = a;
    a = make_float2(t.x + b.x, t.y + b.y);
    b = make_float2(t.x - b.x, t.y - b.y);
  } else if constexpr (M == 8) {  // w = -i : w b = (b.y, -b.x)
    const float2 t = a, s = b;
    a = make_float2(t.x + s.y, t.y - s.x);
    b = make_float2(t.x - s.y, t.y + s.x);
  } else {
    constexpr float c = cos32(M), s = sin32(M);  // w = c - i s
    float2 t;
    t.x = fmaf(c, b.x, a.x);
    t.x = fmaf(s, b.y, t.x);
    t.y = fmaf(c, b.y, a.y);
    t.y = fmaf(-s, b.x, t.y);
    b = make_float2(fmaf(2.0f, a.x, -t.x), fmaf(2.0f, a.y, -t.y));
    a = t;
  }
}

// one DIT level of an R-point FFT held in registers (bit-reversed order in t)
template <int R, int BITS, int LEN, int JNZ>
__device__ __forceinline__ void dit_level(float2 (&t)[R]) {
  constexpr int half = LEN / 2;
  sfor<0, R / 2>([&](auto bi) {
    constexpr int b = decltype(bi)::value;
    constexpr int blk = b / half, k = b % half;
    constexpr int ia = blk * LEN + k, ib = ia + half;
    constexpr int m = k * (32 / LEN);
    if constexpr (LEN == 2 && brev<BITS>(ib) >= JNZ) {
      t[ib] = t[ia];  // the partner input is a structural zero (zero padding L -> 512)
    } else {
      bfly<m>(t[ia], t[ib]);
    }
  });
}

// y[k] = sum_j x[j] e^{-2 pi i jk/R}; x[j] == 0 for j >= JNZ is known at compile time
template <int R, int BITS, int JNZ>
__device__ __forceinline__ void fft_regs(float2 (&x)[R]) {
  float2 t[R];
  sfor<0, R>([&](auto i) { t[decltype(i)::value] = x[brev<BITS>(decltype(i)::value)]; });
  dit_level<R, BITS, 2, JNZ>(t);
  dit_level<R, BITS, 4, R>(t);
  dit_level<R, BITS, 8, R>(t);
  dit_level<R, BITS, 16, R>(t);
  if constexpr (R == 32) dit_level<R, BITS, 32, R>(t);
  sfor<0, R>([&](auto i) { x[decltype(i)::value] = t[decltype(i)::value]; });
}

__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct FastArgs {
  FrameParams p;
  const float* tab;   // table blob in global memory
  int64_t n_items;    // clips * chunks_per_clip
  int quads_per_clip, chunk_quads, chunks_per_clip;
  int fuse_cmvn;      // 0 none, 1 mean, 2 mean+var: whole clip in one CTA, stash in smem
  int stash_floats;
};

template <int JMAX>
__global__ void __launch_bounds__(kThreads, 3) k_fast512(const FastArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* tab_s = smem;
  float* warp_s = smem + kTabFloats + (threadIdx.x >> 5) * kWarpFloats;
  float* stash = smem + kTabFloats + kWarps * kWarpFloats;  // [T][nc] when fuse_cmvn
  const FrameParams& p = a.p;

  for (int i = threadIdx.x; i < kTabFloats; i += kThreads) tab_s[i] = __ldg(a.tab + i);
  __syncthreads();

  const float2* tw_s = reinterpret_cast<const float2*>(tab_s + kTabTw);
  const float* fw_s = tab_s + kTabFw;
  const float* fw2_s = tab_s + kTabFw2;
  const float* dct_s = tab_s + kTabDct;
  const int* sched_s = reinterpret_cast<const int*>(tab_s + kTabSched);
  const int* nslot_s = reinterpret_cast<const int*>(tab_s + kTabNslot);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = lane >> 4, u = lane & 15;
  const int nf = p.nf, nc = p.nc, L = p.L;
  const int64_t T = p.T;
  const bool isz = (u == 0);
  const int k1b = isz ? 16 : 32 - u;
  const float mscale = 0.5f * p.mag_scale;  // the 1/2 of the A/B separation folded into 1/n_fft
  const bool want_mel = p.mfe || p.lmfe || p.ssc || p.mfcc;

  float2* xch = reinterpret_cast<float2*>(warp_s) + h * kXchPair;  // this pair's transpose buffer
  float* P_s = warp_s + kOffP;
  float* mel_s = warp_s + kOffMel;
  float* ssc_s = warp_s + kOffSsc;
  float* log_s = warp_s + kOffLog;

  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int64_t clip = item / a.chunks_per_clip;
    const int chunk = (int)(item - clip * a.chunks_per_clip);
    const int q_end = min(a.quads_per_clip, (chunk + 1) * a.chunk_quads);
    const float* rowp = p.pcm + clip * p.row_stride;

    for (int q = chunk * a.chunk_quads + warp; q < q_end; q += kWarps) {
      const int64_t fq = 4 * (int64_t)q;  // first frame of the quad
      const int64_t fA = fq + 2 * h, fB = fA + 1;
      const bool vA = fA < T, vB = fB < T;

      // ---- framing: v[j] = (A[u + 16 j], B[u + 16 j]); zero beyond L (processing.rs:147-156)
      float2 v[32];
      {
        const float* pa = rowp + fA * p.S + u;
        const float* pb = pa + p.S;
        sfor<0, 32>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          if constexpr (j < JMAX) {
            const bool in = (u + 16 * j) < L;
            v[j].x = (vA && in) ? __ldg(pa + 16 * j) : 0.0f;
            v[j].y = (vB && in) ? __ldg(pb + 16 * j) : 0.0f;
          } else {
            v[j] = make_float2(0.0f, 0.0f);
          }
        });
      }
      // ---- stage 1: 32-point FFT over j, in registers
      fft_regs<32, 5, JMAX>(v);
      // ---- twiddle W512^{u k1} and transpose through shared memory
      __syncwarp();  // previous quad's readers of this buffer are done
      sfor<0, 32>([&](auto kk) {
        constexpr int k1 = decltype(kk)::value;
        float2 y = v[k1];
        if constexpr (k1 > 0) {
          const float2 w = tw_s[k1 * 16 + u];
          y = make_float2(y.x * w.x - y.y * w.y, y.x * w.y + y.y * w.x);
        }
        xch[k1 * kXs + u] = y;
      });
      __syncwarp();
      float2 g1[16], g2[16];
      sfor<0, 16>([&](auto tt) {
        constexpr int t = decltype(tt)::value;
        g1[t] = xch[u * kXs + t];
        g2[t] = xch[k1b * kXs + t];
      });
      __syncwarp();  // the buffer is reused for P / mel / log below
      // ---- stage 2: 16-point FFTs over t
      fft_regs<16, 4, 16>(g1);
      fft_regs<16, 4, 16>(g2);

      // ---- separate A / B spectra, magnitude (processing.rs:168,180), frame energy
      float eA = 0.f, eB = 0.f;
      float* PA = P_s + (2 * h) * kPs;
      float* PB = PA + kPs;
      sfor<0, 8>([&](auto mm) {
        constexpr int m = decltype(mm)::value;
        {  // bin u + 32 m
          const float2 Z = g1[m];
          const float2 Zm = isz ? g1[(16 - m) & 15] : g2[15 - m];
          const float ax = Z.x + Zm.x, ay = Z.y - Zm.y, bx = Z.y + Zm.y, by = Zm.x - Z.x;
          const float ma = mscale * fast_sqrt(fmaf(ax, ax, ay * ay));
          const float mb = mscale * fast_sqrt(fmaf(bx, bx, by * by));
          PA[u + 32 * m] = ma;
          PB[u + 32 * m] = mb;
          eA += ma;
          eB += mb;
        }
        {  // bin k1b + 32 m
          const float2 Z = g2[m];
          const float2 Zm = isz ? g2[15 - m] : g1[15 - m];
          const float ax = Z.x + Zm.x, ay = Z.y - Zm.y, bx = Z.y + Zm.y, by = Zm.x - Z.x;
          const float ma = mscale * fast_sqrt(fmaf(ax, ax, ay * ay));
          const float mb = mscale * fast_sqrt(fmaf(bx, bx, by * by));
          PA[k1b + 32 * m] = ma;
          PB[k1b + 32 * m] = mb;
          eA += ma;
          eB += mb;
        }
      });
      if (isz) {  // bin 256 (self-mirrored): X_A = Re Z, X_B = Im Z
        const float ma = 2.0f * mscale * fabsf(g1[8].x), mb = 2.0f * mscale * fabsf(g1[8].y);
        PA[256] = ma;
        PB[256] = mb;
        eA += ma;
        eB += mb;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        eA += __shfl_xor_sync(0xffffffffu, eA, o);
        eB += __shfl_xor_sync(0xffffffffu, eB, o);
      }
      eA = zero_handling(eA);  // feature.rs:219
      eB = zero_handling(eB);
      __syncwarp();

      // ---- banded mel reduction (feature.rs:229-230): lane u owns a balanced set of filters
      if (want_mel) {
        const int ns = nslot_s[u];
        for (int s = 0; s < ns; ++s) {
          const int4 e = *reinterpret_cast<const int4*>(sched_s + (u * kSlots + s) * 4);
          const int filt = e.x, b0 = e.y, len = e.z, wo = e.w;
          float accA = 0.f, accB = 0.f, numA = 0.f, numB = 0.f;
          for (int i = 0; i < len; ++i) {
            const float w = fw_s[wo + i], w2 = fw2_s[wo + i];
            const float pa = PA[b0 + i], pb = PB[b0 + i];
            accA = fmaf(w, pa, accA);
            accB = fmaf(w, pb, accB);
            numA = fmaf(w2, pa, numA);
            numB = fmaf(w2, pb, numB);
          }
          accA = zero_handling(accA);
          accB = zero_handling(accB);
          const int oA = (2 * h) * nf + filt, oB = oA + nf;
          mel_s[oA] = accA;
          mel_s[oB] = accB;
          ssc_s[oA] = __fdividef(numA, accA);
          ssc_s[oB] = __fdividef(numB, accB);
          log_s[(2 * h) * kRow + filt] = __logf(accA);
          log_s[(2 * h + 1) * kRow + filt] = __logf(accB);
        }
        // zero the float4 padding of the log rows (the buffer held FFT data a moment ago)
        if (u < ((nf + 3) & ~3) - nf) {
          log_s[(2 * h) * kRow + nf + u] = 0.0f;
          log_s[(2 * h + 1) * kRow + nf + u] = 0.0f;
        }
        __syncwarp();
      }

      // ---- stores: [clip][t][nf] blocks, 4 consecutive frames = 4*nf contiguous floats
      const int64_t fbase = clip * T + fq;
      const int nvalid = (int)min((int64_t)4, T - fq);
      if (p.mfe || p.lmfe || p.ssc) {
        const int tot = nvalid * nf;
        for (int i = lane; i < tot; i += 32) {
          if (p.mfe) p.mfe[fbase * nf + i] = mel_s[i];
          if (p.ssc) p.ssc[fbase * nf + i] = ssc_s[i];
          if (p.lmfe) {
            const int fr = i / nf;
            p.lmfe[fbase * nf + i] = log_s[fr * kRow + (i - fr * nf)];
          }
        }
      }
      if (p.energy && isz) {
        if (vA) p.energy[clip * T + fA] = eA;
        if (vB) p.energy[clip * T + fB] = eB;
      }
      // ---- ln -> DCT-II -> scale -> c0 (feature.rs:105-146); lane = (pair h, cepstrum u)
      if (p.mfcc && u < nc) {
        const float4* la = reinterpret_cast<const float4*>(log_s + (2 * h) * kRow);
        const float4* lb = reinterpret_cast<const float4*>(log_s + (2 * h + 1) * kRow);
        const float4* dr = reinterpret_cast<const float4*>(dct_s + u * kRow);
        float accA = 0.f, accB = 0.f;
        const int n4 = (nf + 3) >> 2;
        for (int i = 0; i < n4; ++i) {
          const float4 d = dr[i], x = la[i], y = lb[i];
          accA = fmaf(d.x, x.x, accA); accA = fmaf(d.y, x.y, accA);
          accA = fmaf(d.z, x.z, accA); accA = fmaf(d.w, x.w, accA);
          accB = fmaf(d.x, y.x, accB); accB = fmaf(d.y, y.y, accB);
          accB = fmaf(d.z, y.z, accB); accB = fmaf(d.w, y.w, accB);
        }
        if (u == 0) {
          if (p.dc_elim) {
            accA = __logf(eA);
            accB = __logf(eB);
          } else if (fA == 0) {
            accA *= p.scale_00;  // only element [0,0] gets the k=0 scale (feature.rs:128)
          }
        } else {
          accA *= p.scale_rest;
          accB *= p.scale_rest;
        }
        if (a.fuse_cmvn) {
          if (vA) stash[fA * nc + u] = accA;
          if (vB) stash[fB * nc + u] = accB;
        } else {
          if (vA) p.mfcc[(clip * T + fA) * nc + u] = accA;
          if (vB) p.mfcc[(clip * T + fB) * nc + u] = accB;
        }
      }
    }

    // ---- per-clip CMVN (processing.rs:265-300) on the stashed [T][nc] block ---------------
    if (a.fuse_cmvn) {
      __syncthreads();
      float* stat = stash + a.stash_floats;  // [nc][2]
      const int Ti = (int)T;
      for (int c = warp; c < nc; c += kWarps) {
        float s = 0.f;
        for (int t = lane; t < Ti; t += 32) s += stash[t * nc + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)Ti;
        float inv = 1.0f;
        if (a.fuse_cmvn == 2) {
          // std of the mean-subtracted column, ddof 0 (two-pass: mean of y, then squares)
          float s1 = 0.f;
          for (int t = lane; t < Ti; t += 32) s1 += stash[t * nc + c] - mean;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          const float my = s1 / (float)Ti;
          float s2 = 0.f;
          for (int t = lane; t < Ti; t += 32) {
            const float d = (stash[t * nc + c] - mean) - my;
            s2 = fmaf(d, d, s2);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          inv = 1.0f / (sqrtf(s2 / (float)Ti) + kCmvnEps);
        }
        if (lane == 0) {
          stat[2 * c] = mean;
          stat[2 * c + 1] = inv;
        }
      }
      __syncthreads();
      float* dst = p.mfcc + clip * T * nc;
      const int tot = Ti * nc;
      for (int i = threadIdx.x; i < tot; i += kThreads) {
        const int c = i % nc;
        dst[i] = (stash[i] - stat[2 * c]) * stat[2 * c + 1];
      }
      __syncthreads();  // the stash is rewritten by the next clip
    }
  }
}

}  // namespace

struct Fast512Tables {
  float* d_tab = nullptr;
};

bool fast512_supported(const Derived& d, const MelBank& mb) {
  if (d.W != 512 || !d.speechpy_ok) return false;
  if (d.nf > kNfMax || d.nc > kNcMax) return false;
  if ((int)mb.w.size() > kNwMax) return false;
  return true;
}

int fast512_build_tables(const Derived& d, const MelBank& mb, Fast512Tables** out) {
  std::vector<float> h(kTabFloats, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  for (int k1 = 0; k1 < 32; ++k1)
    for (int u = 0; u < 16; ++u) {
      const double ang = -2.0 * pi * (double)(u * k1) / 512.0;
      h[kTabTw + 2 * (k1 * 16 + u)] = (float)std::cos(ang);
      h[kTabTw + 2 * (k1 * 16 + u) + 1] = (float)std::sin(ang);
    }
  const float r0 = 1.0f, step = d.F > 1 ? ((float)d.cfg.sample_rate / 2.0f - 1.0f) / (float)(d.F - 1) : 0.f;
  for (int i = 0; i < d.nf; ++i)
    for (int k = 0; k < mb.len[i]; ++k) {
      const float w = mb.w[mb.wofs[i] + k];
      h[kTabFw + mb.wofs[i] + k] = w;
      h[kTabFw2 + mb.wofs[i] + k] = w * (r0 + step * (float)(mb.start[i] + k));
    }
  std::vector<float> dct;
  build_dct(d, &dct);
  for (int c = 0; c < d.nc; ++c)
    for (int n = 0; n < d.nf; ++n) h[kTabDct + c * kRow + n] = dct[(size_t)c * d.nf + n];
  std::vector<std::vector<int>> per_lane;
  balance_filters(mb, 16, &per_lane);
  bool ok = true;
  for (auto& v : per_lane) ok = ok && (int)v.size() <= kSlots;
  if (!ok) {
    per_lane.assign(16, {});
    for (int i = 0; i < d.nf; ++i) per_lane[i % 16].push_back(i);
  }
  int* sched = reinterpret_cast<int*>(h.data() + kTabSched);
  int* nslot = reinterpret_cast<int*>(h.data() + kTabNslot);
  for (int l = 0; l < 16; ++l) {
    nslot[l] = (int)per_lane[l].size();
    for (int s = 0; s < (int)per_lane[l].size(); ++s) {
      const int f = per_lane[l][s];
      int* e = sched + (l * kSlots + s) * 4;
      e[0] = f;
      e[1] = mb.start[f];
      e[2] = mb.len[f];
      e[3] = mb.wofs[f];
    }
  }
  Fast512Tables* t = new Fast512Tables;
  cudaError_t e = cudaMalloc(&t->d_tab, kTabFloats * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), kTabFloats * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return (int)e; }
  *out = t;
  return 0;
}

void fast512_free_tables(Fast512Tables* t) {
  if (!t) return;
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

// returns 0 (done, including any requested cmvn), -1000 (shape declined), -1001 (features
// written, cmvn left to the caller) or a cudaError_t
int launch_fast512(const FrameParams& p, const Fast512Tables* t, int sm_count, void* stream) {
  if (p.N != 512 || p.L > 512 || p.T < 1 || p.T > (1LL << 31) - 8) return -1000;
  FastArgs a;
  a.p = p;
  a.tab = t->d_tab;
  const int64_t qpc = (p.T + 3) / 4;
  a.quads_per_clip = (int)qpc;
  // fused per-clip CMVN when the clip's [T][nc] block fits a 24 KB stash
  const size_t stash_bytes = (size_t)p.T * p.nc * sizeof(float);
  const bool fuse = p.mfcc && p.cmvn_mode && stash_bytes <= 24 * 1024;
  a.fuse_cmvn = fuse ? p.cmvn_mode : 0;
  a.stash_floats = fuse ? (int)(p.T * p.nc) : 0;
  if (fuse) {
    a.chunk_quads = (int)qpc;
  } else {
    // enough items for ~8 waves of 3 CTAs per SM, in multiples of one round of warps
    const int64_t total_quads = qpc * p.n_rows;
    int64_t cq = total_quads / ((int64_t)sm_count * 3 * 8);
    cq = std::max<int64_t>(kWarps, std::min<int64_t>(cq, 8 * kWarps));
    cq = (cq + kWarps - 1) / kWarps * kWarps;
    a.chunk_quads = (int)std::min<int64_t>(cq, std::max<int64_t>(qpc, 1));
  }
  a.chunks_per_clip = (int)((qpc + a.chunk_quads - 1) / a.chunk_quads);
  a.n_items = p.n_rows * a.chunks_per_clip;
  const size_t smem = (size_t)(kTabFloats + kWarps * kWarpFloats + a.stash_floats + (fuse ? 2 * kNcMax : 0)) * sizeof(float);
  const int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sm_count * 3);
  if (grid <= 0) return 0;
  const int jn = (p.L + 15) / 16;
  cudaError_t e;
#define SS_LAUNCH_FAST(J)                                                                                  \
  do {                                                                                                     \
    e = cudaFuncSetAttribute(k_fast512<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    if (e != cudaSuccess) return (int)e;                                                                   \
    k_fast512<J><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);                             \
  } while (0)
  if (jn <= 20) SS_LAUNCH_FAST(20);
  else if (jn <= 25) SS_LAUNCH_FAST(25);
  else SS_LAUNCH_FAST(32);
#undef SS_LAUNCH_FAST
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return (p.cmvn_mode && !fuse) ? -1001 : 0;
}

}  // namespace ss
