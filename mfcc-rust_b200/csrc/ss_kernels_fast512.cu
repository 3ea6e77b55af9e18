// Fused speechpy-framing feature kernel for n_fft = 512 (configs c1 / c2 / c5):
//   framing -> zero-pad -> real FFT -> |X|/N -> frame energy -> banded mel -> ln -> DCT-II
//   (+ SSC, + per-clip CMVN), one pass, every intermediate in registers / shared memory.
// Replaces feature::{mfe, lmfe, mfcc} (feature.rs:99-245) + processing::{stack_frames,
// power_spectrum, cmvn} (processing.rs:65-181,265-300) for this FFT size.
//
// Mapping (see DESIGN.md "fast512"):
//   * one real frame = one 256-point complex FFT of z[m] = x[2m] + i x[2m+1] plus the real-FFT
//     post-processing step, computed by 8 threads (a quarter-warp) holding 32 complex points
//     each.  Frames are NOT packed two-per-complex-FFT: that trick leaks ~1e-7 of a loud frame
//     into its partner, which breaks the reference's exact zero_handling on silent frames;
//   * 256 = 32 x 8 (four-step): stage 1 is a 32-point DIT FFT entirely in registers with
//     compile-time twiddles (6-FFMA butterflies), then the W256^{u*k1} twiddle from a shared
//     table, ONE transpose through shared memory (padded, conflict-free, warp-local: only
//     __syncwarp), stage 2 is four 8-point register FFTs per thread;
//   * thread u owns output groups k1 = {u, 32-u, 16-u, 16+u} (u = 0: {0, 16, 8, 24}), so Z[k] and
//     Z[256-k] sit in the same thread and the post-processing needs no second exchange;
//   * a warp = 4 consecutive frames; warps never synchronise with each other until the optional
//     per-clip CMVN at the end of a clip.
#include <cuda_runtime.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "ss_internal.h"

namespace ss {

namespace {

constexpr int kWarps = 5;
constexpr int kThreads = kWarps * 32;
constexpr int kXs = 10;                      // exchange row stride in float2 (8 + 2 pad: 16-byte rows, conflict-free LDS.128)
constexpr int kXchFrame = 32 * kXs + 8;      // float2 per frame; +8 shifts the next frame by half the banks
constexpr int kWarpFloats = 4 * kXchFrame * 2;  // floats per warp (4 frames) = 2624
constexpr int kPs = 264;                     // P row stride (floats); 264 % 32 == 8
constexpr int kNfMax = 64, kNcMax = 16;
constexpr int kRow = 68;                     // row stride of log_s and the DCT rows (68/4 odd)
// per-warp float offsets after the exchange buffer has been consumed
constexpr int kOffP = 0;
constexpr int kOffMel = 4 * kPs;             // 1056
constexpr int kOffSsc = kOffMel + 4 * kNfMax;  // 1312
constexpr int kOffLog = kOffSsc + 4 * kNfMax;  // 1568
static_assert(kOffLog + 4 * kRow <= kWarpFloats, "per-warp scratch overflows the exchange buffer");

// CTA-shared table blob (float units)
constexpr int kTwRow = 68;                      // floats per thread row of the W256 table (64 + 4 pad)
constexpr int kTabTw = 0;                       // [8][kTwRow]: row u = W256^{u k1}, k1 = 0..31 (float2)
constexpr int kTabTw512 = kTabTw + 8 * kTwRow;  // [256] float2: W512^k, k = 0..255
constexpr int kTabDct = kTabTw512 + 256 * 2;    // [kNcMax][kRow]
constexpr int kTabFixed = kTabDct + kNcMax * kRow;  // floats before the per-plan mel program
// mel program (per plan, nit iterations): int2 desc[nit][8] = {first bin (multiple of 4), flush},
// then float4 w[nit][8], then float4 w2[nit][8] (SSC).  flush: -2 keep accumulating, -1 end of a
// slot on an unused lane, >= 0 end of a slot: filter index to store.
constexpr int kMaxIt = 48;
static_assert(kTabTw512 % 4 == 0 && kTabDct % 4 == 0 && kTabFixed % 4 == 0,
              "16-byte alignment of table sections");

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
  if constexpr (R >= 16) dit_level<R, BITS, 16, R>(t);
  if constexpr (R >= 32) dit_level<R, BITS, 32, R>(t);
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
  int nit;            // iterations of the mel program
  int tab_floats;     // kTabFixed + nit * 8 * (2 + 4 + (SSC ? 4 : 0))
};

// one (k, 256-k) pair of the real-FFT post-processing:
//   E = (Za + conj Zb)/2, O = (Za - conj Zb)/(2i), T = W512^k O, X[k] = E + T, X[256-k] = conj(E - T).
// Returns 2|X[k]|, 2|X[256-k]| (the common 1/2 and the 1/n_fft of processing.rs:180 are folded
// into the mel weights and the energy scale).
__device__ __forceinline__ float2 rfft_pair_mag(float2 za, float2 zb, float2 w) {
  const float ex = za.x + zb.x, ey = za.y - zb.y;
  const float ox = za.y + zb.y, oy = zb.x - za.x;
  const float tx = ox * w.x - oy * w.y, ty = ox * w.y + oy * w.x;
  const float ax = ex + tx, ay = ey + ty, bx = ex - tx, by = ey - ty;
  float2 r;
  r.x = fast_sqrt(fmaf(ax, ax, ay * ay));
  r.y = fast_sqrt(fmaf(bx, bx, by * by));
  return r;
}

// JMAX: number of 16-sample column groups that can be non-zero (ceil(L/16));
// EXACT: L == 16*JMAX and every frame start is 8-byte aligned -> unpredicated float2 loads;
// SSC: also accumulate the spectral-subband-centroid numerator.
// Only bins <= 128 are kept in shared memory: with the reference's filterbank formula
// (feature.rs:69-70, quirk Q8) no filter has a non-zero tap above bin 128 when n_fft = 512;
// fast512_supported() declines any plan where that does not hold.
template <int JMAX, bool EXACT, bool SSC>
#ifndef SS_PREFETCH
#define SS_PREFETCH 0  // software-pipelined global loads: measured slower (register pressure -> spills)
#endif
#ifndef SS_FAST_MINB
#define SS_FAST_MINB 3
#endif
__global__ void __launch_bounds__(kThreads, SS_FAST_MINB) k_fast512(const FastArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* tab_s = smem;
  float* warp_s = smem + a.tab_floats + (threadIdx.x >> 5) * kWarpFloats;
  float* const stash0 = smem + a.tab_floats + kWarps * kWarpFloats;  // 2 x [T][nc] when fuse_cmvn
  float* stash = stash0;
  const FrameParams& p = a.p;

  for (int i = threadIdx.x; i < a.tab_floats; i += kThreads) tab_s[i] = __ldg(a.tab + i);
  __syncthreads();

  const float2* tw512_s = reinterpret_cast<const float2*>(tab_s + kTabTw512);
  const float* dct_s = tab_s + kTabDct;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = lane >> 3, u = lane & 7;  // frame within the quad, thread within the frame
  const float4* twrow = reinterpret_cast<const float4*>(tab_s + kTabTw + u * kTwRow);
  const int nf = p.nf, nc = p.nc, L = p.L;
  const int64_t T = p.T;
  const bool isz = (u == 0);
  // output groups owned by this thread (k = k1 + 32 k2): A/B and C/D are mirror partners
  const int k1B = isz ? 16 : 32 - u, k1C = isz ? 8 : 16 - u, k1D = isz ? 24 : 16 + u;
  const float mscale = 0.5f * p.mag_scale;  // the 1/2 of E/O and 1/n_fft (processing.rs:180)
  const bool want_mel = p.mfe || p.lmfe || p.ssc || p.mfcc;

  float2* xch = reinterpret_cast<float2*>(warp_s) + fi * kXchFrame;  // this frame's transpose buffer
  float* P_s = warp_s + kOffP;
  float* mel_s = warp_s + kOffMel;
  float* ssc_s = warp_s + kOffSsc;
  float* log_s = warp_s + kOffLog;
  float* Pf = P_s + fi * kPs;
  // post-processing addresses as per-thread base pointers + compile-time offsets
  const int kb = isz ? -112 : u;  // k of slots 4..7 is kb + 32 s (u == 0: 16 + 32 (s - 4))
  const float2* twA = tw512_s + u;
  const float2* twB = tw512_s + kb;
  const float2* twC = tw512_s + k1C;
  float* pA = Pf + u;
  float* pC = Pf + k1C;
  float* pCm = Pf + 256 - k1C;
  float* plo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) plo[i] = isz ? Pf + 16 + 32 * i : Pf + 256 - u - 32 * (i + 4);

  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int64_t clip = item / a.chunks_per_clip;
    const int chunk = (int)(item - clip * a.chunks_per_clip);
    const int q_end = min(a.quads_per_clip, (chunk + 1) * a.chunk_quads);
    const float* rowp = p.pcm + clip * p.row_stride;

    // framing: v[j] = x[2m] + i x[2m+1], m = u + 8 j; zero beyond L (processing.rs:147-156).
    // The loads for quad q + kWarps are issued in the middle of quad q (after its FFT registers
    // are dead), so their latency hides behind the mel / DCT / store phase.
    float2 v[32];
    auto load_quad = [&](int q) {
      const int64_t fr = 4 * (int64_t)q + fi;
      // an out-of-range frame of the last quad re-reads frame T-1 (always in bounds); its results
      // are never stored
      const float* px = rowp + (fr < T ? fr : T - 1) * p.S + 2 * u;
      sfor<0, 32>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        if constexpr (j >= JMAX) {
          v[j] = make_float2(0.0f, 0.0f);
        } else if constexpr (EXACT) {
          v[j] = __ldg(reinterpret_cast<const float2*>(px + 16 * j));
        } else {
          const int n0 = 2 * u + 16 * j;
          v[j].x = (n0 < L) ? __ldg(px + 16 * j) : 0.0f;
          v[j].y = (n0 + 1 < L) ? __ldg(px + 16 * j + 1) : 0.0f;
        }
      });
    };
    const int q_first = chunk * a.chunk_quads + warp;
#if SS_PREFETCH
    if (q_first < q_end) load_quad(q_first);
#endif

    for (int q = q_first; q < q_end; q += kWarps) {
      const int64_t fq = 4 * (int64_t)q;  // first frame of the quad
      const int64_t fr = fq + fi;
      const bool valid = fr < T;
#if !SS_PREFETCH
      load_quad(q);
#endif

      // ---- stage 1: 32-point FFT over j, in registers
      fft_regs<32, 5, JMAX>(v);
      // ---- twiddle W256^{u k1} (two per LDS.128) and transpose through shared memory
      __syncwarp();  // previous quad's readers of this buffer are done
      sfor<0, 16>([&](auto kk) {
        constexpr int k1 = 2 * decltype(kk)::value;
        const float4 w = twrow[k1 / 2];
        float2 y0 = v[k1], y1 = v[k1 + 1];
        if constexpr (k1 > 0) y0 = make_float2(y0.x * w.x - y0.y * w.y, y0.x * w.y + y0.y * w.x);
        y1 = make_float2(y1.x * w.z - y1.y * w.w, y1.x * w.w + y1.y * w.z);
        xch[k1 * kXs + u] = y0;
        xch[(k1 + 1) * kXs + u] = y1;
      });
      __syncwarp();
      float2 gA[8], gB[8], gC[8], gD[8];
      {
        const float4* rA = reinterpret_cast<const float4*>(xch + u * kXs);
        const float4* rB = reinterpret_cast<const float4*>(xch + k1B * kXs);
        const float4* rC = reinterpret_cast<const float4*>(xch + k1C * kXs);
        const float4* rD = reinterpret_cast<const float4*>(xch + k1D * kXs);
        sfor<0, 4>([&](auto tt) {
          constexpr int t = decltype(tt)::value;
          const float4 xa = rA[t], xb = rB[t], xc = rC[t], xd = rD[t];
          gA[2 * t] = make_float2(xa.x, xa.y); gA[2 * t + 1] = make_float2(xa.z, xa.w);
          gB[2 * t] = make_float2(xb.x, xb.y); gB[2 * t + 1] = make_float2(xb.z, xb.w);
          gC[2 * t] = make_float2(xc.x, xc.y); gC[2 * t + 1] = make_float2(xc.z, xc.w);
          gD[2 * t] = make_float2(xd.x, xd.y); gD[2 * t + 1] = make_float2(xd.z, xd.w);
        });
      }
      __syncwarp();  // the buffer is reused for P / mel / log below
      // ---- stage 2: 8-point FFTs over t.  g?[k2] = Z[k1? + 32 k2]
      fft_regs<8, 3, 8>(gA);
      fft_regs<8, 3, 8>(gB);
      fft_regs<8, 3, 8>(gC);
      fft_regs<8, 3, 8>(gD);

      // ---- real-FFT post-processing + magnitude (processing.rs:168,180) + frame energy.
      // u != 0: slot s pairs Z[u+32s] (gA[s]) with Z[256-(u+32s)] = gB[7-s]; likewise gC/gD.
      // u == 0: groups 0 and 16 are their own mirrors: s=0 -> k=0 (gives bins 0 and 256),
      //         s=1..3 -> (gA[s], gA[8-s]); s=4..7 -> (gB[s-4], gB[11-s]); bin 128 = |gA[4]| below.
      float esum = 0.f;
      sfor<0, 8>([&](auto ss_) {
        constexpr int s = decltype(ss_)::value;
        {
          float2 za, zb;
          if constexpr (s == 0) {
            za = gA[0];
            zb = isz ? gA[0] : gB[7];
          } else if constexpr (s < 4) {
            za = gA[s];
            zb = isz ? gA[8 - s] : gB[7 - s];
          } else {
            za = isz ? gB[s - 4] : gA[s];
            zb = isz ? gB[11 - s] : gB[7 - s];
          }
          // k = (s < 4 ? u : kb) + 32 s with kb = u, or -112 for u == 0: base pointer + immediate
          const float2 m = rfft_pair_mag(za, zb, (s < 4 ? twA : twB)[32 * s]);
          if constexpr (s < 4) {
            pA[32 * s] = m.x;                    // bin u + 32 s <= 103
          } else {
            *plo[s - 4] = isz ? m.x : m.y;       // the bin of the pair that is <= 128
          }
          esum += m.x + m.y;
        }
        {
          const float2 m = rfft_pair_mag(gC[s], gD[7 - s], twC[32 * s]);
          if constexpr (s < 4) {
            pC[32 * s] = m.x;                    // bin k1C + 32 s <= 111
          } else {
            pCm[-32 * s] = m.y;                  // bin 256 - k1C - 32 s <= 119
          }
          esum += m.x + m.y;
        }
      });
      if (isz) {  // bin 128: X[128] = conj(Z[128]); 2|X| to match the pairs
        const float m = 2.0f * fast_sqrt(fmaf(gA[4].x, gA[4].x, gA[4].y * gA[4].y));
        Pf[128] = m;
        esum += m;
        // tail of the last aligned group of 4 the mel program may read
        Pf[129] = 0.f; Pf[130] = 0.f; Pf[131] = 0.f;
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
      const float energy = zero_handling(mscale * esum);  // feature.rs:216-219
      __syncwarp();
#if SS_PREFETCH
      if (q + kWarps < q_end) load_quad(q + kWarps);  // v[] is dead from here on: prefetch the next quad
#endif

      // ---- banded mel reduction (feature.rs:229-230) as a flat "program": filters are sorted by
      // length and dealt to the 8 lanes slot by slot; every lane of a slot runs the slot's padded
      // length (zero weights beyond its own filter), four taps per iteration with 16-byte loads of
      // both the weights and P.  Padded taps only read bins this quad has just written, so no
      // stale shared memory can leak in through 0 * x.
      if (want_mel) {
        const int2* desc = reinterpret_cast<const int2*>(tab_s + kTabFixed) + u;
        const float4* pw = reinterpret_cast<const float4*>(tab_s + kTabFixed + a.nit * 16) + u;
        const float4* pw2 = pw + a.nit * 8;
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f), num4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float* melf = mel_s + fi * nf;
        float* sscf = ssc_s + fi * nf;
        float* logf_ = log_s + fi * kRow;
#pragma unroll 1
        for (int it = 0; it < a.nit; ++it) {
          const int2 d = desc[it * 8];
          const float4 w = pw[it * 8];
          const float4 pv = *reinterpret_cast<const float4*>(Pf + d.x);
          acc4.x = fmaf(w.x, pv.x, acc4.x); acc4.y = fmaf(w.y, pv.y, acc4.y);  // four independent chains
          acc4.z = fmaf(w.z, pv.z, acc4.z); acc4.w = fmaf(w.w, pv.w, acc4.w);
          if (SSC) {
            const float4 w2 = pw2[it * 8];
            num4.x = fmaf(w2.x, pv.x, num4.x); num4.y = fmaf(w2.y, pv.y, num4.y);
            num4.z = fmaf(w2.z, pv.z, num4.z); num4.w = fmaf(w2.w, pv.w, num4.w);
          }
          if (d.y != -2) {  // end of a slot (the same iteration on every lane)
            if (d.y >= 0) {
              const float acc = zero_handling((acc4.x + acc4.y) + (acc4.z + acc4.w));
              melf[d.y] = acc;
              if (SSC) sscf[d.y] = __fdividef((num4.x + num4.y) + (num4.z + num4.w), acc);
              logf_[d.y] = __logf(acc);
            }
            acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
            num4 = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        // zero the float4 padding of the log rows (the buffer held FFT data a moment ago)
        if (u < ((nf + 3) & ~3) - nf) log_s[fi * kRow + nf + u] = 0.0f;
        __syncwarp();
      }

      // ---- stores: [clip][t][nf] blocks, 4 consecutive frames = 4*nf contiguous floats
      const int64_t fbase = clip * T + fq;
      const int nvalid = (int)min((int64_t)4, T - fq);
      if (p.mfe || p.lmfe || (SSC && p.ssc)) {
        const int tot = nvalid * nf;
        for (int i = lane; i < tot; i += 32) {
          if (p.mfe) p.mfe[fbase * nf + i] = mel_s[i];
          if (SSC && p.ssc) p.ssc[fbase * nf + i] = ssc_s[i];
          if (p.lmfe) {
            const int f = (i >= nf) + (i >= 2 * nf) + (i >= 3 * nf);
            p.lmfe[fbase * nf + i] = log_s[f * kRow + (i - f * nf)];
          }
        }
      }
      if (p.energy && isz && valid) p.energy[clip * T + fr] = energy;
      // ---- ln -> DCT-II -> scale -> c0 (feature.rs:105-146); lane = (frame fi, cepstra u, u+8)
      if (p.mfcc) {
        const float4* lg = reinterpret_cast<const float4*>(log_s + fi * kRow);
        const int n4 = (nf + 3) >> 2;
        const int c1 = u + 8;
        const float4* d0 = reinterpret_cast<const float4*>(dct_s + u * kRow);
        const float4* d1 = reinterpret_cast<const float4*>(dct_s + (c1 < nc ? c1 : u) * kRow);
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int i = 0; i < n4; ++i) {
          const float4 x = lg[i], da = d0[i], db = d1[i];
          s0.x = fmaf(da.x, x.x, s0.x); s1.x = fmaf(db.x, x.x, s1.x);  // eight independent chains
          s0.y = fmaf(da.y, x.y, s0.y); s1.y = fmaf(db.y, x.y, s1.y);
          s0.z = fmaf(da.z, x.z, s0.z); s1.z = fmaf(db.z, x.z, s1.z);
          s0.w = fmaf(da.w, x.w, s0.w); s1.w = fmaf(db.w, x.w, s1.w);
        }
        float acc0 = (s0.x + s0.y) + (s0.z + s0.w), acc1 = (s1.x + s1.y) + (s1.z + s1.w);
        if (u == 0) {
          if (p.dc_elim) acc0 = __logf(energy);
          else if (fr == 0) acc0 *= p.scale_00;  // only element [0,0] gets the k=0 scale (feature.rs:128)
        } else {
          acc0 *= p.scale_rest;
        }
        acc1 *= p.scale_rest;
        if (valid) {
          float* dst = a.fuse_cmvn ? stash + fr * nc : p.mfcc + (clip * T + fr) * nc;
          if (u < nc) dst[u] = acc0;
          if (c1 < nc) dst[c1] = acc1;
        }
      }
    }

    // ---- per-clip CMVN (processing.rs:265-300) on the stashed [T][nc] block.  The stash is
    // double-buffered so ONE barrier per clip suffices: it both publishes this clip's stash and
    // proves every warp has finished reading the other buffer (the previous clip's CMVN).
    if (a.fuse_cmvn) {
      __syncthreads();
      const float* st = stash;
      const int Ti = (int)T;
      float* dst = p.mfcc + clip * T * nc;
      for (int c = warp; c < nc; c += kWarps) {  // a warp owns whole columns: no further barrier
        float s0 = 0.f;
        for (int t = lane; t < Ti; t += 32) s0 += st[t * nc + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        const float mean = s0 / (float)Ti;
        float inv = 1.0f;
        if (a.fuse_cmvn == 2) {
          // std of the mean-subtracted column y, ddof 0: sqrt(mean(y^2) - mean(y)^2), mean(y) ~ 0
          float s1 = 0.f, s2 = 0.f;
          for (int t = lane; t < Ti; t += 32) {
            const float d = st[t * nc + c] - mean;
            s1 += d;
            s2 = fmaf(d, d, s2);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          const float my = s1 / (float)Ti;
          inv = 1.0f / (sqrtf(fmaxf(s2 / (float)Ti - my * my, 0.f)) + kCmvnEps);
        }
        for (int t = lane; t < Ti; t += 32) dst[t * nc + c] = (st[t * nc + c] - mean) * inv;
      }
      stash = (stash == stash0) ? stash0 + a.stash_floats : stash0;
    }
  }
}

}  // namespace

struct Fast512Tables {
  float* d_tab = nullptr;
  bool store_hi = true;
  int nit = 0;  // iterations of the mel program
};

// The mel program: filters sorted by length and dealt round-robin to 8 lanes; slot s holds
// filters 8s..8s+7.  A slot runs ceil(span/4) iterations, span = the longest 4-aligned run
// [start & ~3, start + len) among its filters.
struct MelProgram {
  int nit = 0;
  std::vector<int> desc;     // [nit][8][2]
  std::vector<float> w, w2;  // [nit][8][4]
};
static void build_mel_program(const Derived& d, const MelBank& mb, MelProgram* mp, bool* store_hi_out) {
  const int nf = d.nf;
  std::vector<int> order(nf);
  for (int i = 0; i < nf; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return mb.len[x] > mb.len[y]; });
  const bool store_hi = mb.max_bin > 128;
  const int hi_end = store_hi ? 260 : 132;  // the kernel writes bins [0, hi_end) afresh in every quad
  const float mscale = 0.5f * (1.0f / (float)d.W);  // the kernel keeps 2*n_fft*|X| in shared memory
  const float r0 = 1.0f, step = d.F > 1 ? ((float)d.cfg.sample_rate / 2.0f - 1.0f) / (float)(d.F - 1) : 0.f;
  mp->nit = 0;
  mp->desc.clear(); mp->w.clear(); mp->w2.clear();
  const int ns = (nf + 7) / 8;
  for (int s = 0; s < ns; ++s) {
    int its = 1;
    int b0[8];
    for (int l = 0; l < 8; ++l) {
      b0[l] = 0;
      if (s * 8 + l >= nf) continue;
      const int f = order[s * 8 + l];
      if (mb.len[f] == 0) continue;
      b0[l] = mb.start[f] & ~3;
      its = std::max(its, (mb.start[f] + mb.len[f] - b0[l] + 3) / 4);
    }
    // keep every (padded) read inside the bins the kernel has freshly written
    for (int l = 0; l < 8; ++l) b0[l] = std::max(0, std::min(b0[l], hi_end - 4 * its));
    for (int it = 0; it < its; ++it)
      for (int l = 0; l < 8; ++l) {
        const bool used = s * 8 + l < nf;
        const int f = used ? order[s * 8 + l] : -1;
        mp->desc.push_back(b0[l] + 4 * it);
        mp->desc.push_back(it == its - 1 ? (used ? f : -1) : -2);
        for (int k = 0; k < 4; ++k) {
          const int bin = b0[l] + 4 * it + k;
          float w = 0.f;
          if (used && bin >= mb.start[f] && bin < mb.start[f] + mb.len[f]) w = mb.w[mb.wofs[f] + bin - mb.start[f]];
          mp->w.push_back(w * mscale);
          mp->w2.push_back(w * (r0 + step * (float)bin) * mscale);
        }
      }
    mp->nit += its;
  }
  *store_hi_out = store_hi;
}

bool fast512_supported(const Derived& d, const MelBank& mb) {
  if (d.W != 512 || !d.speechpy_ok) return false;
  if (d.nf > kNfMax || d.nc > kNcMax) return false;
  MelProgram mp;
  bool hi;
  build_mel_program(d, mb, &mp, &hi);
  if (mp.nit > kMaxIt) return false;
  if (hi) return false;                                  // a filter reaches above bin 128: generic kernel
  if (4 * ((mb.max_len + 6) / 4) > 132) return false;    // a padded run must fit the written bins
  return true;
}

int fast512_build_tables(const Derived& d, const MelBank& mb, Fast512Tables** out) {
  MelProgram mp;
  bool store_hi;
  build_mel_program(d, mb, &mp, &store_hi);
  std::vector<float> h(kTabFixed + (size_t)mp.nit * 8 * 10, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  for (int u = 0; u < 8; ++u)
    for (int k1 = 0; k1 < 32; ++k1) {
      const double ang = -2.0 * pi * (double)(u * k1) / 256.0;
      h[kTabTw + u * kTwRow + 2 * k1] = (float)std::cos(ang);
      h[kTabTw + u * kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
    }
  for (int k = 0; k < 256; ++k) {
    const double ang = -2.0 * pi * (double)k / 512.0;
    h[kTabTw512 + 2 * k] = (float)std::cos(ang);
    h[kTabTw512 + 2 * k + 1] = (float)std::sin(ang);
  }
  std::vector<float> dct;
  build_dct(d, &dct);
  for (int c = 0; c < d.nc; ++c)
    for (int n = 0; n < d.nf; ++n) h[kTabDct + c * kRow + n] = dct[(size_t)c * d.nf + n];
  int* desc = reinterpret_cast<int*>(h.data() + kTabFixed);
  for (size_t i = 0; i < mp.desc.size(); ++i) desc[i] = mp.desc[i];
  float* w = h.data() + kTabFixed + (size_t)mp.nit * 16;
  float* w2 = w + (size_t)mp.nit * 32;
  for (size_t i = 0; i < mp.w.size(); ++i) { w[i] = mp.w[i]; w2[i] = mp.w2[i]; }
  Fast512Tables* t = new Fast512Tables;
  t->store_hi = store_hi;
  t->nit = mp.nit;
  cudaError_t e = cudaMalloc(&t->d_tab, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
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
  const bool fuse = p.mfcc && p.cmvn_mode && stash_bytes <= 12 * 1024;
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
  a.nit = t->nit;
  a.tab_floats = kTabFixed + t->nit * 8 * (2 + 4 + (p.ssc ? 4 : 0));
  const size_t smem = (size_t)(a.tab_floats + kWarps * kWarpFloats + 2 * a.stash_floats) * sizeof(float);
  const int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sm_count * SS_FAST_MINB);
  if (grid <= 0) return 0;
  // float2 loads need every frame start 8-byte aligned: even stride / step and an aligned base
  const bool vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  const bool ssc = p.ssc != nullptr;
  cudaError_t e;
#define SS_LAUNCH_FAST(J, X, C)                                                                            \
  do {                                                                                                     \
    e = cudaFuncSetAttribute(k_fast512<J, X, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e != cudaSuccess) return (int)e;                                                                   \
    k_fast512<J, X, C><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);                       \
  } while (0)
  if (vec2 && p.L == 320) { if (ssc) SS_LAUNCH_FAST(20, true, true); else SS_LAUNCH_FAST(20, true, false); }
  else if (vec2 && p.L == 400) { if (ssc) SS_LAUNCH_FAST(25, true, true); else SS_LAUNCH_FAST(25, true, false); }
  else { if (ssc) SS_LAUNCH_FAST(32, false, true); else SS_LAUNCH_FAST(32, false, false); }
#undef SS_LAUNCH_FAST
  count_launch();
  note_kernel("k_fast512");
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return (p.cmvn_mode && !fuse) ? -1001 : 0;
}

}  // namespace ss
