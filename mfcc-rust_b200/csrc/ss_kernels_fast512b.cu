// Fused speechpy-framing feature kernel for n_fft = 512, second generation (c1 / c2 / c5):
//   framing -> zero-pad -> real FFT -> |X|/N -> frame energy -> banded mel -> ln -> DCT-II
//   (+ SSC, + per-clip CMVN) in one pass; every intermediate in registers / shared memory.
// Replaces feature::{mfe, lmfe, mfcc} (feature.rs:99-245) + processing::{stack_frames,
// power_spectrum, cmvn} (processing.rs:65-181,265-300) for this FFT size.
//
// Versus the first generation (removed; 8 lanes x 32 points per frame, 128
// registers, 15 warps/SM) this kernel uses 16 lanes x 16 points per frame:
//   * real FFT(512) = complex FFT(256) of z[m] = x[2m] + i x[2m+1] + split; 256 = 16 x 16:
//     16-point register FFT (compile-time twiddles, zero-padded inputs pruned), W256^{u k1}
//     twiddle (two per LDS.128), one warp-private conflict-free transpose, second 16-point FFT;
//   * lane u then holds Z[u + 16 k2]; its mirror Z[256 - k] sits in lane 16 - u and is fetched
//     with 16 warp shuffles (as in k_stft2048), so no second shared-memory pass;
//   * half the registers -> twice the resident warps, and the epilogue (mel program, DCT, stores)
//     spreads over twice the lanes per frame: ~20 % fewer warp instructions per frame.
// A warp owns 2 consecutive frames; a CTA of 7 warps covers a 97-frame clip in exactly 7 rounds.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

constexpr int kWarps = 7, kThreads = kWarps * 32;
constexpr int kXs = 18;                      // transpose row stride in float2 (144 B: conflict-free LDS.128)
constexpr int kXchFrame = 16 * kXs;          // 288 float2 per frame
constexpr int kWarpFloats = 2 * kXchFrame * 2;  // 1152 floats per warp (2 frames)
constexpr int kPs = 144;                     // P row: bins 0..128 + zero pad to 131; 144 % 32 == 16 keeps the two frames on disjoint banks
constexpr int kNfMax = 64, kNcMax = 16, kMaxIt = 40, kSlots = 4;  // nf <= 64 -> at most 4 slots of 16 filters
constexpr int kRow = 68;                     // row stride of log_s and the DCT rows
// per-warp overlay once the transpose tile is consumed (floats)
constexpr int kOffP = 0;                     // [2][kPs]
constexpr int kOffMel = 2 * kPs;             // 264: [2][nf]
constexpr int kOffSsc = kOffMel + 2 * kNfMax;  // 392
constexpr int kOffLog = kOffSsc + 2 * kNfMax;  // 520: [2][kRow]
static_assert(kOffLog + 2 * kRow <= kWarpFloats, "per-warp scratch overflows the transpose tile");
// CTA-shared table blob (floats)
constexpr int kTwRow = 36;                        // floats per lane row of the W256 table (16 float2 + pad)
constexpr int kTabTw = 0;                         // [16][kTwRow]
constexpr int kTabDct = kTabTw + 16 * kTwRow;     // [kNcMax][kRow]
constexpr int kTabSlot = kTabDct + kNcMax * kRow; // int [kSlots][16][2] = first bin, filter (-1: none) per slot and lane
constexpr int kTabFixed = kTabSlot + kSlots * 32; // then float4 w[nit][16], float4 w2[nit][16]
static_assert(kTabDct % 4 == 0 && kTabSlot % 4 == 0 && kTabFixed % 4 == 0, "table alignment");

__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct FastArgs {
  FrameParams p;
  const float* tab;
  int64_t n_items;    // clips * chunks_per_clip
  int pairs_per_clip, chunk_pairs, chunks_per_clip;
  int fuse_cmvn;      // 0 none, 1 mean, 2 mean+var: whole clip in one CTA, stash in smem
  int stash_floats;
  int nit, tab_floats;
  int vec2;           // every frame start is 8-byte aligned
  int n_slots;        // mel program: number of slots and iterations per slot (warp-uniform)
  int slot_its[kSlots];
};

// JMAX: column groups (32 samples each) that can be non-zero; EXACT: JMAX == ceil(L / 32), so only
// the last group needs a bounds check; SSC: also the centroid numerator
template <int JMAX, bool EXACT, bool SSC>
__global__ void __launch_bounds__(kThreads, 4) k_fast512b(const FastArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* tab_s = smem;
  float* warp_s = smem + a.tab_floats + (threadIdx.x >> 5) * kWarpFloats;
  float* const stash0 = smem + a.tab_floats + kWarps * kWarpFloats;  // 2 x [T][nc] when fuse_cmvn
  float* stash = stash0;
  const FrameParams& p = a.p;
  for (int i = threadIdx.x; i < a.tab_floats; i += kThreads) tab_s[i] = __ldg(a.tab + i);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = lane >> 4, u = lane & 15;
  const bool isz = (u == 0);
  const int partner = (lane & 16) | ((16 - u) & 15);
  const int nf = p.nf, nc = p.nc, L = p.L;
  const int64_t T = p.T;
  const float mscale = 0.5f * p.mag_scale;  // the 1/2 of E, O and 1/n_fft (processing.rs:180)
  const bool want_mel = p.mfe || p.lmfe || p.ssc || p.mfcc;

  const float4* twrow = reinterpret_cast<const float4*>(tab_s + kTabTw + u * kTwRow);
  const float* dct_s = tab_s + kTabDct;
  const float4* pw = reinterpret_cast<const float4*>(tab_s + kTabFixed) + u;
  const float4* pw2 = pw + a.nit * 16;
  // W512^{u + 16 k2} = W512^u * W32^{k2}: one per-lane register pair times a compile-time constant
  float2 wu;
  sincospif(-(float)u / 256.0f, &wu.y, &wu.x);
  // mel program: per slot this lane's first bin and filter index (registers, loaded once)
  int slot_b0[kSlots], slot_f[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int2 e = reinterpret_cast<const int2*>(tab_s + kTabSlot)[s * 16 + u];
    slot_b0[s] = e.x;
    slot_f[s] = e.y;
  }

  float2* xch = reinterpret_cast<float2*>(warp_s) + fi * kXchFrame;
  float* Pf = warp_s + kOffP + fi * kPs;
  float* pA = Pf + u;
  float* mel_s = warp_s + kOffMel;
  float* ssc_s = warp_s + kOffSsc;
  float* log_s = warp_s + kOffLog;
  // does this lane's last column group (j = JMAX - 1) lie inside the frame?
  const bool last0 = 2 * u + 32 * (JMAX - 1) < L, last1 = 2 * u + 32 * (JMAX - 1) + 1 < L;

  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int64_t clip = item / a.chunks_per_clip;
    const int chunk = (int)(item - clip * a.chunks_per_clip);
    const int q_end = min(a.pairs_per_clip, (chunk + 1) * a.chunk_pairs);
    const float* rowp = p.pcm + clip * p.row_stride;

    for (int q = chunk * a.chunk_pairs + warp; q < q_end; q += kWarps) {
      const int64_t fq = 2 * (int64_t)q;  // first frame of the pair
      const int64_t fr = fq + fi;
      const bool valid = fr < T;

      // ---- framing: v[j] = x[2m] + i x[2m+1], m = u + 16 j; zero beyond L (processing.rs:147-156).
      // An out-of-range frame of the last pair re-reads frame T-1 (in bounds); it is never stored.
      float2 v[16];
      {
        const float* px = rowp + (valid ? fr : T - 1) * p.S + 2 * u;
        sfor<0, 16>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          if constexpr (j >= JMAX) {
            v[j] = make_float2(0.0f, 0.0f);
          } else if constexpr (!EXACT) {
            v[j].x = (2 * u + 32 * j < L) ? __ldg(px + 32 * j) : 0.0f;
            v[j].y = (2 * u + 32 * j + 1 < L) ? __ldg(px + 32 * j + 1) : 0.0f;
          } else if constexpr (j == JMAX - 1) {
            v[j].x = last0 ? __ldg(px + 32 * j) : 0.0f;
            v[j].y = last1 ? __ldg(px + 32 * j + 1) : 0.0f;
          } else {
            if (a.vec2) v[j] = __ldg(reinterpret_cast<const float2*>(px + 32 * j));
            else v[j] = make_float2(__ldg(px + 32 * j), __ldg(px + 32 * j + 1));
          }
        });
      }
      // ---- stage 1: 16-point FFT over j
      ssfft::fft_pow2<16, JMAX>(v);
      // ---- twiddle W256^{u k1} and transpose
      __syncwarp();
      sfor<0, 8>([&](auto kk) {
        constexpr int k1 = 2 * decltype(kk)::value;
        const float4 w = twrow[k1 / 2];
        float2 y0 = v[k1], y1 = v[k1 + 1];
        if constexpr (k1 > 0) y0 = ssfft::cmulf(y0, make_float2(w.x, w.y));
        y1 = ssfft::cmulf(y1, make_float2(w.z, w.w));
        xch[k1 * kXs + u] = y0;
        xch[(k1 + 1) * kXs + u] = y1;
      });
      __syncwarp();
      float2 g[16];
      {
        const float4* r4 = reinterpret_cast<const float4*>(xch + u * kXs);
        sfor<0, 8>([&](auto tt) {
          constexpr int i = decltype(tt)::value;
          const float4 x = r4[i];
          g[2 * i] = make_float2(x.x, x.y);
          g[2 * i + 1] = make_float2(x.z, x.w);
        });
      }
      __syncwarp();  // the tile is reused for P / mel / log below
      // ---- stage 2: 16-point FFT over the lane index.  g[k2] = Z[u + 16 k2]
      ssfft::fft_pow2<16>(g);

      // ---- real-FFT split for k = u + 16 k2 <= 127 (k2 < 8): Z[256 - k] is g[15 - k2] of lane
      // 16 - u (lane 0: its own g[16 - k2]).  Each pair yields |X[k]| and |X[256 - k]|; both go
      // into the frame energy, only the bins a filter can reach (<= 128) into shared memory.
      float esum = 0.f;
      sfor<0, 8>([&](auto kk) {
        constexpr int k2 = decltype(kk)::value;
        float2 zb;
        zb.x = __shfl_sync(0xffffffffu, g[15 - k2].x, partner);
        zb.y = __shfl_sync(0xffffffffu, g[15 - k2].y, partner);
        if (isz) zb = g[(16 - k2) & 15];
        const float2 za = g[k2];
        const float2 e = __fadd2_rn(za, make_float2(zb.x, -zb.y));                         // E' = za + conj zb
        const float2 od = __fadd2_rn(make_float2(za.y, -za.x), make_float2(zb.y, zb.x));   // O' = (za - conj zb) / i
        const float2 tt = ssfft::cmulf(ssfft::twmul<32, k2>(od), wu);                      // O' W32^{k2} W512^u
        const float2 xa = ssfft::cadd(e, tt), xb = ssfft::csub(e, tt);
        const float ma = fast_sqrt(fmaf(xa.x, xa.x, xa.y * xa.y));  // 2 n_fft |X[k]| (scale folded into the weights)
        const float mb = fast_sqrt(fmaf(xb.x, xb.x, xb.y * xb.y));  // 2 n_fft |X[256-k]|
        pA[16 * k2] = ma;
        esum += ma + mb;
      });
      if (isz) {  // bin 128 (self-mirrored): X[128] = conj Z[128]; x2 to match the pairs
        const float m = 2.0f * fast_sqrt(fmaf(g[8].x, g[8].x, g[8].y * g[8].y));
        Pf[128] = m;
        Pf[129] = 0.f; Pf[130] = 0.f; Pf[131] = 0.f;  // tail of the last aligned group of 4
        esum += m;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
      const float energy = zero_handling(mscale * esum);  // feature.rs:216-219
      __syncwarp();

      // ---- banded mel reduction (feature.rs:229-230) as a flat program on 16 lanes: filters sorted
      // by length and dealt to the lanes slot by slot, four taps per iteration with 16-byte loads of
      // weights and P, warp-uniform trip counts.  Padded taps only read bins written in this pair.
      if (want_mel) {
        float* melf = mel_s + fi * nf;
        float* sscf = ssc_s + fi * nf;
        float* logf_ = log_s + fi * kRow;
        int it = 0;
#pragma unroll
        for (int sidx = 0; sidx < kSlots; ++sidx) {
          if (sidx < a.n_slots) {
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f), num4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* pp = Pf + slot_b0[sidx];
            const int its = a.slot_its[sidx];
#pragma unroll 1
            for (int i = 0; i < its; ++i, ++it) {
              const float4 w = pw[it * 16];
              const float4 pv = *reinterpret_cast<const float4*>(pp + 4 * i);
              acc4.x = fmaf(w.x, pv.x, acc4.x); acc4.y = fmaf(w.y, pv.y, acc4.y);
              acc4.z = fmaf(w.z, pv.z, acc4.z); acc4.w = fmaf(w.w, pv.w, acc4.w);
              if (SSC) {
                const float4 w2 = pw2[it * 16];
                num4.x = fmaf(w2.x, pv.x, num4.x); num4.y = fmaf(w2.y, pv.y, num4.y);
                num4.z = fmaf(w2.z, pv.z, num4.z); num4.w = fmaf(w2.w, pv.w, num4.w);
              }
            }
            if (slot_f[sidx] >= 0) {
              const float acc = zero_handling((acc4.x + acc4.y) + (acc4.z + acc4.w));
              melf[slot_f[sidx]] = acc;
              if (SSC) sscf[slot_f[sidx]] = __fdividef((num4.x + num4.y) + (num4.z + num4.w), acc);
              logf_[slot_f[sidx]] = __logf(acc);
            }
          }
        }
        if (u < ((nf + 3) & ~3) - nf) logf_[nf + u] = 0.0f;  // float4 padding of the log row
        __syncwarp();
      }

      // ---- stores: [clip][t][nf], 2 consecutive frames = 2*nf contiguous floats
      const int64_t fbase = clip * T + fq;
      const int nvalid = (int)min((int64_t)2, T - fq);
      if (p.mfe || p.lmfe || (SSC && p.ssc)) {
        const int tot = nvalid * nf;
        for (int i = lane; i < tot; i += 32) {
          if (p.mfe) p.mfe[fbase * nf + i] = mel_s[i];
          if (SSC && p.ssc) p.ssc[fbase * nf + i] = ssc_s[i];
          if (p.lmfe) {
            const int f = (i >= nf);
            p.lmfe[fbase * nf + i] = log_s[f * kRow + (i - f * nf)];
          }
        }
      }
      if (p.energy && isz && valid) p.energy[clip * T + fr] = energy;
      // ---- ln -> DCT-II -> scale -> c0 (feature.rs:105-146); lane = (frame fi, cepstrum u)
      if (p.mfcc && u < nc) {
        const float4* lg = reinterpret_cast<const float4*>(log_s + fi * kRow);
        const float4* dr = reinterpret_cast<const float4*>(dct_s + u * kRow);
        const int n4 = (nf + 3) >> 2;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int i = 0; i < n4; ++i) {
          const float4 x = lg[i], d = dr[i];
          s4.x = fmaf(d.x, x.x, s4.x); s4.y = fmaf(d.y, x.y, s4.y);
          s4.z = fmaf(d.z, x.z, s4.z); s4.w = fmaf(d.w, x.w, s4.w);
        }
        float acc = (s4.x + s4.y) + (s4.z + s4.w);
        if (u == 0) {
          if (p.dc_elim) acc = __logf(energy);
          else if (fr == 0) acc *= p.scale_00;  // only element [0,0] gets the k=0 scale (feature.rs:128)
        } else {
          acc *= p.scale_rest;
        }
        if (valid) {
          if (a.fuse_cmvn) stash[fr * nc + u] = acc;
          else p.mfcc[(clip * T + fr) * nc + u] = acc;
        }
      }
    }

    // ---- per-clip CMVN (processing.rs:265-300): one barrier per clip, double-buffered stash, each
    // warp owns whole columns (mean, variance of the mean-subtracted column, normalise, store)
    if (a.fuse_cmvn) {
      __syncthreads();
      const float* st = stash;
      const int Ti = (int)T;
      float* dst = p.mfcc + clip * T * nc;
      for (int c = warp; c < nc; c += kWarps) {
        float s0 = 0.f;
        for (int t = lane; t < Ti; t += 32) s0 += st[t * nc + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        const float mean = s0 / (float)Ti;
        float inv = 1.0f;
        if (a.fuse_cmvn == 2) {
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

struct MelProgramB {
  int nit = 0, n_slots = 0;
  int slot_its[kSlots] = {0, 0, 0, 0};
  std::vector<int> slot;     // [kSlots][16][2] = first bin, filter
  std::vector<float> w, w2;  // [nit][16][4]
};

// Filters sorted by length, 16 per slot.  Inside a slot the filters are dealt to the lanes so that
// the 16-byte P reads of each quarter-warp (lanes 0-7, 8-15) fall into distinct bank groups.
static bool build_program(const Derived& d, const MelBank& mb, MelProgramB* mp) {
  const int nf = d.nf, NL = 16;
  if (mb.max_bin > 128) return false;  // only bins <= 128 are kept (always true for the reference's bank at n_fft 512)
  std::vector<int> order(nf);
  for (int i = 0; i < nf; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return mb.len[x] > mb.len[y]; });
  const float mscale = 0.5f * (1.0f / (float)d.W);  // the kernel keeps 2 n_fft |X| in shared memory
  const float r0 = 1.0f, step = d.F > 1 ? ((float)d.cfg.sample_rate / 2.0f - 1.0f) / (float)(d.F - 1) : 0.f;
  const int ns = (nf + NL - 1) / NL;
  if (ns > kSlots) return false;
  mp->nit = 0;
  mp->n_slots = ns;
  mp->slot.assign(kSlots * 32, 0);
  for (int s = 0; s < kSlots; ++s)
    for (int l = 0; l < NL; ++l) mp->slot[(s * NL + l) * 2 + 1] = -1;
  for (int s = 0; s < ns; ++s) {
    int its = 1;
    std::vector<int> fl, b0s;  // filters of this slot and their aligned first bins
    for (int l = 0; l < NL && s * NL + l < nf; ++l) {
      const int f = order[s * NL + l];
      const int b = mb.len[f] ? (mb.start[f] & ~3) : 0;
      fl.push_back(f);
      b0s.push_back(b);
      if (mb.len[f]) its = std::max(its, (mb.start[f] + mb.len[f] - b + 3) / 4);
    }
    if (4 * its > kPs) return false;
    for (auto& b : b0s) b = std::max(0, std::min(b, 132 - 4 * its));  // bins [0, 132) are written afresh per pair
    // greedy lane assignment: two quarters of 8 lanes, each wants distinct (b0 / 4) % 8
    int lane_f[16], lane_b[16];
    bool used_lane[16] = {false};
    for (int l = 0; l < NL; ++l) { lane_f[l] = -1; lane_b[l] = 0; }
    for (size_t i = 0; i < fl.size(); ++i) {
      int best = -1, best_cost = 1 << 30;
      for (int l = 0; l < NL; ++l) {
        if (used_lane[l]) continue;
        int cost = 0;
        for (int m = (l & 8); m < (l & 8) + 8; ++m)
          if (used_lane[m] && lane_f[m] >= 0 && ((lane_b[m] >> 2) & 7) == ((b0s[i] >> 2) & 7) && lane_b[m] != b0s[i]) ++cost;
        if (cost < best_cost) { best_cost = cost; best = l; }
      }
      used_lane[best] = true;
      lane_f[best] = fl[i];
      lane_b[best] = b0s[i];
    }
    mp->slot_its[s] = its;
    for (int l = 0; l < NL; ++l) {
      mp->slot[(s * NL + l) * 2] = lane_b[l];
      mp->slot[(s * NL + l) * 2 + 1] = lane_f[l];
    }
    for (int it = 0; it < its; ++it)
      for (int l = 0; l < NL; ++l) {
        const int f = lane_f[l];
        for (int k = 0; k < 4; ++k) {
          const int bin = lane_b[l] + 4 * it + k;
          float w = 0.f;
          if (f >= 0 && bin >= mb.start[f] && bin < mb.start[f] + mb.len[f]) w = mb.w[mb.wofs[f] + bin - mb.start[f]];
          mp->w.push_back(w * mscale);
          mp->w2.push_back(w * (r0 + step * (float)bin) * mscale);
        }
      }
    mp->nit += its;
  }
  return mp->nit <= kMaxIt;
}

}  // namespace

struct Fast512bTables {
  float* d_tab = nullptr;
  int nit = 0, n_slots = 0;
  int slot_its[4] = {0, 0, 0, 0};
};

bool fast512b_supported(const Derived& d, const MelBank& mb) {
  if (d.W != 512 || !d.speechpy_ok || d.nf > kNfMax || d.nc > kNcMax) return false;
  MelProgramB mp;
  return build_program(d, mb, &mp);
}

int fast512b_build_tables(const Derived& d, const MelBank& mb, Fast512bTables** out) {
  MelProgramB mp;
  if (!build_program(d, mb, &mp)) return (int)cudaErrorInvalidValue;
  std::vector<float> h(kTabFixed + (size_t)mp.nit * 16 * 8, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  for (int u = 0; u < 16; ++u)
    for (int k1 = 0; k1 < 16; ++k1) {
      const double ang = -2.0 * pi * (double)(u * k1) / 256.0;
      h[kTabTw + u * kTwRow + 2 * k1] = (float)std::cos(ang);
      h[kTabTw + u * kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
    }
  std::vector<float> dct;
  build_dct(d, &dct);
  for (int c = 0; c < d.nc; ++c)
    for (int n = 0; n < d.nf; ++n) h[kTabDct + c * kRow + n] = dct[(size_t)c * d.nf + n];
  int* slot = reinterpret_cast<int*>(h.data() + kTabSlot);
  for (size_t i = 0; i < mp.slot.size(); ++i) slot[i] = mp.slot[i];
  float* w = h.data() + kTabFixed;
  float* w2 = w + (size_t)mp.nit * 64;
  for (size_t i = 0; i < mp.w.size(); ++i) { w[i] = mp.w[i]; w2[i] = mp.w2[i]; }
  Fast512bTables* t = new Fast512bTables;
  t->nit = mp.nit;
  t->n_slots = mp.n_slots;
  for (int i = 0; i < 4; ++i) t->slot_its[i] = mp.slot_its[i];
  cudaError_t e = cudaMalloc(&t->d_tab, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return (int)e; }
  *out = t;
  return 0;
}

void fast512b_free_tables(Fast512bTables* t) {
  if (!t) return;
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

// returns 0 (done, including any requested cmvn), -1000 (shape declined), -1001 (features
// written, cmvn left to the caller) or a cudaError_t
int launch_fast512b(const FrameParams& p, const Fast512bTables* t, int sm_count, void* stream) {
  if (p.N != 512 || p.L > 512 || p.T < 1 || p.T > (1LL << 31) - 8) return -1000;
  FastArgs a;
  a.p = p;
  a.tab = t->d_tab;
  a.nit = t->nit;
  const int64_t ppc = (p.T + 1) / 2;
  a.pairs_per_clip = (int)ppc;
  const size_t stash_bytes = (size_t)p.T * p.nc * sizeof(float);
  const bool fuse = p.mfcc && p.cmvn_mode && stash_bytes <= 12 * 1024;
  a.fuse_cmvn = fuse ? p.cmvn_mode : 0;
  a.stash_floats = fuse ? (int)(p.T * p.nc) : 0;
  if (fuse) {
    a.chunk_pairs = (int)ppc;
  } else {
    const int64_t total = ppc * p.n_rows;
    int64_t cq = total / ((int64_t)sm_count * 4 * 8);
    cq = std::max<int64_t>(kWarps, std::min<int64_t>(cq, 8 * kWarps));
    cq = (cq + kWarps - 1) / kWarps * kWarps;
    a.chunk_pairs = (int)std::min<int64_t>(cq, std::max<int64_t>(ppc, 1));
  }
  a.chunks_per_clip = (int)((ppc + a.chunk_pairs - 1) / a.chunk_pairs);
  a.n_items = p.n_rows * a.chunks_per_clip;
  const bool ssc = p.ssc != nullptr;
  a.tab_floats = kTabFixed + t->nit * 16 * (4 + (ssc ? 4 : 0));
  a.n_slots = t->n_slots;
  for (int i = 0; i < kSlots; ++i) a.slot_its[i] = t->slot_its[i];
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  const size_t smem = (size_t)(a.tab_floats + kWarps * kWarpFloats + 2 * a.stash_floats) * sizeof(float);
  const int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sm_count * 4);
  if (grid <= 0) return 0;
  const int jn = (p.L + 31) / 32;
  cudaError_t e;
#define SS_LAUNCH_FASTB(J, X, C)                                                                            \
  do {                                                                                                      \
    e = cudaFuncSetAttribute(k_fast512b<J, X, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e != cudaSuccess) return (int)e;                                                                    \
    k_fast512b<J, X, C><<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(a);                       \
  } while (0)
  if (jn == 10) { if (ssc) SS_LAUNCH_FASTB(10, true, true); else SS_LAUNCH_FASTB(10, true, false); }        // L = 289..320
  else if (jn == 13) { if (ssc) SS_LAUNCH_FASTB(13, true, true); else SS_LAUNCH_FASTB(13, true, false); }   // L = 385..416
  else { if (ssc) SS_LAUNCH_FASTB(16, false, true); else SS_LAUNCH_FASTB(16, false, false); }
#undef SS_LAUNCH_FASTB
  count_launch();
  note_kernel("k_fast512b");
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return (p.cmvn_mode && !fuse) ? -1001 : 0;
}

}  // namespace ss
