// Fused speechpy-framing feature kernel for n_fft = 512, third generation (c1 / c2 / c5):
//   framing -> zero-pad -> real FFT -> |X|/N -> frame energy -> mel filterbank -> ln -> DCT-II
//   (+ SSC, + per-clip CMVN) in one pass; every intermediate in registers / shared memory.
// Replaces feature::{mfe, lmfe, mfcc} (feature.rs:99-245) + processing::{stack_frames,
// power_spectrum, cmvn} (processing.rs:65-181,265-300) for this FFT size.
//
// FFT stage as in the second generation (ss_kernels_fast512b.cu): 16 lanes x 16 points per frame,
// 256 = 16 x 16 four-step complex FFT in registers, mirror bin by warp shuffle, two frames per warp.
// New here: everything after the magnitudes runs on the tensor cores with warp-level
// mma.sync.m16n8k8 TF32, split 3-way (hi x hi + lo x hi + hi x lo, ~2^-21 relative) so the f32
// tolerances hold:
//   * a warp collects |X| of 8 consecutive frames (4 FFT pairs) in a private [8][132] tile and
//     multiplies it by the banded filterbank in n-tiles of 8 filters, visiting only the 8-bin k-steps
//     the band touches (22 of 85 for the 40-filter bank).  The M rows of the MMA are the 8 frames split
//     in place (rows 0..7 the TF32 high parts, rows 8..15 the remainders): two MMAs and one 16-byte
//     weight load per k-step.  The profile of the second generation had the per-pair CUDA-core mel
//     program at 23 % and the DCT at 11 % of all issued instructions;
//   * the accumulator fragment of the mel product (frame g, filters 2t, 2t+1) *is* the B fragment of
//     the transposed DCT product ceps^T = D x logmel^T once the contraction index is permuted (the
//     DCT matrix fragments are permuted on the host instead), so ln -> DCT needs no data movement;
//   * with the spectral subband centroids the same weights run against P[k] * R[k] (two more MMAs
//     per k-step);
//   * a clip's cepstra are normalised in place by the warp that wrote them (per-clip CMVN: column
//     statistics over the L2-resident [T][nc] block, no block-wide barrier) when every resident warp
//     has at least one clip.
// The fragment tables (12 KB mel + 5 KB DCT for the 40-filter bank) and the twiddles sit in shared
// memory once per SM (one TMA bulk copy per CTA); a warp's scratch is 8.9 KB (transpose tile 4.6 KB,
// |X| tile 4.2 KB): 198 KB for the 20 warps of the one persistent CTA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

#ifndef SS_F512C_WARPS
#define SS_F512C_WARPS 20
#endif
// One persistent CTA per SM: the warps are independent after the table load, and a single CTA
// needs only one shared-memory copy of the fragment tables (17 KB for the 40-filter bank).
constexpr int kMaxWarps = SS_F512C_WARPS, kMaxThreads = kMaxWarps * 32;
constexpr int kSmemBudget = 226 * 1024;       // of the 227 KB an sm_100 CTA may use
constexpr int kXs = 18;                       // transpose row stride in float2 (144 B: conflict-free LDS.128)
constexpr int kXchFrame = 16 * kXs;           // 288 float2 per frame
constexpr int kXchFloats = 2 * kXchFrame * 2; // 1152 floats per warp (2 frames)
constexpr int kTileRows = 8;                  // frames per MMA tile
constexpr int kPs = 148;                      // P row stride: bins 0..128 + 19 zero pads, so that the last k-step (bins 128..135) and the
                                              // zero-weight block of an odd band (bins 136..143) stay inside the row: a NaN / Inf frame
                                              // cannot reach its neighbours through 0 x Inf.  148 % 32 == 20 -> the 4-byte A-fragment
                                              // loads (row g, bin t) still hit 32 distinct banks
constexpr int kKs = 17;                       // k-steps of 8 bins
constexpr int kKsTab = kKs + 1;               // + one zero block: bands are padded to an even number of k-steps
constexpr int kWarpFloats = kXchFloats + kTileRows * kPs + kTileRows + 4;  // + 8 frame energies (+ 4: 16-byte granularity)
static_assert(kWarpFloats % 4 == 0, "per-warp scratch must stay 16-byte aligned");
constexpr int kTwRow = 36;                    // floats per lane row of the W256 table (16 float2 + pad)
constexpr int kTabTw = 16 * kTwRow;           // CTA-shared twiddle table
constexpr int kMbarFloats = 2 * 34;           // one 8-byte mbarrier per warp + one for the table copy, after the tables
constexpr int kNtMax = 8, kNcMax = 16;        // nf <= 64, nc <= 16

__device__ __forceinline__ float zero_handling(float v) { return v == 0.0f ? kEps : v; }
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// ln x = log2 x * ln 2 with the MUFU approximation (as __logf, minus its denormal-input fix-up:
// the arguments here are zero_handling()-ed mel energies)
__device__ __forceinline__ float fast_ln(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.693147180559945309f;
}
// x = hi + lo with hi = x truncated to TF32 (10 mantissa bits); lo is exact in f32 (|lo| < 2^-10 |x|)
// and the tensor core drops its low 13 mantissa bits, i.e. 2^-20 |x| in the worst case
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// one 4-byte shared-memory load that the compiler must not merge with its neighbour: the two P values of a
// lane go to slots 0 and 2 of the MMA's A register quad, and an 8-byte load would land them side by side
// (ptxas then spends four moves per k-step on rebuilding the quad)
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// D += A(16x8, row) * B(8x8, col), TF32 inputs, f32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
        "r"(b0), "r"(b1));
}

struct FastArgsC {
  FrameParams p;
  const float* tw;         // [16][kTwRow] W256^{u k1}
  const float* tab;        // fragment tables, copied to shared memory once per CTA:
                           //   mel [blocks][32] float4 {hi b0, hi b1, lo b0, lo b1}: B fragments of the filterbank
                           //   (8 bins x 8 filters), only the k-steps inside each 8-filter group's band, then
                           //   dct [NT][hi, lo][32] float4: A fragments of the DCT matrix (16 cepstra x 8 filters)
  int tab_floats;          // size of the two tables
  int dct_ofs;             // float offset of the dct table inside tab
  int mel_base[kNtMax];    // first block of n-tile nt inside the mel table
  int64_t n_units;         // clips * units_per_clip: the work is cut into units (one 8-frame tile, or the whole clip
  int units_per_clip;      // when the CMVN is fused) and every resident warp takes one contiguous, equally long
  int unit_frames;         // run of them (static and balanced per SM); a run is processed one clip segment at a time
  int fuse_cmvn;           // 0 none, 1 mean, 2 mean+var: one warp takes the whole clip and normalises it in place
  int vec2;                // every frame start is 8-byte aligned
  int stage;               // > 0: 16-byte chunks per staged pair (frame starts 16-byte aligned, S + L fits the buffer)
  int nf_even;             // nf is even: float2 stores of the mel rows
  int ks_lo[kNtMax], ks_n2[kNtMax];  // first k-step (8 bins each) and number of k-step pairs of the band, per n-tile of 8 filters
};

// JMAX: column groups (32 samples each) that can be non-zero; EXACT: JMAX == ceil(L / 32), so only
// the last group needs a bounds check; SSC: also the centroid numerator; NT: n-tiles of 8 filters
template <int JMAX, bool EXACT, bool SSC, int NT>
__global__ void __launch_bounds__(kMaxThreads, 1) k_fast512c(const FastArgsC a) {
  extern __shared__ __align__(16) float smem[];
  float* tw_s = smem;
  float* tab_s = smem + kTabTw;
  uint64_t* mbar_all = reinterpret_cast<uint64_t*>(tab_s + a.tab_floats);
  float* warp_s = tab_s + a.tab_floats + kMbarFloats + (threadIdx.x >> 5) * kWarpFloats;
  const FrameParams& p = a.p;
  // the tables (twiddles + fragments, contiguous in global and in shared memory) arrive by one TMA bulk copy
  const unsigned mbar_tab = (unsigned)__cvta_generic_to_shared(mbar_all + 32);
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(kTabTw + a.tab_floats) * 4u;
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar_tab) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_tab), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(tw_s)),
                 "l"(a.tw), "r"(bytes), "r"(mbar_tab)
                 : "memory");
  }
  // the tile's pad columns and the energies behind the last row (read by the zero-weight end of the last k-step) must be finite
  for (int i = (threadIdx.x & 31); i < kWarpFloats; i += 32) warp_s[i] = 0.0f;
  // the staged samples of the next pair arrive by TMA bulk copy and complete on this warp's mbarrier
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_all + (threadIdx.x >> 5));
  if ((threadIdx.x & 31) == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;  // parity of the mbarrier phase the next staged pair completes
  __syncthreads();  // the only block-wide barrier: from here on the warps are independent
  {
    unsigned done;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(mbar_tab)
          : "memory");
    } while (!done);
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = lane >> 4, u = lane & 15;   // FFT stage: frame of the pair, lane within the frame
  const int g = lane >> 2, t = lane & 3;     // MMA stage: fragment row group, thread in group
  const bool isz = (u == 0);
  const int partner = (lane & 16) | ((16 - u) & 15);
  const int nf = p.nf, nc = p.nc, L = p.L;
  const int T = (int)p.T;
  const float mscale = 0.5f * p.mag_scale;  // the 1/2 of E, O and 1/n_fft (processing.rs:180)

  const float4* twrow = reinterpret_cast<const float4*>(tw_s + u * kTwRow);
  // W512^{u + 16 k2} = W512^u * W32^{k2}: one per-lane register pair times a compile-time constant
  float2 wu;
  sincospif(-(float)u / 256.0f, &wu.y, &wu.x);

  float2* xch = reinterpret_cast<float2*>(warp_s) + fi * kXchFrame;
  float* stage_s = warp_s + kXchFloats / 2;   // landing buffer of the next pair's samples
  float* Pt = warp_s + kXchFloats;            // [kTileRows][kPs]
  float* e_s = Pt + kTileRows * kPs;          // [kTileRows] frame energies
  // does this lane's last column group (j = JMAX - 1) lie inside the frame?
  const bool last0 = 2 * u + 32 * (JMAX - 1) < L, last1 = 2 * u + 32 * (JMAX - 1) + 1 < L;
  const int n_warps = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * n_warps + warp, n_gw = (int64_t)gridDim.x * n_warps;
  const int64_t uend = (gw + 1) * a.n_units / n_gw;
  for (int64_t ubeg = gw * a.n_units / n_gw; ubeg < uend;) {
    const int64_t clip = ubeg / a.units_per_clip;
    const int u0 = (int)(ubeg - clip * a.units_per_clip);
    const int u1 = (int)min((int64_t)a.units_per_clip, u0 + (uend - ubeg));
    ubeg += u1 - u0;
    const float* rowp0 = p.pcm + clip * p.row_stride;
    const float* rowp = rowp0 + 2 * u;
    bool staged = false;  // the first pair of an item is loaded directly
    const int64_t obase = clip * T;  // first output row of the clip
    const int fbeg = u0 * a.unit_frames;
    const int fend = (int)min((int64_t)T, (int64_t)u1 * a.unit_frames);

#pragma unroll 1
    for (int tile_f0 = fbeg; tile_f0 < fend; tile_f0 += kTileRows) {
      const int npairs = min(kTileRows / 2, (fend - tile_f0 + 1) >> 1);
#pragma unroll 1
      for (int pr = 0; pr < npairs; ++pr) {
        const int fr = tile_f0 + 2 * pr + fi;
        const bool valid = fr < fend;
        const int row = 2 * pr + fi;  // tile rows in frame order

        // ---- framing: v[j] = x[2m] + i x[2m+1], m = u + 16 j; zero beyond L (processing.rs:147-156).
        // An out-of-range frame of the last pair re-reads frame T-1 (in bounds); it is never stored.
        // A staged pair was fetched by TMA bulk copy into the upper half of the transpose tile while
        // the previous pair was finishing (see below).
        float2 v[16];
        if (staged) {
          {
            unsigned done;
            do {
              asm volatile(
                  "{\n\t.reg .pred p;\n\t"
                  "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                  "selp.b32 %0, 1, 0, p;\n\t}"
                  : "=r"(done)
                  : "r"(mbar), "r"(phase)
                  : "memory");
            } while (!done);
            phase ^= 1u;
          }
          const float* sx = stage_s + fi * p.S + 2 * u;
          sfor<0, 16>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            if constexpr (j >= JMAX) {
              v[j] = make_float2(0.0f, 0.0f);
            } else if constexpr (!EXACT) {
              v[j].x = (2 * u + 32 * j < L) ? sx[32 * j] : 0.0f;
              v[j].y = (2 * u + 32 * j + 1 < L) ? sx[32 * j + 1] : 0.0f;
            } else if constexpr (j == JMAX - 1) {
              v[j].x = last0 ? sx[32 * j] : 0.0f;
              v[j].y = last1 ? sx[32 * j + 1] : 0.0f;
            } else {
              v[j] = *reinterpret_cast<const float2*>(sx + 32 * j);
            }
          });
          __syncwarp();  // the buffer is part of the transpose tile again
        } else {
          const float* px = rowp + (int64_t)(valid ? fr : T - 1) * p.S;
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
        float2 gg[16];
        {
          const float4* r4 = reinterpret_cast<const float4*>(xch + u * kXs);
          sfor<0, 8>([&](auto tt) {
            constexpr int i = decltype(tt)::value;
            const float4 x = r4[i];
            gg[2 * i] = make_float2(x.x, x.y);
            gg[2 * i + 1] = make_float2(x.z, x.w);
          });
        }
        __syncwarp();  // the transpose tile is free
        // ---- request the samples of the next pair (S + L contiguous floats) while this pair is
        // finished: one bulk copy into the upper half of the tile (the lower half holds the
        // log-mel tile during the epilogue).  Only whole, in-row pairs are staged.
        {
          const int nf0 = tile_f0 + 2 * pr + 2;  // first frame of the next pair
          staged = a.stage && nf0 + 1 < fend;
          if (staged && lane == 0) {
            // one TMA bulk copy (UBLKCP) per pair: no LSU wavefronts, no per-lane address arithmetic.  The tile
            // was last touched through the generic proxy (ordered by the __syncwarp above): proxy fence first.
            const float* src = rowp0 + (int64_t)nf0 * p.S;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stage_s);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(16 * a.stage) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src), "r"(16 * a.stage), "r"(mbar)
                         : "memory");
          }
        }
        // ---- stage 2: 16-point FFT over the lane index.  gg[k2] = Z[u + 16 k2]
        ssfft::fft_pow2<16>(gg);

        // ---- real-FFT split for k = u + 16 k2 <= 127 (k2 < 8): Z[256 - k] is gg[15 - k2] of lane
        // 16 - u (lane 0: its own gg[16 - k2]).  Each pair yields |X[k]| and |X[256 - k]|; both go
        // into the frame energy, only the bins a filter can reach (<= 128) into the tile.
        float* Prow = Pt + row * kPs;
        float* pA = Prow + u;
        float esum = 0.f;
        sfor<0, 8>([&](auto kk) {
          constexpr int k2 = decltype(kk)::value;
          float2 zb;
          zb.x = __shfl_sync(0xffffffffu, gg[15 - k2].x, partner);
          zb.y = __shfl_sync(0xffffffffu, gg[15 - k2].y, partner);
          if (isz) zb = gg[(16 - k2) & 15];
          const float2 za = gg[k2];
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
          const float m = 2.0f * fast_sqrt(fmaf(gg[8].x, gg[8].x, gg[8].y * gg[8].y));
          *reinterpret_cast<float4*>(Prow + 128) = make_float4(m, 0.f, 0.f, 0.f);
          esum += m;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
        if (isz) {
          const float energy = zero_handling(mscale * esum);  // feature.rs:216-219
          e_s[2 * pr + fi] = energy;
          if (p.energy && valid) p.energy[obase + fr] = energy;
        }
      }
      __syncwarp();

      // ---- mel filterbank on the tensor cores (feature.rs:229-230):
      //   mel [8 frames][8 filters of n-tile nt] = P [8][132] x W^T (banded, pre-split on the host).
      // The M rows of the MMA are the tile's 8 frames split in place: rows 0..7 their TF32 high parts, rows
      // 8..15 the remainders, so A = [P_hi; P_lo] against W_hi yields hi x hi and lo x hi at once and a second
      // MMA against W_lo adds hi x lo: two MMAs per k-step (8 bins x 8 filters), one 16-byte load of weights.
      // Only the k-steps an 8-filter group's band touches are visited (22 of 85 for the 40-filter bank).
      // Contraction slot t <-> bin 8 ks + t, slot t + 4 <-> bin 8 ks + 4 + t; a lane's two P values are two
      // 4-byte loads (conflict-free: bank 4 g + t with the 132-float row stride).  The accumulator rows g / g + 8 hold (frame g, filters 8 nt + 2t [+ 1]):
      // zero_handling -> stores -> ln follow on the fragment (feature.rs:216-230, 242-245), and the two
      // log-mel values are exactly this lane's B fragment of the DCT product below (contraction slot t <->
      // filter 8 nt + 2t, slot t + 4 <-> filter 8 nt + 2t + 1; the DCT matrix fragments are permuted on the
      // host), so ln -> DCT needs no data movement.
      // DCT-II (feature.rs:120-123), transposed as well:
      //   ceps^T [16][8 frames] = D [16][8 NT filters] (pre-split on the host) x logmel^T.
      // One m-tile covers all (<= 16) cepstra; the accumulator holds (cepstrum g [+ 8], frame 2t [+ 1]).
      float dB[4] = {0.f, 0.f, 0.f, 0.f}, dS[4] = {0.f, 0.f, 0.f, 0.f}, dT[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const int fr_g = tile_f0 + g;
        const bool vg = fr_g < fend;
        const int64_t og = (obase + fr_g) * nf;
        const unsigned prow_a = (unsigned)__cvta_generic_to_shared(Pt + g * kPs + t);
        const float4* dp = reinterpret_cast<const float4*>(tab_s + a.dct_ofs) + lane;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float accH[4] = {0.f, 0.f, 0.f, 0.f}, accL[4] = {0.f, 0.f, 0.f, 0.f};
          float numH[4] = {0.f, 0.f, 0.f, 0.f}, numL[4] = {0.f, 0.f, 0.f, 0.f};
          const int k0 = a.ks_lo[nt];
          unsigned pp = prow_a + 32 * k0;  // shared-memory byte address of P[g][8 k0 + t]
          const float4* bp = reinterpret_cast<const float4*>(tab_s) + a.mel_base[nt] * 32 + lane;
          float rb = SSC ? fmaf(p.ssc_step, (float)(8 * k0 + t), p.ssc_r0) : 0.f;  // R[k] of this lane's first bin
          // two k-steps per iteration (the band is padded to an even count with zero-weight blocks;
          // the tile columns those read are finite)
#pragma unroll 1
          for (int n2 = a.ks_n2[nt]; n2 > 0; --n2) {
            const float2 pv0 = make_float2(lds_f32(pp), lds_f32(pp + 16));        // bins 8 ks + t, 8 ks + 4 + t
            const float2 pv1 = make_float2(lds_f32(pp + 32), lds_f32(pp + 48));
            const float4 w0 = bp[0], w1 = bp[32];
            pp += 64;
            bp += 64;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float2 pv = h ? pv1 : pv0;
              const float4 w = h ? w1 : w0;
              uint32_t h0, l0, h2, l2;
              split_tf32(pv.x, h0, l0);
              split_tf32(pv.y, h2, l2);
              const float4 av = make_float4(__uint_as_float(h0), __uint_as_float(l0), __uint_as_float(h2), __uint_as_float(l2));
              mma_tf32(accH, av, __float_as_uint(w.x), __float_as_uint(w.y));  // [hi; lo] x hi
              mma_tf32(accL, av, __float_as_uint(w.z), __float_as_uint(w.w));  // [hi; lo] x lo
              if (SSC) {  // numerator: the same weights against P[k] R[k]
                split_tf32(pv.x * rb, h0, l0);
                split_tf32(pv.y * fmaf(4.0f, p.ssc_step, rb), h2, l2);
                rb = fmaf(8.0f, p.ssc_step, rb);
                const float4 nv = make_float4(__uint_as_float(h0), __uint_as_float(l0), __uint_as_float(h2), __uint_as_float(l2));
                mma_tf32(numH, nv, __float_as_uint(w.x), __float_as_uint(w.y));
                mma_tf32(numL, nv, __float_as_uint(w.z), __float_as_uint(w.w));
              }
            }
          }
          const int f0 = 8 * nt + 2 * t;
          const float m0 = zero_handling(accH[0] + (accH[2] + (accL[0] + accL[2])));
          const float m1 = zero_handling(accH[1] + (accH[3] + (accL[1] + accL[3])));
          const float lg0 = fast_ln(m0), lg1 = fast_ln(m1);
          if (vg && f0 < nf) {
            const int64_t o = og + f0;
            const float s0 = SSC ? __fdividef(numH[0] + (numH[2] + (numL[0] + numL[2])), m0) : 0.f;
            const float s1 = SSC ? __fdividef(numH[1] + (numH[3] + (numL[1] + numL[3])), m1) : 0.f;
            if (a.nf_even) {  // f0 is even: 8-byte aligned pairs
              if (p.mfe) *reinterpret_cast<float2*>(p.mfe + o) = make_float2(m0, m1);
              if (p.lmfe) *reinterpret_cast<float2*>(p.lmfe + o) = make_float2(lg0, lg1);
              if (SSC && p.ssc) *reinterpret_cast<float2*>(p.ssc + o) = make_float2(s0, s1);
            } else {
              const bool v1 = f0 + 1 < nf;
              if (p.mfe) { p.mfe[o] = m0; if (v1) p.mfe[o + 1] = m1; }
              if (p.lmfe) { p.lmfe[o] = lg0; if (v1) p.lmfe[o + 1] = lg1; }
              if (SSC && p.ssc) { p.ssc[o] = s0; if (v1) p.ssc[o + 1] = s1; }
            }
          }
          // ---- DCT-II on the tensor cores: the k-step of this n-tile's 8 filters
          if (p.mfcc) {
            const float4 wh = dp[(2 * nt) * 32], wl = dp[(2 * nt + 1) * 32];
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(lg0, bh0, bl0);
            split_tf32(lg1, bh1, bl1);
            mma_tf32(dB, wh, bh0, bh1);
            mma_tf32(dS, wl, bh0, bh1);
            mma_tf32(dT, wh, bl0, bl1);
          }
        }
      }

      if (p.mfcc) {
        // scale (feature.rs:126-131), c0 <- ln(energy) (feature.rs:137-146) and stores; with a
        // fused CMVN the block is normalised in place below
        const int fr_a = tile_f0 + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = g + 8 * (i >> 1), fr = fr_a + (i & 1);
          float x = dB[i] + (dS[i] + dT[i]);
          if (c == 0) {
            if (p.dc_elim) x = fast_ln(e_s[2 * t + (i & 1)]);
            else if (fr == 0) x *= p.scale_00;  // only element [0,0] gets the k=0 scale (feature.rs:128)
          } else {
            x *= p.scale_rest;
          }
          if (fr < fend && c < nc) p.mfcc[(obase + fr) * nc + c] = x;
        }
      }
      __syncwarp();  // the tile and the energies are rewritten by the next tile's pairs
    }

    // ---- per-clip CMVN (processing.rs:265-300): the warp re-reads the [T][nc] block it has just
    // written (L2-resident), lane = (row parity, column), 16 rows in flight per lane: column mean
    // and variance of the column (shifted by its first row: a one-sweep, cancellation-safe form of the
    // reference's mean-subtract-then-std), then normalise in place.
    if (a.fuse_cmvn) {
      __syncwarp();
      const int c = lane & 15, r0 = lane >> 4;
      const bool act = c < nc;
      float* col = p.mfcc + obase * nc + (act ? c : 0);
      const float K = __ldcg(col);  // row 0 of the column
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int rb = r0; rb < T; rb += 32) {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = (rb + 2 * j < T) ? __ldcg(col + (rb + 2 * j) * nc) : K;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float d = x[j] - K;
          s1 += d;
          s2 = fmaf(d, d, s2);
        }
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
      const float invT = 1.0f / (float)T;
      const float mean = fmaf(s1, invT, K);
      float inv = 1.0f;  // std with ddof 0, divide by (std + 2^-30) (processing.rs:283-292)
      if (a.fuse_cmvn == 2) inv = 1.0f / (sqrtf(fmaxf(s2 - s1 * s1 * invT, 0.f) * invT) + kCmvnEps);
      if (act) {
#pragma unroll 1
        for (int rb = r0; rb < T; rb += 32) {
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = (rb + 2 * j < T) ? __ldcg(col + (rb + 2 * j) * nc) : 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (rb + 2 * j < T) col[(rb + 2 * j) * nc] = (x[j] - mean) * inv;
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

struct FragTables {
  int NT = 0;
  int ks_lo[kNtMax] = {0}, ks_n2[kNtMax] = {0}, mel_base[kNtMax] = {0};
  std::vector<float> mel, dct;  // [blocks][32][4], [NT][2][32][4]
};

// A fragment of mma.m16n8k8 (row-major 16 x 8) for lane (g, t): rows g, g + 8; contraction slots t, t + 4.
// `at(row, slot)` returns the matrix element; the fragment is stored pre-split as a TF32 hi quad and
// the f32 remainder quad.
template <class F>
static void put_a_frag(float* hi, float* lo, int lane, F at) {
  const int g = lane >> 2, t = lane & 3;
  const float v[4] = {at(g, t), at(g + 8, t), at(g, t + 4), at(g + 8, t + 4)};
  for (int i = 0; i < 4; ++i) {
    hi[i] = tf32_rna(v[i]);
    lo[i] = v[i] - hi[i];
  }
}

static bool build_frags(const Derived& d, const MelBank& mb, FragTables* ft) {
  const int nf = d.nf;
  if (mb.max_bin > 128) return false;  // only bins <= 128 are kept (always true for the reference's bank at n_fft 512)
  if (nf > 8 * kNtMax || d.nc > kNcMax || nf < 1) return false;
  const int NT = nf <= 40 ? 5 : 8;
  ft->NT = NT;
  const float mscale = 0.5f * (1.0f / (float)d.W);  // the kernel keeps 2 n_fft |X| in the tile
  const int ncol = 8 * kKsTab;
  std::vector<float> dense((size_t)8 * NT * ncol, 0.f);  // [filter][bin]
  for (int f = 0; f < nf; ++f)
    for (int i = 0; i < mb.len[f]; ++i) dense[(size_t)f * ncol + mb.start[f] + i] = mb.w[mb.wofs[f] + i] * mscale;
  ft->mel.clear();
  for (int nt = 0; nt < kNtMax; ++nt) { ft->ks_lo[nt] = 0; ft->ks_n2[nt] = 0; ft->mel_base[nt] = 0; }
  int n_blocks = 0;
  for (int nt = 0; nt < NT; ++nt) {
    int lo = kKs, hi = 0;
    for (int ks = 0; ks < kKs; ++ks) {
      bool any = false;
      for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) any = any || dense[(size_t)(8 * nt + r) * ncol + 8 * ks + c] != 0.f;
      if (any) { lo = std::min(lo, ks); hi = std::max(hi, ks + 1); }
    }
    if (hi <= lo) continue;  // 8 empty filters: no k-steps
    const int n2 = (hi - lo + 1) / 2;  // k-step pairs; an odd band gets one zero-weight block (ks <= 17 < kKsTab)
    ft->ks_lo[nt] = lo;
    ft->ks_n2[nt] = n2;
    ft->mel_base[nt] = n_blocks;
    for (int ks = lo; ks < lo + 2 * n2; ++ks, ++n_blocks) {
      // B fragment of mma.m16n8k8 (col-major 8 bins x 8 filters) for lane (g, t): filter 8 nt + g, contraction
      // slots t, t + 4 <-> bins 8 ks + t, 8 ks + 4 + t;
      // stored pre-split as {hi b0, hi b1, lo b0, lo b1}
      const size_t o = ft->mel.size();
      ft->mel.resize(o + 32 * 4, 0.f);
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        for (int i = 0; i < 2; ++i) {
          const float v = dense[(size_t)(8 * nt + g) * ncol + 8 * ks + t + 4 * i];
          const float h = tf32_rna(v);
          ft->mel[o + lane * 4 + i] = h;
          ft->mel[o + lane * 4 + 2 + i] = v - h;
        }
      }
    }
  }
  std::vector<float> dct;
  build_dct(d, &dct);  // [nc][nf]
  ft->dct.assign((size_t)NT * 2 * 32 * 4, 0.f);
  for (int nt = 0; nt < NT; ++nt)
    for (int lane = 0; lane < 32; ++lane) {
      const size_t o = (((size_t)nt * 2) * 32 + lane) * 4;
      put_a_frag(&ft->dct[o], &ft->dct[o + 32 * 4], lane, [&](int r, int s) {
        const int f = 8 * nt + 2 * (s & 3) + (s >> 2);
        return (r < d.nc && f < nf) ? dct[(size_t)r * nf + f] : 0.f;
      });
    }
  return true;
}

}  // namespace

struct Fast512cTables {
  float* d_tab = nullptr;  // tw | mel fragments | dct fragments
  int tab_floats = 0, dct_ofs = 0;
  int NT = 0;
  int ks_lo[kNtMax] = {0}, ks_n2[kNtMax] = {0}, mel_base[kNtMax] = {0};
};

bool fast512c_supported(const Derived& d, const MelBank& mb) {
  if (d.W != 512 || !d.speechpy_ok) return false;
  FragTables ft;
  return build_frags(d, mb, &ft);
}

int fast512c_build_tables(const Derived& d, const MelBank& mb, Fast512cTables** out) {
  FragTables ft;
  if (!build_frags(d, mb, &ft)) return (int)cudaErrorInvalidValue;
  std::vector<float> h(kTabTw, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  for (int u = 0; u < 16; ++u)
    for (int k1 = 0; k1 < 16; ++k1) {
      const double ang = -2.0 * pi * (double)(u * k1) / 256.0;
      h[u * kTwRow + 2 * k1] = (float)std::cos(ang);
      h[u * kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
    }
  Fast512cTables* t = new Fast512cTables;
  t->NT = ft.NT;
  for (int i = 0; i < kNtMax; ++i) { t->ks_lo[i] = ft.ks_lo[i]; t->ks_n2[i] = ft.ks_n2[i]; t->mel_base[i] = ft.mel_base[i]; }
  h.insert(h.end(), ft.mel.begin(), ft.mel.end());
  t->dct_ofs = (int)ft.mel.size();
  h.insert(h.end(), ft.dct.begin(), ft.dct.end());
  t->tab_floats = (int)(ft.mel.size() + ft.dct.size());
  cudaError_t e = cudaMalloc(&t->d_tab, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return (int)e; }
  *out = t;
  return 0;
}

void fast512c_free_tables(Fast512cTables* t) {
  if (!t) return;
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

// returns 0 (done, including any requested cmvn), -1000 (shape declined), -1001 (features
// written, cmvn left to the caller) or a cudaError_t
int launch_fast512c(const FrameParams& p, const Fast512cTables* t, int sm_count, void* stream) {
  if (p.N != 512 || p.L > 512 || p.T < 1 || p.T > (1LL << 31) - 64) return -1000;
  FastArgsC a;
  a.p = p;
  a.tw = t->d_tab;
  a.tab = t->d_tab + kTabTw;
  a.tab_floats = t->tab_floats;
  a.dct_ofs = t->dct_ofs;
  for (int i = 0; i < kNtMax; ++i) { a.ks_lo[i] = t->ks_lo[i]; a.ks_n2[i] = t->ks_n2[i]; a.mel_base[i] = t->mel_base[i]; }
  // as many warps as the shared memory left by the tables allows (21 for the 40-filter bank)
  const size_t fixed = (size_t)(kTabTw + t->tab_floats + kMbarFloats) * sizeof(float);
  const int n_warps = (int)std::min<size_t>(kMaxWarps, (kSmemBudget - fixed) / (kWarpFloats * sizeof(float)));
  if (n_warps < 4) return -1000;
  const int64_t warps_resident = (int64_t)sm_count * n_warps;
  // a fused CMVN needs the whole clip in one warp; worth it while there are enough clips to fill the GPU
  const bool fuse = p.mfcc && p.cmvn_mode && p.T <= 4096 && p.n_rows >= warps_resident;
  a.fuse_cmvn = fuse ? p.cmvn_mode : 0;
  // latency case: fewer tiles than SMs -> single pairs as units, so that more SMs share the few frames
  const int uf = p.n_rows * ((p.T + kTileRows - 1) / kTileRows) < sm_count ? 2 : kTileRows;
  a.unit_frames = fuse ? (int)p.T : uf;
  a.units_per_clip = fuse ? 1 : (int)((p.T + uf - 1) / uf);
  a.n_units = p.n_rows * a.units_per_clip;
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  // TMA bulk-copy staging of whole pairs: 16-byte aligned frame starts, the pair's S + L samples (rounded up to
  // 16 bytes) fit the upper half of the transpose tile and stay inside the row for every staged pair
  {
    const int chunks = (p.S + p.L + 3) / 4;
    const bool fits = 4 * chunks <= kXchFloats / 2 && (p.T - 2) * (int64_t)p.S + 4 * chunks <= p.row_len;
    const bool aligned = (p.S % 4 == 0) && (p.row_stride % 4 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 16 == 0);
    static const bool no_stage = getenv("SS_F512C_STAGE") && atoi(getenv("SS_F512C_STAGE")) == 0;  // A/B switch
    a.stage = (fits && aligned && p.T >= 2 && !no_stage) ? chunks : 0;
  }
  auto al8 = [](const float* q) { return reinterpret_cast<uintptr_t>(q) % 8 == 0; };
  a.nf_even = (p.nf % 2 == 0) && al8(p.mfe) && al8(p.lmfe) && al8(p.ssc);
  const bool ssc = p.ssc != nullptr;
  // small jobs: fewer warps per CTA so that the items spread over the SMs instead of queueing on a few
  const int wl = (int)std::min<int64_t>(n_warps, std::max<int64_t>(1, (a.n_units + sm_count - 1) / sm_count));
  const size_t smem = fixed + (size_t)wl * kWarpFloats * sizeof(float);
  const int64_t grid = std::min<int64_t>((a.n_units + wl - 1) / wl, (int64_t)sm_count);
  if (grid <= 0) return 0;
  const int jn = (p.L + 31) / 32;
  cudaError_t e;
#define SS_LAUNCH_FASTC(J, X, C, N)                                                                            \
  do {                                                                                                         \
    e = cudaFuncSetAttribute(k_fast512c<J, X, C, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e != cudaSuccess) return (int)e;                                                                       \
    k_fast512c<J, X, C, N><<<(unsigned)grid, wl * 32, smem, (cudaStream_t)stream>>>(a);                       \
  } while (0)
#define SS_LAUNCH_FASTC_N(J, X, C)                                          \
  do {                                                                      \
    if (t->NT == 5) SS_LAUNCH_FASTC(J, X, C, 5); else SS_LAUNCH_FASTC(J, X, C, 8); \
  } while (0)
  if (jn == 10) { if (ssc) SS_LAUNCH_FASTC_N(10, true, true); else SS_LAUNCH_FASTC_N(10, true, false); }        // L = 289..320
  else if (jn == 13) { if (ssc) SS_LAUNCH_FASTC_N(13, true, true); else SS_LAUNCH_FASTC_N(13, true, false); }   // L = 385..416
  else { if (ssc) SS_LAUNCH_FASTC_N(16, false, true); else SS_LAUNCH_FASTC_N(16, false, false); }
#undef SS_LAUNCH_FASTC_N
#undef SS_LAUNCH_FASTC
  count_launch();
  note_kernel("k_fast512c");
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return (p.cmvn_mode && !fuse) ? -1001 : 0;
}

}  // namespace ss
