// Fused mel-spectrogram kernel for the STFT path with fft_points = 400, second generation
// (config c4, the Whisper-shaped log-mel: 16 kHz, hop 160, 80 mels):
//   stream framing (functions.rs:86-160) -> Vorbis window -> real FFT(400) -> |X wnorm|^2 ->
//   mel filterbank (feature.rs:163-174) -> optional ln(zero_handling) -> time-minor store.
// Replaces feature::mel_spectrogram{1,2} and functions::stft{1,2} (complex output, written straight
// from the registers of the real-FFT split) for this FFT size.
//
// FFT stage as in the first generation (ss_kernels_stft400.cu): 10 lanes x 20 complex points per
// frame, 200 = 20 x 10 four-step FFT in registers (radix 5 / 2, packed FFMA2 butterflies), three
// frames per warp.  New here, following k_fast512c:
//   * the warps are independent (one persistent CTA per SM, no block-wide barrier after the table
//     load): an item is a run of frames of one row; a warp works through it three frames per FFT
//     round, the |X|^2 rows go to a private 12-row ring in shared memory and every time 8 more
//     frames are complete the ring feeds one 8-frame MMA tile;
//   * the samples of the next round are fetched by TMA bulk copies (one per frame) into the
//     transpose tile while the current round finishes, so the global-load latency is off the
//     critical path and the window multiply reads shared memory;
//   * the filterbank runs on the tensor cores, transposed: mel^T [16 MT filters][8 frames] =
//     W [16 MT][104 bins] x P^T with mma.sync.m16n8k8 TF32 split 3-way (hi x hi + lo x hi + hi x lo,
//     ~2^-21 relative), visiting only the 8-bin k-steps a 16-filter group's band touches; the
//     pre-split weight fragments live in shared memory, one copy per SM;
//   * the accumulator fragment holds (filter g [+ 8], frames 2t, 2t + 1), i.e. two neighbouring
//     elements of the time-minor output [row][m][t] (feature.rs:173): they are stored straight from
//     registers as 8-byte pairs, no staging tile and no __syncthreads.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

constexpr int kW = 400, kM = 200, kT = 10, kR1 = 20;
constexpr int kFpw = 3;                  // frames per warp and FFT round
constexpr int kTileFrames = 8;           // frames per MMA tile
constexpr int kRing = 12;                // rows of the |X|^2 ring: one tile plus the 2 frames a round may run ahead, rounded up
constexpr int kChunkAlign = 24;          // items start at multiples of 24 frames (whole tiles, 96-byte aligned output runs)
#ifndef SS_STFT400M_WARPS
#define SS_STFT400M_WARPS 20
#endif
constexpr int kMaxWarps = SS_STFT400M_WARPS, kMaxThreads = kMaxWarps * 32;
constexpr int kSmemBudget = 226 * 1024;
constexpr int kXchFrame = 202;           // float2 per frame: 20 rows x 10 + 2 (frame stride = 10 mod 16 banks)
constexpr int kXchFloats = kFpw * kXchFrame * 2;  // 1212
// |X|^2 tile geometry.  Reference semantics: the bank reaches bins 0..100 only (quirk Q8) -> 104 columns;
// librosa plans (SS_FLAG_LIBROSA): the Slaney bank spans all 201 bins -> 208 columns.  Both strides are
// 8 mod 32 words, so the LDS.64 B-fragment loads stay conflict-free.
template <bool FULL>
struct Geo {
  static constexpr int kPs = FULL ? 208 : 104;     // P row stride
  static constexpr int kKs = kPs / 8;              // k-steps of 8 bins
  static constexpr int kTileFloats = kRing * kPs + 8;  // + slack for the padded k-step of the last row
  static constexpr int kWarpFloats = kFpw * 202 * 2 + kTileFloats;
};
constexpr int kStage = kW + (kFpw - 1) * 160;    // samples of one round when hop = 160 (staging needs hop == 160)
// The staged samples of frame fi land in that frame's slot of the transpose tile (404 floats apart, i.e. 20 banks:
// the LDS.64 of the three frames do not collide).
static_assert(kW <= 2 * kXchFrame, "a staged frame fits its transpose slot");
// Physical row of logical row k1 in a frame's transpose tile.  Lane u reads rows u and 20 - u (10 for u = 0)
// with LDS.128; this placement makes the first conflict-free across the three frames of a warp (4 wavefronts)
// and leaves the second at 6 (natural order: 7 and 10).  Found by exhaustive search over 16-byte bank groups.
__device__ constexpr int kTwSlot[10] = {0, 1, 4, 7, 3, 2, 6, 5, 8, 9};
__device__ constexpr int kRowOf[20] = {0, 5, 4, 3, 1, 6, 2, 7, 8, 11, 16, 18, 17, 15, 14, 13, 12, 10, 9, 19};
static_assert(Geo<false>::kWarpFloats % 4 == 0 && Geo<true>::kWarpFloats % 4 == 0, "per-warp scratch must stay 16-byte aligned");
static_assert(Geo<false>::kWarpFloats == kXchFloats + Geo<false>::kTileFloats, "transpose tile size");
constexpr int kMtMax = 8;                // nf <= 128
constexpr int kMbarFloats = 2 * 34;      // one 8-byte mbarrier per warp + one for the table copy, at the start of the dynamic shared memory
constexpr int kTwRow = 44;               // floats per lane row of the W200 table (20 float2 + pad)
// fixed table blob (floats)
// W200 table: one copy per frame of the round, lane u's row in slot kTwSlot[u], the copies 452 floats apart:
// the 30 lanes' LDS.128 then fall into distinct 16-byte bank groups in every quarter-warp (4 wavefronts per
// load instead of 7 with a single row-ordered copy; exhaustive search like kRowOf)
constexpr int kTwCopy = 452;
constexpr int kTabTw = 0;                                          // [3][10 slots][kTwRow]
constexpr int kTabTw400 = kTabTw + 2 * kTwCopy + kT * kTwRow;      // [200] float2 W400^k (entries 100 + 20 m: W400^(10 + 20 m))
constexpr int kTabWin = kTabTw400 + 2 * kM;       // [200] float2 window pairs
constexpr int kTabFixed = kTabWin + 2 * kM;       // then the mel A fragments [blocks][hi, lo][32] float4
static_assert(kTabTw400 % 4 == 0 && kTabWin % 4 == 0 && kTabFixed % 4 == 0, "table alignment");

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

struct Stft400mArgs {
  FrameParams p;
  const float* tab;     // fixed tables + mel fragments, copied to shared memory once per CTA
  int tab_floats;
  int64_t n_units;      // rows * units_per_row: the work is cut into units of kChunkAlign frames and every
  int units_per_row;    // resident warp takes one contiguous, equally long run of units (static, balanced)
  int vec2;             // frames start 8-byte aligned
  int stage;            // hop == 160 and frames start 16-byte aligned: bulk-copy staging of whole rounds
  int pair_ok;          // T_out is even and the output is 8-byte aligned: float2 stores
  int n_mt;             // m-tiles of 16 filters
  int ks_lo[kMtMax], ks_n2[kMtMax], mel_base[kMtMax];
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

// E' = za + conj zb, O' = (za - conj zb)/i, T' = W400^k O': xa = E' + T' = 2 X[k], xb = E' - T' = 2 conj X[200-k]
__device__ __forceinline__ void pair_split(float2 za, float2 zb, float2 w, float2& xa, float2& xb) {
  const float2 e = __fadd2_rn(za, make_float2(zb.x, -zb.y));
  const float2 od = __fadd2_rn(make_float2(za.y, -za.x), make_float2(zb.y, zb.x));
  const float2 tt = ssfft::cmulf(od, w);
  xa = ssfft::cadd(e, tt);
  xb = ssfft::csub(e, tt);
}

// STFT: also / only the complex output of functions::stft1/2.  FULL: librosa plan -- centred frames with
// reflect padding, every one of the 201 bins kept for the full-spectrum Slaney bank, optional power_to_db.
template <bool STFT, bool FULL>
__global__ void __launch_bounds__(kMaxThreads, 1) k_stft400m(const Stft400mArgs a) {
  constexpr int kPs = Geo<FULL>::kPs, kWarpFloats = Geo<FULL>::kWarpFloats;
  extern __shared__ __align__(16) float smem[];
  // layout: mbarriers, the warps' scratch, then the tables (whose size depends on the bank): the per-warp addresses are
  // compile-time offsets from the warp index and need no register
  uint64_t* mbar_all = reinterpret_cast<uint64_t*>(smem);  // one mbarrier per warp
  float* warp_s = smem + kMbarFloats + (threadIdx.x >> 5) * kWarpFloats;
  float* tab_s = smem + kMbarFloats + (blockDim.x >> 5) * kWarpFloats;
  const FrameParams& p = a.p;
  const int nf = p.nf;
  // the tables arrive by one TMA bulk copy
  const unsigned mbar_tab = (unsigned)__cvta_generic_to_shared(mbar_all + 32);
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)a.tab_floats * 4u;
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar_tab) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_tab), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(tab_s)),
                 "l"(a.tab), "r"(bytes), "r"(mbar_tab)
                 : "memory");
  }
  // the pad bins of the ring rows are never written again and must be finite
  for (int i = (threadIdx.x & 31); i < kWarpFloats; i += 32) warp_s[i] = 0.0f;
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_all + (threadIdx.x >> 5));
  if ((threadIdx.x & 31) == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;  // parity of the mbarrier phase the next staged round completes
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
  const int sl = lane < 30 ? lane : lane - 2;   // lanes 30, 31 shadow lanes 28, 29 (same addresses: broadcasts)
  const int fi = sl / 10, u = sl - fi * 10;     // FFT stage: frame of the round, lane within the frame
  const int g = lane >> 2, t = lane & 3;        // MMA stage: fragment row group, thread in group
  const bool real_lane = lane < 30;
  const bool isz = (u == 0);
  const int k1B = isz ? 10 : 20 - u;
  const int rowA = kRowOf[u], rowB = kRowOf[k1B];
  const int kb = isz ? -90 : u;  // k of slots 5..9 is kb + 20 e (u == 0: 10 + 20 (e - 5))

  const float4* twrow = reinterpret_cast<const float4*>(tab_s + kTabTw + fi * kTwCopy + kTwSlot[u] * kTwRow);
  const float2* tw400 = reinterpret_cast<const float2*>(tab_s + kTabTw400);
  const float2* win2 = reinterpret_cast<const float2*>(tab_s + kTabWin) + u;
  const float4* frag_s = reinterpret_cast<const float4*>(tab_s + kTabFixed) + lane;

  float2* xch = reinterpret_cast<float2*>(warp_s) + fi * kXchFrame;
  float* Pt = warp_s + kXchFloats;  // [8][kPs]
  // slot e of lane u pairs bins k = u + 20 e and 200 - k and needs W400^k; lane 0's slots 5..9 pair
  // k = 10 + 20 (e - 5) instead, whose twiddles sit in the table's otherwise unused entries 20 e
  const float2* twA = tw400 + u;
  int plo[5];  // column of the low bin of slots 5..9
#pragma unroll
  for (int i = 0; i < 5; ++i) plo[i] = isz ? 10 + 20 * i : 200 - u - 20 * (i + 5);

  const int64_t span = p.Tp * (int64_t)p.S;
  const int64_t memlen = p.N - p.S;
  // row-relative position of frame 0's first sample (<= 0): the stream model of the reference, or the centre
  // of frame 0 on sample 0 (librosa.stft(center=True))
  const int c0 = FULL ? -kW / 2 : (p.n_pad + 1) * p.S - p.N;
  const int T_out = (int)p.T_out, Tc = (int)p.T;
  const int n_warps = blockDim.x >> 5;

  // this warp's run of units; a run that crosses row boundaries is processed one row segment (item) at a time
  const int64_t gw = (int64_t)blockIdx.x * n_warps + warp, n_gw = (int64_t)gridDim.x * n_warps;
  const int64_t uend = (gw + 1) * a.n_units / n_gw;
  for (int64_t ubeg = gw * a.n_units / n_gw; ubeg < uend;) {
    const int64_t row = ubeg / a.units_per_row;
    const int u0 = (int)(ubeg - row * a.units_per_row);
    const int u1 = (int)min((int64_t)a.units_per_row, u0 + (uend - ubeg));
    ubeg += u1 - u0;
    const int fbeg = u0 * kChunkAlign;
    const int fend = min(T_out, u1 * kChunkAlign);
    const float* rowp = p.pcm + row * p.row_stride;
    float* outp = p.mel_out ? p.mel_out + row * nf * (int64_t)T_out : nullptr;

    const int nfr = fend - fbeg;                    // frames of this item
    const int n_rounds = (nfr + kFpw - 1) / kFpw;
    float* stage_s = warp_s;                        // the transpose tile doubles as the landing buffer of the next round
    // a round can be staged when its three frames are whole, in-row and computed
    auto stageable = [&](int rd) {
      const int f0 = fbeg + kFpw * rd;
      const int64_t o0 = (int64_t)f0 * p.S + c0;
      return a.stage && f0 + kFpw <= Tc && o0 >= 0 && o0 + kStage <= p.row_len;
    };
    // TMA bulk copies (UBLKCP) issued by one lane: no LSU wavefronts, no per-lane address arithmetic.  The
    // tile was last touched through the generic proxy (ordered before this point by __syncwarp), hence the
    // proxy fence before the asynchronous proxy writes it.
    auto issue_stage = [&](int rd) {
      if (lane == 0) {
        const float* src = rowp + (int64_t)(fbeg + kFpw * rd) * p.S + c0;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(stage_s);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // one copy per frame (400 samples each; the overlaps come twice, out of L2) into the frames' own 404-float slots:
        // three UBLKCP per round instead of five 160-sample segments with padding (each costs ~14 issue slots of
        // uniform-register setup), and the window pass reads plain frame offsets
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(kFpw * kW * 4) : "memory");
#pragma unroll
        for (int sg = 0; sg < kFpw; ++sg)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           dst + sg * 2 * kXchFrame * 4),
                       "l"(src + sg * 160), "r"(kW * 4), "r"(mbar)
                       : "memory");
      }
    };
    bool staged = n_rounds > 0 && stageable(0);
    if (staged) issue_stage(0);
    int tile = 0;  // next 8-frame tile to run through the filterbank

#pragma unroll 1
    for (int rd = 0; rd < n_rounds; ++rd) {
      {
        const int lf = kFpw * rd + fi;      // frame within the item
        const int tf = fbeg + lf;           // this lane group's output frame
        // frames in [T, T_out) are the reference's zero tail; frames >= fend are never stored
        const bool computed = tf < Tc;
        float* Pf = Pt + (lf % kRing) * kPs;

        // ---- framing + window: v[j] = (x w)[2m] + i (x w)[2m+1], m = u + 10 j
        float2 v[kR1];
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
          const float* sx = stage_s + fi * 2 * kXchFrame + 2 * u;
          sfor<0, kR1>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            v[j] = __fmul2_rn(*reinterpret_cast<const float2*>(sx + 20 * j), win2[10 * j]);
          });
          __syncwarp();  // the buffer is the transpose tile again
        } else {
          const int64_t o0 = (int64_t)tf * p.S + c0;  // first sample of the frame, relative to the row start
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
              float xa, xb;
              if (FULL) {  // pad_mode="reflect"
                const int64_t n = p.row_len;
                int64_t ia = o < 0 ? -o : o, ib = o + 1 < 0 ? -(o + 1) : o + 1;
                if (ia >= n) ia = 2 * (n - 1) - ia;
                if (ib >= n) ib = 2 * (n - 1) - ib;
                xa = rowp[ia];
                xb = rowp[ib];
              } else {
                xa = o >= 0 ? (o < p.row_len ? rowp[o] : 0.f) : stream_at(p, q);
                xb = o + 1 >= 0 ? (o + 1 < p.row_len ? rowp[o + 1] : 0.f) : stream_at(p, q + 1);
              }
              v[j] = make_float2(xa * w.x, xb * w.y);
            });
          }
        }
        // ---- stage 1: 20-point FFT over j, in registers
        ssfft::fft_5x2<kR1>(v);
        // ---- twiddle W200^{u k1} and transpose
        sfor<0, kR1 / 2>([&](auto kk) {
          constexpr int k1 = 2 * decltype(kk)::value;
          const float4 w = twrow[k1 / 2];
          float2 y0 = v[k1], y1 = v[k1 + 1];
          if constexpr (k1 > 0) y0 = ssfft::cmulf(y0, make_float2(w.x, w.y));
          y1 = ssfft::cmulf(y1, make_float2(w.z, w.w));
          xch[kRowOf[k1] * kT + u] = y0;
          xch[kRowOf[k1 + 1] * kT + u] = y1;
        });
        __syncwarp();
        float2 gA[kT], gB[kT];
        {
          const float4* rA = reinterpret_cast<const float4*>(xch + rowA * kT);
          const float4* rB = reinterpret_cast<const float4*>(xch + rowB * kT);
          sfor<0, kT / 2>([&](auto tt) {
            constexpr int i = decltype(tt)::value;
            const float4 xa = rA[i], xb = rB[i];
            gA[2 * i] = make_float2(xa.x, xa.y); gA[2 * i + 1] = make_float2(xa.z, xa.w);
            gB[2 * i] = make_float2(xb.x, xb.y); gB[2 * i + 1] = make_float2(xb.z, xb.w);
          });
        }
        __syncwarp();  // the transpose tile is free: it receives the samples of the next round
        staged = rd + 1 < n_rounds && stageable(rd + 1);
        if (staged) issue_stage(rd + 1);
        // ---- stage 2: 10-point FFTs over the lane index.  g?[k2] = Z[k1? + 20 k2]
        ssfft::fft_5x2<kT>(gA);
        ssfft::fft_5x2<kT>(gB);

        // ---- real-FFT post-processing, |X|^2 of the bins a mel filter can reach (<= 100).
        // u != 0: slot e pairs Z[u+20e] (gA[e]) with Z[200-(u+20e)] = gB[9-e].
        // u == 0: groups 0 and 10 mirror onto themselves: e=0 -> k=0; e=1..4 -> (gA[e], gA[10-e]);
        //         e=5..9 -> (gB[e-5], gB[14-e]) with k = 10 + 20 (e-5); bin 100 = |gA[5]|^2 below.
        if (real_lane) {
          float* pA = Pf + u;
          // complex output of functions::stft1/2 (functions.rs:86-123,199-233): [row][T_out][201] (re, im),
          // X wnorm for computed frames, zeros in the reference's zero tail (v = 0 there)
          float2* so = (STFT && p.stft_out && tf < fend) ? p.stft_out + (row * (int64_t)T_out + tf) * (kM + 1) : nullptr;
          const float hs = 0.5f * p.wnorm;
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
            float2 xa, xb;
            pair_split(za, zb, twA[20 * e], xa, xb);
            const float2 m = make_float2(fmaf(xa.x, xa.x, xa.y * xa.y), fmaf(xb.x, xb.x, xb.y * xb.y));
            if constexpr (FULL) {  // both bins of the pair: k and 200 - k
              const int k = (e < 5 ? u : kb) + 20 * e;
              Pf[k] = m.x;
              Pf[kM - k] = m.y;
            } else if constexpr (e < 5) {
              pA[20 * e] = m.x;                           // bin u + 20 e <= 89
            } else {
              Pf[plo[e - 5]] = isz ? m.x : m.y;           // the bin of the pair that is <= 100
            }
            if (STFT && so) {
              const int k = (e < 5 ? u : kb) + 20 * e;    // xa is bin k, xb the conjugate of bin 200 - k
              so[k] = make_float2(xa.x * hs, xa.y * hs);
              so[kM - k] = make_float2(xb.x * hs, -xb.y * hs);
            }
          });
          if (isz) {  // X[100] = conj Z[100]; x4 to match the pairs; the pad bins stay zero
            Pf[100] = 4.0f * fmaf(gA[5].x, gA[5].x, gA[5].y * gA[5].y);
            if (STFT && so) so[100] = make_float2(gA[5].x * p.wnorm, -gA[5].y * p.wnorm);
          }
        }
      }
      if (STFT && !p.mel_out) continue;  // stft only
      // a tile is complete once 8 more frames exist (or the item ends)
      const int produced = min(nfr, kFpw * (rd + 1));
      if (produced < kTileFrames * (tile + 1) && rd + 1 < n_rounds) continue;
      __syncwarp();

      // ---- mel filterbank on the tensor cores (feature.rs:163-174), transposed:
      //   mel^T [16 MT filters][8 frames] = W [16 MT][104 bins] (banded, pre-split on the host) x P^T,
      // contraction slot t <-> bin 8 ks + 2t, slot t + 4 <-> bin 8 ks + 2t + 1 (one 8-byte load).
      // The accumulator holds (filter 16 mt + g [+ 8], frames tile_f0 + 2t [+ 1]): pairs along the
      // time-minor output [row][m][t].
#pragma unroll 1
      for (; kTileFrames * tile < produced && (kTileFrames * (tile + 1) <= produced || rd + 1 == n_rounds); ++tile) {
        const int fa = fbeg + kTileFrames * tile + 2 * t;  // first frame of this lane's pair
        const bool va = fa < fend, vb = fa + 1 < fend;
        const float* prow = Pt + ((kTileFrames * tile + g) % kRing) * kPs + 2 * t;
#pragma unroll 1
        for (int mt = 0; mt < a.n_mt; ++mt) {
          float accB[4] = {0.f, 0.f, 0.f, 0.f}, accS[4] = {0.f, 0.f, 0.f, 0.f}, accT[4] = {0.f, 0.f, 0.f, 0.f};
          const float* pp = prow + 8 * a.ks_lo[mt];
          const float4* ap = frag_s + a.mel_base[mt] * 64;
          // two k-steps per iteration (the band is padded to an even count with zero-weight blocks)
#pragma unroll 1
          for (int n2 = a.ks_n2[mt]; n2 > 0; --n2) {
            const float2 pv0 = *reinterpret_cast<const float2*>(pp);
            const float2 pv1 = *reinterpret_cast<const float2*>(pp + 8);
            const float4 h0 = ap[0], l0 = ap[32], h1 = ap[64], l1 = ap[96];
            pp += 16;
            ap += 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float2 pv = h ? pv1 : pv0;
              const float4 wh = h ? h1 : h0, wl = h ? l1 : l0;
              uint32_t bh0, bl0, bh1, bl1;
              split_tf32(pv.x, bh0, bl0);
              split_tf32(pv.y, bh1, bl1);
              mma_tf32(accB, wh, bh0, bh1);  // hi x hi
              mma_tf32(accS, wl, bh0, bh1);  // lo x hi
              mma_tf32(accT, wh, bl0, bl1);  // hi x lo
            }
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int f = 16 * mt + g + 8 * hh;
            float m0 = accB[2 * hh] + (accS[2 * hh] + accT[2 * hh]);
            float m1 = accB[2 * hh + 1] + (accS[2 * hh + 1] + accT[2 * hh + 1]);
            if (FULL && p.log_mode == 2) {  // librosa.power_to_db(ref=1, amin=1e-10, top_db=None)
              m0 = 10.0f * log10f(fmaxf(m0, 1e-10f));
              m1 = 10.0f * log10f(fmaxf(m1, 1e-10f));
            } else if (p.log_mode) {
              m0 = fast_ln(zero_handling(m0));
              m1 = fast_ln(zero_handling(m1));
            }
            if (f < nf) {
              float* dst = outp + (int64_t)f * T_out + fa;
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
      __syncwarp();  // ring rows are rewritten by later rounds
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

struct FragTables400 {
  int n_mt = 0;
  int ks_lo[kMtMax] = {0}, ks_n2[kMtMax] = {0}, mel_base[kMtMax] = {0};
  std::vector<float> mel;  // [blocks][hi, lo][32][4]
};

// A fragments of mma.m16n8k8 (row-major 16 filters x 8 bins) for lane (g, t): rows g, g + 8;
// contraction slot s <-> bin 8 ks + 2 (s % 4) + s / 4.  Only the k-steps inside each 16-filter
// group's band are stored, padded to an even count.
static bool build_frags(const Derived& d, const MelBank& mb, float pscale, FragTables400* ft) {
  const int nf = d.nf;
  const int kKs = d.librosa ? Geo<true>::kKs : Geo<false>::kKs;
  if (mb.max_bin > (d.librosa ? kM : kM / 2) || nf > 16 * kMtMax || nf < 1) return false;
  const int MT = (nf + 15) / 16, ncol = 8 * (kKs + 1);  // bands are padded to an even number of k-steps
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

struct Stft400mTables {
  float* d_tab = nullptr;
  int tab_floats = 0;
  int n_mt = 0;
  int ks_lo[kMtMax] = {0}, ks_n2[kMtMax] = {0}, mel_base[kMtMax] = {0};
  bool full = false;  // librosa plan: full-spectrum tile, centred reflect-padded frames
};

bool stft400m_supported(const Derived& d, const MelBank& mb) {
  if (d.W != kW || !d.stft_ok) return false;
  FragTables400 ft;
  if (!build_frags(d, mb, 1.0f, &ft)) return false;
  const size_t fixed = (size_t)(kTabFixed + ft.mel.size() + kMbarFloats) * sizeof(float);
  const size_t per_warp = (d.librosa ? Geo<true>::kWarpFloats : Geo<false>::kWarpFloats) * sizeof(float);
  return fixed + 4 * per_warp <= (size_t)kSmemBudget;
}

int stft400m_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft400mTables** out) {
  FragTables400 ft;
  const float pscale = 0.25f * d.wnorm * d.wnorm;  // (wnorm / 2)^2: the 1/2 of E, O and the STFT normalisation (functions.rs:166-169)
  if (!build_frags(d, mb, pscale, &ft)) return (int)cudaErrorInvalidValue;
  std::vector<float> h(kTabFixed, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  static const int tw_slot[10] = {0, 1, 4, 7, 3, 2, 6, 5, 8, 9};  // = kTwSlot
  for (int c = 0; c < kFpw; ++c)
    for (int u = 0; u < kT; ++u)
      for (int k1 = 0; k1 < kR1; ++k1) {
        const double ang = -2.0 * pi * (double)(u * k1) / (double)kM;
        h[kTabTw + c * kTwCopy + tw_slot[u] * kTwRow + 2 * k1] = (float)std::cos(ang);
        h[kTabTw + c * kTwCopy + tw_slot[u] * kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
      }
  for (int k = 0; k < kM; ++k) {
    // entries 100, 120, .., 180 are read by lane 0 only, for the pairs (10 + 20 m, 190 - 20 m)
    const int kk = (k >= 100 && k % 20 == 0) ? 10 + (k - 100) : k;
    const double ang = -2.0 * pi * (double)kk / (double)kW;
    h[kTabTw400 + 2 * k] = (float)std::cos(ang);
    h[kTabTw400 + 2 * k + 1] = (float)std::sin(ang);
  }
  for (int n = 0; n < kW; ++n) h[kTabWin + n] = window[n];
  h.insert(h.end(), ft.mel.begin(), ft.mel.end());
  Stft400mTables* t = new Stft400mTables;
  t->tab_floats = (int)h.size();
  t->n_mt = ft.n_mt;
  t->full = d.librosa;
  for (int i = 0; i < kMtMax; ++i) { t->ks_lo[i] = ft.ks_lo[i]; t->ks_n2[i] = ft.ks_n2[i]; t->mel_base[i] = ft.mel_base[i]; }
  cudaError_t e = cudaMalloc(&t->d_tab, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete t; return (int)e; }
  e = cudaMemcpy(t->d_tab, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t->d_tab); delete t; return (int)e; }
  *out = t;
  return 0;
}

void stft400m_free_tables(Stft400mTables* t) {
  if (!t) return;
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

// mel_spectrogram (p.mel_out) and / or the complex STFT (p.stft_out); returns -1000 if the shape is declined
int launch_stft400m(const FrameParams& p, const Stft400mTables* t, int sm_count, void* stream) {
  if (p.N != kW || (!p.mel_out && !p.stft_out) || p.T_out < 1 || p.T_out > (1LL << 31) - 64) return -1000;
  if (t->full != (p.mode == kModeLibrosa) || (t->full && (p.stft_out || !p.mel_out))) return -1000;
  const int kWarpFloats = t->full ? Geo<true>::kWarpFloats : Geo<false>::kWarpFloats;
  Stft400mArgs a;
  a.p = p;
  a.tab = t->d_tab;
  a.tab_floats = t->tab_floats;
  a.n_mt = t->n_mt;
  for (int i = 0; i < kMtMax; ++i) { a.ks_lo[i] = t->ks_lo[i]; a.ks_n2[i] = t->ks_n2[i]; a.mel_base[i] = t->mel_base[i]; }
  const size_t fixed = (size_t)(t->tab_floats + kMbarFloats) * sizeof(float);
  const int n_warps = (int)std::min<size_t>(kMaxWarps, (kSmemBudget - fixed) / (kWarpFloats * sizeof(float)));
  if (n_warps < 4) return -1000;
  // units of 24 frames (whole tiles, 96-byte aligned output runs), dealt out in equal contiguous runs
  a.units_per_row = (int)((p.T_out + kChunkAlign - 1) / kChunkAlign);
  a.n_units = p.n_rows * (int64_t)a.units_per_row;
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  // whole rounds can be staged by bulk copies when every frame starts 16-byte aligned
  const int64_t c0 = t->full ? -kW / 2 : (int64_t)(p.n_pad + 1) * p.S - p.N;
  a.stage = p.S == 160 && (p.row_stride % 4 == 0) && (((c0 % 4) + 4) % 4 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 16 == 0);
  a.pair_ok = (p.T_out % 2 == 0) && (reinterpret_cast<uintptr_t>(p.mel_out) % 8 == 0);
  // small jobs: fewer warps per CTA so that the items spread over the SMs instead of queueing on a few
  const int wl = (int)std::min<int64_t>(n_warps, std::max<int64_t>(1, (a.n_units + sm_count - 1) / sm_count));
  const size_t smem = fixed + (size_t)wl * kWarpFloats * sizeof(float);
  const int64_t grid = std::min<int64_t>((a.n_units + wl - 1) / wl, (int64_t)sm_count);
  if (grid <= 0) return 0;
  cudaError_t e;
  auto go = [&](auto kern) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) kern<<<(unsigned)grid, wl * 32, smem, (cudaStream_t)stream>>>(a);
  };
  if (t->full) go(k_stft400m<false, true>);
  else if (p.stft_out) go(k_stft400m<true, false>);
  else go(k_stft400m<false, false>);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  note_kernel(t->full ? "k_stft400m<librosa>" : "k_stft400m");
  return (int)cudaGetLastError();
}

}  // namespace ss
