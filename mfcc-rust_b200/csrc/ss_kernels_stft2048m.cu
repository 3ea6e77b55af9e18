// Fused mel-spectrogram kernel for the STFT path with fft_points = 2048, second generation
// (config c3: 22.05 kHz, hop 512, 128 mels):  stream framing (functions.rs:86-160) -> Vorbis window
// -> real FFT(2048) -> |X wnorm|^2 -> mel filterbank (feature.rs:163-174) -> optional
// ln(zero_handling) -> time-minor store.
//
// FFT stage as in the first generation (ss_kernels_stft2048.cu): one warp per frame, 32 complex
// points per lane, 1024 = 32 x 32 four-step FFT in registers (packed FFMA2 butterflies), mirror bin
// by warp shuffle.  New here: the filterbank runs on the tensor cores.
//   * The CTA's 16 warps form two independent groups of 8.  A group transforms 8 consecutive frames
//     of one row per round; every warp writes the |X|^2 row of its frame into the group's [8][516]
//     tile (double-buffered, so one named barrier of 256 threads per round is enough; the two groups
//     drift apart in phase and hide each other's load and barrier latency).
//   * The tile then feeds mel [8 frames][128 filters] = P [8][520 bins] x W^T on mma.sync.m16n8k8 TF32 split
//     3-way (hi x hi + lo x hi + hi x lo) in n-tiles of 8 filters: the M rows are the 8 frames' high parts
//     and, in rows 8..15, their low parts, so two MMAs per k-step cover the three products.  An n-tile visits only
//     the 8-bin k-steps its 8 filters' band touches (83 of 1,040 blocks for the 128-filter bank), and the
//     bands widen with the filter index, so a warp takes n-tile w and n-tile NT - 1 - w: 8 .. 14 k-steps
//     per warp and round.  (The first cut had the filters as M rows, one 16-filter m-tile per warp: 2 .. 22
//     k-steps, and every round's barrier waited for the warp with the widest band; its weight fragments
//     were also twice the bytes.)  The pre-split weight fragments (42 KB) are read from global memory:
//     they are L2 / L1-resident and every warp needs only its own n-tiles' share per round.  (A per-warp
//     |X|^2 tile as in k_fast512c would be 16.6 KB for 8 frames of 513 bins; sixteen of them do not fit
//     next to the transpose tiles.)
//   * The accumulator fragment holds (frame g, filters 2t, 2t + 1): the 8 lanes of a filter store 32
//     consecutive bytes of the time-minor output [row][m][t] straight from registers.
//   * The samples of a warp's next frame (8 KB, contiguous) are requested by one TMA bulk copy into the
//     warp's transpose tile as soon as the transpose of the current frame has been read, and complete on
//     the warp's own mbarrier.
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

constexpr int kW = 2048, kM = 1024, kR = 32;
constexpr int kWarps = 16, kThreads = kWarps * 32;
constexpr int kGroup = 8, kGroups = kWarps / kGroup, kFpr = kGroup;  // 8 frames per group and round
constexpr int kXs = 34;                    // transpose row stride in float2 (272 B: 16-byte rows, conflict-free LDS.128)
constexpr int kWarpFloats = kR * kXs * 2;  // 2176 floats per warp
constexpr int kPs = 516;                   // P row: bins 0..512 + 3 zero pads; 516 % 32 == 4 -> the 4-byte A-fragment loads (row g, bin t) hit 32 banks
constexpr int kKs = 65;                    // k-steps of 8 bins; the last one runs 4 floats into the next row / the slack (finite values x zero weights)
constexpr int kKsTab = kKs + 1;            // bands are padded to an even number of k-steps
constexpr int kTileFloats = kFpr * kPs + 12;  // + slack: the zero-weight block of an odd band reads bins 520..527 of the last row
constexpr int kNtMax = 16;                 // n-tiles of 8 filters: nf <= 128
constexpr int kTwRow = 68;                 // floats per lane row of the W1024 table (32 float2 + pad)
constexpr int kTabTw = 0;                          // [32][kTwRow]
constexpr int kTabTw2048 = kTabTw + kR * kTwRow;   // [512] float2 W2048^k
constexpr int kTabWin = kTabTw2048 + 2 * 512;      // [1024] float2 window pairs
constexpr int kTabFixed = kTabWin + 2 * kM;        // 5248 floats
constexpr int kMbarFloats = 2 * (kWarps + kGroups);  // one 8-byte mbarrier per warp (staged frames) and per group (tile complete), after the tables
static_assert((kTabFixed + kMbarFloats) % 4 == 0, "the transpose tiles start 16-byte aligned");
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

struct Stft2048mArgs {
  FrameParams p;
  const float* tab;        // fixed tables (twiddles, window), copied to shared memory once per CTA
  const float4* mel_frag;  // [blocks][32] {hi b0, hi b1, lo b0, lo b1}: B fragments of the filterbank (8 bins x 8 filters), global
  int64_t n_items;         // rows * rounds_per_row
  int rounds_per_row;
  int vec2;                // frames start 8-byte aligned
  int stage;               // vec2 and a 16-byte aligned base: the next round's frame is staged by TMA bulk copy
  int n_nt;                // n-tiles of 8 filters
  int ks_lo[kNtMax], ks_n2[kNtMax], mel_base[kNtMax];
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

__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned phase) {
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
}

// The filterbank product of round r runs in the middle of round r + 1's FFT, at the one point where the
// transform lives in shared memory (the transpose tile) and the registers are free.  A warp announces its tile
// row on the group's mbarrier and carries on with the next frame's load, window and first FFT stage; only then
// does it wait for the other seven rows, so the group tolerates ~40 % of a round of skew between its warps
// instead of stalling all eight at a bar.sync per round.
__global__ void __launch_bounds__(kThreads, 1) k_stft2048m(const Stft2048mArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* tab_s = smem;
  float* warp_s = smem + kTabFixed + kMbarFloats + (threadIdx.x >> 5) * kWarpFloats;
  float* tile0 = smem + kTabFixed + kMbarFloats + kWarps * kWarpFloats + (threadIdx.x >> 8) * 2 * kTileFloats;  // this group's 2 x [8][kPs] (+ slack)
  const FrameParams& p = a.p;
  const int nf = p.nf;
  for (int i = threadIdx.x; i < kTabFixed; i += kThreads) tab_s[i] = __ldg(a.tab + i);
  // the samples of a warp's next frame arrive by TMA bulk copy and complete on the warp's own mbarrier
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(reinterpret_cast<uint64_t*>(smem + kTabFixed) + (threadIdx.x >> 5));
  if ((threadIdx.x & 31) == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const unsigned gbar = (unsigned)__cvta_generic_to_shared(reinterpret_cast<uint64_t*>(smem + kTabFixed) + kWarps + (threadIdx.x >> 8));
  if ((threadIdx.x & 255) == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(gbar), "r"(kGroup * 32) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned gphase = 0;  // parity of the group barrier phase the pending tile completes
  unsigned phase = 0;   // parity of the mbarrier phase the next staged frame completes
  bool staged = false;  // this round's frame sits in the transpose tile (requested during the previous round)
  int sshift = 0;       // ... starting at this float offset (the copy starts at the 16-byte boundary below the frame)
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
  const int group = warp >> 3, wg = warp & 7;
  int buf = 0, rot = 0;

  // ---- mel filterbank on the tensor cores (feature.rs:163-174):
  //   mel [8 frames][8 filters of n-tile nt] = P [8][band] x W^T,
  // contraction slot t <-> bin 8 ks + t, slot t + 4 <-> bin 8 ks + 4 + t (two conflict-free 4-byte loads of P).
  // An n-tile visits only the 8-bin k-steps its 8 filters' band touches; the bands widen with the filter
  // index (1 .. 13 k-steps for the 128-filter bank), so a warp takes n-tile w and, beyond the first
  // eight, NT - 1 - w: the widest with the narrowest (8 .. 14 k-steps per warp and round; with one m-tile
  // of 16 filters per warp it was 2 .. 22, and every round waited for the last warp).
  // Rows 8..15 of the A operand carry the low parts of the same frames: with A = [P_hi; P_lo] one MMA
  // against W_hi yields hi x hi (rows 0..7) and lo x hi (rows 8..15), a second one against W_lo adds
  // hi x lo (and the negligible lo x lo): two MMAs per k-step instead of three, no idle rows, and the split
  // results are the operand registers as they are.
  // The weight fragments come from L2 (the L1 left beside 227 KB of shared memory does not hold the 46 KB
  // table): the FFT's registers are free at this point, so the first kPre double k-steps of a job are
  // requested at once -- job 0's before the wait for the group's tile -- instead of one iteration ahead.
  constexpr int kPre = 3;
  auto mma_pair = [&](float (&accH)[4], float (&accL)[4], unsigned pp, const float4& w0, const float4& w1) {
    const float2 pv0 = make_float2(lds_f32(pp), lds_f32(pp + 16));  // bins 8 ks + t, 8 ks + 4 + t
    const float2 pv1 = make_float2(lds_f32(pp + 32), lds_f32(pp + 48));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 pv = h ? pv1 : pv0;
      const float4 w = h ? w1 : w0;
      uint32_t ah0, al0, ah2, al2;
      split_tf32(pv.x, ah0, al0);
      split_tf32(pv.y, ah2, al2);
      const float4 av = make_float4(__uint_as_float(ah0), __uint_as_float(al0), __uint_as_float(ah2), __uint_as_float(al2));
      mma_tf32(accH, av, __float_as_uint(w.x), __float_as_uint(w.y));  // [hi; lo] x hi
      mma_tf32(accL, av, __float_as_uint(w.z), __float_as_uint(w.w));  // [hi; lo] x lo
    }
  };
  // out0: (frame t_round + g) of filter 0 in the time-minor output; fr_ok: that frame exists
  // wr: which pair of n-tiles this warp takes.  It rotates with the round, so every warp sees every band width in turn:
  // with a fixed pair the warp that owns the widest bands (16 k-steps against 8) is the last to arrive in every round.
  auto mel_stage = [&](const float* tile, float* const out0, const bool fr_ok, unsigned wait_phase, int wr) {
#pragma unroll 1
    for (int job = 0; job < 2; ++job) {
      const int nt = job == 0 ? wr : a.n_nt - 1 - wr;
      const bool active = nt < a.n_nt && !(job == 1 && nt < kGroup);
      const int n2 = active ? a.ks_n2[nt] : 0;
      const float4* bp = a.mel_frag + (int64_t)(active ? a.mel_base[nt] : 0) * 32 + lane;
      float4 wq[2 * kPre];
      sfor<0, kPre>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        if (i < n2) { wq[2 * i] = __ldg(bp + 64 * i); wq[2 * i + 1] = __ldg(bp + 64 * i + 32); }
      });
      if (job == 0) mbar_wait(gbar, wait_phase);  // the group's eight tile rows are written
      if (!active) continue;
      float accH[4] = {0.f, 0.f, 0.f, 0.f}, accL[4] = {0.f, 0.f, 0.f, 0.f};
      unsigned pp = (unsigned)__cvta_generic_to_shared(tile + g * kPs + t4 + 8 * a.ks_lo[nt]);
      // a ring of kPre double k-steps: each slot is refilled kPre steps ahead as soon as it has been consumed
#pragma unroll 1
      for (int base = 0; base < n2; base += kPre) {
        sfor<0, kPre>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          if (base + i < n2) {
            mma_pair(accH, accL, pp + 64 * i, wq[2 * i], wq[2 * i + 1]);
            if (base + i + kPre < n2) {
              wq[2 * i] = __ldg(bp + 64 * (i + kPre));
              wq[2 * i + 1] = __ldg(bp + 64 * (i + kPre) + 32);
            }
          }
        });
        pp += 64 * kPre;
        bp += 64 * kPre;
      }
      // accumulator rows g / g + 8: (frame t_round + g, filters 8 nt + 2 t4 [+ 1]) from the high / low parts; the
      // 8 lanes of a filter write 32 consecutive bytes of the time-minor output [row][m][t]
      if (fr_ok) {
        float* outp = out0 + (int64_t)(8 * nt + 2 * t4) * p.T_out;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float m = accH[i] + (accH[i + 2] + (accL[i] + accL[i + 2]));
          if (p.log_mode) m = fast_ln(zero_handling(m));
          if (8 * nt + 2 * t4 + i < nf) outp[(int64_t)i * p.T_out] = m;
        }
      }
    }
  };
  bool have_prev = false;
  const float* prev_tile = nullptr;
  float* prev_out0 = nullptr;
  bool prev_ok = false;
  // item = row * rounds_per_row + rr walks by item_step; (row, rr) advance by its quotient and remainder instead of
  // a division per round, in 32 bits (the launcher declines n_items >= 2^31): this state lives across the whole FFT
  const int rpr = a.rounds_per_row, n_rows = (int)p.n_rows;
  const int item_step = (int)gridDim.x * kGroups, step_q = item_step / rpr, step_r = item_step - step_q * rpr;
  int row = ((int)blockIdx.x * kGroups + group) / rpr, rr = ((int)blockIdx.x * kGroups + group) - row * rpr;
  for (; row < n_rows; buf ^= 1) {
    const int t_round = rr * kFpr;
    const float* rowp = p.pcm + (int64_t)row * p.row_stride;
    float* tile = tile0 + buf * kTileFloats;
    {
      const int64_t t = t_round + wg;
      const bool computed = t < p.T;  // frames in [T, T_out) are the reference's zero tail

      // ---- framing + window: v[j] = (x w)[2m] + i (x w)[2m+1], m = u + 32 j
      float2 v[kR];
      {
        const int64_t o0 = t * p.S + c0;
        if (staged) {
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
          const float* sx = warp_s + sshift + 2 * u;
          sfor<0, kR>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const float2 x = *reinterpret_cast<const float2*>(sx + 64 * j);
            const float2 w = win2[32 * j];
            v[j] = __fmul2_rn(x, w);
          });
          __syncwarp();  // the buffer is the transpose tile again
        } else if (!computed) {
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
      if (have_prev) {  // the previous round's filterbank product, while this frame sits in the transpose tile
        mel_stage(prev_tile, prev_out0, prev_ok, gphase, (wg + rot) & (kGroup - 1));
        gphase ^= 1u;
        ++rot;
      }
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
      __syncwarp();  // the transpose tile is free until the next round
      // ---- request this warp's frame of the group's next item (2048 contiguous samples) while the second
      // stage, the split and the filterbank of this one run: one TMA bulk copy (UBLKCP) into the transpose
      // tile.  Only whole interior frames are staged; a source that is 8 but not 16 bytes aligned is copied
      // from the 16-byte boundary below it (16 bytes more).
      {
        int nrow = row + step_q, nrr = rr + step_r;
        if (nrr >= rpr) { nrr -= rpr; ++nrow; }
        staged = false;
        if (a.stage && nrow < n_rows) {
          const int64_t nt = nrr * kFpr + wg;
          const int64_t no0 = nt * p.S + c0;
          if (nt < p.T && no0 >= 0 && no0 + kW + 4 <= p.row_len) {
            const float* src = p.pcm + (int64_t)nrow * p.row_stride + no0;
            sshift = (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 2);  // 0 or 2
            staged = true;
            if (lane == 0) {
              const unsigned bytes = kW * 4 + (sshift ? 16 : 0);
              const unsigned dst = (unsigned)__cvta_generic_to_shared(warp_s);
              // the tile was last touched through the generic proxy (ordered by the __syncwarp above)
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                           "l"(src - sshift), "r"(bytes), "r"(mbar)
                           : "memory");
            }
          }
        }
      }
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
    // the warp's tile row is written (the other buffer belongs to the next round): every lane announces its own stores
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gbar) : "memory");
    have_prev = true;
    prev_tile = tile;
    prev_out0 = p.mel_out + (int64_t)row * nf * p.T_out + t_round + g;
    prev_ok = t_round + g < p.T_out;
    row += step_q;
    rr += step_r;
    if (rr >= rpr) { rr -= rpr; ++row; }
  }
  if (have_prev) mel_stage(prev_tile, prev_out0, prev_ok, gphase, (wg + rot) & (kGroup - 1));
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
  int n_nt = 0;
  int ks_lo[kNtMax] = {0}, ks_n2[kNtMax] = {0}, mel_base[kNtMax] = {0};
  std::vector<float> mel;  // [blocks][32][4]
};

// B fragments of mma.m16n8k8 (col-major 8 bins x 8 filters) for lane (g, t): filter 8 nt + g, contraction
// slots t and t + 4, i.e. bins 8 ks + t and 8 ks + 4 + t, each as {hi, lo}: one float4
// {hi b0, hi b1, lo b0, lo b1} per lane and block.  Only the k-steps inside each 8-filter group's band are
// stored, padded to an even count.
static bool build_frags(const Derived& d, const MelBank& mb, float pscale, FragTables2048* ft) {
  const int nf = d.nf;
  if (mb.max_bin > 512 || nf > 8 * kNtMax || nf < 1) return false;
  const int NT = (nf + 7) / 8, ncol = 8 * kKsTab;
  ft->n_nt = NT;
  std::vector<float> dense((size_t)8 * NT * ncol, 0.f);  // [filter][bin]
  for (int f = 0; f < nf; ++f)
    for (int i = 0; i < mb.len[f]; ++i) dense[(size_t)f * ncol + mb.start[f] + i] = mb.w[mb.wofs[f] + i] * pscale;
  int n_blocks = 0;
  ft->mel.clear();
  for (int nt = 0; nt < NT; ++nt) {
    int lo = kKs, hi = 0;
    for (int ks = 0; ks < kKs; ++ks) {
      bool any = false;
      for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) any = any || dense[(size_t)(8 * nt + r) * ncol + 8 * ks + c] != 0.f;
      if (any) { lo = std::min(lo, ks); hi = std::max(hi, ks + 1); }
    }
    ft->ks_lo[nt] = 0;
    ft->ks_n2[nt] = 0;
    ft->mel_base[nt] = n_blocks;
    if (hi <= lo) continue;  // 8 empty filters: no k-steps, the outputs are exact zeros
    const int n2 = (hi - lo + 1) / 2;
    ft->ks_lo[nt] = lo;
    ft->ks_n2[nt] = n2;
    for (int ks = lo; ks < lo + 2 * n2; ++ks, ++n_blocks) {
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
  return true;
}

}  // namespace

struct Stft2048mTables {
  float* d_tab = nullptr;
  size_t o_mel = 0;
  int n_nt = 0;
  int ks_lo[kNtMax] = {0}, ks_n2[kNtMax] = {0}, mel_base[kNtMax] = {0};
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
  t->n_nt = ft.n_nt;
  for (int i = 0; i < kNtMax; ++i) { t->ks_lo[i] = ft.ks_lo[i]; t->ks_n2[i] = ft.ks_n2[i]; t->mel_base[i] = ft.mel_base[i]; }
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
  a.n_nt = t->n_nt;
  for (int i = 0; i < kNtMax; ++i) { a.ks_lo[i] = t->ks_lo[i]; a.ks_n2[i] = t->ks_n2[i]; a.mel_base[i] = t->mel_base[i]; }
  a.rounds_per_row = (int)((p.T_out + kFpr - 1) / kFpr);
  a.n_items = p.n_rows * (int64_t)a.rounds_per_row;
  if (a.n_items + (int64_t)kGroups * sm_count >= (1LL << 31)) return -1000;  // 32-bit item arithmetic in the kernel
  a.vec2 = (p.S % 2 == 0) && (p.row_stride % 2 == 0) && (reinterpret_cast<uintptr_t>(p.pcm) % 8 == 0);
  static const bool no_stage = getenv("SS_STFT2048M_STAGE") && atoi(getenv("SS_STFT2048M_STAGE")) == 0;  // A/B switch
  a.stage = a.vec2 && !no_stage && (reinterpret_cast<uintptr_t>(p.pcm) % 16 == 0);
  const size_t smem = (size_t)(kTabFixed + kMbarFloats + kWarps * kWarpFloats + kGroups * 2 * kTileFloats) * sizeof(float);
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
