// k_pspec512tc -- the 512-point real FFT + |X|/N of speechpy frames on the 5th-generation tensor cores.
//
// processing::power_spectrum (processing.rs:143-181) for n_fft = 512 as two real GEMM stages on
// tcgen05.mma (kind::tf32, 3-pass split precision), accumulators in TMEM, operands in shared memory:
//
//   n = 16 n1 + n2 (n1 < 32, n2 < 16),  k = k1 + 32 k2 (k1 < 32, k2 < 16)
//   stage 1  Y[k1][n2]  = sum_n1 x[16 n1 + n2] W32^{n1 k1}          real input -> k1 = 0..16 suffice
//   twiddle  T[k1][n2]  = Y[k1][n2] W512^{n2 k1}                     CUDA cores, on the TMEM -> register path
//   stage 2  X[k1+32k2] = sum_n2 T[k1][n2] W16^{n2 k2}               rows (frame, k1): 17 per frame
//
// Stage 1 reads the PCM where the TMA put it: with n = 16 n1 + n2 a frame is 25 (L / 16) rows of 16 floats, and a
// 64-byte-swizzled MN-major operand (atom = 16 M-elements x 8 K-rows, K-rows 64 bytes apart) is exactly that
// layout -- M = n2, K = n1, one atom column per frame, `frame stride` bytes between the atoms.  The tensor
// core takes the upper 19 bits of an f32 word (tf32), so the raw samples ARE the high operand; the low
// operand x - tf32(x) is one pass over the tile on the CUDA cores.  Overlapping frames share their rows.
// Stage 2's operand is written by the epilogue threads of stage 1 (K-major, 128-byte swizzle): thread
// (frame, n2) owns the 17 values Y[.][n2] of its row, twiddles them and stores (re, im) pairs.  The epilogue
// of stage 2 takes sqrt(re^2 + im^2) on the accumulators (the 1/512 is folded into the DFT-16 matrix).
//
// One CTA per SM: G groups of four warps (one TMEM lane quarter each) own a tile slot each and run
// lo-split -> epilogue 1 -> epilogue 2 for their tiles; one more warp issues every tcgen05.mma.  A tile is 7
// frames: 7 x 16 stage-1 rows (one idle atom) and 7 x 17 = 119 stage-2 rows of a 128-row MMA.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include <algorithm>
#include <cstdint>

// standalone micro-benchmark build (profiles/micro/tcgen05_fft512/Makefile): not part of libspeechsauce_cuda.so
namespace ss {
static void count_launch(int = 1) {}
static void note_kernel(const char*) {}
}  // namespace ss

namespace ss {
namespace {

constexpr int kFPT = 7;          // frames per tile
constexpr int kRows2 = 17;       // stage-2 rows per frame (k1 = 0..16)
constexpr int kMaxGroups = 4;
constexpr int kA2Bytes = 128 * 128;   // 128 rows x 32 tf32
constexpr int kBBytes = 32 * 128;     // 32 rows (N) x 32 tf32 (K)
// shared memory map (bytes from the 1024-aligned base)
constexpr int kOffB1Hi = 0, kOffB1Lo = kBBytes, kOffB2Hi = 2 * kBBytes, kOffB2Lo = 3 * kBBytes;
constexpr int kTwRow = 34;                                // floats per n2 row: 8-byte loads of 16 rows hit 32 distinct banks
constexpr int kOffTw = 4 * kBBytes;                       // [16 n2][kTwRow]: W512^{n2 k1} for k1 = 1..15 as (re, im); slots 0,1: W32^{n2}
constexpr int kTabBytes = kOffTw + 16 * kTwRow * 4;       // what the host uploads
constexpr int kOffBars = (kTabBytes + 127) & ~127;        // mbarriers + tmem base
constexpr int kOffGroups = ((kOffBars + 512) + 1023) & ~1023;
constexpr int kSmemBudget = 227 * 1024;

struct TcArgs {
  CUtensorMap map;      // PCM as [rows][16] floats, 64-byte swizzle, box = [16][box_rows]
  int64_t n_clips, T;   // frames per clip
  int64_t clip_rows;    // rows of 16 floats between clips
  int frame_rows;       // rows between frames (S / 16)
  int k_rows;           // ceil(L / 16): rows of a frame that carry samples
  int n_groups;
  int box_rows;
  int in_slot_bytes;    // bytes of one input slot (hi or lo), multiple of 1024
  const float* tab;     // kTabBytes of tables in global memory
  int64_t n_tiles, tiles_per_clip;
  float* pspec;         // [n_clips * T][257] or null
  float* energy;        // [n_clips * T] or null
  int variant;          // bit 0: base_offset from the start address
  float* dbg;           // optional: stage-1 / stage-2 accumulators of tile 0
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | base_offset << 49 | layout << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout,
                                              uint32_t base_offset) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) |
         ((uint64_t)layout << 61);
}
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw64 = 4;
// instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A/B tf32, M = 128, N = 32
constexpr uint32_t make_idesc(bool a_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | (0u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// tile -> first 16-float row of its first frame, frames in it, index of its first frame
__device__ __forceinline__ void tile_geometry(const TcArgs& a, int64_t tile, int64_t* row0, int* frames, int64_t* frame0) {
  const int64_t clip = tile / a.tiles_per_clip, t0 = (tile % a.tiles_per_clip) * kFPT;
  *row0 = clip * a.clip_rows + t0 * a.frame_rows;
  *frames = (int)min((int64_t)kFPT, a.T - t0);
  *frame0 = clip * a.T + t0;
}

__global__ void __launch_bounds__(32 * (4 * kMaxGroups + 1), 1) k_pspec512tc(const __grid_constant__ TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.n_groups;
  const int mma_warp = 4 * G;
  // barriers: per group g: [0] in_full, [1] lo_ready, [2] d1_full, [3] a2_ready, [4] d2_full
  const uint32_t bars = base + kOffBars;
  auto bar = [&](int g, int i) { return bars + 8u * (g * 5 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBars + 8 * 5 * kMaxGroups);

  const int group_bytes = 2 * a.in_slot_bytes + 2 * kA2Bytes;  // in_hi, in_lo, a2_hi, a2_lo
  for (int i = threadIdx.x; i < kTabBytes / 16; i += blockDim.x)
    reinterpret_cast<float4*>(sm)[i] = __ldg(reinterpret_cast<const float4*>(a.tab) + i);
  fence_async_smem();  // the DFT matrices are read through the async proxy (tcgen05.mma)
  if (threadIdx.x == 0) {
    for (int g = 0; g < G; ++g) {
      mbar_init(bar(g, 0), 1);
      mbar_init(bar(g, 1), 4);
      mbar_init(bar(g, 2), 1);
      mbar_init(bar(g, 3), 4);
      mbar_init(bar(g, 4), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == mma_warp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t gstride = (int64_t)gridDim.x * G;
  if (warp == mma_warp) {
    // ===================================================== the tcgen05.mma issuer (one elected lane)
    if (lane == 0) {
      const uint32_t bo = a.variant & 1;
      const uint32_t lbo1 = (uint32_t)a.frame_rows * 64u;
      for (int64_t it = 0;; ++it) {
        bool any = false;
        const uint32_t ph = (uint32_t)(it & 1);
        for (int g = 0; g < G; ++g) {  // stage 1 of every group's tile `it`
          if ((int64_t)blockIdx.x * G + g + it * gstride >= a.n_tiles) continue;
          any = true;
          mbar_wait(bar(g, 0), ph);  // the TMA tile has landed ...
          mbar_wait(bar(g, 1), ph);  // ... and its low operand is written
          tc_fence_after();
          const uint32_t in_hi = base + kOffGroups + g * group_bytes, in_lo = in_hi + a.in_slot_bytes;
          const uint32_t d1 = tmem + (uint32_t)(64 * g);
          uint32_t acc = 0;
          for (int pass = 0; pass < 3; ++pass) {  // hi.hi, lo.hi, hi.lo
            const uint32_t A = pass == 1 ? in_lo : in_hi;
            const uint32_t B = base + (pass == 2 ? kOffB1Lo : kOffB1Hi);
            for (int ks = 0; ks < 4; ++ks) {
              if (ks * 8 >= a.k_rows) break;  // rows beyond the frame only meet zero weights
              const uint32_t aa = A + ks * 512;
              mma_tf32(d1, make_desc(aa, lbo1, 512, kLayoutSw64, bo ? (aa >> 7) & 7 : 0),
                       make_desc(B + ks * 32, 16, 1024, kLayoutSw128, 0), make_idesc(true), acc);
              acc = 1;
            }
          }
          mma_commit(bar(g, 2));
        }
        for (int g = 0; g < G; ++g) {  // stage 2
          if ((int64_t)blockIdx.x * G + g + it * gstride >= a.n_tiles) continue;
          mbar_wait(bar(g, 3), ph);
          tc_fence_after();
          const uint32_t a2_hi = base + kOffGroups + g * group_bytes + 2 * a.in_slot_bytes, a2_lo = a2_hi + kA2Bytes;
          const uint32_t d2 = tmem + (uint32_t)(64 * g + 32);
          uint32_t acc = 0;
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t A = pass == 1 ? a2_lo : a2_hi;
            const uint32_t B = base + (pass == 2 ? kOffB2Lo : kOffB2Hi);
            for (int ks = 0; ks < 4; ++ks) {
              mma_tf32(d2, make_desc(A + ks * 32, 16, 1024, kLayoutSw128, 0), make_desc(B + ks * 32, 16, 1024, kLayoutSw128, 0),
                       make_idesc(false), acc);
              acc = 1;
            }
          }
          mma_commit(bar(g, 4));
        }
        if (!any) break;
      }
    }
  } else if (warp < 4 * G) {
    // ===================================================== group g: lo-split, epilogue 1, epilogue 2
    const int g = warp >> 2, q = warp & 3;
    const int gl = q * 32 + lane;                 // TMEM lane = row of the 128-row tiles
    uint8_t* grp = sm + kOffGroups + g * group_bytes;
    float* in_hi = reinterpret_cast<float*>(grp);
    float* in_lo = reinterpret_cast<float*>(grp + a.in_slot_bytes);
    uint8_t* a2_hi = grp + 2 * a.in_slot_bytes;
    uint8_t* a2_lo = a2_hi + kA2Bytes;
    const uint32_t in_hi_u = base + kOffGroups + g * group_bytes;
    const bool leader = q == 0 && lane == 0;
    const int box_bytes = a.box_rows * 64;
    const int64_t first = (int64_t)blockIdx.x * G + g;
    // stage-1 row of this thread: frame slot j1, n2; its twiddles W512^{n2 k1}
    const int j1 = gl >> 4, n2 = gl & 15;
    const float* tw = reinterpret_cast<const float*>(sm + kOffTw) + n2 * kTwRow;
    // stage-2 row of this thread: r = 17 j2 + k1b
    const int j2 = gl / kRows2, k1b = gl % kRows2;

    int it_dbg = 0;
    auto issue_tma = [&](int64_t tile) {
      int64_t row0, f0;
      int fr;
      tile_geometry(a, tile, &row0, &fr, &f0);
      mbar_expect_tx(bar(g, 0), (uint32_t)box_bytes);
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(in_hi_u),
                   "l"(&a.map), "r"(0), "r"((int)row0), "r"(bar(g, 0))
                   : "memory");
    };
    // low operand of stage 1: x - tf32(x), element-wise at the same (swizzled) addresses
    auto lo_split = [&](uint32_t ph) {
      mbar_wait(bar(g, 0), ph);
      if (a.dbg && first == 0 && it_dbg == 0) {  // the raw tile as the TMA left it + the DFT-32 matrix as the MMA sees it
        for (int i = gl; i < box_bytes / 4; i += 128) a.dbg[8192 + i] = in_hi[i];
        for (int i = gl; i < 1024; i += 128) a.dbg[8192 + 4096 + i] = reinterpret_cast<const float*>(sm + kOffB1Hi)[i];
      }
      const float4* src = reinterpret_cast<const float4*>(in_hi);
      float4* dst = reinterpret_cast<float4*>(in_lo);
      const int n4 = box_bytes / 16;
      for (int i = gl; i < n4; i += 128) {
        float4 v = src[i];
        v.x = tf32_lo(v.x);
        v.y = tf32_lo(v.y);
        v.z = tf32_lo(v.z);
        v.w = tf32_lo(v.w);
        dst[i] = v;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(g, 1));
    };
    if (first < a.n_tiles) {
      if (leader) issue_tma(first);
      lo_split(0);
    }
    // byte offsets of this thread's 17 (re, im) slots in the stage-2 operand (K-major, 128-byte swizzle)
    uint32_t a2off[kRows2];
#pragma unroll
    for (int k = 0; k < kRows2; ++k) {
      const int r = kRows2 * j1 + k;
      a2off[k] = (uint32_t)r * 128u + ((((uint32_t)n2 >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)n2 & 1u) * 8u;
    }
    int64_t it = 0;
    for (int64_t tile = first; tile < a.n_tiles; tile += gstride, ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const bool more = tile + gstride < a.n_tiles;
      int64_t row0, frame0;
      int frames;
      tile_geometry(a, tile, &row0, &frames, &frame0);
      // ---- epilogue 1: D1 row (j1, n2) -> twiddle -> rows 17 j1 + k1 of the stage-2 operand
      mbar_wait(bar(g, 2), ph);
      tc_fence_after();
      if (leader && more) issue_tma(tile + gstride);  // stage 1 has consumed the input slot
      {
        float y[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * g), y);
        if (a.dbg && tile == 0)
          for (int i = 0; i < 32; ++i) a.dbg[gl * 32 + i] = y[i];
        if (j1 < kFPT) {
          // y[0] = Y[0] (real), y[1] = Y[16] (real), y[2k], y[2k+1] = Re, Im Y[k] for k = 1..15
          float2 t[kRows2];
          t[0] = make_float2(y[0], 0.f);
          {
            const float2 w = *reinterpret_cast<const float2*>(tw);
            t[16] = make_float2(y[1] * w.x, y[1] * w.y);
          }
#pragma unroll
          for (int k = 1; k < 16; ++k) {
            const float2 w = *reinterpret_cast<const float2*>(tw + 2 * k);
            t[k] = make_float2(y[2 * k] * w.x - y[2 * k + 1] * w.y, y[2 * k] * w.y + y[2 * k + 1] * w.x);
          }
#pragma unroll
          for (int k = 0; k < kRows2; ++k) {
            *reinterpret_cast<float2*>(a2_hi + a2off[k]) = t[k];
            *reinterpret_cast<float2*>(a2_lo + a2off[k]) = make_float2(tf32_lo(t[k].x), tf32_lo(t[k].y));
          }
        }
      }
      if (a.dbg && tile == 0)
        for (int i = 0; i < 32; ++i) a.dbg[8192 + 4096 + 1024 + gl * 32 + i] = reinterpret_cast<const float*>(a2_hi)[gl * 32 + i];
      it_dbg = 1;
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(g, 3));
      // ---- the next tile's low operand, while the tensor cores run stage 2 of this one
      if (more) lo_split(ph ^ 1);
      // ---- epilogue 2: D2 row (j2, k1) holds X[k1 + 32 k2] / 512, k2 = 0..15
      mbar_wait(bar(g, 4), ph);
      tc_fence_after();
      {
        float x[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * g + 32), x);
        if (a.dbg && tile == 0)
          for (int i = 0; i < 32; ++i) a.dbg[4096 + gl * 32 + i] = x[i];
        float e = 0.f;
        const bool live = gl < kFPT * kRows2 && j2 < frames;
        if (live) {
          float* prow = a.pspec ? a.pspec + (frame0 + j2) * 257 : nullptr;
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const int k = k1b + 32 * k2;
            const bool dup = (k1b == 0 && k2 > 8) || (k1b == 16 && k2 > 7);
            const float m = fast_sqrt(fmaf(x[2 * k2], x[2 * k2], x[2 * k2 + 1] * x[2 * k2 + 1]));
            if (!dup) {
              e += m;
              if (prow) prow[k <= 256 ? k : 512 - k] = m;
            }
          }
        }
        if (a.energy) {
          // the 17 rows of a frame are consecutive lanes: segmented shuffle sum, one atomic per run inside a warp
          const unsigned full = 0xffffffffu;
          float s = e;
          const int seg = live ? j2 : -1 - lane;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const float o = __shfl_down_sync(full, s, d);
            const int os = __shfl_down_sync(full, seg, d);
            if (lane + d < 32 && os == seg) s += o;
          }
          const int prev = __shfl_up_sync(full, seg, 1);
          if (live && (lane == 0 || prev != seg)) atomicAdd(a.energy + frame0 + j2, s);
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == mma_warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---- descriptor self-test (development hook): TMEM st/ld round trip and three operand layouts against
// D[r][n] = A[r][n] (n + 1) (B = diag(1..32)).  out: 4 x [128][32].
__global__ void __launch_bounds__(128, 1) k_tc_selftest(float* out, int base_offset_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  // layout: A 16 KB | B 4 KB | bar | tmem slot
  float* A = reinterpret_cast<float*>(sm);
  float* B = reinterpret_cast<float*>(sm + 16384);
  const uint32_t bar = base + 16384 + 4096;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + 16384 + 4096 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t my = tmem + ((uint32_t)(warp * 32) << 16);
  // B (N x K, K-major SW128): B[n][k] = (n == k) ? n + 1 : 0
  for (int i = tid; i < 1024; i += 128) B[i] = 0.f;
  __syncthreads();
  if (tid < 32) {
    const int n = tid, k = tid;
    B[(n * 128 + ((((k >> 2) ^ (n & 7))) << 4) + (k & 3) * 4) / 4] = (float)(n + 1);
  }
  // ---- test 0: tcgen05.st -> tcgen05.ld
  {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(tid * 100 + i));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(my), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float y[32];
    tmem_ld32(my, y);
    for (int i = 0; i < 32; ++i) out[tid * 32 + i] = y[i];
  }
  uint32_t phase = 0;
  for (int test = 1; test <= 3; ++test) {
    __syncthreads();
    for (int i = tid; i < 4096; i += 128) A[i] = 0.f;
    __syncthreads();
    // logical A[m][k], m < 128, k < 32
    for (int e = tid; e < 128 * 32; e += 128) {
      const int m = e >> 5, k = e & 31;
      uint32_t off;
      float val;
      if (test == 1) {         // K-major, 128-byte swizzle: row m = 128 bytes
        val = (float)((m * 32 + k) % 251);
        off = m * 128 + ((((k >> 2) ^ (m & 7))) << 4) + (k & 3) * 4;
      } else if (test == 2) {  // MN-major, 64-byte swizzle, aligned atoms: M-atom a at a * 2048, K-group at + 512
        val = (float)((m * 32 + k) % 251);
        off = (m >> 4) * 2048 + (k >> 3) * 512 + (k & 7) * 64 + (m & 15) * 4;
        off ^= ((off >> 7) & 3) << 4;
      } else {                 // MN-major, 64-byte swizzle, overlapping atoms 640 bytes apart (the PCM tile)
        const int row = 10 * (m >> 4) + k, col = m & 15;
        val = (float)((row * 16 + col) % 251);
        off = row * 64 + col * 4;
        off ^= ((off >> 7) & 3) << 4;
      }
      A[off / 4] = val;
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      uint32_t acc = 0;
      for (int ks = 0; ks < 4; ++ks) {
        uint64_t da;
        if (test == 1) da = make_desc(base + ks * 32, 16, 1024, kLayoutSw128, 0);
        else if (test == 2) da = make_desc(base + ks * 512, 2048, 512, kLayoutSw64, base_offset_mode ? ((base + ks * 512) >> 7) & 7 : 0);
        else da = make_desc(base + ks * 512, 640, 512, kLayoutSw64, base_offset_mode ? ((base + ks * 512) >> 7) & 7 : 0);
        mma_tf32(tmem, da, make_desc(base + 16384 + ks * 32, 16, 1024, kLayoutSw128, 0), make_idesc(test != 1), acc);
        acc = 1;
      }
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    float y[32];
    tmem_ld32(my, y);
    for (int i = 0; i < 32; ++i) out[test * 4096 + tid * 32 + i] = y[i];
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

// ---- host side -----------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

inline float tf32_hi_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}
// B operand "N x K, K-major, 128-byte swizzle": element (n, k) at n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4
void put_b(std::vector<float>& hi, std::vector<float>& lo, int n, int k, double v) {
  const float f = (float)v, h = tf32_hi_host(f), l = tf32_hi_host(f - h);
  const size_t off = ((size_t)n * 128 + ((((size_t)k >> 2) ^ ((size_t)n & 7)) << 4) + ((size_t)k & 3) * 4) / 4;
  hi[off] = h;
  lo[off] = l;
}

}  // namespace

struct Pspec512TcTables {
  float* d_tab = nullptr;  // B1 hi/lo, B2 hi/lo, twiddles (kTabBytes)
};

bool pspec512tc_supported(int64_t L, int64_t frame_stride, int64_t n_fft) {
  return n_fft == 512 && L > 0 && L <= 512 && L % 16 == 0 && frame_stride > 0 && frame_stride % 16 == 0 && frame_stride <= 512 &&
         get_encode() != nullptr;
}

int pspec512tc_build_tables(int64_t L, Pspec512TcTables** out) {
  std::vector<float> b1h(kBBytes / 4, 0.f), b1l(kBBytes / 4, 0.f), b2h(kBBytes / 4, 0.f), b2l(kBBytes / 4, 0.f), tw(16 * kTwRow, 0.f);
  const double pi = 3.14159265358979323846264338327950288;
  const int k_rows = (int)((L + 15) / 16);
  // stage 1: D1[(f, n2)][col] = sum_n1 x[16 n1 + n2] B1[n1][col]; col 0: Y[0], col 1: Y[16], col 2k, 2k+1: Re, Im Y[k]
  for (int n1 = 0; n1 < k_rows; ++n1) {
    put_b(b1h, b1l, 0, n1, 1.0);
    put_b(b1h, b1l, 1, n1, (n1 & 1) ? -1.0 : 1.0);
    for (int k = 1; k < 16; ++k) {
      const double ang = -2.0 * pi * (double)((n1 * k) % 32) / 32.0;
      put_b(b1h, b1l, 2 * k, n1, std::cos(ang));
      put_b(b1h, b1l, 2 * k + 1, n1, std::sin(ang));
    }
  }
  // stage 2: D2[(f, k1)][(k2, c')] = sum_(n2, c) T[(n2, c)] B2[(n2, c)][(k2, c')], w = W16^{n2 k2} / 512
  for (int n2 = 0; n2 < 16; ++n2)
    for (int k2 = 0; k2 < 16; ++k2) {
      const double ang = -2.0 * pi * (double)((n2 * k2) % 16) / 16.0, wr = std::cos(ang) / 512.0, wi = std::sin(ang) / 512.0;
      put_b(b2h, b2l, 2 * k2, 2 * n2, wr);
      put_b(b2h, b2l, 2 * k2, 2 * n2 + 1, -wi);
      put_b(b2h, b2l, 2 * k2 + 1, 2 * n2, wi);
      put_b(b2h, b2l, 2 * k2 + 1, 2 * n2 + 1, wr);
    }
  for (int n2 = 0; n2 < 16; ++n2) {
    tw[n2 * kTwRow + 0] = (float)std::cos(-2.0 * pi * n2 / 32.0);  // W512^{16 n2} = W32^{n2}
    tw[n2 * kTwRow + 1] = (float)std::sin(-2.0 * pi * n2 / 32.0);
    for (int k = 1; k < 16; ++k) {
      const double ang = -2.0 * pi * (double)(n2 * k) / 512.0;
      tw[n2 * kTwRow + 2 * k] = (float)std::cos(ang);
      tw[n2 * kTwRow + 2 * k + 1] = (float)std::sin(ang);
    }
  }
  std::vector<float> h(kTabBytes / 4, 0.f);
  std::memcpy(h.data() + kOffB1Hi / 4, b1h.data(), kBBytes);
  std::memcpy(h.data() + kOffB1Lo / 4, b1l.data(), kBBytes);
  std::memcpy(h.data() + kOffB2Hi / 4, b2h.data(), kBBytes);
  std::memcpy(h.data() + kOffB2Lo / 4, b2l.data(), kBBytes);
  std::memcpy(h.data() + kOffTw / 4, tw.data(), 16 * kTwRow * 4);
  auto* t = new Pspec512TcTables();
  cudaError_t e = cudaMalloc(&t->d_tab, kTabBytes);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_tab, h.data(), kTabBytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete t;
    return (int)e;
  }
  *out = t;
  return 0;
}
void pspec512tc_free_tables(Pspec512TcTables* t) {
  if (!t) return;
  cudaFree(t->d_tab);
  delete t;
}

// frames: n_clips clips of T frames, frame t of clip c at pcm + c * clip_stride + t * frame_stride (floats), L samples
// each.  pspec [n_clips * T][257] and / or energy [n_clips * T] (energy is accumulated: the caller zeroes it).
int launch_pspec512tc(const float* pcm, int64_t n_clips, int64_t T, int64_t clip_stride, int64_t frame_stride, int64_t L,
                      int64_t total_floats, const Pspec512TcTables* t, float* pspec, float* energy, int sm_count, int variant,
                      float* dbg, void* stream) {
  if (!pspec512tc_supported(L, frame_stride, 512) || clip_stride % 16 || reinterpret_cast<uintptr_t>(pcm) % 64) return -1000;
  if (n_clips <= 0 || T <= 0) return 0;
  TcArgs a;
  std::memset(&a, 0, sizeof a);
  const cuuint64_t rows = (cuuint64_t)(total_floats / 16);
  const int frame_rows = (int)(frame_stride / 16);
  const int box_rows = (kFPT - 1) * frame_rows + 32;
  if (box_rows > 256 || rows < 1 || rows > 0x7fffffff) return -1000;
  // a slot holds the box and the idle eighth atom column (rows 7 frame_rows .. +32) the 128-row MMA also reads
  const int in_slot = (std::max(box_rows * 64, 7 * frame_rows * 64 + 2048) + 1023) & ~1023;
  const cuuint64_t gdim[2] = {16, rows};
  const cuuint64_t gstr[1] = {64};
  const cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&a.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pcm), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  a.n_clips = n_clips;
  a.T = T;
  a.clip_rows = clip_stride / 16;
  a.frame_rows = frame_rows;
  a.k_rows = (int)((L + 15) / 16);
  a.box_rows = box_rows;
  a.in_slot_bytes = in_slot;
  a.tab = t->d_tab;
  a.tiles_per_clip = (T + kFPT - 1) / kFPT;
  a.n_tiles = n_clips * a.tiles_per_clip;
  a.pspec = pspec;
  a.energy = energy;
  a.variant = variant;
  a.dbg = dbg;
  const int group_bytes = 2 * in_slot + 2 * kA2Bytes;
  int G = std::min(kMaxGroups, (kSmemBudget - 1024 - kOffGroups) / group_bytes);
  if (variant & 0x100) G = 1;
  if (G < 1) return -1000;
  a.n_groups = G;
  const size_t smem = 1024 + kOffGroups + (size_t)G * group_bytes;
  cudaError_t e = cudaFuncSetAttribute(k_pspec512tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int64_t grid = std::min<int64_t>(sm_count, (a.n_tiles + G - 1) / G);
  k_pspec512tc<<<(unsigned)grid, 32 * (4 * G + 1), smem, (cudaStream_t)stream>>>(a);
  count_launch();
  note_kernel("k_pspec512tc");
  return (int)cudaGetLastError();
}

int launch_tc_selftest(float* out, int base_offset_mode, void* stream) {
  const size_t smem = 1024 + 16384 + 4096 + 256;
  cudaError_t e = cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_tc_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(out, base_offset_mode);
  return (int)cudaGetLastError();
}

}  // namespace ss

// ---- C entry point of the probe (tc_probe.py) -------------------------------------------------------------
extern "C" int tcfft_run(void* stream, const float* d_pcm, int64_t n_clips, int64_t T, int64_t clip_stride, int64_t frame_stride,
                         int64_t L, int64_t total_floats, float* d_pspec, float* d_energy, int variant, float* d_dbg) {
  if (variant & 0x1000) return ss::launch_tc_selftest(d_dbg, variant & 1, stream);
  static ss::Pspec512TcTables* tab = nullptr;
  static int64_t tab_L = -1;
  if (!ss::pspec512tc_supported(L, frame_stride, 512)) return -1000;
  if (tab_L != L) {
    if (tab) ss::pspec512tc_free_tables(tab);
    tab = nullptr;
    if (int e = ss::pspec512tc_build_tables(L, &tab)) return e;
    tab_L = L;
  }
  int dev = 0, smc = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, dev);
  return ss::launch_pspec512tc(d_pcm, n_clips, T, clip_stride, frame_stride, L, total_floats, tab, d_pspec, d_energy, smc, variant, d_dbg, stream);
}
