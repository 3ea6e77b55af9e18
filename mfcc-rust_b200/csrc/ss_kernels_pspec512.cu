// k_pspec512 -- processing::power_spectrum (processing.rs:143-181) for n_fft = 512 on pre-framed rows.
//
// The register FFT of k_fast512c (ss_kernels_fast512c.cu) without the feature epilogue: a half-warp owns a row
// (a frame of L <= 512 samples), a warp two.  Real FFT(512) = complex FFT(256) of z[m] = x[2m] + i x[2m+1] + the
// real-FFT split; 256 = 16 x 16 four-step: a 16-point radix-2 FFT in registers (packed FP32, inputs beyond L
// pruned), the W256^{u k1} twiddle, one transpose through a warp-private padded tile, a second 16-point FFT.
// Lane u then holds Z[u + 16 k2]; its mirror Z[256 - k] sits in lane 16 - u and arrives by warp shuffle.
// Every lane stores |X[k]| / 512 for k = u + 16 k2 and 256 - k: the 16 lanes of a frame write 64 contiguous
// bytes per store instruction.  HBM-bound by design (1.6 KB in, 1 KB out per row); no shared-memory staging of
// the input because rows do not overlap.
#include <cuda_runtime.h>

#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "fft_ct.cuh"
#include "ss_internal.h"

namespace ss {

namespace {

using ssfft::sfor;

constexpr int kWarps = 8;
constexpr int kXs = 18;                       // transpose row stride in float2 (144 B: conflict-free LDS.128)
constexpr int kXchFrame = 16 * kXs;           // float2 per frame
constexpr int kWarpFloats = 2 * kXchFrame * 2;
constexpr int kTwRow = 36;                    // floats per lane row of the W256 table (16 float2 + pad)
constexpr int kTabTw = 16 * kTwRow;

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// JMAX: column groups of 32 samples that carry input (ceil(L / 32)); the FFT prunes the rest
template <int JMAX>
__global__ void __launch_bounds__(kWarps * 32) k_pspec512(const float* __restrict__ frames, int64_t rows, int L, float* __restrict__ out,
                                                           const float* __restrict__ tw_tab, float mscale, int vec2) {
  __shared__ __align__(16) float tw_s[kTabTw];
  __shared__ __align__(16) float xch_s[kWarps * kWarpFloats];
  for (int i = threadIdx.x; i < kTabTw; i += kWarps * 32) tw_s[i] = __ldg(tw_tab + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = lane >> 4, u = lane & 15;
  const bool isz = (u == 0);
  const int partner = (lane & 16) | ((16 - u) & 15);
  const float4* twrow = reinterpret_cast<const float4*>(tw_s + u * kTwRow);
  float2 wu;  // W512^u
  sincospif(-(float)u / 256.0f, &wu.y, &wu.x);
  float2* xch = reinterpret_cast<float2*>(xch_s + warp * kWarpFloats) + fi * kXchFrame;

  const int64_t n_pairs = (rows + 1) >> 1;
  for (int64_t pr = (int64_t)blockIdx.x * kWarps + warp; pr < n_pairs; pr += (int64_t)gridDim.x * kWarps) {
    const int64_t row = 2 * pr + fi;
    const bool valid = row < rows;
    const float* px = frames + (valid ? row : rows - 1) * (int64_t)L + 2 * u;
    // ---- v[j] = x[2m] + i x[2m+1], m = u + 16 j; zero beyond L (util::pad, processing.rs:147-156)
    float2 v[16];
    sfor<0, 16>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      if constexpr (j >= JMAX) {
        v[j] = make_float2(0.0f, 0.0f);
      } else if constexpr (j == JMAX - 1) {
        v[j].x = (2 * u + 32 * j < L) ? __ldg(px + 32 * j) : 0.0f;
        v[j].y = (2 * u + 32 * j + 1 < L) ? __ldg(px + 32 * j + 1) : 0.0f;
      } else {
        if (vec2) v[j] = __ldg(reinterpret_cast<const float2*>(px + 32 * j));
        else v[j] = make_float2(__ldg(px + 32 * j), __ldg(px + 32 * j + 1));
      }
    });
    // ---- stage 1: 16-point FFT over j; twiddle W256^{u k1}; transpose
    ssfft::fft_pow2<16, JMAX>(v);
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
    __syncwarp();
    // ---- stage 2: 16-point FFT over the lane index.  gg[k2] = Z[u + 16 k2]
    ssfft::fft_pow2<16>(gg);
    // ---- real-FFT split for k = u + 16 k2 <= 127: Z[256 - k] is gg[15 - k2] of lane 16 - u (lane 0: its own
    // gg[16 - k2]); each pair yields |X[k]| and |X[256 - k]| (processing.rs:168,180)
    float* orow = out + row * 257;
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
      const float ma = mscale * fast_sqrt(fmaf(xa.x, xa.x, xa.y * xa.y));  // xa = 2 X[k]
      const float mb = mscale * fast_sqrt(fmaf(xb.x, xb.x, xb.y * xb.y));  // xb = 2 conj X[256 - k]
      if (valid) {
        orow[u + 16 * k2] = ma;
        orow[256 - u - 16 * k2] = mb;
      }
    });
    if (isz && valid) orow[128] = 2.0f * mscale * fast_sqrt(fmaf(gg[8].x, gg[8].x, gg[8].y * gg[8].y));  // X[128] = conj Z[128]
  }
}

struct TwEntry {
  float* d = nullptr;
};
std::mutex g_mu;
std::map<int, TwEntry> g_tw;  // per device

}  // namespace

bool pspec512_supported(int64_t L, int64_t n_fft) { return n_fft == 512 && L >= 1 && L <= 512; }

int launch_pspec512(const float* frames, int64_t rows, int64_t L, float* out, void* stream) {
  if (rows <= 0) return 0;
  int dev = 0, smc = 148;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, dev);
  const float* tw = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    TwEntry& t = g_tw[dev];
    if (!t.d) {
      std::vector<float> h(kTabTw, 0.f);
      const double pi = 3.14159265358979323846264338327950288;
      for (int u = 0; u < 16; ++u)
        for (int k1 = 0; k1 < 16; ++k1) {
          const double ang = -2.0 * pi * (double)(u * k1) / 256.0;
          h[u * kTwRow + 2 * k1] = (float)std::cos(ang);
          h[u * kTwRow + 2 * k1 + 1] = (float)std::sin(ang);
        }
      if ((e = cudaMalloc(&t.d, kTabTw * sizeof(float))) != cudaSuccess) return (int)e;
      // synchronous upload, once per device: every stream that later uses the table is ordered behind it
      if ((e = cudaMemcpy(t.d, h.data(), kTabTw * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    }
    tw = t.d;
  }
  const float mscale = 0.5f / 512.0f;  // the 1/2 of E', O' and processing.rs:180
  const int vec2 = (L % 2 == 0) && (reinterpret_cast<uintptr_t>(frames) % 8 == 0);
  const int64_t pairs = (rows + 1) / 2;
  const int64_t grid = std::min<int64_t>((pairs + kWarps - 1) / kWarps, 8LL * smc);
  const int jn = (int)((L + 31) / 32);
  cudaStream_t s = (cudaStream_t)stream;
#define SS_PSPEC_CASE(J) \
  case J: k_pspec512<J><<<(unsigned)grid, kWarps * 32, 0, s>>>(frames, rows, (int)L, out, tw, mscale, vec2); break;
  switch (jn) {
    SS_PSPEC_CASE(1) SS_PSPEC_CASE(2) SS_PSPEC_CASE(3) SS_PSPEC_CASE(4) SS_PSPEC_CASE(5) SS_PSPEC_CASE(6) SS_PSPEC_CASE(7) SS_PSPEC_CASE(8)
    SS_PSPEC_CASE(9) SS_PSPEC_CASE(10) SS_PSPEC_CASE(11) SS_PSPEC_CASE(12) SS_PSPEC_CASE(13) SS_PSPEC_CASE(14) SS_PSPEC_CASE(15)
    default: k_pspec512<16><<<(unsigned)grid, kWarps * 32, 0, s>>>(frames, rows, (int)L, out, tw, mscale, vec2); break;
  }
#undef SS_PSPEC_CASE
  count_launch();
  note_kernel("k_pspec512");
  return (int)cudaGetLastError();
}

}  // namespace ss
