// HBM-bound standalone kernels behind the reference's config-less processing functions:
//   processing::preemphasis   processing.rs:31-53
//   processing::stack_frames  processing.rs:65-129
//   processing::cmvn          processing.rs:265-300
// All are pure streaming kernels: coalesced (float4 where alignment allows) loads and stores,
// grids sized in multiples of the SM count, no intermediate buffers except the cmvn partial
// statistics.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "ss_internal.h"

namespace ss {

namespace {

constexpr int kSmCountGuess = 148;

// ------------------------------------------------------------------------- preemphasis
// out[i] = in[i] - cof * in[(i - shift) mod n]   (np.roll semantics, circular, Q17)
__global__ void __launch_bounds__(256) k_preemphasis(const float* __restrict__ in, float* __restrict__ out,
                                                     int64_t n_rows, int64_t n, int64_t row_stride,
                                                     int64_t shift_mod, float cof, int vec_ok, int bx) {
  const int64_t row = blockIdx.x / bx;
  const float* x = in + row * row_stride;
  float* y = out + row * row_stride;
  const int64_t tid = (blockIdx.x % bx) * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)bx * blockDim.x;
  if (vec_ok) {
    // 4 outputs per thread; the rolled operand is generally misaligned, so it is read as
    // scalars (served by L1 from the lines the aligned stream just brought in).
    const int64_t n4 = n >> 2;
    for (int64_t v = tid; v < n4; v += nthreads) {
      const int64_t i = v << 2;
      const float4 a = *reinterpret_cast<const float4*>(x + i);
      int64_t j = i - shift_mod;
      if (j < 0) j += n;
      float r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int64_t jj = j + u;
        if (jj >= n) jj -= n;
        r[u] = __ldg(x + jj);
      }
      float4 o;
      // two roundings like the reference's `signal - (cof * rolled)`: no FMA contraction
      o.x = __fsub_rn(a.x, __fmul_rn(cof, r[0]));
      o.y = __fsub_rn(a.y, __fmul_rn(cof, r[1]));
      o.z = __fsub_rn(a.z, __fmul_rn(cof, r[2]));
      o.w = __fsub_rn(a.w, __fmul_rn(cof, r[3]));
      *reinterpret_cast<float4*>(y + i) = o;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthreads) {
      int64_t j = i - shift_mod;
      if (j < 0) j += n;
      y[i] = __fsub_rn(x[i], __fmul_rn(cof, x[j]));
    }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) {
      int64_t j = i - shift_mod;
      if (j < 0) j += n;
      y[i] = __fsub_rn(x[i], __fmul_rn(cof, x[j]));
    }
  }
}

// ------------------------------------------------------------------------- stack_frames
// out[f][j] = x[f*S + j] (zero beyond n: the zero_padding=true tail), or the literal rows.
__global__ void __launch_bounds__(256) k_stack_frames(const float* __restrict__ in, float* __restrict__ out,
                                                      int64_t n, int64_t L, int64_t S, int64_t T, uint32_t flags) {
  const int64_t total = T * L;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t f = i / L, j = i - f * L;
    float v;
    if (flags & SS_FLAG_FRAMING_LITERAL) {
      v = (T <= 2 && j < 2 * (L / 2) && j < n) ? in[j] : 0.0f;
    } else {
      const int64_t s = f * S + j;
      v = s < n ? __ldg(in + s) : 0.0f;
    }
    out[i] = v;
  }
}

// 16-byte path: L, S multiples of 4 and 16-byte aligned bases.  A block copies kSfFrames consecutive frames;
// thread i moves float4 i of the block's [frames][L / 4] slab (the overlapping part of neighbouring frames is
// re-read through L1 / L2: DRAM sees every sample once).
constexpr int kSfFrames = 16;
__global__ void __launch_bounds__(256) k_stack_frames_v4(const float4* __restrict__ in, float4* __restrict__ out, int64_t n4,
                                                         int L4, int S4, int64_t T) {
  for (int64_t f0 = (int64_t)blockIdx.x * kSfFrames; f0 < T; f0 += (int64_t)gridDim.x * kSfFrames) {
    const int frames = (int)min((int64_t)kSfFrames, T - f0);
    const int total = frames * L4;
    const float4* src = in + f0 * S4;
    float4* dst = out + f0 * L4;
    const int64_t left = n4 - f0 * S4;  // float4s of the signal from this slab's first sample on (zero padding beyond)
#pragma unroll 4
    for (int i = threadIdx.x; i < total; i += 256) {
      const int f = i / L4, j = i - f * L4;
      const int o = f * S4 + j;
      dst[i] = o < left ? __ldg(src + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------- cmvn
// Small matrices (the per-clip cepstra: 97 x 13): one warp per matrix, the matrix parked in shared memory between
// the statistics and the normalisation -> one DRAM read and one write, no scratch.  Columns are summed down the
// rows in order (ndarray's mean_axis(0) is a sequential row accumulation, processing.rs:270), by 32 / cols row
// groups side by side; the variance is the two-pass form of std_axis(0, ddof 0) (processing.rs:283).
constexpr int kCmvnSmallWarps = 8;
__global__ void __launch_bounds__(kCmvnSmallWarps * 32) k_cmvn_small(const float* in, float* out,  // in == out is allowed
                                                                     int64_t n_mats, int rows, int cols, int varnorm,
                                                                     int warp_floats) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* x = sm + (size_t)warp * warp_floats;  // [rows * cols] then [R][cols] partials, [cols] mean, [cols] denominator
  const int n = rows * cols;
  const int R = cols <= 16 ? 32 / cols : 1;    // row groups working side by side
  float* part = x + ((n + 3) & ~3);
  float* mean_s = part + R * cols;
  float* den_s = mean_s + cols;
  const int rg = lane / cols, c0 = lane - rg * cols;
  const int rpg = (rows + R - 1) / R;
  for (int64_t m = (int64_t)blockIdx.x * kCmvnSmallWarps + warp; m < n_mats; m += (int64_t)gridDim.x * kCmvnSmallWarps) {
    const float* src = in + m * n;
    float* dst = out + m * n;
    for (int i = lane; i < n; i += 32) x[i] = src[i];  // plain loads: the host layer runs this in place
    __syncwarp();
    // pass 0: column sums -> mean.  pass 1: sum (x - mean)^2 -> std.  (std_axis subtracts the mean of y = x - mean
    // once more; that residual is a rounding error of ~1e-8 sigma and changes nothing at f32 resolution.)
    for (int pass = 0; pass < (varnorm ? 2 : 1); ++pass) {
      if (cols <= 16) {
        if (rg < R) {
          const int r0 = rg * rpg, r1 = min(rows, r0 + rpg);
          const float mu = pass ? mean_s[c0] : 0.f;
          float acc = 0.f;
          for (int r = r0; r < r1; ++r) {
            const float v = x[r * cols + c0] - mu;
            acc += pass ? v * v : v;
          }
          part[rg * cols + c0] = acc;
        }
      } else {
        for (int c = lane; c < cols; c += 32) {
          const float mu = pass ? mean_s[c] : 0.f;
          float acc = 0.f;
          for (int r = 0; r < rows; ++r) {
            const float v = x[r * cols + c] - mu;
            acc += pass ? v * v : v;
          }
          part[c] = acc;
        }
      }
      __syncwarp();
      for (int c = lane; c < cols; c += 32) {
        float t = part[c];
        for (int g = 1; g < R; ++g) t += part[g * cols + c];
        if (pass == 0) mean_s[c] = t / (float)rows;
        else den_s[c] = 1.0f / (sqrtf(t / (float)rows) + kCmvnEps);  // processing.rs:283-292
      }
      __syncwarp();
    }
    {
      int c = lane % cols;
      const int step = 32 % cols;
      for (int i = lane; i < n; i += 32) {
        const float d = x[i] - mean_s[c];
        dst[i] = varnorm ? d * den_s[c] : d;
        c += step;
        if (c >= cols) c -= cols;
      }
    }
    __syncwarp();
  }
}

// Smaller still (rows * COLS <= kCmvnRegJ * stride; COLS one of the usual cepstral counts: the per-clip cepstra,
// 97 x 13): the matrix lives in the registers of one warp.  The first stride = (32 / COLS) * COLS lanes walk it
// with that stride, a multiple of COLS, so a lane stays in one column: no index arithmetic (COLS is a template
// parameter: every load and store is base + immediate), no shared memory, and all of a lane's loads are in flight
// at once (k_cmvn_small spends two thirds of its issue slots on addressing and shared-memory traffic).  The
// 32 / COLS lanes of a column add their sums in lane order; variance in the two-pass form as above.
constexpr int kCmvnRegJ = 64;
template <int COLS>
__global__ void __launch_bounds__(256, 2) k_cmvn_reg(const float* in, float* out, int64_t n_mats, int rows, int varnorm) {  // in == out is allowed
  constexpr int R = 32 / COLS, stride = R * COLS;
  const int lane = threadIdx.x & 31;
  const int n = rows * COLS;
  const int c = lane % COLS;
  const bool act = lane < stride;
  const int nj = act ? (n - lane + stride - 1) / stride : 0;  // this lane's elements
  const float inv_rows = 1.0f / (float)rows;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < n_mats; m += warps) {
    const float* ps = in + m * n + lane;
    float* pd = out + m * n + lane;
    float v[kCmvnRegJ];
#pragma unroll
    for (int j = 0; j < kCmvnRegJ; ++j) v[j] = j < nj ? ps[j * stride] : 0.f;  // plain loads: the host layer runs this in place
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kCmvnRegJ; ++j) s += v[j];
    float tot = 0.f;
#pragma unroll
    for (int g = 0; g < R; ++g) tot += __shfl_sync(0xffffffffu, s, c + g * COLS);
    const float mean = tot * inv_rows;
    float inv = 1.0f;
    if (varnorm) {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < kCmvnRegJ; ++j) {
        const float d = v[j] - mean;
        q += j < nj ? d * d : 0.f;
      }
      float tq = 0.f;
#pragma unroll
      for (int g = 0; g < R; ++g) tq += __shfl_sync(0xffffffffu, q, c + g * COLS);
      inv = 1.0f / (sqrtf(tq * inv_rows) + kCmvnEps);  // processing.rs:283-292
    }
#pragma unroll
    for (int j = 0; j < kCmvnRegJ; ++j)
      if (j < nj) pd[j * stride] = (v[j] - mean) * inv;
  }
}
template <int COLS>
static int launch_cmvn_reg(const float* in, float* out, int64_t n_mats, int64_t rows, int varnorm, void* stream) {
  const int64_t blocks = std::min<int64_t>((n_mats + 7) / 8, 16LL * kSmCountGuess);
  k_cmvn_reg<COLS><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n_mats, (int)rows, varnorm);
  count_launch();
  return (int)cudaGetLastError();
}

// Pass 1: per (matrix, row block) Welford statistics per column, merged across the block's
// threads with Chan's formula.  Pass 2: merge the row blocks of the matrix, then normalise.
// partial layout: [n_mats][nblk][cols][3] = (count, mean, M2)
struct Wf {
  float n, mean, m2;
};
__device__ __forceinline__ Wf wf_merge(Wf a, Wf b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  Wf r;
  r.n = a.n + b.n;
  const float d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}

// The pivot of a row block's shifted sums: per column, the median of the block's first, middle and last row.  Any
// single row may be an outlier (a click, a padded edge); a pivot far from the data would put the whole block's
// squares at |pivot|^2 and lose the variance to cancellation.
__device__ __forceinline__ float cmvn_pivot(const float* x, int64_t r0, int64_t r1, int cols, int c) {
  const float p0 = x[r0 * cols + c], p1 = x[((r0 + r1) >> 1) * cols + c], p2 = x[(r1 - 1) * cols + c];
  return fmaxf(fminf(p0, p1), fminf(fmaxf(p0, p1), p2));
}

__global__ void __launch_bounds__(256) k_cmvn_partial(const float* __restrict__ in, float* __restrict__ partial,
                                                      int64_t rows, int cols, int64_t rows_per_blk, int nblk) {
  extern __shared__ float sm[];  // [lanes][cols][3]
  const int64_t mat = blockIdx.x;
  const int blk = blockIdx.y;
  const int lanes = blockDim.x / cols;  // row lanes; threads beyond lanes*cols idle
  const int lane = threadIdx.x / cols, c = threadIdx.x - lane * cols;
  const float* x = in + mat * rows * cols;
  const int64_t r0 = blk * rows_per_blk;
  const int64_t r1 = min(rows, r0 + rows_per_blk);
  // sums of (v - pivot) and (v - pivot)^2 with a typical row of the block (cmvn_pivot) as the pivot of every column: three
  // operations per element and no division (a per-element Welford update made this pass compute-bound);
  // the pivot keeps the squares small, and the row blocks are merged with Chan's formula as before
  float n_t = 0.f, s1 = 0.f, s2 = 0.f;
  if (lane < lanes && r0 < r1) {
    const float pivot = cmvn_pivot(x, r0, r1, cols, c);
    int64_t r = r0 + lane;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {  // four independent loads in flight
      const float v0 = __ldg(x + r * cols + c) - pivot, v1 = __ldg(x + (r + lanes) * cols + c) - pivot;
      const float v2 = __ldg(x + (r + 2 * lanes) * cols + c) - pivot, v3 = __ldg(x + (r + 3 * lanes) * cols + c) - pivot;
      s1 += (v0 + v1) + (v2 + v3);
      s2 += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
      n_t += 4.f;
    }
    for (; r < r1; r += lanes) {
      const float v = __ldg(x + r * cols + c) - pivot;
      s1 += v;
      s2 += v * v;
      n_t += 1.f;
    }
    float* s = sm + ((size_t)lane * cols + c) * 3;
    s[0] = n_t; s[1] = s1; s[2] = s2;
  } else if (lane < lanes) {
    float* s = sm + ((size_t)lane * cols + c) * 3;
    s[0] = 0.f; s[1] = 0.f; s[2] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x < cols) {
    float N = 0.f, S1 = 0.f, S2 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      const float* s = sm + ((size_t)l * cols + threadIdx.x) * 3;
      N += s[0]; S1 += s[1]; S2 += s[2];
    }
    float* o = partial + (((size_t)mat * nblk + blk) * cols + threadIdx.x) * 3;
    const float pivot = r0 < r1 ? cmvn_pivot(x, r0, r1, cols, threadIdx.x) : 0.f;
    o[0] = N;
    o[1] = N > 0.f ? pivot + S1 / N : 0.f;
    o[2] = N > 0.f ? fmaxf(S2 - S1 * (S1 / N), 0.f) : 0.f;
  }
}

// Pass 1b: one block per matrix merges the row blocks' statistics (Chan's formula, a tree over up to 256 / cols
// groups) into stats[mat][cols][2] = mean, 1 / (std + eps).  (The normalising blocks used to repeat this
// serial merge each: with hundreds of row blocks per matrix that was most of their run time.)
__global__ void __launch_bounds__(256) k_cmvn_merge(const float* __restrict__ partial, float* __restrict__ stats, int cols, int nblk,
                                                    int varnorm) {
  extern __shared__ float sm[];  // [groups][cols][3]
  const int64_t mat = blockIdx.x;
  const int groups = 256 / cols;
  const int g = threadIdx.x / cols, c = threadIdx.x - g * cols;
  if (g < groups) {
    Wf acc = {0.f, 0.f, 0.f};
    for (int b = g; b < nblk; b += groups) {
      const float* s = partial + (((size_t)mat * nblk + b) * cols + c) * 3;
      acc = wf_merge(acc, Wf{__ldg(s), __ldg(s + 1), __ldg(s + 2)});
    }
    float* s = sm + ((size_t)g * cols + c) * 3;
    s[0] = acc.n; s[1] = acc.mean; s[2] = acc.m2;
  }
  __syncthreads();
  if (threadIdx.x < cols) {
    Wf acc = {0.f, 0.f, 0.f};
    for (int l = 0; l < groups; ++l) {
      const float* s = sm + ((size_t)l * cols + threadIdx.x) * 3;
      acc = wf_merge(acc, Wf{s[0], s[1], s[2]});
    }
    float* o = stats + ((size_t)mat * cols + threadIdx.x) * 2;
    o[0] = acc.mean;
    // processing.rs:283-292: std with ddof 0, divide by (std + 2^-30)
    o[1] = varnorm ? 1.0f / (sqrtf(acc.m2 / acc.n) + kCmvnEps) : 1.0f;
  }
}

__global__ void __launch_bounds__(256) k_cmvn_apply(const float* in, float* out, const float* __restrict__ stats, int64_t rows, int cols,
                                                    int64_t rows_per_blk, int varnorm) {
  extern __shared__ float sm[];  // [cols][2] = mean, 1/(std+eps)
  const int64_t mat = blockIdx.x;
  const int blk = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x) sm[i] = __ldg(stats + (size_t)mat * cols * 2 + i);
  __syncthreads();
  const float* x = in + mat * rows * cols;
  float* y = out + mat * rows * cols;
  const int64_t e0 = blk * rows_per_blk * cols;
  const int64_t e1 = min(rows, (blk + 1) * rows_per_blk) * cols;
  x += e0;
  y += e0;
  const int n = (int)(e1 - e0);  // rows_per_blk * cols < 2^31 (launch_cmvn)
  int c = (int)((e0 + threadIdx.x) % cols);
  const int step = (int)(blockDim.x % cols);
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = x[i] - sm[2 * c];  // plain load: the host layer runs this pass in place (in == out)
    y[i] = varnorm ? d * sm[2 * c + 1] : d;
    c += step;
    if (c >= cols) c -= cols;
  }
}

// Long matrices (hours of cepstra per matrix, e.g. 360,000 x 13 = 18.7 MB): one persistent, co-resident grid
// (cooperative launch) pipelines the matrices through the L2 so that each crosses the DRAM bus once per direction
// instead of twice in.  Worker blocks own a fixed row range: they sum it for matrix i (shifted sums, as
// k_cmvn_partial) and announce it on the matrix's counter, then normalise the same rows of matrix i - lag (lag = 1) out of
// L2 once its statistics are published.  A few merger blocks (matrix m belongs to merger m mod kCoopMergers) wait for
// a matrix's counter, merge the workers' partials in block order (Chan's formula; the bits do not depend on the
// batch) and publish mean and 1 / (std + eps) as tagged 8-byte records.  No grid-wide barrier.  Measured on
// 64 x (359,997 x 13): DRAM reads 1.21 GB = the input once (three-pass kernels: 2.33 GB), 0.71 ms against 0.89 ms.  What is left is not bandwidth and not bytes in flight: every block walks every matrix, so a
// matrix costs one block's chain of dependent round trips (DRAM batches, three block reductions, fence + counter,
// flag, L2 batches: ~11 us).  Two matrices per step (G = 2) buy 1 % and lose the L2 reuse; four spill; 16-byte vectors (one trip per
// matrix, 128 bytes per thread in flight) measured 0.82 ms; prefetching the next matrix's first batch lost to register
// spills, and letting every thread poll the statistics records to 150,000 threads polling the L2.
constexpr int kCoopThreads = 512, kCoopBlocksPerSm = 2, kCoopMergers = 4, kCoopGroupMax = 4;
constexpr int64_t kCoopMinMatBytes = 3LL << 20, kCoopMaxMatBytes = 40LL << 20, kCoopL2Bytes = 64LL << 20;
struct CoopArgs {
  const float* in;
  float* out;
  float* partial;   // [ring][workers][cols][3]
  unsigned long long* stats;  // [n_mats][cols][2] records: bits of mean / 1 / (std + eps) | (m + 1) << 32 -- each 8-byte record
                              // carries its own "published" tag (zeroed before the launch)
  int* counter;     // [n_mats] workers done with the sums of matrix m   (zeroed before the launch)
  int64_t n_mats, rows;
  int cols, lag, ring, varnorm;
};
__device__ __forceinline__ void spin_until(const int* p, int want) {
  while (*reinterpret_cast<const volatile int*>(p) < want) __nanosleep(100);
}
// the merger blocks of the cooperative kernels: matrix m belongs to merger m mod kCoopMergers
__device__ __forceinline__ void coop_merger(const CoopArgs& a, float* red, float* red2, int workers) {
  const int cols = a.cols;
  const int lanes = kCoopThreads / cols;
  const int lane = threadIdx.x / cols, c = threadIdx.x - lane * cols;
    const int me = blockIdx.x - workers;
    for (int64_t m = me; m < a.n_mats; m += kCoopMergers) {
      if (threadIdx.x == 0) spin_until(a.counter + m, workers);
      __syncthreads();
      __threadfence();
      const float* pbuf = a.partial + (size_t)(m % a.ring) * workers * cols * 3;
      if (lane < lanes) {
        Wf acc = {0.f, 0.f, 0.f};
        // eight partials (24 clamped, unconditional L2 loads) per round trip, merged in block order: with ~300 workers
        // and 39 lanes that is one trip per merge -- the merge sits on the workers' path when the lag is one matrix
        for (int b0 = lane; b0 < workers; b0 += 8 * lanes) {
          Wf w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int bb = b0 + q * lanes;
            const float* s = pbuf + ((size_t)min(bb, workers - 1) * cols + c) * 3;
            w[q] = Wf{bb < workers ? __ldcg(s) : 0.f, __ldcg(s + 1), __ldcg(s + 2)};
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) acc = wf_merge(acc, w[q]);  // n == 0: skipped
        }
        float* s = red + ((size_t)lane * cols + c) * 3;
        s[0] = acc.n; s[1] = acc.mean; s[2] = acc.m2;
      }
      __syncthreads();
      if (threadIdx.x < 8 * cols) {  // 8 column-parallel strands over the lanes, then one thread per column
        const int part = threadIdx.x / cols, cc = threadIdx.x - part * cols;
        Wf acc = {0.f, 0.f, 0.f};
        for (int l = part; l < lanes; l += 8) {
          const float* s = red + ((size_t)l * cols + cc) * 3;
          acc = wf_merge(acc, Wf{s[0], s[1], s[2]});
        }
        float* s = red2 + ((size_t)part * cols + cc) * 3;
        s[0] = acc.n; s[1] = acc.mean; s[2] = acc.m2;
      }
      __syncthreads();
      if (threadIdx.x < cols) {
        Wf acc = {0.f, 0.f, 0.f};
        for (int l = 0; l < 8; ++l) {
          const float* s = red2 + ((size_t)l * cols + threadIdx.x) * 3;
          acc = wf_merge(acc, Wf{s[0], s[1], s[2]});
        }
        // processing.rs:283-292: std with ddof 0, divide by (std + 2^-30)
        const float inv = a.varnorm ? 1.0f / (sqrtf(acc.m2 / acc.n) + kCmvnEps) : 1.0f;
        // value and tag land together (aligned 8-byte stores): no flag, no fence, and the workers' 2 * cols polling
        // threads get the statistics with the round trip that tells them they are there
        volatile unsigned long long* o = a.stats + ((size_t)m * cols + threadIdx.x) * 2;
        const unsigned long long tag = (unsigned long long)(unsigned)(m + 1) << 32;
        o[0] = tag | __float_as_uint(acc.mean);
        o[1] = tag | __float_as_uint(inv);
      }
      __syncthreads();  // red / red2 are rewritten for the next matrix
    }
}

// G: matrices per pipeline step.  A step's loads for all G matrices are issued together, so the chain of dependent round
// trips (the pipeline's clock, below) is paid once per G matrices.
template <int G>
__global__ void __launch_bounds__(kCoopThreads, kCoopBlocksPerSm) k_cmvn_coop(const CoopArgs a) {
  constexpr int kCoopBatch = G == 1 ? 16 : 8;  // loads per thread, matrix and round trip
  extern __shared__ float sm[];  // G x ([lanes][cols][3], then [8][cols][3])
  const int cols = a.cols;
  const int lanes = kCoopThreads / cols;
  const int lane = threadIdx.x / cols, c = threadIdx.x - lane * cols;
  float* red = sm;
  float* red2 = sm + (size_t)G * lanes * cols * 3;
  const int workers = gridDim.x - kCoopMergers;
  if ((int)blockIdx.x >= workers) {
    coop_merger(a, red, red2, workers);
    return;
  }
  // ---- worker: rows [r0, r1) of every matrix.  The first lanes * cols threads walk the block's n = (r1 - r0) * cols
  // elements with a stride of lanes * cols, a multiple of cols: a thread stays in one column, its statistics sit in
  // registers and an element costs one 32-bit multiply-add of index arithmetic.  A block walks every matrix, so its
  // per-matrix chain of dependent memory round trips is the pipeline's clock: the loads of a batch (kCoopBatch per
  // trip, clamped and unconditional so that they issue back to back) and the pivot rows go out before anything waits.
  const int64_t rpb = (a.rows + workers - 1) / workers;
  const int64_t r0 = min(a.rows, blockIdx.x * rpb), r1 = min(a.rows, r0 + rpb);
  const int n = (int)((r1 - r0) * cols);  // rpb * cols < 2^31 (launch_cmvn)
  const int stride = lanes * cols, t = threadIdx.x;
  const bool act = t < stride && n > 0;
  const int mid = (int)((r1 - r0) >> 1) * cols + c, last = n - cols + c;  // cmvn_pivot's rows, block-relative
  const int64_t n_steps = (a.n_mats + G - 1) / G;
  const int red_sz = stride * 3, red2_sz = 8 * cols * 3;
  for (int64_t st = 0; st < n_steps + a.lag; ++st) {
    if (st < n_steps) {
      const int64_t m0 = st * G;
      const int g_n = (int)min((int64_t)G, a.n_mats - m0);
      const float* x0 = a.in + (m0 * a.rows + r0) * cols;
      const int64_t mstride = a.rows * cols;  // matrix to matrix
      float s1[G], s2[G], pivot[G];
#pragma unroll
      for (int g = 0; g < G; ++g) { s1[g] = 0.f; s2[g] = 0.f; pivot[g] = 0.f; }
      float n_t = 0.f;
      if (act) {
        float p0[G], p1[G], p2[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float* x = x0 + (g < g_n ? g : 0) * mstride;  // a missing matrix of the last step re-reads the first one
          p0[g] = x[c]; p1[g] = x[mid]; p2[g] = x[last];
        }
        for (int k = t; k < n; k += kCoopBatch * stride) {
          float v[G][kCoopBatch];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float* x = x0 + (g < g_n ? g : 0) * mstride;
#pragma unroll
            for (int q = 0; q < kCoopBatch; ++q) v[g][q] = x[min(k + q * stride, n - 1)];
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            pivot[g] = fmaxf(fminf(p0[g], p1[g]), fminf(fmaxf(p0[g], p1[g]), p2[g]));  // cmvn_pivot
#pragma unroll
            for (int q = 0; q < kCoopBatch; ++q) v[g][q] = k + q * stride < n ? v[g][q] - pivot[g] : 0.f;
#pragma unroll
            for (int q = 0; q < kCoopBatch; q += 4) {
              s1[g] += (v[g][q] + v[g][q + 1]) + (v[g][q + 2] + v[g][q + 3]);
              s2[g] += (v[g][q] * v[g][q] + v[g][q + 1] * v[g][q + 1]) + (v[g][q + 2] * v[g][q + 2] + v[g][q + 3] * v[g][q + 3]);
            }
          }
        }
        n_t = (float)((n - t + stride - 1) / stride);  // elements of this thread
      }
      if (t < stride) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float* s = red + g * red_sz + (size_t)t * 3;
          s[0] = n_t; s[1] = s1[g]; s[2] = s2[g];
        }
      }
      __syncthreads();
      if (t < 8 * cols) {
        const int part = t / cols, cc = t - part * cols;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float N = 0.f, S1 = 0.f, S2 = 0.f;
          for (int l = part; l < lanes; l += 8) {
            const float* s = red + g * red_sz + ((size_t)l * cols + cc) * 3;
            N += s[0]; S1 += s[1]; S2 += s[2];
          }
          float* s = red2 + g * red2_sz + ((size_t)part * cols + cc) * 3;
          s[0] = N; s[1] = S1; s[2] = S2;
        }
      }
      __syncthreads();
      if (t < cols) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < g_n) {
            float N = 0.f, S1 = 0.f, S2 = 0.f;
            for (int l = 0; l < 8; ++l) {
              const float* s = red2 + g * red2_sz + ((size_t)l * cols + t) * 3;
              N += s[0]; S1 += s[1]; S2 += s[2];
            }
            float* o = a.partial + (((size_t)((m0 + g) % a.ring) * workers + blockIdx.x) * cols + t) * 3;
            o[0] = N;  // t < cols: this thread's pivot is column t's
            o[1] = N > 0.f ? pivot[g] + S1 / N : 0.f;
            o[2] = N > 0.f ? fmaxf(S2 - S1 * (S1 / N), 0.f) : 0.f;
          }
        }
        __threadfence();
      }
      __syncthreads();
      if (t < g_n) atomicAdd(a.counter + m0 + t, 1);
    }
    if (st >= a.lag) {
      const int64_t m0 = (st - a.lag) * G;
      const int g_n = (int)min((int64_t)G, a.n_mats - m0);
      if (t < 2 * cols) {  // 2 * cols threads per block poll the tagged records and park the values in shared memory
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < g_n) {
            const volatile unsigned long long* rp = a.stats + (size_t)(m0 + g) * cols * 2 + t;
            unsigned long long r = *rp;
            while ((unsigned)(r >> 32) != (unsigned)(m0 + g + 1)) {
              __nanosleep(100);
              r = *rp;
            }
            red2[g * 2 * cols + t] = __uint_as_float((unsigned)r);
          }
        }
      }
      __syncthreads();
      float mean[G], inv[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int gg = g < g_n ? g : 0;
        mean[g] = red2[gg * 2 * cols + 2 * c];
        inv[g] = red2[gg * 2 * cols + 2 * c + 1];  // 1 when !varnorm
      }
      __syncthreads();  // everyone has its statistics: red2 may be rewritten (in the drain steps by the next poll right away)
      if (act) {
        const float* x0 = a.in + (m0 * a.rows + r0) * cols;
        float* y0 = a.out + (m0 * a.rows + r0) * cols;
        const int64_t mstride = a.rows * cols;
        for (int k = t; k < n; k += kCoopBatch * stride) {  // one round trip (to L2) per batch, as above
          float v[G][kCoopBatch];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float* x = x0 + (g < g_n ? g : 0) * mstride;
#pragma unroll
            for (int q = 0; q < kCoopBatch; ++q) v[g][q] = __ldcs(x + min(k + q * stride, n - 1));  // last use of the line: evict first
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (g < g_n) {
              float* y = y0 + g * mstride;
#pragma unroll
              for (int q = 0; q < kCoopBatch; ++q) {
                const float d = v[g][q] - mean[g];
                if (k + q * stride < n) __stcs(y + k + q * stride, a.varnorm ? d * inv[g] : d);  // streaming: the L2 belongs to the sums' reads
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ derivative features
// processing::derivative_extraction (processing.rs:222-254) with numpy 'edge' padding along the
// FEATURE axis (the reference's util::pad edge mode is degenerate):
//   out[r][c] = (sum_{q=1..dw} (F[r][clamp(c+q)] * q - F[r][clamp(c-q)])) / (2 sum q^2)
// Unfused multiplies / adds in the reference's order, so results are bit-identical to it.
__device__ __forceinline__ float delta_at(const float* __restrict__ row, int cols, int c, int dw, float scale) {
  float dif = 0.f;
  for (int q = 1; q <= dw; ++q) {
    const float a = __ldg(row + min(c + q, cols - 1)), b = __ldg(row + max(c - q, 0));
    dif = __fadd_rn(dif, __fsub_rn(__fmul_rn(a, (float)q), b));
  }
  return __fdiv_rn(dif, scale);
}

__global__ void __launch_bounds__(256) k_derivative(const float* __restrict__ in, float* __restrict__ out,
                                                    int64_t rows, int cols, int dw, float scale) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    out[i] = delta_at(in + r * cols, cols, c, dw, scale);
  }
}

// feature::extract_derivative_feature (feature.rs:253-269): [rows][cols] -> [rows][cols][3] =
// (static, delta, delta-delta) with DeltaWindows = 2; delta-delta is the delta of the delta row.
__global__ void __launch_bounds__(256) k_derivative_cube(const float* __restrict__ in, float* __restrict__ out,
                                                         int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float* row = in + r * cols;
    float d[5];  // delta at clamp(c-2) .. clamp(c+2)
#pragma unroll
    for (int k = -2; k <= 2; ++k) d[k + 2] = delta_at(row, cols, min(max(c + k, 0), cols - 1), 2, 10.0f);
    float dd = 0.f;
    dd = __fadd_rn(dd, __fsub_rn(__fmul_rn(d[3], 1.0f), d[1]));
    dd = __fadd_rn(dd, __fsub_rn(__fmul_rn(d[4], 2.0f), d[0]));
    out[i * 3] = __ldg(row + c);
    out[i * 3 + 1] = d[2];
    out[i * 3 + 2] = __fdiv_rn(dd, 10.0f);
  }
}

// ------------------------------------------------------------------------- cmvnw
// processing::cmvnw (processing.rs:315-371) with numpy 'symmetric' padding along time.  One thread
// owns (column, chunk of rows): it sums its first window directly, then slides.  Sums are kept in
// double so the result does not depend on the chunking.  Pass 1: y = x - window mean; pass 2
// (variance normalisation): y / (window std of the symmetric-padded y + 2^-30).
__device__ __forceinline__ int64_t reflect_sym(int64_t i, int64_t n) {
  const int64_t p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}

template <bool VAR>
__global__ void __launch_bounds__(256) k_cmvnw(const float* __restrict__ in, float* __restrict__ out, int64_t rows,
                                               int cols, int win, int chunk) {
  const int64_t mat = blockIdx.y;
  const int colsp = blockDim.x >= 32 ? 1 : 1;
  (void)colsp;
  const int c = threadIdx.x % cols;          // threads beyond a multiple of cols idle
  const int lane = threadIdx.x / cols;
  const int lanes = blockDim.x / cols;
  if (lane >= lanes) return;
  const float* x = in + mat * rows * cols;
  float* y = out + mat * rows * cols;
  const int64_t r0 = ((int64_t)blockIdx.x * lanes + lane) * chunk;
  if (r0 >= rows) return;
  const int64_t r1 = min(rows, r0 + chunk);
  const int p = (win - 1) / 2;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t j = r0 - p; j <= r0 + p; ++j) {
    const double v = (double)__ldg(x + reflect_sym(j, rows) * cols + c);
    s1 += v;
    if (VAR) s2 += v * v;
  }
  for (int64_t r = r0; r < r1; ++r) {
    const float v = __ldg(x + r * cols + c);
    if (!VAR) {
      y[r * cols + c] = v - (float)(s1 / (double)win);
    } else {
      const double m = s1 / (double)win;
      const double var = fmax(s2 / (double)win - m * m, 0.0);
      y[r * cols + c] = v / ((float)sqrt(var) + kCmvnEps);
    }
    const double a = (double)__ldg(x + reflect_sym(r + p + 1, rows) * cols + c);
    const double b = (double)__ldg(x + reflect_sym(r - p, rows) * cols + c);
    s1 += a - b;
    if (VAR) s2 += a * a - b * b;
  }
}

inline int64_t cmvn_rows_per_blk(int64_t n_mats, int64_t rows) {
  // many small blocks: at least ~4 waves of 8 x 148 resident blocks (a grid of 1.03 waves ran for two), at most
  // 2048 rows per block, at least 64
  const int64_t want_blocks = 4 * 8 * kSmCountGuess;
  int64_t rpb = (rows * n_mats + want_blocks - 1) / want_blocks;
  rpb = std::min<int64_t>(rpb, 2048);
  rpb = std::max<int64_t>(rpb, 64);
  rpb = std::max<int64_t>(rpb, (rows + 65534) / 65535);  // gridDim.y
  if (rpb > rows) rpb = rows;
  return rpb < 1 ? 1 : rpb;
}

}  // namespace

int launch_preemphasis(const float* in, float* out, int64_t n_rows, int64_t n, int64_t row_stride,
                       int64_t shift, float cof, void* stream) {
  if (n_rows <= 0 || n <= 0) return 0;
  int64_t sm = shift % n;
  if (sm < 0) sm += n;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0) &&
                     (row_stride % 4 == 0);
  int64_t bx = (n / 4 + 255) / 256;
  const int64_t cap = (8LL * kSmCountGuess + n_rows - 1) / n_rows;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  if (bx * n_rows > 2147483647LL) return (int)cudaErrorInvalidValue;
  k_preemphasis<<<(unsigned)(bx * n_rows), 256, 0, (cudaStream_t)stream>>>(in, out, n_rows, n, row_stride, sm,
                                                                           cof, vec_ok, (int)bx);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_stack_frames(const float* in, int64_t n, int64_t L, int64_t S, int64_t T, uint32_t flags,
                        float* out, void* stream) {
  if (T <= 0) return 0;
  const bool v4 = !(flags & SS_FLAG_FRAMING_LITERAL) && L % 4 == 0 && S % 4 == 0 && n % 4 == 0 && L / 4 < (1 << 20) && S / 4 < (1 << 20) &&
                  reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  if (v4) {
    const int64_t blocks = std::min<int64_t>((T + kSfFrames - 1) / kSfFrames, 32LL * kSmCountGuess);
    k_stack_frames_v4<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                                        n / 4, (int)(L / 4), (int)(S / 4), T);
    count_launch();
    return (int)cudaGetLastError();
  }
  int64_t blocks = (T * L + 255) / 256;
  if (blocks > 16LL * kSmCountGuess) blocks = 16LL * kSmCountGuess;
  k_stack_frames<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n, L, S, T, flags);
  count_launch();
  return (int)cudaGetLastError();
}

// ---- processing::log_power_spectrum (processing.rs:194-214) on an already computed power spectrum:
// x > 1e-20 ? 10 log10 x : -200, and (quirk Q7) the normaliser is the LAST value above the threshold in
// row-major order, not the maximum.  scratch[0] = 1 + index of that element (0: none), scratch[1] = bits of mx.
__global__ void k_logps_map(float* __restrict__ x, int64_t n, unsigned long long* __restrict__ scratch) {
  unsigned long long last = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (v > 1e-20f) {
      x[i] = 10.0f * log10f(v);
      last = (unsigned long long)i + 1;  // i grows along the grid-stride loop
    } else {
      x[i] = -200.0f;  // 10 log10(1e-20)
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if ((threadIdx.x & 31) == 0 && last) atomicMax(scratch, last);
}
__global__ void k_logps_pick(const float* __restrict__ x, unsigned long long* __restrict__ scratch) {
  const float mx = scratch[0] ? x[scratch[0] - 1] : 1e-20f;  // `let mut mx = 1e-20_f32`
  scratch[1] = (unsigned long long)__float_as_uint(mx);
}
__global__ void k_logps_sub(float* __restrict__ x, int64_t n, const unsigned long long* __restrict__ scratch) {
  const float mx = __uint_as_float((unsigned)scratch[1]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = x[i] - mx;
}

// scratch: 16 bytes of device memory
int launch_log_power(float* x, int64_t n, int normalize, void* scratch, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* sc = static_cast<unsigned long long*>(scratch);
  cudaError_t e = cudaMemsetAsync(sc, 0, 16, st);
  if (e != cudaSuccess) return (int)e;
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, 16LL * kSmCountGuess);
  k_logps_map<<<(unsigned)blocks, 256, 0, st>>>(x, n, sc);
  count_launch();
  if (normalize) {
    k_logps_pick<<<1, 1, 0, st>>>(x, sc);
    k_logps_sub<<<(unsigned)blocks, 256, 0, st>>>(x, n, sc);
    count_launch(2);
  }
  return (int)cudaGetLastError();
}

int launch_derivative(const float* in, int64_t rows, int64_t cols, int dw, float* out, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  float scale = 0.f;
  for (int q = 1; q <= dw; ++q) scale += 2.0f * (float)(q * q);  // processing.rs:249
  int64_t blocks = std::min<int64_t>((rows * cols + 255) / 256, 16LL * kSmCountGuess);
  k_derivative<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, rows, (int)cols, dw, scale);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_derivative_cube(const float* in, int64_t rows, int64_t cols, float* out, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  int64_t blocks = std::min<int64_t>((rows * cols + 255) / 256, 16LL * kSmCountGuess);
  k_derivative_cube<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, rows, (int)cols);
  count_launch();
  return (int)cudaGetLastError();
}

// scratch: n_mats * rows * cols floats when variance_normalization (the mean-subtracted matrix)
int launch_cmvnw(const float* in, int64_t n_mats, int64_t rows, int64_t cols, int win, int varnorm, float* out,
                 float* scratch, void* stream) {
  if (n_mats <= 0 || rows <= 0 || cols <= 0) return 0;
  if (cols > 256 || n_mats > 65535) return (int)cudaErrorInvalidValue;
  const int lanes = 256 / (int)cols;
  const int chunk = 256;
  const int64_t nchunks = (rows + chunk - 1) / chunk;
  dim3 grid((unsigned)((nchunks + lanes - 1) / lanes), (unsigned)n_mats);
  float* y = varnorm ? scratch : out;
  k_cmvnw<false><<<grid, 256, 0, (cudaStream_t)stream>>>(in, y, rows, (int)cols, win, chunk);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !varnorm) return (int)e;
  k_cmvnw<true><<<grid, 256, 0, (cudaStream_t)stream>>>(y, out, rows, (int)cols, win, chunk);
  count_launch();
  return (int)cudaGetLastError();
}

// the cooperative long-matrix kernel: how many matrices the sums run ahead of the normalisation (0: shape declined)
static int cmvn_coop_lag(int64_t rows, int64_t cols) {
  static const bool off = getenv("SS_CMVN_COOP") && atoi(getenv("SS_CMVN_COOP")) == 0;  // A/B switch
  const int64_t mat_bytes = rows * cols * (int64_t)sizeof(float);
  if (off || cols > 64 || mat_bytes < kCoopMinMatBytes || mat_bytes > kCoopMaxMatBytes) return 0;
  static const int lag_env = getenv("SS_CMVN_LAG") ? atoi(getenv("SS_CMVN_LAG")) : 0;  // measurements
  if (lag_env > 0) return std::min(lag_env, 8);
  // the sums run as few matrices ahead as the merge latency allows: each one in between has to stay in the L2 next to the
  // output stream (18.7 MB matrices: one ahead re-reads nothing from DRAM, two ahead a third of the matrix)
  return (int)std::min<int64_t>(8, std::max<int64_t>(1, kCoopL2Bytes / mat_bytes - 2));
}
// matrices per pipeline step and the lag in steps
static void cmvn_coop_steps(int lag_mats, int* group, int* lag_steps) {
  static const int g_env = getenv("SS_CMVN_GROUP") ? atoi(getenv("SS_CMVN_GROUP")) : 0;  // measurements
  // default 1: two matrices per step are 1 % faster (0.746 against 0.753 ms) but put four matrices between the sums and
  // the normalisation, which no longer fit the L2 next to the output stream: DRAM reads 2.23 GB instead of 1.20 GB
  const int g = g_env == 1 || g_env == 2 || g_env == 4 ? g_env : 1;
  *group = g;
  *lag_steps = std::max(1, (lag_mats + 1) / g - 1);
}
static size_t coop_align(size_t b) { return (b + 255) & ~(size_t)255; }

size_t cmvn_scratch_bytes(int64_t n_mats, int64_t rows, int64_t cols) {
  const int64_t rpb = cmvn_rows_per_blk(n_mats, rows);
  const int64_t nblk = (rows + rpb - 1) / rpb;
  // + 2 floats per matrix and column for the merged statistics
  const size_t three_pass = (size_t)(n_mats * nblk + n_mats) * cols * 3 * sizeof(float);
  const int lag = cmvn_coop_lag(rows, cols);
  if (!lag) return three_pass;
  // partial ring + statistics + counters and flags (launch_cmvn lays them out in this order)
  const int64_t coop_blocks = (int64_t)kCoopBlocksPerSm * kSmCountGuess;
  const size_t coop = coop_align((size_t)(lag + 1) * kCoopGroupMax * coop_blocks * cols * 3 * sizeof(float)) + coop_align((size_t)n_mats * cols * 2 * sizeof(unsigned long long)) +
                      coop_align((size_t)n_mats * sizeof(int));
  return std::max(three_pass, coop);
}

int launch_cmvn(const float* in, int64_t n_mats, int64_t rows, int64_t cols, int varnorm, float* out,
                float* scratch, void* stream) {
  if (n_mats <= 0 || rows <= 0 || cols <= 0) return 0;
  if (cols > 256 || n_mats > 2147483647LL) return (int)cudaErrorInvalidValue;
  // matrices that fit a warp's slice of shared memory: the fused one-read / one-write kernel
  // (chosen by the matrix shape alone, so that a batch gives the same bits however the host layer chunks it)
  static const bool no_reg = getenv("SS_CMVN_REG") && atoi(getenv("SS_CMVN_REG")) == 0;  // A/B switch
  if (cols <= 32 && rows * cols <= kCmvnRegJ * ((32 / cols) * cols) && !no_reg) {
    switch (cols) {  // the usual cepstral counts; other widths take the shared-memory kernel below
      case 12: return launch_cmvn_reg<12>(in, out, n_mats, rows, varnorm, stream);
      case 13: return launch_cmvn_reg<13>(in, out, n_mats, rows, varnorm, stream);
      case 20: return launch_cmvn_reg<20>(in, out, n_mats, rows, varnorm, stream);
      default: break;
    }
  }
  if (rows * cols <= 5120) {
    const int n = (int)(rows * cols), R = cols <= 16 ? 32 / (int)cols : 1;
    const int warp_floats = ((n + 3) & ~3) + (R + 2) * (int)cols + 4;
    const size_t smem = (size_t)kCmvnSmallWarps * warp_floats * sizeof(float);
    if (smem > 48 * 1024) {  // per device, so not cached in a static
      cudaError_t ea = cudaFuncSetAttribute(k_cmvn_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (ea != cudaSuccess) return (int)ea;
    }
    // more blocks than resident slots: the hardware scheduler balances the tail (exactly one wave measured 11 % slower)
    const int64_t blocks = std::min<int64_t>((n_mats + kCmvnSmallWarps - 1) / kCmvnSmallWarps, 8LL * kSmCountGuess);
    k_cmvn_small<<<(unsigned)blocks, kCmvnSmallWarps * 32, smem, (cudaStream_t)stream>>>(in, out, n_mats, (int)rows, (int)cols, varnorm,
                                                                                         warp_floats);
    count_launch();
    return (int)cudaGetLastError();
  }
  if (const int lag = cmvn_coop_lag(rows, cols)) {
    // long matrices: the cooperative one-DRAM-read pipeline, when the device can hold its whole grid
    int dev = 0, sms = 0, coop_ok = 0, per_sm = 0;
    const int lanes = kCoopThreads / (int)cols;
    int G = 1, lag_steps = 1;
    cmvn_coop_steps(lag, &G, &lag_steps);
    const size_t smem = (size_t)G * ((size_t)lanes * cols * 3 + (size_t)8 * cols * 3) * sizeof(float);
    const void* kern = G == 1 ? (const void*)k_cmvn_coop<1> : G == 2 ? (const void*)k_cmvn_coop<2> : (const void*)k_cmvn_coop<4>;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, dev);
    if (e == cudaSuccess && smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCoopThreads, smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t nblk = std::min<int64_t>((int64_t)std::min(per_sm, kCoopBlocksPerSm) * sms, (int64_t)kCoopBlocksPerSm * kSmCountGuess);
    const int64_t workers = nblk - kCoopMergers;
    if (coop_ok && workers >= 32 && ((rows + workers - 1) / workers) * cols < (1LL << 31)) {
      CoopArgs a;
      a.in = in;
      a.out = out;
      a.n_mats = n_mats;
      a.rows = rows;
      a.cols = (int)cols;
      a.lag = lag_steps;
      a.ring = (lag_steps + 1) * G;  // <= (lag + 1) * kCoopGroupMax slots (cmvn_scratch_bytes)
      a.varnorm = varnorm;
      char* sp = reinterpret_cast<char*>(scratch);
      a.partial = reinterpret_cast<float*>(sp);
      sp += coop_align((size_t)(lag + 1) * kCoopGroupMax * kCoopBlocksPerSm * kSmCountGuess * cols * 3 * sizeof(float));
      a.stats = reinterpret_cast<unsigned long long*>(sp);
      const size_t rec_bytes = coop_align((size_t)n_mats * cols * 2 * sizeof(unsigned long long));
      sp += rec_bytes;
      a.counter = reinterpret_cast<int*>(sp);
      e = cudaMemsetAsync(a.stats, 0, rec_bytes + (size_t)n_mats * sizeof(int), (cudaStream_t)stream);
      if (e != cudaSuccess) return (int)e;
      void* args[] = {(void*)&a};
      e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)nblk), dim3(kCoopThreads), args, smem, (cudaStream_t)stream);
      count_launch();
      return (int)e;
    }
  }
  const int64_t rpb = cmvn_rows_per_blk(n_mats, rows);
  const int nblk = (int)((rows + rpb - 1) / rpb);
  if (nblk > 65535) return (int)cudaErrorInvalidValue;
  dim3 grid((unsigned)n_mats, (unsigned)nblk);
  const int lanes = 256 / (int)cols;
  k_cmvn_partial<<<grid, 256, (size_t)lanes * cols * 3 * sizeof(float), (cudaStream_t)stream>>>(
      in, scratch, rows, (int)cols, rpb, nblk);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  // the merged statistics live behind the partials in the scratch buffer
  float* stats = scratch + (size_t)n_mats * nblk * cols * 3;
  k_cmvn_merge<<<(unsigned)n_mats, 256, (size_t)lanes * cols * 3 * sizeof(float), (cudaStream_t)stream>>>(scratch, stats, (int)cols, nblk, varnorm);
  count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  if (rpb * cols >= (1LL << 31)) return (int)cudaErrorInvalidValue;
  k_cmvn_apply<<<grid, 256, (size_t)cols * 2 * sizeof(float), (cudaStream_t)stream>>>(in, out, stats, rows, (int)cols, rpb, varnorm);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace ss
