// Micro-benchmark: issue rate of the sm_100 packed FP32 instructions (FFMA2 / FADD2, PTX
// fma.rn.f32x2 / add.rn.f32x2) next to scalar FFMA, alone and interleaved with integer work,
// to decide whether the complex butterflies of the fused MFCC kernel should use them.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float2* out, int iters, float2 w) {
  float2 d[8];
  int acc[8];
  for (int c = 0; c < 8; ++c) { d[c] = make_float2(c + threadIdx.x, c * 0.5f); acc[c] = c + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (MODE == 0) { d[c].x = fmaf(d[c].x, w.x, w.y); d[c].y = fmaf(d[c].y, w.x, w.y); }            // 2 FFMA
      if (MODE == 1) d[c] = __ffma2_rn(d[c], w, make_float2(w.y, w.x));                                 // 1 FFMA2
      if (MODE == 2) { d[c] = __ffma2_rn(d[c], w, make_float2(w.y, w.x)); acc[c] = acc[c] * 3 + it; }   // FFMA2 + IMAD
      if (MODE == 3) { d[c].x = fmaf(d[c].x, w.x, w.y); d[c].y = fmaf(d[c].y, w.x, w.y); acc[c] = acc[c] * 3 + it; }  // 2 FFMA + IMAD
      if (MODE == 4) d[c] = __fadd2_rn(d[c], w);                                                         // 1 FADD2
      if (MODE == 5) { d[c] = __ffma2_rn(d[c], w, make_float2(w.y, w.x)); acc[c] = (acc[c] ^ it) + (acc[c] >> 3); }  // FFMA2 + 2-3 ALU
    }
  }
  float2 s = make_float2(0.f, 0.f);
  int si = 0;
  for (int c = 0; c < 8; ++c) { s.x += d[c].x; s.y += d[c].y; si += acc[c]; }
  s.x += (float)si;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_it(K kern, int blocks, int threads, float2* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, iters, make_float2(0.999f, 1e-3f));
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(out, iters, make_float2(0.999f, 1e-3f));
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = pr.multiProcessorCount;
  float2* out; cudaMalloc(&out, sizeof(float2) * sms * 1024);
  const int iters = 20000;
  const double clk = clk_khz * 1e3;
  const char* names[] = {"2xFFMA", "FFMA2", "FFMA2+IMAD", "2xFFMA+IMAD", "FADD2", "FFMA2+3ALU"};
  for (int warps : {8, 16, 32}) {
    float ms[6];
    ms[0] = time_it(k<0>, sms, warps * 32, out, iters);
    ms[1] = time_it(k<1>, sms, warps * 32, out, iters);
    ms[2] = time_it(k<2>, sms, warps * 32, out, iters);
    ms[3] = time_it(k<3>, sms, warps * 32, out, iters);
    ms[4] = time_it(k<4>, sms, warps * 32, out, iters);
    ms[5] = time_it(k<5>, sms, warps * 32, out, iters);
    for (int m = 0; m < 6; ++m)
      printf("warps/SM %2d  %-12s  %.3f clk per (warp, complex element) per SM-quarter -> %.2f complex FMA/clk/SM\n", warps,
             names[m], ms[m] * 1e-3 * clk / (iters * 8.0 * warps / 4.0), warps * 8.0 * iters / (ms[m] * 1e-3 * clk));
  }
  return 0;
}
