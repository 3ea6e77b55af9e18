// Micro-benchmark: issue rate of mma.sync.m16n8k8 TF32 (legacy tensor path, HMMA) on sm_100a,
// next to FFMA, to decide whether the banded mel / DCT stages of the fused MFCC kernel can ride on it.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_tf32_rate mma_tf32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void k_mma(float* out, int iters) {
  float d[CHAINS][4];
  unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 5u, 7u}, b[2] = {threadIdx.x, 11u};
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma_tf32(d[c], a, b);
  }
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_ffma(float* out, int iters) {
  float d[CHAINS];
  float a = threadIdx.x * 1e-3f, b = 0.999f;
  for (int c = 0; c < CHAINS; ++c) d[c] = c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) d[c] = fmaf(d[c], b, a);
  }
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) s += d[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_it(K k, int blocks, int threads, float* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(out, iters);
  cudaEventRecord(e0);
  k<<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = pr.multiProcessorCount;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
  const int iters = 20000;
  printf("%s, %d SMs, %d kHz\n", pr.name, sms, clk_khz);
  for (int warps : {4, 8, 16, 32}) {
    float ms1 = time_it(k_mma<1>, sms, warps * 32, out, iters);
    float ms4 = time_it(k_mma<4>, sms, warps * 32, out, iters);
    float ms8 = time_it(k_mma<8>, sms, warps * 32, out, iters);
    double clk = clk_khz * 1e3;
    // MMAs per SM per clock
    printf("warps/SM %2d: mma/clk/SM chains1 %.3f chains4 %.3f chains8 %.3f  (dep latency ~%.1f clk)\n", warps,
           warps * (double)iters / (ms1 * 1e-3 * clk), warps * 4.0 * iters / (ms4 * 1e-3 * clk),
           warps * 8.0 * iters / (ms8 * 1e-3 * clk), ms1 * 1e-3 * clk / iters);
    float f8 = time_it(k_ffma<8>, sms, warps * 32, out, iters);
    printf("             ffma warp-instr/clk/SM chains8 %.3f\n", warps * 8.0 * iters / (f8 * 1e-3 * clk));
  }
  // TF32 dense TFLOP/s implied: mma/clk/SM * 2048 MAC * 2 * SMs * clk
  return 0;
}
