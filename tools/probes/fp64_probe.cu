// DFMA throughput / latency probe for B200 (sm_100a): nvcc -arch=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  if (s == 12345.678) out[0] = s;
}

template <int CHAINS>
void run(int threads, int blocks_per_sm, int sms) {
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_kernel<CHAINS><<<sms * blocks_per_sm, threads>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  dfma_kernel<CHAINS><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)sms * blocks_per_sm * threads * iters * CHAINS;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3;
  printf("chains %2d threads/SM %4d : %.2f TFLOP/s, %.2f DFMA lanes/clk/SM, %.1f clk per dependent DFMA per warp-chain set\n",
         CHAINS, threads * blocks_per_sm, 2 * fmas / ms / 1e9, fmas / cyc / sms, cyc / iters);
  cudaFree(out);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  run<1>(32, 1, sms);      // latency: one warp, one chain
  run<1>(128, 1, sms);     // 1 warp per scheduler, 1 chain
  run<2>(128, 1, sms);
  run<4>(128, 1, sms);
  run<8>(128, 1, sms);
  run<16>(128, 1, sms);
  run<8>(256, 1, sms);     // 2 warps per scheduler
  run<8>(512, 1, sms);
  run<8>(1024, 1, sms);
  run<4>(1024, 2, sms);
  return 0;
}
