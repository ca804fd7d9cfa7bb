// MUFU throughput probe (diagnostic): lane-ops per clock per SM for ex2 / rcp / tanh and for the swish sequence
// y * rcp(1 + ex2(-y log2 e)), with 4 / 8 / 16 warps per scheduler-quarter's worth of CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/diag_mufu tools/diag_mufu.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) mufu_kernel(float* out, int iters, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      else if (OP == 3) {   // swish: fmul, ex2, fadd, rcp, fmul
        float e = x[i] * -1.4426950408889634f;
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e));
        e += 1.f;
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(e));
        x[i] = x[i] * e + 0.5f;
      } else {              // FFMA only (reference for the issue rate)
        x[i] = fmaf(x[i], 1.0001f, 0.5f);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;
}

template <int OP>
void run(const char* name, int mufu_per_elem, float* d, int threads, int sms, double ghz_hint) {
  const int iters = 4000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  mufu_kernel<OP><<<sms, threads>>>(d, iters, 0.3f);
  cudaEventRecord(a);
  mufu_kernel<OP><<<sms, threads>>>(d, iters, 0.3f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double elems = (double)sms * threads * iters * 8;
  printf("%-6s threads/SM %4d: %7.3f ms  %6.2f elem/ns  = %5.2f elem/clk/SM @%.2f GHz (%d MUFU per elem -> %5.2f MUFU lanes/clk/SM)\n",
         name, threads, ms, elems / (ms * 1e6), elems / (ms * 1e6) / sms / ghz_hint, ghz_hint, mufu_per_elem,
         mufu_per_elem * elems / (ms * 1e6) / sms / ghz_hint);
}

int main() {
  float* d; cudaMalloc(&d, 4);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double ghz = clk * 1e-6;
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("ex2", 1, d, threads, 148, ghz);
    run<1>("rcp", 1, d, threads, 148, ghz);
    run<2>("tanh", 1, d, threads, 148, ghz);
    run<3>("swish", 2, d, threads, 148, ghz);
    run<4>("ffma", 0, d, threads, 148, ghz);
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
