// Store-pattern probe (diagnostic, not part of the library): how fast can 148 persistent CTAs write an [M][N] fp32 matrix
// when every warp-wide store instruction covers (a) 8 rows x 32 B, (b) 8 rows x 64 B, (c) 4 rows x 128 B, (d) 1 x 512 B?
// The GEMM epilogues write pattern (a)/(b) (tcgen05.ld 16x256b fragments); a plain copy kernel writes (d).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/diag_store tools/diag_store.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) store_kernel(float* __restrict__ out, long long M, int N, int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const long long tiles = M / 32;   // a warp handles 32 rows x 32 columns per step
  const int cgroups = N / 32;
  for (long long t = (long long)blockIdx.x * warps + warp; t < tiles * cgroups; t += (long long)gridDim.x * warps) {
    const long long r0 = (t / cgroups) * 32;
    const int c0 = (int)(t % cgroups) * 32;
    float v = (float)(t & 1023) * 0.001f + lane;
    if (act) v = __fdividef(v, 1.f + __expf(-v));
    if (MODE == 0) {   // 8 rows x 32 B per instruction: lane (lr = lane / 4, q = lane % 4) writes float2 at (row, 8 i + 2 q)
      const int lr = lane >> 2, q = lane & 3;
#pragma unroll
      for (int h = 0; h < 4; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float2*>(out + (r0 + 8 * h + lr) * N + c0 + 8 * i + 2 * q) = make_float2(v, v + i);
    } else if (MODE == 1) {   // 8 rows x 64 B: float4 at (row, 16 j + 4 q)
      const int lr = lane >> 2, q = lane & 3;
#pragma unroll
      for (int h = 0; h < 4; ++h)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          *reinterpret_cast<float4*>(out + (r0 + 8 * h + lr) * N + c0 + 16 * j + 4 * q) = make_float4(v, v + j, v, v);
    } else if (MODE == 2) {   // 4 rows x 128 B: float4 at (row = lane / 8, 4 (lane % 8))
      const int lr = lane >> 3, q = lane & 7;
#pragma unroll
      for (int h = 0; h < 8; ++h)
        *reinterpret_cast<float4*>(out + (r0 + 4 * h + lr) * N + c0 + 4 * q) = make_float4(v, v + h, v, v);
    } else {   // 512 contiguous bytes per instruction (flat: ignores the row structure)
      float* base = out + (r0 * N) + (long long)c0 * 32;
#pragma unroll
      for (int h = 0; h < 8; ++h) *reinterpret_cast<float4*>(base + h * 128 + 4 * lane) = make_float4(v, v + h, v, v);
    }
  }
}

template <int MODE>
float run(float* d, long long M, int N, int act, int ctas) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) store_kernel<MODE><<<ctas, 512>>>(d, M, N, act);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) store_kernel<MODE><<<ctas, 512>>>(d, M, N, act);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return (float)((double)M * N * 4 * 5 / (ms * 1e-3) / 1e9);
}

int main() {
  const int Ns[] = {96, 160, 480, 1152};   // multiples of 32 near the layer widths (144 -> 160)
  for (int N : Ns) {
    const long long M = (long long)(3.2e9 / (N * 4)) / 32 * 32;   // ~3.2 GB per pass: far beyond L2
    float* d; cudaMalloc(&d, M * N * 4);
    for (int ctas : {148, 296}) for (int act = 0; act < 2; ++act)
      printf("N=%4d ctas=%d act=%d | 8x32B %6.0f | 8x64B %6.0f | 4x128B %6.0f | 512B flat %6.0f GB/s\n", N, ctas, act,
             run<0>(d, M, N, act, ctas), run<1>(d, M, N, act, ctas), run<2>(d, M, N, act, ctas), run<3>(d, M, N, act, ctas));
    cudaFree(d);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
