// Pointwise (1x1 conv / Linear) GEMM on the CUDA cores with exact fp32 FMA arithmetic.
//
//   out[m][n] = act( (sum_k A[m][k] * gate[m / HW][k] * W[n][k]) * scale[n] + bias[n] ) + res[m][n]
//
// This is the fp32-exact engine: MC_MODE_FP32 uses it for every 1x1 conv (the parity
// configuration, max-abs 1e-3 against the CPU path), the MLP head uses it for its Linear
// layers, and it is the on-device cross-check for the tcgen05 kernels in pw_tc.cuh.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, k-major shared tiles.
#pragma once
#include "common.cuh"

namespace mc {

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2 };

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

// Requires K % 4 == 0 and N % 4 == 0 (every B0 channel count is a multiple of 8; the MLP
// head pads its dims on the host).
template <typename TA, typename TO, int ACT, bool GATE, bool RES>
__global__ void __launch_bounds__(256)
pw_simt_kernel(const TA* __restrict__ A, const float* __restrict__ W, const float* __restrict__ scale,
               const float* __restrict__ bias, const float* __restrict__ gate, const TO* __restrict__ res,
               TO* __restrict__ out, int64_t M, int N, int K, int HW) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t am = m0 + lrow;
  const bool a_ok = am < M;
  const int bn = n0 + lrow;
  const bool b_ok = bn < N;
  const float* grow = nullptr;
  if (GATE && a_ok) grow = gate + (am / HW) * K;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + lk;
    if (k < K) {
      if (a_ok) {
        load4<TA>(A + am * K + k, a);
        if (GATE) {
          float g[4];
          load4<float>(grow + k, g);
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] *= g[e];
        }
      }
      if (b_ok) load4<float>(W + (int64_t)bn * K + k, b);
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[lk + e][lrow] = a[e];
      Bs[lk + e][lrow] = b[e];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  const int n = n0 + tx * 4;
  if (n >= N) return;
  float sc[4] = {1.f, 1.f, 1.f, 1.f}, bi[4];
  if (scale) load4<float>(scale + n, sc);
  load4<float>(bias + n, bi);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y = fmaf(acc[i][j], sc[j], bi[j]);
      if (ACT == ACT_SILU) y = silu_f(y);
      if (ACT == ACT_RELU) y = fmaxf(y, 0.f);
      v[j] = y;
    }
    if (RES) {
      float r[4];
      load4<TO>(res + m * N + n, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r[j];
    }
    store4<TO>(out + m * N + n, v);
  }
}

}  // namespace mc
