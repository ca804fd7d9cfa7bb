// K5: squeeze-excite FCs (pool partials -> mean -> FC + bias -> swish -> FC + bias -> sigmoid) and the global
// average pool of the head conv.  The depthwise kernels (K4) that produce the pool partials live in dw_tma.cuh.
#pragma once
#include "common.cuh"
#include "pw_simt.cuh"

namespace mc {

// One CTA per patch.  gate[n][c] = sigmoid(W2 . swish(W1 . mean + b1) + b2).
__global__ void __launch_bounds__(256)
se_kernel(const float* __restrict__ pool_partial, int nbands, float inv_hw, const float* __restrict__ w1,
          const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
          float* __restrict__ gate, __nv_bfloat16* __restrict__ gate_h, int C, int Cse) {
  extern __shared__ float sm[];  // pooled[C] + hidden[Cse]
  float* pooled = sm;
  float* hidden = sm + C;
  const int64_t n = blockIdx.x;
  const int tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) {
    float s = 0.f;
    for (int b = 0; b < nbands; ++b) s += pool_partial[(n * nbands + b) * C + c];
    pooled[c] = s * inv_hw;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp; j < Cse; j += 8) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w1[j * C + c], pooled[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) hidden[j] = silu_f(s + b1[j]);
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float s = b2[c];
    for (int j = 0; j < Cse; ++j) s = fmaf(w2[j * C + c], hidden[j], s);  // w2 stored [Cse][C]: coalesced
    const float gv = sigmoid_f(s);
    gate[n * C + c] = gv;
    if (gate_h) gate_h[n * C + c] = __float2bfloat16_rn(gv);  // bf16 copy for the tcgen05 project conv
  }
}

// Global average pool of the head conv output: feats[n][c] = mean_p x[n][p][c] (fp32 out).
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, float* __restrict__ feats, int HW, int C) {
  const int64_t n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const T* p = x + n * (int64_t)HW * C + c;
  float s = 0.f;
  for (int i = 0; i < HW; ++i) s += to_f(p[(int64_t)i * C]);
  feats[n * C + c] = s / (float)HW;
}

}  // namespace mc
