// K5: squeeze-excite FCs (pool partials -> mean -> FC + bias -> swish -> FC + bias -> sigmoid) and the global
// average pool of the head conv.  The depthwise kernels (K4) that produce the pool partials live in dw_tma.cuh.
#pragma once
#include "common.cuh"
#include "pw_simt.cuh"

namespace mc {

// One CTA per P patches.  gate[n][c] = sigmoid(W2 . swish(W1 . mean + b1) + b2).  The two FC weight matrices (2 x Cse x C
// floats: 442 KB for the 1152-channel blocks) are read from L2 once per CTA, so P patches per CTA divide that traffic by P;
// every patch keeps its own accumulators and the same summation order as P = 1 (bit-identical gates).  Kept as an experiment
// switch (MC_SE_P): P = 1 is the fastest on B200, see se_launch.
template <int P>
__global__ void __launch_bounds__(256)
se_kernel(const float* __restrict__ pool_partial, int nbands, float inv_hw, const float* __restrict__ w1,
          const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
          float* __restrict__ gate, __nv_bfloat16* __restrict__ gate_h, int C, int Cse, int nb) {
  extern __shared__ float sm[];  // pooled[P][C] + hidden[P][Cse]
  float* pooled = sm;
  float* hidden = sm + P * C;
  const int64_t n0 = (int64_t)blockIdx.x * P;
  const int np = (int)min((int64_t)P, (int64_t)nb - n0);   // patches of this CTA (the last one may hold fewer)
  const int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();   // the pool partials come from the depthwise kernel right before
  for (int c = tid; c < C; c += 256) {
#pragma unroll
    for (int q = 0; q < P; ++q) {
      float s = 0.f;
      if (q < np)
        for (int b = 0; b < nbands; ++b) s += pool_partial[((n0 + q) * nbands + b) * C + c];
      pooled[q * C + c] = s * inv_hw;
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp; j < Cse; j += 8) {
    float s[P];
#pragma unroll
    for (int q = 0; q < P; ++q) s[q] = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float w = w1[j * C + c];
#pragma unroll
      for (int q = 0; q < P; ++q) s[q] = fmaf(w, pooled[q * C + c], s[q]);
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
      if (lane == 0) hidden[q * Cse + j] = silu_f(s[q] + b1[j]);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float s[P];
    const float bias = b2[c];
#pragma unroll
    for (int q = 0; q < P; ++q) s[q] = bias;
    for (int j = 0; j < Cse; ++j) {
      const float w = w2[j * C + c];  // w2 stored [Cse][C]: coalesced
#pragma unroll
      for (int q = 0; q < P; ++q) s[q] = fmaf(w, hidden[q * Cse + j], s[q]);
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (q < np) {
        const float gv = sigmoid_f(s[q]);
        gate[(n0 + q) * C + c] = gv;
        if (gate_h) gate_h[(n0 + q) * C + c] = __float2bfloat16_rn(gv);  // bf16 copy for the tcgen05 project conv
      }
    }
  }
}

// patches per CTA: MC_SE_P overrides (1, 2 or 4)
inline void se_launch(const float* pool_partial, int nbands, float inv_hw, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* gate, __nv_bfloat16* gate_h, int C, int Cse, int nb, cudaStream_t st) {
  static const int env_p = getenv("MC_SE_P") ? atoi(getenv("MC_SE_P")) : 0;
  // measured (B200, per-layer CUDA events): P = 2 is 10-30 % SLOWER than P = 1 on every block and P = 4 50-80 % slower -- the
  // kernel is bound by the latency of its dependent loops at ~7 resident CTAs per SM, not by the L2 re-reads
  const int P = env_p ? env_p : 1;
  const size_t smem = (size_t)P * (C + Cse) * sizeof(float);
  if (P >= 4) launch_pdl(PDL_SE, se_kernel<4>, dim3(cdiv(nb, 4)), dim3(256), smem, st, pool_partial, nbands, inv_hw, w1, b1, w2, b2, gate, gate_h, C, Cse, nb);
  else if (P == 2) launch_pdl(PDL_SE, se_kernel<2>, dim3(cdiv(nb, 2)), dim3(256), smem, st, pool_partial, nbands, inv_hw, w1, b1, w2, b2, gate, gate_h, C, Cse, nb);
  else launch_pdl(PDL_SE, se_kernel<1>, dim3(nb), dim3(256), smem, st, pool_partial, nbands, inv_hw, w1, b1, w2, b2, gate, gate_h, C, Cse, nb);
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
