// K8 row kernel: softmax -> per-class Platt sigmoid -> row-normalise -> overshoot clip ->
// argmax / top-k, one warp per feature row (the Linear/ReLU chain in front of it runs on
// the GEMM kernels).
//
// Restates CalibratedHead.forward (mermaid_classifier/pyspacer/inference/head.py:66-89)
// and, when no Platt parameters are given, TorchMLPClassifier._forward_probs
// (mermaid_classifier/pyspacer/torch_classifier.py:332-370: fp32 softmax, fp64 renorm).
#pragma once
#include "common.cuh"

namespace mc {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// np.argmax semantics over the warp's strided slice: larger value wins, ties -> lower index.
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) {
      v = ov;
      i = oi;
    }
  }
}

__global__ void head_rows_kernel(const float* __restrict__ logits, int ld, int K,
                                 const float* __restrict__ pa, const float* __restrict__ pb,
                                 double* __restrict__ proba, int32_t* __restrict__ labels, int topk,
                                 int32_t* __restrict__ topk_idx, float* __restrict__ topk_val, int64_t n) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * warps + warp;
  if (row >= n) return;
  float* s = sm + (size_t)warp * K;
  const float* x = logits + row * ld;

  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) {
    const float v = x[k];
    s[k] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float e = expf(s[k] - mx);
    s[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);

  if (pa != nullptr) {
    // calibrated: c_k = sigmoid(-(a_k p_k + b_k)); proba = c / sum(c)  (uniform when sum == 0)
    float csum = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float p = s[k] / sum;
      const float c = 1.f / (1.f + expf(fmaf(pa[k], p, pb[k])));
      s[k] = c;
      csum += c;
    }
    csum = warp_sum(csum);
    const bool nz = csum != 0.f;
    const float uni = 1.f / (float)K;
    for (int k = lane; k < K; k += 32) {
      float pr = nz ? s[k] / csum : uni;
      if (pr > 1.f && pr <= 1.f + 1e-5f) pr = 1.f;
      s[k] = pr;
      if (proba) proba[row * K + k] = (double)pr;
    }
  } else {
    // uncalibrated: fp32 softmax, then renormalise in fp64 so rows sum to exactly 1
    double dsum = 0.0;
    for (int k = lane; k < K; k += 32) {
      const float p = s[k] / sum;
      s[k] = p;
      dsum += (double)p;
    }
    dsum = warp_sum_d(dsum);
    if (proba)
      for (int k = lane; k < K; k += 32) proba[row * K + k] = (double)s[k] / dsum;
  }
  __syncwarp();

  const int rounds = topk > 0 ? topk : (labels ? 1 : 0);
  for (int r = 0; r < rounds; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
      const float v = s[k];
      if (v > bv) {
        bv = v;
        bi = k;
      }
    }
    warp_argmax(bv, bi);
    if (lane == 0) {
      if (r == 0 && labels) labels[row] = bi;
      if (topk > 0) {
        topk_idx[row * topk + r] = bi;
        if (topk_val) topk_val[row * topk + r] = bv;
      }
      if (bi < K) s[bi] = -INFINITY;
    }
    __syncwarp();
  }
}

// Copy an (n x d) fp32 matrix into a zero-padded (n x dp) one.
__global__ void pad_rows_kernel(const float* __restrict__ in, int d, float* __restrict__ out, int dp, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * dp) return;
  const int64_t r = t / dp;
  const int c = (int)(t % dp);
  out[t] = c < d ? in[r * d + c] : 0.f;
}

template <typename T>
__global__ void to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = to_f(in[t]);
}

}  // namespace mc
