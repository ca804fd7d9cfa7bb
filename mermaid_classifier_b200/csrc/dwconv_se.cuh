// K4: depthwise kxk conv (TF-SAME asymmetric pad) + BN + swish + deterministic SE partial pool.
// K5: squeeze-excite FCs (pool -> FC + bias -> swish -> FC + bias -> sigmoid).
//
// Layout: NHWC.  A thread owns one 16-byte channel vector (4 fp32 / 8 bf16 channels) and a
// strip of TW adjacent output pixels, so every global access is a coalesced 16-byte vector
// and the sliding window is reused from registers along x.
//
// Block = (CGT channel-group threads) x (PT position threads); blockIdx.x = band of output
// rows, blockIdx.y = patch.  Each (position-thread, channel) pair accumulates its own pool
// sum, the CTA reduces them in a fixed order and writes one partial per (patch, band, channel):
// no atomics, so features are bit-reproducible run to run.
#pragma once
#include "common.cuh"

namespace mc {

template <typename T, int K, int S, int TW>
__global__ void __launch_bounds__(256)
dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w,  // [K*K][C]
              const float* __restrict__ scale, const float* __restrict__ bias, T* __restrict__ out,
              float* __restrict__ pool_partial,  // [n][nbands][C]
              int C, int Hin, int Hout, int pad, int rows_per_band, int nbands) {
  constexpr int VN = Vec<T>::N;
  constexpr int NCOL = (TW - 1) * S + K;
  extern __shared__ float pool_s[];  // [PT][C]
  const int CG = C / VN;
  const int CGT = blockDim.x, PT = blockDim.y;
  const int band = blockIdx.x;
  const int64_t n = blockIdx.y;
  const int y0 = band * rows_per_band;
  const int y1 = min(Hout, y0 + rows_per_band);
  const int nstrips = (Hout + TW - 1) / TW;  // square maps: Wout == Hout
  const int npos = (y1 - y0) * nstrips;
  const T* in_n = in + n * (int64_t)Hin * Hin * C;
  T* out_n = out + n * (int64_t)Hout * Hout * C;

  for (int cg = threadIdx.x; cg < CG; cg += CGT) {
    const int c0 = cg * VN;
    float sc[VN], bi[VN], psum[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      sc[e] = scale[c0 + e];
      bi[e] = bias[c0 + e];
      psum[e] = 0.f;
    }
    for (int pos = threadIdx.y; pos < npos; pos += PT) {
      const int oy = y0 + pos / nstrips;
      const int ox0 = (pos % nstrips) * TW;
      float acc[TW][VN];
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[t][e] = 0.f;
      const int ix0 = ox0 * S - pad;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S + ky - pad;
        if (iy < 0 || iy >= Hin) continue;
        const T* row = in_n + (int64_t)iy * Hin * C + c0;
        Vec<T> v[NCOL];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
          const int ix = ix0 + j;
          if (ix >= 0 && ix < Hin) {
            v[j] = Vec<T>::load(row + (int64_t)ix * C);
          } else {
#pragma unroll
            for (int e = 0; e < VN; ++e) v[j].v[e] = 0.f;
          }
        }
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float wv[VN];
          const float* wp = w + (ky * K + kx) * C + c0;
#pragma unroll
          for (int e = 0; e < VN; e += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(wp + e));
            wv[e] = t4.x; wv[e + 1] = t4.y; wv[e + 2] = t4.z; wv[e + 3] = t4.w;
          }
#pragma unroll
          for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int e = 0; e < VN; ++e) acc[t][e] = fmaf(v[t * S + kx].v[e], wv[e], acc[t][e]);
        }
      }
#pragma unroll
      for (int t = 0; t < TW; ++t) {
        const int ox = ox0 + t;
        if (ox < Hout) {
          Vec<T> o;
#pragma unroll
          for (int e = 0; e < VN; ++e) {
            const float y = silu_f(fmaf(acc[t][e], sc[e], bi[e]));
            o.v[e] = y;
            psum[e] += y;
          }
          o.store(out_n + ((int64_t)oy * Hout + ox) * C + c0);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) pool_s[threadIdx.y * C + c0 + e] = psum[e];
  }
  __syncthreads();
  const int tid = threadIdx.y * CGT + threadIdx.x;
  for (int c = tid; c < C; c += CGT * PT) {
    float s = 0.f;
    for (int p = 0; p < PT; ++p) s += pool_s[p * C + c];
    pool_partial[(n * nbands + band) * C + c] = s;
  }
}

// One CTA per patch.  gate[n][c] = sigmoid(W2 . swish(W1 . mean + b1) + b2).
__global__ void __launch_bounds__(256)
se_kernel(const float* __restrict__ pool_partial, int nbands, float inv_hw, const float* __restrict__ w1,
          const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
          float* __restrict__ gate, int C, int Cse) {
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
    for (int j = 0; j < Cse; ++j) s = fmaf(w2[c * Cse + j], hidden[j], s);
    gate[n * C + c] = sigmoid_f(s);
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
