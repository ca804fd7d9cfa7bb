// K4: depthwise kxk conv (TF-SAME asymmetric pad) + BN + swish + deterministic SE partial pool.
// K5: squeeze-excite FCs (pool -> FC + bias -> swish -> FC + bias -> sigmoid).
//
// Layout: NHWC.  A thread owns a 4-channel vector (16 B fp32 / 8 B bf16) and a strip of TW adjacent
// output pixels, so global accesses are coalesced vectors and the sliding window is reused from
// registers along x and (rolling accumulators) along y.
//
// Each thread accumulates its own pool sum, the CTA reduces them in a fixed order and writes one
// partial per (patch, part, channel): no atomics, so features are bit-reproducible run to run.
#pragma once
#include "common.cuh"
#include "pw_simt.cuh"

namespace mc {

// Four channels as loaded from memory (kept packed while in flight).
template <typename T>
struct RawVec4;
template <>
struct RawVec4<float> {
  typedef float4 type;
  __device__ __forceinline__ static float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ static float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ static void unpack(const float4& r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
};
template <>
struct RawVec4<__nv_bfloat16> {
  typedef uint2 type;
  __device__ __forceinline__ static uint2 load(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ static uint2 zero() { return make_uint2(0u, 0u); }
  __device__ __forceinline__ static void unpack(const uint2& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x << 16);
    v[1] = __uint_as_float(r.x & 0xFFFF0000u);
    v[2] = __uint_as_float(r.y << 16);
    v[3] = __uint_as_float(r.y & 0xFFFF0000u);
  }
};

// Depthwise conv with ROLLING row accumulators: every input row is loaded once per thread and
// folded into the ceil(K/S) output rows it contributes to, so the k x k window costs one global
// read per input element instead of k (the 5x5 layers were L2-bound on re-reads).
// Thread = (4-channel group, strip of TW output columns); it walks down its band of rows keeping
// NL x TW x 4 fp32 accumulators in registers, NL = ceil(K/S) live output rows (slot indices are
// compile-time after unrolling the row loop S*NL-fold).  Weights of the CTA's channel slice sit
// in shared memory.  TF-"SAME" asymmetric padding enters through `pad` (leading pad).
//   blockDim = (CGT channel groups, PT strips); grid = (row-band x strip-chunk, patch, channel chunk)
template <typename T, int K, int S, int TW>
__global__ void __launch_bounds__(256, 2)
dwconv_roll_kernel(const T* __restrict__ in, const float* __restrict__ w,  // [K*K][C]
                   const float* __restrict__ scale, const float* __restrict__ bias, T* __restrict__ out,
                   float* __restrict__ pool_partial,  // [n][nparts][C]
                   int C, int Hin, int Hout, int pad, int rows_per_band, int nxchunks) {
  constexpr int NL = (K + S - 1) / S;      // live output rows
  constexpr int P = S * NL;                // unroll period of the input-row loop
  constexpr int NCOL = (TW - 1) * S + K;   // input columns per strip
  extern __shared__ __align__(16) float dsm[];
  const int CGT = blockDim.x, PT = blockDim.y;
  float* w_s = dsm;                       // [K*K][CGT*4]
  float* pool_s = dsm + K * K * CGT * 4;  // [PT][CGT*4]
  const int part = blockIdx.x;
  const int band = part / nxchunks, xchunk = part % nxchunks;
  const int64_t n = blockIdx.y;
  const int c0 = (blockIdx.z * CGT + threadIdx.x) * 4;
  const int y0 = band * rows_per_band, y1 = min(Hout, y0 + rows_per_band);
  const int nstrips = (Hout + TW - 1) / TW;
  const int strip = xchunk * PT + threadIdx.y;
  const int tid = threadIdx.y * CGT + threadIdx.x;

  for (int i = tid; i < K * K * CGT; i += CGT * PT) {
    const int tap = i / CGT, g = i % CGT;
    const float4 t4 = *reinterpret_cast<const float4*>(w + (int64_t)tap * C + (blockIdx.z * CGT + g) * 4);
    *reinterpret_cast<float4*>(w_s + (tap * CGT + g) * 4) = t4;
  }
  __syncthreads();

  float psum[4] = {0.f, 0.f, 0.f, 0.f};
  if (strip < nstrips) {
    float sc[4], bi[4];
    load4<float>(scale + c0, sc);
    load4<float>(bias + c0, bi);
    const int ox0 = strip * TW;
    const int ix0 = ox0 * S - pad;
    const T* in_n = in + n * (int64_t)Hin * Hin * C + c0;
    T* out_n = out + n * (int64_t)Hout * Hout * C + c0;
    float acc[NL][TW][4];
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[a][t][e] = 0.f;
    // step t handles input row iy = y0*S - pad + t; output row y0 + (t - ky)/S takes tap ky.
    // Rows are software-pipelined: the raw (still packed) vectors of row t+1 are in flight while
    // row t is folded into the accumulators.
    const int nsteps = (y1 - 1 - y0) * S + K;
    using Raw = typename RawVec4<T>::type;
    Raw rawA[NCOL], rawB[NCOL];
    auto load_row = [&](int t, Raw(&raw)[NCOL]) {
      const int iy = y0 * S - pad + t;
      const bool ok = t < nsteps && iy >= 0 && iy < Hin;
      const T* row = in_n + (int64_t)iy * Hin * C;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const int ix = ix0 + j;
        if (ok && ix >= 0 && ix < Hin) raw[j] = RawVec4<T>::load(row + (int64_t)ix * C);
        else raw[j] = RawVec4<T>::zero();
      }
    };
    load_row(0, rawA);
    for (int t0 = 0; t0 < nsteps; t0 += 2 * P) {
#pragma unroll
      for (int r = 0; r < 2 * P; ++r) {
        const int t = t0 + r;
        if (t < nsteps) {
          Raw(&cur)[NCOL] = (r % 2 == 0) ? rawA : rawB;
          Raw(&nxt)[NCOL] = (r % 2 == 0) ? rawB : rawA;
          load_row(t + 1, nxt);
          const int iy = y0 * S - pad + t;
          if (iy >= 0 && iy < Hin) {
            float v[NCOL][4];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) RawVec4<T>::unpack(cur[j], v[j]);
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
              if ((r - ky + P * 4) % S == 0) {  // compile-time: this input row feeds tap ky of some output row
                const int slot = (((r - ky + P * 4) / S) % NL);
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                  const float4 w4 = *reinterpret_cast<const float4*>(w_s + ((ky * K + kx) * CGT + threadIdx.x) * 4);
#pragma unroll
                  for (int q = 0; q < TW; ++q) {
                    acc[slot][q][0] = fmaf(v[q * S + kx][0], w4.x, acc[slot][q][0]);
                    acc[slot][q][1] = fmaf(v[q * S + kx][1], w4.y, acc[slot][q][1]);
                    acc[slot][q][2] = fmaf(v[q * S + kx][2], w4.z, acc[slot][q][2]);
                    acc[slot][q][3] = fmaf(v[q * S + kx][3], w4.w, acc[slot][q][3]);
                  }
                }
              }
            }
          }
          // the output row whose last tap (ky = K-1) is this input row is complete
          if ((r - (K - 1) + P * 4) % S == 0) {  // compile-time
            const int done = (((r - (K - 1) + P * 4) / S) % NL);
            const int td = t - (K - 1);
            const int oy = y0 + td / S;
            if (td >= 0 && oy < y1) {
#pragma unroll
              for (int q = 0; q < TW; ++q) {
                const int ox = ox0 + q;
                if (ox < Hout) {
                  float y[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    y[e] = silu_f(fmaf(acc[done][q][e], sc[e], bi[e]));
                    psum[e] += y[e];
                  }
                  store4<T>(out_n + ((int64_t)oy * Hout + ox) * C, y);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < TW; ++q)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[done][q][e] = 0.f;
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) pool_s[(threadIdx.y * CGT + threadIdx.x) * 4 + e] = psum[e];
  __syncthreads();
  for (int i = tid; i < CGT * 4; i += CGT * PT) {
    float s = 0.f;
    for (int pp = 0; pp < PT; ++pp) s += pool_s[pp * CGT * 4 + i];
    pool_partial[(n * gridDim.x + part) * C + blockIdx.z * CGT * 4 + i] = s;
  }
}

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
