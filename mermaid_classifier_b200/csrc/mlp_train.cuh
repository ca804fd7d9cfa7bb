// K9: MLP-head training step kernels (fp32, CUDA cores).
//
// Restates the inner loop of TorchMLPClassifier.partial_fit
// (mermaid_classifier/pyspacer/torch_classifier.py:270-297): per mini-batch
//   logits = MLP(x);  loss = CE(logits, y, weight=class_w) + 0.5*alpha/mb * sum(W^2)
//   backward;  Adam(lr, beta1, beta2, eps)
// with the data-parallel split made explicit: every rank produces UN-NORMALISED sums
//   G = sum_rows w_r * d(-log p_r[y_r])/dtheta,  wsum = sum_rows w_r,  lsum = sum_rows w_r * (-log p_r[y_r]),
//   nrows = rows in this rank's share
// in one flat buffer, the buffer is summed over ranks (NCCL all-reduce), and the Adam kernel
// forms  g = G / wsum + alpha / nrows * W  -- exactly F.cross_entropy's weighted-mean reduction
// (torch_classifier.py:283) and the per-mini-batch L2 term (:288) for the GLOBAL mini-batch.
//
// Flat parameter / gradient layout: [W_0 (out0 x in0) | b_0 | W_1 | b_1 | ... ] with every width
// padded to a multiple of 4 (zero weights; padded classes are excluded from the softmax), then
// 4 statistics floats [wsum, lsum, nrows, 0] at the end of the gradient buffer only.
#pragma once
#include "common.cuh"
#include "head.cuh"

namespace mc {

enum { MLP_EPI_NONE = 0, MLP_EPI_BIAS_RELU = 1, MLP_EPI_BIAS = 2, MLP_EPI_RELU_MASK = 3 };

// C[m][n] = sum_k A(m,k) * B(k,n)      (64 x 64 x 16 tiles, 256 threads, 4x4 micro-tile)
//   A_MC == false: A[m * lda + k]   (k contiguous)      A_MC == true: A[k * lda + m]
//   B_NC == false: B[n * ldb + k]   (k contiguous)      B_NC == true: B[k * ldb + n]
// Contiguous dimensions must be multiples of 4 (16-byte vector loads); the other one is free.
// splits > 1 = split-K: slice z accumulates k in [z*k_per_split, (z+1)*k_per_split) and stores its raw partial to
// part + z * M * N; the LAST slice of a tile to finish (a ticket per tile) sums the partials in slice order -- the same
// fixed order whichever CTA happens to be last, so the result is deterministic -- and applies the epilogue.  (Round 1 ran
// that reduction as a second launch per GEMM: seven ~4 us launches per Adam step.)
// colsum (A_MC only): CTAs with by == 0 also write colsum[m] = sum_k A(m,k)  (bias gradient).
struct MlpGemmP {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  float* part;            // split-K partials [splits][M][N], then [splits][M] partial column sums
  int* tickets;           // one per output tile, zero between launches
  int M, N, K, k_per_split, splits, epi;
  const float* bias;
  const float* mask; int ldmask;
  float* colsum;
  int gx, gy;             // tile grid
};

template <bool A_MC, bool B_NC>
__device__ __forceinline__ void mlp_gemm_body(const MlpGemmP& p, const int bx, const int by, const int bz) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ int s_ticket;
  const float* __restrict__ A = p.A;
  const float* __restrict__ B = p.B;
  const int lda = p.lda, ldb = p.ldb, M = p.M, N = p.N, K = p.K;
  const int tid = threadIdx.x;
  const int m0 = bx * BM, n0 = by * BN;
  const int kbeg = bz * p.k_per_split, kend = min(K, kbeg + p.k_per_split);
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float csum = 0.f;
  const bool do_colsum = A_MC && p.colsum != nullptr && by == 0;

  // global -> register loads of one k-tile (issued one tile ahead of the math: the loop is latency-bound otherwise, a
  // 64 x 64 x 200 tile spent ~1 us per 16-deep k step waiting for L2)
  auto load_tile = [&](int k0, float4& a4, float4& b4) {
    a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (A_MC) {
      const int k = k0 + (tid >> 4), m = m0 + (tid & 15) * 4;
      if (k < kend && m < M) a4 = *reinterpret_cast<const float4*>(A + (int64_t)k * lda + m);
    } else {
      const int m = m0 + (tid >> 2), k = k0 + (tid & 3) * 4;
      if (m < M && k < kend) a4 = *reinterpret_cast<const float4*>(A + (int64_t)m * lda + k);
    }
    if (B_NC) {
      const int k = k0 + (tid >> 4), n = n0 + (tid & 15) * 4;
      if (k < kend && n < N) b4 = *reinterpret_cast<const float4*>(B + (int64_t)k * ldb + n);
    } else {
      const int n = n0 + (tid >> 2), k = k0 + (tid & 3) * 4;
      if (n < N && k < kend) b4 = *reinterpret_cast<const float4*>(B + (int64_t)n * ldb + k);
    }
  };
  float4 a4, b4;
  load_tile(kbeg, a4, b4);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    __syncthreads();   // the previous tile's math is done with As / Bs
    if (A_MC) {
      *reinterpret_cast<float4*>(&As[tid >> 4][(tid & 15) * 4]) = a4;
    } else {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      As[kk + 0][r] = a4.x; As[kk + 1][r] = a4.y; As[kk + 2][r] = a4.z; As[kk + 3][r] = a4.w;
    }
    if (B_NC) {
      *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = b4;
    } else {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      Bs[kk + 0][r] = b4.x; Bs[kk + 1][r] = b4.y; Bs[kk + 2][r] = b4.z; Bs[kk + 3][r] = b4.w;
    }
    __syncthreads();
    if (k0 + BK < kend) load_tile(k0 + BK, a4, b4);   // in flight during the math below
    if (do_colsum && tid < BM) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) csum += As[kk][tid];   // rows past kend were stored as zeros
    }
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
  const int64_t MN = (int64_t)M * N;
  if (do_colsum && tid < BM && m0 + tid < M) {
    if (p.splits > 1) p.part[(int64_t)p.splits * MN + (int64_t)bz * M + m0 + tid] = csum;   // partial column sums after the tile partials
    else p.colsum[m0 + tid] = csum;
  }
  const int n = n0 + tx * 4;
  const bool col_ok = n < N;
  const bool split = p.splits > 1;
  const int epi = p.epi;
  auto epilogue_store = [&](int m, float (&v)[4]) {
    if (epi == MLP_EPI_BIAS_RELU || epi == MLP_EPI_BIAS) {
      const float4 t = *reinterpret_cast<const float4*>(p.bias + n);
      v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
      if (epi == MLP_EPI_BIAS_RELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      }
    } else if (epi == MLP_EPI_RELU_MASK) {
      const float4 mk = *reinterpret_cast<const float4*>(p.mask + (int64_t)m * p.ldmask + n);
      v[0] = mk.x > 0.f ? v[0] : 0.f; v[1] = mk.y > 0.f ? v[1] : 0.f;
      v[2] = mk.z > 0.f ? v[2] : 0.f; v[3] = mk.w > 0.f ? v[3] : 0.f;
    }
    *reinterpret_cast<float4*>(p.C + (int64_t)m * p.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
  };
  if (!split) {
    if (col_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m < M) {
          float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
          epilogue_store(m, v);
        }
      }
    }
    return;
  }
  // split-K: raw partial of this slice, then the last slice of the tile reduces
  if (col_ok) {
    float* Cz = p.part + (int64_t)bz * MN;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < M) *reinterpret_cast<float4*>(Cz + (int64_t)m * N + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(&p.tickets[by * p.gx + bx], 1);
  __syncthreads();
  if (s_ticket != p.splits - 1) return;
  __threadfence();
  if (col_ok) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < M) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int z = 0; z < p.splits; ++z) {   // slice order: the sum does not depend on which slice finished last
          const float4 q = __ldcg(reinterpret_cast<const float4*>(p.part + (int64_t)z * MN + (int64_t)m * N + n));
          v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
        }
        epilogue_store(m, v);
      }
    }
  }
  if (do_colsum && tid < BM && m0 + tid < M) {
    float cs = 0.f;
    for (int z = 0; z < p.splits; ++z) cs += __ldcg(p.part + (int64_t)p.splits * MN + (int64_t)z * M + m0 + tid);
    p.colsum[m0 + tid] = cs;
  }
  if (tid == 0) p.tickets[by * p.gx + bx] = 0;   // ready for the next launch on this stream
}

template <bool A_MC, bool B_NC>
__global__ void __launch_bounds__(256) mlp_gemm_kernel(const MlpGemmP p) {
  mlp_gemm_body<A_MC, B_NC>(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// The two GEMMs that consume one layer's delta -- dW_i = delta_i^T . input_i (+ bias gradient) and
// dX = delta_i . W_i (ReLU mask, split-K) -- are independent: one launch, the CTAs of both side by side.
__global__ void __launch_bounds__(256) mlp_gemm_pair_kernel(const MlpGemmP p0, const MlpGemmP p1) {
  int b = blockIdx.x;
  const int n0 = p0.gx * p0.gy * p0.splits;
  if (b < n0) {
    mlp_gemm_body<true, true>(p0, b % p0.gx, (b / p0.gx) % p0.gy, b / (p0.gx * p0.gy));
  } else {
    b -= n0;
    mlp_gemm_body<false, true>(p1, b % p1.gx, (b / p1.gx) % p1.gy, b / (p1.gx * p1.gy));
  }
}

// One warp per mini-batch row: log-softmax over the K real classes, un-normalised delta
//   delta[r][k] = w_r * (softmax_k - [k == y_r])     (padded columns K..Kp-1 = 0)
//   row_stat[r] = (w_r, w_r * -log p[y_r])
// The last CTA to finish (ticket) then reduces the row statistics in a fixed order into the 4 statistics floats of the
// gradient buffer (round 1: a second, single-CTA launch).
__device__ __forceinline__ void mlp_stats_reduce(const float2* __restrict__ row_stat, int rows, float* __restrict__ stats);

__global__ void __launch_bounds__(256)
mlp_ce_kernel(const float* __restrict__ logits, int Kp, int K, const int32_t* __restrict__ y,
              const float* __restrict__ class_w, float* __restrict__ delta,
              float2* __restrict__ row_stat, int rows, float* __restrict__ stats, int* __restrict__ ticket) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r < rows) {
  const float* x = logits + (int64_t)r * Kp;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, x[k]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int k = lane; k < K; k += 32) sum += expf(x[k] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  const int yr = y[r];
  const float w = class_w ? class_w[yr] : 1.f;
  float* d = delta + (int64_t)r * Kp;
  for (int k = lane; k < Kp; k += 32) {
    float v = 0.f;
    if (k < K) v = w * (expf(x[k] - lse) - (k == yr ? 1.f : 0.f));
    d[k] = v;
  }
  if (lane == 0) row_stat[r] = make_float2(w, w * (lse - x[yr]));
  }
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  mlp_stats_reduce(row_stat, rows, stats);
  if (threadIdx.x == 0) *ticket = 0;
}

// Fixed-order reduction of the row statistics by the first 256 threads of the CTA: thread t sums rows t, t + 256, ..., then a
// tree over the threads.  (Callers with more than 256 threads pass only those through; the barriers inside are named so
// that the other warps need not take part.)
__device__ __forceinline__ void mlp_stats_reduce_n(const float2* __restrict__ row_stat, int rows, float* __restrict__ stats, int nthreads) {
  __shared__ float sw[256], sl[256];
  float w = 0.f, l = 0.f;
  for (int r = threadIdx.x; r < rows; r += 256) {
    const float2 s = __ldcg(row_stat + r);
    w += s.x;
    l += s.y;
  }
  sw[threadIdx.x] = w;
  sl[threadIdx.x] = l;
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sw[threadIdx.x] += sw[threadIdx.x + o];
      sl[threadIdx.x] += sl[threadIdx.x + o];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
  }
  if (threadIdx.x == 0) {
    stats[0] = sw[0];
    stats[1] = sl[0];
    stats[2] = (float)rows;
    stats[3] = 0.f;
  }
}
__device__ __forceinline__ void mlp_stats_reduce(const float2* __restrict__ row_stat, int rows, float* __restrict__ stats) {
  __shared__ float sw[256], sl[256];
  float w = 0.f, l = 0.f;
  for (int r = threadIdx.x; r < rows; r += 256) {
    const float2 s = __ldcg(row_stat + r);
    w += s.x;
    l += s.y;
  }
  sw[threadIdx.x] = w;
  sl[threadIdx.x] = l;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sw[threadIdx.x] += sw[threadIdx.x + o];
      sl[threadIdx.x] += sl[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = sw[0];
    stats[1] = sl[0];
    stats[2] = (float)rows;
    stats[3] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-local middle of the step.  At mini-batch 200 everything between the first layer's GEMM and the weight-gradient
// GEMMs is a chain of small dependent products per ROW: layers 1..L-1 forward, softmax / cross-entropy, and the deltas
// back down to layer 0.  As GEMM launches that chain was 4 + 3 launches of 6-17 us each (grids of 8-60 CTAs, latency- and
// launch-bound); here one CTA carries R rows through the whole chain with the activations in shared memory and the
// weights streamed from L2, and writes what the weight-gradient GEMMs need (activations 1..L-2, every delta) plus the
// row statistics.  Forward: a warp per output neuron, lanes split the reduction (coalesced weight rows).  Backward: a
// thread per input neuron, loop over the outputs (coalesced again, no reduction).
constexpr int MLP_RL_MAX_LAYERS = 8;
struct MlpRowLocalP {
  int L, K, rows;
  int dims_p[MLP_RL_MAX_LAYERS + 1];
  int64_t w_off[MLP_RL_MAX_LAYERS], b_off[MLP_RL_MAX_LAYERS];
  const float* params;
  const float* act0;                  // [rows][dims_p[1]]: layer 0 output (bias + ReLU applied)
  float* act[MLP_RL_MAX_LAYERS];      // act[i], i = 1..L-2: written (input of the weight gradient of layer i + 1)
  float* delta[MLP_RL_MAX_LAYERS];    // delta[i], i = 0..L-1: written
  const int32_t* y;
  const float* class_w;
  float2* row_stat;
  float* stats;
  int* ticket;
};

template <int R>
__global__ void __launch_bounds__(512) mlp_rowlocal_kernel(const MlpRowLocalP p) {
  extern __shared__ float rl_smem[];
  // h[i] = activations of layer i (i = 0..L-1, the last one = logits), then two delta buffers of the widest layer
  float* h[MLP_RL_MAX_LAYERS];
  int maxw = 0;
  {
    float* q = rl_smem;
    for (int i = 0; i < p.L; ++i) {
      h[i] = q;
      q += R * p.dims_p[i + 1];
      maxw = max(maxw, p.dims_p[i + 1]);
    }
    h[p.L] = q;   // (unused slot keeps the indexing below simple)
  }
  float* dbuf[2] = {h[p.L], h[p.L] + R * maxw};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int r0 = blockIdx.x * R;
  const int nr = min(R, p.rows - r0);

  // layer-0 activations of this CTA's rows
  {
    const int w0 = p.dims_p[1];
    for (int t = tid; t < R * w0; t += blockDim.x) {
      const int r = t / w0, k = t - r * w0;
      h[0][t] = r < nr ? p.act0[(int64_t)(r0 + r) * w0 + k] : 0.f;
    }
  }
  __syncthreads();
  // forward, layers 1 .. L-1
  for (int i = 1; i < p.L; ++i) {
    const int Ki = p.dims_p[i], Ni = p.dims_p[i + 1];
    const float* __restrict__ W = p.params + p.w_off[i];
    const float* __restrict__ bias = p.params + p.b_off[i];
    const float* hin = h[i - 1];
    float* hout = h[i];
    const bool relu = i < p.L - 1;
    // NB output neurons per warp at a time: NB independent weight rows in flight per lane (the loop is bound by the L2
    // latency of the weight stream, one row at a time left the SM's ~64 B/clk L2 port mostly idle)
    constexpr int NB = 4;
    for (int n0 = warp * NB; n0 < Ni; n0 += nwarps * NB) {
      float acc[NB][R];
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[j][r] = 0.f;
#pragma unroll 2
      for (int k = lane; k < Ki; k += 32) {
        float w[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) w[j] = n0 + j < Ni ? W[(int64_t)(n0 + j) * Ki + k] : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float hv = hin[r * Ki + k];
#pragma unroll
          for (int j = 0; j < NB; ++j) acc[j][r] = fmaf(w[j], hv, acc[j][r]);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int n = n0 + j;
        if (n < Ni) {   // warp-uniform
#pragma unroll
          for (int r = 0; r < R; ++r) acc[j][r] = warp_sum(acc[j][r]);
          const float b = bias[n];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float v = acc[j][r] + b;
            if (relu) v = fmaxf(v, 0.f);
            if (lane == r) {
              hout[r * Ni + n] = v;
              if (relu && r < nr) p.act[i][(int64_t)(r0 + r) * Ni + n] = v;
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // softmax / cross-entropy of the rows: warp r handles row r (as mlp_ce_kernel)
  {
    const int Kp = p.dims_p[p.L], K = p.K;
    float* d = dbuf[0];
    for (int r = warp; r < R; r += nwarps) {
      const float* x = h[p.L - 1] + r * Kp;
      if (r < nr) {
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, x[k]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int k = lane; k < K; k += 32) sum += expf(x[k] - mx);
        sum = warp_sum(sum);
        const float lse = mx + logf(sum);
        const int yr = p.y[r0 + r];
        const float w = p.class_w ? p.class_w[yr] : 1.f;
        float* dg = p.delta[p.L - 1] + (int64_t)(r0 + r) * Kp;
        for (int k = lane; k < Kp; k += 32) {
          float v = 0.f;
          if (k < K) v = w * (expf(x[k] - lse) - (k == yr ? 1.f : 0.f));
          d[r * Kp + k] = v;
          dg[k] = v;
        }
        if (lane == 0) p.row_stat[r0 + r] = make_float2(w, w * (lse - x[yr]));
      } else {
        for (int k = lane; k < Kp; k += 32) d[r * Kp + k] = 0.f;
      }
    }
  }
  __syncthreads();
  // deltas down to layer 0: delta[i-1][r][k] = (act[i-1][r][k] > 0) * sum_n delta[i][r][n] W_i[n][k]
  for (int i = p.L - 1; i >= 1; --i) {
    const int Ki = p.dims_p[i], Ni = p.dims_p[i + 1];
    const float* __restrict__ W = p.params + p.w_off[i];
    const float* din = dbuf[(p.L - 1 - i) & 1];
    float* dout = dbuf[(p.L - i) & 1];
    const float* hmask = h[i - 1];
    for (int k = tid; k < Ki; k += blockDim.x) {
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
      // 16 weight loads in flight per thread (Ni is a multiple of 4; the tail runs one by one)
      int n = 0;
      for (; n + 16 <= Ni; n += 16) {
        float w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = W[(int64_t)(n + j) * Ki + k];
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fmaf(din[r * Ni + n + j], w[j], acc[r]);
      }
      for (; n < Ni; ++n) {
        const float w = W[(int64_t)n * Ki + k];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(din[r * Ni + n], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float v = hmask[r * Ki + k] > 0.f ? acc[r] : 0.f;
        dout[r * Ki + k] = v;
        if (r < nr) p.delta[i - 1][(int64_t)(r0 + r) * Ki + k] = v;
      }
    }
    __syncthreads();
  }
  // the last CTA reduces the row statistics (fixed order)
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(p.ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 256) mlp_stats_reduce_n(p.row_stat, p.rows, p.stats, 256);
  if (tid == 0) *p.ticket = 0;
}

// the weight-gradient GEMMs of ALL layers in one launch (they only depend on the deltas and activations written above)
constexpr int MLP_MULTI_MAX = MLP_RL_MAX_LAYERS;
struct MlpGemmMulti {
  int n;
  int first[MLP_MULTI_MAX + 1];   // first CTA of each problem
  MlpGemmP p[MLP_MULTI_MAX];
};
__global__ void __launch_bounds__(256) mlp_gemm_multi_kernel(const MlpGemmMulti mp) {
  int q = 0;
  while (q + 1 < mp.n && (int)blockIdx.x >= mp.first[q + 1]) ++q;
  const MlpGemmP& p = mp.p[q];
  const int b = blockIdx.x - mp.first[q];
  mlp_gemm_body<true, true>(p, b % p.gx, (b / p.gx) % p.gy, b / (p.gx * p.gy));
}

struct MlpSegs {
  int n_layers;
  int64_t w_off[8], b_off[8], end;  // weights of layer i: [w_off[i], b_off[i]); biases: [b_off[i], w_off[i+1] or end)
};

constexpr int MLP_ADAM_THREADS = 256;

// sum of squares of the weight entries (not biases), one partial per block, fixed order inside a block
__device__ __forceinline__ void block_ssq_store(float ssq, float* out) {
  __shared__ float red[MLP_ADAM_THREADS];
  red[threadIdx.x] = ssq;
  __syncthreads();
  for (int o = MLP_ADAM_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

__device__ __forceinline__ bool mlp_is_weight(const MlpSegs& s, int64_t i) {
  bool w = false;
  for (int l = 0; l < s.n_layers; ++l) w = w || (i >= s.w_off[l] && i < s.b_off[l]);
  return w;
}

__global__ void __launch_bounds__(MLP_ADAM_THREADS)
mlp_ssq_kernel(const float* __restrict__ p, MlpSegs segs, float* __restrict__ ssq_part) {
  const int64_t i = (int64_t)blockIdx.x * MLP_ADAM_THREADS + threadIdx.x;
  float s = 0.f;
  if (i < segs.end && mlp_is_weight(segs, i)) s = p[i] * p[i];
  block_ssq_store(s, ssq_part + blockIdx.x);
}

// Device-side cursor of a run of equal-sized Adam steps replayed from a CUDA graph: everything that changes from step to
// step -- which rows of the call's (xs, ys) form the mini-batch, Adam's bias corrections -- is read through it, so the
// kernels of a step keep the same arguments and one captured pair of steps can be replayed for the whole run.
struct MlpCtl {
  long long step;            // index into offs / bc of the step being executed
  const float* xs;           // the call's padded, ordered feature rows
  const int32_t* ys;
  const int64_t* offs;       // [n_steps + 1] row offsets of the call's steps
  const float2* bc;          // [n_steps] (1 - beta1^t, sqrt(1 - beta2^t)) computed on the host exactly as for a launched step
};

// mini-batch of the cursor's step -> the fixed staging buffers the step's kernels read (rows x Dp floats, rows labels)
__global__ void mlp_stage_kernel(const MlpCtl* __restrict__ ctl, float4* __restrict__ xb, int32_t* __restrict__ yb, int rows,
                                 int Dp4) {
  const int64_t off = ctl->offs[ctl->step];
  const float4* src = reinterpret_cast<const float4*>(ctl->xs) + off * Dp4;
  const int n4 = rows * Dp4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) xb[i] = src[i];
  if (blockIdx.x == 0)
    for (int r = threadIdx.x; r < rows; r += blockDim.x) yb[r] = ctl->ys[off + r];
}

__global__ void mlp_advance_kernel(MlpCtl* ctl) { ctl->step += 1; }

// Adam step on the flat parameter vector (torch.optim.Adam semantics: eps added to sqrt(v_hat)).
//   g = G / wsum + (alpha / nrows) * W   for weights,  G / wsum for biases.
// Block 0 also books the mini-batch loss:  loss_acc += (lsum / wsum + 0.5 * alpha / nrows * ssq_prev) * nrows,
// rows_acc += nrows, where ssq_prev = sum(W^2) BEFORE this update (partials written by the previous step).
__global__ void __launch_bounds__(MLP_ADAM_THREADS)
mlp_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ G,
                MlpSegs segs, const float* __restrict__ stats, float lr, float alpha, float beta1, float beta2, float eps,
                float bc1, float bc2_sqrt, const float* __restrict__ ssq_prev, int n_ssq, float* __restrict__ ssq_next,
                double* __restrict__ loss_acc, const MlpCtl* __restrict__ ctl) {
  if (ctl != nullptr) {   // graph replay: this step's bias corrections come through the cursor
    const float2 bc = ctl->bc[ctl->step];
    bc1 = bc.x;
    bc2_sqrt = bc.y;
  }
  const float wsum = stats[0], nrows = stats[2];
  const int64_t i = (int64_t)blockIdx.x * MLP_ADAM_THREADS + threadIdx.x;
  float s = 0.f;
  if (i < segs.end && nrows > 0.f) {
    const bool isw = mlp_is_weight(segs, i);
    const float pi = p[i];
    float g = wsum != 0.f ? G[i] / wsum : 0.f;
    if (isw) g = fmaf(alpha / nrows, pi, g);
    const float mi = beta1 * m[i] + (1.f - beta1) * g;
    const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pn = pi - (lr / bc1) * (mi / denom);
    p[i] = pn;
    if (isw) s = pn * pn;
  } else if (i < segs.end && mlp_is_weight(segs, i)) {
    s = p[i] * p[i];
  }
  block_ssq_store(s, ssq_next + blockIdx.x);
  if (blockIdx.x == 0 && nrows > 0.f) {
    // sum(W^2) before this update: the per-block partials of the previous step, reduced by the whole block in a fixed
    // order (a single thread walking ~3 400 partials was 70 of the kernel's 76 us)
    __shared__ float red2[MLP_ADAM_THREADS];
    float part = 0.f;
    for (int b = threadIdx.x; b < n_ssq; b += MLP_ADAM_THREADS) part += ssq_prev[b];
    __syncthreads();
    red2[threadIdx.x] = part;
    __syncthreads();
    for (int o = MLP_ADAM_THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) red2[threadIdx.x] += red2[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float ssq = red2[0];
      const double data = wsum != 0.f ? (double)stats[1] / (double)wsum : 0.0;
      loss_acc[0] += (data + 0.5 * (double)alpha / (double)nrows * (double)ssq) * (double)nrows;
      loss_acc[1] += (double)nrows;
    }
  }
}

// Rows of X (and their labels) selected by an index vector -- the shuffled order of one partial_fit
// call (torch_classifier.py:258-264) -- into the zero-padded training layout.  order == nullptr: identity.
__global__ void mlp_gather_rows_kernel(const float* __restrict__ X, const int32_t* __restrict__ y,
                                       const int64_t* __restrict__ order, int D, int Dp, float* __restrict__ xs,
                                       int32_t* __restrict__ ys, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * Dp) return;
  const int64_t r = t / Dp;
  const int c = (int)(t % Dp);
  const int64_t src = order ? order[r] : r;
  xs[t] = c < D ? X[src * D + c] : 0.f;
  if (c == 0) ys[r] = y[src];
}

}  // namespace mc
