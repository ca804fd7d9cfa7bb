// K4: depthwise k x k conv (TF-"SAME" asymmetric pad) + BN + swish + deterministic SE partial pool,
// with the input staged by TMA.
//
//   out[n][oy][ox][c] = swish( scale[c] * sum_{ky,kx} in[n][oy*S - pad + ky][ox*S - pad + kx][c] * w[ky][kx][c] + bias[c] )
//
// One CTA = (band of output rows, patch, slice of CB = 4*CGT channels).  A producer warp streams the
// band's input rows through a ring of shared-memory stages with 4-D TMA boxes
// {CB channels, BWIN columns, 1 row, 1 patch}; rows and columns outside the image are ZERO-FILLED by
// the TMA unit, so the SAME padding costs no bounds checks and no address arithmetic in the math
// loop, and the bytes in flight per SM are set by the ring depth, not by registers.
// Consumer thread = (4-channel group, strip of TW output columns): for every input row it reads its
// (TW-1)*S + K column vectors from shared memory (16 B fp32 / 8 B bf16, conflict-free: a 4-channel
// group per lane, CGT consecutive lanes = one contiguous pixel) and folds the row into the
// ceil(K/S) output rows it contributes to (rolling accumulators in registers), so each input element
// is read from shared memory once per thread and from HBM once per CTA.
//
// Each thread accumulates its own pool sum, the CTA reduces them in a fixed order and writes one
// partial per (patch, band, channel): no atomics, so features are bit-reproducible run to run.
#pragma once
#include "common.cuh"
#include "pw_simt.cuh"
#include "pw_tc.cuh"

namespace mc {

namespace ptx {
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
}  // namespace ptx

constexpr int DW_MAX_STAGES = 8;

struct DwArgs {
  const float* w;      // [K*K][C]
  const float* scale;  // [C]
  const float* bias;   // [C]
  void* out;           // [n][Hout][Hout][C]
  float* pool_partial; // [n][gridDim.x][C]
  int C, Hin, Hout, pad, rows_per_band;
  int cgt, pt;         // channel groups / column strips per CTA (consumer threads = cgt * pt)
  int bwin;            // input columns per staged row = (pt*TW - 1)*S + K
  int stages, row_bytes;  // ring depth; bytes between stages (box bytes rounded up to 128)
  int box_bytes;          // bytes one TMA box delivers = bwin * 4*cgt * sizeof(T)
  int n_off;           // first patch of this launch inside the tensor map
};

template <typename T>
__device__ __forceinline__ void dw_load_vec(uint32_t addr, float (&v)[4]);
template <>
__device__ __forceinline__ void dw_load_vec<float>(uint32_t addr, float (&v)[4]) {
  const uint4 r = ptx::lds128(addr);
  v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
}
template <>
__device__ __forceinline__ void dw_load_vec<__nv_bfloat16>(uint32_t addr, float (&v)[4]) {
  const uint2 r = ptx::lds64(addr);
  v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xFFFF0000u);
  v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xFFFF0000u);
}

template <typename T, int K, int S, int TW>
__global__ void __launch_bounds__(256, 2)
dw_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const DwArgs a) {
  constexpr int NL = (K + S - 1) / S;     // live output rows
  constexpr int P = S * NL;               // unroll period of the input-row loop
  constexpr int NCOL = (TW - 1) * S + K;  // input columns per strip
  constexpr int ES = (int)sizeof(T);
  extern __shared__ uint8_t dw_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)dw_smem_raw + 127) & ~(uintptr_t)127);
  const int CGT = a.cgt, PT = a.pt, CB = CGT * 4;
  const int n_cons = CGT * PT;
  uint8_t* ring = smem;                                            // [stages][row_bytes]
  float* w_s = (float*)(ring + (size_t)a.stages * a.row_bytes);    // [K*K][CB]
  float* pool_s = w_s + K * K * CB;                                // [PT][CB]
  uint64_t* full = (uint64_t*)(pool_s + PT * CB);                  // [stages]
  uint64_t* empty = full + DW_MAX_STAGES;                          // [stages]

  const int tid = threadIdx.x;
  const int band = blockIdx.x;
  const int n = blockIdx.y;
  const int cb0 = blockIdx.z * CB;  // first channel of this CTA
  const int y0 = band * a.rows_per_band, y1 = min(a.Hout, y0 + a.rows_per_band);
  const int nsteps = (y1 - 1 - y0) * S + K;
  const int iy0 = y0 * S - a.pad;

  for (int i = tid; i < K * K * CGT; i += blockDim.x) {
    const int tap = i / CGT, g = i % CGT;
    *reinterpret_cast<float4*>(w_s + tap * CB + g * 4) = *reinterpret_cast<const float4*>(a.w + (int64_t)tap * a.C + cb0 + g * 4);
  }
  if (tid == 0) {
    for (int s = 0; s < DW_MAX_STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], (uint32_t)n_cons);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmIn);
  }
  __syncthreads();

  const bool producer = tid >= (int)blockDim.x - 32;
  float psum[4] = {0.f, 0.f, 0.f, 0.f};
  if (producer) {
    if (tid == (int)blockDim.x - 32) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < nsteps; ++t) {
        ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&full[s], (uint32_t)a.box_bytes);
        ptx::tma_load_4d(ring + (size_t)s * a.row_bytes, &tmIn, &full[s], cb0, -a.pad, iy0 + t, a.n_off + n);
        if (++s == a.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (tid < n_cons) {
    const int cg = tid % CGT, strip = tid / CGT;
    const int c0 = cb0 + cg * 4;
    float sc[4], bi[4];
    load4<float>(a.scale + c0, sc);
    load4<float>(a.bias + c0, bi);
    const int ox0 = strip * TW;
    const uint32_t ring_u32 = ptx::smem_u32(ring) + (uint32_t)((strip * TW * S * CB + cg * 4) * ES);
    const uint32_t col_pitch = (uint32_t)(CB * ES);
    const float* w_t = w_s + cg * 4;
    T* out_n = (T*)a.out + (int64_t)n * a.Hout * a.Hout * a.C + c0;
    float acc[NL][TW][4];
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
      for (int q = 0; q < TW; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[l][q][e] = 0.f;
    int s = 0;
    uint32_t ph = 0;
    // step t handles input row iy = iy0 + t; output row y0 + (t - ky)/S takes tap row ky from it
    for (int t0 = 0; t0 < nsteps; t0 += P) {
#pragma unroll
      for (int r = 0; r < P; ++r) {
        const int t = t0 + r;
        if (t < nsteps) {
          ptx::mbar_wait(&full[s], ph);
          const uint32_t rowbase = ring_u32 + (uint32_t)s * (uint32_t)a.row_bytes;
          float v[NCOL][4];
#pragma unroll
          for (int j = 0; j < NCOL; ++j) dw_load_vec<T>(rowbase + (uint32_t)j * col_pitch, v[j]);
          const int iy = iy0 + t;
          if (iy >= 0 && iy < a.Hin) {   // zero-filled rows contribute nothing
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
              if ((r - ky + P * 4) % S == 0) {  // compile-time: this input row feeds tap row ky of some output row
                const int slot = (((r - ky + P * 4) / S) % NL);
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                  const float4 w4 = *reinterpret_cast<const float4*>(w_t + (ky * K + kx) * CB);
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
          ptx::mbar_arrive(&empty[s]);   // the row is in registers / folded in: release the stage
          if (++s == a.stages) {
            s = 0;
            ph ^= 1;
          }
          // the output row whose last tap row (ky = K-1) is this input row is complete
          if ((r - (K - 1) + P * 4) % S == 0) {  // compile-time
            const int done = (((r - (K - 1) + P * 4) / S) % NL);
            const int td = t - (K - 1);
            const int oy = y0 + td / S;
            if (td >= 0 && oy < y1) {
#pragma unroll
              for (int q = 0; q < TW; ++q) {
                const int ox = ox0 + q;
                if (ox < a.Hout) {
                  float y[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    y[e] = silu_f(fmaf(acc[done][q][e], sc[e], bi[e]));
                    psum[e] += y[e];
                  }
                  store4<T>(out_n + ((int64_t)oy * a.Hout + ox) * a.C, y);
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
  if (tid < n_cons) {
#pragma unroll
    for (int e = 0; e < 4; ++e) pool_s[tid * 4 + e] = psum[e];   // tid = strip * CGT + cg
  }
  __syncthreads();
  for (int i = tid; i < CB; i += blockDim.x) {
    float s = 0.f;
    for (int pp = 0; pp < PT; ++pp) s += pool_s[pp * CB + i];
    a.pool_partial[((int64_t)n * gridDim.x + band) * a.C + cb0 + i] = s;
  }
}

// ---------------------------------------------------------------------------------------
// host side: per-block launch plans
// ---------------------------------------------------------------------------------------
struct DwLayer {
  int K = 0, S = 0, TW = 0, C = 0, Hin = 0, Hout = 0, pad = 0;
  int cgt = 0, pt = 0, cz = 0, bwin = 0, stages = 0, row_bytes = 0, box_bytes = 0, rows_per_band = 0, nbands = 0, threads = 0;
  size_t smem = 0;
  const void* in_ptr = nullptr;  // the tensor map below describes this buffer
  CUtensorMap tm;
};

inline int dw_pick_tw(int K, int S, int Hout) {
  if (S == 1 && K == 3 && Hout >= 28) return 4;
  return 2;
}

// 4-D map over an NHWC activation buffer: dims {C, W, H, N}; box {CB, BWIN, 1, 1}; no swizzle; OOB -> 0.
inline int dw_make_map(CUtensorMap* map, bool f32, const void* base, int C, int H, int64_t n_patches, int cb, int bwin) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)H, (cuuint64_t)n_patches};
  cuuint64_t gstride[3] = {(cuuint64_t)C * es, (cuuint64_t)H * C * es, (cuuint64_t)H * H * C * es};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)bwin, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled (depthwise) failed with CUresult " + std::to_string((int)r));
  return MC_OK;
}

// Channel groups per CTA: the largest divisor of C/4 with cgt * pt <= 224 consumer threads (+ one producer warp = 256) and 4*cgt <= 256 (TMA box limit).
// The box's inner extent must be a multiple of 16 bytes: any cgt in fp32 (16 B per group), even cgt in bf16.
inline int dw_pick_cgt(int cg_total, int pt, bool f32) {
  int best = 0;
  for (int d = 1; d <= cg_total && d <= 64; ++d)
    if (cg_total % d == 0 && d * pt <= 224 && (f32 || d % 2 == 0)) best = d;
  return best;
}

inline int dw_plan_layer(DwLayer* l, const BlockCfg& b, bool f32) {
  l->K = b.k; l->S = b.stride; l->C = b.c_mid; l->Hin = b.h_in; l->Hout = b.h_out; l->pad = b.pad;
  l->TW = dw_pick_tw(b.k, b.stride, b.h_out);
  l->pt = (b.h_out + l->TW - 1) / l->TW;
  l->cgt = dw_pick_cgt(b.c_mid / 4, l->pt, f32);
  if (l->cgt == 0) return fail(MC_ERR_UNSUPPORTED, "depthwise: no channel-slice size satisfies the TMA box alignment");
  l->cz = b.c_mid / 4 / l->cgt;
  l->bwin = (l->pt * l->TW - 1) * b.stride + b.k;
  const int es = f32 ? 4 : 2;
  l->box_bytes = l->bwin * l->cgt * 4 * es;
  l->row_bytes = (l->box_bytes + 127) / 128 * 128;
  l->rows_per_band = b.h_out >= 112 ? 16 : (b.h_out >= 56 ? 14 : b.h_out);
  l->nbands = (b.h_out + l->rows_per_band - 1) / l->rows_per_band;
  l->threads = (l->cgt * l->pt + 31) / 32 * 32 + 32;
  const size_t fixed = 128 + (size_t)(b.k * b.k + l->pt) * l->cgt * 4 * sizeof(float) + 2 * DW_MAX_STAGES * sizeof(uint64_t);
  // ring depth: as many rows as fit in ~100 KB (two CTAs per SM), at most DW_MAX_STAGES
  int stages = (int)((100 * 1024 - fixed) / l->row_bytes);
  l->stages = stages < 2 ? 2 : (stages > DW_MAX_STAGES ? DW_MAX_STAGES : stages);
  l->smem = fixed + (size_t)l->stages * l->row_bytes;
  if (l->bwin > 256 || l->cgt * 4 > 256) return fail(MC_ERR_UNSUPPORTED, "depthwise tile exceeds the TMA box limits");
  return MC_OK;
}

template <typename T>
inline int dw_tma_launch(DwLayer& l, const CUtensorMap& tm, const DwArgs& a, int nb, cudaStream_t st) {
  dim3 grid(l.nbands, nb, l.cz), block(l.threads);
#define DW_CASE(KK, SS, TT)                                                                                          \
  if (l.K == KK && l.S == SS && l.TW == TT) {                                                                        \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      MC_CUDA(cudaFuncSetAttribute(dw_tma_kernel<T, KK, SS, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    dw_tma_kernel<T, KK, SS, TT><<<grid, block, l.smem, st>>>(tm, a);                                                 \
    MC_CHECK_LAUNCH();                                                                                                \
    return MC_OK;                                                                                                     \
  }
  DW_CASE(3, 1, 4)
  DW_CASE(3, 1, 2)
  DW_CASE(5, 1, 2)
  DW_CASE(3, 2, 2)
  DW_CASE(5, 2, 2)
#undef DW_CASE
  return fail(MC_ERR_UNSUPPORTED, "depthwise kernel/stride combination");
}

}  // namespace mc
