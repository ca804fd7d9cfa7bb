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
  uint8_t* smem = dw_smem_raw + ((128u - (ptx::smem_u32(dw_smem_raw) & 127u)) & 127u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
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
  pdl_trigger();
  pdl_wait();   // the input map (and the output buffer's previous readers) belong to earlier kernels

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
                    y[e] = bn_silu<T>(acc[done][q][e], sc[e], bi[e]);
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
// Compile-time layer shapes of the register-weight kernel: the ten distinct depthwise layers of
// EfficientNet-B0 past the two 112-wide ones.  Everything the kernel indexes with (channel count, map
// size, slice width, ring geometry) is a constant, so shared-memory and global offsets fold into
// immediates instead of per-row integer arithmetic (the run-time-shaped version spent ~145 address
// instructions per input row against 200 FMAs).
struct DwShape {
  int K, S, C, Hin;                 // kernel, stride, channels, input map size
  int Hout, pad, TW, CH;            // output map size, leading SAME pad, columns / channels per thread
  int cb, ptc, nxc, cgt, cz;        // channels per CTA, strips per CTA, column chunks, channel groups per CTA, slices
  int bwin, box_bytes, row_bytes;   // staged input columns, bytes per TMA box, ring pitch
  int rows_per_band, nbands, threads, stages, smem;
};

constexpr int DW_N_SHAPES = 12;
constexpr int DW_SHAPE_KSCH[DW_N_SHAPES][4] = {{3, 1, 144, 56},  {5, 2, 144, 56},  {5, 1, 240, 28}, {3, 2, 240, 28},
                                               {3, 1, 480, 14},  {5, 1, 480, 14},  {5, 1, 672, 14}, {5, 2, 672, 14},
                                               {5, 1, 1152, 7},  {3, 1, 1152, 7},
                                               // the two 112-wide layers: used in bf16 mode only (in fp32 the 4-channel-vector
                                               // kernel already runs them at ~90 % of the HBM peak)
                                               {3, 1, 32, 112},  {3, 2, 96, 112}};

constexpr int dw_shape_index(int K, int S, int C, int Hin) {
  for (int i = 0; i < DW_N_SHAPES; ++i)
    if (DW_SHAPE_KSCH[i][0] == K && DW_SHAPE_KSCH[i][1] == S && DW_SHAPE_KSCH[i][2] == C && DW_SHAPE_KSCH[i][3] == Hin) return i;
  return -1;
}

// Slice / strip selection (same rule at compile time for the kernel and at run time for the host plan).
__host__ __device__ constexpr DwShape dw_make_shape(int K, int S, int C, int Hin, int es) {
  DwShape d{};
  d.K = K; d.S = S; d.C = C; d.Hin = Hin;
  d.Hout = (Hin + S - 1) / S;
  int total = (d.Hout - 1) * S + K - Hin;
  if (total < 0) total = 0;
  d.pad = total / 2;
  d.CH = 2;
  d.TW = (K == 3 && S == 1) ? 8 : 4;
  const int pt = (d.Hout + d.TW - 1) / d.TW;
  double best = -1.0;
  for (int cb = 8; cb <= 256 && cb <= C; cb += 8) {
    if (C % cb) continue;
    const int cgt = cb / d.CH;
    int ptc = pt < 224 / cgt ? pt : 224 / cgt;
    if (ptc < 1) continue;
    const int nxc = (pt + ptc - 1) / ptc;
    ptc = (pt + nxc - 1) / nxc;
    // prefer full CTAs, one column chunk, and slices that are whole 128-byte lines of a pixel
    double score = (double)cgt * ptc * (nxc == 1 ? 1.0 : 0.93) * ((cb * es) % 128 == 0 ? 1.0 : 0.9) *
                   (cb * es >= 128 ? 1.0 : 0.6);
    // 5x5 stride 1 over the 28-wide map in bf16: three column chunks re-stage the 4-column halo (16 input columns per 12
    // outputs); the whole row in one CTA with 96-byte slices reads a third less through shared memory
    if (K == 5 && S == 1 && Hin == 28 && es == 2 && cb == 48) score = 1e9;
    if (score > best) {
      best = score;
      d.cb = cb; d.ptc = ptc; d.nxc = nxc;
    }
  }
  d.cgt = d.cb / d.CH;
  d.cz = C / d.cb;
  d.bwin = (d.ptc * d.TW - 1) * S + K;
  d.box_bytes = d.bwin * d.cb * es;
  d.row_bytes = (d.box_bytes + 127) / 128 * 128;
  d.rows_per_band = d.Hout >= 112 ? 16 : (d.Hout >= 56 ? 14 : d.Hout);
  d.nbands = (d.Hout + d.rows_per_band - 1) / d.rows_per_band;
  d.threads = (d.cgt * d.ptc + 31) / 32 * 32 + 32;
  const int fixed = 128 + 2 * d.ptc * d.cb * 4 + 2 * 8 * 8;
  int stages = (64 * 1024 - fixed) / d.row_bytes;
  d.stages = stages < 2 ? 2 : (stages > 8 ? 8 : stages);
  d.smem = fixed + d.stages * d.row_bytes;
  return d;
}

// ---------------------------------------------------------------------------------------
// Register-weight variant for every layer past the two 112-wide ones.
//
// The kernel above re-reads its k*k weight vectors from shared memory for every input row (one
// 16-byte load per 4*TW FMAs), which makes the 5x5 and the narrow 3x3 layers shared-memory
// bandwidth bound (ncu: l1tex 77 %, fma 30 %).  Here a thread owns CH channels and a strip of TW
// output columns and keeps its k*k*CH weights in REGISTERS for the whole launch, so an input row
// costs (TW-1)*S + K vector loads against K*K*TW*CH/S FMAs.  A CTA walks several patches
// (grid-stride over blockIdx.z), so the weights, the barrier set-up and the TMA ring are amortised
// over them; the ring keeps streaming across patch boundaries.  Channel slices are the FASTEST grid
// dimension: sibling slices of one patch run together and share the 128-byte lines a slice
// boundary splits.
struct DwRegArgs {
  const float* w;
  const float* scale;
  const float* bias;
  void* out;
  float* pool_partial;  // [n][gridDim.y][C]
  int n_off, nb;        // first patch inside the tensor map; patches in this launch
};

template <typename T, int CH>
__device__ __forceinline__ void dw_load_ch(uint32_t addr, float (&v)[CH]);
template <>
__device__ __forceinline__ void dw_load_ch<float, 2>(uint32_t addr, float (&v)[2]) {
  const uint2 r = ptx::lds64(addr);
  v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y);
}
template <>
__device__ __forceinline__ void dw_load_ch<float, 4>(uint32_t addr, float (&v)[4]) {
  dw_load_vec<float>(addr, v);
}
template <>
__device__ __forceinline__ void dw_load_ch<__nv_bfloat16, 2>(uint32_t addr, float (&v)[2]) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  v[0] = __uint_as_float(r << 16); v[1] = __uint_as_float(r & 0xFFFF0000u);
}
template <>
__device__ __forceinline__ void dw_load_ch<__nv_bfloat16, 4>(uint32_t addr, float (&v)[4]) {
  dw_load_vec<__nv_bfloat16>(addr, v);
}
template <typename T, int CH>
__device__ __forceinline__ void dw_store_ch(T* p, const float (&y)[CH]);
template <>
__device__ __forceinline__ void dw_store_ch<float, 2>(float* p, const float (&y)[2]) {
  *reinterpret_cast<float2*>(p) = make_float2(y[0], y[1]);
}
template <>
__device__ __forceinline__ void dw_store_ch<float, 4>(float* p, const float (&y)[4]) { store4<float>(p, y); }
template <>
__device__ __forceinline__ void dw_store_ch<__nv_bfloat16, 2>(__nv_bfloat16* p, const float (&y)[2]) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(y[0], y[1]);
}
template <>
__device__ __forceinline__ void dw_store_ch<__nv_bfloat16, 4>(__nv_bfloat16* p, const float (&y)[4]) { store4<__nv_bfloat16>(p, y); }

template <typename T, int SHAPE>
__global__ void __launch_bounds__(256, 2)
dw_reg_kernel(const __grid_constant__ CUtensorMap tmIn, const DwRegArgs a) {
  constexpr DwShape SH = dw_make_shape(DW_SHAPE_KSCH[SHAPE][0], DW_SHAPE_KSCH[SHAPE][1], DW_SHAPE_KSCH[SHAPE][2],
                                       DW_SHAPE_KSCH[SHAPE][3], (int)sizeof(T));
  constexpr int K = SH.K, S = SH.S, TW = SH.TW, CH = SH.CH;
  constexpr int NL = (K + S - 1) / S;
  constexpr int P = S * NL;
  constexpr int NCOL = (TW - 1) * S + K;
  constexpr int ES = (int)sizeof(T);
  extern __shared__ uint8_t dw_smem_raw[];
  uint8_t* smem = dw_smem_raw + ((128u - (ptx::smem_u32(dw_smem_raw) & 127u)) & 127u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  constexpr int CGT = SH.cgt, PTC = SH.ptc, CB = SH.cb, NXC = SH.nxc, C = SH.C, HIN = SH.Hin, HOUT = SH.Hout, PAD = SH.pad;
  constexpr int RPB = SH.rows_per_band, STAGES = SH.stages, ROW_BYTES = SH.row_bytes, BOX_BYTES = SH.box_bytes;
  constexpr int n_cons = CGT * PTC;
  const int cons_threads = (int)blockDim.x - 32;                    // whole consumer warps
  uint8_t* ring = smem;                                              // [stages][row_bytes]
  float* pool_s = (float*)(ring + (size_t)STAGES * ROW_BYTES);   // [2][PTC][CB]
  uint64_t* full = (uint64_t*)(pool_s + 2 * PTC * CB);
  uint64_t* empty = full + DW_MAX_STAGES;

  const int tid = threadIdx.x;
  const int band = blockIdx.y / NXC, xchunk = blockIdx.y % NXC;
  const int cb0 = blockIdx.x * CB;
  const int y0 = band * RPB, y1 = min(HOUT, y0 + RPB);
  const int nsteps = (y1 - 1 - y0) * S + K;
  const int iy0 = y0 * S - PAD;
  const int strip0 = xchunk * PTC;
  const int x_start = strip0 * TW * S - PAD;

  if (tid == 0) {
    for (int s = 0; s < DW_MAX_STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], (uint32_t)n_cons);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmIn);
  }
  __syncthreads();
  pdl_trigger();
  if (tid >= cons_threads) pdl_wait();   // the producer waits here; the consumers load their weights (constants) first

  if (tid >= cons_threads) {
    // ================================ TMA producer ================================
    if (tid == cons_threads) {
      int s = 0;
      uint32_t ph = 0;
      for (int n = blockIdx.z; n < a.nb; n += gridDim.z) {
        for (int t = 0; t < nsteps; ++t) {
          ptx::mbar_wait(&empty[s], ph ^ 1);
          ptx::mbar_expect_tx(&full[s], (uint32_t)BOX_BYTES);
          ptx::tma_load_4d(ring + (size_t)s * ROW_BYTES, &tmIn, &full[s], cb0, x_start, iy0 + t, a.n_off + n);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    return;
  }
  // ================================== consumers ===================================
  const bool active = tid < n_cons;           // the last consumer warp may be partly idle
  const int cg = active ? tid % CGT : 0, sl = active ? tid / CGT : 0;
  const int c = cb0 + cg * CH;
  // CH == 2: the channel pair rides in 64-bit register pairs and every tap is ONE packed FMA (FFMA2, sm_100):
  // the three-register scalar FFMA issues every other cycle per scheduler, which capped the 5x5 layers at ~45 % FMA-pipe
  // utilisation with the memory system half idle.
  static_assert(CH == 2, "dw_reg_kernel packs a channel pair per thread");
  float2 w[K * K];
  float sc[CH], bi[CH];
#pragma unroll
  for (int i = 0; i < K * K; ++i) w[i] = make_float2(a.w[(int64_t)i * C + c], a.w[(int64_t)i * C + c + 1]);
#pragma unroll
  for (int e = 0; e < CH; ++e) {
    sc[e] = a.scale[c + e];
    bi[e] = a.bias[c + e];
  }
  const int ox0 = (strip0 + sl) * TW;
  const uint32_t ring_u32 = ptx::smem_u32(ring) + (uint32_t)((sl * TW * S * CB + cg * CH) * ES);
  const uint32_t col_pitch = (uint32_t)(CB * ES);
  int s = 0;
  uint32_t ph = 0;
  int pbuf = 0;
  pdl_wait();
  for (int n = blockIdx.z; n < a.nb; n += gridDim.z) {
    T* out_n = (T*)a.out + ((int64_t)n * HOUT * HOUT + ox0) * C + c;
    float2 acc[NL][TW];
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
      for (int q = 0; q < TW; ++q) acc[l][q] = make_float2(0.f, 0.f);
    float psum[CH];
#pragma unroll
    for (int e = 0; e < CH; ++e) psum[e] = 0.f;
    auto row_block = [&](const int t0) {
#pragma unroll
      for (int r = 0; r < P; ++r) {
        const int t = t0 + r;
        if (t < nsteps) {
          ptx::mbar_wait(&full[s], ph);
          const uint32_t rowbase = ring_u32 + (uint32_t)s * (uint32_t)ROW_BYTES;
          float2 v[NCOL];
#pragma unroll
          for (int j = 0; j < NCOL; ++j) {
            float t2[2];
            dw_load_ch<T, CH>(rowbase + (uint32_t)j * col_pitch, t2);
            v[j] = make_float2(t2[0], t2[1]);
          }
          const int iy = iy0 + t;
          if (iy >= 0 && iy < HIN) {
            // tap column outermost: consecutive FMAs go to all NL x TW accumulators in turn (each accumulator still takes its
            // taps in the same order, so the sums are bit-identical to the row-major loop)
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
              for (int ky = 0; ky < K; ++ky)
                if ((r - ky + P * 4) % S == 0) {
                  const int slot = (((r - ky + P * 4) / S) % NL);
#pragma unroll
                  for (int q = 0; q < TW; ++q) acc[slot][q] = __ffma2_rn(v[q * S + kx], w[ky * K + kx], acc[slot][q]);
                }
          }
          if (active) ptx::mbar_arrive(&empty[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
          if ((r - (K - 1) + P * 4) % S == 0) {
            const int done = (((r - (K - 1) + P * 4) / S) % NL);
            const int td = t - (K - 1);
            const int oy = y0 + td / S;
            if (td >= 0 && oy < y1 && active) {
              T* orow = out_n + (int64_t)oy * HOUT * C;
#pragma unroll
              for (int q = 0; q < TW; ++q) {
                if (ox0 + q < HOUT) {
                  float y[CH];
                  y[0] = bn_silu<T>(acc[done][q].x, sc[0], bi[0]);
                  y[1] = bn_silu<T>(acc[done][q].y, sc[1], bi[1]);
                  psum[0] += y[0];
                  psum[1] += y[1];
                  dw_store_ch<T, CH>(orow + q * C, y);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < TW; ++q) acc[done][q] = make_float2(0.f, 0.f);
          }
        }
      }
    };
    // Short bands (the 14x14 and 7x7 maps, <= 22 input rows) are unrolled completely; longer ones keep the row loop
    // rolled at its P-step period -- fully unrolled, the 28x28 layer is ~150 KB of code and thrashes the instruction cache.
    constexpr int NSTEPS_MAX = (RPB - 1) * S + K;
    if constexpr (SH.nbands == 1 && NSTEPS_MAX <= 22) {
#pragma unroll
      for (int t0 = 0; t0 < NSTEPS_MAX; t0 += P) row_block(t0);
    } else {
#pragma unroll 1
      for (int t0 = 0; t0 < nsteps; t0 += P) row_block(t0);
    }
    // per-patch pool partial: fixed-order reduction over the CTA's strips
    float* ps = pool_s + pbuf * PTC * CB;
    if (active) {
#pragma unroll
      for (int e = 0; e < CH; ++e) ps[sl * CB + cg * CH + e] = psum[e];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(cons_threads) : "memory");
    for (int i = tid; i < CB; i += cons_threads) {
      float sum = 0.f;
      for (int pp = 0; pp < PTC; ++pp) sum += ps[pp * CB + i];
      a.pool_partial[((int64_t)n * gridDim.y + blockIdx.y) * C + cb0 + i] = sum;
    }
    pbuf ^= 1;
  }
}

// ---------------------------------------------------------------------------------------
// host side: per-block launch plans
// ---------------------------------------------------------------------------------------
struct DwLayer {
  int K = 0, S = 0, TW = 0, C = 0, Hin = 0, Hout = 0, pad = 0;
  int cgt = 0, pt = 0, cz = 0, bwin = 0, stages = 0, row_bytes = 0, box_bytes = 0, rows_per_band = 0, nbands = 0, threads = 0;
  size_t smem = 0;
  bool lane = false;  // register-weight kernel (dw_reg_kernel): cb channels, ptc strips per CTA, nxc column chunks
  int cb = 0, ptc = 0, nxc = 1, ch = 4, shape = -1;
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

// Register-weight plan: copy the compile-time shape (dw_make_shape) of this layer into the run-time plan.
inline int dw_plan_reg(DwLayer* l, const BlockCfg& b, bool f32) {
  const int idx = dw_shape_index(b.k, b.stride, b.c_mid, b.h_in);
  if (idx < 0) return fail(MC_ERR_UNSUPPORTED, "depthwise layer shape is not in the compiled EfficientNet-B0 table");
  const DwShape d = dw_make_shape(b.k, b.stride, b.c_mid, b.h_in, f32 ? 4 : 2);
  if (d.Hout != b.h_out || d.pad != b.pad) return fail(MC_ERR_UNSUPPORTED, "depthwise shape table disagrees with the layer table");
  l->lane = true;
  l->shape = idx;
  l->TW = d.TW; l->ch = d.CH; l->cb = d.cb; l->ptc = d.ptc; l->nxc = d.nxc; l->cgt = d.cgt; l->cz = d.cz;
  l->pt = (d.Hout + d.TW - 1) / d.TW;
  l->bwin = d.bwin; l->box_bytes = d.box_bytes; l->row_bytes = d.row_bytes;
  l->rows_per_band = d.rows_per_band; l->nbands = d.nbands; l->threads = d.threads; l->stages = d.stages;
  l->smem = (size_t)d.smem;
  if (l->bwin > 256 || l->cb > 256) return fail(MC_ERR_UNSUPPORTED, "depthwise tile exceeds the TMA box limits");
  return MC_OK;
}

inline int dw_plan_layer(DwLayer* l, const BlockCfg& b, bool f32) {
  l->K = b.k; l->S = b.stride; l->C = b.c_mid; l->Hin = b.h_in; l->Hout = b.h_out; l->pad = b.pad;
  static const bool vec_env = getenv("MC_DW_VEC") != nullptr;   // experiment switch: 4-channel-vector kernel everywhere
  // fp32: the two 112-wide layers keep the 4-channel-vector kernel (HBM-bound there); bf16: register-weight kernel everywhere
  if (!vec_env && (b.h_in < 112 || !f32)) return dw_plan_reg(l, b, f32);
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
    static std::atomic<unsigned long long> attr_mask{0};                                                              \
    if (first_use_on_device(attr_mask))                                                                               \
      MC_CUDA(cudaFuncSetAttribute(dw_tma_kernel<T, KK, SS, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); \
    MC_CUDA(launch_pdl(PDL_DW, dw_tma_kernel<T, KK, SS, TT>, grid, block, l.smem, st, tm, a));                               \
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

template <typename T, int SHAPE>
inline int dw_reg_launch_shape(DwLayer& l, const CUtensorMap& tm, const DwRegArgs& a, dim3 grid, dim3 block, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_mask{0};
  if (first_use_on_device(attr_mask))
    MC_CUDA(cudaFuncSetAttribute(dw_reg_kernel<T, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  MC_CUDA(launch_pdl(PDL_DWREG, dw_reg_kernel<T, SHAPE>, grid, block, l.smem, st, tm, a));
  MC_CHECK_LAUNCH();
  return MC_OK;
}

template <typename T>
inline int dw_reg_launch(DwLayer& l, const CUtensorMap& tm, const DwRegArgs& a, int nb, cudaStream_t st) {
  const int per_patch = l.nbands * l.nxc * l.cz;
  const int gz = std::max(1, std::min(nb, 3000 / per_patch));   // CTAs walk patches grid-stride
  dim3 grid(l.cz, l.nbands * l.nxc, gz), block(l.threads);
  switch (l.shape) {
    case 0: return dw_reg_launch_shape<T, 0>(l, tm, a, grid, block, st);
    case 1: return dw_reg_launch_shape<T, 1>(l, tm, a, grid, block, st);
    case 2: return dw_reg_launch_shape<T, 2>(l, tm, a, grid, block, st);
    case 3: return dw_reg_launch_shape<T, 3>(l, tm, a, grid, block, st);
    case 4: return dw_reg_launch_shape<T, 4>(l, tm, a, grid, block, st);
    case 5: return dw_reg_launch_shape<T, 5>(l, tm, a, grid, block, st);
    case 6: return dw_reg_launch_shape<T, 6>(l, tm, a, grid, block, st);
    case 7: return dw_reg_launch_shape<T, 7>(l, tm, a, grid, block, st);
    case 8: return dw_reg_launch_shape<T, 8>(l, tm, a, grid, block, st);
    case 9: return dw_reg_launch_shape<T, 9>(l, tm, a, grid, block, st);
    case 10: return dw_reg_launch_shape<T, 10>(l, tm, a, grid, block, st);
    case 11: return dw_reg_launch_shape<T, 11>(l, tm, a, grid, block, st);
  }
  return fail(MC_ERR_UNSUPPORTED, "depthwise layer shape");
}

}  // namespace mc
