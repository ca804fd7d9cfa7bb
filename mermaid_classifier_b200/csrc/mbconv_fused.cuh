// WORK IN PROGRESS (round 2) -- compiled, NOT yet run on a GPU, off unless MC_FUSE_MASK selects a block.
//
// Expand 1x1 + BN + swish fused into the depthwise kernel (fp32 mode, blocks b1..b4).
//
// dw_reg_kernel consumes a shared-memory ring of input rows ([column][channel slice], one row per stage) that a TMA
// producer fills from the expanded map in HBM.  Here the ring is filled by a small GEMM pipeline instead, so the
// expanded map -- 55 % of all HBM bytes of the network -- is never written or read:
//
//   TMA warp        block-input tile x[R rows x W pixels = 112 pixels][C_in] (128B-swizzled, K zero-filled to 32)
//   transform warps lo = x - tf32(x)                                             (3xTF32 split, as pw_tc_kernel)
//   MMA warp        D[128 pixels x CB] = x * W_slice^T (3 terms), W slice resident in shared memory, D in TMEM (x2)
//   epilogue warps  TMEM lane = pixel: tcgen05.ld 32 channels -> BN + swish -> the pixel's channel slice into the ring
//                   row of its input row (zeros for rows outside the image: SAME padding pads the EXPANDED map)
//   consumer warps  dw_reg_kernel's loop: k x k taps out of the ring with weights in registers, BN + swish, store,
//                   SE pool partials
//
// All four shapes have R * W = 112 pixels per tile (R = 1, 2, 2, 4 input rows).  Ring pixel pitch = CB * 4 + 16 bytes
// keeps the epilogue's 16-byte stores of neighbouring pixels on different banks.
#pragma once
#include "dw_tma.cuh"
#include "pw_tc.cuh"

namespace mc {

struct FusedShape {
  int K, S, Cin, C, Hin, Hout, pad, TW, pt, CB, cgt, cz, bwin, pitch, row_bytes, stages, R, rows_per_band, nbands, kchunks;
  int a_stage_bytes, w_bytes, smem;
};

//                                         K  S  Cin  C   Hin
constexpr int FUSED_KSCH[4][5] = {{3, 2, 16, 96, 112}, {3, 1, 24, 144, 56}, {5, 2, 24, 144, 56}, {5, 1, 40, 240, 28}};
constexpr int FUSED_MAX_STAGES = 12;
constexpr int FUSED_TILE_PIX = 112;
constexpr int FUSED_CONS_WARPS = 8, FUSED_THREADS = (FUSED_CONS_WARPS + 4 + 4 + 2) * 32;   // 576

__host__ __device__ constexpr FusedShape fused_make_shape(int K, int S, int Cin, int C, int Hin) {
  FusedShape d{};
  d.K = K; d.S = S; d.Cin = Cin; d.C = C; d.Hin = Hin;
  d.Hout = (d.Hin + d.S - 1) / d.S;
  int total = (d.Hout - 1) * d.S + d.K - d.Hin;
  if (total < 0) total = 0;
  d.pad = total / 2;
  d.TW = (d.K == 3 && d.S == 1) ? 8 : 4;
  d.pt = (d.Hout + d.TW - 1) / d.TW;
  d.CB = d.C == 96 ? 32 : 48;
  d.cgt = d.CB / 2;
  d.cz = d.C / d.CB;
  d.bwin = (d.pt * d.TW - 1) * d.S + d.K;
  d.pitch = d.CB * 4 + 16;
  d.row_bytes = (d.bwin * d.pitch + 127) / 128 * 128;
  d.R = FUSED_TILE_PIX / d.Hin;
  d.stages = 2 * d.R + 2 < 4 ? 4 : 2 * d.R + 2;
  d.rows_per_band = d.Hout >= 56 ? 14 : d.Hout;
  d.nbands = (d.Hout + d.rows_per_band - 1) / d.rows_per_band;
  d.kchunks = (d.Cin + 31) / 32;
  d.a_stage_bytes = d.kchunks * 2 * TC_BM * 128;                        // hi + lo tiles of every k-chunk
  d.w_bytes = d.kchunks * 2 * ((d.CB * 128 + 1023) / 1024 * 1024);      // hi + lo slice of every k-chunk
  d.smem = 1024 + d.w_bytes + 2 * d.a_stage_bytes + d.stages * d.row_bytes + 2 * d.pt * d.CB * 4 + 2 * d.CB * 4 + 512;
  return d;
}

struct FusedArgs {
  const float* w_dw;       // depthwise weights [K*K][C]
  const float* s_dw;       // depthwise BN
  const float* b_dw;
  const float* s_exp;      // expand BN
  const float* b_exp;
  float* out;              // [n][Hout][Hout][C]
  float* pool_partial;     // [n][nbands][C]
  int nb;
};

template <int SHAPE>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
mbconv_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmWlo, const FusedArgs a) {
  constexpr FusedShape SH = fused_make_shape(FUSED_KSCH[SHAPE][0], FUSED_KSCH[SHAPE][1], FUSED_KSCH[SHAPE][2],
                                             FUSED_KSCH[SHAPE][3], FUSED_KSCH[SHAPE][4]);
  constexpr int K = SH.K, S = SH.S, TW = SH.TW, CB = SH.CB, C = SH.C, HIN = SH.Hin, HOUT = SH.Hout, PAD = SH.pad;
  constexpr int CGT = SH.cgt, PTC = SH.pt, RPB = SH.rows_per_band, STAGES = SH.stages, ROW_BYTES = SH.row_bytes;
  constexpr int PITCH = SH.pitch, R = SH.R, KCH = SH.kchunks, BWIN = SH.bwin;
  constexpr int NL = (K + S - 1) / S, P = S * NL, NCOL = (TW - 1) * S + K;
  constexpr int n_cons = CGT * PTC;
  constexpr int A_TILE = TC_BM * 128;
  constexpr int W_TILE = (CB * 128 + 1023) / 1024 * 1024;
  static_assert(R * HIN == FUSED_TILE_PIX, "a tile is R whole input rows of 112 pixels in total");
  static_assert(n_cons <= FUSED_CONS_WARPS * 32, "consumer threads");
  static_assert(STAGES <= FUSED_MAX_STAGES, "ring depth");

  extern __shared__ uint8_t fused_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)fused_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* w_s = smem;                                   // [kchunk][hi, lo] W slice tiles
  uint8_t* a_s = w_s + SH.w_bytes;                       // [2 stages][kchunk][hi, lo] x tiles
  uint8_t* ring = a_s + 2 * SH.a_stage_bytes;            // [STAGES][ROW_BYTES]
  float* pool_s = (float*)(ring + STAGES * ROW_BYTES);   // [2][PTC][CB]
  float* sc_s = pool_s + 2 * PTC * CB;                   // expand BN of this slice
  float* bi_s = sc_s + CB;
  uint64_t* bars = (uint64_t*)(bi_s + CB);
  uint64_t* r_full = bars;                               // [STAGES]  ring row written (HIN arrivals: one per pixel)
  uint64_t* r_empty = r_full + FUSED_MAX_STAGES;         // [STAGES]  ring row drained (n_cons arrivals)
  uint64_t* a_full = r_empty + FUSED_MAX_STAGES;         // [2] x tile landed
  uint64_t* a_ready = a_full + 2;                        // [2] lo operand written
  uint64_t* a_empty = a_ready + 2;                       // [2] MMAs reading the tile retired
  uint64_t* t_full = a_empty + 2;                        // [2] accumulator complete
  uint64_t* t_empty = t_full + 2;                        // [2] accumulator drained
  uint64_t* w_bar = t_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int WARP_EPI0 = FUSED_CONS_WARPS;            // 8..11: warp & 3 = TMEM quarter
  constexpr int WARP_TR0 = WARP_EPI0 + 4;                // 12..15
  constexpr int WARP_MMA = WARP_TR0 + 4, WARP_TMA = WARP_MMA + 1;

  const int band = blockIdx.y;
  const int cb0 = blockIdx.x * CB;
  const int y0 = band * RPB, y1 = min(HOUT, y0 + RPB);
  const int nsteps = (y1 - 1 - y0) * S + K;              // input rows this band consumes
  const int ntiles = (nsteps + R - 1) / R;
  const int iy0 = y0 * S - PAD;

  for (int i = tid; i < CB; i += FUSED_THREADS) {
    sc_s[i] = a.s_exp[cb0 + i];
    bi_s[i] = a.b_exp[cb0 + i];
  }
  // pad columns of every ring row are zero for the whole launch (the epilogue only writes image columns)
  for (int i = tid; i < STAGES * (BWIN - HIN) * CB; i += FUSED_THREADS) {
    const int st = i / ((BWIN - HIN) * CB), rem = i - st * (BWIN - HIN) * CB;
    const int pc = rem / CB, c = rem - pc * CB;
    const int col = pc < PAD ? pc : pc + HIN;
    *(float*)(ring + (size_t)st * ROW_BYTES + (size_t)col * PITCH + c * 4) = 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < FUSED_MAX_STAGES; ++s) {
      ptx::mbar_init(&r_full[s], (uint32_t)HIN);
      ptx::mbar_init(&r_empty[s], (uint32_t)n_cons);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_ready[s], 128);
      ptx::mbar_init(&a_empty[s], 1);
      ptx::mbar_init(&t_full[s], 1);
      ptx::mbar_init(&t_empty[s], 128);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmWlo);
  }
  if (warp == WARP_MMA) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_TMA) {
    // ================================ TMA producer ================================
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_bar, (uint32_t)(KCH * 2 * CB * 128));
      for (int kc = 0; kc < KCH; ++kc) {
        ptx::tma_load_2d(w_s + (size_t)(kc * 2) * W_TILE, &tmW, w_bar, kc * 32, cb0);
        ptx::tma_load_2d(w_s + (size_t)(kc * 2 + 1) * W_TILE, &tmWlo, w_bar, kc * 32, cb0);
      }
      int li = 0;
      for (int n = blockIdx.z; n < a.nb; n += gridDim.z) {
        for (int tt = 0; tt < ntiles; ++tt, ++li) {
          const int as = li & 1;
          ptx::mbar_wait(&a_empty[as], ((li >> 1) & 1) ^ 1);
          ptx::mbar_expect_tx(&a_full[as], (uint32_t)(KCH * FUSED_TILE_PIX * 128));
          const int pix0 = (n * HIN + iy0 + tt * R) * HIN;       // may be negative / beyond the tensor: TMA zero-fills
          for (int kc = 0; kc < KCH; ++kc)
            ptx::tma_load_2d(a_s + (size_t)as * SH.a_stage_bytes + (size_t)(kc * 2) * A_TILE, &tmX, &a_full[as], kc * 32, pix0);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer ==================================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CB >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      ptx::mbar_wait(w_bar, 0);
      ptx::tc_fence_after();
      int li = 0;
      for (int n = blockIdx.z; n < a.nb; n += gridDim.z) {
        for (int tt = 0; tt < ntiles; ++tt, ++li) {
          const int as = li & 1;
          const uint32_t use = (uint32_t)(li >> 1) & 1u;
          ptx::mbar_wait(&t_empty[as], use ^ 1);
          ptx::mbar_wait(&a_ready[as], use);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * 128);
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc) {
            const uint32_t a_addr = ptx::smem_u32(a_s + (size_t)as * SH.a_stage_bytes + (size_t)(kc * 2) * A_TILE);
            const uint32_t w_addr = ptx::smem_u32(w_s + (size_t)(kc * 2) * W_TILE);
            constexpr int krem_full = SH.Cin;
            const int krem = krem_full - kc * 32;
            const int ksteps = ((krem < 32 ? krem : 32) + 7) / 8;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t koff = (uint32_t)ks * 32u;
              const uint64_t ahi = umma_desc_sw128(a_addr + koff), alo = umma_desc_sw128(a_addr + A_TILE + koff);
              const uint64_t whi = umma_desc_sw128(w_addr + koff), wlo = umma_desc_sw128(w_addr + W_TILE + koff);
              ptx::mma_ss<true>(d_tmem, alo, whi, idesc, (uint32_t)((kc | ks) != 0));
              ptx::mma_ss<true>(d_tmem, ahi, wlo, idesc, 1u);
              ptx::mma_ss<true>(d_tmem, ahi, whi, idesc, 1u);
            }
          }
          ptx::mma_commit(&a_empty[as]);
          ptx::mma_commit(&t_full[as]);
        }
      }
    }
  } else if (warp >= WARP_TR0) {
    // ============================ operand transform ================================
    // one thread per tile row: lo = x - tf32(x); the raw tile stays in place as the hi operand
    const int r = tid - WARP_TR0 * 32;
    const uint32_t row_off = (uint32_t)r * 128u, xr = (uint32_t)(r & 7);
    int li = 0;
    for (int n = blockIdx.z; n < a.nb; n += gridDim.z) {
      for (int tt = 0; tt < ntiles; ++tt, ++li) {
        const int as = li & 1;
        ptx::mbar_wait(&a_full[as], (uint32_t)(li >> 1) & 1u);
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc) {
          const uint32_t a_hi = ptx::smem_u32(a_s + (size_t)as * SH.a_stage_bytes + (size_t)(kc * 2) * A_TILE) + row_off;
          const int nch = min(8, (SH.Cin - kc * 32) / 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t phys = a_hi + (((uint32_t)j ^ xr) << 4);
            uint4 lo = make_uint4(0u, 0u, 0u, 0u);
            if (j < nch) {
              const uint4 raw = ptx::lds128(phys);
              lo.x = __float_as_uint(__uint_as_float(raw.x) - __uint_as_float(raw.x & 0xFFFFE000u));
              lo.y = __float_as_uint(__uint_as_float(raw.y) - __uint_as_float(raw.y & 0xFFFFE000u));
              lo.z = __float_as_uint(__uint_as_float(raw.z) - __uint_as_float(raw.z & 0xFFFFE000u));
              lo.w = __float_as_uint(__uint_as_float(raw.w) - __uint_as_float(raw.w & 0xFFFFE000u));
            }
            ptx::sts128(phys + A_TILE, lo);
          }
        }
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&a_ready[as]);
      }
    }
  } else if (warp >= WARP_EPI0) {
    // ================================== epilogue ====================================
    // TMEM lane = pixel of the tile; pixel p lies in input row p / HIN of the tile, column p % HIN.
    const int p = (warp & 3) * 32 + lane;
    const bool pix_ok = p < FUSED_TILE_PIX;
    const int pr = pix_ok ? p / HIN : 0, px = pix_ok ? p - pr * HIN : 0;
    int li = 0;
    int g_rows = 0;                                       // ring rows produced before this (patch, tile)
    for (int n = blockIdx.z; n < a.nb; n += gridDim.z) {
      for (int tt = 0; tt < ntiles; ++tt, ++li) {
        const int as = li & 1;
        const int t_my = tt * R + pr;                     // row step of this pixel inside the band
        const bool row_ok = pix_ok && t_my < nsteps;
        const int g = g_rows + t_my;
        const int slot = g % STAGES;
        if (row_ok) ptx::mbar_wait(&r_empty[slot], (uint32_t)((g / STAGES) & 1) ^ 1u);
        ptx::mbar_wait(&t_full[as], (uint32_t)(li >> 1) & 1u);
        __syncwarp();                                     // lanes waited on different ring rows: reconverge for the aligned TMEM loads
        ptx::tc_fence_after();
        const int iy = iy0 + t_my;
        const bool in_img = iy >= 0 && iy < HIN;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(as * 128);
        const uint32_t dst = ptx::smem_u32(ring + (size_t)slot * ROW_BYTES) + (uint32_t)((PAD + px) * PITCH);
#pragma unroll
        for (int c0 = 0; c0 < CB; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld32x32b_x16(taddr + (uint32_t)c0, v);
          ptx::tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              uint32_t* op = &o.x;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = c0 + 4 * q + e;
                const float y = bn_silu<float>(__uint_as_float(v[4 * q + e]), sc_s[c], bi_s[c]);
                op[e] = in_img ? __float_as_uint(y) : 0u;
              }
              ptx::sts128(dst + (uint32_t)((c0 + 4 * q) * 4), o);
            }
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&t_empty[as]);
        if (row_ok) ptx::mbar_arrive(&r_full[slot]);
      }
      g_rows += nsteps;
    }
  } else {
    // ================================== consumers ===================================
    const bool active = tid < n_cons;
    const int cg = active ? tid % CGT : 0, sl = active ? tid / CGT : 0;
    const int c = cb0 + cg * 2;
    float2 w[K * K];
    float sc[2], bi[2];
#pragma unroll
    for (int i = 0; i < K * K; ++i) w[i] = make_float2(a.w_dw[(int64_t)i * C + c], a.w_dw[(int64_t)i * C + c + 1]);
    sc[0] = a.s_dw[c]; sc[1] = a.s_dw[c + 1];
    bi[0] = a.b_dw[c]; bi[1] = a.b_dw[c + 1];
    const int ox0 = sl * TW;
    const uint32_t ring_u32 = ptx::smem_u32(ring) + (uint32_t)(sl * TW * S * PITCH + cg * 8);
    int s = 0;
    uint32_t ph = 0;
    int pbuf = 0;
    constexpr int cons_threads = FUSED_CONS_WARPS * 32;
    for (int n = blockIdx.z; n < a.nb; n += gridDim.z) {
      float* out_n = a.out + ((int64_t)n * HOUT * HOUT + ox0) * C + c;
      float2 acc[NL][TW];
#pragma unroll
      for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int q = 0; q < TW; ++q) acc[l][q] = make_float2(0.f, 0.f);
      float psum[2] = {0.f, 0.f};
#pragma unroll 1
      for (int t0 = 0; t0 < nsteps; t0 += P) {
#pragma unroll
        for (int r = 0; r < P; ++r) {
          const int t = t0 + r;
          if (t < nsteps) {
            ptx::mbar_wait(&r_full[s], ph);
            const uint32_t rowbase = ring_u32 + (uint32_t)s * (uint32_t)ROW_BYTES;
            float2 v[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
              const uint2 raw = ptx::lds64(rowbase + (uint32_t)(j * PITCH));
              v[j] = make_float2(__uint_as_float(raw.x), __uint_as_float(raw.y));
            }
            // rows outside the image are zeros in the ring: no row test needed here
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
              if ((r - ky + P * 4) % S == 0) {
                const int slot = (((r - ky + P * 4) / S) % NL);
#pragma unroll
                for (int kx = 0; kx < K; ++kx)
#pragma unroll
                  for (int q = 0; q < TW; ++q) acc[slot][q] = __ffma2_rn(v[q * S + kx], w[ky * K + kx], acc[slot][q]);
              }
            }
            if (active) ptx::mbar_arrive(&r_empty[s]);
            if (++s == STAGES) {
              s = 0;
              ph ^= 1;
            }
            if ((r - (K - 1) + P * 4) % S == 0) {
              const int done = (((r - (K - 1) + P * 4) / S) % NL);
              const int td = t - (K - 1);
              const int oy = y0 + td / S;
              if (td >= 0 && oy < y1 && active) {
                float* orow = out_n + (int64_t)oy * HOUT * C;
#pragma unroll
                for (int q = 0; q < TW; ++q) {
                  if (ox0 + q < HOUT) {
                    const float u0 = bn_silu<float>(acc[done][q].x, sc[0], bi[0]);
                    const float u1 = bn_silu<float>(acc[done][q].y, sc[1], bi[1]);
                    psum[0] += u0;
                    psum[1] += u1;
                    *reinterpret_cast<float2*>(orow + q * C) = make_float2(u0, u1);
                  }
                }
              }
#pragma unroll
              for (int q = 0; q < TW; ++q) acc[done][q] = make_float2(0.f, 0.f);
            }
          }
        }
      }
      float* ps = pool_s + pbuf * PTC * CB;
      if (active) {
        ps[sl * CB + cg * 2] = psum[0];
        ps[sl * CB + cg * 2 + 1] = psum[1];
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
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
struct FusedLayer {
  bool present = false;
  int shape = -1;
  CUtensorMap tmW, tmWlo;
  const void* x_ptr[2] = {nullptr, nullptr};
  int64_t x_rows[2] = {0, 0};
  CUtensorMap tmX[2];
};

inline FusedShape fused_shape_of(int i) {
  return fused_make_shape(FUSED_KSCH[i][0], FUSED_KSCH[i][1], FUSED_KSCH[i][2], FUSED_KSCH[i][3], FUSED_KSCH[i][4]);
}

inline int fused_shape_index(int K, int S, int Cin, int C, int Hin) {
  for (int i = 0; i < 4; ++i)
    if (FUSED_KSCH[i][0] == K && FUSED_KSCH[i][1] == S && FUSED_KSCH[i][2] == Cin && FUSED_KSCH[i][3] == C && FUSED_KSCH[i][4] == Hin)
      return i;
  return -1;
}

// 2-D map over a row-major [rows][K] fp32 matrix with a {32, box_rows} box (128B swizzle, zero OOB fill)
inline int fused_make_map(CUtensorMap* map, const void* base, int64_t rows, int K, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled (fused) failed with CUresult " + std::to_string((int)r));
  return MC_OK;
}

// w_hi / w_lo: the expand layer's device weight copies of the pointwise plan ([C][Cin] fp32 hi / lo parts)
inline int fused_plan_layer(FusedLayer* l, int K, int S, int Cin, int C, int Hin, const void* w_hi, const void* w_lo) {
  l->shape = fused_shape_index(K, S, Cin, C, Hin);
  if (l->shape < 0 || !w_hi || !w_lo) return MC_OK;   // not a fusable block: the two-kernel path stays
  const FusedShape sh = fused_shape_of(l->shape);
  int rc;
  if ((rc = fused_make_map(&l->tmW, w_hi, C, Cin, sh.CB)) || (rc = fused_make_map(&l->tmWlo, w_lo, C, Cin, sh.CB))) return rc;
  l->present = true;
  return MC_OK;
}

template <int SHAPE>
inline int fused_launch_shape(FusedLayer& l, int slot, const FusedArgs& a, cudaStream_t st) {
  constexpr FusedShape sh = fused_make_shape(FUSED_KSCH[SHAPE][0], FUSED_KSCH[SHAPE][1], FUSED_KSCH[SHAPE][2],
                                             FUSED_KSCH[SHAPE][3], FUSED_KSCH[SHAPE][4]);
  static std::atomic<unsigned long long> attr_mask{0};
  if (first_use_on_device(attr_mask))
    MC_CUDA(cudaFuncSetAttribute(mbconv_fused_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, sh.smem));
  const int per_patch = sh.cz * sh.nbands;
  const int gz = std::max(1, std::min(a.nb, 3000 / per_patch));
  mbconv_fused_kernel<SHAPE><<<dim3(sh.cz, sh.nbands, gz), FUSED_THREADS, sh.smem, st>>>(l.tmX[slot], l.tmW, l.tmWlo, a);
  return MC_OK;
}

// x: block input [nb * Hin * Hin][Cin] fp32
inline int fused_launch(FusedLayer& l, const float* x, int Cin, int Hin, const FusedArgs& a, cudaStream_t st) {
  const int64_t rows = (int64_t)a.nb * Hin * Hin;
  int slot = -1;
  for (int i = 0; i < 2; ++i)
    if (l.x_ptr[i] == x && l.x_rows[i] == rows) slot = i;
  if (slot < 0) {
    slot = l.x_ptr[0] == nullptr || l.x_ptr[0] == x ? 0 : 1;
    if (int rc = fused_make_map(&l.tmX[slot], x, rows, Cin, FUSED_TILE_PIX)) return rc;
    l.x_ptr[slot] = x;
    l.x_rows[slot] = rows;
  }
  switch (l.shape) {
    case 0: return fused_launch_shape<0>(l, slot, a, st);
    case 1: return fused_launch_shape<1>(l, slot, a, st);
    case 2: return fused_launch_shape<2>(l, slot, a, st);
    case 3: return fused_launch_shape<3>(l, slot, a, st);
  }
  return fail(MC_ERR_UNSUPPORTED, "fused_launch: no such shape");
}

}  // namespace mc
