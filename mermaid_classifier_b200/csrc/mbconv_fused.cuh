// Expand 1x1 + BN + swish fused into the depthwise kernel (front MBConv blocks): the expanded map -- more than half of
// all HBM bytes of the network -- is produced and consumed inside one CTA and never touches global memory.
//
// dw_reg_kernel consumes a shared-memory ring of input rows ([column][channel slice], one row per stage) that a TMA
// producer fills from the expanded map in HBM.  Here the ring is filled by a small GEMM pipeline instead:
//
//   TMA warp         block-input tile x[R rows x W pixels = 112 pixels][C_in] (128B-swizzled, K zero-filled to one row)
//   transform warps  fp32 mode only: lo = x - tf32(x)                           (3xTF32 split, as pw_tc_kernel)
//   MMA warp         D[128 pixels x CB] = x * W_slice^T, W slice (BN scale folded in) resident in shared memory,
//                    D double-buffered in TMEM
//   epilogue warps   8 warps: two per TMEM lane quarter, each half of the slice's channels.  TMEM lane = pixel:
//                    tcgen05.ld -> + BN bias -> swish -> the pixel's channels into the ring row of its input row
//                    (zeros for rows outside the image: SAME padding pads the EXPANDED map)
//   consumer warps   dw_reg_kernel's loop: k x k taps out of the ring with weights in registers, BN + swish, store,
//                    SE pool partials
//
// The kernel is bound by the XU (MUFU) pipe: one swish per expanded element (ex2 + rcp in fp32, one tanh.approx in
// bf16) at 16 lanes/clk/SM.  That is why the epilogue is spread over two warps per scheduler (ncu on the first
// draft with one epilogue warp per scheduler: epilogue busy 92 %, everything else waiting on it, XU 30 %).
//
// All shapes have R * W = 112 pixels per tile (R = 1, 2, 2, 4 input rows).  Ring pixel pitch = CB * sizeof(T) + 16 bytes
// keeps the epilogue's 16-byte stores of neighbouring pixels on different banks.
#pragma once
#include "dw_tma.cuh"
#include "pw_tc.cuh"

namespace mc {

namespace ptx {
// 32 lanes x 8 consecutive fp32 columns.  No wait inside.
__device__ __forceinline__ void tmem_ld32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
}  // namespace ptx

struct FusedShape {
  int K, S, Cin, C, Hin, Hout, pad, TW, pt, CB, cgt, cz, bwin, pitch, row_bytes, stages, R, rows_per_band, nbands, kchunks;
  int nop, nh, a_stage_bytes, w_tile, w_bytes, smem;
};

//                                         K  S  Cin  C   Hin
constexpr int FUSED_KSCH[4][5] = {{3, 2, 16, 96, 112}, {3, 1, 24, 144, 56}, {5, 2, 24, 144, 56}, {5, 1, 40, 240, 28}};
constexpr int FUSED_MAX_STAGES = 12;
constexpr int FUSED_TILE_PIX = 112;
constexpr int FUSED_CONS_WARPS = 8, FUSED_EPI_WARPS = 8, FUSED_TR_WARPS = 2;
constexpr int FUSED_THREADS = (FUSED_CONS_WARPS + FUSED_TR_WARPS + FUSED_EPI_WARPS + 2) * 32;   // 640

__host__ __device__ constexpr FusedShape fused_make_shape(int K, int S, int Cin, int C, int Hin, int es) {
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
  d.pitch = d.CB * es + 16;
  d.row_bytes = (d.bwin * d.pitch + 127) / 128 * 128;
  d.R = FUSED_TILE_PIX / d.Hin;
  d.stages = 2 * d.R + 2 < 4 ? 4 : 2 * d.R + 2;
  d.rows_per_band = d.Hout >= 56 ? (d.S == 1 ? 28 : 14) : d.Hout;   // halo rows are recomputed: fewer bands at stride 1
  d.nbands = (d.Hout + d.rows_per_band - 1) / d.rows_per_band;
  const int kc = 128 / es;                                               // K elements per 128-byte operand row
  d.kchunks = (d.Cin + kc - 1) / kc;
  d.nop = es == 4 ? 2 : 1;                                               // operand copies: hi + lo in fp32 mode
  d.a_stage_bytes = d.kchunks * TC_BM * 128;                             // one operand copy of one tile
  d.w_tile = (d.CB * 128 + 1023) / 1024 * 1024;
  d.w_bytes = d.kchunks * d.nop * d.w_tile;
  // x tiles: a ring of NH stages filled by TMA (deep enough to cover the L2 / HBM latency), plus -- fp32 mode -- two
  // stages of the TF32 lo operand written by the transform warps
  d.nh = d.kchunks == 1 ? 4 : 3;
  const int other = 1024 + d.w_bytes + (d.nh + (es == 4 ? 2 : 0)) * d.a_stage_bytes + (2 * d.pt + 1) * d.CB * 4 + 512;
#ifdef MC_FUSED_DEEP_RING
  // experiment (-DMC_FUSED_DEEP_RING): as many ring rows as fit.  Measured: b1 / b3 unchanged, b2 8 % SLOWER (12 rows for 6):
  // ring depth is not what the fused kernels wait for
  {
    int fit = (227 * 1024 - other) / d.row_bytes;
    if (fit > FUSED_MAX_STAGES) fit = FUSED_MAX_STAGES;
    if (fit > d.stages) d.stages = fit;
  }
#endif
  d.smem = other + d.stages * d.row_bytes;
  return d;
}

struct FusedArgs {
  const float* w_dw;       // depthwise weights [K*K][C]
  const float* s_dw;       // depthwise BN
  const float* b_dw;
  const float* b_exp;      // expand BN bias (the scale is folded into the weight slice)
  void* out;               // [n][Hout][Hout][C]
  float* pool_partial;     // [n][nbands][C]
  int nb;
  int num_sms = 148;       // host side only: the launch takes one CTA per SM
};

template <typename T, int SHAPE>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
mbconv_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmWlo, const FusedArgs a) {
  constexpr int ES = (int)sizeof(T);
  constexpr bool F32 = ES == 4;
  constexpr FusedShape SH = fused_make_shape(FUSED_KSCH[SHAPE][0], FUSED_KSCH[SHAPE][1], FUSED_KSCH[SHAPE][2],
                                             FUSED_KSCH[SHAPE][3], FUSED_KSCH[SHAPE][4], ES);
  constexpr int K = SH.K, S = SH.S, TW = SH.TW, CB = SH.CB, C = SH.C, HIN = SH.Hin, HOUT = SH.Hout, PAD = SH.pad;
  constexpr int CGT = SH.cgt, PTC = SH.pt, RPB = SH.rows_per_band, STAGES = SH.stages, ROW_BYTES = SH.row_bytes;
  constexpr int PITCH = SH.pitch, R = SH.R, KCH = SH.kchunks, BWIN = SH.bwin, NOP = SH.nop;
  constexpr int KC = 128 / ES, UK = 32 / ES;            // K per operand row / per MMA
  constexpr int NL = (K + S - 1) / S, P = S * NL, NCOL = (TW - 1) * S + K;
  constexpr int n_cons = CGT * PTC;
  constexpr int A_TILE = TC_BM * 128, A_STAGE = SH.a_stage_bytes, NH = SH.nh;
  constexpr int W_TILE = SH.w_tile;
  static_assert(SH.Cin % 8 == 0, "C_in must be whole MMA k-steps");
  static_assert(R * HIN == FUSED_TILE_PIX, "a tile is R whole input rows of 112 pixels in total");
  static_assert(n_cons <= FUSED_CONS_WARPS * 32, "consumer threads");
  static_assert(STAGES <= FUSED_MAX_STAGES, "ring depth");
  static_assert(CB == 32 || CB == 48, "epilogue column passes");

  extern __shared__ uint8_t fused_smem_raw[];
  uint8_t* smem = fused_smem_raw + ((1024u - (ptx::smem_u32(fused_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* w_s = smem;                                   // [kchunk][hi, lo] W slice tiles
  uint8_t* a_s = w_s + SH.w_bytes;                       // [NH stages][kchunk] x tiles (TMA)
  uint8_t* lo_s = a_s + NH * A_STAGE;                    // [2 stages][kchunk] TF32 lo operand (fp32 mode)
  uint8_t* ring = lo_s + (F32 ? 2 : 0) * A_STAGE;        // [STAGES][ROW_BYTES]
  float* pool_s = (float*)(ring + STAGES * ROW_BYTES);   // [2][PTC][CB]
  float* bias_s = pool_s + 2 * PTC * CB;                 // [CB] expand BN bias of this slice
  uint64_t* bars = (uint64_t*)(bias_s + CB);
  uint64_t* r_full = bars;                               // [STAGES]  ring row written (HIN arrivals: one per pixel)
  uint64_t* r_empty = r_full + FUSED_MAX_STAGES;         // [STAGES]  ring row drained (n_cons arrivals)
  uint64_t* a_full = r_empty + FUSED_MAX_STAGES;         // [NH] x tile landed
  uint64_t* a_empty = a_full + 4;                        // [NH] MMAs reading the tile retired
  uint64_t* a_ready = a_empty + 4;                       // [2] lo operand written (fp32 mode)
  uint64_t* lo_empty = a_ready + 2;                      // [2] MMAs reading the lo operand retired
  uint64_t* t_full = lo_empty + 2;                       // [2] accumulator complete
  uint64_t* t_empty = t_full + 2;                        // [2] accumulator drained
  uint64_t* w_bar = t_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the warp arbiter favours the highest warp id: single-thread roles last, then the epilogue (the critical path)
  constexpr int WARP_TR0 = FUSED_CONS_WARPS;                  // 8, 9
  constexpr int WARP_EPI0 = WARP_TR0 + FUSED_TR_WARPS;        // 10..17: warp & 3 = TMEM lane quarter
  constexpr int WARP_MMA = WARP_EPI0 + FUSED_EPI_WARPS, WARP_TMA = WARP_MMA + 1;

  // Work unit = (patch, band) of this CTA's channel slice.  The grid is (slices, workers) with slices * workers <= the SM
  // count: one CTA per SM for the whole launch, so barrier / TMEM / weight set-up and the pipeline fill are paid once and not
  // once per unit.  Worker w takes units w, w + workers, ...; the CTAs of one worker (the slices) walk the same units at the
  // same time and share the x tiles in L2.
  constexpr int NBANDS = SH.nbands;
  const int cb0 = blockIdx.x * CB;
  const int worker = blockIdx.y, nworkers = gridDim.y;
  const int nunits = a.nb * NBANDS;
  struct Unit {
    int n, band, y0, y1, nsteps, ntiles, iy0;
  };
  auto unit_of = [&](int u) {
    Unit q;
    q.n = u / NBANDS;
    q.band = u - q.n * NBANDS;
    q.y0 = q.band * RPB;
    q.y1 = min(HOUT, q.y0 + RPB);
    q.nsteps = (q.y1 - 1 - q.y0) * S + K;                // input rows this band consumes
    q.ntiles = (q.nsteps + R - 1) / R;
    q.iy0 = q.y0 * S - PAD;
    return q;
  };

  for (int i = tid; i < CB; i += FUSED_THREADS) bias_s[i] = a.b_exp[cb0 + i] * (F32 ? 1.f : 0.5f);
  // pad columns of every ring row are zero for the whole launch (the epilogue only writes image columns)
  for (int i = tid; i < STAGES * (BWIN - HIN) * CB; i += FUSED_THREADS) {
    const int st = i / ((BWIN - HIN) * CB), rem = i - st * (BWIN - HIN) * CB;
    const int pc = rem / CB, c = rem - pc * CB;
    const int col = pc < PAD ? pc : pc + HIN;
    *(T*)(ring + (size_t)st * ROW_BYTES + (size_t)col * PITCH + c * ES) = from_f<T>(0.f);
  }
  if (tid == 0) {
    for (int s = 0; s < FUSED_MAX_STAGES; ++s) {
      ptx::mbar_init(&r_full[s], (uint32_t)HIN);
      ptx::mbar_init(&r_empty[s], (uint32_t)n_cons);
    }
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&a_ready[s], FUSED_TR_WARPS * 32);
      ptx::mbar_init(&lo_empty[s], 1);
      ptx::mbar_init(&t_full[s], 1);
      ptx::mbar_init(&t_empty[s], FUSED_EPI_WARPS / 2 * 32);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
    if (F32) ptx::prefetch_tmap(&tmWlo);
  }
  if (warp == WARP_MMA) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch (common.cuh): the producer loads the weight slice (constants) before it waits
  pdl_trigger();
  if (warp != WARP_TMA) pdl_wait();

  if (warp == WARP_TMA) {
    // ================================ TMA producer ================================
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_bar, (uint32_t)(KCH * NOP * CB * 128));
      for (int kc = 0; kc < KCH; ++kc) {
        ptx::tma_load_2d(w_s + (size_t)(kc * NOP) * W_TILE, &tmW, w_bar, kc * KC, cb0);
        if (F32) ptx::tma_load_2d(w_s + (size_t)(kc * NOP + 1) * W_TILE, &tmWlo, w_bar, kc * KC, cb0);
      }
      pdl_wait();
      int hs = 0;
      uint32_t hph = 0;
      for (int u = worker; u < nunits; u += nworkers) {
        const Unit q = unit_of(u);
        const int n = q.n, ntiles = q.ntiles, iy0 = q.iy0;
        for (int tt = 0; tt < ntiles; ++tt) {
          ptx::mbar_wait(&a_empty[hs], hph ^ 1);
          ptx::mbar_expect_tx(&a_full[hs], (uint32_t)(KCH * FUSED_TILE_PIX * 128));
          const int pix0 = (n * HIN + iy0 + tt * R) * HIN;       // may be negative / beyond the tensor: TMA zero-fills
          for (int kc = 0; kc < KCH; ++kc)
            ptx::tma_load_2d(a_s + (size_t)hs * A_STAGE + (size_t)kc * A_TILE, &tmX, &a_full[hs], kc * KC, pix0);
          if (++hs == NH) {
            hs = 0;
            hph ^= 1;
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer ==================================
    if (ptx::elect_one()) {
      constexpr uint32_t FMT = F32 ? 2u : 1u;   // TF32 / BF16
      constexpr uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(CB >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      ptx::mbar_wait(w_bar, 0);
      ptx::tc_fence_after();
      int li = 0, hs = 0;
      uint32_t hph = 0;
      for (int u = worker; u < nunits; u += nworkers) {
        const int ntiles = unit_of(u).ntiles;
        for (int tt = 0; tt < ntiles; ++tt, ++li) {
          const int as = li & 1;
          const uint32_t use = (uint32_t)(li >> 1) & 1u;
          ptx::mbar_wait(&t_empty[as], use ^ 1);
          if (F32) ptx::mbar_wait(&a_ready[as], use);      // implies a_full[hs]: the transform read the landed tile
          else ptx::mbar_wait(&a_full[hs], hph);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * 128);
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc) {
            const uint32_t a_addr = ptx::smem_u32(a_s + (size_t)hs * A_STAGE + (size_t)kc * A_TILE);
            const uint32_t l_addr = ptx::smem_u32(lo_s + (size_t)as * A_STAGE + (size_t)kc * A_TILE);
            const uint32_t w_addr = ptx::smem_u32(w_s + (size_t)(kc * NOP) * W_TILE);
            constexpr int krem_full = SH.Cin;
            const int krem = krem_full - kc * KC;
            const int ksteps = ((krem < KC ? krem : KC) + UK - 1) / UK;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t koff = (uint32_t)ks * 32u;
              const uint32_t acc = (uint32_t)((kc | ks) != 0);
              if constexpr (F32) {
                const uint64_t ahi = umma_desc_sw128(a_addr + koff), alo = umma_desc_sw128(l_addr + koff);
                const uint64_t whi = umma_desc_sw128(w_addr + koff), wlo = umma_desc_sw128(w_addr + W_TILE + koff);
                ptx::mma_ss<true>(d_tmem, alo, whi, idesc, acc);
                ptx::mma_ss<true>(d_tmem, ahi, wlo, idesc, 1u);
                ptx::mma_ss<true>(d_tmem, ahi, whi, idesc, 1u);
              } else {
                ptx::mma_ss<false>(d_tmem, umma_desc_sw128(a_addr + koff), umma_desc_sw128(w_addr + koff), idesc, acc);
              }
            }
          }
          ptx::mma_commit(&a_empty[hs]);
          if (F32) ptx::mma_commit(&lo_empty[as]);
          ptx::mma_commit(&t_full[as]);
          if (++hs == NH) {
            hs = 0;
            hph ^= 1;
          }
        }
      }
    }
  } else if (warp >= WARP_EPI0) {
    // ================================== epilogue ====================================
    // TMEM lane = pixel of the tile; pixel p lies in input row p / HIN of the tile, column p % HIN.
    // Two sets of four warps (one warp per lane quarter) ALTERNATE tiles: set 0 drains accumulator stage 0 (even tiles),
    // set 1 stage 1 (odd tiles).  The sets run out of phase, so the MUFU phase of one overlaps the TMEM load / ring
    // store / barrier phase of the other on every scheduler (in phase -- both halves of one tile at once, as the first
    // version did -- the XU pipe idled half of the time: ncu XU 51 %).
    const int quarter = warp & 3, set = (warp - WARP_EPI0) >> 2;
    const int p = quarter * 32 + lane;
    const bool pix_ok = p < FUSED_TILE_PIX;
    const int pr = pix_ok ? p / HIN : 0, px = pix_ok ? p - pr * HIN : 0;
    const uint32_t bias_u32 = ptx::smem_u32(bias_s);
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 128);
    int li = 0;
    int g_rows = 0;                                       // ring rows produced before this (patch, tile)
    uint32_t use = 0;                                     // phase of this set's accumulator stage
    for (int u = worker; u < nunits; u += nworkers) {
      const Unit q = unit_of(u);
      const int ntiles = q.ntiles, nsteps = q.nsteps, iy0 = q.iy0;
      for (int tt = 0; tt < ntiles; ++tt, ++li) {
        if ((li & 1) != set) continue;
        const int t_my = tt * R + pr;                     // row step of this pixel inside the band
        const bool row_ok = pix_ok && t_my < nsteps;
        const int g = g_rows + t_my;
        const int slot = g % STAGES;
        if (row_ok) ptx::mbar_wait(&r_empty[slot], (uint32_t)((g / STAGES) & 1) ^ 1u);
        ptx::mbar_wait(&t_full[set], use);
        use ^= 1u;
        __syncwarp();                                     // lanes waited on different ring rows: reconverge for the aligned TMEM loads
        ptx::tc_fence_after();
        uint32_t v[CB];
        {
          uint32_t(&v0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[0]);
          uint32_t(&v1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[16]);
          ptx::tmem_ld32x32b_x16(taddr0, v0);
          ptx::tmem_ld32x32b_x16(taddr0 + 16u, v1);
          if constexpr (CB == 48) {
            uint32_t(&v2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[32]);
            ptx::tmem_ld32x32b_x16(taddr0 + 32u, v2);
          }
          ptx::tmem_ld_wait();
        }
        // the accumulator is in registers: hand the TMEM stage back before the math
        ptx::tc_fence_before();
        ptx::mbar_arrive(&t_empty[set]);
        if (row_ok) {
          const int iy = iy0 + t_my;
          const uint32_t dst = ptx::smem_u32(ring + (size_t)slot * ROW_BYTES) + (uint32_t)((PAD + px) * PITCH);
          if (iy >= 0 && iy < HIN) {
#pragma unroll
            for (int c0 = 0; c0 < CB; c0 += 16) {
              // BN bias of 16 channels (the scale lives in the weights; bf16: pre-halved for the tanh form of swish)
              float bq[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 t4 = ptx::lds128(bias_u32 + (uint32_t)((c0 + 4 * q) * 4));
                bq[4 * q] = __uint_as_float(t4.x); bq[4 * q + 1] = __uint_as_float(t4.y);
                bq[4 * q + 2] = __uint_as_float(t4.z); bq[4 * q + 3] = __uint_as_float(t4.w);
              }
              float y[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float x = __uint_as_float(v[c0 + i]) + bq[i];
                if constexpr (F32) y[i] = __fdividef(x, 1.f + __expf(-x));
                else y[i] = fmaf(x, ptx::tanh_approx(x), x);
              }
              if constexpr (F32) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  ptx::sts128(dst + (uint32_t)((c0 + 4 * q) * 4), make_uint4(__float_as_uint(y[4 * q]), __float_as_uint(y[4 * q + 1]),
                                                                             __float_as_uint(y[4 * q + 2]), __float_as_uint(y[4 * q + 3])));
              } else {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  uint4 o;
                  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                  for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(y[8 * q + 2 * e], y[8 * q + 2 * e + 1]);
                  ptx::sts128(dst + (uint32_t)((c0 + 8 * q) * 2), o);
                }
              }
            }
          } else {
            // SAME padding pads the expanded map: rows outside the image are zeros, not swish(bias)
#pragma unroll
            for (int q = 0; q < CB * ES / 16; ++q) ptx::sts128(dst + (uint32_t)(q * 16), make_uint4(0u, 0u, 0u, 0u));
          }
          ptx::mbar_arrive(&r_full[slot]);
        }
      }
      g_rows += nsteps;
    }
  } else if (warp >= WARP_TR0) {
    // ============================ operand transform (fp32) ==========================
    // lo = x - tf32(x); the raw tile stays in place as the hi operand.  64 threads, two tile rows each.
    if constexpr (F32) {
      const int r0 = tid - WARP_TR0 * 32;
      int li = 0, hs = 0;
      uint32_t hph = 0;
      constexpr int NCH = SH.Cin / 4;                    // 16-byte chunks of a row that hold real data
      for (int u = worker; u < nunits; u += nworkers) {
        const int ntiles = unit_of(u).ntiles;
        for (int tt = 0; tt < ntiles; ++tt, ++li) {
          const int as = li & 1;
          ptx::mbar_wait(&lo_empty[as], ((uint32_t)(li >> 1) & 1u) ^ 1u);
          ptx::mbar_wait(&a_full[hs], hph);
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = r0 + rr * 64;
            if (r < FUSED_TILE_PIX) {
              const uint32_t row_off = (uint32_t)r * 128u, xr = (uint32_t)(r & 7);
              const uint32_t hi0 = ptx::smem_u32(a_s + (size_t)hs * A_STAGE) + row_off;
              const uint32_t lo0 = ptx::smem_u32(lo_s + (size_t)as * A_STAGE) + row_off;
              // all loads first: the inline-asm shared-memory accesses are volatile and keep their program order
              uint4 raw[NCH];
#pragma unroll
              for (int j = 0; j < NCH; ++j) raw[j] = ptx::lds128(hi0 + (uint32_t)((j >> 3) * A_TILE) + ((((uint32_t)j & 7u) ^ xr) << 4));
              // C_in is a multiple of 8 (one MMA k-step): the k-steps issued never read past chunk NCH - 1, no zero tail needed
#pragma unroll
              for (int j = 0; j < NCH; ++j) {
                uint4 lo;
                lo.x = tf32_lo_bits(raw[j].x);
                lo.y = tf32_lo_bits(raw[j].y);
                lo.z = tf32_lo_bits(raw[j].z);
                lo.w = tf32_lo_bits(raw[j].w);
                ptx::sts128(lo0 + (uint32_t)((j >> 3) * A_TILE) + ((((uint32_t)j & 7u) ^ xr) << 4), lo);
              }
            }
          }
          if (++hs == NH) {
            hs = 0;
            hph ^= 1;
          }
          ptx::fence_proxy_async();
          ptx::mbar_arrive(&a_ready[as]);
        }
      }
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
    const uint32_t ring_u32 = ptx::smem_u32(ring) + (uint32_t)(sl * TW * S * PITCH + cg * 2 * ES);
    int s = 0;
    uint32_t ph = 0;
    int pbuf = 0;
    constexpr int cons_threads = FUSED_CONS_WARPS * 32;
    for (int u = worker; u < nunits; u += nworkers) {
      const Unit q = unit_of(u);
      const int n = q.n, y0 = q.y0, y1 = q.y1, nsteps = q.nsteps;
      T* out_n = (T*)a.out + ((int64_t)n * HOUT * HOUT + ox0) * C + c;
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
              float t2[2];
              dw_load_ch<T, 2>(rowbase + (uint32_t)(j * PITCH), t2);
              v[j] = make_float2(t2[0], t2[1]);
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
                T* orow = out_n + (int64_t)oy * HOUT * C;
#pragma unroll
                for (int q = 0; q < TW; ++q) {
                  if (ox0 + q < HOUT) {
                    float u[2];
                    u[0] = bn_silu<T>(acc[done][q].x, sc[0], bi[0]);
                    u[1] = bn_silu<T>(acc[done][q].y, sc[1], bi[1]);
                    psum[0] += u[0];
                    psum[1] += u[1];
                    dw_store_ch<T, 2>(orow + q * C, u);
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
        a.pool_partial[((int64_t)n * NBANDS + q.band) * C + cb0 + i] = sum;
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
  void* d_w = nullptr;     // expand weight with the BN scale folded in: bf16, or the TF32 hi part
  void* d_wlo = nullptr;   // TF32 lo part (fp32 mode)
  CUtensorMap tmW, tmWlo;
  const void* x_ptr[2] = {nullptr, nullptr};
  int64_t x_rows[2] = {0, 0};
  CUtensorMap tmX[2];
};

inline void fused_free(FusedLayer& l) {
  if (l.d_w) cudaFree(l.d_w);
  if (l.d_wlo) cudaFree(l.d_wlo);
  l.d_w = l.d_wlo = nullptr;
  l.present = false;
}

inline FusedShape fused_shape_of(int i, int es) {
  return fused_make_shape(FUSED_KSCH[i][0], FUSED_KSCH[i][1], FUSED_KSCH[i][2], FUSED_KSCH[i][3], FUSED_KSCH[i][4], es);
}

inline int fused_shape_index(int K, int S, int Cin, int C, int Hin) {
  for (int i = 0; i < 4; ++i)
    if (FUSED_KSCH[i][0] == K && FUSED_KSCH[i][1] == S && FUSED_KSCH[i][2] == Cin && FUSED_KSCH[i][3] == C && FUSED_KSCH[i][4] == Hin)
      return i;
  return -1;
}

// `w_host` [C][Cin], `scale_host` [C]: the expand conv and its folded BN scale.  The product is rounded to fp32 once, then
// split into TF32 hi / lo parts (fp32 mode) or rounded to bf16 (bf16 mode, with the 1/2 of the tanh form of swish).
inline int fused_plan_layer(FusedLayer* l, bool f32, int K, int S, int Cin, int C, int Hin, const float* w_host,
                            const float* scale_host) {
  l->shape = fused_shape_index(K, S, Cin, C, Hin);
  if (l->shape < 0) return MC_OK;   // not a fusable block: the two-kernel path stays
  const FusedShape sh = fused_shape_of(l->shape, f32 ? 4 : 2);
  if (sh.smem > TC_SMEM_BUDGET) return MC_OK;   // does not fit (b4 in fp32 mode): the two-kernel path stays
  const size_t n = (size_t)C * Cin;
  int rc;
  if (f32) {
    std::vector<float> hi(n), lo(n);
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < Cin; ++k) {
        const float wf = w_host[(size_t)c * Cin + k] * scale_host[c];
        uint32_t bits, bits0;
        memcpy(&bits, &wf, 4);
        bits0 = bits;
        bits &= 0xFFFFE000u;
        float h;
        memcpy(&h, &bits, 4);
        hi[(size_t)c * Cin + k] = h;
        const uint32_t lb = tf32_lo_bits(bits0);
        memcpy(&lo[(size_t)c * Cin + k], &lb, 4);
      }
    MC_CUDA(cudaMalloc(&l->d_w, n * 4));
    MC_CUDA(cudaMalloc(&l->d_wlo, n * 4));
    MC_CUDA(cudaMemcpy(l->d_w, hi.data(), n * 4, cudaMemcpyHostToDevice));
    MC_CUDA(cudaMemcpy(l->d_wlo, lo.data(), n * 4, cudaMemcpyHostToDevice));
    if ((rc = make_map(&l->tmW, true, l->d_w, C, Cin, sh.CB)) || (rc = make_map(&l->tmWlo, true, l->d_wlo, C, Cin, sh.CB))) return rc;
  } else {
    std::vector<__nv_bfloat16> wb(n);
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < Cin; ++k) wb[(size_t)c * Cin + k] = __float2bfloat16_rn(w_host[(size_t)c * Cin + k] * scale_host[c] * 0.5f);
    MC_CUDA(cudaMalloc(&l->d_w, n * 2));
    MC_CUDA(cudaMemcpy(l->d_w, wb.data(), n * 2, cudaMemcpyHostToDevice));
    if ((rc = make_map(&l->tmW, false, l->d_w, C, Cin, sh.CB))) return rc;
    l->tmWlo = l->tmW;
  }
  l->present = true;
  return MC_OK;
}

template <typename T, int SHAPE>
inline int fused_launch_shape(FusedLayer& l, int slot, const FusedArgs& a, cudaStream_t st) {
  constexpr FusedShape sh = fused_make_shape(FUSED_KSCH[SHAPE][0], FUSED_KSCH[SHAPE][1], FUSED_KSCH[SHAPE][2],
                                             FUSED_KSCH[SHAPE][3], FUSED_KSCH[SHAPE][4], (int)sizeof(T));
  if constexpr (sh.smem > TC_SMEM_BUDGET) {
    return fail(MC_ERR_UNSUPPORTED, "fused_launch: shape does not fit in shared memory");   // fused_plan_layer never selects it
  } else {
    static std::atomic<unsigned long long> attr_mask{0};
    if (first_use_on_device(attr_mask))
      MC_CUDA(cudaFuncSetAttribute(mbconv_fused_kernel<T, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, sh.smem));
    // one CTA per SM (the kernel's shared memory allows no second one): slices x workers <= SMs.  MC_FUSED_WORKERS overrides
    // the workers per slice (experiments: 1000 ~ the earlier one-or-two-units-per-CTA grid)
    static const int workers_env = getenv("MC_FUSED_WORKERS") ? atoi(getenv("MC_FUSED_WORKERS")) : 0;
    int workers = workers_env > 0 ? workers_env : std::max(1, a.num_sms / sh.cz);
    workers = std::max(1, std::min(workers, a.nb * sh.nbands));
    MC_CUDA(launch_pdl(PDL_FUSED, mbconv_fused_kernel<T, SHAPE>, dim3(sh.cz, workers), dim3(FUSED_THREADS), sh.smem, st, l.tmX[slot], l.tmW, l.tmWlo, a));
    MC_CHECK_LAUNCH();
    return MC_OK;
  }
}

// x: block input [nb * Hin * Hin][Cin]
template <typename T>
inline int fused_launch(FusedLayer& l, const T* x, int Cin, int Hin, const FusedArgs& a, cudaStream_t st) {
  const int64_t rows = (int64_t)a.nb * Hin * Hin;
  int slot = -1;
  for (int i = 0; i < 2; ++i)
    if (l.x_ptr[i] == x && l.x_rows[i] == rows) slot = i;
  if (slot < 0) {
    slot = l.x_ptr[0] == nullptr || l.x_ptr[0] == x ? 0 : 1;
    if (int rc = make_map(&l.tmX[slot], sizeof(T) == 4, x, rows, Cin, FUSED_TILE_PIX)) return rc;
    l.x_ptr[slot] = x;
    l.x_rows[slot] = rows;
  }
  switch (l.shape) {
    case 0: return fused_launch_shape<T, 0>(l, slot, a, st);
    case 1: return fused_launch_shape<T, 1>(l, slot, a, st);
    case 2: return fused_launch_shape<T, 2>(l, slot, a, st);
    case 3: return fused_launch_shape<T, 3>(l, slot, a, st);
  }
  return fail(MC_ERR_UNSUPPORTED, "fused_launch: no such shape");
}

}  // namespace mc
