// K1 + K2 on the tensor cores: reflect-crop gather, normalisation and the 3x3 / stride-2 stem conv as an implicit GEMM.
//
// Reference semantics (oracle/crop.py, oracle/effnet.py): patch[i][j][c] = img[R(row-112+i)][R(col-112+j)][c]
// (np.pad mode='reflect'), x = (u8/255 - mean)/std, stem = swish(bn0(conv3x3 s2, TF-SAME pad (0,1))).
//
// The normalisation is affine per input channel, so it folds into the conv:
//   conv(W, (u8/255 - m)/s) = sum_taps W' * u8  -  C,     W' = W / (255 s_ci),  C = sum_{taps inside the patch} W m_ci / s_ci
// The left term is a GEMM whose A operand holds RAW BYTES: 0..255 are exact in bf16, so A needs no hi/lo split and half the
// shared-memory bytes of fp32.  W' is split on the host into three bf16 parts (24 mantissa bits, products with an 8-bit
// integer are exact), accumulated in fp32 by three kind::f16 MMAs per k-step: fp32-class arithmetic on the bf16 pipe.
// The SAME zero padding (patch row / column 224) pads the NORMALISED input, so a padded tap must drop out of C as well:
// four bias vectors (interior, last column, last row, corner) cover it.
//
// One persistent CTA per SM walks (patch, band of 16 output rows) items:
//   builder warps (8)  stage the band's 33 source rows in shared memory with 16-byte coalesced loads (rows that reflect
//                      in x: per-pixel gathers), then write im2col rows -- 27 taps of one output pixel, bf16, K padded to
//                      32 -- straight into the 128B-swizzled A tiles (128 output pixels each)
//   MMA warp           six tcgen05.mma (M 128, N 32, K 16) per tile into one of two TMEM accumulator stages
//   epilogue warps (8) two sets alternating tiles: tcgen05.ld -> BN + swish -> NHWC stores (a quad writes 64 B of a pixel)
// The next band's source rows are loaded into registers before the current band is built and stored afterwards, so the
// HBM latency of the gather hides behind the im2col work.
#pragma once
#include "crop_stem.cuh"
#include "pw_tc.cuh"

namespace mc {

constexpr int STEM_BUILD_WARPS = 8, STEM_EPI_WARPS = 8;
constexpr int STEM_TC_THREADS = (STEM_BUILD_WARPS + STEM_EPI_WARPS + 1) * 32;   // 544
constexpr int STEM_BAND = 16;                        // output rows per item
constexpr int STEM_IN_ROWS = 2 * STEM_BAND + 1;      // 33 source rows
constexpr int STEM_ROW_PITCH = 720;                  // bytes per staged row: 43 chunks of 16 (672 + up to 15 bytes of misalignment)
constexpr int STEM_ROW_CHUNKS = 43;
constexpr int STEM_TILES = STEM_BAND * 112 / 128;    // 14 tiles of 128 output pixels per band
constexpr int STEM_A_STAGES = 4;
constexpr int STEM_PRE = (STEM_IN_ROWS * STEM_ROW_CHUNKS + STEM_BUILD_WARPS * 32 - 1) / (STEM_BUILD_WARPS * 32);   // 6 chunks per builder thread
constexpr int STEM_SMEM = 1024 + 3 * 4096 + STEM_A_STAGES * TC_BM * 128 + 2 * (STEM_IN_ROWS * STEM_ROW_PITCH + 160) + 5 * 32 * 4 + 256;

struct StemTcArgs {
  const __nv_bfloat16* w_parts;   // [3][32][32]: hi / mid / lo bf16 parts of W'[co][k], k = (ky*3 + kx)*3 + ci, zero for k >= 27
  const float* scale;             // [32] folded BN scale
  const float* bias_v;            // [4][32] folded BN bias minus scale * C for: interior, ox == 111, oy == 111, both
  void* out;                      // [n][112][112][32]
  int nb;
};

template <typename T>
__global__ void __launch_bounds__(STEM_TC_THREADS, 1)
stem_tc_kernel(const mc_image* __restrict__ images, const mc_point* __restrict__ points, const StemTcArgs a) {
  extern __shared__ uint8_t stem_smem_raw[];
  uint8_t* smem = stem_smem_raw + ((1024u - (ptx::smem_u32(stem_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* w_s = smem;                                             // 3 parts x [32 rows x 128 B] (first 64 B of a row used)
  uint8_t* a_s = w_s + 3 * 4096;                                   // [4 stages][128 rows x 128 B] (first 64 B of a row used)
  uint8_t* rows_s = a_s + STEM_A_STAGES * TC_BM * 128;             // [2][33][720] staged source bytes
  int* roff_s = (int*)(rows_s + 2 * STEM_IN_ROWS * STEM_ROW_PITCH);   // [2][40] byte offset of patch column 0 inside a staged row
  float* sc_s = (float*)(roff_s + 80);                             // [32]
  float* bv_s = sc_s + 32;                                         // [4][32]
  uint64_t* bars = (uint64_t*)(bv_s + 128);
  uint64_t* a_full = bars;                                         // [4] im2col tile written (128 builder arrivals)
  uint64_t* a_empty = a_full + STEM_A_STAGES;                      // [4] MMAs reading the tile retired
  uint64_t* t_full = a_empty + STEM_A_STAGES;                      // [2] accumulator complete
  uint64_t* t_empty = t_full + 2;                                  // [2] accumulator drained (128 arrivals)
  uint32_t* tmem_slot = (uint32_t*)(t_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int WARP_EPI0 = STEM_BUILD_WARPS, WARP_MMA = STEM_BUILD_WARPS + STEM_EPI_WARPS;
  const int items = a.nb * (112 / STEM_BAND);

  // W' parts into the swizzled operand layout (row = output channel, 16-byte chunk c of the row at (c ^ (row & 7)) << 4)
  for (int i = tid; i < 3 * 32 * 4; i += STEM_TC_THREADS) {
    const int part = i / 128, row = (i >> 2) & 31, c = i & 3;
    const uint4 v = *reinterpret_cast<const uint4*>(a.w_parts + (part * 32 + row) * 32 + c * 8);
    *reinterpret_cast<uint4*>(w_s + part * 4096 + row * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  // bf16 swish is x * sigmoid(x) = h + h * tanh(h), h = x / 2: fold the 1/2 into scale and bias
  const float fold = (sizeof(T) == 2 && MC_BF16_TANH) ? 0.5f : 1.f;
  for (int i = tid; i < 32; i += STEM_TC_THREADS) sc_s[i] = a.scale[i] * fold;
  for (int i = tid; i < 128; i += STEM_TC_THREADS) bv_s[i] = a.bias_v[i] * fold;
  if (tid == 0) {
    for (int s = 0; s < STEM_A_STAGES; ++s) {
      ptx::mbar_init(&a_full[s], 128);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&t_full[s], 1);
      ptx::mbar_init(&t_empty[s], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == WARP_MMA) ptx::tmem_alloc(tmem_slot, 64);
  ptx::fence_proxy_async();   // the W' tiles were written through the generic proxy; the MMAs read them through the async proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();   // programmatic dependent launch (common.cuh): weights, scale and bias above are constants
  pdl_wait();      // the output buffer may still be read by the previous sub-batch's kernels

  if (warp < STEM_BUILD_WARPS) {
    // ================================== builders ====================================
    constexpr int NB_THREADS = STEM_BUILD_WARPS * 32;
    const int grp = tid >> 7, pl = tid & 127;      // tile parity this thread builds; its row of the A tile
    // geometry of an item's source rows
    struct RowGeo {
      const uint8_t* data;
      int64_t pitch;
      const uint8_t* end;     // one past the last byte of the image
      int H, W, row0, x0;
      bool fast;
    };
    auto row_src = [&](const RowGeo& g, int r) {
      return g.data + (int64_t)reflect_fast(g.row0 + r, g.H) * g.pitch + (int64_t)g.x0 * 3;
    };
    auto row_safe = [&](const RowGeo& g, const uint8_t* src) {
      const uintptr_t a0 = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15;
      return a0 >= reinterpret_cast<uintptr_t>(g.data) && a0 + STEM_ROW_CHUNKS * 16 <= reinterpret_cast<uintptr_t>(g.end);
    };
    auto geo_of = [&](int it) {
      RowGeo g;
      const int n = it / (112 / STEM_BAND), band = it - n * (112 / STEM_BAND);
      const mc_point pt = points[n];
      const mc_image im = images[pt.image];
      g.data = im.data;
      g.pitch = im.row_pitch;
      g.H = im.height;
      g.W = im.width;
      g.row0 = pt.row - 112 + 2 * STEM_BAND * band;   // image row of staged row 0 (before reflection)
      g.x0 = pt.col - 112;
      // fast items: no reflection in x, the 224 patch columns of a row are 672 contiguous bytes of the image.  The aligned
      // 16-byte chunks that cover them reach up to 15 bytes before and 31 bytes past the row segment: a row whose chunks would
      // leave the image buffer (first / last image row) is gathered per pixel instead (row_safe)
      g.fast = g.x0 >= 0 && g.x0 + 224 <= g.W;
      g.end = g.data + (int64_t)(g.H - 1) * g.pitch + (int64_t)g.W * 3;
      return g;
    };
    // staged rows of an item: r = 0..32, patch row 32*band + r; patch row 224 is the SAME pad (never staged, never read)
    auto n_rows_of = [&](int it) { return (it % (112 / STEM_BAND)) == 112 / STEM_BAND - 1 ? STEM_IN_ROWS - 1 : STEM_IN_ROWS; };
    uint4 pre[STEM_PRE];
    auto prefetch = [&](const RowGeo& g, int nrows) {
#pragma unroll
      for (int q = 0; q < STEM_PRE; ++q) {
        const int idx = tid + q * NB_THREADS;
        const int r = idx / STEM_ROW_CHUNKS, c = idx - r * STEM_ROW_CHUNKS;
        pre[q] = make_uint4(0u, 0u, 0u, 0u);
        if (g.fast && r < nrows) {
          const uint8_t* src = row_src(g, r);
          const uintptr_t a0 = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15;
          if (row_safe(g, src)) pre[q] = __ldg(reinterpret_cast<const uint4*>(a0) + c);
        }
      }
    };
    auto commit = [&](const RowGeo& g, int nrows, int buf) {
      uint8_t* rb = rows_s + buf * STEM_IN_ROWS * STEM_ROW_PITCH;
      if (g.fast) {
#pragma unroll
        for (int q = 0; q < STEM_PRE; ++q) {
          const int idx = tid + q * NB_THREADS;
          const int r = idx / STEM_ROW_CHUNKS, c = idx - r * STEM_ROW_CHUNKS;
          if (r < nrows && row_safe(g, row_src(g, r))) *reinterpret_cast<uint4*>(rb + r * STEM_ROW_PITCH + c * 16) = pre[q];
        }
        // rows whose aligned chunks would leave the image buffer: per-byte copy of the 672 bytes (one warp per such row)
        for (int r = tid >> 5; r < nrows; r += NB_THREADS / 32) {
          const uint8_t* src = row_src(g, r);
          if (!row_safe(g, src)) {
            const int o = (int)(reinterpret_cast<uintptr_t>(src) & 15);
            for (int b = tid & 31; b < 672; b += 32) rb[r * STEM_ROW_PITCH + o + b] = src[b];
          }
        }
        if (tid < nrows) roff_s[buf * 40 + tid] = (int)(reinterpret_cast<uintptr_t>(row_src(g, tid)) & 15);
      } else {
        // the window crosses the left / right image border: per-pixel reflect gather
        for (int idx = tid; idx < nrows * 224; idx += NB_THREADS) {
          const int r = idx / 224, j = idx - r * 224;
          const uint8_t* s = g.data + (int64_t)reflect_fast(g.row0 + r, g.H) * g.pitch + (int64_t)reflect_fast(g.x0 + j, g.W) * 3;
          uint8_t* d = rb + r * STEM_ROW_PITCH + j * 3;
          d[0] = s[0];
          d[1] = s[1];
          d[2] = s[2];
        }
        if (tid < nrows) roff_s[buf * 40 + tid] = 0;
      }
    };
    int it = blockIdx.x, k_it = 0;
    if (it < items) {
      const RowGeo g0 = geo_of(it);
      prefetch(g0, n_rows_of(it));
      commit(g0, n_rows_of(it), 0);
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NB_THREADS) : "memory");
    for (; it < items; it += gridDim.x, ++k_it) {
      const int buf = k_it & 1;
      const int it_n = it + gridDim.x;
      RowGeo gn{};
      int nrows_n = 0;
      if (it_n < items) {
        gn = geo_of(it_n);
        nrows_n = n_rows_of(it_n);
        prefetch(gn, nrows_n);
      }
      const bool last_band = (it % (112 / STEM_BAND)) == 112 / STEM_BAND - 1;
      const uint32_t rb_u32 = ptx::smem_u32(rows_s + buf * STEM_IN_ROWS * STEM_ROW_PITCH);
      const uint32_t ro_u32 = ptx::smem_u32(roff_s + buf * 40);
#pragma unroll 1
      for (int j = grp; j < STEM_TILES; j += 2) {
        const int li = k_it * STEM_TILES + j;          // tile counter of this CTA
        const int st = li & (STEM_A_STAGES - 1);
        const int p = j * 128 + pl;                   // output pixel of the band
        const int oy = p / 112, ox = p - oy * 112;
        uint32_t f[27];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int r = 2 * oy + ky;
          uint32_t b0 = 0u, b1 = 0u, b2 = 0u;          // bytes 0..3, 4..7, 8 of the 3 pixels x 3 channels under this tap row
          if (!(last_band && r == STEM_IN_ROWS - 1)) { // patch row 224 = zero pad
            const int o = 6 * ox + (int)ptx::lds32(ro_u32 + (uint32_t)(r * 4));
            const uint32_t wa = rb_u32 + (uint32_t)(r * STEM_ROW_PITCH + (o & ~3));
            const uint32_t w0 = ptx::lds32(wa), w1 = ptx::lds32(wa + 4), w2 = ptx::lds32(wa + 8);
            const uint32_t sh = (uint32_t)(o & 3) * 8u;
            b0 = __funnelshift_r(w0, w1, sh);
            b1 = __funnelshift_r(w1, w2, sh);
            b2 = w2 >> sh;
            if (ox == 111) {                           // patch column 224 = zero pad: taps kx == 2 (bytes 6, 7, 8)
              b1 &= 0x0000FFFFu;
              b2 = 0u;
            }
          }
          // byte v -> float(v): 0x4B0000vv is 8388608 + v
#pragma unroll
          for (int e = 0; e < 4; ++e) f[9 * ky + e] = __float_as_uint(__uint_as_float(__byte_perm(b0, 0x4B000000u, 0x7440 + e)) - 8388608.f);
#pragma unroll
          for (int e = 0; e < 4; ++e) f[9 * ky + 4 + e] = __float_as_uint(__uint_as_float(__byte_perm(b1, 0x4B000000u, 0x7440 + e)) - 8388608.f);
          f[9 * ky + 8] = __float_as_uint(__uint_as_float(__byte_perm(b2, 0x4B000000u, 0x7440)) - 8388608.f);
        }
        // bf16 of an integer <= 255 = the upper half of its fp32 pattern; pairs packed k even low, k odd high
        uint32_t w16[16];
#pragma unroll
        for (int i = 0; i < 13; ++i) w16[i] = __byte_perm(f[2 * i], f[2 * i + 1], 0x7632);
        w16[13] = f[26] >> 16;
        w16[14] = 0u;
        w16[15] = 0u;
        ptx::mbar_wait(&a_empty[st], (uint32_t)((li >> 2) & 1) ^ 1u);
        const uint32_t dst = ptx::smem_u32(a_s + st * (TC_BM * 128)) + (uint32_t)pl * 128u;
        const uint32_t xr = (uint32_t)(pl & 7);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          ptx::sts128(dst + (((uint32_t)c ^ xr) << 4), make_uint4(w16[4 * c], w16[4 * c + 1], w16[4 * c + 2], w16[4 * c + 3]));
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&a_full[st]);
      }
      if (it_n < items) commit(gn, nrows_n, buf ^ 1);
      asm volatile("bar.sync 1, %0;" ::"r"(NB_THREADS) : "memory");
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer ==================================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t w_addr = ptx::smem_u32(w_s);
      int li = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        for (int j = 0; j < STEM_TILES; ++j, ++li) {
          const int st = li & (STEM_A_STAGES - 1), ts = li & 1;
          ptx::mbar_wait(&t_empty[ts], (uint32_t)((li >> 1) & 1) ^ 1u);
          ptx::mbar_wait(&a_full[st], (uint32_t)(li >> 2) & 1u);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(a_s + st * (TC_BM * 128));
          const uint32_t d_tmem = tmem_base + (uint32_t)(ts * 32);
#pragma unroll
          for (int part = 0; part < 3; ++part)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              ptx::mma_ss<false>(d_tmem, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(w_addr + part * 4096 + ks * 32), idesc,
                                 (uint32_t)((part | ks) != 0));
          ptx::mma_commit(&a_empty[st]);
          ptx::mma_commit(&t_full[ts]);
        }
      }
    }
  } else {
    // ================================== epilogue ====================================
    // Accumulator in the mma fragment layout (tcgen05.ld 16x256b), one lane exchange so a thread owns four consecutive
    // channels of a pixel: a quad then writes 64 contiguous bytes of the pixel's 128-byte NHWC line.  (A thread-per-pixel
    // layout needs no exchange but its 16-byte stores land 128 bytes apart: measured 25 % slower.)
    const int set = (warp - WARP_EPI0) >> 2, quarter = warp & 3;
    const int lr = lane >> 2, q = lane & 3;
    const bool odd = (q & 1) != 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 32);
    const uint32_t sc_u32 = ptx::smem_u32(sc_s), bv_u32 = ptx::smem_u32(bv_s);
    // this thread's columns in the two 16-column halves, their scale and the interior bias (variant 0)
    const int cb0 = odd ? 8 + 2 * (q - 1) : 2 * q;
    float sc[2][4], b0v[2][4];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const uint4 s4 = ptx::lds128(sc_u32 + (uint32_t)((16 * pr + cb0) * 4));
      const uint4 b4 = ptx::lds128(bv_u32 + (uint32_t)((16 * pr + cb0) * 4));
      sc[pr][0] = __uint_as_float(s4.x); sc[pr][1] = __uint_as_float(s4.y); sc[pr][2] = __uint_as_float(s4.z); sc[pr][3] = __uint_as_float(s4.w);
      b0v[pr][0] = __uint_as_float(b4.x); b0v[pr][1] = __uint_as_float(b4.y); b0v[pr][2] = __uint_as_float(b4.z); b0v[pr][3] = __uint_as_float(b4.w);
    }
    uint32_t use = 0;
    int li = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int n = it / (112 / STEM_BAND), band = it - n * (112 / STEM_BAND);
      T* out_b = (T*)a.out + ((int64_t)n * 12544 + (int64_t)band * STEM_BAND * 112) * 32;
      for (int j = 0; j < STEM_TILES; ++j, ++li) {
        if ((li & 1) != set) continue;
        ptx::mbar_wait(&t_full[set], use);
        use ^= 1u;
        ptx::tc_fence_after();
        uint32_t v[2][16];
        ptx::tmem_ld16x256b_x4(taddr, v[0]);
        ptx::tmem_ld16x256b_x4(taddr + (16u << 16), v[1]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&t_empty[set]);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const int p = j * 128 + quarter * 32 + 16 * h2 + 8 * rh + lr;       // output pixel of the band
            const int oy = p / 112, ox = p - oy * 112;
            const int variant = (ox == 111 ? 1 : 0) | ((band == 112 / STEM_BAND - 1 && oy == STEM_BAND - 1) ? 2 : 0);
            T* o = out_b + (int64_t)p * 32 + cb0;
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              const uint32_t a0 = v[h2][8 * pr + 2 * rh], a1 = v[h2][8 * pr + 2 * rh + 1];
              const uint32_t b0 = v[h2][8 * pr + 4 + 2 * rh], b1 = v[h2][8 * pr + 4 + 2 * rh + 1];
              const uint32_t s0 = odd ? a0 : b0, s1 = odd ? a1 : b1;
              const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
              float y[4];
              y[0] = __uint_as_float(odd ? r0 : a0);
              y[1] = __uint_as_float(odd ? r1 : a1);
              y[2] = __uint_as_float(odd ? b0 : r0);
              y[3] = __uint_as_float(odd ? b1 : r1);
              float bb[4] = {b0v[pr][0], b0v[pr][1], b0v[pr][2], b0v[pr][3]};
              if (variant != 0) {   // last column / last row of the patch: taps on the SAME pad drop out of the constant term
                const uint4 b4 = ptx::lds128(bv_u32 + (uint32_t)((variant * 32 + 16 * pr + cb0) * 4));
                bb[0] = __uint_as_float(b4.x); bb[1] = __uint_as_float(b4.y); bb[2] = __uint_as_float(b4.z); bb[3] = __uint_as_float(b4.w);
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) y[e] = fmaf(y[e], sc[pr][e], bb[e]);
              if (sizeof(T) == 4 || !MC_BF16_TANH) {
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = __fdividef(y[e], 1.f + __expf(-y[e]));
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = fmaf(y[e], ptx::tanh_approx(y[e]), y[e]);
              }
              store4<T>(o + 16 * pr, y);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 64);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
struct StemTcPlan {
  __nv_bfloat16* d_w = nullptr;   // [3][32][32]
  float* d_bias_v = nullptr;      // [4][32]
  int num_sms = 148;
};

inline void stem_tc_free(StemTcPlan& p) {
  if (p.d_w) cudaFree(p.d_w);
  if (p.d_bias_v) cudaFree(p.d_bias_v);
  p.d_w = nullptr;
  p.d_bias_v = nullptr;
}

// w_stem [27][32] (k = (ky*3 + kx)*3 + ci major, output channel minor), scale / bias [32]: the packed-blob segments.
inline int stem_tc_plan(StemTcPlan* p, const float* w_stem, const float* scale, const float* bias, int device) {
  const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
  std::vector<__nv_bfloat16> parts(3 * 32 * 32, __float2bfloat16_rn(0.f));
  std::vector<float> bias_v(4 * 32);
  for (int co = 0; co < 32; ++co) {
    double c_var[4] = {0, 0, 0, 0};
    for (int k = 0; k < 27; ++k) {
      const int ci = k % 3, kx = (k / 3) % 3, ky = k / 9;
      const double w = (double)w_stem[k * 32 + co];
      float rem = (float)(w / (255.0 * stdv[ci]));
      for (int part = 0; part < 3; ++part) {
        // truncate to bf16 (the remainder stays exact); the last part is rounded
        uint32_t bits;
        memcpy(&bits, &rem, 4);
        __nv_bfloat16 h;
        if (part < 2) {
          const uint32_t hb = bits & 0xFFFF0000u;
          float hf;
          memcpy(&hf, &hb, 4);
          h = __float2bfloat16_rn(hf);   // exact: hf has 8 significant bits
          rem -= hf;
        } else {
          h = __float2bfloat16_rn(rem);
        }
        parts[(part * 32 + co) * 32 + k] = h;
      }
      const double c = w * mean[ci] / stdv[ci];
      for (int v = 0; v < 4; ++v) {
        const bool dropped = ((v & 1) && kx == 2) || ((v & 2) && ky == 2);
        if (!dropped) c_var[v] += c;
      }
    }
    for (int v = 0; v < 4; ++v) bias_v[v * 32 + co] = (float)((double)bias[co] - (double)scale[co] * c_var[v]);
  }
  MC_CUDA(cudaMalloc((void**)&p->d_w, parts.size() * sizeof(__nv_bfloat16)));
  MC_CUDA(cudaMalloc((void**)&p->d_bias_v, bias_v.size() * sizeof(float)));
  MC_CUDA(cudaMemcpy(p->d_w, parts.data(), parts.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  MC_CUDA(cudaMemcpy(p->d_bias_v, bias_v.data(), bias_v.size() * sizeof(float), cudaMemcpyHostToDevice));
  cudaDeviceProp prop;
  MC_CUDA(cudaGetDeviceProperties(&prop, device));
  p->num_sms = prop.multiProcessorCount;
  MC_CUDA(cudaFuncSetAttribute(stem_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM));
  MC_CUDA(cudaFuncSetAttribute(stem_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM));
  return MC_OK;
}

template <typename T>
inline int stem_tc_launch(const StemTcPlan& p, const mc_image* d_images, const mc_point* d_points, const float* d_scale, T* out,
                          int nb, cudaStream_t st) {
  StemTcArgs a;
  a.w_parts = p.d_w;
  a.scale = d_scale;
  a.bias_v = p.d_bias_v;
  a.out = out;
  a.nb = nb;
  const int items = nb * (112 / STEM_BAND);
  MC_CUDA(launch_pdl(PDL_STEM, stem_tc_kernel<T>, dim3(std::min(items, p.num_sms)), dim3(STEM_TC_THREADS), STEM_SMEM, st, d_images, d_points, a));
  MC_CHECK_LAUNCH();
  return MC_OK;
}

}  // namespace mc
