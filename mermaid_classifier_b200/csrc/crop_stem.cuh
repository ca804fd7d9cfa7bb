// K1 (point-patch gather / reflect crop / normalise) and K2 (stem conv) kernels.
//
// Reference semantics restated (see oracle/crop.py):
//   patch[i][j][c] = img[R(row-112+i, H)][R(col-112+j, W)][c]   (np.pad mode='reflect')
//   x[c][i][j]     = (u8/255 - mean[c]) / std[c]                (ToTensor + Normalize)
//   stem           = swish(bn0(conv3x3 s2, TF-SAME pad (0,1)))  (lukemelas EfficientNet)
// The normalisation is a 3x256-entry lookup table computed on the host in fp32 with the
// exact op order torch uses, so the device result is bit-identical by construction.
#pragma once
#include "common.cuh"

namespace mc {

// ---- synthetic image (bench/test input generator) ------------------------------------
__global__ void synth_image_kernel(uint8_t* img, int H, int W, int64_t pitch, uint32_t key) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const uint32_t pos = mix32(((uint32_t)y << 16) | (uint32_t)x);
  const uint32_t ramp = (((uint32_t)x & 63u) + ((uint32_t)y & 63u)) >> 2;
  uint8_t* p = img + (int64_t)y * pitch + (int64_t)x * 3;
#pragma unroll
  for (uint32_t c = 0; c < 3; ++c) {
    const uint32_t cell =
        mix32(key ^ (((uint32_t)y >> 6) * 0x85EBCA77u) ^ (((uint32_t)x >> 6) * 0xC2B2AE3Du) ^ (c * 0x27D4EB2Fu));
    const uint32_t base = 40u + ((((cell >> 8) & 0xFFFFu) * 160u) >> 16);
    const uint32_t noise = mix32((key + c * 0x632BE5ABu) ^ pos) & 31u;
    p[c] = (uint8_t)(base + ramp + noise - 16u);
  }
}

// ---- K1 stand-alone: bit-exact crop (parity gate + pre-cropped patch producer) ---------
// grid (8 row bands, n points), 192 threads: a CTA copies 28 patch rows of 672 bytes.
// Interior rows (no column reflection) are contiguous in the source image: thread t produces output word t
// from two aligned 4-byte loads and a funnel shift (vectorised, coalesced; the source alignment is arbitrary
// because a pixel is 3 bytes).  Rows that reflect in x fall back to per-pixel byte gathers through shared memory.
// Row reflection in y only selects the source row.
__global__ void __launch_bounds__(192) crop_kernel(const mc_image* __restrict__ images, const mc_point* __restrict__ points,
                                                   uint8_t* __restrict__ out) {
  constexpr int ROWS = 28;
  const int64_t k = blockIdx.y;
  const mc_point pt = points[k];
  const mc_image im = images[pt.image];
  const int x0 = pt.col - 112;
  // one pixel of slack on either side: the first / last aligned word of a misaligned segment reaches up to 3 bytes
  // before / past it, and must stay inside the row even when the image base or pitch is not 4-byte aligned
  const bool interior = x0 >= 1 && x0 + 225 <= im.width;
  const int tid = threadIdx.x;
  __shared__ uint8_t row_s[224 * 3];
#pragma unroll 7
  for (int r = 0; r < ROWS; ++r) {
    const int i = blockIdx.x * ROWS + r;
    const int y = reflect_idx(pt.row - 112 + i, im.height);
    const uint8_t* src_row = im.data + (int64_t)y * im.row_pitch;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (k * 224 + i) * 672);
    if (interior) {
      if (tid < 168) {
        const uint8_t* s = src_row + (int64_t)x0 * 3 + tid * 4;        // first source byte of output word `tid`
        const uintptr_t a = reinterpret_cast<uintptr_t>(s);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8u;
        const uint32_t w0 = __ldg(w);
        // the second word is only dereferenced when the segment is misaligned (never reads past the last needed byte's word)
        const uint32_t w1 = sh ? __ldg(w + 1) : 0u;
        dst[tid] = __funnelshift_r(w0, w1, sh);
      }
    } else {
      for (int j = tid; j < 224; j += 192) {
        const int x = reflect_idx(x0 + j, im.width);
        const uint8_t* s = src_row + (int64_t)x * 3;
        row_s[j * 3 + 0] = s[0];
        row_s[j * 3 + 1] = s[1];
        row_s[j * 3 + 2] = s[2];
      }
      __syncthreads();
      if (tid < 168) dst[tid] = reinterpret_cast<const uint32_t*>(row_s)[tid];
      __syncthreads();
    }
  }
}

// ---- A3 stand-alone: ToTensor + Normalize (parity gate only) ---------------------------
__global__ void normalize_kernel(const uint8_t* __restrict__ patches, const float* __restrict__ lut,
                                 float* __restrict__ out, int64_t n_pix_total) {
  // one thread per (patch, i, j)
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pix_total) return;
  const int64_t k = t / (224 * 224);
  const int64_t ij = t % (224 * 224);
  const uint8_t* p = patches + t * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[(k * 3 + c) * (224 * 224) + ij] = lut[c * 256 + p[c]];
}

// ---- fused K1+K2: gather + normalise + stem conv3x3/s2 + BN + swish --------------------
// One CTA = one 16x16 tile of the 112x112 stem output of one patch (grid 49 x n).
// The 33x33x3 input window is gathered straight from the source image (reflect), normalised
// through the LUT into shared memory, and each thread produces the 32 output channels of one
// output pixel; the tile is then written out NHWC, fully coalesced.
// The 27x32 weights and the folded BN live in the kernel's PARAMETER space (constant bank): every
// FFMA takes its weight as a constant operand, so the math loop issues no weight loads at all
// (the shared-memory version spent 216 LDS.128 per thread on them and was smem-bandwidth bound).
struct StemParams {
  float w[27 * 32];   // [ky][kx][ci][co]
  float scale[32];
  float bias[32];
};

// np.pad(mode='reflect') index with a fast path for the usual single reflection
__device__ __forceinline__ int reflect_fast(int t, int n) {
  if (t < 0) t = -t;
  if (t >= n) t = 2 * (n - 1) - t;
  return (t >= 0 && t < n) ? t : reflect_idx(t, n);
}

template <typename T>
__global__ void __launch_bounds__(256, 3) stem_kernel(const mc_image* __restrict__ images,
                                                   const mc_point* __restrict__ points,
                                                   const __grid_constant__ StemParams P,
                                                   const float* __restrict__ lut,    // [3][256]
                                                   T* __restrict__ out) {            // [n][112][112][32]
  // One CTA = one ROW of seven 16x16 output tiles of one patch (grid 7 x n).  The source bytes of tile t+1 are
  // gathered into registers while tile t is convolved, so the gather's DRAM latency hides behind the FMAs.
  constexpr int TS = 16, IN = 2 * TS + 1;  // 33
  constexpr int NT = 7;                    // tiles per row
  __shared__ float in_s[IN][IN * 3 + 1];
  __shared__ float lut_s[768];
  const int tid = threadIdx.x;
  const int oy0 = blockIdx.x * TS;
  const int64_t k = blockIdx.y;
  const mc_point pt = points[k];
  const mc_image im = images[pt.image];
  constexpr int NIT = (IN * IN + 255) / 256;   // 5 window pixels per thread
  uint8_t px[NIT][3];
  bool ok[NIT];
  auto gather = [&](const int ox0) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int t = tid + it * 256;
      const int r = t / IN, cpx = t - r * IN;
      const int pi = 2 * oy0 + r, pj = 2 * ox0 + cpx;   // patch coordinates; 224 == the SAME zero pad
      ok[it] = t < IN * IN && pi < 224 && pj < 224;
      px[it][0] = px[it][1] = px[it][2] = 0;
      if (ok[it]) {
        const uint8_t* s = im.data + (int64_t)reflect_fast(pt.row - 112 + pi, im.height) * im.row_pitch +
                           (int64_t)reflect_fast(pt.col - 112 + pj, im.width) * 3;
        px[it][0] = s[0];
        px[it][1] = s[1];
        px[it][2] = s[2];
      }
    }
  };
  gather(0);
  for (int t = tid; t < 768; t += 256) lut_s[t] = lut[t];
  __syncthreads();
  const int ty = tid / TS, tx = tid % TS;
  for (int tile = 0; tile < NT; ++tile) {
    const int ox0 = tile * TS;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int t = tid + it * 256;
      if (t < IN * IN) {
        const int r = t / IN, cpx = t - r * IN;
        in_s[r][cpx * 3 + 0] = ok[it] ? lut_s[px[it][0]] : 0.f;
        in_s[r][cpx * 3 + 1] = ok[it] ? lut_s[256 + px[it][1]] : 0.f;
        in_s[r][cpx * 3 + 2] = ok[it] ? lut_s[512 + px[it][2]] : 0.f;
      }
    }
    __syncthreads();
    if (tile + 1 < NT) gather(ox0 + TS);   // in flight during the convolution below
    // output channels in pairs: one packed FMA (FFMA2) per pair and tap, the weight pair a 64-bit constant operand
    float2 acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float v = in_s[2 * ty + ky][(2 * tx + kx) * 3 + ci];
          const float2 vv = make_float2(v, v);
          const float* wt = P.w + ((ky * 3 + kx) * 3 + ci) * 32;
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = __ffma2_rn(vv, make_float2(wt[2 * c], wt[2 * c + 1]), acc[c]);
        }
      }
    }
    T* o = out + (((k * 112 + (oy0 + ty)) * 112) + (ox0 + tx)) * 32;
    constexpr int VN = Vec<T>::N;
#pragma unroll
    for (int q = 0; q < 32 / VN; ++q) {
      Vec<T> v;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const int c = q * VN + e;
        v.v[e] = bn_silu<T>((c & 1) ? acc[c >> 1].y : acc[c >> 1].x, P.scale[c], P.bias[c]);
      }
      v.store(o + q * VN);
    }
    __syncthreads();   // everyone is done reading in_s before the next tile overwrites it
  }
}

// Crop a crop x crop window around each point (same reflect-padded window rule as crop_kernel, centre offset crop / 2) and
// resize it to 224 x 224 with bilinear interpolation -- the patch-size != 224 path (model.json config.patch_size,
// inference/export.py:77).  The arithmetic is torch.nn.functional.interpolate(mode="bilinear", align_corners=False,
// antialias=False) on the float patch, reproduced to the bit: source index = fma(scale, i + 0.5, -0.5) clamped at 0, weights
// l1 = index - floor(index), l0 = 1 - l1, value = fma(l0, a, l1 * b) along x and then along y (the contraction torch's CPU
// kernel compiles to; checked against torch in tests/test_oracle_crop.py); the result is rounded half-to-even to uint8.
// One CTA = one band of eight output rows of one patch, one thread = one output pixel (its four taps are byte gathers through
// the read-only path: neighbouring threads read neighbouring pixels, and this is the side path -- 224-pixel patches take the
// coalesced crop / the fused stem).
__global__ void __launch_bounds__(224) crop_resize_kernel(const mc_image* __restrict__ images, const mc_point* __restrict__ points,
                                                          int crop, uint8_t* __restrict__ out) {
  constexpr int ROWS = 8;
  const int64_t k = blockIdx.y;
  const mc_point pt = points[k];
  const mc_image im = images[pt.image];
  const int half = crop / 2;
  const int j = threadIdx.x;                       // output column
  const float scale = (float)crop / 224.f;
  // x taps of this thread
  const float rx = fmaxf(__fmaf_rn(scale, (float)j + 0.5f, -0.5f), 0.f);
  const int x0c = min((int)rx, crop - 1);
  const float lx1 = fminf(fmaxf(rx - (float)x0c, 0.f), 1.f), lx0 = 1.f - lx1;
  const int x1c = min(x0c + 1, crop - 1);
  const int xa = reflect_idx(pt.col - half + x0c, im.width) * 3, xb = reflect_idx(pt.col - half + x1c, im.width) * 3;
#pragma unroll 2
  for (int r = 0; r < ROWS; ++r) {
    const int i = blockIdx.x * ROWS + r;
    const float ry = fmaxf(__fmaf_rn(scale, (float)i + 0.5f, -0.5f), 0.f);
    const int y0c = min((int)ry, crop - 1);
    const float ly1 = fminf(fmaxf(ry - (float)y0c, 0.f), 1.f), ly0 = 1.f - ly1;
    const int y1c = min(y0c + 1, crop - 1);
    const uint8_t* row0 = im.data + (int64_t)reflect_idx(pt.row - half + y0c, im.height) * im.row_pitch;
    const uint8_t* row1 = im.data + (int64_t)reflect_idx(pt.row - half + y1c, im.height) * im.row_pitch;
    uint8_t* dst = out + ((k * 224 + i) * 224 + j) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = (float)__ldg(row0 + xa + c), b = (float)__ldg(row0 + xb + c);
      const float d = (float)__ldg(row1 + xa + c), e = (float)__ldg(row1 + xb + c);
      const float t0 = __fmaf_rn(lx0, a, __fmul_rn(lx1, b));
      const float t1 = __fmaf_rn(lx0, d, __fmul_rn(lx1, e));
      const float v = __fmaf_rn(ly0, t0, __fmul_rn(ly1, t1));
      dst[c] = (uint8_t)min(max(__float2int_rn(v), 0), 255);
    }
  }
}

}  // namespace mc
