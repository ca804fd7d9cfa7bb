// mc_jpeg_decode_exact: baseline JPEG -> RGB8 on the device, BIT FOR BIT what libjpeg-turbo (hence PIL, hence the
// reference's spacer.storage.load_image, call site mermaid_classifier/pyspacer/annotation.py:235) produces.
//
// nvJPEG (mc_jpeg_decode) uses its own inverse DCT and chroma interpolation: its bytes differ from PIL's by a grey level or
// two.  Here the arithmetic of libjpeg-turbo's default decode path is restated -- jidctint.c jpeg_idct_islow (accurate
// integer IDCT), jdsample.c h2v1 / h2v2 "fancy" (triangle) upsampling with its narrow-image and edge rules, jdcolor.c
// fixed-point YCbCr -> RGB -- and checked against PIL through oracle/jpeg.py (tests/test_oracle_jpeg.py pins the oracle to
// PIL byte for byte; tests/test_gpu_decode.py pins these kernels to both).
//
//   host thread   marker parsing + Huffman decoding (sequential by nature) into a SPARSE coefficient stream: one 32-bit
//                 word per non-zero coefficient (natural-order position << 16 | value) and one offset per 8x8 block, in
//                 pinned memory -- for a 12 MP 4:2:0 photograph ~10 MB instead of the 36 MB of decoded pixels
//   idct kernel   one thread per block: scatter into shared memory, dequantise, two LL&M butterfly passes, range limit
//   colour kernel one thread per pixel: chroma upsampled on the fly from the component planes, YCbCr -> RGB, 3 bytes out
//
// Covered: baseline / extended-sequential and progressive (jdphuff.c: DC / AC first and refinement scans, accumulated on the
// host into dense coefficients, then sparsified) Huffman streams, 8-bit, grayscale or YCbCr with chroma 1x1 and luma 1x1 / 2x1
// / 2x2 (4:4:4, 4:2:2, 4:2:0), restart intervals.  Anything else returns MC_ERR_UNSUPPORTED (callers fall back to
// mc_jpeg_decode or PIL).
namespace {

constexpr int JPX_ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,
                                7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JpxComp {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int bx = 0, by = 0;          // block grid of this component (whole MCUs)
};

struct JpxHuff {
  // 9-bit lookahead: (length << 8) | symbol, 0 = longer code; canonical tables for the slow path
  uint16_t look[512];
  int maxcode[18];
  int valptr[17];
  int mincode[17];
  uint8_t vals[256];
  // AC tables only: when the code AND the magnitude bits of a coefficient fit in the 9-bit lookahead, the decoded value, its
  // zero run and the total bit count in one entry -- (value << 8) | (run << 4) | bits, 0 = take the general path
  int16_t fast_ac[512];
  bool present = false;
};

struct JpxHeader {
  int width = 0, height = 0, ncomp = 0, ri = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  bool progressive = false;
  JpxComp comp[3];
  uint16_t qt[4][64];          // natural order
  bool qt_present[4] = {false, false, false, false};
  JpxHuff dc[4], ac[4];
  size_t data_pos = 0;
};

void jpx_build_huff(JpxHuff* h, const uint8_t* counts, const uint8_t* symbols, int nsym) {
  memset(h->look, 0, sizeof(h->look));
  memcpy(h->vals, symbols, (size_t)nsym);
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    h->valptr[len] = k;
    h->mincode[len] = code;
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      if (len <= 9) {
        const int shift = 9 - len;
        for (int fill = 0; fill < (1 << shift); ++fill) h->look[(code << shift) | fill] = (uint16_t)((len << 8) | symbols[k]);
      }
    }
    h->maxcode[len] = counts[len - 1] ? code - 1 : -1;
    code <<= 1;
  }
  h->maxcode[17] = 0x7fffffff;
  for (int i = 0; i < 512; ++i) {
    h->fast_ac[i] = 0;
    const int e = h->look[i];
    if (!e) continue;
    const int len = e >> 8, rs = e & 0xFF, run = rs >> 4, mag = rs & 15;
    if (mag && len + mag <= 9) {
      int k = ((i << len) & 511) >> (9 - mag);          // the magnitude bits that follow the code
      if (k < (1 << (mag - 1))) k -= (1 << mag) - 1;      // EXTEND
      if (k >= -128 && k <= 127) h->fast_ac[i] = (int16_t)(k * 256 + run * 16 + (len + mag));
    }
  }
  h->present = true;
}

// returns MC_OK, or a failure code with the message set
int jpx_parse(const uint8_t* d, size_t n, JpxHeader* H) {
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail(MC_ERR_BAD_ARG, "not a JPEG stream");
  size_t pos = 2;
  bool have_frame = false;
  int adobe_transform = -1;
  while (pos + 4 <= n) {
    if (d[pos] != 0xFF) return fail(MC_ERR_BAD_ARG, "JPEG: marker expected");
    while (pos < n && d[pos] == 0xFF) ++pos;
    if (pos >= n) break;
    const int m = d[pos++];
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (pos + 2 > n) break;
    const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
    if (len < 2 || pos + len > n) return fail(MC_ERR_BAD_ARG, "JPEG: truncated segment");
    const uint8_t* s = d + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {
      size_t i = 0;
      while (i < sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > sl) return fail(MC_ERR_BAD_ARG, "JPEG: bad DQT");
        for (int k = 0; k < 64; ++k) {
          const int v = pq ? ((s[i + 2 * k] << 8) | s[i + 2 * k + 1]) : s[i + k];
          H->qt[tq][JPX_ZIGZAG[k]] = (uint16_t)v;
        }
        H->qt_present[tq] = true;
        i += pq ? 128 : 64;
      }
    } else if (m == 0xC4) {
      size_t i = 0;
      while (i + 17 <= sl) {
        const int tc = s[i] >> 4, th = s[i] & 15;
        int nsym = 0;
        for (int k = 0; k < 16; ++k) nsym += s[i + 1 + k];
        if (tc > 1 || th > 3 || nsym > 256 || i + 17 + (size_t)nsym > sl) return fail(MC_ERR_BAD_ARG, "JPEG: bad DHT");
        jpx_build_huff(tc ? &H->ac[th] : &H->dc[th], s + i + 1, s + i + 17, nsym);
        i += 17 + (size_t)nsym;
      }
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
      H->progressive = m == 0xC2;
      if (sl < 6 || s[0] != 8) return fail(MC_ERR_UNSUPPORTED, "JPEG: sample precision other than 8 bits");
      H->height = (s[1] << 8) | s[2];
      H->width = (s[3] << 8) | s[4];
      H->ncomp = s[5];
      if (H->ncomp != 1 && H->ncomp != 3) return fail(MC_ERR_UNSUPPORTED, "JPEG: only grayscale and YCbCr streams are decoded exactly");
      if (sl < 6 + 3 * (size_t)H->ncomp) return fail(MC_ERR_BAD_ARG, "JPEG: bad SOF");
      for (int c = 0; c < H->ncomp; ++c) {
        H->comp[c].id = s[6 + 3 * c];
        H->comp[c].h = s[7 + 3 * c] >> 4;
        H->comp[c].v = s[7 + 3 * c] & 15;
        H->comp[c].tq = s[8 + 3 * c];
      }
      have_frame = true;
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return fail(MC_ERR_UNSUPPORTED, "JPEG: not a Huffman-coded baseline or progressive stream");
    } else if (m == 0xDD) {
      if (sl >= 2) H->ri = (s[0] << 8) | s[1];
    } else if (m == 0xEE && sl >= 12 && memcmp(s, "Adobe", 5) == 0) {
      adobe_transform = s[11];
    } else if (m == 0xDA) {
      if (have_frame && H->progressive) {   // the scans are walked by jpx_decode_progressive
        H->data_pos = pos + len;
        break;
      }
      if (!have_frame || sl < 1 || s[0] != H->ncomp || sl < 1 + 2 * (size_t)H->ncomp)
        return fail(MC_ERR_UNSUPPORTED, "JPEG: multi-scan sequential streams are not decoded exactly");
      for (int c = 0; c < H->ncomp; ++c) {
        int ci = -1;
        for (int k = 0; k < H->ncomp; ++k)
          if (H->comp[k].id == s[1 + 2 * c]) ci = k;
        if (ci != c) return fail(MC_ERR_UNSUPPORTED, "JPEG: scan component order");
        H->comp[c].td = s[2 + 2 * c] >> 4;
        H->comp[c].ta = s[2 + 2 * c] & 15;
      }
      H->data_pos = pos + len;
      break;
    } else if (m == 0xD9) {
      return fail(MC_ERR_BAD_ARG, "JPEG: no scan");
    }
    pos += len;
  }
  if (!H->data_pos || H->width < 1 || H->height < 1) return fail(MC_ERR_BAD_ARG, "JPEG: no scan data");
  if (H->ncomp == 3) {
    if (adobe_transform == 0) return fail(MC_ERR_UNSUPPORTED, "JPEG: RGB (Adobe transform 0) streams are not decoded exactly");
    if (H->comp[1].h != 1 || H->comp[1].v != 1 || H->comp[2].h != 1 || H->comp[2].v != 1)
      return fail(MC_ERR_UNSUPPORTED, "JPEG: chroma sampling factors other than 1x1");
    const int lh = H->comp[0].h, lv = H->comp[0].v;
    if (!((lh == 1 && lv == 1) || (lh == 2 && lv == 1) || (lh == 2 && lv == 2)))
      return fail(MC_ERR_UNSUPPORTED, "JPEG: luma sampling factors other than 1x1, 2x1, 2x2");
    H->hmax = lh;
    H->vmax = lv;
    H->mcux = (H->width + 8 * H->hmax - 1) / (8 * H->hmax);
    H->mcuy = (H->height + 8 * H->vmax - 1) / (8 * H->vmax);
    for (int c = 0; c < 3; ++c) {
      H->comp[c].bx = H->mcux * H->comp[c].h;
      H->comp[c].by = H->mcuy * H->comp[c].v;
    }
  } else {
    H->hmax = H->vmax = 1;
    H->comp[0].h = H->comp[0].v = 1;   // a single-component scan is never interleaved
    H->mcux = (H->width + 7) / 8;
    H->mcuy = (H->height + 7) / 8;
    H->comp[0].bx = H->mcux;
    H->comp[0].by = H->mcuy;
  }
  for (int c = 0; c < H->ncomp; ++c) {
    if (H->comp[c].tq > 3 || !H->qt_present[H->comp[c].tq]) return fail(MC_ERR_BAD_ARG, "JPEG: missing quantisation table");
    if (!H->progressive &&
        (H->comp[c].td > 3 || H->comp[c].ta > 3 || !H->dc[H->comp[c].td].present || !H->ac[H->comp[c].ta].present))
      return fail(MC_ERR_BAD_ARG, "JPEG: missing Huffman table");
  }
  return MC_OK;
}

struct JpxBits {
  const uint8_t* d;
  size_t pos, n;
  uint64_t acc = 0;
  int cnt = 0;
  bool hit_marker = false;
  inline void fill() {
    while (cnt <= 56) {
      uint8_t b = 0;
      if (!hit_marker && pos < n) {
        b = d[pos];
        if (b == 0xFF) {
          const uint8_t nx = pos + 1 < n ? d[pos + 1] : 0xD9;
          if (nx == 0) {
            pos += 2;
          } else {   // a marker: feed zeros until the caller handles it (libjpeg does the same)
            hit_marker = true;
            b = 0;
          }
        } else {
          ++pos;
        }
      }
      acc |= (uint64_t)b << (56 - cnt);
      cnt += 8;
    }
  }
  inline int peek(int k) { return (int)(acc >> (64 - k)); }
  inline void skip(int k) {
    acc <<= k;
    cnt -= k;
  }
  inline int get(int k) {
    if (k == 0) return 0;
    if (cnt < k) fill();
    const int v = peek(k);
    skip(k);
    return v;
  }
  // byte-align and step over the next RSTn marker
  void restart() {
    acc = 0;
    cnt = 0;
    hit_marker = false;
    while (pos + 1 < n && !(d[pos] == 0xFF && d[pos + 1] >= 0xD0 && d[pos + 1] <= 0xD7)) ++pos;
    if (pos + 1 < n) pos += 2;
  }
};

inline int jpx_decode_symbol(JpxBits& br, const JpxHuff& h) {
  if (br.cnt < 16) br.fill();
  const int e = h.look[br.peek(9)];
  if (e) {
    br.skip(e >> 8);
    return e & 0xFF;
  }
  int code = br.peek(10);
  for (int len = 10; len <= 16; ++len) {
    if (code <= h.maxcode[len]) {
      br.skip(len);
      return h.vals[h.valptr[len] + code - h.mincode[len]];
    }
    code = br.peek(len + 1);
  }
  br.skip(16);
  return 0;   // corrupt stream: libjpeg warns and returns 0
}

inline int jpx_extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------------------
namespace mc {

struct JpxKernelArgs {
  int ncomp;
  int blk_base[4];              // first block (in the offset table) of each component; [ncomp] = total
  int bx[3], by[3];             // block grids
  int64_t plane_off[3];         // byte offset of each component plane inside `planes`
  uint16_t qt[3][64];           // natural order
};

__device__ __forceinline__ int jpx_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jidctint.c: one butterfly pass on eight values (CONST_BITS 13)
__device__ __forceinline__ void jpx_idct_1d(int (&v)[8], const int shift) {
  constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
                F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069,
                F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;
  int z2 = v[2], z3 = v[6];
  int z1 = (z2 + z3) * F_0_541196100;
  int tmp2 = z1 + z3 * (-F_1_847759065);
  int tmp3 = z1 + z2 * F_0_765366865;
  int tmp0 = (v[0] + v[4]) << 13;
  int tmp1 = (v[0] - v[4]) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = v[7];
  tmp1 = v[5];
  tmp2 = v[3];
  tmp3 = v[1];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * F_1_175875602;
  tmp0 *= F_0_298631336;
  tmp1 *= F_2_053119869;
  tmp2 *= F_3_072711026;
  tmp3 *= F_1_501321110;
  z1 *= -F_0_899976223;
  z2 *= -F_2_562915447;
  z3 = z3 * (-F_1_961570560) + z5;
  z4 = z4 * (-F_0_390180644) + z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  v[0] = jpx_descale(tmp10 + tmp3, shift);
  v[7] = jpx_descale(tmp10 - tmp3, shift);
  v[1] = jpx_descale(tmp11 + tmp2, shift);
  v[6] = jpx_descale(tmp11 - tmp2, shift);
  v[2] = jpx_descale(tmp12 + tmp1, shift);
  v[5] = jpx_descale(tmp12 - tmp1, shift);
  v[3] = jpx_descale(tmp13 + tmp0, shift);
  v[4] = jpx_descale(tmp13 - tmp0, shift);
}

// range_limit[x & RANGE_MASK] of jdmaster.c's table: x + 128 clamped to 0..255 for x in [-512, 511], wrap-around outside
__device__ __forceinline__ uint32_t jpx_range_limit(int x) {
  const int i = x & 1023;
  return (uint32_t)(i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896)));
}

constexpr int JPX_IDCT_THREADS = 128;

// one thread per 8x8 block: sparse entries -> shared memory -> dequantise -> IDCT -> 8 rows of 8 bytes
__global__ void __launch_bounds__(JPX_IDCT_THREADS)
jpx_idct_kernel(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets, uint8_t* __restrict__ planes,
                const JpxKernelArgs a) {
  __shared__ short coef[64][JPX_IDCT_THREADS];   // [position][thread]: conflict-free in both phases
  __shared__ uint16_t qt_s[3][64];
  for (int i = threadIdx.x; i < 3 * 64; i += JPX_IDCT_THREADS) qt_s[i / 64][i % 64] = a.qt[i / 64][i % 64];
  const int b = blockIdx.x * JPX_IDCT_THREADS + threadIdx.x;
  const int total = a.blk_base[a.ncomp];
#pragma unroll
  for (int k = 0; k < 64; ++k) coef[k][threadIdx.x] = 0;
  __syncthreads();
  if (b >= total) return;
  int c = 0;
  while (c + 1 < a.ncomp && b >= a.blk_base[c + 1]) ++c;
  const uint32_t e0 = offsets[2 * b], e1 = offsets[2 * b + 1];   // [begin, end): an interleaved scan does not visit blocks in index order
  for (uint32_t e = e0; e < e1; ++e) {
    const uint32_t w = entries[e];
    coef[w >> 16][threadIdx.x] = (short)(w & 0xFFFFu);
  }
  // pass 1: columns
  int ws[8][8];
#pragma unroll
  for (int col = 0; col < 8; ++col) {
    int v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = (int)coef[r * 8 + col][threadIdx.x] * (int)qt_s[c][r * 8 + col];
    jpx_idct_1d(v, 13 - 2);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r][col] = v[r];
  }
  // pass 2: rows
  const int lb = b - a.blk_base[c];
  const int byi = lb / a.bx[c], bxi = lb - byi * a.bx[c];
  const int pitch = a.bx[c] * 8;
  uint8_t* dst = planes + a.plane_off[c] + (int64_t)(byi * 8) * pitch + bxi * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ws[r][k];
    jpx_idct_1d(v, 13 + 2 + 3);
    const uint32_t lo = jpx_range_limit(v[0]) | (jpx_range_limit(v[1]) << 8) | (jpx_range_limit(v[2]) << 16) | (jpx_range_limit(v[3]) << 24);
    const uint32_t hi = jpx_range_limit(v[4]) | (jpx_range_limit(v[5]) << 8) | (jpx_range_limit(v[6]) << 16) | (jpx_range_limit(v[7]) << 24);
    *reinterpret_cast<uint2*>(dst + (int64_t)r * pitch) = make_uint2(lo, hi);
  }
}

// jdsample.c fancy upsampling of one chroma sample at full-resolution position (y, x), evaluated on the fly
__device__ __forceinline__ int jpx_chroma(const uint8_t* __restrict__ p, int pitch, int cw, int ch, int hs, int vs, int y, int x) {
  if (hs == 1) return p[(int64_t)y * pitch + x];
  const int cx = x >> 1;
  if (cw <= 2) return p[(int64_t)(vs == 2 ? y >> 1 : y) * pitch + cx];   // narrow components are replicated, not interpolated
  if (vs == 1) {   // h2v1
    const uint8_t* row = p + (int64_t)y * pitch;
    const int v = row[cx];
    if (!(x & 1)) return cx == 0 ? v : (3 * v + row[cx - 1] + 1) >> 2;
    return cx == cw - 1 ? v : (3 * v + row[cx + 1] + 2) >> 2;
  }
  // h2v2: the nearer row weighs 3, the farther 1 (rows above the first / below the last are the edge rows themselves)
  const int cy = y >> 1;
  const int oy = (y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);
  const uint8_t* r0 = p + (int64_t)cy * pitch;
  const uint8_t* r1 = p + (int64_t)oy * pitch;
  const int cs = 3 * r0[cx] + r1[cx];
  if (!(x & 1)) return cx == 0 ? (cs * 4 + 8) >> 4 : (cs * 3 + (3 * r0[cx - 1] + r1[cx - 1]) + 8) >> 4;
  return cx == cw - 1 ? (cs * 4 + 7) >> 4 : (cs * 3 + (3 * r0[cx + 1] + r1[cx + 1]) + 7) >> 4;
}

// one thread per pixel: jdcolor.c ycc_rgb_convert with its fixed-point tables evaluated directly
__global__ void __launch_bounds__(256)
jpx_color_kernel(const uint8_t* __restrict__ planes, const JpxKernelArgs a, int width, int height, int hs, int vs,
                 uint8_t* __restrict__ rgb, int64_t row_pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= width) return;
  const int yv = planes[a.plane_off[0] + (int64_t)y * (a.bx[0] * 8) + x];
  uint8_t* o = rgb + (int64_t)y * row_pitch + (int64_t)x * 3;
  if (a.ncomp == 1) {
    o[0] = o[1] = o[2] = (uint8_t)yv;
    return;
  }
  const int cw = (width + hs - 1) / hs, ch = (height + vs - 1) / vs;
  const int cb = jpx_chroma(planes + a.plane_off[1], a.bx[1] * 8, cw, ch, hs, vs, y, x) - 128;
  const int cr = jpx_chroma(planes + a.plane_off[2], a.bx[2] * 8, cw, ch, hs, vs, y, x) - 128;
  // FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768, SCALEBITS 16
  const int r = yv + ((91881 * cr + 32768) >> 16);
  const int g = yv + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  const int bl = yv + ((116130 * cb + 32768) >> 16);
  o[0] = (uint8_t)min(max(r, 0), 255);
  o[1] = (uint8_t)min(max(g, 0), 255);
  o[2] = (uint8_t)min(max(bl, 0), 255);
}

}  // namespace mc

// scratch of the exact path, owned by the mc_jpeg handle
struct JpxScratch {
  std::vector<int16_t> dense;      // progressive streams: coefficients accumulate over the scans before they are sparsified
  uint32_t* h_entries = nullptr;   // pinned
  uint32_t* h_offsets = nullptr;   // pinned
  uint32_t* d_entries = nullptr;
  uint32_t* d_offsets = nullptr;
  uint8_t* d_planes = nullptr;
  size_t cap_entries = 0, cap_offsets = 0, cap_planes = 0;
  cudaEvent_t h2d_done = nullptr;
  bool pending = false;
};

namespace {

void jpx_free(JpxScratch* s) {
  if (s->h_entries) cudaFreeHost(s->h_entries);
  if (s->h_offsets) cudaFreeHost(s->h_offsets);
  if (s->d_entries) cudaFree(s->d_entries);
  if (s->d_offsets) cudaFree(s->d_offsets);
  if (s->d_planes) cudaFree(s->d_planes);
  if (s->h2d_done) cudaEventDestroy(s->h2d_done);
  s->h_entries = s->h_offsets = s->d_entries = s->d_offsets = nullptr;
  s->d_planes = nullptr;
  s->h2d_done = nullptr;
  s->cap_entries = s->cap_offsets = s->cap_planes = 0;
  s->pending = false;
  s->dense.clear();
  s->dense.shrink_to_fit();
}

int jpx_reserve(JpxScratch* s, size_t entries, size_t offsets, size_t planes) {
  if (!s->h2d_done) MC_CUDA(cudaEventCreateWithFlags(&s->h2d_done, cudaEventDisableTiming));
  if (entries > s->cap_entries) {
    const size_t cap = entries + entries / 4 + 4096;
    if (s->h_entries) cudaFreeHost(s->h_entries);
    if (s->d_entries) cudaFree(s->d_entries);
    s->h_entries = nullptr;
    s->d_entries = nullptr;
    s->cap_entries = 0;
    MC_CUDA(cudaMallocHost((void**)&s->h_entries, cap * sizeof(uint32_t)));
    MC_CUDA(cudaMalloc((void**)&s->d_entries, cap * sizeof(uint32_t)));
    s->cap_entries = cap;
  }
  if (offsets > s->cap_offsets) {
    if (s->h_offsets) cudaFreeHost(s->h_offsets);
    if (s->d_offsets) cudaFree(s->d_offsets);
    s->h_offsets = nullptr;
    s->d_offsets = nullptr;
    s->cap_offsets = 0;
    MC_CUDA(cudaMallocHost((void**)&s->h_offsets, offsets * sizeof(uint32_t)));
    MC_CUDA(cudaMalloc((void**)&s->d_offsets, offsets * sizeof(uint32_t)));
    s->cap_offsets = offsets;
  }
  if (planes > s->cap_planes) {
    if (s->d_planes) cudaFree(s->d_planes);
    s->d_planes = nullptr;
    s->cap_planes = 0;
    MC_CUDA(cudaMalloc((void**)&s->d_planes, planes));
    s->cap_planes = planes;
  }
  return MC_OK;
}

// Huffman-decode the whole scan into the sparse stream (at most min(64 per block, ~4 per compressed byte) words).
int jpx_entropy_decode(const uint8_t* d, size_t n, const JpxHeader& H, const int* blk_base, uint32_t* entries, size_t cap_entries,
                       uint32_t* offsets, size_t* n_entries) {
  JpxBits br{d, H.data_pos, n};
  int pred[3] = {0, 0, 0};
  // block (component c, row y, column x) has index blk_base[c] + y * bx + x; the scan visits them MCU by MCU, so the
  // entries of a block are written at the running position and its offset recorded -- offsets[] is then NOT monotonic in
  // block index for interleaved scans; the kernel needs [begin, end) per block, so both are stored: offsets[2 b], [2 b + 1]
  size_t pos = 0;
  long count = 0;
  for (int my = 0; my < H.mcuy; ++my) {
    for (int mx = 0; mx < H.mcux; ++mx) {
      if (H.ri && count && count % H.ri == 0) {
        br.restart();
        pred[0] = pred[1] = pred[2] = 0;
      }
      ++count;
      for (int c = 0; c < H.ncomp; ++c) {
        const JpxComp& cp = H.comp[c];
        const JpxHuff& hd = H.dc[cp.td];
        const JpxHuff& ha = H.ac[cp.ta];
        for (int by = 0; by < cp.v; ++by) {
          for (int bx = 0; bx < cp.h; ++bx) {
            const int b = blk_base[c] + (my * cp.v + by) * cp.bx + (mx * cp.h + bx);
            if (pos + 64 > cap_entries) return fail(MC_ERR_BAD_ARG, "JPEG: more coefficients than the stream can hold (corrupt data)");
            offsets[2 * b] = (uint32_t)pos;
            int s = jpx_decode_symbol(br, hd);
            if (s > 15) s = 0;
            const int diff = s ? jpx_extend(br.get(s), s) : 0;
            pred[c] += diff;
            if (pred[c]) entries[pos++] = (0u << 16) | ((uint32_t)pred[c] & 0xFFFFu);
            for (int k = 1; k < 64;) {
              if (br.cnt < 16) br.fill();
              const int fa = ha.fast_ac[br.peek(9)];
              if (fa) {   // code + magnitude inside the lookahead: one table hit per coefficient
                k += (fa >> 4) & 15;
                if (k > 63) break;   // corrupt stream
                br.skip(fa & 15);
                entries[pos++] = ((uint32_t)JPX_ZIGZAG[k] << 16) | ((uint32_t)(fa >> 8) & 0xFFFFu);
                ++k;
                continue;
              }
              const int rs = jpx_decode_symbol(br, ha);
              const int r = rs >> 4, sz = rs & 15;
              if (sz == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) break;   // corrupt stream
              const int v = jpx_extend(br.get(sz), sz);
              entries[pos++] = ((uint32_t)JPX_ZIGZAG[k] << 16) | ((uint32_t)v & 0xFFFFu);
              ++k;
            }
            offsets[2 * b + 1] = (uint32_t)pos;
          }
        }
      }
    }
  }
  *n_entries = pos;
  return MC_OK;
}

// jdphuff.c: every scan of a progressive stream (DC / AC, first / refinement) accumulated into dense coefficients
// coef[block][64] (natural order, zero-initialised by the caller; block = blk_base[c] + y * bx + x).  Huffman tables and the
// restart interval may change between scans, so the markers are walked again from the top.
int jpx_decode_progressive(const uint8_t* d, size_t n, JpxHeader& H, const int* blk_base, int16_t* coef) {
  size_t pos = 2;
  int ri = 0;
  int scans = 0;
  while (pos + 4 <= n) {
    if (d[pos] != 0xFF) return fail(MC_ERR_BAD_ARG, "JPEG: marker expected");
    while (pos < n && d[pos] == 0xFF) ++pos;
    if (pos >= n) break;
    const int m = d[pos++];
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (m == 0xD9) break;
    if (pos + 2 > n) break;
    const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
    if (len < 2 || pos + len > n) return fail(MC_ERR_BAD_ARG, "JPEG: truncated segment");
    const uint8_t* s = d + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xC4) {
      size_t i = 0;
      while (i + 17 <= sl) {
        const int tc = s[i] >> 4, th = s[i] & 15;
        int nsym = 0;
        for (int k = 0; k < 16; ++k) nsym += s[i + 1 + k];
        if (tc > 1 || th > 3 || nsym > 256 || i + 17 + (size_t)nsym > sl) return fail(MC_ERR_BAD_ARG, "JPEG: bad DHT");
        jpx_build_huff(tc ? &H.ac[th] : &H.dc[th], s + i + 1, s + i + 17, nsym);
        i += 17 + (size_t)nsym;
      }
    } else if (m == 0xDD) {
      if (sl >= 2) ri = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {
      if (sl < 1) return fail(MC_ERR_BAD_ARG, "JPEG: bad SOS");
      const int ns = s[0];
      if (ns < 1 || ns > H.ncomp || sl < 4 + 2 * (size_t)ns) return fail(MC_ERR_BAD_ARG, "JPEG: bad SOS");
      int ci[3], td[3], ta[3];
      for (int c = 0; c < ns; ++c) {
        ci[c] = -1;
        for (int k = 0; k < H.ncomp; ++k)
          if (H.comp[k].id == s[1 + 2 * c]) ci[c] = k;
        if (ci[c] < 0) return fail(MC_ERR_BAD_ARG, "JPEG: scan names an unknown component");
        td[c] = s[2 + 2 * c] >> 4;
        ta[c] = s[2 + 2 * c] & 15;
        if (td[c] > 3 || ta[c] > 3) return fail(MC_ERR_BAD_ARG, "JPEG: bad table index");
      }
      const int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
      if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13)
        return fail(MC_ERR_BAD_ARG, "JPEG: bad progressive scan parameters");
      for (int c = 0; c < ns; ++c) {
        if (ss == 0 && ah == 0 && !H.dc[td[c]].present) return fail(MC_ERR_BAD_ARG, "JPEG: missing Huffman table");
        if (ss > 0 && !H.ac[ta[c]].present) return fail(MC_ERR_BAD_ARG, "JPEG: missing Huffman table");
      }
      ++scans;
      JpxBits br{d, pos + len, n};
      // blocks of this scan: MCU-interleaved over the padded grids, or the component's own (unpadded) block grid
      int gw, gh;
      if (ns > 1) {
        gw = H.mcux;
        gh = H.mcuy;
      } else {
        const JpxComp& cp = H.comp[ci[0]];
        const int cw = H.ncomp == 1 ? H.width : (H.width * cp.h + H.hmax - 1) / H.hmax;
        const int chh = H.ncomp == 1 ? H.height : (H.height * cp.v + H.vmax - 1) / H.vmax;
        gw = (cw + 7) / 8;
        gh = (chh + 7) / 8;
      }
      int pred[3] = {0, 0, 0};
      int eobrun = 0;
      const int p1 = 1 << al, m1 = -(1 << al);
      long count = 0;
      for (int gy = 0; gy < gh; ++gy) {
        for (int gx = 0; gx < gw; ++gx) {
          if (ri && count && count % ri == 0) {
            br.restart();
            pred[0] = pred[1] = pred[2] = 0;
            eobrun = 0;
          }
          ++count;
          for (int c = 0; c < ns; ++c) {
            const JpxComp& cp = H.comp[ci[c]];
            const int vb = ns > 1 ? cp.v : 1, hb = ns > 1 ? cp.h : 1;
            for (int by = 0; by < vb; ++by) {
              for (int bx = 0; bx < hb; ++bx) {
                int16_t* blk = coef + (size_t)(blk_base[ci[c]] + (gy * vb + by) * cp.bx + (gx * hb + bx)) * 64;
                if (ss == 0) {
                  if (ah == 0) {
                    int sz = jpx_decode_symbol(br, H.dc[td[c]]);
                    if (sz > 15) sz = 0;
                    pred[c] += sz ? jpx_extend(br.get(sz), sz) : 0;
                    blk[0] = (int16_t)(pred[c] * (1 << al));
                  } else if (br.get(1)) {
                    blk[0] |= (int16_t)p1;
                  }
                  continue;
                }
                const JpxHuff& ha = H.ac[ta[c]];
                if (ah == 0) {   // AC first
                  if (eobrun > 0) {
                    --eobrun;
                    continue;
                  }
                  for (int k = ss; k <= se; ++k) {
                    const int rs = jpx_decode_symbol(br, ha);
                    const int r = rs >> 4, sz = rs & 15;
                    if (sz) {
                      k += r;
                      if (k > 63) break;
                      blk[JPX_ZIGZAG[k]] = (int16_t)(jpx_extend(br.get(sz), sz) * (1 << al));
                    } else if (r == 15) {
                      k += 15;
                    } else {
                      eobrun = (1 << r) + (r ? br.get(r) : 0) - 1;
                      break;
                    }
                  }
                  continue;
                }
                // AC refinement (decode_mcu_AC_refine)
                int k = ss;
                if (eobrun == 0) {
                  for (; k <= se; ++k) {
                    const int rs = jpx_decode_symbol(br, ha);
                    int r = rs >> 4, sz = rs & 15;
                    if (sz) {
                      sz = br.get(1) ? p1 : m1;   // the size must be 1; any other value is treated the same way (libjpeg warns)
                    } else if (r != 15) {
                      eobrun = (1 << r) + (r ? br.get(r) : 0);
                      break;
                    }
                    do {
                      int16_t* cf = blk + JPX_ZIGZAG[k];
                      if (*cf != 0) {
                        if (br.get(1) && (*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
                      } else if (--r < 0) {
                        break;
                      }
                      ++k;
                    } while (k <= se);
                    if (sz && k <= 63) blk[JPX_ZIGZAG[k]] = (int16_t)sz;
                  }
                }
                if (eobrun > 0) {
                  for (; k <= se; ++k) {
                    int16_t* cf = blk + JPX_ZIGZAG[k];
                    if (*cf != 0 && br.get(1) && (*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
                  }
                  --eobrun;
                }
              }
            }
          }
        }
      }
      // on to the next marker after the entropy-coded segment
      size_t q = pos + len;
      while (q + 1 < n && !(d[q] == 0xFF && d[q + 1] != 0 && !(d[q + 1] >= 0xD0 && d[q + 1] <= 0xD7))) ++q;
      pos = q;
      continue;
    }
    pos += len;
  }
  if (!scans) return fail(MC_ERR_BAD_ARG, "JPEG: no scan");
  return MC_OK;
}

// dense coefficients -> the sparse stream of the kernels
int jpx_sparsify(const int16_t* coef, int total, uint32_t* entries, size_t cap_entries, uint32_t* offsets, size_t* n_entries) {
  size_t pos = 0;
  for (int b = 0; b < total; ++b) {
    if (pos + 64 > cap_entries) return fail(MC_ERR_BAD_ARG, "JPEG: more coefficients than the stream can hold (corrupt data)");
    offsets[2 * b] = (uint32_t)pos;
    const int16_t* blk = coef + (size_t)b * 64;
    for (int k = 0; k < 64; ++k)
      if (blk[k]) entries[pos++] = ((uint32_t)k << 16) | ((uint32_t)blk[k] & 0xFFFFu);
    offsets[2 * b + 1] = (uint32_t)pos;
  }
  *n_entries = pos;
  return MC_OK;
}

}  // namespace
