// EfficientNet-B0 layer table and the canonical packed-parameter layout.
//
// The layout MUST match mermaid_classifier_b200/weights.py::pack_backbone:
//   stem:   W[ky][kx][ci][co] (27 x 32), scale[32], bias[32]
//   block b (16 of them):
//     expand (absent when expand_ratio == 1): W[c_mid][c_in], scale[c_mid], bias[c_mid]
//     depthwise: W[k*k][c_mid], scale[c_mid], bias[c_mid]
//     se_reduce: W[c_se][c_mid], b[c_se];  se_expand: W^T[c_se][c_mid] (transposed), b[c_mid]
//     project:   W[c_out][c_mid], scale[c_out], bias[c_out]
//   head:   W[1280][320], scale[1280], bias[1280]
// Every segment is zero-padded to a multiple of 4 floats so each starts 16-byte aligned.
// scale/bias are the inference-form BatchNorm folded to y = conv * scale + bias
// (eps = 1e-3; pyspacer's vendored lukemelas EfficientNet).
#pragma once
#include <stdint.h>

#include <vector>

namespace mc {

struct BlockCfg {
  int k, stride, expand, c_in, c_out, c_mid, c_se;
  int h_in, h_out;  // square feature maps
  int pad;          // TF-"SAME" leading pad of the depthwise conv (top == left)
  bool skip;
  // parameter offsets (floats) into the packed blob
  int64_t w_exp, s_exp, b_exp, w_dw, s_dw, b_dw, w_se1, b_se1, w_se2, b_se2, w_proj, s_proj, b_proj;
};

struct NetCfg {
  std::vector<BlockCfg> blocks;
  int64_t w_stem, s_stem, b_stem, w_head, s_head, b_head;
  int64_t n_params;
  int64_t max_in_out;  // per-patch elements of the largest block input/output map
  int64_t max_mid;     // per-patch elements of the largest expanded map
  int64_t max_dw;      // per-patch elements of the largest depthwise output map
  int max_c_mid;
};

inline int same_pad_before(int i, int k, int s) {
  int o = (i + s - 1) / s;
  int total = (o - 1) * s + k - i;
  if (total < 0) total = 0;
  return total / 2;
}

inline NetCfg make_b0() {
  static const int stages[7][6] = {{1, 3, 1, 1, 32, 16},  {2, 3, 2, 6, 16, 24},   {2, 5, 2, 6, 24, 40},
                                   {3, 3, 2, 6, 40, 80},  {3, 5, 1, 6, 80, 112},  {4, 5, 2, 6, 112, 192},
                                   {1, 3, 1, 6, 192, 320}};
  NetCfg net;
  int64_t off = 0;
  // every segment starts on a 16-byte boundary (float4 loads)
  auto take = [&](int64_t n) {
    int64_t o = off;
    off += (n + 3) / 4 * 4;
    return o;
  };
  net.w_stem = take(27 * 32);
  net.s_stem = take(32);
  net.b_stem = take(32);
  int h = 112;
  net.max_in_out = (int64_t)112 * 112 * 32;
  net.max_mid = 0;
  net.max_dw = 0;
  net.max_c_mid = 0;
  for (int s = 0; s < 7; ++s) {
    for (int j = 0; j < stages[s][0]; ++j) {
      BlockCfg b{};
      b.k = stages[s][1];
      b.stride = j == 0 ? stages[s][2] : 1;
      b.expand = stages[s][3];
      b.c_in = j == 0 ? stages[s][4] : stages[s][5];
      b.c_out = stages[s][5];
      b.c_mid = b.c_in * b.expand;
      b.c_se = b.c_in / 4 > 1 ? b.c_in / 4 : 1;  // max(1, int(c_in * 0.25))
      b.h_in = h;
      b.h_out = (h + b.stride - 1) / b.stride;
      b.pad = same_pad_before(h, b.k, b.stride);
      b.skip = b.stride == 1 && b.c_in == b.c_out;
      if (b.expand != 1) {
        b.w_exp = take((int64_t)b.c_mid * b.c_in);
        b.s_exp = take(b.c_mid);
        b.b_exp = take(b.c_mid);
      } else {
        b.w_exp = b.s_exp = b.b_exp = -1;
      }
      b.w_dw = take((int64_t)b.k * b.k * b.c_mid);
      b.s_dw = take(b.c_mid);
      b.b_dw = take(b.c_mid);
      b.w_se1 = take((int64_t)b.c_se * b.c_mid);
      b.b_se1 = take(b.c_se);
      b.w_se2 = take((int64_t)b.c_mid * b.c_se);
      b.b_se2 = take(b.c_mid);
      b.w_proj = take((int64_t)b.c_out * b.c_mid);
      b.s_proj = take(b.c_out);
      b.b_proj = take(b.c_out);
      int64_t in_e = (int64_t)b.h_in * b.h_in * b.c_in, out_e = (int64_t)b.h_out * b.h_out * b.c_out;
      int64_t mid_e = (int64_t)b.h_in * b.h_in * b.c_mid, dw_e = (int64_t)b.h_out * b.h_out * b.c_mid;
      if (in_e > net.max_in_out) net.max_in_out = in_e;
      if (out_e > net.max_in_out) net.max_in_out = out_e;
      if (b.expand != 1 && mid_e > net.max_mid) net.max_mid = mid_e;
      if (dw_e > net.max_dw) net.max_dw = dw_e;
      if (b.c_mid > net.max_c_mid) net.max_c_mid = b.c_mid;
      h = b.h_out;
      net.blocks.push_back(b);
    }
  }
  net.w_head = take((int64_t)1280 * 320);
  net.s_head = take(1280);
  net.b_head = take(1280);
  net.n_params = off;
  return net;
}

}  // namespace mc
