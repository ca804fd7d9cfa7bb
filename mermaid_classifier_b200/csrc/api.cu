// C ABI of libmermaid_b200.so: handles, workspaces, layer orchestration.
// Declarations (with the reference interface each one replaces) live in include/mermaid_b200.h.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "crop_stem.cuh"
#include "dwconv_se.cuh"
#include "dw_tma.cuh"
#include "head.cuh"
#include "platt.cuh"
#include "layers.h"
#include "mlp_train.cuh"
#include "pw_simt.cuh"
#include "pw_tc.cuh"
#include "mbconv_fused.cuh"
#include "stem_tc.cuh"

using namespace mc;

#ifndef MC_FUSE_DEFAULT
#define MC_FUSE_DEFAULT 0x6u   // blocks fused by default in bf16 mode (bit b = block b): b1 (-33 % against expand + depthwise) and, since the
                               // persistent one-CTA-per-SM form of the fused kernel, b2 (+1.8 % on the step: 116.6 -> 118.7 k patches/s);
                               // b3 on top measures the same (118.8 k) and stays on the two-kernel path; b4 does not fit in fp32
#endif
#ifndef MC_FUSE_DEFAULT_FP32
// fp32 mode runs into the board's power cap (~1 kW, SM clock 1.72-1.75 GHz): with b2 and b3 fused as well the step moves
// 4.2 MB less per patch through HBM, the clock settles at 1.81-1.86 GHz and the whole step is 1.3-1.5 % faster (76.4 ->
// 77.5 k patches/s, interleaved runs on one box) although the two fused kernels are no faster than the pairs they replace.
// bf16 mode is not power-capped (1.965 GHz either way) and keeps b1 only.
#define MC_FUSE_DEFAULT_FP32 0xEu
#endif

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ToTensor + Normalize lookup table, built with the exact fp32 op order torch uses:
// (float(u8) / 255.f - mean) / std.
void build_norm_lut(float* lut) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      volatile float x = (float)v / 255.0f;
      volatile float y = x - mean[c];
      volatile float z = y / stdv[c];
      lut[c * 256 + v] = z;
    }
}

int check_points(const mc_image* images, int32_t n_images, const mc_point* pts, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const mc_point& p = pts[i];
    if (p.image < 0 || p.image >= n_images)
      return fail(MC_ERR_BAD_ARG, "point " + std::to_string(i) + " references image " + std::to_string(p.image));
    const mc_image& im = images[p.image];
    if (p.row < 0 || p.row > im.height - 1 || p.col < 0 || p.col > im.width - 1)
      return fail(MC_ERR_POINT_BOUNDS, "point (" + std::to_string(p.row) + ", " + std::to_string(p.col) +
                                           ") outside image of " + std::to_string(im.height) + " x " +
                                           std::to_string(im.width));
  }
  return MC_OK;
}

int check_images(const mc_image* images, int32_t n_images) {
  for (int i = 0; i < n_images; ++i) {
    const mc_image& im = images[i];
    if (!im.data || im.height < 1 || im.width < 1 || im.row_pitch < (int64_t)im.width * 3)
      return fail(MC_ERR_BAD_ARG, "image " + std::to_string(i) + " has a null pointer, empty shape or short pitch");
  }
  return MC_OK;
}

template <typename T>
int grow(T** p, int64_t* cap, int64_t need) {
  if (need <= *cap) return MC_OK;
  if (*p) MC_CUDA(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  int64_t n = std::max<int64_t>(need, 16);
  cudaError_t e = cudaMalloc((void**)p, (size_t)n * sizeof(T));
  if (e != cudaSuccess) return fail(MC_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  *cap = n;
  return MC_OK;
}

}  // namespace

// =====================================================================================
// extractor
// =====================================================================================
struct HostPipe;   // host_pipe.inl: staging ring + streams of mc_extract_images_host
namespace {
void host_pipe_free(HostPipe* p);
}
struct mc_extractor {
  int device = 0, mode = 0, max_batch = 0;
  NetCfg net;
  float* d_params = nullptr;
  float* d_lut = nullptr;
  void *bufX = nullptr, *bufY = nullptr, *bufE = nullptr, *bufD = nullptr, *bufH = nullptr;
  float *d_pool = nullptr, *d_gate = nullptr;
  __nv_bfloat16* d_gate_h = nullptr;  // bf16 copy of the SE gate (bf16 mode)
  mc_image* d_images = nullptr;
  int64_t cap_images = 0;
  mc_point* d_points = nullptr;
  int64_t cap_points = 0;
  uint8_t* d_img = nullptr;  // staging image for mc_extract_image_host
  int64_t cap_img = 0;
  float* d_feats = nullptr;  // staging features for mc_extract_image_host
  int64_t cap_feats = 0;
  std::vector<mc_point> h_points;
  PwTcPlan* tc = nullptr;  // tcgen05 GEMM plans (pw_tc.cuh)
  StemParams stem;          // stem weights + folded BN, passed by value (kernel parameter / constant bank): CUDA-core stem (MC_STEM_SIMT)
  StemTcPlan stem_tc;       // tensor-core stem (stem_tc.cuh), the default
  bool stem_simt = false;
  std::vector<DwLayer> dw;  // TMA-staged depthwise plans (dw_tma.cuh), one per block
  std::vector<FusedLayer> fused;  // expand + depthwise in one kernel (mbconv_fused.cuh); fuse_mask bit b = block b
  unsigned fuse_mask = 0;
  HostPipe* pipe = nullptr;        // mc_extract_images_host
  int64_t launches = 0;
  int64_t l2_budget = 0;  // bytes of per-chunk working set kept L2-resident (0 = no chunking)
  bool no_pool_fusion = false;   // MC_NO_POOL_FUSION: head conv and average pool as two launches (A/B timing)
  bool sparse_h2d = true;        // MC_SPARSE_H2D=0: the host pipeline uploads every image whole (A/B of the per-point window upload)
  int tap_layer = -1;
  float* tap_out = nullptr;
  int64_t tap_cap = 0;
  // per-layer CUDA-event profiler (mc_extractor_profile): -1 off, -2 every layer, >= 0 one layer
  int prof_layer = -1;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_ids;
  size_t prof_used = 0;
  double prof_ms[MC_N_LAYERS] = {0};
  int64_t prof_cnt[MC_N_LAYERS] = {0};
};

namespace {

constexpr int MAX_BANDS = 14;

struct ProfScope {
  mc_extractor* h;
  cudaStream_t st;
  bool on;
  size_t slot = 0;
  ProfScope(mc_extractor* h_, int layer, cudaStream_t st_) : h(h_), st(st_) {
    on = h->prof_layer == -2 || h->prof_layer == layer;
    if (!on) return;
    if (h->prof_used + 2 > h->prof_ev.size()) {
      for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->prof_ev.push_back(e);
      }
    }
    slot = h->prof_used;
    h->prof_used += 2;
    h->prof_ids.push_back(layer);
    cudaEventRecord(h->prof_ev[slot], st);
  }
  ~ProfScope() {
    if (on) cudaEventRecord(h->prof_ev[slot + 1], st);
  }
};

int prof_collect(mc_extractor* h, cudaStream_t st) {
  if (h->prof_used == 0) return MC_OK;
  MC_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < h->prof_ids.size(); ++i) {
    float ms = 0.f;
    MC_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
    h->prof_ms[h->prof_ids[i]] += ms;
    h->prof_cnt[h->prof_ids[i]] += 1;
  }
  h->prof_used = 0;
  h->prof_ids.clear();
  return MC_OK;
}

template <typename T>
int launch_dw(mc_extractor* h, const BlockCfg& b, int n_off, T* out, int nb, cudaStream_t st) {
  const float* P = h->d_params;
  const int C = b.c_mid;
  const int bidx = (int)(&b - &h->net.blocks[0]);
  DwLayer& l = h->dw[bidx];
  const int nparts = l.nbands * l.nxc;
  if (l.lane) {
    ProfScope ps_dw(h, 2 + 4 * bidx, st);
    DwRegArgs a;
    a.w = P + b.w_dw;
    a.scale = P + b.s_dw;
    a.bias = P + b.b_dw;
    a.out = out;
    a.pool_partial = h->d_pool;
    a.n_off = n_off;
    a.nb = nb;
    if (int rc = dw_reg_launch<T>(l, l.tm, a, nb, st)) return rc;
  } else {
    ProfScope ps_dw(h, 2 + 4 * bidx, st);
    DwArgs a;
    a.w = P + b.w_dw;
    a.scale = P + b.s_dw;
    a.bias = P + b.b_dw;
    a.out = out;
    a.pool_partial = h->d_pool;
    a.C = C;
    a.Hin = b.h_in;
    a.Hout = b.h_out;
    a.pad = b.pad;
    a.rows_per_band = l.rows_per_band;
    a.cgt = l.cgt;
    a.pt = l.pt;
    a.bwin = l.bwin;
    a.stages = l.stages;
    a.row_bytes = l.row_bytes;
    a.box_bytes = l.box_bytes;
    a.n_off = n_off;
    if (int rc = dw_tma_launch<T>(l, l.tm, a, nb, st)) return rc;
  }
  MC_CHECK_LAUNCH();
  h->launches++;
  {
    ProfScope ps(h, 3 + 4 * bidx, st);
    se_launch(h->d_pool, nparts, 1.f / (float)(b.h_out * b.h_out), P + b.w_se1, P + b.b_se1, P + b.w_se2, P + b.b_se2,
              h->d_gate, h->d_gate_h, C, b.c_se, nb, st);
  }
  MC_CHECK_LAUNCH();
  h->launches++;
  return MC_OK;
}

// SIMT fp32-exact pointwise conv.
template <typename T, int ACT, bool GATE, bool RES>
int launch_pw_simt(mc_extractor* h, const T* A, const float* W, const float* sc, const float* bi, const float* gate,
                   const T* res, T* out, int64_t M, int N, int K, int HW, cudaStream_t st) {
  dim3 grid(cdiv(M, 64), cdiv(N, 64));
  pw_simt_kernel<T, T, ACT, GATE, RES><<<grid, 256, 0, st>>>(A, W, sc, bi, gate, res, out, M, N, K, HW);
  MC_CHECK_LAUNCH();
  h->launches++;
  return MC_OK;
}

// MC_DEBUG_SYNC=1: synchronise after every layer and name the layer that faulted (bring-up aid).
int debug_sync(int layer, cudaStream_t st) {
  static const bool on = getenv("MC_DEBUG_SYNC") != nullptr;
  if (!on) return MC_OK;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "layer " + std::to_string(layer) + ": " + cudaGetErrorString(e));
  return MC_OK;
}

template <typename T>
int tap(mc_extractor* h, int layer, const T* src, int64_t n_elems, cudaStream_t st) {
  if (int rc = debug_sync(layer, st)) return rc;
  if (h->tap_layer != layer || !h->tap_out) return MC_OK;
  const int64_t n = std::min(n_elems, h->tap_cap);
  to_f32_kernel<T><<<cdiv(n, 256), 256, 0, st>>>(src, h->tap_out, n);
  MC_CHECK_LAUNCH();
  h->tap_layer = -1;
  return MC_OK;
}

// One sub-batch: points already on the device (h->d_points[0..nb)), images table in h->d_images.
template <typename T>
int forward(mc_extractor* h, int nb, float* feats_dev, cudaStream_t st) {
  const float* P = h->d_params;
  const NetCfg& net = h->net;
  T* X = (T*)h->bufX;
  T* Y = (T*)h->bufY;
  T* E = (T*)h->bufE;
  T* D = (T*)h->bufD;
  int rc;
  {
    ProfScope ps(h, 0, st);
    if (h->stem_simt) stem_kernel<T><<<dim3(7, nb), 256, 0, st>>>(h->d_images, h->d_points, h->stem, h->d_lut, X);
    else if ((rc = stem_tc_launch<T>(h->stem_tc, h->d_images, h->d_points, P + net.s_stem, X, nb, st))) return rc;
  }
  MC_CHECK_LAUNCH();
  h->launches++;
  if ((rc = tap<T>(h, 0, X, (int64_t)nb * 112 * 112 * 32, st))) return rc;
  // L2-resident chunked execution: each MBConv block runs expand -> depthwise -> SE -> project over
  // chunks of patches small enough that the expanded map E and the depthwise output D of a chunk
  // (which every chunk re-uses at the same addresses) stay in the 126 MB L2 between producer and
  // consumer; only the block input X and output Y stream through HBM.
  const bool tapping = h->tap_layer >= 0 && h->tap_out != nullptr;
  for (size_t bi = 0; bi < net.blocks.size(); ++bi) {
    const BlockCfg& b = net.blocks[bi];
    const int HWi = b.h_in * b.h_in, HW = b.h_out * b.h_out;
    const int64_t per_patch = ((int64_t)HWi * (b.c_in + (b.expand != 1 ? b.c_mid : 0)) + (int64_t)HW * (b.c_mid + b.c_out)) *
                              (int64_t)sizeof(T);
    int chunk = nb;
    if (h->l2_budget > 0 && !tapping) {
      chunk = (int)std::max<int64_t>(1, std::min<int64_t>(nb, h->l2_budget / per_patch));
      chunk = cdiv(nb, cdiv(nb, chunk));  // equal-sized chunks
    }
    for (int c0 = 0; c0 < nb; c0 += chunk) {
      const int cn = std::min(chunk, nb - c0);
      const int64_t Min = (int64_t)cn * HWi, Mout = (int64_t)cn * HW;
      const T* Xc = X + (int64_t)c0 * HWi * b.c_in;
      T* Yc = Y + (int64_t)c0 * HW * b.c_out;
      const T* dw_in = Xc;
      // fused expand + depthwise (mbconv_fused.cuh): whole batch, no taps (the expanded map does not exist to be tapped)
      if (b.expand != 1 && ((h->fuse_mask >> bi) & 1u) && chunk == nb && !tapping && h->fused[bi].present) {
        FusedLayer& fl = h->fused[bi];
        FusedArgs fa;
        fa.w_dw = P + b.w_dw;
        fa.s_dw = P + b.s_dw;
        fa.b_dw = P + b.b_dw;
        fa.b_exp = P + b.b_exp;
        fa.out = D;
        fa.pool_partial = h->d_pool;
        fa.nb = nb;
        fa.num_sms = h->tc ? h->tc->num_sms : 148;
        {
          ProfScope ps(h, 2 + 4 * (int)bi, st);
          if ((rc = fused_launch<T>(fl, Xc, b.c_in, b.h_in, fa, st))) return rc;
        }
        h->launches++;
        {
          ProfScope ps(h, 3 + 4 * (int)bi, st);
          se_launch(h->d_pool, fused_shape_of(fl.shape, (int)sizeof(T)).nbands, 1.f / (float)(b.h_out * b.h_out), P + b.w_se1,
                    P + b.b_se1, P + b.w_se2, P + b.b_se2, h->d_gate, h->d_gate_h, b.c_mid, b.c_se, nb, st);
        }
        MC_CHECK_LAUNCH();
        h->launches++;
        if ((rc = debug_sync(2 + 4 * (int)bi, st))) return rc;
        goto project;
      }
      if (b.expand != 1) {
        ProfScope ps(h, 1 + 4 * (int)bi, st);
        if (h->tc && pw_tc_has(h->tc, (int)bi * 2)) {
          if ((rc = pw_tc_run(h->tc, (int)bi * 2, X, (int64_t)c0 * HWi, nullptr, nullptr, E, Min, HWi, st))) return rc;
          h->launches++;
        } else if ((rc = launch_pw_simt<T, ACT_SILU, false, false>(h, Xc, P + b.w_exp, P + b.s_exp, P + b.b_exp, nullptr,
                                                                   nullptr, E, Min, b.c_mid, b.c_in, 1, st)))
          return rc;
        dw_in = E;
      }
      if (b.expand != 1 && (rc = tap<T>(h, 1 + 4 * (int)bi, E, Min * b.c_mid, st))) return rc;
      if ((rc = launch_dw<T>(h, b, b.expand != 1 ? 0 : c0, D, cn, st))) return rc;
      if ((rc = tap<T>(h, 2 + 4 * (int)bi, D, Mout * b.c_mid, st))) return rc;
      if ((rc = tap<float>(h, 3 + 4 * (int)bi, h->d_gate, (int64_t)cn * b.c_mid, st))) return rc;
    project:
      {
        ProfScope ps_proj(h, 4 + 4 * (int)bi, st);
        if (h->tc && pw_tc_has(h->tc, (int)bi * 2 + 1)) {
          const void* gp = h->mode == MC_MODE_BF16 ? (const void*)h->d_gate_h : (const void*)h->d_gate;
          if ((rc = pw_tc_run(h->tc, (int)bi * 2 + 1, D, 0, gp, b.skip ? Xc : nullptr, Yc, Mout, HW, st))) return rc;
          h->launches++;
        } else if (b.skip) {
          if ((rc = launch_pw_simt<T, ACT_NONE, true, true>(h, D, P + b.w_proj, P + b.s_proj, P + b.b_proj, h->d_gate, Xc, Yc,
                                                            Mout, b.c_out, b.c_mid, HW, st)))
            return rc;
        } else {
          if ((rc = launch_pw_simt<T, ACT_NONE, true, false>(h, D, P + b.w_proj, P + b.s_proj, P + b.b_proj, h->d_gate,
                                                             nullptr, Yc, Mout, b.c_out, b.c_mid, HW, st)))
            return rc;
        }
      }
      if ((rc = tap<T>(h, 4 + 4 * (int)bi, Yc, Mout * b.c_out, st))) return rc;
    }
    std::swap(X, Y);
  }
  const int64_t Mh = (int64_t)nb * 49;
  // K7: head conv + BN + swish + global average pool in ONE launch (pw_tc_kernel POOL instantiation): the 49 x 1280 map
  // per patch never reaches HBM.  The two-launch form (conv to bufH, then avgpool_kernel) only remains for the layer-65
  // tap and for MC_TC_MASK bisecting with the CUDA-core GEMM; bufH is allocated on first use.
  const bool want_map = (h->tap_layer == 65 && h->tap_out != nullptr) || !(h->tc && pw_tc_has(h->tc, 32)) || h->no_pool_fusion;
  if (!want_map) {
    ProfScope ps_head(h, 65, st);
    if ((rc = pw_tc_run_pool(h->tc, 32, X, nb, feats_dev, st))) return rc;
    h->launches++;
    return debug_sync(65, st);
  }
  if (!h->bufH) {
    const size_t es = h->mode == MC_MODE_FP32 ? 4 : 2;
    cudaError_t e = cudaMalloc(&h->bufH, (size_t)h->max_batch * 49 * 1280 * es);
    if (e != cudaSuccess) return fail(MC_ERR_NOMEM, std::string("cudaMalloc head-conv map: ") + cudaGetErrorString(e));
  }
  T* Hb = (T*)h->bufH;
  {
    ProfScope ps_head(h, 65, st);
    if (h->tc && pw_tc_has(h->tc, 32)) {
      if ((rc = pw_tc_run(h->tc, 32, X, 0, nullptr, nullptr, Hb, Mh, 49, st))) return rc;
      h->launches++;
    } else if ((rc = launch_pw_simt<T, ACT_SILU, false, false>(h, X, P + net.w_head, P + net.s_head, P + net.b_head,
                                                               nullptr, nullptr, Hb, Mh, 1280, 320, 1, st)))
      return rc;
  }
  if ((rc = tap<T>(h, 65, Hb, Mh * 1280, st))) return rc;
  {
    ProfScope ps(h, 66, st);
    avgpool_kernel<T><<<dim3(cdiv(1280, 128), nb), 128, 0, st>>>(Hb, feats_dev, 49, 1280);
  }
  MC_CHECK_LAUNCH();
  h->launches++;
  return MC_OK;
}

int forward_any(mc_extractor* h, int nb, float* feats_dev, cudaStream_t st) {
  return h->mode == MC_MODE_FP32 ? forward<float>(h, nb, feats_dev, st)
                                 : forward<__nv_bfloat16>(h, nb, feats_dev, st);
}

}  // namespace

extern "C" {

int mc_abi_version(void) { return MC_ABI_VERSION; }
const char* mc_last_error(void) { return last_error().c_str(); }

int mc_synth_image(uint8_t* img_dev, int32_t height, int32_t width, int64_t row_pitch, uint32_t seed,
                   uint32_t image_id, void* stream) {
  if (!img_dev || height < 1 || width < 1 || height >= 65536 || width >= 65536 || row_pitch < (int64_t)width * 3)
    return fail(MC_ERR_BAD_ARG, "mc_synth_image: bad shape");
  const uint32_t key = mix32(seed ^ (image_id * 0x9E3779B1u));
  synth_image_kernel<<<dim3(cdiv(width, 256), height), 256, 0, (cudaStream_t)stream>>>(img_dev, height, width,
                                                                                        row_pitch, key);
  MC_CHECK_LAUNCH();
  return MC_OK;
}

int mc_check_extract_inputs(int32_t height, int32_t width, const int32_t* rowcols, int64_t n, int64_t max_pixels,
                            int64_t max_points) {
  if (max_pixels <= 0) max_pixels = 100000000;
  if (max_points <= 0) max_points = 1000;
  if ((int64_t)height * width > max_pixels)
    return fail(MC_ERR_DATA_LIMIT, "image has " + std::to_string((int64_t)height * width) + " pixels, max " +
                                       std::to_string(max_pixels));
  if (n > max_points)
    return fail(MC_ERR_DATA_LIMIT, std::to_string(n) + " points, max " + std::to_string(max_points));
  for (int64_t i = 0; i < n; ++i) {
    const int r = rowcols[2 * i], c = rowcols[2 * i + 1];
    if (r < 0 || r > height - 1)
      return fail(MC_ERR_POINT_BOUNDS, "row " + std::to_string(r) + " outside [0, " + std::to_string(height - 1) + "]");
    if (c < 0 || c > width - 1)
      return fail(MC_ERR_POINT_BOUNDS, "col " + std::to_string(c) + " outside [0, " + std::to_string(width - 1) + "]");
  }
  return MC_OK;
}

int mc_crop_patches(const mc_image* images, int32_t n_images, const mc_point* points, int64_t n, uint8_t* patches_dev,
                    void* stream) {
  if (n == 0) return MC_OK;
  if (!images || !points || !patches_dev || n_images < 1 || n < 0) return fail(MC_ERR_BAD_ARG, "mc_crop_patches: null");
  int rc;
  if ((rc = check_images(images, n_images)) || (rc = check_points(images, n_images, points, n))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  mc_image* d_im = nullptr;
  mc_point* d_pt = nullptr;
  MC_CUDA(cudaMallocAsync((void**)&d_im, n_images * sizeof(mc_image), st));
  MC_CUDA(cudaMallocAsync((void**)&d_pt, n * sizeof(mc_point), st));
  MC_CUDA(cudaMemcpyAsync(d_im, images, n_images * sizeof(mc_image), cudaMemcpyHostToDevice, st));
  MC_CUDA(cudaMemcpyAsync(d_pt, points, n * sizeof(mc_point), cudaMemcpyHostToDevice, st));
  for (int64_t s = 0; s < n; s += 32768) {
    const int nb = (int)std::min<int64_t>(32768, n - s);
    crop_kernel<<<dim3(8, nb), 192, 0, st>>>(d_im, d_pt + s, patches_dev + s * 224 * 224 * 3);
    MC_CHECK_LAUNCH();
  }
  MC_CUDA(cudaFreeAsync(d_im, st));
  MC_CUDA(cudaFreeAsync(d_pt, st));
  return MC_OK;
}

int mc_crop_resize_patches(const mc_image* images, int32_t n_images, const mc_point* points, int64_t n, int32_t crop_size,
                           uint8_t* patches_dev, void* stream) {
  if (n == 0) return MC_OK;
  if (!images || !points || !patches_dev || n_images < 1 || n < 0) return fail(MC_ERR_BAD_ARG, "mc_crop_resize_patches: null");
  if (crop_size < 2 || crop_size > 4096 || (crop_size & 1))
    return fail(MC_ERR_BAD_ARG, "mc_crop_resize_patches: crop_size must be an even number in [2, 4096]");
  if (crop_size == 224) return mc_crop_patches(images, n_images, points, n, patches_dev, stream);
  int rc;
  if ((rc = check_images(images, n_images)) || (rc = check_points(images, n_images, points, n))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  mc_image* d_im = nullptr;
  mc_point* d_pt = nullptr;
  MC_CUDA(cudaMallocAsync((void**)&d_im, n_images * sizeof(mc_image), st));
  MC_CUDA(cudaMallocAsync((void**)&d_pt, n * sizeof(mc_point), st));
  MC_CUDA(cudaMemcpyAsync(d_im, images, n_images * sizeof(mc_image), cudaMemcpyHostToDevice, st));
  MC_CUDA(cudaMemcpyAsync(d_pt, points, n * sizeof(mc_point), cudaMemcpyHostToDevice, st));
  for (int64_t s = 0; s < n; s += 32768) {
    const int nb = (int)std::min<int64_t>(32768, n - s);
    crop_resize_kernel<<<dim3(28, nb), 224, 0, st>>>(d_im, d_pt + s, crop_size, patches_dev + s * 224 * 224 * 3);
    MC_CHECK_LAUNCH();
  }
  MC_CUDA(cudaFreeAsync(d_im, st));
  MC_CUDA(cudaFreeAsync(d_pt, st));
  return MC_OK;
}

int mc_normalize_patches(const uint8_t* patches_dev, int64_t n, float* out_dev, void* stream) {
  if (n == 0) return MC_OK;
  if (!patches_dev || !out_dev || n < 0) return fail(MC_ERR_BAD_ARG, "mc_normalize_patches: null");
  cudaStream_t st = (cudaStream_t)stream;
  float lut[768];
  build_norm_lut(lut);
  float* d_lut = nullptr;
  MC_CUDA(cudaMallocAsync((void**)&d_lut, sizeof(lut), st));
  MC_CUDA(cudaMemcpyAsync(d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice, st));
  MC_CUDA(cudaStreamSynchronize(st));  // `lut` is a stack array
  const int64_t total = n * 224 * 224;
  normalize_kernel<<<cdiv(total, 256), 256, 0, st>>>(patches_dev, d_lut, out_dev, total);
  MC_CHECK_LAUNCH();
  MC_CUDA(cudaFreeAsync(d_lut, st));
  return MC_OK;
}

int64_t mc_backbone_param_count(void) {
  static const int64_t n = make_b0().n_params;
  return n;
}

int mc_extractor_create(const float* params, int64_t n_params, int32_t mode, int32_t device, int32_t max_batch,
                        mc_extractor** out) {
  if (!params || !out) return fail(MC_ERR_BAD_ARG, "mc_extractor_create: null argument");
  if (mode != MC_MODE_FP32 && mode != MC_MODE_BF16) return fail(MC_ERR_BAD_ARG, "mc_extractor_create: bad mode");
  if (max_batch < 1 || max_batch > 65535) return fail(MC_ERR_BAD_ARG, "mc_extractor_create: max_batch in [1, 65535]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MC_ERR_CUDA, "no CUDA device: libmermaid_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(MC_ERR_BAD_ARG, "mc_extractor_create: bad device index");
  NetCfg net = make_b0();
  if (n_params != net.n_params)
    return fail(MC_ERR_BAD_ARG, "parameter blob has " + std::to_string(n_params) + " floats, expected " +
                                    std::to_string(net.n_params));
  DeviceGuard g(device);
  cudaDeviceProp prop;
  MC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MC_ERR_UNSUPPORTED, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                        std::to_string(prop.minor) + "; this library is built for sm_100a (B200) only");
  mc_extractor* h = new mc_extractor();
  h->device = device;
  h->mode = mode;
  h->max_batch = max_batch;
  h->net = net;
  const size_t es = mode == MC_MODE_FP32 ? 4 : 2;
  const int64_t nb = max_batch;
  auto dmalloc = [&](void** p, size_t bytes) -> int {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(MC_ERR_NOMEM, std::string("cudaMalloc workspace: ") + cudaGetErrorString(e));
    return MC_OK;
  };
  int rc = MC_OK;
  float lut[768];
  build_norm_lut(lut);
  if ((rc = dmalloc((void**)&h->d_params, n_params * sizeof(float))) ||
      (rc = dmalloc((void**)&h->d_lut, sizeof(lut))) ||
      (rc = dmalloc(&h->bufX, nb * net.max_in_out * es)) || (rc = dmalloc(&h->bufY, nb * net.max_in_out * es)) ||
      (rc = dmalloc(&h->bufE, nb * net.max_mid * es)) || (rc = dmalloc(&h->bufD, nb * net.max_dw * es)) ||
      (rc = dmalloc((void**)&h->d_pool, nb * MAX_BANDS * net.max_c_mid * sizeof(float))) ||
      (rc = dmalloc((void**)&h->d_gate, nb * net.max_c_mid * sizeof(float))) ||
      (mode == MC_MODE_BF16 && (rc = dmalloc((void**)&h->d_gate_h, nb * net.max_c_mid * sizeof(__nv_bfloat16))))) {
    mc_extractor_destroy(h);
    return rc;
  }
  cudaError_t e = cudaMemcpy(h->d_params, params, n_params * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    mc_extractor_destroy(h);
    return fail(MC_ERR_CUDA, std::string("parameter upload: ") + cudaGetErrorString(e));
  }
  // MC_L2_MB: working-set budget of the chunked execution in MB.  Default 0 = off: measured on B200, the
  // per-launch latency chain of ~900 small dependent launches per sub-batch costs more than the HBM
  // traffic it saves (25 k vs 44 k patches/s at 64 MB); kept as an experiment switch.
  h->l2_budget = 0;
  if (const char* env = getenv("MC_L2_MB")) h->l2_budget = (int64_t)atoll(env) << 20;
  // MC_TC_MASK (hex, bit 2b = expand of block b, 2b+1 = project, 32 = head conv) selects which 1x1
  // convs run on the tcgen05 kernel; default: all of them.  Bring-up / bisecting aid only.
  unsigned long long tc_mask = ~0ull;
  if (const char* env = getenv("MC_TC_MASK")) tc_mask = strtoull(env, nullptr, 16);
  // MC_FUSE_MASK (hex, bit b = block b): which MBConv blocks run expand + depthwise as one kernel (mbconv_fused.cuh)
  h->no_pool_fusion = getenv("MC_NO_POOL_FUSION") != nullptr;
  h->fuse_mask = mode == MC_MODE_FP32 ? MC_FUSE_DEFAULT_FP32 : MC_FUSE_DEFAULT;
  if (const char* env = getenv("MC_FUSE_MASK")) h->fuse_mask = (unsigned)strtoul(env, nullptr, 16);
  if (const char* env = getenv("MC_SPARSE_H2D")) h->sparse_h2d = atoi(env) != 0;
  if ((rc = pw_tc_build(&h->tc, h->net, params, h->d_params, mode, max_batch, device, (unsigned)(tc_mask & 0xffffffffu),
                        (unsigned)(tc_mask >> 32)))) {
    mc_extractor_destroy(h);
    return rc;
  }
  h->stem_simt = getenv("MC_STEM_SIMT") != nullptr;   // A/B switch: the CUDA-core stem kernel of round 1
  if ((rc = stem_tc_plan(&h->stem_tc, params + h->net.w_stem, params + h->net.s_stem, params + h->net.b_stem, device))) {
    mc_extractor_destroy(h);
    return rc;
  }
  memcpy(h->stem.w, params + h->net.w_stem, sizeof(h->stem.w));
  memcpy(h->stem.scale, params + h->net.s_stem, sizeof(h->stem.scale));
  memcpy(h->stem.bias, params + h->net.b_stem, sizeof(h->stem.bias));
  h->fused.resize(h->net.blocks.size());
  for (size_t bi = 0; bi < h->net.blocks.size(); ++bi) {
    const BlockCfg& b = h->net.blocks[bi];
    if (b.expand != 1 && ((h->fuse_mask >> bi) & 1u) &&
        (rc = fused_plan_layer(&h->fused[bi], mode == MC_MODE_FP32, b.k, b.stride, b.c_in, b.c_mid, b.h_in, params + b.w_exp,
                               params + b.s_exp))) {
      mc_extractor_destroy(h);
      return rc;
    }
  }
  h->dw.resize(h->net.blocks.size());
  for (size_t bi = 0; bi < h->net.blocks.size(); ++bi) {
    const BlockCfg& b = h->net.blocks[bi];
    DwLayer& l = h->dw[bi];
    l.in_ptr = b.expand != 1 ? h->bufE : h->bufX;  // block 0 (no expand) reads the stem output
    if ((rc = dw_plan_layer(&l, b, mode == MC_MODE_FP32)) ||
        (rc = dw_make_map(&l.tm, mode == MC_MODE_FP32, l.in_ptr, b.c_mid, b.h_in, max_batch, l.lane ? l.cb : l.cgt * 4, l.bwin))) {
      mc_extractor_destroy(h);
      return rc;
    }
  }
  *out = h;
  return MC_OK;
}

int mc_extractor_destroy(mc_extractor* h) {
  if (!h) return MC_OK;
  DeviceGuard g(h->device);
  pw_tc_free(h->tc);
  host_pipe_free(h->pipe);
  stem_tc_free(h->stem_tc);
  for (auto& fl : h->fused) fused_free(fl);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  void* ptrs[] = {h->d_params, h->d_lut, h->bufX, h->bufY,    h->bufE, h->bufD,
                  h->bufH,     h->d_pool, h->d_gate, h->d_gate_h, h->d_images, h->d_points, h->d_img, h->d_feats};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete h;
  return MC_OK;
}

int mc_extractor_mode(const mc_extractor* h) { return h ? h->mode : -1; }
int64_t mc_extractor_launches(const mc_extractor* h) { return h ? h->launches : 0; }

int mc_extractor_profile(mc_extractor* h, int32_t layer) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (layer < -2 || layer >= MC_N_LAYERS) return fail(MC_ERR_BAD_ARG, "mc_extractor_profile: bad layer");
  h->prof_layer = layer;
  return MC_OK;
}

int mc_extractor_profile_read(mc_extractor* h, double* ms_out, int64_t* count_out, int32_t reset) {
  if (!h || !ms_out || !count_out) return fail(MC_ERR_BAD_ARG, "null argument");
  for (int i = 0; i < MC_N_LAYERS; ++i) {
    ms_out[i] = h->prof_ms[i];
    count_out[i] = h->prof_cnt[i];
    if (reset) {
      h->prof_ms[i] = 0;
      h->prof_cnt[i] = 0;
    }
  }
  return MC_OK;
}

int mc_extractor_set_tap(mc_extractor* h, int32_t layer, float* out_dev, int64_t capacity) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  h->tap_layer = layer;
  h->tap_out = out_dev;
  h->tap_cap = capacity;
  return MC_OK;
}

int mc_extract_points(mc_extractor* h, const mc_image* images, int32_t n_images, const mc_point* points, int64_t n,
                      float* feats_dev, void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (n == 0) return MC_OK;
  if (!images || !points || !feats_dev || n_images < 1 || n < 0) return fail(MC_ERR_BAD_ARG, "mc_extract_points: null");
  int rc;
  if ((rc = check_images(images, n_images)) || (rc = check_points(images, n_images, points, n))) return rc;
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = grow(&h->d_images, &h->cap_images, n_images)) || (rc = grow(&h->d_points, &h->cap_points, n))) return rc;
  MC_CUDA(cudaMemcpyAsync(h->d_images, images, n_images * sizeof(mc_image), cudaMemcpyHostToDevice, st));
  MC_CUDA(cudaMemcpyAsync(h->d_points, points, n * sizeof(mc_point), cudaMemcpyHostToDevice, st));
  mc_point* base = h->d_points;
  for (int64_t s = 0; s < n; s += h->max_batch) {
    const int nb = (int)std::min<int64_t>(h->max_batch, n - s);
    h->d_points = base + s;
    rc = forward_any(h, nb, feats_dev + s * MC_FEATURE_DIM, st);
    h->d_points = base;
    if (rc) return rc;
  }
  return prof_collect(h, st);
}

int mc_extract_patches(mc_extractor* h, const uint8_t* patches_dev, int64_t n, float* feats_dev, void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (n == 0) return MC_OK;
  if (!patches_dev || !feats_dev || n < 0) return fail(MC_ERR_BAD_ARG, "mc_extract_patches: null");
  // A pre-cropped patch is a 224x224 image whose point is its centre: no reflection is ever taken.
  std::vector<mc_image> ims((size_t)n);
  std::vector<mc_point> pts((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    ims[i] = mc_image{patches_dev + i * 224 * 224 * 3, 224, 224, 224 * 3};
    pts[i] = mc_point{(int32_t)i, 112, 112};
  }
  int rc = mc_extract_points(h, ims.data(), (int32_t)n, pts.data(), n, feats_dev, stream);
  if (rc == MC_OK) MC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // host tables die with this frame
  return rc;
}

int mc_extract_image_host(mc_extractor* h, const uint8_t* img_host, int32_t height, int32_t width, int64_t row_pitch,
                          const int32_t* rowcols, int64_t n, float* feats_host, void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (!img_host || height < 1 || width < 1 || row_pitch < (int64_t)width * 3 || n < 0 || (n > 0 && (!rowcols || !feats_host)))
    return fail(MC_ERR_BAD_ARG, "mc_extract_image_host: bad argument");
  if (n == 0) return MC_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const int64_t dpitch = ((int64_t)width * 3 + 255) / 256 * 256;
  if ((rc = grow(&h->d_img, &h->cap_img, dpitch * height)) || (rc = grow(&h->d_feats, &h->cap_feats, n * MC_FEATURE_DIM)))
    return rc;
  h->h_points.resize((size_t)n);
  for (int64_t i = 0; i < n; ++i) h->h_points[i] = mc_point{0, rowcols[2 * i], rowcols[2 * i + 1]};
  mc_image im{h->d_img, height, width, dpitch};
  if ((rc = check_points(&im, 1, h->h_points.data(), n))) return rc;
  MC_CUDA(cudaMemcpy2DAsync(h->d_img, dpitch, img_host, row_pitch, (size_t)width * 3, height, cudaMemcpyHostToDevice, st));
  if ((rc = mc_extract_points(h, &im, 1, h->h_points.data(), n, h->d_feats, st))) return rc;
  MC_CUDA(cudaMemcpyAsync(feats_host, h->d_feats, n * MC_FEATURE_DIM * sizeof(float), cudaMemcpyDeviceToHost, st));
  MC_CUDA(cudaStreamSynchronize(st));
  return MC_OK;
}

}  // extern "C"

// =====================================================================================
// head scoring
// =====================================================================================
struct mc_head {
  int device = 0, n_layers = 0;
  std::vector<int> dims, dims_p;     // logical and padded (multiple of 4) widths
  std::vector<float*> d_w, d_b;      // padded (dims_p[i+1] x dims_p[i]) / dims_p[i+1]
  float *d_a = nullptr, *d_pb = nullptr;
  std::vector<float*> d_act;         // per-layer activation chunks
  float* d_in_pad = nullptr;
  int64_t chunk = 0;
  float* d_feats = nullptr;  // host-call staging
  double* d_proba = nullptr;
  int32_t* d_labels = nullptr;
  int64_t cap_feats = 0, cap_proba = 0, cap_labels = 0;
  int64_t launches = 0;
  // tensor-core Linear chain (tcgen05, 3xTF32): used by the device scoring path unless `exact` is requested
  PwTcPlan* tc = nullptr;
  float* d_ones = nullptr;
  // mc_head_evaluate workspaces: per-row log-loss terms and labels, reduction partials, [loss_sum | hits]
  double* d_row_loss = nullptr;
  int32_t* d_eval_labels = nullptr;
  int64_t cap_row_loss = 0, cap_eval_labels = 0;
  double* d_eval_parts = nullptr;  // EVAL_PARTS doubles, EVAL_PARTS int64, then the two results
  bool exact = false;   // MC_HEAD_EXACT / mc_head_set_exact: run the Linear chain on the exact-fp32 CUDA-core GEMM
};

extern "C" {

int mc_head_create(int32_t n_layers, const int32_t* dims, const float* const* weights, const float* const* biases,
                   const float* pa, const float* pb, int32_t device, mc_head** out) {
  if (n_layers < 1 || !dims || !weights || !biases || !out) return fail(MC_ERR_BAD_ARG, "mc_head_create: null/empty");
  if ((pa == nullptr) != (pb == nullptr)) return fail(MC_ERR_BAD_ARG, "mc_head_create: a and b must both be given or both NULL");
  for (int i = 0; i <= n_layers; ++i)
    if (dims[i] < 1) return fail(MC_ERR_BAD_ARG, "mc_head_create: non-positive layer width");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MC_ERR_CUDA, "no CUDA device: libmermaid_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(MC_ERR_BAD_ARG, "mc_head_create: bad device index");
  DeviceGuard g(device);
  mc_head* h = new mc_head();
  h->device = device;
  h->n_layers = n_layers;
  h->dims.assign(dims, dims + n_layers + 1);
  for (int d : h->dims) h->dims_p.push_back((d + 3) / 4 * 4);
  const int K = h->dims.back();
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < n_layers && e == cudaSuccess; ++i) {
    const int ki = h->dims[i], no = h->dims[i + 1], kp = h->dims_p[i], np_ = h->dims_p[i + 1];
    std::vector<float> wp((size_t)np_ * kp, 0.f), bp((size_t)np_, 0.f);
    for (int r = 0; r < no; ++r) memcpy(&wp[(size_t)r * kp], weights[i] + (size_t)r * ki, ki * sizeof(float));
    memcpy(bp.data(), biases[i], no * sizeof(float));
    float *dw = nullptr, *db = nullptr;
    e = cudaMalloc((void**)&dw, wp.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&db, bp.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(dw, wp.data(), wp.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice);
    h->d_w.push_back(dw);
    h->d_b.push_back(db);
  }
  if (e == cudaSuccess && pa) {
    e = cudaMalloc((void**)&h->d_a, K * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_pb, K * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_a, pa, K * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_pb, pb, K * sizeof(float), cudaMemcpyHostToDevice);
  }
  // activation chunk: bound the workspace to ~1 GB (the caller's rows are read in place unless the width needs padding).
  // At (200, 100, 500) that is 131 072 rows per chunk: with the 32 k-row chunks of round 1 the four launches of a chunk were
  // four to five waves of work items each, and launch gaps + tails were most of the 165 us a chunk took.
  int64_t per_row = h->dims_p[0] != h->dims[0] ? h->dims_p[0] : 0;
  for (int i = 1; i <= n_layers; ++i) per_row += h->dims_p[i];
  h->chunk = std::max<int64_t>(1024, std::min<int64_t>(1 << 17, (int64_t)(1024ll << 20) / (per_row * 4)));
  for (int i = 1; i <= n_layers && e == cudaSuccess; ++i) {
    float* p = nullptr;
    e = cudaMalloc((void**)&p, h->chunk * h->dims_p[i] * sizeof(float));
    h->d_act.push_back(p);
  }
  if (e == cudaSuccess && h->dims_p[0] != h->dims[0])
    e = cudaMalloc((void**)&h->d_in_pad, h->chunk * h->dims_p[0] * sizeof(float));
  if (e != cudaSuccess) {
    mc_head_destroy(h);
    return fail(MC_ERR_CUDA, std::string("mc_head_create: ") + cudaGetErrorString(e));
  }
  // tcgen05 plan of the Linear chain: layer i = (dims_p[i+1] x dims_p[i]) padded weights, scale 1, bias, ReLU between layers
  {
    int max_w = 0;
    for (int d : h->dims_p) max_w = std::max(max_w, d);
    cudaDeviceProp prop;
    bool ok = cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.major == 10 && max_w <= 1280 && getenv("MC_HEAD_EXACT") == nullptr;
    if (ok) {
      std::vector<float> ones((size_t)max_w, 1.f);
      ok = cudaMalloc((void**)&h->d_ones, ones.size() * sizeof(float)) == cudaSuccess &&
           cudaMemcpy(h->d_ones, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (ok) {
      PwTcPlan* plan = new PwTcPlan();
      plan->mode = MC_MODE_FP32;
      plan->relu_variant = true;
      plan->device = device;
      plan->num_sms = prop.multiProcessorCount;
      plan->max_batch = (int)h->chunk;
      plan->layers.resize(n_layers);
      for (int i = 0; i < n_layers && ok; ++i) {
        const int ki = h->dims[i], no = h->dims[i + 1], kp = h->dims_p[i], np_ = h->dims_p[i + 1];
        std::vector<float> wp((size_t)np_ * kp, 0.f);
        for (int r = 0; r < no; ++r) memcpy(&wp[(size_t)r * kp], weights[i] + (size_t)r * ki, ki * sizeof(float));
        ok = pw_tc_add(plan, i, wp.data(), h->d_ones, h->d_b[i], np_, kp, i < n_layers - 1 ? 2 : 0, false) == MC_OK;
      }
      if (ok) h->tc = plan;
      else pw_tc_free(plan);
    }
  }
  *out = h;
  return MC_OK;
}

int mc_head_set_exact(mc_head* h, int32_t exact) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  h->exact = exact != 0;
  return MC_OK;
}

int mc_head_destroy(mc_head* h) {
  if (!h) return MC_OK;
  DeviceGuard g(h->device);
  for (float* p : h->d_w) if (p) cudaFree(p);
  for (float* p : h->d_b) if (p) cudaFree(p);
  for (float* p : h->d_act) if (p) cudaFree(p);
  void* ptrs[] = {h->d_a, h->d_pb, h->d_in_pad, h->d_feats, h->d_proba, h->d_labels, h->d_ones,
                  h->d_row_loss, h->d_eval_labels, h->d_eval_parts};
  for (void* p : ptrs) if (p) cudaFree(p);
  pw_tc_free(h->tc);
  delete h;
  return MC_OK;
}

}  // extern "C"

static int head_scores_impl(mc_head* h, const float* features_dev, int64_t n, double* proba_dev, int32_t* labels_dev,
                            int32_t topk, int32_t* topk_idx_dev, float* topk_val_dev, const int32_t* y_dev,
                            double* row_loss_dev, void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (n == 0) return MC_OK;
  if (!features_dev || n < 0) return fail(MC_ERR_BAD_ARG, "mc_head_scores: null features");
  const int K = h->dims.back();
  if (topk < 0 || topk > K || (topk > 0 && !topk_idx_dev)) return fail(MC_ERR_BAD_ARG, "mc_head_scores: bad topk");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int L = h->n_layers;
  for (int64_t s = 0; s < n; s += h->chunk) {
    const int64_t m = std::min<int64_t>(h->chunk, n - s);
    const float* x = features_dev + s * h->dims[0];
    if (h->d_in_pad) {
      pad_rows_kernel<<<cdiv(m * h->dims_p[0], 256), 256, 0, st>>>(x, h->dims[0], h->d_in_pad, h->dims_p[0], m);
      MC_CHECK_LAUNCH();
      h->launches++;
      x = h->d_in_pad;
    }
    const bool use_tc = h->tc != nullptr && !h->exact;
    const bool x_is_callers = x == features_dev + s * h->dims[0];
    for (int i = 0; i < L; ++i) {
      const int N = h->dims_p[i + 1], Kd = h->dims_p[i];
      if (use_tc) {
        // layer 0 reads the caller's feature rows [s, s + m): the map covers exactly those rows
        const int64_t map_rows = (i == 0 && x_is_callers) ? m : -1;
        int rc = pw_tc_run(h->tc, i, x, 0, nullptr, nullptr, h->d_act[i], m, 1, st, map_rows);
        if (rc) return rc;
        h->launches++;
        x = h->d_act[i];
        continue;
      }
      dim3 grid(cdiv(m, 64), cdiv(N, 64));
      if (i < L - 1)
        pw_simt_kernel<float, float, ACT_RELU, false, false><<<grid, 256, 0, st>>>(x, h->d_w[i], nullptr, h->d_b[i], nullptr,
                                                                                 nullptr, h->d_act[i], m, N, Kd, 1);
      else
        pw_simt_kernel<float, float, ACT_NONE, false, false><<<grid, 256, 0, st>>>(x, h->d_w[i], nullptr, h->d_b[i], nullptr,
                                                                                 nullptr, h->d_act[i], m, N, Kd, 1);
      MC_CHECK_LAUNCH();
      h->launches++;
      x = h->d_act[i];
    }
    int warps = (int)std::min<int64_t>(8, (96 * 1024) / ((int64_t)K * 4));
    if (warps < 1) return fail(MC_ERR_UNSUPPORTED, "mc_head_scores: too many classes for the row kernel");
    const size_t smem = (size_t)warps * K * sizeof(float);
    // labels / top-k only on a calibrated head through the tensor-core chain: the FAST row kernel (head.cuh)
    const bool fast_rows = use_tc && h->d_a != nullptr && proba_dev == nullptr && y_dev == nullptr && row_loss_dev == nullptr;
    if (smem > 48 * 1024) {
      MC_CUDA(cudaFuncSetAttribute(head_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MC_CUDA(cudaFuncSetAttribute(head_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
#define MC_HEAD_ROWS_ARGS                                                                                              \
  x, h->dims_p[L], K, h->d_a, h->d_pb, proba_dev ? proba_dev + s * K : nullptr, labels_dev ? labels_dev + s : nullptr, topk,   \
      topk_idx_dev ? topk_idx_dev + s * topk : nullptr, topk_val_dev ? topk_val_dev + s * topk : nullptr, m,                  \
      y_dev ? y_dev + s : nullptr, row_loss_dev ? row_loss_dev + s : nullptr
    if (fast_rows) head_rows_kernel<true><<<cdiv(m, warps), warps * 32, smem, st>>>(MC_HEAD_ROWS_ARGS);
    else head_rows_kernel<false><<<cdiv(m, warps), warps * 32, smem, st>>>(MC_HEAD_ROWS_ARGS);
#undef MC_HEAD_ROWS_ARGS
    MC_CHECK_LAUNCH();
    h->launches++;
  }
  return MC_OK;
}

extern "C" {

int mc_head_scores(mc_head* h, const float* features_dev, int64_t n, double* proba_dev, int32_t* labels_dev,
                   int32_t topk, int32_t* topk_idx_dev, float* topk_val_dev, void* stream) {
  return head_scores_impl(h, features_dev, n, proba_dev, labels_dev, topk, topk_idx_dev, topk_val_dev, nullptr, nullptr,
                          stream);
}

int mc_head_evaluate(mc_head* h, const float* features_dev, const int32_t* y_dev, int64_t n, int64_t* n_correct,
                     double* loss_sum, void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (!n_correct || !loss_sum) return fail(MC_ERR_BAD_ARG, "mc_head_evaluate: null outputs");
  *n_correct = 0;
  *loss_sum = 0.0;
  if (n == 0) return MC_OK;
  if (!features_dev || !y_dev || n < 0) return fail(MC_ERR_BAD_ARG, "mc_head_evaluate: null features / targets");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = grow(&h->d_row_loss, &h->cap_row_loss, n))) return rc;
  if ((rc = grow(&h->d_eval_labels, &h->cap_eval_labels, n))) return rc;
  if (!h->d_eval_parts) MC_CUDA(cudaMalloc(&h->d_eval_parts, (2 * EVAL_PARTS + 2) * sizeof(double)));
  if ((rc = head_scores_impl(h, features_dev, n, nullptr, h->d_eval_labels, 0, nullptr, nullptr, y_dev, h->d_row_loss,
                             stream)))
    return rc;
  double* part_loss = h->d_eval_parts;
  long long* part_hits = reinterpret_cast<long long*>(h->d_eval_parts + EVAL_PARTS);
  double* res = h->d_eval_parts + 2 * EVAL_PARTS;
  eval_partial_kernel<<<EVAL_PARTS, 256, 0, st>>>(h->d_row_loss, h->d_eval_labels, y_dev, n, part_loss, part_hits);
  MC_CHECK_LAUNCH();
  eval_final_kernel<<<1, 32, 0, st>>>(part_loss, part_hits, EVAL_PARTS, res, reinterpret_cast<long long*>(res + 1));
  MC_CHECK_LAUNCH();
  h->launches += 2;
  double out[2];
  MC_CUDA(cudaMemcpyAsync(out, res, sizeof(out), cudaMemcpyDeviceToHost, st));
  MC_CUDA(cudaStreamSynchronize(st));
  *loss_sum = out[0];
  long long hits;
  memcpy(&hits, &out[1], sizeof(hits));
  *n_correct = hits;
  return MC_OK;
}

int mc_head_scores_host(mc_head* h, const float* features_host, int64_t n, double* proba_host, int32_t* labels_host,
                        void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (n == 0) return MC_OK;
  if (!features_host || n < 0) return fail(MC_ERR_BAD_ARG, "mc_head_scores_host: null features");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int K = h->dims.back(), D = h->dims[0];
  int rc;
  // walk in chunks so a 10 M-row call does not need a 40 GB proba buffer
  const int64_t step = h->chunk;
  if ((rc = grow(&h->d_feats, &h->cap_feats, step * D))) return rc;
  if (proba_host && (rc = grow(&h->d_proba, &h->cap_proba, step * K))) return rc;
  if (labels_host && (rc = grow(&h->d_labels, &h->cap_labels, step))) return rc;
  for (int64_t s = 0; s < n; s += step) {
    const int64_t m = std::min<int64_t>(step, n - s);
    MC_CUDA(cudaMemcpyAsync(h->d_feats, features_host + s * D, m * D * sizeof(float), cudaMemcpyHostToDevice, st));
    if ((rc = mc_head_scores(h, h->d_feats, m, proba_host ? h->d_proba : nullptr, labels_host ? h->d_labels : nullptr, 0,
                             nullptr, nullptr, st)))
      return rc;
    if (proba_host)
      MC_CUDA(cudaMemcpyAsync(proba_host + s * K, h->d_proba, m * K * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (labels_host)
      MC_CUDA(cudaMemcpyAsync(labels_host + s, h->d_labels, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaStreamSynchronize(st));
  }
  return MC_OK;
}

int64_t mc_head_launches(const mc_head* h) { return h ? h->launches : 0; }

}  // extern "C"

#include "host_pipe.inl"
#include "jpeg_exact.inl"
#include "jpeg_api.inl"
#include "mlp_api.inl"
#include "calib_api.inl"
