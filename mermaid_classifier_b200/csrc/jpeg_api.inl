// mc_jpeg_*: image decode feeding K1.  The reference loads every image with
// spacer.storage.load_image (call site mermaid_classifier/pyspacer/annotation.py:235; inside spacer.tasks.extract_features,
// scripts/build_feature_bucket.py:775): PIL decodes the JPEG on the CPU, converts to RGB, and the whole decoded image
// (36 MB for 12 MP) is then copied around.  Here the compressed bytes go to nvJPEG (resolved at run time from the CUDA
// toolkit's libnvjpeg, so the library itself carries no link dependency): Huffman decoding on the calling host thread,
// inverse DCT / upsampling / colour conversion on the GPU, RGB8 interleaved straight into a device buffer the crop/stem
// kernel reads -- the decoded image never crosses PCIe.  Decoder handles are per thread: run one per worker of a pool.
#include <dlfcn.h>
#include <nvjpeg.h>

namespace {

typedef nvjpegStatus_t (*nvj_create_simple_fn)(nvjpegHandle_t*);
typedef nvjpegStatus_t (*nvj_destroy_fn)(nvjpegHandle_t);
typedef nvjpegStatus_t (*nvj_state_create_fn)(nvjpegHandle_t, nvjpegJpegState_t*);
typedef nvjpegStatus_t (*nvj_state_destroy_fn)(nvjpegJpegState_t);
typedef nvjpegStatus_t (*nvj_get_info_fn)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*);
typedef nvjpegStatus_t (*nvj_decode_fn)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t,
                                        nvjpegImage_t*, cudaStream_t);

struct NvjpegApi {
  void* lib = nullptr;
  nvj_create_simple_fn create = nullptr;
  nvj_destroy_fn destroy = nullptr;
  nvj_state_create_fn state_create = nullptr;
  nvj_state_destroy_fn state_destroy = nullptr;
  nvj_get_info_fn get_info = nullptr;
  nvj_decode_fn decode = nullptr;
};

NvjpegApi* nvjpeg_api() {
  static NvjpegApi api;
  static std::atomic<int> state{0};   // 0 untried, 1 ready, 2 unavailable
  if (state.load() == 1) return &api;
  if (state.load() == 2) return nullptr;
  static std::atomic_flag busy = ATOMIC_FLAG_INIT;
  while (busy.test_and_set()) {
  }
  if (state.load() == 0) {
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    bool ok = api.lib != nullptr;
    if (ok) {
      api.create = (nvj_create_simple_fn)dlsym(api.lib, "nvjpegCreateSimple");
      api.destroy = (nvj_destroy_fn)dlsym(api.lib, "nvjpegDestroy");
      api.state_create = (nvj_state_create_fn)dlsym(api.lib, "nvjpegJpegStateCreate");
      api.state_destroy = (nvj_state_destroy_fn)dlsym(api.lib, "nvjpegJpegStateDestroy");
      api.get_info = (nvj_get_info_fn)dlsym(api.lib, "nvjpegGetImageInfo");
      api.decode = (nvj_decode_fn)dlsym(api.lib, "nvjpegDecode");
      ok = api.create && api.destroy && api.state_create && api.state_destroy && api.get_info && api.decode;
    }
    state.store(ok ? 1 : 2);
  }
  busy.clear();
  return state.load() == 1 ? &api : nullptr;
}

}  // namespace

struct mc_jpeg {
  int device = 0;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  JpxScratch jpx;   // mc_jpeg_decode_exact
};

extern "C" {

int mc_jpeg_create(int32_t device, mc_jpeg** out) {
  if (!out) return fail(MC_ERR_BAD_ARG, "mc_jpeg_create: null");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MC_ERR_CUDA, "no CUDA device: libmermaid_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(MC_ERR_BAD_ARG, "mc_jpeg_create: bad device index");
  NvjpegApi* api = nvjpeg_api();
  if (!api) return fail(MC_ERR_UNSUPPORTED, "libnvjpeg.so.12 could not be loaded (CUDA toolkit library)");
  DeviceGuard g(device);
  mc_jpeg* d = new mc_jpeg();
  d->device = device;
  nvjpegStatus_t rc = api->create(&d->handle);
  if (rc == NVJPEG_STATUS_SUCCESS) rc = api->state_create(d->handle, &d->state);
  if (rc != NVJPEG_STATUS_SUCCESS) {
    if (d->handle) api->destroy(d->handle);
    delete d;
    return fail(MC_ERR_CUDA, "nvjpeg handle creation failed with status " + std::to_string((int)rc));
  }
  *out = d;
  return MC_OK;
}

int mc_jpeg_destroy(mc_jpeg* d) {
  if (!d) return MC_OK;
  NvjpegApi* api = nvjpeg_api();
  DeviceGuard g(d->device);
  if (api) {
    if (d->state) api->state_destroy(d->state);
    if (d->handle) api->destroy(d->handle);
  }
  jpx_free(&d->jpx);
  delete d;
  return MC_OK;
}

int mc_jpeg_info(mc_jpeg* d, const uint8_t* data, int64_t len, int32_t* height, int32_t* width, int32_t* components) {
  if (!d || !data || len <= 0 || !height || !width) return fail(MC_ERR_BAD_ARG, "mc_jpeg_info: null argument");
  NvjpegApi* api = nvjpeg_api();
  int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  const nvjpegStatus_t rc = api->get_info(d->handle, data, (size_t)len, &nc, &ss, ws, hs);
  if (rc != NVJPEG_STATUS_SUCCESS) return fail(MC_ERR_BAD_ARG, "not a decodable JPEG stream (nvjpeg status " + std::to_string((int)rc) + ")");
  *height = hs[0];
  *width = ws[0];
  if (components) *components = nc;
  return MC_OK;
}

int mc_jpeg_decode(mc_jpeg* d, const uint8_t* data, int64_t len, uint8_t* rgb_dev, int64_t row_pitch, int32_t height,
                   int32_t width, void* stream) {
  if (!d || !data || len <= 0 || !rgb_dev) return fail(MC_ERR_BAD_ARG, "mc_jpeg_decode: null argument");
  int32_t h = 0, w = 0, nc = 0;
  if (int rc = mc_jpeg_info(d, data, len, &h, &w, &nc)) return rc;
  if (h != height || w != width || row_pitch < (int64_t)width * 3)
    return fail(MC_ERR_BAD_ARG, "mc_jpeg_decode: destination is " + std::to_string(height) + " x " + std::to_string(width) +
                                    ", the stream holds " + std::to_string(h) + " x " + std::to_string(w));
  if (nc != 1 && nc != 3) return fail(MC_ERR_UNSUPPORTED, "mc_jpeg_decode: " + std::to_string(nc) + "-component JPEG (CMYK / YCCK) is not supported");
  NvjpegApi* api = nvjpeg_api();
  DeviceGuard g(d->device);
  nvjpegImage_t dst;
  memset(&dst, 0, sizeof(dst));
  dst.channel[0] = rgb_dev;
  dst.pitch[0] = (size_t)row_pitch;
  // NVJPEG_OUTPUT_RGBI: interleaved RGB; a grayscale stream is expanded to three equal channels (PIL's convert("RGB"))
  const nvjpegStatus_t rc = api->decode(d->handle, d->state, data, (size_t)len, NVJPEG_OUTPUT_RGBI, &dst, (cudaStream_t)stream);
  if (rc != NVJPEG_STATUS_SUCCESS) return fail(MC_ERR_CUDA, "nvjpegDecode failed with status " + std::to_string((int)rc));
  return MC_OK;
}

int mc_jpeg_decode_exact(mc_jpeg* d, const uint8_t* data, int64_t len, uint8_t* rgb_dev, int64_t row_pitch, int32_t height,
                         int32_t width, void* stream) {
  if (!d || !data || len <= 0 || !rgb_dev) return fail(MC_ERR_BAD_ARG, "mc_jpeg_decode_exact: null argument");
  JpxHeader H;
  if (int rc = jpx_parse(data, (size_t)len, &H)) return rc;
  if ((int64_t)H.height * H.width > 100000000ll)   // pyspacer's MAX_IMAGE_PIXELS (check_extract_inputs): also bounds the scratch
    return fail(MC_ERR_DATA_LIMIT, "mc_jpeg_decode_exact: " + std::to_string(H.width) + " x " + std::to_string(H.height) + " pixels exceed the 1e8 limit");
  if (H.height != height || H.width != width || row_pitch < (int64_t)width * 3)
    return fail(MC_ERR_BAD_ARG, "mc_jpeg_decode_exact: destination is " + std::to_string(height) + " x " + std::to_string(width) +
                                    ", the stream holds " + std::to_string(H.height) + " x " + std::to_string(H.width));
  DeviceGuard g(d->device);
  cudaStream_t st = (cudaStream_t)stream;
  JpxKernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.ncomp = H.ncomp;
  int total = 0;
  int64_t plane_bytes = 0;
  for (int c = 0; c < H.ncomp; ++c) {
    ka.blk_base[c] = total;
    ka.bx[c] = H.comp[c].bx;
    ka.by[c] = H.comp[c].by;
    ka.plane_off[c] = plane_bytes;
    total += H.comp[c].bx * H.comp[c].by;
    plane_bytes += (int64_t)H.comp[c].bx * 8 * H.comp[c].by * 8;
    memcpy(ka.qt[c], H.qt[H.comp[c].tq], sizeof(ka.qt[c]));
  }
  ka.blk_base[H.ncomp] = total;
  const size_t cap = std::min<size_t>((size_t)total * 64, (size_t)len * 4 + (size_t)total) + 64;
  JpxScratch* s = &d->jpx;
  if (int rc = jpx_reserve(s, cap, (size_t)total * 2, (size_t)plane_bytes)) return rc;
  if (s->pending) {   // the previous call's copies out of the pinned buffers must have finished before they are overwritten
    MC_CUDA(cudaEventSynchronize(s->h2d_done));
    s->pending = false;
  }
  size_t n_entries = 0;
  if (H.progressive) {
    s->dense.assign((size_t)total * 64, 0);
    if (int rc = jpx_decode_progressive(data, (size_t)len, H, ka.blk_base, s->dense.data())) return rc;
    if (int rc = jpx_sparsify(s->dense.data(), total, s->h_entries, s->cap_entries, s->h_offsets, &n_entries)) return rc;
  } else if (int rc = jpx_entropy_decode(data, (size_t)len, H, ka.blk_base, s->h_entries, s->cap_entries, s->h_offsets, &n_entries)) {
    return rc;
  }
  MC_CUDA(cudaMemcpyAsync(s->d_entries, s->h_entries, std::max<size_t>(n_entries, 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  MC_CUDA(cudaMemcpyAsync(s->d_offsets, s->h_offsets, (size_t)total * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  MC_CUDA(cudaEventRecord(s->h2d_done, st));
  s->pending = true;
  jpx_idct_kernel<<<cdiv(total, JPX_IDCT_THREADS), JPX_IDCT_THREADS, 0, st>>>(s->d_entries, s->d_offsets, s->d_planes, ka);
  MC_CHECK_LAUNCH();
  jpx_color_kernel<<<dim3(cdiv(width, 256), height), 256, 0, st>>>(s->d_planes, ka, width, height, H.hmax, H.vmax, rgb_dev, row_pitch);
  MC_CHECK_LAUNCH();
  return MC_OK;
}

// Host-only half of mc_jpeg_decode_exact: the quantised DCT coefficients of a baseline stream, dense, natural order,
// component after component over each component's whole block grid ([by][bx][64] int16).  info = {height, width, components,
// bx0, by0, bx1, by1, bx2, by2, restart interval}.  No CUDA device needed (the CPU test-suite checks the entropy decoder
// against oracle/jpeg.py with it); coef_out == NULL only fills `info`.
int mc_jpeg_coefficients_host(const uint8_t* data, int64_t len, int16_t* coef_out, int64_t capacity_blocks, int32_t* info) {
  if (!data || len <= 0 || !info) return fail(MC_ERR_BAD_ARG, "mc_jpeg_coefficients_host: null argument");
  JpxHeader H;
  if (int rc = jpx_parse(data, (size_t)len, &H)) return rc;
  int blk_base[4] = {0, 0, 0, 0};
  int total = 0;
  for (int c = 0; c < H.ncomp; ++c) {
    blk_base[c] = total;
    total += H.comp[c].bx * H.comp[c].by;
  }
  blk_base[H.ncomp] = total;
  info[0] = H.height; info[1] = H.width; info[2] = H.ncomp;
  for (int c = 0; c < 3; ++c) {
    info[3 + 2 * c] = c < H.ncomp ? H.comp[c].bx : 0;
    info[4 + 2 * c] = c < H.ncomp ? H.comp[c].by : 0;
  }
  info[9] = H.ri;
  if (!coef_out) return MC_OK;
  if (capacity_blocks < total) return fail(MC_ERR_BAD_ARG, "mc_jpeg_coefficients_host: room for " + std::to_string(capacity_blocks) + " blocks, the stream holds " + std::to_string(total));
  if (H.progressive) {
    memset(coef_out, 0, (size_t)total * 64 * sizeof(int16_t));
    return jpx_decode_progressive(data, (size_t)len, H, blk_base, coef_out);
  }
  const size_t cap = std::min<size_t>((size_t)total * 64, (size_t)len * 4 + (size_t)total) + 64;
  std::vector<uint32_t> entries(cap), offsets((size_t)total * 2);
  size_t n_entries = 0;
  if (int rc = jpx_entropy_decode(data, (size_t)len, H, blk_base, entries.data(), cap, offsets.data(), &n_entries)) return rc;
  memset(coef_out, 0, (size_t)total * 64 * sizeof(int16_t));
  for (int b = 0; b < total; ++b)
    for (uint32_t e = offsets[2 * b]; e < offsets[2 * b + 1]; ++e) coef_out[(size_t)b * 64 + (entries[e] >> 16)] = (int16_t)(entries[e] & 0xFFFFu);
  return MC_OK;
}

}  // extern "C"
