// mc_extract_images_host: the bulk reference-facing call with HOST buffers.
//
// The reference walks images one at a time -- load, crop on the CPU, batches of 10 patches to the device, features back
// (scripts/build_feature_bucket.py:749-788; extract + classify of one image: pyspacer/annotation.py:235-251).  Here a
// whole list of decoded images goes through a three-slot pipeline owned by the extractor handle:
//
//   h2d stream      per group: the group's image / point tables, then ONE cudaMemcpyAsync per image into the slot's arena
//                   (pageable sources are first copied into the slot's pinned staging buffer by the calling thread)
//   compute stream  (the caller's) waits for the slot's copies, runs the backbone over the group's points in sub-batches
//                   of max_batch patches, then the head when one is given
//   d2h stream      features / labels of the group back into the caller's arrays
//
// A group is a run of consecutive images whose points fill at most one sub-batch, so the copy of group g+1 overlaps
// the compute of group g and the read-back of group g-1.  Events order the slot reuse; the only host waits are the
// pinned-staging reuse (pageable sources) and the final drain.
#pragma once

struct HostPipe {
  static constexpr int SLOTS = 3;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  struct Slot {
    uint8_t* d_arena = nullptr;
    int64_t cap_arena = 0;
    uint8_t* pinned = nullptr;     // staging for pageable sources
    int64_t cap_pinned = 0;
    uint8_t* tab_host = nullptr;   // pinned: [mc_image table | mc_point table] of the group
    int64_t cap_tab_host = 0;
    uint8_t* tab_dev = nullptr;
    int64_t cap_tab = 0;
    float* d_feats = nullptr;
    int64_t cap_feats = 0;
    int32_t* d_labels = nullptr;
    int64_t cap_labels = 0;
    cudaEvent_t copied = nullptr, freed = nullptr, scored = nullptr, drained = nullptr;
    bool used = false;
  } slot[SLOTS];
  int64_t h2d_bytes = 0, d2h_bytes = 0, groups = 0;   // totals of the last call (bench accounting)
};

namespace {

void host_pipe_free(HostPipe* p) {
  if (!p) return;
  for (auto& s : p->slot) {
    if (s.d_arena) cudaFree(s.d_arena);
    if (s.pinned) cudaFreeHost(s.pinned);
    if (s.tab_host) cudaFreeHost(s.tab_host);
    if (s.tab_dev) cudaFree(s.tab_dev);
    if (s.d_feats) cudaFree(s.d_feats);
    if (s.d_labels) cudaFree(s.d_labels);
    for (cudaEvent_t e : {s.copied, s.freed, s.scored, s.drained})
      if (e) cudaEventDestroy(e);
  }
  if (p->h2d) cudaStreamDestroy(p->h2d);
  if (p->d2h) cudaStreamDestroy(p->d2h);
  delete p;
}

int host_pipe_get(mc_extractor* h, HostPipe** out) {
  if (!h->pipe) {
    HostPipe* p = new HostPipe();
    h->pipe = p;
    MC_CUDA(cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking));
    MC_CUDA(cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking));
    for (auto& s : p->slot) {
      MC_CUDA(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
      MC_CUDA(cudaEventCreateWithFlags(&s.freed, cudaEventDisableTiming));
      MC_CUDA(cudaEventCreateWithFlags(&s.scored, cudaEventDisableTiming));
      MC_CUDA(cudaEventCreateWithFlags(&s.drained, cudaEventDisableTiming));
    }
  }
  *out = h->pipe;
  return MC_OK;
}

// device buffer growth for a slot that may still be in flight: drain its events first
template <typename T>
int slot_grow(HostPipe::Slot& s, T** p, int64_t* cap, int64_t need) {
  if (need <= *cap) return MC_OK;
  if (s.used) {
    MC_CUDA(cudaEventSynchronize(s.freed));
    MC_CUDA(cudaEventSynchronize(s.drained));
  }
  return grow(p, cap, need + need / 4);
}

int pinned_grow(uint8_t** p, int64_t* cap, int64_t need) {
  if (need <= *cap) return MC_OK;
  if (*p) MC_CUDA(cudaFreeHost(*p));
  *p = nullptr;
  *cap = 0;
  const int64_t n = need + need / 4;
  cudaError_t e = cudaHostAlloc((void**)p, (size_t)n, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(MC_ERR_NOMEM, std::string("cudaHostAlloc staging: ") + cudaGetErrorString(e));
  *cap = n;
  return MC_OK;
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace

static int head_scores_impl(mc_head* h, const float* features_dev, int64_t n, double* proba_dev, int32_t* labels_dev,
                            int32_t topk, int32_t* topk_idx_dev, float* topk_val_dev, const int32_t* y_dev,
                            double* row_loss_dev, void* stream);

extern "C" int mc_extract_images_host(mc_extractor* h, mc_head* head, const mc_image* images, int32_t n_images,
                                      const mc_point* points, int64_t n, float* feats_host, int32_t* labels_host,
                                      void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (n == 0) return MC_OK;
  if (!images || !points || n_images < 1 || n < 0 || (!feats_host && !labels_host))
    return fail(MC_ERR_BAD_ARG, "mc_extract_images_host: null argument");
  if (labels_host && !head) return fail(MC_ERR_BAD_ARG, "mc_extract_images_host: labels requested without a head");
  if (head && head->device != h->device) return fail(MC_ERR_BAD_ARG, "mc_extract_images_host: head lives on another device");
  int rc;
  if ((rc = check_images(images, n_images)) || (rc = check_points(images, n_images, points, n))) return rc;
  for (int64_t i = 1; i < n; ++i)
    if (points[i].image < points[i - 1].image)
      return fail(MC_ERR_BAD_ARG, "mc_extract_images_host: points must be grouped by image (non-decreasing image index)");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  HostPipe* P = nullptr;
  if ((rc = host_pipe_get(h, &P))) return rc;
  P->h2d_bytes = P->d2h_bytes = P->groups = 0;

  // the streams of the pipeline start after whatever the caller already queued on `st`
  cudaEvent_t& enter = P->slot[0].scored;   // any idle event will do before the first group
  mc_image* const saved_images = h->d_images;
  mc_point* const saved_points = h->d_points;
  auto restore = [&]() {
    h->d_images = saved_images;
    h->d_points = saved_points;
  };

  int64_t p0 = 0;   // first point of the current group
  int gi = 0;
  bool first = true;
  while (p0 < n) {
    // ---- group: consecutive images while the point count stays within one sub-batch --------------------------------
    int64_t p1 = p0;
    while (p1 < n) {
      const int im = points[p1].image;
      int64_t q = p1;
      while (q < n && points[q].image == im) ++q;
      if (p1 > p0 && q - p0 > h->max_batch) break;
      p1 = q;
    }
    const int64_t ng = p1 - p0;
    HostPipe::Slot& s = P->slot[gi % HostPipe::SLOTS];
    // images of the group that carry points, their arena offsets
    std::vector<int> ims;
    for (int64_t q = p0; q < p1; ++q)
      if (ims.empty() || ims.back() != points[q].image) ims.push_back(points[q].image);
    std::vector<int64_t> off(ims.size());
    std::vector<int64_t> dpitch(ims.size());
    int64_t arena = 0, stage = 0;
    for (size_t k = 0; k < ims.size(); ++k) {
      const mc_image& im = images[ims[k]];
      const bool contiguous = im.row_pitch == (int64_t)im.width * 3;
      dpitch[k] = contiguous ? im.row_pitch : ((int64_t)im.width * 3 + 255) / 256 * 256;
      off[k] = arena;
      arena += (dpitch[k] * im.height + 255) / 256 * 256;
      stage += ((int64_t)im.width * 3 * im.height + 255) / 256 * 256;
    }
    const int64_t tab_bytes = (int64_t)(ims.size() * sizeof(mc_image) + 15) / 16 * 16 + ng * (int64_t)sizeof(mc_point);
    if ((rc = slot_grow(s, &s.d_arena, &s.cap_arena, arena)) || (rc = slot_grow(s, &s.d_feats, &s.cap_feats, ng * MC_FEATURE_DIM)) ||
        (head && (rc = slot_grow(s, &s.d_labels, &s.cap_labels, ng))) || (rc = slot_grow(s, &s.tab_dev, &s.cap_tab, tab_bytes))) {
      restore();
      return rc;
    }
    // the pinned table / staging buffers are rewritten by the host: the slot's previous H2D must have finished
    if (s.used) MC_CUDA(cudaEventSynchronize(s.copied));
    if ((rc = pinned_grow(&s.tab_host, &s.cap_tab_host, tab_bytes))) {
      restore();
      return rc;
    }
    mc_image* t_im = reinterpret_cast<mc_image*>(s.tab_host);
    mc_point* t_pt = reinterpret_cast<mc_point*>(s.tab_host + (ims.size() * sizeof(mc_image) + 15) / 16 * 16);
    for (size_t k = 0; k < ims.size(); ++k) {
      const mc_image& im = images[ims[k]];
      t_im[k] = mc_image{s.d_arena + off[k], im.height, im.width, dpitch[k]};
    }
    {
      size_t k = 0;
      for (int64_t q = 0; q < ng; ++q) {
        while (ims[k] != points[p0 + q].image) ++k;
        t_pt[q] = mc_point{(int32_t)k, points[p0 + q].row, points[p0 + q].col};
      }
    }
    // ---- h2d stream ---------------------------------------------------------------------------------------------------
    if (first) {
      MC_CUDA(cudaEventRecord(enter, st));
      MC_CUDA(cudaStreamWaitEvent(P->h2d, enter, 0));
      first = false;
    }
    if (s.used) MC_CUDA(cudaStreamWaitEvent(P->h2d, s.freed, 0));   // the arena's previous group has been convolved
    MC_CUDA(cudaMemcpyAsync(s.tab_dev, s.tab_host, (size_t)tab_bytes, cudaMemcpyHostToDevice, P->h2d));
    int64_t stage_off = 0;
    for (size_t k = 0; k < ims.size(); ++k) {
      const mc_image& im = images[ims[k]];
      const int64_t row = (int64_t)im.width * 3;
      const uint8_t* src = im.data;
      int64_t spitch = im.row_pitch;
      if (!is_pinned_host(im.data)) {
        // pageable source: rows into the slot's pinned staging buffer (packed), DMA from there
        if ((rc = pinned_grow(&s.pinned, &s.cap_pinned, stage))) {
          restore();
          return rc;
        }
        uint8_t* dst = s.pinned + stage_off;
        if (im.row_pitch == row) memcpy(dst, im.data, (size_t)(row * im.height));
        else
          for (int y = 0; y < im.height; ++y) memcpy(dst + (int64_t)y * row, im.data + (int64_t)y * im.row_pitch, (size_t)row);
        src = dst;
        spitch = row;
        stage_off += (row * im.height + 255) / 256 * 256;
      }
      if (spitch == row && dpitch[k] == row)
        MC_CUDA(cudaMemcpyAsync(s.d_arena + off[k], src, (size_t)(row * im.height), cudaMemcpyHostToDevice, P->h2d));
      else
        MC_CUDA(cudaMemcpy2DAsync(s.d_arena + off[k], (size_t)dpitch[k], src, (size_t)spitch, (size_t)row, im.height,
                                  cudaMemcpyHostToDevice, P->h2d));
      P->h2d_bytes += row * im.height;
    }
    P->h2d_bytes += tab_bytes;
    MC_CUDA(cudaEventRecord(s.copied, P->h2d));
    // ---- compute stream -----------------------------------------------------------------------------------------------
    MC_CUDA(cudaStreamWaitEvent(st, s.copied, 0));
    if (s.used) MC_CUDA(cudaStreamWaitEvent(st, s.drained, 0));   // the slot's previous features have been read back
    h->d_images = reinterpret_cast<mc_image*>(s.tab_dev);
    mc_point* const d_pts = reinterpret_cast<mc_point*>(s.tab_dev + (ims.size() * sizeof(mc_image) + 15) / 16 * 16);
    for (int64_t q = 0; q < ng; q += h->max_batch) {
      const int nb = (int)std::min<int64_t>(h->max_batch, ng - q);
      h->d_points = d_pts + q;
      if ((rc = forward_any(h, nb, s.d_feats + q * MC_FEATURE_DIM, st))) {
        restore();
        return rc;
      }
    }
    MC_CUDA(cudaEventRecord(s.freed, st));
    if (labels_host && (rc = head_scores_impl(head, s.d_feats, ng, nullptr, s.d_labels, 0, nullptr, nullptr, nullptr, nullptr, st))) {
      restore();
      return rc;
    }
    MC_CUDA(cudaEventRecord(s.scored, st));
    // ---- d2h stream ---------------------------------------------------------------------------------------------------
    MC_CUDA(cudaStreamWaitEvent(P->d2h, s.scored, 0));
    if (feats_host) {
      MC_CUDA(cudaMemcpyAsync(feats_host + p0 * MC_FEATURE_DIM, s.d_feats, (size_t)ng * MC_FEATURE_DIM * sizeof(float),
                              cudaMemcpyDeviceToHost, P->d2h));
      P->d2h_bytes += ng * MC_FEATURE_DIM * (int64_t)sizeof(float);
    }
    if (labels_host) {
      MC_CUDA(cudaMemcpyAsync(labels_host + p0, s.d_labels, (size_t)ng * sizeof(int32_t), cudaMemcpyDeviceToHost, P->d2h));
      P->d2h_bytes += ng * (int64_t)sizeof(int32_t);
    }
    MC_CUDA(cudaEventRecord(s.drained, P->d2h));
    s.used = true;
    P->groups++;
    p0 = p1;
    ++gi;
  }
  restore();
  MC_CUDA(cudaStreamSynchronize(P->d2h));
  MC_CUDA(cudaStreamSynchronize(st));
  return prof_collect(h, st);
}

extern "C" int mc_extractor_pipe_stats(const mc_extractor* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* groups) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  const HostPipe* p = h->pipe;
  if (h2d_bytes) *h2d_bytes = p ? p->h2d_bytes : 0;
  if (d2h_bytes) *d2h_bytes = p ? p->d2h_bytes : 0;
  if (groups) *groups = p ? p->groups : 0;
  return MC_OK;
}
