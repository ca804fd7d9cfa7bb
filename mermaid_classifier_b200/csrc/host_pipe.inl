// mc_extract_images_host: the bulk reference-facing call with HOST buffers.
//
// The reference walks images one at a time -- load, crop on the CPU, batches of 10 patches to the device, features back
// (scripts/build_feature_bucket.py:749-788; extract + classify of one image: pyspacer/annotation.py:235-251).  Here a
// whole list of decoded images goes through a three-slot pipeline owned by the extractor handle:
//
//   h2d stream      per group: the group's image / point tables, then ONE cudaMemcpyAsync per image into the slot's arena --
//                   or, when an image's points need under 60 % of its pixels, one 2-D copy per point of just the window
//                   that point's patch reads (pageable sources are first copied into the slot's pinned staging buffer by
//                   the calling thread)
//   compute stream  (the caller's) waits for the slot's copies, runs the backbone over the group's points in sub-batches
//                   of max_batch patches, then the head when one is given
//   d2h stream      features / labels of the group back into the caller's arrays
//
// A group is a run of consecutive images whose points fill at most one sub-batch, so the copy of group g+1 overlaps
// the compute of group g and the read-back of group g-1.  Events order the slot reuse; the only host waits are the
// pinned-staging reuse (pageable sources) and the final drain.
#pragma once

constexpr int MC_CROP_HALF = MC_CROP_SIZE / 2;

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
  int64_t copies = 0;                                // H2D copy operations of the last call
  double issue_ms = 0;                               // host time spent issuing them (MC_PIPE_DEBUG prints both)
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

// upload window of a patch centre along one axis: [win_lo, win_hi)
inline int win_lo(int centre) { return std::max(0, centre - MC_CROP_HALF); }
inline int win_hi(int centre, int size) { return std::min(size, centre + MC_CROP_HALF + (centre == 0 ? 1 : 0)); }

// Upload plan of one image: every point's own window, with windows MERGED into their bounding box while that adds few bytes
// -- on annotation-dense images many windows overlap, and every 2-D copy is a driver call and a DMA descriptor chain of
// ~224 short rows (measured: 35-55 % fewer copies for the same bytes on the bench's point sets; the end-to-end rate is bound
// by the strided DMA itself, ~20 GB/s, and did not move).  A merged box contains each member's own window, touches an image
// border exactly where a member's window was clipped there, and is otherwise never crossed by a member's patch: the reflect
// argument of a single window carries over.  Greedy, in point order: a point joins the box whose growth is smallest if that
// growth exceeds the point's own window by at most MC_MERGE_SLACK bytes.
struct UploadWin {
  int r0, c0, r1, c1;   // rows [r0, r1) x columns [c0, c1)
  int64_t bytes() const { return (int64_t)(r1 - r0) * (c1 - c0) * 3; }
};
constexpr int64_t MC_MERGE_SLACK = 128 * 1024;

void plan_uploads(int height, int width, const mc_point* pts, int64_t n, std::vector<UploadWin>& wins, int32_t* idx) {
  wins.clear();
  for (int64_t t = 0; t < n; ++t) {
    const UploadWin own{win_lo(pts[t].row), win_lo(pts[t].col), win_hi(pts[t].row, height), win_hi(pts[t].col, width)};
    int best = -1;
    int64_t best_growth = own.bytes() + MC_MERGE_SLACK;
    for (size_t k = 0; k < wins.size(); ++k) {
      const UploadWin& w = wins[k];
      const UploadWin u{std::min(w.r0, own.r0), std::min(w.c0, own.c0), std::max(w.r1, own.r1), std::max(w.c1, own.c1)};
      const int64_t growth = u.bytes() - w.bytes();
      if (growth <= best_growth) {
        best_growth = growth;
        best = (int)k;
      }
    }
    if (best < 0) {
      idx[t] = (int32_t)wins.size();
      wins.push_back(own);
    } else {
      UploadWin& w = wins[(size_t)best];
      w = UploadWin{std::min(w.r0, own.r0), std::min(w.c0, own.c0), std::max(w.r1, own.r1), std::max(w.c1, own.c1)};
      idx[t] = best;
    }
  }
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
  P->h2d_bytes = P->d2h_bytes = P->groups = P->copies = 0;
  P->issue_ms = 0;

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
    // Virtual images of the group.  An image whose points need only a fraction of its pixels is not uploaded whole: each of its
    // points gets its own WINDOW -- rows [row - 112, row + 112) x columns [col - 112, col + 112) clipped to the image -- as a
    // small image of its own, the point re-expressed in window coordinates.  The crop reflects at the IMAGE border
    // (crop_patches); a window border is either that same border (clipped side) or is never crossed: a patch
    // covers [centre - 112, centre + 111], and an index reflected at border 0 is at most 112 - centre <= centre + 111 unless
    // the centre lies ON the border (index 112: one more row / column, `win_hi`); reflected at the far border it is at least
    // 2 (H - 1) - (centre + 111) >= centre - 112.  So the patch bytes are identical.
    // C3 shape (50 points on 4000 x 3000): 7.5 MB instead of 36 MB per image over PCIe; C2 (100 points): 15 MB.
    struct VImg {
      int src, r0, c0, h, w;
      int64_t off, dpitch;
    };
    std::vector<VImg> vims;
    std::vector<UploadWin> plan;
    std::vector<int32_t> pt_v((size_t)ng);
    int64_t arena = 0, stage = 0;
    for (int64_t q = p0; q < p1;) {
      const int src = points[q].image;
      int64_t e = q;
      while (e < p1 && points[e].image == src) ++e;
      const mc_image& im = images[src];
      const int64_t row_b = (int64_t)im.width * 3;
      int64_t win_bytes = 0;
      if (h->sparse_h2d) {
        plan_uploads(im.height, im.width, points + q, e - q, plan, pt_v.data() + (q - p0));
        for (const UploadWin& w : plan) win_bytes += w.bytes();
      }
      if (h->sparse_h2d && win_bytes * 10 <= row_b * im.height * 6) {
        const int32_t base = (int32_t)vims.size();
        for (int64_t t = q; t < e; ++t) pt_v[(size_t)(t - p0)] += base;
        for (const UploadWin& w : plan) {
          VImg v;
          v.src = src;
          v.r0 = w.r0;
          v.c0 = w.c0;
          v.h = w.r1 - w.r0;
          v.w = w.c1 - w.c0;
          v.dpitch = ((int64_t)v.w * 3 + 15) / 16 * 16;
          v.off = arena;
          arena += (v.dpitch * v.h + 255) / 256 * 256;
          stage += ((int64_t)v.w * 3 * v.h + 255) / 256 * 256;
          vims.push_back(v);
        }
      } else {
        VImg v;
        v.src = src;
        v.r0 = v.c0 = 0;
        v.h = im.height;
        v.w = im.width;
        v.dpitch = im.row_pitch == row_b ? im.row_pitch : (row_b + 255) / 256 * 256;
        v.off = arena;
        arena += (v.dpitch * v.h + 255) / 256 * 256;
        stage += (row_b * v.h + 255) / 256 * 256;
        for (int64_t t = q; t < e; ++t) pt_v[(size_t)(t - p0)] = (int32_t)vims.size();
        vims.push_back(v);
      }
      q = e;
    }
    const int64_t im_tab = (int64_t)(vims.size() * sizeof(mc_image) + 15) / 16 * 16;
    const int64_t tab_bytes = im_tab + ng * (int64_t)sizeof(mc_point);
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
    mc_point* t_pt = reinterpret_cast<mc_point*>(s.tab_host + im_tab);
    for (size_t k = 0; k < vims.size(); ++k) t_im[k] = mc_image{s.d_arena + vims[k].off, vims[k].h, vims[k].w, vims[k].dpitch};
    for (int64_t q = 0; q < ng; ++q) {
      const VImg& v = vims[(size_t)pt_v[(size_t)q]];
      t_pt[q] = mc_point{pt_v[(size_t)q], points[p0 + q].row - v.r0, points[p0 + q].col - v.c0};
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
    const auto t_issue0 = std::chrono::steady_clock::now();
    int pinned_src = -1;
    bool src_is_pinned = false;
    for (size_t k = 0; k < vims.size(); ++k) {
      const VImg& v = vims[k];
      const mc_image& im = images[v.src];
      const int64_t row = (int64_t)v.w * 3;
      const uint8_t* src = im.data + (int64_t)v.r0 * im.row_pitch + (int64_t)v.c0 * 3;
      int64_t spitch = im.row_pitch;
      if (v.src != pinned_src) {
        src_is_pinned = is_pinned_host(im.data);
        pinned_src = v.src;
      }
      if (!src_is_pinned) {
        // pageable source: rows into the slot's pinned staging buffer (packed), DMA from there
        if ((rc = pinned_grow(&s.pinned, &s.cap_pinned, stage))) {
          restore();
          return rc;
        }
        uint8_t* dst = s.pinned + stage_off;
        if (spitch == row) memcpy(dst, src, (size_t)(row * v.h));
        else
          for (int y = 0; y < v.h; ++y) memcpy(dst + (int64_t)y * row, src + (int64_t)y * spitch, (size_t)row);
        src = dst;
        spitch = row;
        stage_off += (row * v.h + 255) / 256 * 256;
      }
      if (spitch == row && v.dpitch == row)
        MC_CUDA(cudaMemcpyAsync(s.d_arena + v.off, src, (size_t)(row * v.h), cudaMemcpyHostToDevice, P->h2d));
      else
        MC_CUDA(cudaMemcpy2DAsync(s.d_arena + v.off, (size_t)v.dpitch, src, (size_t)spitch, (size_t)row, v.h,
                                  cudaMemcpyHostToDevice, P->h2d));
      P->h2d_bytes += row * v.h;
    }
    P->copies += (int64_t)vims.size();
    P->issue_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_issue0).count();
    P->h2d_bytes += tab_bytes;
    MC_CUDA(cudaEventRecord(s.copied, P->h2d));
    // ---- compute stream -----------------------------------------------------------------------------------------------
    MC_CUDA(cudaStreamWaitEvent(st, s.copied, 0));
    if (s.used) MC_CUDA(cudaStreamWaitEvent(st, s.drained, 0));   // the slot's previous features have been read back
    h->d_images = reinterpret_cast<mc_image*>(s.tab_dev);
    mc_point* const d_pts = reinterpret_cast<mc_point*>(s.tab_dev + im_tab);
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
  if (getenv("MC_PIPE_DEBUG"))
    fprintf(stderr, "[mc pipe] %lld points, %lld groups, %lld H2D copies issued in %.2f ms (%.2f us each), %.1f MB\n", (long long)n,
            (long long)P->groups, (long long)P->copies, P->issue_ms, 1e3 * P->issue_ms / (double)std::max<int64_t>(1, P->copies), P->h2d_bytes / 1e6);
  return prof_collect(h, st);
}

extern "C" int mc_upload_window(int32_t height, int32_t width, int32_t row, int32_t col, int32_t* r0, int32_t* c0, int32_t* hh,
                                int32_t* ww) {
  if (!r0 || !c0 || !hh || !ww) return fail(MC_ERR_BAD_ARG, "mc_upload_window: null output");
  if (height < 1 || width < 1 || row < 0 || row >= height || col < 0 || col >= width)
    return fail(MC_ERR_POINT_BOUNDS, "mc_upload_window: point outside the image");
  *r0 = win_lo(row);
  *c0 = win_lo(col);
  *hh = win_hi(row, height) - *r0;
  *ww = win_hi(col, width) - *c0;
  return MC_OK;
}

extern "C" int mc_plan_uploads(int32_t height, int32_t width, const int32_t* rowcols, int64_t n, int32_t* windows_out,
                               int32_t* index_out, int64_t* n_windows) {
  if (!rowcols || !windows_out || !index_out || !n_windows || n < 0) return fail(MC_ERR_BAD_ARG, "mc_plan_uploads: null argument");
  if (height < 1 || width < 1) return fail(MC_ERR_BAD_ARG, "mc_plan_uploads: empty image");
  std::vector<mc_point> pts((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const int r = rowcols[2 * i], c = rowcols[2 * i + 1];
    if (r < 0 || r >= height || c < 0 || c >= width) return fail(MC_ERR_POINT_BOUNDS, "mc_plan_uploads: point outside the image");
    pts[(size_t)i] = mc_point{0, r, c};
  }
  std::vector<UploadWin> wins;
  plan_uploads(height, width, pts.data(), n, wins, index_out);
  for (size_t k = 0; k < wins.size(); ++k) {
    windows_out[4 * k + 0] = wins[k].r0;
    windows_out[4 * k + 1] = wins[k].c0;
    windows_out[4 * k + 2] = wins[k].r1 - wins[k].r0;
    windows_out[4 * k + 3] = wins[k].c1 - wins[k].c0;
  }
  *n_windows = (int64_t)wins.size();
  return MC_OK;
}

extern "C" int mc_extractor_pipe_stats(const mc_extractor* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* groups) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  const HostPipe* p = h->pipe;
  if (h2d_bytes) *h2d_bytes = p ? p->h2d_bytes : 0;
  if (d2h_bytes) *d2h_bytes = p ? p->d2h_bytes : 0;
  if (groups) *groups = p ? p->groups : 0;
  return MC_OK;
}
