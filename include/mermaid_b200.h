/*
 * mermaid_b200.h -- C ABI of libmermaid_b200.so (sm_100a CUDA, B200).
 *
 * One data-parallel hot path of data-mermaid/mermaid-classifier, rebuilt for B200:
 * point-patch crop -> normalise -> EfficientNet-B0 -> 1280-d features -> MLP/Platt head.
 *
 * Every entry point replaces a Python-level call of the reference (there is no native
 * interface in the reference; the FFI a maintainer would add is the ctypes stub shown
 * in INTEGRATION.md).  Each declaration cites the reference code it stands in for
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  "dev" pointers are CUDA device
 *     pointers on the extractor's device, "host" pointers are ordinary host memory.
 *   - the caller owns every input/output buffer; handles own weights and workspaces.
 *   - every call returns an int status (MC_OK == 0); mc_last_error() gives the
 *     thread-local message of the last failure.
 *   - calls taking `stream` (a cudaStream_t passed as void*, NULL = default stream)
 *     are asynchronous on that stream and never call cudaDeviceSynchronize; calls
 *     with `_host` in the name synchronise the stream before returning because they
 *     hand back host results.
 *   - handles are not thread-safe: one handle per (device, stream).
 */
#ifndef MERMAID_B200_H
#define MERMAID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_ABI_VERSION 1

/* status codes; the Python wrapper maps them to the reference's exception types */
#define MC_OK 0
#define MC_ERR_BAD_ARG 1       /* ValueError                                              */
#define MC_ERR_POINT_BOUNDS 2  /* spacer RowColumnInvalidError (check_extract_inputs)     */
#define MC_ERR_DATA_LIMIT 3    /* spacer DataLimitError (check_extract_inputs)            */
#define MC_ERR_CUDA 4          /* RuntimeError                                            */
#define MC_ERR_UNSUPPORTED 5   /* RuntimeError                                            */
#define MC_ERR_NOMEM 6         /* MemoryError                                             */

/* arithmetic modes of the backbone */
#define MC_MODE_FP32 0 /* fp32 activations, fp32-accurate arithmetic (parity config C1/C2) */
#define MC_MODE_BF16 1 /* bf16 activations + bf16 tensor-core GEMMs, fp32 accumulate (C3)  */

#define MC_FEATURE_DIM 1280
#define MC_CROP_SIZE 224

typedef struct mc_extractor mc_extractor;
typedef struct mc_head mc_head;
typedef struct mc_mlp mc_mlp;

/* A decoded RGB8 image, HWC, rows `row_pitch` bytes apart. */
typedef struct mc_image {
  const uint8_t* data;
  int32_t height;
  int32_t width;
  int64_t row_pitch;
} mc_image;

/* A point annotation: index into the image table + (row, col) of the patch centre. */
typedef struct mc_point {
  int32_t image;
  int32_t row;
  int32_t col;
} mc_point;

int mc_abi_version(void);
const char* mc_last_error(void);

/* ---- synthetic inputs (bench/test data generator; same integer hash as
 *      mermaid_classifier_b200/synth.py::synth_image) ------------------------------- */
int mc_synth_image(uint8_t* img_dev, int32_t height, int32_t width, int64_t row_pitch,
                   uint32_t seed, uint32_t image_id, void* stream);

/* ---- A1: spacer.task_utils.check_extract_inputs (call site:
 *      mermaid_classifier/pyspacer/annotation.py:240).  Host-only validation. -------- */
int mc_check_extract_inputs(int32_t height, int32_t width, const int32_t* rowcols_host,
                            int64_t n_points, int64_t max_pixels, int64_t max_points);

/* ---- A2: spacer crop_patches / crop_simple (invoked through extractor(img, rowcols),
 *      mermaid_classifier/pyspacer/annotation.py:241).  Bit-exact reflect-padded gather:
 *      patches_dev[k][i][j][c] = img[R(row_k-112+i, H)][R(col_k-112+j, W)][c]. ---------- */
int mc_crop_patches(const mc_image* images_host, int32_t n_images, const mc_point* points_host,
                    int64_t n_points, uint8_t* patches_dev /* n x 224 x 224 x 3 */, void* stream);

/* ---- A2': the patch-size != 224 path (model.json "config": {"patch_size": ...}, mermaid_classifier/pyspacer/inference/
 *      export.py:77; north_star's "crop / bilinear-resize"): a crop_size x crop_size window around each point (same
 *      reflect rule, centre offset crop_size / 2; even sizes) resized to 224 x 224 with bilinear interpolation in the
 *      arithmetic of torch.nn.functional.interpolate(mode="bilinear", align_corners=False) and rounded half-to-even to
 *      uint8.  crop_size == 224 is mc_crop_patches.  The patches then go through mc_extract_patches. ------------- */
int mc_crop_resize_patches(const mc_image* images_host, int32_t n_images, const mc_point* points_host,
                           int64_t n_points, int32_t crop_size, uint8_t* patches_dev /* n x 224 x 224 x 3 */,
                           void* stream);

/* ---- A3: torch_extractors.transformation() (scripts/build_feature_bucket.py:420-431):
 *      ToTensor + Normalize, HWC u8 -> CHW fp32.  Exposed for the parity gate only; the
 *      extraction path fuses it into the stem kernel. -------------------------------- */
int mc_normalize_patches(const uint8_t* patches_dev, int64_t n, float* out_dev /* n x 3 x 224 x 224 */,
                         void* stream);

/* ---- EfficientNetExtractor (spacer.extractors; constructed at
 *      mermaid_classifier/pyspacer/annotation.py:236-238 and
 *      scripts/build_feature_bucket.py:854-859).
 *      `params_host`: BN-folded parameters packed by
 *      mermaid_classifier_b200/weights.py::pack_backbone (fp32, canonical order);
 *      `n_params` must equal mc_backbone_param_count(). ------------------------------ */
int64_t mc_backbone_param_count(void);
int mc_extractor_create(const float* params_host, int64_t n_params, int32_t mode, int32_t device,
                        int32_t max_batch, mc_extractor** out);
int mc_extractor_destroy(mc_extractor* h);
int mc_extractor_mode(const mc_extractor* h);
int64_t mc_extractor_launches(const mc_extractor* h); /* kernels launched so far by this handle */

/* ---- A2+A3+A4: extractor.__call__(image, rowcols) (annotation.py:241) for a table of
 *      device-resident images: crop + normalise + EfficientNet-B0 extract_features.
 *      feats_dev is n_points x 1280 fp32, row k = point k.  Any n_points; the library
 *      walks it in sub-batches of at most max_batch patches. ------------------------- */
int mc_extract_points(mc_extractor* h, const mc_image* images_host, int32_t n_images,
                      const mc_point* points_host, int64_t n_points, float* feats_dev, void* stream);

/* ---- A3+A4: TorchExtractor.patches_to_features(patch_list)
 *      (scripts/build_feature_bucket.py:415-446) for pre-cropped 224x224x3 u8 patches. */
int mc_extract_patches(mc_extractor* h, const uint8_t* patches_dev, int64_t n, float* feats_dev,
                       void* stream);

/* ---- The reference-facing call with HOST buffers: one image + its rowcols in, features
 *      out (spacer.tasks.extract_features minus storage I/O,
 *      scripts/build_feature_bucket.py:775).  H2D of the image, D2H of the features and a
 *      stream synchronise happen inside.  rowcols_host is n x 2 int32 (row, col). ------- */
int mc_extract_image_host(mc_extractor* h, const uint8_t* img_host, int32_t height, int32_t width,
                          int64_t row_pitch, const int32_t* rowcols_host, int64_t n_points,
                          float* feats_host, void* stream);

/* ---- The bulk reference-facing call with HOST buffers: the per-image loop of process_source
 *      (scripts/build_feature_bucket.py:749-788: load image, extractor(img, rowcols), store) and, with a
 *      head, the extract + classify sequence of AnnotationRun (mermaid_classifier/pyspacer/annotation.py:235-251)
 *      for a whole list of decoded images in one call.  images_host[i].data are HOST pointers (pinned memory is
 *      DMA'd directly, pageable memory goes through the handle's pinned staging buffers); points_host must be
 *      grouped by image (non-decreasing .image).  Inside: a three-slot pipeline on the handle's own copy streams --
 *      one cudaMemcpyAsync per image (or, for sparsely annotated images, one 2-D copy per point of the window its
 *      patch reads: mc_upload_window) into a device arena, the backbone (+ head) on `stream`, features / labels
 *      back -- so the copy of one group of images overlaps the compute of the previous one.  feats_host
 *      (n_points x 1280 fp32) and labels_host (n_points int32, needs `head`) may each be NULL.  Synchronises
 *      before returning. ------------------------------------------------------------------------------------ */
int mc_extract_images_host(mc_extractor* h, mc_head* head, const mc_image* images_host, int32_t n_images,
                           const mc_point* points_host, int64_t n_points, float* feats_host,
                           int32_t* labels_host, void* stream);
/* Bytes copied host->device / device->host and image groups of the last mc_extract_images_host call. */
int mc_extractor_pipe_stats(const mc_extractor* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* groups);
/* Host-only.  The window of a height x width image that mc_extract_images_host uploads for a point when the image's points
 * need under 60 % of its pixels: rows [*r0, *r0 + *h) x columns [*c0, *c0 + *w) hold every pixel crop_patches reads for the
 * 224 x 224 patch centred on (row, col) -- reflection at the image border included -- so cropping the window around
 * (row - *r0, col - *c0) gives the patch of the whole image byte for byte (tests/test_cabi.py checks this against the oracle
 * for every centre of small images).  Replaces nothing in the reference (it pads and slices whole images on the host,
 * SURVEY 8a A2); it is the data-movement rule of the host pipeline, exported so that it can be tested without a GPU. */
int mc_upload_window(int32_t height, int32_t width, int32_t row, int32_t col, int32_t* r0, int32_t* c0, int32_t* h,
                     int32_t* w);
/* Host-only.  The upload plan of one image as mc_extract_images_host builds it: the points' windows (mc_upload_window), merged
 * into bounding boxes while a merge costs fewer bytes than another copy costs time.  windows_out receives up to n rows of
 * (r0, c0, h, w), index_out[i] the window of point i, *n_windows the number of windows.  Every point's own window lies inside
 * its planned window, so the patch cropped from the planned window around (row - r0, col - c0) is the patch of the whole image. */
int mc_plan_uploads(int32_t height, int32_t width, const int32_t* rowcols, int64_t n_points, int32_t* windows_out,
                    int32_t* index_out, int64_t* n_windows);

/* ---- (f)2: image decode feeding the crop kernel: spacer.storage.load_image (call site
 *      mermaid_classifier/pyspacer/annotation.py:235; inside spacer.tasks.extract_features,
 *      scripts/build_feature_bucket.py:775) for JPEG streams.  Huffman decoding runs on the calling host thread,
 *      IDCT / upsampling / colour conversion on the GPU (nvJPEG), RGB8 interleaved straight into a device buffer that
 *      mc_extract_points reads: the decoded image never crosses PCIe.  A grayscale stream becomes three equal channels
 *      (PIL convert("RGB")); CMYK streams are rejected with MC_ERR_UNSUPPORTED.  mc_jpeg_decode is asynchronous on
 *      `stream` once the host part is done.  One decoder handle per host thread. ------------------------------------ */
typedef struct mc_jpeg mc_jpeg;
int mc_jpeg_create(int32_t device, mc_jpeg** out);
int mc_jpeg_destroy(mc_jpeg* d);
int mc_jpeg_info(mc_jpeg* d, const uint8_t* jpeg_host, int64_t n_bytes, int32_t* height, int32_t* width,
                 int32_t* components);
int mc_jpeg_decode(mc_jpeg* d, const uint8_t* jpeg_host, int64_t n_bytes, uint8_t* rgb_dev, int64_t row_pitch,
                   int32_t height, int32_t width, void* stream);
/* The same, BIT FOR BIT what PIL / libjpeg-turbo decode (the reference's spacer.storage.load_image): the entropy decoding runs
 * on the calling thread, libjpeg-turbo's integer IDCT (jidctint.c), fancy chroma upsampling (jdsample.c) and YCbCr -> RGB
 * tables (jdcolor.c) are restated as kernels.  Baseline / extended-sequential and progressive (jdphuff.c) Huffman streams,
 * grayscale or YCbCr 4:4:4 / 4:2:2 / 4:2:0; anything else returns MC_ERR_UNSUPPORTED (use mc_jpeg_decode or a host decoder).  Asynchronous on `stream`
 * after the host-side entropy decoding. */
int mc_jpeg_decode_exact(mc_jpeg* d, const uint8_t* jpeg_host, int64_t n_bytes, uint8_t* rgb_dev, int64_t row_pitch,
                         int32_t height, int32_t width, void* stream);
/* Host-only half of the exact decoder: quantised DCT coefficients (dense int16 [block][64], natural order, component after
 * component over each component's whole MCU-padded block grid); info = {height, width, components, bx0, by0, bx1, by1, bx2,
 * by2, restart interval}; coef_out == NULL only fills info.  Needs no CUDA device. */
int mc_jpeg_coefficients_host(const uint8_t* jpeg_host, int64_t n_bytes, int16_t* coef_out, int64_t capacity_blocks,
                              int32_t* info /* [10] */);

/* Debug/parity tap: during the NEXT extract call, copy one internal NHWC activation of the
 * first sub-batch to `out_dev` as fp32 (at most `capacity` elements).
 * layer: 0 = stem, 1+4*b+{0,1,2,3} = block b {expand, depthwise, SE gate, block out},
 * 65 = head conv (pre-pool).  layer < 0 disarms the tap. */
int mc_extractor_set_tap(mc_extractor* h, int32_t layer, float* out_dev, int64_t capacity);

/* Per-layer CUDA-event timing on the launching stream.  layer: -1 off (default), -2 every
 * layer, >= 0 that layer only (ids as for mc_extractor_set_tap; +2 = SE FCs, 66 = pool).
 * While enabled, extract calls synchronise the stream before returning.
 * mc_extractor_profile_read copies accumulated milliseconds / launch counts (67 entries each). */
#define MC_N_LAYERS 67
int mc_extractor_profile(mc_extractor* h, int32_t layer);
int mc_extractor_profile_read(mc_extractor* h, double* ms_out, int64_t* count_out, int32_t reset);

/* ---- A6: CalibratedHead (mermaid_classifier/pyspacer/inference/head.py:25-89) behind
 *      Predictor.predict_proba (inference/loader.py:30-35).
 *      weights_host[i] is (dims[i+1] x dims[i]) row-major fp32, biases_host[i] dims[i+1];
 *      a/b are the per-class Platt parameters, or NULL for the uncalibrated softmax path
 *      of TorchMLPClassifier._forward_probs (torch_classifier.py:332-370). ------------ */
int mc_head_create(int32_t n_layers, const int32_t* dims /* n_layers + 1 */,
                   const float* const* weights_host, const float* const* biases_host,
                   const float* platt_a_host, const float* platt_b_host, int32_t device,
                   mc_head** out);
int mc_head_destroy(mc_head* h);
/* The Linear/ReLU chain runs on the tensor cores (tcgen05, 3xTF32 split, fp32-class) by default; exact != 0 selects the
 * exact-fp32 CUDA-core GEMM (used by predict_proba, whose contract is the reference's 1e-6 export gate). */
int mc_head_set_exact(mc_head* h, int32_t exact);

/* features_dev: n x dims[0] fp32.  Any of the outputs may be NULL.
 *   proba_dev  : n x K float64 (predict_proba)
 *   labels_dev : n int32, np.argmax tie-break (lowest index)
 *   topk_*     : n x topk, descending, stable (annotation.py:253-259) */
int mc_head_scores(mc_head* h, const float* features_dev, int64_t n, double* proba_dev,
                   int32_t* labels_dev, int32_t topk, int32_t* topk_idx_dev, float* topk_val_dev,
                   void* stream);
int mc_head_scores_host(mc_head* h, const float* features_host, int64_t n, double* proba_host,
                        int32_t* labels_host, void* stream);
int64_t mc_head_launches(const mc_head* h); /* kernels launched so far by this handle */

/* ---- A8: batched evaluation of MermaidTrainer (mermaid_classifier/pyspacer/trainer.py:295-342):
 *      accuracy_score of the argmax labels and sklearn.metrics.log_loss(labels=classes) of
 *      predict_proba against integer targets y_dev (positions in classes_), without materialising the
 *      (n x K) matrix.  Returns host values (synchronises `stream`): n_correct = #(argmax == y),
 *      loss_sum = sum_i -log(clip(p[i, y_i], eps, 1 - eps)), eps = DBL_EPSILON; a target outside
 *      [0, K) makes loss_sum NaN.  Fixed-order reduction: bit-reproducible. ------------------- */
int mc_head_evaluate(mc_head* h, const float* features_dev, const int32_t* y_dev, int64_t n,
                     int64_t* n_correct, double* loss_sum, void* stream);

/* ---- (f)3: Platt calibration of MermaidTrainer._calibrate_in_batches (trainer.py:344-396), i.e.
 *      sklearn 1.5.2 _sigmoid_calibration(proba[:, k], y == k) for every class k at once.
 *      proba_dev: n x K float64 (mc_head_scores of the uncalibrated head), y_dev: n int32 targets.
 *      Newton + backtracking on the device until every class has |grad|_inf < gtol (sum form) or
 *      max_passes matrix passes are spent.  a_out/b_out (host, K doubles): sigmoid(-(a p + b));
 *      loss_out (host, K, or NULL): the per-class objective at (a, b); passes_out: passes used. */
int mc_platt_fit(const double* proba_dev, const int32_t* y_dev, int64_t n, int32_t n_classes,
                 int32_t device, double gtol, int32_t max_passes, double* a_out, double* b_out,
                 double* loss_out, int32_t* passes_out, void* stream);

/* ---- A7: TorchMLPClassifier.partial_fit inner loop
 *      (mermaid_classifier/pyspacer/torch_classifier.py:226-303): per mini-batch
 *      weighted CE + 0.5*alpha/mb*sum(W^2), backward, Adam.  Parameters and Adam state live
 *      on the device.  weights_host[i] is (dims[i+1] x dims[i]) row-major (the xavier_uniform /
 *      zero-bias init of torch_classifier.py:62-73 is done by the caller with torch's RNG so it
 *      is bit-identical to the reference's). ------------------------------------------------- */
int mc_mlp_create(int32_t n_layers, const int32_t* dims, const float* const* weights_host,
                  const float* const* biases_host, const float* class_weight_host /* K or NULL */,
                  float lr, float alpha, float beta1, float beta2, float eps, int32_t device,
                  mc_mlp** out);
int mc_mlp_destroy(mc_mlp* h);

/* Data-parallel communicator for the gradient all-reduce (NCCL over NVLink; one process per
 * GPU).  Rank 0 calls mc_dp_unique_id and hands the 128 bytes to every rank out of band
 * (torch.distributed broadcast in the Python mirror); every rank then calls mc_dp_create. */
typedef struct mc_dp mc_dp;
int mc_dp_unique_id(char* id_out_128);
int mc_dp_create(const char* id_128, int32_t rank, int32_t world, int32_t device, mc_dp** out);
int mc_dp_destroy(mc_dp* d);
int mc_dp_all_reduce_sum(mc_dp* d, float* buf_dev, int64_t n, void* stream);

/* One partial_fit pass = n_steps Adam steps.  Step s trains on THIS RANK's rows
 * [step_offsets_host[s], step_offsets_host[s+1]) of x_dev (n x dims[0] fp32) / y_dev (n int32
 * class indices), taken through order_dev (int64 row indices, the shuffled order; NULL =
 * identity).  A step may have zero local rows (ragged tail under data parallelism).
 * After each backward the flat gradient buffer -- un-normalised sums followed by the 4
 * statistics [sum w, sum w*nll, rows, 0], mc_mlp_grad_size() floats -- is all-reduced over `dp`
 * (NULL = single GPU) and then handed to the optional `grad_sync` hook on the stream; the Adam
 * kernel normalises by the GLOBAL statistics, so the update equals the reference's for the
 * global mini-batch.  *loss_out_host receives the loss_curve_ entry (sample-weighted mean of
 * the regularised mini-batch losses, torch_classifier.py:295-301); passing it synchronises.
 * Without `dp` and `grad_sync`, runs of >= 8 equal-sized steps are replayed from a captured CUDA graph of two steps (the
 * same kernels on a staged copy of each mini-batch: bit-identical parameters; the call then synchronises `stream` once
 * per run to hand the step table to the device); MC_MLP_GRAPH=0 launches every step kernel by kernel. */
typedef void (*mc_grad_sync_fn)(float* grad_dev, int64_t n_grad, void* stream, void* user);
int mc_mlp_partial_fit(mc_mlp* h, const float* x_dev, const int32_t* y_dev, const int64_t* order_dev,
                       const int64_t* step_offsets_host, int32_t n_steps, mc_dp* dp,
                       mc_grad_sync_fn grad_sync, void* user, double* loss_out_host, void* stream);
int mc_mlp_get_params(mc_mlp* h, float* const* weights_host, float* const* biases_host);
/* Adam first/second moments and step count (pickling: torch_classifier.py:412-444). */
int mc_mlp_get_adam(mc_mlp* h, float* const* m_w_host, float* const* m_b_host, float* const* v_w_host,
                    float* const* v_b_host, int64_t* t_out);
int mc_mlp_set_adam(mc_mlp* h, const float* const* m_w_host, const float* const* m_b_host,
                    const float* const* v_w_host, const float* const* v_b_host, int64_t t);
int64_t mc_mlp_steps(const mc_mlp* h);     /* Adam steps taken */
int64_t mc_mlp_launches(const mc_mlp* h);  /* kernels launched so far by this handle */
int64_t mc_mlp_graph_steps(const mc_mlp* h);  /* Adam steps of this handle taken by CUDA-graph replay (runs of >= 8 equal-sized steps) */
int64_t mc_mlp_grad_size(const mc_mlp* h); /* floats in the flat gradient buffer (incl. 4 statistics) */

#ifdef __cplusplus
}
#endif
#endif /* MERMAID_B200_H */
