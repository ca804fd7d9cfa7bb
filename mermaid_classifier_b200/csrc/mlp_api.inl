// mc_mlp_* / mc_dp_* entry points: the MLP-head training inner loop (A7) and its data-parallel
// gradient all-reduce (NCCL over NVLink, resolved at run time from the libnccl the process has
// already loaded -- torch's bundled copy -- so the library itself carries no link dependency).
#include <dlfcn.h>

#include "mlp_train.cuh"

namespace {

// ---- minimal NCCL surface (nccl.h 2.27: ncclUniqueId is 128 opaque bytes; ncclFloat32 = 7, ncclSum = 0)
struct NcclId { char internal[128]; };
typedef int (*nccl_get_unique_id_fn)(NcclId*);
typedef int (*nccl_comm_init_rank_fn)(void**, int, NcclId, int);
typedef int (*nccl_comm_destroy_fn)(void*);
typedef int (*nccl_all_reduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_get_error_string_fn)(int);

struct NcclApi {
  void* lib = nullptr;
  nccl_get_unique_id_fn get_unique_id = nullptr;
  nccl_comm_init_rank_fn comm_init_rank = nullptr;
  nccl_comm_destroy_fn comm_destroy = nullptr;
  nccl_all_reduce_fn all_reduce = nullptr;
  nccl_get_error_string_fn error_string = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy torch already mapped, if any
    if (api.lib) break;
  }
  if (!api.lib)
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
  if (!api.lib) return nullptr;
  api.get_unique_id = (nccl_get_unique_id_fn)dlsym(api.lib, "ncclGetUniqueId");
  api.comm_init_rank = (nccl_comm_init_rank_fn)dlsym(api.lib, "ncclCommInitRank");
  api.comm_destroy = (nccl_comm_destroy_fn)dlsym(api.lib, "ncclCommDestroy");
  api.all_reduce = (nccl_all_reduce_fn)dlsym(api.lib, "ncclAllReduce");
  api.error_string = (nccl_get_error_string_fn)dlsym(api.lib, "ncclGetErrorString");
  if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_reduce) {
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

int nccl_fail(NcclApi* api, const char* what, int rc) {
  return fail(MC_ERR_CUDA, std::string(what) + ": NCCL error " + std::to_string(rc) +
                               (api && api->error_string ? std::string(" (") + api->error_string(rc) + ")" : ""));
}

}  // namespace

struct mc_dp {
  void* comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

struct mc_mlp {
  int device = 0, L = 0;
  std::vector<int> dims, dims_p;
  MlpSegs segs{};
  int64_t n_flat = 0;
  float *d_p = nullptr, *d_m = nullptr, *d_v = nullptr, *d_g = nullptr, *d_cw = nullptr;
  float* d_ssq[2] = {nullptr, nullptr};
  int n_blocks = 0, ssq_cur = 0;
  double* d_loss = nullptr;
  std::vector<float*> d_act, d_delta;  // [L] each: outputs of layer i / gradient w.r.t. them
  float* d_part[2] = {nullptr, nullptr};   // split-K partials; slot 1 = the second GEMM of a paired launch
  int64_t cap_part[2] = {0, 0};
  bool rl_attr_set = false;   // dynamic shared-memory attribute of mlp_rowlocal_kernel set on this handle's device
  int* d_tickets = nullptr;   // split-K tile tickets, n_tickets per slot (zero between launches)
  int n_tickets = 0;
  float2* d_rowstat = nullptr;
  int cap_rows = 0;
  float* d_xs = nullptr;
  int64_t cap_xs = 0;
  int32_t* d_ys = nullptr;
  int64_t cap_ys = 0;
  float lr = 1e-3f, alpha = 1e-4f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
  int64_t t = 0;          // Adam steps taken
  int64_t launches = 0;   // kernels launched
  // CUDA-graph replay of runs of equal-sized steps (mc_mlp_partial_fit): one captured PAIR of steps (the weight-norm partials
  // ping-pong between two buffers), a device cursor, the mini-batch staging buffers the captured kernels read
  cudaStream_t cap_st = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int g_rows = 0, g_parity = 0;
  const void* g_ptrs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  MlpCtl* d_ctl = nullptr;
  float* d_xb = nullptr;
  int64_t cap_xb = 0;
  int32_t* d_yb = nullptr;
  int64_t cap_yb = 0;
  int64_t* d_offs = nullptr;
  int64_t cap_offs = 0;
  float2* d_bc = nullptr;
  int64_t cap_bc = 0;
  int64_t graph_steps = 0;   // Adam steps taken through graph replay (diagnostics / tests)
};

namespace {

int mlp_ensure_rows(mc_mlp* h, int rows) {
  if (rows <= h->cap_rows) return MC_OK;
  const int cap = std::max(rows, 256);
  for (int i = 0; i < h->L; ++i) {
    if (h->d_act[i]) cudaFree(h->d_act[i]);
    if (h->d_delta[i]) cudaFree(h->d_delta[i]);
    h->d_act[i] = h->d_delta[i] = nullptr;
    MC_CUDA(cudaMalloc((void**)&h->d_act[i], (size_t)cap * h->dims_p[i + 1] * sizeof(float)));
    MC_CUDA(cudaMalloc((void**)&h->d_delta[i], (size_t)cap * h->dims_p[i + 1] * sizeof(float)));
  }
  if (h->d_rowstat) cudaFree(h->d_rowstat);
  h->d_rowstat = nullptr;
  MC_CUDA(cudaMalloc((void**)&h->d_rowstat, (size_t)cap * sizeof(float2)));
  h->cap_rows = cap;
  return MC_OK;
}

// C = A * B with the layout flags of mlp_gemm_kernel; split-K when the tile grid would leave most SMs idle.
int mlp_gemm_plan(mc_mlp* h, MlpGemmP* p, bool a_mc, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M,
                  int N, int K, int epi, const float* bias, const float* mask, int ldmask, float* colsum, int slot = 0) {
  const int tiles = cdiv(M, 64) * cdiv(N, 64);
  int splits = 1;
  // few tiles and a long reduction (forward layers and the delta back-propagation at mini-batch 200): split K so that
  // ~300 CTAs (two per SM) work instead of 8-40; the partials are summed in slice order (deterministic) by the last slice
  // Measured: splitting further (640 CTAs, and the weight-gradient GEMMs over the 200 mini-batch rows as well) is SLOWER
  // (6.9 k -> 6.5 k Adam steps/s): these GEMMs are bound by the shared-memory traffic of the 4x4 micro-tile (128 KB per
  // 64 x 64 x 16 step) plus ~5 us of launch / prologue each, not by their serial k-steps, and every extra slice adds a
  // partial tile of traffic.  MC_MLP_SPLIT_CTAS overrides the CTA target.
  static const int split_ctas = getenv("MC_MLP_SPLIT_CTAS") ? atoi(getenv("MC_MLP_SPLIT_CTAS")) : 296;
  if (!a_mc && tiles < 64 && tiles <= h->n_tickets && K >= 256) {
    splits = std::min(cdiv(K, 64), std::max(1, split_ctas / tiles));
  }
  int kps = K;
  if (splits > 1) {
    kps = cdiv(cdiv(K, splits), 16) * 16;
    splits = cdiv(K, kps);
  }
  if (splits > 1) {
    if (tiles > h->n_tickets) return fail(MC_ERR_UNSUPPORTED, "mlp_gemm: more output tiles than split-K tickets");
    int rc = grow(&h->d_part[slot], &h->cap_part[slot], (int64_t)splits * M * N + (int64_t)splits * M);
    if (rc) return rc;
  }
  *p = MlpGemmP{A, lda, B, ldb, C, ldc, h->d_part[slot], h->d_tickets + slot * h->n_tickets, M, N, K, kps, splits, epi, bias, mask, ldmask, colsum,
                cdiv(M, 64), cdiv(N, 64)};
  return MC_OK;
}

int mlp_gemm(mc_mlp* h, bool a_mc, bool b_nc, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M,
             int N, int K, int epi, const float* bias, const float* mask, int ldmask, float* colsum, cudaStream_t st) {
  MlpGemmP p;
  int rc = mlp_gemm_plan(h, &p, a_mc, A, lda, B, ldb, C, ldc, M, N, K, epi, bias, mask, ldmask, colsum);
  if (rc) return rc;
  dim3 grid(p.gx, p.gy, p.splits);
  if (!a_mc && !b_nc) mlp_gemm_kernel<false, false><<<grid, 256, 0, st>>>(p);
  else if (a_mc && b_nc) mlp_gemm_kernel<true, true><<<grid, 256, 0, st>>>(p);
  else if (!a_mc && b_nc) mlp_gemm_kernel<false, true><<<grid, 256, 0, st>>>(p);
  else return fail(MC_ERR_UNSUPPORTED, "mlp_gemm: layout combination");
  MC_CHECK_LAUNCH();
  h->launches++;
  return MC_OK;
}

// dW (A_MC, B_NC, bias gradient) and dX (B_NC, ReLU mask, split-K) of one layer in ONE launch
int mlp_gemm_pair(mc_mlp* h, const MlpGemmP& dw, const MlpGemmP& dx, cudaStream_t st) {
  mlp_gemm_pair_kernel<<<dw.gx * dw.gy * dw.splits + dx.gx * dx.gy * dx.splits, 256, 0, st>>>(dw, dx);
  MC_CHECK_LAUNCH();
  h->launches++;
  return MC_OK;
}

int mlp_ssq_init(mc_mlp* h, cudaStream_t st) {
  mlp_ssq_kernel<<<h->n_blocks, MLP_ADAM_THREADS, 0, st>>>(h->d_p, h->segs, h->d_ssq[h->ssq_cur]);
  MC_CHECK_LAUNCH();
  h->launches++;
  return MC_OK;
}

// The captured pair of steps for mini-batches of `rows` rows: built on first use, rebuilt when `rows` or any buffer the
// captured kernels address has changed.  `body(stream)` issues the two steps on the capture stream.
template <typename Body>
int mlp_graph_prepare(mc_mlp* h, int rows, int Dp, Body body) {
  int rc;
  if (!h->cap_st) MC_CUDA(cudaStreamCreateWithFlags(&h->cap_st, cudaStreamNonBlocking));
  if (!h->d_ctl) MC_CUDA(cudaMalloc((void**)&h->d_ctl, sizeof(MlpCtl)));
  if ((rc = grow(&h->d_xb, &h->cap_xb, (int64_t)rows * Dp)) || (rc = grow(&h->d_yb, &h->cap_yb, (int64_t)rows))) return rc;
  const void* now[8] = {h->d_part[0], h->d_part[1], h->d_act[0], h->d_delta[0], h->d_rowstat, h->d_xb, h->d_yb, h->d_ctl};
  bool valid = h->gexec != nullptr && h->g_rows == rows;
  for (int i = 0; i < 8 && valid; ++i) valid = h->g_ptrs[i] == now[i];
  if (valid) return MC_OK;
  if (h->gexec) {
    cudaGraphExecDestroy(h->gexec);
    h->gexec = nullptr;
  }
  const int parity = h->ssq_cur;
  MC_CUDA(cudaStreamBeginCapture(h->cap_st, cudaStreamCaptureModeThreadLocal));
  rc = body(h->cap_st);
  cudaGraph_t graph = nullptr;
  const cudaError_t e_end = cudaStreamEndCapture(h->cap_st, &graph);   // always: the stream must leave capture mode
  h->ssq_cur = parity;   // a captured pair toggles twice; a failed capture may have toggled once
  if (rc != MC_OK || e_end != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc != MC_OK ? rc : fail(MC_ERR_CUDA, std::string("mc_mlp_partial_fit: graph capture: ") + cudaGetErrorString(e_end));
  }
  const cudaError_t e_inst = cudaGraphInstantiate(&h->gexec, graph, 0);
  cudaGraphDestroy(graph);
  if (e_inst != cudaSuccess) {
    h->gexec = nullptr;
    return fail(MC_ERR_CUDA, std::string("mc_mlp_partial_fit: cudaGraphInstantiate: ") + cudaGetErrorString(e_inst));
  }
  h->g_rows = rows;
  h->g_parity = parity;
  for (int i = 0; i < 8; ++i) h->g_ptrs[i] = now[i];
  return MC_OK;
}

// pack / unpack between per-layer (out x in) host arrays and the padded flat device layout
void mlp_pack(const mc_mlp* h, const float* const* w, const float* const* b, std::vector<float>& flat) {
  flat.assign((size_t)h->n_flat, 0.f);
  for (int i = 0; i < h->L; ++i) {
    const int ki = h->dims[i], no = h->dims[i + 1], kp = h->dims_p[i];
    for (int r = 0; r < no; ++r) memcpy(&flat[h->segs.w_off[i] + (size_t)r * kp], w[i] + (size_t)r * ki, ki * sizeof(float));
    if (b) memcpy(&flat[h->segs.b_off[i]], b[i], no * sizeof(float));
  }
}
void mlp_unpack(const mc_mlp* h, const std::vector<float>& flat, float* const* w, float* const* b) {
  for (int i = 0; i < h->L; ++i) {
    const int ki = h->dims[i], no = h->dims[i + 1], kp = h->dims_p[i];
    for (int r = 0; r < no; ++r) memcpy(w[i] + (size_t)r * ki, &flat[h->segs.w_off[i] + (size_t)r * kp], ki * sizeof(float));
    if (b) memcpy(b[i], &flat[h->segs.b_off[i]], no * sizeof(float));
  }
}

}  // namespace

extern "C" {

// ---- data-parallel communicator -----------------------------------------------------------
int mc_dp_unique_id(char* id_out_128) {
  if (!id_out_128) return fail(MC_ERR_BAD_ARG, "mc_dp_unique_id: null");
  NcclApi* api = nccl_api();
  if (!api) return fail(MC_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
  NcclId id;
  int rc = api->get_unique_id(&id);
  if (rc) return nccl_fail(api, "ncclGetUniqueId", rc);
  memcpy(id_out_128, id.internal, 128);
  return MC_OK;
}

int mc_dp_create(const char* id_128, int32_t rank, int32_t world, int32_t device, mc_dp** out) {
  if (!id_128 || !out || world < 1 || rank < 0 || rank >= world) return fail(MC_ERR_BAD_ARG, "mc_dp_create: bad argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(MC_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
  DeviceGuard g(device);
  NcclId id;
  memcpy(id.internal, id_128, 128);
  mc_dp* d = new mc_dp();
  d->rank = rank;
  d->world = world;
  d->device = device;
  int rc = api->comm_init_rank(&d->comm, world, id, rank);
  if (rc) {
    delete d;
    return nccl_fail(api, "ncclCommInitRank", rc);
  }
  *out = d;
  return MC_OK;
}

int mc_dp_destroy(mc_dp* d) {
  if (!d) return MC_OK;
  NcclApi* api = nccl_api();
  if (api && d->comm) api->comm_destroy(d->comm);
  delete d;
  return MC_OK;
}

int mc_dp_all_reduce_sum(mc_dp* d, float* buf_dev, int64_t n, void* stream) {
  if (!d || !buf_dev || n < 0) return fail(MC_ERR_BAD_ARG, "mc_dp_all_reduce_sum: bad argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(MC_ERR_UNSUPPORTED, "libnccl.so.2 is not loadable in this process");
  int rc = api->all_reduce(buf_dev, buf_dev, (size_t)n, 7 /*ncclFloat32*/, 0 /*ncclSum*/, d->comm, (cudaStream_t)stream);
  if (rc) return nccl_fail(api, "ncclAllReduce", rc);
  return MC_OK;
}

// ---- trainer ------------------------------------------------------------------------------------
int mc_mlp_create(int32_t n_layers, const int32_t* dims, const float* const* weights, const float* const* biases,
                  const float* class_weight, float lr, float alpha, float beta1, float beta2, float eps, int32_t device,
                  mc_mlp** out) {
  if (n_layers < 1 || n_layers > 8 || !dims || !weights || !biases || !out)
    return fail(MC_ERR_BAD_ARG, "mc_mlp_create: null argument or layer count outside [1, 8]");
  for (int i = 0; i <= n_layers; ++i)
    if (dims[i] < 1) return fail(MC_ERR_BAD_ARG, "mc_mlp_create: non-positive layer width");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MC_ERR_CUDA, "no CUDA device: libmermaid_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(MC_ERR_BAD_ARG, "mc_mlp_create: bad device index");
  DeviceGuard g(device);
  mc_mlp* h = new mc_mlp();
  h->device = device;
  h->L = n_layers;
  h->dims.assign(dims, dims + n_layers + 1);
  for (int d : h->dims) h->dims_p.push_back((d + 3) / 4 * 4);
  h->lr = lr; h->alpha = alpha; h->beta1 = beta1; h->beta2 = beta2; h->eps = eps;
  int64_t off = 0;
  h->segs.n_layers = n_layers;
  for (int i = 0; i < n_layers; ++i) {
    h->segs.w_off[i] = off;
    off += (int64_t)h->dims_p[i + 1] * h->dims_p[i];
    h->segs.b_off[i] = off;
    off += h->dims_p[i + 1];
  }
  h->segs.end = off;
  h->n_flat = off;
  h->n_blocks = cdiv(off, MLP_ADAM_THREADS);
  h->d_act.assign(n_layers, nullptr);
  h->d_delta.assign(n_layers, nullptr);
  std::vector<float> flat;
  mlp_pack(h, weights, biases, flat);
  cudaError_t e = cudaMalloc((void**)&h->d_p, off * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_m, off * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_v, off * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_g, (off + 4) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_ssq[0], h->n_blocks * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_ssq[1], h->n_blocks * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_loss, 2 * sizeof(double));
  h->n_tickets = 256;  // split-K only runs on grids of at most this many tiles
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_tickets, (2 * h->n_tickets + 1) * sizeof(int));   // + the CE kernel's
  if (e == cudaSuccess) e = cudaMemset(h->d_tickets, 0, (2 * h->n_tickets + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_p, flat.data(), off * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_m, 0, off * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(h->d_v, 0, off * sizeof(float));
  if (e == cudaSuccess && class_weight) {
    e = cudaMalloc((void**)&h->d_cw, h->dims.back() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_cw, class_weight, h->dims.back() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    mc_mlp_destroy(h);
    return fail(MC_ERR_CUDA, std::string("mc_mlp_create: ") + cudaGetErrorString(e));
  }
  int rc = mlp_ssq_init(h, nullptr);
  if (rc == MC_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(MC_ERR_CUDA, "mc_mlp_create: ssq init failed");
  if (rc) {
    mc_mlp_destroy(h);
    return rc;
  }
  *out = h;
  return MC_OK;
}

int mc_mlp_destroy(mc_mlp* h) {
  if (!h) return MC_OK;
  DeviceGuard g(h->device);
  for (float* p : h->d_act) if (p) cudaFree(p);
  for (float* p : h->d_delta) if (p) cudaFree(p);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  if (h->cap_st) cudaStreamDestroy(h->cap_st);
  void* ptrs[] = {h->d_p, h->d_m, h->d_v, h->d_g, h->d_cw, h->d_ssq[0], h->d_ssq[1], h->d_loss,
                  h->d_part[0], h->d_part[1], h->d_rowstat, h->d_xs, h->d_ys, h->d_tickets,
                  h->d_ctl, h->d_xb, h->d_yb, h->d_offs, h->d_bc};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
  return MC_OK;
}

int mc_mlp_partial_fit(mc_mlp* h, const float* x_dev, const int32_t* y_dev, const int64_t* order_dev,
                       const int64_t* step_offsets, int32_t n_steps, mc_dp* dp, mc_grad_sync_fn grad_sync, void* user,
                       double* loss_out_host, void* stream) {
  if (!h) return fail(MC_ERR_BAD_ARG, "null handle");
  if (n_steps < 0 || (n_steps > 0 && !step_offsets)) return fail(MC_ERR_BAD_ARG, "mc_mlp_partial_fit: bad steps");
  const int64_t n = n_steps > 0 ? step_offsets[n_steps] : 0;
  if (n_steps > 0 && step_offsets[0] != 0) return fail(MC_ERR_BAD_ARG, "mc_mlp_partial_fit: step_offsets[0] must be 0");
  int max_rows = 0;
  for (int s = 0; s < n_steps; ++s) {
    const int64_t r = step_offsets[s + 1] - step_offsets[s];
    if (r < 0 || r > (1 << 20)) return fail(MC_ERR_BAD_ARG, "mc_mlp_partial_fit: step size outside [0, 2^20]");
    max_rows = std::max<int>(max_rows, (int)r);
  }
  if (n > 0 && (!x_dev || !y_dev)) return fail(MC_ERR_BAD_ARG, "mc_mlp_partial_fit: null data");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const int L = h->L, D = h->dims[0], Dp = h->dims_p[0], K = h->dims.back(), Kp = h->dims_p.back();
  if ((rc = mlp_ensure_rows(h, max_rows))) return rc;
  const float* xs = x_dev;
  const int32_t* ys = y_dev;
  if (n > 0 && (order_dev || D != Dp)) {
    if ((rc = grow(&h->d_xs, &h->cap_xs, n * Dp)) || (rc = grow(&h->d_ys, &h->cap_ys, n))) return rc;
    mlp_gather_rows_kernel<<<cdiv(n * Dp, 256), 256, 0, st>>>(x_dev, y_dev, order_dev, D, Dp, h->d_xs, h->d_ys, n);
    MC_CHECK_LAUNCH();
    h->launches++;
    xs = h->d_xs;
    ys = h->d_ys;
  }
  MC_CUDA(cudaMemsetAsync(h->d_loss, 0, 2 * sizeof(double), st));
  // graph replay (below) redirects a step to the staging buffers, the cursor and the capture stream
  const float* x_over = nullptr;
  const int32_t* y_over = nullptr;
  const MlpCtl* ctl_arg = nullptr;
  auto run_step = [&](int s) -> int {
    const int rows = (int)(step_offsets[s + 1] - step_offsets[s]);
    const float* x = x_over ? x_over : xs + step_offsets[s] * Dp;
    const int32_t* y_step = y_over ? y_over : ys + step_offsets[s];
    // Experiment, off by default (MC_MLP_ROWLOCAL=1; MC_MLP_RL_R = rows per CTA): first-layer GEMM -> ONE launch for layers
    // 1..L-1, the loss and every delta (mlp_rowlocal_kernel) -> one launch for all weight gradients: 4 launches per step
    // instead of 10.  Measured at (500,300,100) / 500 classes / mini-batch 200: 6.0 k Adam steps/s (R = 2) against 7.1 k for
    // the GEMM chain -- every CTA streams all 1.84 MB of the layer 1..L-1 weights from L2 (forward + backward), 184 MB per
    // step at R = 2, and without the row-tiling of a GEMM that stream, not the launch count, is the bound.
    static const bool no_rl = getenv("MC_MLP_ROWLOCAL") == nullptr;
    static const int rl_r = getenv("MC_MLP_RL_R") ? atoi(getenv("MC_MLP_RL_R")) : 2;
    int rl_width = 0, rl_maxw = 0;
    for (int i = 1; i <= L; ++i) {
      rl_width += h->dims_p[i];
      rl_maxw = std::max(rl_maxw, h->dims_p[i]);
    }
    const size_t rl_smem = (size_t)rl_r * (rl_width + 2 * rl_maxw) * sizeof(float);
    const bool use_rl = !no_rl && L >= 2 && L <= MLP_RL_MAX_LAYERS && rows > 0 && rows <= 4096 && rl_smem <= 200 * 1024 &&
                        (rl_r == 2 || rl_r == 4 || rl_r == 8);
    if (use_rl) {
      const int N0 = h->dims_p[1];
      if ((rc = mlp_gemm(h, false, false, x, Dp, h->d_p + h->segs.w_off[0], Dp, h->d_act[0], N0, rows, N0, Dp, MLP_EPI_BIAS_RELU,
                         h->d_p + h->segs.b_off[0], nullptr, 0, nullptr, st)))
        return rc;
      MlpRowLocalP rp{};
      rp.L = L; rp.K = K; rp.rows = rows;
      for (int i = 0; i <= L; ++i) rp.dims_p[i] = h->dims_p[i];
      for (int i = 0; i < L; ++i) {
        rp.w_off[i] = h->segs.w_off[i];
        rp.b_off[i] = h->segs.b_off[i];
        rp.act[i] = h->d_act[i];
        rp.delta[i] = h->d_delta[i];
      }
      rp.params = h->d_p;
      rp.act0 = h->d_act[0];
      rp.y = y_step;
      rp.class_w = h->d_cw;
      rp.row_stat = h->d_rowstat;
      rp.stats = h->d_g + h->n_flat;
      rp.ticket = h->d_tickets + 2 * h->n_tickets;
      if (!h->rl_attr_set) {
        MC_CUDA(cudaFuncSetAttribute(mlp_rowlocal_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MC_CUDA(cudaFuncSetAttribute(mlp_rowlocal_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MC_CUDA(cudaFuncSetAttribute(mlp_rowlocal_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        h->rl_attr_set = true;
      }
      if (rl_r == 2) mlp_rowlocal_kernel<2><<<cdiv(rows, 2), 512, rl_smem, st>>>(rp);
      else if (rl_r == 4) mlp_rowlocal_kernel<4><<<cdiv(rows, 4), 512, rl_smem, st>>>(rp);
      else mlp_rowlocal_kernel<8><<<cdiv(rows, 8), 512, rl_smem, st>>>(rp);
      MC_CHECK_LAUNCH();
      h->launches++;
      MlpGemmMulti mp{};
      mp.n = L;
      int first = 0;
      for (int i = 0; i < L; ++i) {
        const int Ki = h->dims_p[i], Ni = h->dims_p[i + 1];
        const float* lin = i == 0 ? x : h->d_act[i - 1];
        if ((rc = mlp_gemm_plan(h, &mp.p[i], true, h->d_delta[i], Ni, lin, Ki, h->d_g + h->segs.w_off[i], Ki, Ni, Ki, rows,
                                MLP_EPI_NONE, nullptr, nullptr, 0, h->d_g + h->segs.b_off[i])))
          return rc;
        mp.first[i] = first;
        first += mp.p[i].gx * mp.p[i].gy * mp.p[i].splits;
      }
      mp.first[L] = first;
      mlp_gemm_multi_kernel<<<first, 256, 0, st>>>(mp);
      MC_CHECK_LAUNCH();
      h->launches++;
    } else if (rows > 0) {
      // forward
      const float* in = x;
      for (int i = 0; i < L; ++i) {
        const int Ki = h->dims_p[i], Ni = h->dims_p[i + 1];
        if ((rc = mlp_gemm(h, false, false, in, Ki, h->d_p + h->segs.w_off[i], Ki, h->d_act[i], Ni, rows, Ni, Ki,
                           i < L - 1 ? MLP_EPI_BIAS_RELU : MLP_EPI_BIAS, h->d_p + h->segs.b_off[i], nullptr, 0, nullptr, st)))
          return rc;
        in = h->d_act[i];
      }
      // loss + un-normalised output delta + statistics
      mlp_ce_kernel<<<cdiv(rows, 8), 256, 0, st>>>(h->d_act[L - 1], Kp, K, y_step, h->d_cw, h->d_delta[L - 1],
                                                  h->d_rowstat, rows, h->d_g + h->n_flat, h->d_tickets + 2 * h->n_tickets);
      MC_CHECK_LAUNCH();
      h->launches++;
      // backward
      for (int i = L - 1; i >= 0; --i) {
        const int Ki = h->dims_p[i], Ni = h->dims_p[i + 1];
        const float* lin = i == 0 ? x : h->d_act[i - 1];
        if (i == 0) {
          if ((rc = mlp_gemm(h, true, true, h->d_delta[i], Ni, lin, Ki, h->d_g + h->segs.w_off[i], Ki, Ni, Ki, rows,
                             MLP_EPI_NONE, nullptr, nullptr, 0, h->d_g + h->segs.b_off[i], st)))
            return rc;
        } else {
          // dW_i and the delta of the layer below both read delta_i and nothing of each other: one launch
          MlpGemmP dw, dx;
          if ((rc = mlp_gemm_plan(h, &dw, true, h->d_delta[i], Ni, lin, Ki, h->d_g + h->segs.w_off[i], Ki, Ni, Ki, rows,
                                  MLP_EPI_NONE, nullptr, nullptr, 0, h->d_g + h->segs.b_off[i])) ||
              (rc = mlp_gemm_plan(h, &dx, false, h->d_delta[i], Ni, h->d_p + h->segs.w_off[i], Ki, h->d_delta[i - 1], Ki, rows,
                                  Ki, Ni, MLP_EPI_RELU_MASK, nullptr, h->d_act[i - 1], Ki, nullptr, 1)) ||
              (rc = mlp_gemm_pair(h, dw, dx, st)))
            return rc;
        }
      }
    } else {
      MC_CUDA(cudaMemsetAsync(h->d_g, 0, (h->n_flat + 4) * sizeof(float), st));
    }
    if (dp && dp->world > 1 && (rc = mc_dp_all_reduce_sum(dp, h->d_g, h->n_flat + 4, st))) return rc;
    if (grad_sync) grad_sync(h->d_g, h->n_flat + 4, stream, user);
    h->t++;
    const double bc1 = 1.0 - pow((double)h->beta1, (double)h->t);
    const double bc2 = 1.0 - pow((double)h->beta2, (double)h->t);
    mlp_adam_kernel<<<h->n_blocks, MLP_ADAM_THREADS, 0, st>>>(h->d_p, h->d_m, h->d_v, h->d_g, h->segs, h->d_g + h->n_flat,
                                                             h->lr, h->alpha, h->beta1, h->beta2, h->eps, (float)bc1,
                                                             (float)sqrt(bc2), h->d_ssq[h->ssq_cur], h->n_blocks,
                                                             h->d_ssq[h->ssq_cur ^ 1], h->d_loss, ctl_arg);
    MC_CHECK_LAUNCH();
    h->launches++;
    h->ssq_cur ^= 1;
    return MC_OK;
  };
  auto rows_of = [&](int s) { return (int)(step_offsets[s + 1] - step_offsets[s]); };
  // ---- runs of equal-sized steps: CUDA-graph replay -----------------------------------------------------------------------
  // A step is ten dependent launches of 5-25 us; on a host that needs longer than that per launch the loop is launch-bound
  // (measured on two boxes of the same pool: 7.1 k and 2.8 k Adam steps/s).  A run of >= 8 equal-sized steps of a single
  // process is therefore replayed from ONE captured pair of steps: a stage kernel copies the cursor's mini-batch into fixed
  // buffers, the step's kernels read those (same kernels, same order, same operands: bit-identical weights), Adam takes its
  // bias corrections through the cursor, a one-thread kernel advances it.  The first two steps of a run are launched the
  // ordinary way (lazy allocations); MC_MLP_GRAPH=0 turns the replay off.
  static const bool graph_on = !(getenv("MC_MLP_GRAPH") && atoi(getenv("MC_MLP_GRAPH")) == 0) && getenv("MC_MLP_ROWLOCAL") == nullptr;
  const bool graph_ok = graph_on && !(dp && dp->world > 1) && !grad_sync && (reinterpret_cast<uintptr_t>(xs) & 15) == 0;
  bool cursor_ready = false;
  int s = 0;
  while (s < n_steps) {
    const int rows = rows_of(s);
    int e = s;
    while (e < n_steps && rows_of(e) == rows) ++e;
    if (graph_ok && rows > 0 && e - s >= 8) {
      if ((rc = run_step(s++)) || (rc = run_step(s++))) return rc;
      if ((rc = mlp_graph_prepare(h, rows, Dp, [&](cudaStream_t cs) -> int {
             // body of the capture: two steps on the capture stream, reading the staging buffers and the cursor
             const cudaStream_t keep = st;
             const int64_t t_keep = h->t, l_keep = h->launches;
             st = cs;
             x_over = h->d_xb;
             y_over = h->d_yb;
             ctl_arg = h->d_ctl;
             int r2 = MC_OK;
             for (int k = 0; k < 2 && r2 == MC_OK; ++k) {
               mlp_stage_kernel<<<148, 256, 0, cs>>>(h->d_ctl, reinterpret_cast<float4*>(h->d_xb), h->d_yb, rows, Dp / 4);
               r2 = run_step(s);   // same shapes as step s; its pointers and bias corrections are overridden
               mlp_advance_kernel<<<1, 1, 0, cs>>>(h->d_ctl);
             }
             st = keep;
             x_over = nullptr;
             y_over = nullptr;
             ctl_arg = nullptr;
             h->t = t_keep;
             h->launches = l_keep;
             return r2;
           })))
        return rc;
      if (h->ssq_cur != h->g_parity && (rc = run_step(s++))) return rc;
      const int pairs = (e - s) / 2;
      if (pairs > 0) {
        if (!cursor_ready) {
          // the call's step table and bias corrections, computed exactly as run_step computes them
          std::vector<float2> bc((size_t)n_steps);
          for (int i = 0; i < n_steps; ++i) {
            const int64_t t_i = h->t - s + i + 1;   // Adam step number of the call's step i
            bc[(size_t)i] = make_float2((float)(1.0 - pow((double)h->beta1, (double)t_i)),
                                        (float)sqrt(1.0 - pow((double)h->beta2, (double)t_i)));
          }
          if ((rc = grow(&h->d_offs, &h->cap_offs, (int64_t)n_steps + 1)) || (rc = grow(&h->d_bc, &h->cap_bc, (int64_t)n_steps))) return rc;
          MC_CUDA(cudaMemcpyAsync(h->d_offs, step_offsets, ((size_t)n_steps + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
          MC_CUDA(cudaMemcpyAsync(h->d_bc, bc.data(), (size_t)n_steps * sizeof(float2), cudaMemcpyHostToDevice, st));
          MC_CUDA(cudaStreamSynchronize(st));   // `bc` is a local: the copy must have left it
          cursor_ready = true;
        }
        const MlpCtl ctl{(long long)s, xs, ys, h->d_offs, h->d_bc};
        MC_CUDA(cudaMemcpyAsync(h->d_ctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, st));
        MC_CUDA(cudaStreamSynchronize(st));     // `ctl` is a local
        for (int k = 0; k < pairs; ++k) MC_CUDA(cudaGraphLaunch(h->gexec, st));
        h->t += 2 * pairs;
        h->launches += (int64_t)pairs * 2 * (2 * L + 4);   // per step: stage, L forward, loss, L backward, Adam, advance
        h->graph_steps += 2 * pairs;
        s += 2 * pairs;
      }
    }
    while (s < e)
      if ((rc = run_step(s++))) return rc;
  }
  if (loss_out_host) {
    double acc[2] = {0.0, 0.0};
    MC_CUDA(cudaMemcpyAsync(acc, h->d_loss, sizeof(acc), cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaStreamSynchronize(st));
    *loss_out_host = acc[1] > 0.0 ? acc[0] / acc[1] : 0.0;
  }
  return MC_OK;
}

int mc_mlp_get_params(mc_mlp* h, float* const* weights_host, float* const* biases_host) {
  if (!h || !weights_host || !biases_host) return fail(MC_ERR_BAD_ARG, "mc_mlp_get_params: null");
  DeviceGuard g(h->device);
  std::vector<float> flat((size_t)h->n_flat);
  MC_CUDA(cudaMemcpy(flat.data(), h->d_p, h->n_flat * sizeof(float), cudaMemcpyDeviceToHost));
  mlp_unpack(h, flat, weights_host, biases_host);
  return MC_OK;
}

int mc_mlp_get_adam(mc_mlp* h, float* const* m_w, float* const* m_b, float* const* v_w, float* const* v_b, int64_t* t_out) {
  if (!h || !m_w || !m_b || !v_w || !v_b || !t_out) return fail(MC_ERR_BAD_ARG, "mc_mlp_get_adam: null");
  DeviceGuard g(h->device);
  std::vector<float> flat((size_t)h->n_flat);
  MC_CUDA(cudaMemcpy(flat.data(), h->d_m, h->n_flat * sizeof(float), cudaMemcpyDeviceToHost));
  mlp_unpack(h, flat, m_w, m_b);
  MC_CUDA(cudaMemcpy(flat.data(), h->d_v, h->n_flat * sizeof(float), cudaMemcpyDeviceToHost));
  mlp_unpack(h, flat, v_w, v_b);
  *t_out = h->t;
  return MC_OK;
}

int mc_mlp_set_adam(mc_mlp* h, const float* const* m_w, const float* const* m_b, const float* const* v_w,
                    const float* const* v_b, int64_t t) {
  if (!h || !m_w || !m_b || !v_w || !v_b || t < 0) return fail(MC_ERR_BAD_ARG, "mc_mlp_set_adam: bad argument");
  DeviceGuard g(h->device);
  std::vector<float> flat;
  mlp_pack(h, m_w, m_b, flat);
  MC_CUDA(cudaMemcpy(h->d_m, flat.data(), h->n_flat * sizeof(float), cudaMemcpyHostToDevice));
  mlp_pack(h, v_w, v_b, flat);
  MC_CUDA(cudaMemcpy(h->d_v, flat.data(), h->n_flat * sizeof(float), cudaMemcpyHostToDevice));
  h->t = t;
  return MC_OK;
}

int64_t mc_mlp_steps(const mc_mlp* h) { return h ? h->t : 0; }
int64_t mc_mlp_launches(const mc_mlp* h) { return h ? h->launches : 0; }
int64_t mc_mlp_graph_steps(const mc_mlp* h) { return h ? h->graph_steps : 0; }
int64_t mc_mlp_grad_size(const mc_mlp* h) { return h ? h->n_flat + 4 : 0; }

}  // extern "C"
