// C ABI of the device Platt fit (platt.cuh).  Included at the end of api.cu.

extern "C" {

int mc_platt_fit(const double* proba_dev, const int32_t* y_dev, int64_t n, int32_t n_classes, int32_t device, double gtol,
                 int32_t max_passes, double* a_out, double* b_out, double* loss_out, int32_t* passes_out, void* stream) {
  using namespace mc;
  if (!proba_dev || !y_dev || !a_out || !b_out) return fail(MC_ERR_BAD_ARG, "mc_platt_fit: null argument");
  if (n < 1 || n_classes < 1) return fail(MC_ERR_BAD_ARG, "mc_platt_fit: empty problem");
  if (!(gtol > 0.0) || max_passes < 1) return fail(MC_ERR_BAD_ARG, "mc_platt_fit: gtol and max_passes must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MC_ERR_CUDA, "no CUDA device: libmermaid_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(MC_ERR_BAD_ARG, "mc_platt_fit: bad device index");
  DeviceGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  const int K = n_classes;

  unsigned long long* d_cnt = nullptr;
  PlattState* d_state = nullptr;
  double* d_part = nullptr;
  int* d_active = nullptr;
  auto release = [&]() {
    cudaFree(d_cnt);
    cudaFree(d_state);
    cudaFree(d_part);
    cudaFree(d_active);
  };
  cudaError_t e = cudaMalloc(&d_cnt, (size_t)K * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&d_state, (size_t)K * sizeof(PlattState));
  if (e == cudaSuccess) e = cudaMalloc(&d_part, (size_t)PLATT_SLICES * PLATT_TERMS * K * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_active, sizeof(int));
  if (e != cudaSuccess) {
    release();
    return fail(MC_ERR_NOMEM, std::string("mc_platt_fit: cudaMalloc: ") + cudaGetErrorString(e));
  }
  auto bail = [&](cudaError_t err, const char* what) {
    release();
    return fail(MC_ERR_CUDA, std::string("mc_platt_fit: ") + what + ": " + cudaGetErrorString(err));
  };

  if ((e = cudaMemsetAsync(d_cnt, 0, (size_t)K * sizeof(unsigned long long), st)) != cudaSuccess) return bail(e, "memset");
  platt_count_kernel<<<148 * 4, 256, 0, st>>>(y_dev, n, K, d_cnt);
  platt_init_kernel<<<cdiv(K, 128), 128, 0, st>>>(d_cnt, n, K, d_state);
  if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "init launch");

  const dim3 grid(cdiv(K, 128), PLATT_SLICES);
  int passes = 0, active = 1;
  while (passes < max_passes && active > 0) {
    // a few passes per host round trip: finished classes skip their columns, so late passes are cheap
    const int burst = std::min(4, max_passes - passes);
    for (int b = 0; b < burst; ++b) {
      if ((e = cudaMemsetAsync(d_active, 0, sizeof(int), st)) != cudaSuccess) return bail(e, "memset");
      platt_pass_kernel<<<grid, 128, 0, st>>>(proba_dev, y_dev, n, K, d_state, d_part);
      platt_update_kernel<<<cdiv(K, 128), 128, 0, st>>>(d_part, K, d_state, gtol, 100, d_active);
      ++passes;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "pass launch");
    if ((e = cudaMemcpyAsync(&active, d_active, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return bail(e, "copy");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail(e, "sync");
  }

  std::vector<PlattState> host(K);
  if ((e = cudaMemcpyAsync(host.data(), d_state, (size_t)K * sizeof(PlattState), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
    return bail(e, "copy");
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail(e, "sync");
  release();
  for (int k = 0; k < K; ++k) {
    a_out[k] = host[k].A;
    b_out[k] = host[k].B;
    if (loss_out) loss_out[k] = host[k].fval;
  }
  if (passes_out) *passes_out = passes;
  return MC_OK;
}

}  // extern "C"
