// Per-class Platt scaling (sigmoid calibration) of an (n x K) probability matrix on the device.
//
// What it replaces: MermaidTrainer._calibrate_in_batches
// (mermaid_classifier/pyspacer/trainer.py:344-396) -> sklearn.calibration._fit_calibrator(method="sigmoid")
// -> one _sigmoid_calibration(F = proba[:, k], y = (label == k)) per class (scikit-learn 1.5.2,
// sklearn/calibration.py): Platt targets T+ = (N+ + 1)/(N+ + 2), T- = 1/(N- + 2), start
// AB0 = (0, log((N- + 1)/(N+ + 1))), minimise sum_i [log(1 + exp(r_i)) - T_i r_i] with
// r_i = -(A F_i + B).  sklearn hands that convex 2-parameter problem to L-BFGS-B; here all K
// problems advance together with Newton steps and a backtracking line search (Lin, Lin & Weng 2007):
// every pass streams the matrix once and produces, per class, the objective, gradient and Hessian
// at that class's trial point.
//
// HBM-bound by design (8 B per element per pass, coalesced across classes); fixed launch shape and
// fixed-order partial sums make the result bit-reproducible.
#pragma once
#include "common.cuh"

namespace mc {

constexpr int PLATT_SLICES = 296;  // row slices: two CTAs per SM for each 128-class column block
constexpr int PLATT_TERMS = 6;     // L, gA, gB, hAA, hAB, hBB

struct PlattState {
  double A, B, fval;     // accepted point and its objective
  double tA, tB;         // trial point evaluated by the next pass
  double dA, dB, gd;     // current Newton direction and its slope g.d
  double step;
  double t_pos, t_neg;   // Platt targets
  int done, accepted, passes, first;
};

__global__ void platt_count_kernel(const int32_t* __restrict__ y, int64_t n, int K, unsigned long long* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = y[i];
    if (t >= 0 && t < K) atomicAdd(&cnt[t], 1ull);  // integer adds: order-independent
  }
}

__global__ void platt_init_kernel(const unsigned long long* __restrict__ cnt, int64_t n, int K, PlattState* __restrict__ st) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double prior1 = (double)cnt[k], prior0 = (double)n - prior1;
  PlattState s;
  s.t_pos = (prior1 + 1.0) / (prior1 + 2.0);
  s.t_neg = 1.0 / (prior0 + 2.0);
  s.A = 0.0;
  s.B = log((prior0 + 1.0) / (prior1 + 1.0));
  s.tA = s.A;
  s.tB = s.B;
  s.fval = 0.0;
  s.dA = s.dB = s.gd = 0.0;
  s.step = 1.0;
  s.done = 0;
  s.accepted = 0;
  s.passes = 0;
  s.first = 1;
  st[k] = s;
}

// grid (ceil(K / 128), PLATT_SLICES), 128 threads: thread = class column, CTA row slice i = blockIdx.y (mod PLATT_SLICES).
__global__ void __launch_bounds__(128) platt_pass_kernel(const double* __restrict__ proba, const int32_t* __restrict__ y,
                                                         int64_t n, int K, const PlattState* __restrict__ st,
                                                         double* __restrict__ part /* [slice][term][K] */) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= K) return;
  const PlattState s = st[k];
  if (s.done) return;
  double L = 0, gA = 0, gB = 0, hAA = 0, hAB = 0, hBB = 0;
  for (int64_t i = blockIdx.y; i < n; i += PLATT_SLICES) {
    const double f = proba[i * K + k];
    const double t = (y[i] == k) ? s.t_pos : s.t_neg;
    const double r = -(s.tA * f + s.tB);
    const double e = exp(-fabs(r));
    const double p = (r >= 0.0 ? 1.0 : e) / (1.0 + e);  // sigmoid(r)
    L += log1p(e) + fmax(r, 0.0) - t * r;               // log(1 + exp(r)) - t r
    const double g = p - t, h = p * (1.0 - p);
    gA -= g * f;
    gB -= g;
    hAA += h * f * f;
    hAB += h * f;
    hBB += h;
  }
  double* o = part + (size_t)blockIdx.y * PLATT_TERMS * K + k;
  o[0 * K] = L;
  o[1 * K] = gA;
  o[2 * K] = gB;
  o[3 * K] = hAA;
  o[4 * K] = hAB;
  o[5 * K] = hBB;
}

__global__ void platt_update_kernel(const double* __restrict__ part, int K, PlattState* __restrict__ st, double gtol,
                                    int max_accepted, int* __restrict__ n_active) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  PlattState s = st[k];
  if (s.done) return;
  double v[PLATT_TERMS] = {0, 0, 0, 0, 0, 0};
  for (int sl = 0; sl < PLATT_SLICES; ++sl)
#pragma unroll
    for (int t = 0; t < PLATT_TERMS; ++t) v[t] += part[((size_t)sl * PLATT_TERMS + t) * K + k];
  s.passes++;
  const double L = v[0], gA = v[1], gB = v[2];
  const bool accept = s.first || L < s.fval + 1e-4 * s.step * s.gd;
  if (accept) {
    s.first = 0;
    s.A = s.tA;
    s.B = s.tB;
    s.fval = L;
    s.accepted++;
    if ((fabs(gA) < gtol && fabs(gB) < gtol) || s.accepted > max_accepted) {
      s.done = 1;
    } else {
      // Newton direction with a tiny ridge (Lin et al.: H + sigma I, sigma = 1e-12)
      const double hAA = v[3] + 1e-12, hAB = v[4], hBB = v[5] + 1e-12;
      const double det = hAA * hBB - hAB * hAB;
      s.dA = -(hBB * gA - hAB * gB) / det;
      s.dB = -(-hAB * gA + hAA * gB) / det;
      s.gd = gA * s.dA + gB * s.dB;
      s.step = 1.0;
      if (!(s.gd < 0.0) || !isfinite(s.dA) || !isfinite(s.dB)) {
        s.done = 1;  // no descent left at this precision
      } else if (-s.gd < 1e-11 * fmax(1.0, fabs(s.fval))) {
        // the predicted decrease is below what the summed objective can resolve: take the full Newton step as a
        // final polish (quadratic convergence: the gradient drops to rounding level) instead of line-searching noise
        s.A += s.dA;
        s.B += s.dB;
        s.done = 1;
      }
    }
  } else {
    s.step *= 0.5;
    if (s.step < 1e-10) s.done = 1;  // line search exhausted: keep the accepted point
  }
  if (!s.done) {
    s.tA = s.A + s.step * s.dA;
    s.tB = s.B + s.step * s.dB;
    atomicAdd(n_active, 1);
  }
  st[k] = s;
}

}  // namespace mc
