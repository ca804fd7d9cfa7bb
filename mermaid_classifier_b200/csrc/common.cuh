// Shared helpers for libmermaid_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/mermaid_b200.h"

namespace mc {

// ---- error plumbing -------------------------------------------------------------
inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const std::string& msg) {
  last_error() = msg;
  return code;
}
#define MC_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::mc::fail(MC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
  } while (0)
#define MC_CHECK_LAUNCH() MC_CUDA(cudaGetLastError())

// ---- math -----------------------------------------------------------------------
// backbone activations: fast intrinsics by name (the library is NOT built with --use_fast_math)
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// Experiment (MC_SWISH_NR=1, off): swish of a pair with ONE MUFU op per value -- e = 2^(-x log2 e) on the XU, the reciprocal of
// 1 + e on the FMA pipe (integer seed, three Newton steps r <- r (2 - d r) as packed FFMA2 / FMUL2: 5 % -> 2.5e-3 -> 6.6e-6
// -> 1.2e-7 relative error).  Measured on B200 in the GEMM epilogue: expand layers 4-11 % SLOWER, although the XU pipe is
// the busiest unit of that epilogue (ncu 46-57 %) -- the extra nine FMA-pipe instructions per pair cost more issue slots
// than the MUFU.RCP they replace frees.  The exponent is clamped at 2^60 so that 1 + e stays finite.
#ifndef MC_SWISH_NR
#define MC_SWISH_NR 0
#endif
__device__ __forceinline__ float2 silu2_nr(float2 x) {
  float ex, ey;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fminf(x.x * -1.4426950408889634f, 60.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(fminf(x.y * -1.4426950408889634f, 60.f)));
  const float2 d = make_float2(1.f + ex, 1.f + ey);
  const float2 nd = make_float2(-d.x, -d.y);
  float2 r = make_float2(__int_as_float(0x7EF311C7 - __float_as_int(d.x)), __int_as_float(0x7EF311C7 - __float_as_int(d.y)));
  const float2 two = make_float2(2.f, 2.f);
#pragma unroll
  for (int it = 0; it < 3; ++it) r = __fmul2_rn(r, __ffma2_rn(nd, r, two));
  return __fmul2_rn(x, r);
}
// swish of bn = acc * scale + bias in the activation mode of T.
//   fp32 mode: exact form, 2 MUFU ops (ex2, rcp).
//   bf16 mode: x * sigmoid(x) = h + h * tanh(h) with h = x / 2 and ONE MUFU op (tanh.approx, rel. error 2^-11,
//   below the bf16 rounding of the stored result); the 1/2 is folded into the BN FMA.
template <typename T>
__device__ __forceinline__ float bn_silu(float acc, float scale, float bias);
template <>
__device__ __forceinline__ float bn_silu<float>(float acc, float scale, float bias) {
  return silu_f(fmaf(acc, scale, bias));
}
template <>
__device__ __forceinline__ float bn_silu<__nv_bfloat16>(float acc, float scale, float bias) {
  const float h = fmaf(acc, 0.5f * scale, 0.5f * bias);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// ---- 16-byte activation vectors ---------------------------------------------------
template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ static Vec load(const float* p) {
    Vec r;
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ static Vec load(const __nv_bfloat16* p) {
    Vec r;
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      r.v[2 * i] = f.x;
      r.v[2 * i + 1] = f.y;
    }
    return r;
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// Four channels as loaded from memory (kept packed while in flight).
template <typename T>
struct RawVec4;
template <>
struct RawVec4<float> {
  typedef float4 type;
  __device__ __forceinline__ static float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ static float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ static void unpack(const float4& r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
};
template <>
struct RawVec4<__nv_bfloat16> {
  typedef uint2 type;
  __device__ __forceinline__ static uint2 load(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ static uint2 zero() { return make_uint2(0u, 0u); }
  __device__ __forceinline__ static void unpack(const uint2& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x << 16);
    v[1] = __uint_as_float(r.x & 0xFFFF0000u);
    v[2] = __uint_as_float(r.y << 16);
    v[3] = __uint_as_float(r.y & 0xFFFF0000u);
  }
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ float from_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// 3xTF32 split of an fp32 value a (bit pattern `bits`): hi = a truncated to TF32 (what the tensor core reads when it is
// handed the raw fp32: it ignores the low 13 mantissa bits), lo = a - hi, exact in fp32 with up to 13 significant bits.
// The tensor core would TRUNCATE lo to TF32's 11 bits as well -- a bias toward zero of up to 2^-21 |a| on every product;
// rounding lo to nearest TF32 here halves that error and removes its sign.
#ifndef MC_TF32_LO_RN
#define MC_TF32_LO_RN 1
#endif
__host__ __device__ __forceinline__ uint32_t tf32_lo_bits(uint32_t bits) {
#ifdef __CUDA_ARCH__
  const float lo = __uint_as_float(bits) - __uint_as_float(bits & 0xFFFFE000u);
  uint32_t lb = __float_as_uint(lo);
#else
  float a, h;
  const uint32_t hb = bits & 0xFFFFE000u;
  memcpy(&a, &bits, 4);
  memcpy(&h, &hb, 4);
  const float lo = a - h;
  uint32_t lb;
  memcpy(&lb, &lo, 4);
#endif
#if MC_TF32_LO_RN
  lb = (lb + 0x1000u) & 0xFFFFE000u;
#endif
  return lb;
}

// NumPy mode='reflect' index (periodic mirror without repeating the edge sample).
__device__ __forceinline__ int reflect_idx(int t, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  t %= period;
  if (t < 0) t += period;
  return t >= n ? period - t : t;
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Function attributes (max dynamic shared memory) are per DEVICE: a `static bool` would set them on the first device
// of the process only.  `mask` is a function-local static; true exactly once per (call site, current device).
inline bool first_use_on_device(std::atomic<unsigned long long>& mask) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  return (mask.fetch_or(bit) & bit) == 0;
}

// Programmatic dependent launch -- an EXPERIMENT, off by default (MC_PDL_MASK=<family bits> turns it on).  The backbone is a
// chain of ~65 dependent launches per sub-batch, and between two of them the GPU drains the tail of one grid, launches the
// next and runs its prologue (barrier init, TMEM allocation, resident weights into shared memory, scale / bias) before any
// useful byte moves.  Kernels launched through launch_pdl with their family enabled may start while their predecessor in the
// stream is still running; each does its input-independent prologue, then pdl_wait() -- which returns once the predecessor
// grid has COMPLETED and its writes are visible -- before it touches anything a previous kernel reads or writes;
// pdl_trigger() lets the next kernel be scheduled as soon as this grid's CTAs are all resident.
// Measured on B200 (200 images x 100 points, same run): fp32 79.5 -> 79.7 k patches/s (noise), bf16 115.8 -> 117.4 k (+1.4 %)
// with the GEMM / SE / stem / fused families; the MLP training step gets 3x SLOWER (2.3 k for 6.7 k Adam steps/s: ten 5-25 us
// kernels per step, the attributed launches cost more host time than the step has).  And with dw_reg_kernel enabled the
// bit-reproducibility test (the same patches in reversed order, test_full_size_sub_batch_properties) FAILS, every time,
// although every access of that kernel sits behind its wait.  Narrowed down: it only fails when the predecessor GEMM runs with
// the per-n-block weight residency, i.e. on a grid of 144-147 CTAs that leaves SMs idle, so that the depthwise CTAs are
// resident (and waiting) from the start of the GEMM instead of from its tail; waiting in every thread, __threadfence and
// fence.proxy.async after the wait change nothing.  Unexplained, so nothing is enabled by default.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// kernel families (bits of MC_PDL_MASK, default all): bisecting aid
enum { PDL_GEMM = 1, PDL_DW = 2, PDL_SE = 4, PDL_STEM = 8, PDL_FUSED = 16, PDL_DWREG = 64 };
inline bool pdl_enabled(int family) {
  static const int mask = getenv("MC_PDL_MASK") ? (int)strtol(getenv("MC_PDL_MASK"), nullptr, 0) : 0;
  return (mask & family) != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int family, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(family) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mc
