// Pointwise (1x1 conv) GEMM on the 5th-generation tensor cores: tcgen05.mma with TMEM
// accumulators, operands staged by TMA into 128B-swizzled shared memory.
//
//   out[m][n] = act( (sum_k A[m][k] * gate[m / HW][k] * W[n][k]) * scale[n] + bias[n] ) + res[m][n]
//
// A is the NHWC activation matrix (M = patches*H*W rows, K = C_in, K-major), W the conv weight
// (N = C_out rows, K-major).  One persistent CTA per SM walks (m-tile, n-block) work items:
//
//   warp 0      TMA producer: A [128 x KC] and W [BN x KC] boxes into a ring of stages
//   warp 1      MMA issuer (one elected lane): tcgen05.mma.cta_group::1, M = 128, N = BN,
//               accumulating over the K chunks into one of two TMEM accumulator stages
//   warps 2-5   operand transform (only when needed): SE gate applied to the A tile in shared
//               memory (project convs) and, in fp32 mode, the 3xTF32 split a = hi + lo
//   warps 6-13  epilogue, two groups of four warps alternating over the TMEM stages:
//               tcgen05.ld -> BN-fold scale/bias -> swish -> (+ residual) -> 16-byte stores
//
// bf16 mode: kind::f16 (bf16 x bf16 -> fp32), KC = 64.  fp32 mode: kind::tf32 with the
// 3-term split  A_hi W_hi + A_lo W_hi + A_hi W_lo  (error ~2^-21, fp32-class), KC = 32.
//
// Every mbarrier wait carries a clock-based timeout that traps, so a protocol bug faults the
// launch instead of hanging the GPU.
#pragma once
#include <cuda.h>

#include <vector>

#include "common.cuh"
#include "layers.h"
#include "pw_simt.cuh"

#ifndef MC_TC_TIMING
#define MC_TC_TIMING 0   // 1: per-role wait/total cycle counters (MC_TC_DBG=<layer>); costs registers, off in production
#endif
#ifndef MC_BF16_TANH
#define MC_BF16_TANH 1   // bf16 mode: swish with one MUFU op (tanh.approx) instead of two (ex2 + rcp)
#endif

namespace mc {

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~4 s at 2 GHz, then trap (a faulted launch, never a hung GPU).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();
  }
}

// wait + cycles spent waiting (role-level pipeline diagnosis, MC_TC_DBG)
__device__ __forceinline__ long long mbar_wait_timed(uint64_t* bar, uint32_t parity) {
#if MC_TC_TIMING
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  return clock64() - t0;
#else
  mbar_wait(bar, parity);
  return 0;
#endif
}
__device__ __forceinline__ long long tc_clock() {
#if MC_TC_TIMING
  return clock64();
#else
  return 0;
#endif
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]
template <bool TF32>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 rows = lanes, K 32-bit columns) is read from tensor memory, so an
// MMA pulls only its B tile through the shared-memory pipe
__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same for 16-bit operands: two consecutive K elements per 32-bit TMEM column (K = 16 per MMA = 8 columns)
__device__ __forceinline__ void mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: thread t of the warp writes lane (base + t), 16 consecutive 32-bit columns.  No wait inside.
__device__ __forceinline__ void tmem_st32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Convergent-warp forms: the WHOLE warp executes these and one elected lane issues.  Under a divergent
// `if (lane == 0)` ptxas cannot prove the uniform-register operands warp-uniform and wraps every tcgen05 instruction in an
// elect / broadcast / branch waterfall (6 extra instructions each, all on the slow uniform datapath): measured ~150
// cycles of issue work per MMA, more than a 128 x 128 x 8 TF32 MMA takes to execute.
// a_lo / b_lo: low descriptor words ((smem address >> 4) | LBO << 16); hi: the constant high word.
template <bool TF32>
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                             uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p, el;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "elect.sync _|el, 0xffffffff;\n\t"
        "@el tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, el;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "elect.sync _|el, 0xffffffff;\n\t"
        "@el tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred el;\n\telect.sync _|el, 0xffffffff;\n\t"
      "@el tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base + t), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"  // same asm statement: the outputs are only defined after the wait
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns (thread t: lane base + t, columns c..c+15).  No wait inside.
__device__ __forceinline__ void tmem_ld32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bits, repeated 4 times (32 fp32 columns): the mma-style fragment layout.  For
// repetition i (columns 8i .. 8i+7) thread t holds  v[4i+0..1] = (lane t/4,     columns 8i + 2(t%4) + {0,1})
//                                                   v[4i+2..3] = (lane t/4 + 8, same columns).
// No wait inside: issue several, then tmem_ld_wait() once.
__device__ __forceinline__ void tmem_ld16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace ptx

// four consecutive residual values as loaded from memory (kept packed while in flight)
template <typename T>
struct ResVec;
template <>
struct ResVec<float> {
  typedef float4 type;
  __device__ __forceinline__ static float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ static float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ static void unpack(const float4& v, float (&r)[4]) { r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
};
template <>
struct ResVec<__nv_bfloat16> {
  typedef uint2 type;
  __device__ __forceinline__ static uint2 zero() { return make_uint2(0u, 0u); }
  __device__ __forceinline__ static uint2 load(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ static void unpack(const uint2& v, float (&r)[4]) {
    r[0] = __uint_as_float(v.x << 16); r[1] = __uint_as_float(v.x & 0xFFFF0000u);
    r[2] = __uint_as_float(v.y << 16); r[3] = __uint_as_float(v.y & 0xFFFF0000u);
  }
};

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row (1024 B) swizzle atoms stacked
// along M/N.  start address >> 4 | LBO(ignored)=1 | SBO = 1024 >> 4 | version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
constexpr int TC_EPI_GROUPS = 4;                       // epilogue warp groups (4 warps each)
// warps per CTA: TMA + MMA + transform + epilogue groups of four.  Ungated layers (expand, head conv): 4 transform
// warps + 4 epilogue groups = 22 warps (80 registers); gated layers (project): 8 transform warps + 2 epilogue groups =
// 18 warps, which leaves 112 registers per thread for the gate double-buffer.
template <bool GATED>
__host__ __device__ constexpr int tc_threads() { return GATED ? (2 + 8 + 4 * 2) * 32 : (2 + 4 + 4 * 4) * 32; }
constexpr int TC_BM = 128;
constexpr int TC_ACC_COLS = 256;  // TMEM columns per accumulator stage (2 stages = 512)

struct PwTcArgs {
  const float* scale;
  const float* bias;
  const void* gate;   // [n_img][K] (fp32 in fp32 mode, bf16 in bf16 mode) or null
  const void* res;    // [M][N] or null
  void* out;          // [M][N]
  int64_t M;
  int N, K, HW, BN, n_blocks, k_chunks, act, stages;
  long long* dbg;     // MC_TC_DBG: per-role wait/total cycle counters of CTA 0 (null = off)
  int exp_flags;      // MC_TC_EXP experiments: 1 no activation, 2 no global stores, 4 no operand transform, 16 no MMA, 32 no tcgen05.ld,
                      // 64 no W lo fetch, 256 single-pass TF32 in the TS form (hi x hi only: a precision experiment)
  int w_res;          // 1: the layer's whole weight (all n-blocks x k-chunks, hi+lo) stays resident in smem
                      // 2: ONE n-block's weight (all k-chunks) stays resident: the grid is a multiple of n_blocks, so CTA c only
                      //    ever sees n-block c % n_blocks (item it = c + j * grid -> n-block it % n_blocks)
  int a_row_off;      // first row of this launch inside the activation tensor map (chunked execution)
  int64_t m_tiles;
  int pool_nb;        // POOL instantiation: patches in this launch (a work item's 128-row tile = two 49-row patches)
};

// POOL instantiation (head conv + global average pool, K7): rows of a patch's 7x7 map per tile half, and the shared
// memory the four epilogue groups use to combine the column sums of their warps
constexpr int TC_POOL_HW = 49;
constexpr int TC_POOL_BYTES = TC_EPI_GROUPS * 4 * 128 * 4;

template <typename T>
struct TcCfg;
template <>
struct TcCfg<__nv_bfloat16> {
  static constexpr bool TF32 = false;
  static constexpr int KC = 64;       // elements per 128-byte row
  static constexpr int UK = 16;       // K per MMA
  static constexpr int BN_MAX = 128;   // <= 128 so four accumulator stages fit in TMEM (one per epilogue group)
  static constexpr int A_BYTES = TC_BM * 128;
  static constexpr int NA = 1, NW = 1;  // A / W operand copies per stage
  static constexpr int EPI_COLS = 64;   // output columns staged per epilogue pass (128 bytes per row)
  static constexpr uint32_t FMT = 1;    // BF16
};
template <>
struct TcCfg<float> {
  static constexpr bool TF32 = true;
  static constexpr int KC = 32;
  static constexpr int UK = 8;
  static constexpr int BN_MAX = 128;
  static constexpr int A_BYTES = TC_BM * 128;
  static constexpr int NA = 2, NW = 2;  // A_hi, A_lo, W_hi, W_lo
  static constexpr int EPI_COLS = 32;
  static constexpr uint32_t FMT = 2;    // TF32
};

constexpr int TC_MAX_STAGES = 8;
constexpr int TC_SMEM_BUDGET = 227 * 1024;
constexpr int TC_EPI_BYTES = 0;                                                  // the epilogue stores straight from registers
constexpr int TC_GATE_ROWS = 4;                                                   // patches a 128-row tile can span (7x7 maps: 49 rows each)
constexpr int TC_GATE_BYTES = TC_MAX_STAGES * TC_GATE_ROWS * 128;                 // SE gate chunks of every stage (project layers)
constexpr int TC_FIXED_BYTES = 1024 /*align slack*/ + TC_EPI_BYTES + TC_GATE_BYTES + 2 * 1280 * 4 + 512;

// bytes of one pipeline stage for a layer with BN output columns per block (1024-aligned)
// The TF32 lo operand of A is produced by the transform warps, not by TMA: it lives in its own two-slot ring, so the TMA
// ring holds one A tile (+ the W tiles) per stage and is a stage deeper for the same shared memory.  With hi + lo of A AND W in
// every stage the late layers fitted three stages; their time per k-chunk follows (TMA latency + transform + MMA) / stages
// (measured: removing the MMAs altogether only took b15.project from 0.242 to 0.163 ms), so depth is what they lack.
template <typename T>
constexpr int tc_lo_ring_bytes() { return TcCfg<T>::TF32 ? 2 * TcCfg<T>::A_BYTES : 0; }
template <typename T>
constexpr int tc_stage_bytes(int BN) {
  return TcCfg<T>::A_BYTES + TcCfg<T>::NW * ((BN * 128 + 1023) / 1024 * 1024);
}
template <typename T>
inline int tc_num_stages(int BN, int extra_fixed = 0, bool no_lo_ring = false) {
  int s = (TC_SMEM_BUDGET - TC_FIXED_BYTES - extra_fixed - (no_lo_ring ? 0 : tc_lo_ring_bytes<T>())) / tc_stage_bytes<T>(BN);
  return s > TC_MAX_STAGES ? TC_MAX_STAGES : s;
}
// resident-weight layout: W region of n_blocks * k_chunks * NW tiles, stages carry the A operands only
constexpr int TC_W_RES_MAX = 96 * 1024;
template <typename T>
inline int tc_w_res_bytes(int BN, int n_blocks, int k_chunks) {
  return n_blocks * k_chunks * TcCfg<T>::NW * ((BN * 128 + 1023) / 1024 * 1024);
}
template <typename T>
inline int tc_num_stages_res(int w_bytes, bool no_lo_ring = false) {
  int s = (TC_SMEM_BUDGET - TC_FIXED_BYTES - w_bytes - (no_lo_ring ? 0 : tc_lo_ring_bytes<T>())) / TcCfg<T>::A_BYTES;
  return s > TC_MAX_STAGES ? TC_MAX_STAGES : s;
}

// RELU: a separate instantiation for the MLP-head Linear layers.  The conv instantiations must not even carry the
// (never taken) ReLU branch: b1.expand sits on a scheduling knife-edge between 4.7 and 3.1 TB/s, and that branch alone
// tipped it (measured, round 1).
// POOL: the head conv with the global average pool in its epilogue (K7).  A work item's 128-row tile holds TWO patches:
// rows 0..48 = the 7x7 map of patch 2t, rows 64..112 = patch 2t+1 (two TMA boxes of 49 rows; a patch starts on a warp
// boundary of the epilogue, the other rows of the tile are never read back).  The epilogue applies BN + swish, sums each
// column over the patch's 49 rows in a fixed order (in-thread, lane shuffles, then the two warps of the patch through
// shared memory) and writes mean features [patch][N] in fp32: the 49 x 1280 map per patch never exists in global memory.
// TS (gated fp32 layers): the transform warps write the gated TF32 hi / lo operands of A into TENSOR MEMORY (tcgen05.st, a
// thread = a tile row = a TMEM lane) instead of back into shared memory, and the MMAs take A from there.  With both
// operands in shared memory the K >= 480 project layers were bound by the shared-memory pipe (128 B/clk per SM), not by the
// tensor cores: per 128 x 96 x 8 MMA 4 KB of A + 3 KB of W operand reads, three MMAs per k-step, plus the TMA fills (40 KB
// per k-chunk) and the transform's 16 KB of loads and 32 KB of stores -- ~115 smem-cycles per MMA against 49 tensor-cycles
// (ncu: tensor pipe 41 % on b15.project; role timers: the issuer busy 75 %, no barrier waits).  A from TMEM removes the A
// operand reads and the transform's stores: ~60 smem-cycles per MMA.  TMEM: accumulators in columns [0, 256) (two stages of
// 128, or four of 64 when BN <= 64), A ring of four slots x (32 hi + 32 lo columns) in [256, 512).
template <typename T, bool GATED, bool RELU = false, bool POOL = false, bool TS = false>
__global__ void __launch_bounds__(tc_threads<GATED>(), 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CUtensorMap tmG, const PwTcArgs p) {
  using Cfg = TcCfg<T>;
  const int S = p.stages;
  const int W_BYTES = (p.BN * 128 + 1023) / 1024 * 1024;       // one W operand tile (1024-aligned)
  // Weight residency: when the whole layer weight fits, it is loaded ONCE per CTA and the ring carries activations only.
  // Re-fetching the same W tile for every work item made 148 CTAs hammer a handful of L2 lines: throughput of the
  // front layers then depended on which L2 slices the weight allocation happened to hash to (4.8 vs 3.2 TB/s on b1.expand).
  // Past the front layers the whole weight no longer fits, and streaming W hi + lo tiles with every A tile made the expand
  // layers L2 -> shared-memory fill-bound (b9.expand: 160 KB of fills per 128 x 96 output tile, 60 % of it weights; the chip's
  // L2 delivers ~6300 B/clk in total).  Mode 2 pins each CTA to one n-block and keeps that block's weight resident.
  const bool w_res = p.w_res != 0;
  const bool w_res_nb = p.w_res == 2;
  const int W_RES_BYTES = w_res ? (w_res_nb ? 1 : p.n_blocks) * p.k_chunks * Cfg::NW * W_BYTES : 0;
  const int STAGE_BYTES = Cfg::A_BYTES + (w_res ? 0 : Cfg::NW * W_BYTES);   // A (hi) tile [+ W hi, W lo tiles]
  static_assert(!TS || Cfg::TF32 || GATED, "TS: fp32 layers, and the gated bf16 ones (an ungated bf16 layer has no transform at all)");
  constexpr int LO_SLOTS = (Cfg::TF32 && !TS) ? 2 : 0;                       // ring of the TF32 lo operand of A (shared memory)
  // TS: ring of the A operands in tensor memory.  Gated (two epilogue groups): accumulators in columns [0, 256), four slots
  // in [256, 512).  Ungated (four groups, one 96-column accumulator each, BN <= 96): [0, 384), two slots in [384, 512).
  constexpr uint32_t TS_A_COL0 = GATED ? 256 : 384, TS_SLOTS = GATED ? 4 : 2, TS_SLOT_COLS = 64;
  constexpr uint32_t TS_SLOT_SHIFT = GATED ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* w_base = smem;                                      // resident weights (1024-aligned tiles), may be empty
  uint8_t* stage_base = smem + W_RES_BYTES;
  uint8_t* lo_base = stage_base + (size_t)S * STAGE_BYTES;
  uint8_t* epi_base = lo_base + LO_SLOTS * Cfg::A_BYTES;
  uint8_t* gate_s = epi_base;                                  // GATED: [stage][4 patches][128 B] SE gate chunk of the stage's k-chunk
  float* pool_s = (float*)(epi_base + TC_GATE_BYTES);          // POOL: [group][warp of the group][128 columns]
  float* sc_s = (float*)(epi_base + TC_EPI_BYTES + TC_GATE_BYTES + (POOL ? TC_POOL_BYTES : 0));
  float* bi_s = sc_s + 1280;
  uint64_t* bars = (uint64_t*)(bi_s + 1280);
  uint64_t* full = bars;                          // [S]   TMA landed
  uint64_t* ready = bars + TC_MAX_STAGES;         // [S]   transform done (when a transform runs)
  uint64_t* empty = bars + 2 * TC_MAX_STAGES;     // [S]   MMAs reading the stage retired
  uint64_t* tfull = bars + 3 * TC_MAX_STAGES;     // [4]   accumulator complete
  uint64_t* tempty = tfull + 4;                   // [4]   accumulator drained
  uint64_t* wbar = tempty + 4;                    // resident weights landed
  uint64_t* lo_empty = wbar + 1;                  // [2]   MMAs reading the lo slot retired ([4] TS: the TMEM A slot)
  uint32_t* tmem_slot = (uint32_t*)(lo_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // accumulator stages in TMEM: 4 x 128 columns when the block fits, else 2 x 256
  // warp roles: [0, ntw) transform, then the epilogue groups of four warps, and the two single-thread roles LAST
  // (MMA issuer, TMA producer): the SM's warp arbiter favours the highest warp id on each sub-partition, and
  // these two are the critical path -- as warps 0/1 they were starved by the busy transform/epilogue warps.
  constexpr bool gated_layer = GATED;
  constexpr int TC_THREADS = tc_threads<GATED>();
  constexpr int ntw = GATED ? 8 : 4;
  constexpr int n_groups = (TC_THREADS / 32 - 2 - ntw) / 4;   // 4 ungated, 2 gated
  constexpr int WARP_MMA = TC_THREADS / 32 - 2, WARP_TMA = TC_THREADS / 32 - 1;
  // Accumulator stages of 128 TMEM columns each (BN <= 128): the four 128-column blocks are shared out evenly, so a group of
  // the gated kernel (2 groups) owns TWO stages and the MMAs of its next item overlap the epilogue of the current one; the
  // ungated kernel (4 groups) has one stage per group.  Item li -> group li % n_groups, that group's (li / n_groups)-th item
  // -> stage group + n_groups * (j % ACC_DEPTH), use j / ACC_DEPTH.  A stage is only ever waited on by its own group, which
  // sees every one of its phases (a parity wait is only meaningful for the current or the immediately preceding phase).
  constexpr int NAS = 4;
  const int ACC_DEPTH = TS ? ((GATED && p.BN <= 64) ? 2 : 1) : NAS / n_groups;
  const int acc_cols = TS ? (GATED ? (p.BN <= 64 ? 64 : 128) : 96) : 128;
  constexpr bool transform = Cfg::TF32 || GATED;
  // 32-bit work-item arithmetic throughout: a 64-bit divide by a run-time value is a ~100-instruction
  // subroutine, and every role used to pay several of them per item.
  const int items = (int)(p.m_tiles * p.n_blocks);
  const uint32_t nblk = (uint32_t)p.n_blocks;

  // bf16 swish uses x*sigmoid(x) = h + h*tanh(h) with h = x/2: fold the 1/2 into scale and bias
  const float fold = (MC_BF16_TANH && !Cfg::TF32 && p.act == 1) ? 0.5f : 1.f;
  for (int i = threadIdx.x; i < p.N; i += TC_THREADS) {
    sc_s[i] = p.scale[i] * fold;
    bi_s[i] = p.bias[i] * fold;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_MAX_STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&ready[s], (uint32_t)(ntw * 32));
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(wbar, 1);
    for (int s = 0; s < 4; ++s) ptx::mbar_init(&lo_empty[s], 1);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], 128);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    if (Cfg::TF32) ptx::prefetch_tmap(&tmWlo);
  }
  if (warp == WARP_MMA) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above is independent of the previous kernel.  The producer still loads the
  // resident weights (constants) before it waits; every other role waits here.
  pdl_trigger();
  if (warp != WARP_TMA) pdl_wait();

  if (warp == WARP_TMA) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t w_tile_tx = (uint32_t)p.BN * 128u;   // bytes one W box delivers
      const bool skip_wlo = (p.exp_flags & 64) != 0;   // timing experiment: do not fetch the W lo tile (results wrong by design)
      const uint32_t tx = Cfg::A_BYTES + (w_res ? 0u : w_tile_tx * ((Cfg::TF32 && !skip_wlo) ? 2u : 1u));
      if (w_res) {
        const int nb_lo = w_res_nb ? (int)(blockIdx.x % nblk) : 0, nb_hi = w_res_nb ? nb_lo + 1 : p.n_blocks;
        ptx::mbar_expect_tx(wbar, w_tile_tx * (uint32_t)((nb_hi - nb_lo) * p.k_chunks * Cfg::NW));
        for (int nb_ = nb_lo; nb_ < nb_hi; ++nb_)
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            uint8_t* wt = w_base + (size_t)(((nb_ - nb_lo) * p.k_chunks + kc) * Cfg::NW) * W_BYTES;
            ptx::tma_load_2d(wt, &tmW, wbar, kc * Cfg::KC, nb_ * p.BN);
            if (Cfg::TF32) ptx::tma_load_2d(wt + W_BYTES, &tmWlo, wbar, kc * Cfg::KC, nb_ * p.BN);
          }
      }
      pdl_wait();   // activations, gate: written by the previous kernels
      long long w_empty = 0;
      const long long t_begin = ptx::tc_clock();
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const uint32_t mt = (uint32_t)it / nblk;
        const int m0 = (int)mt * TC_BM + p.a_row_off;
        const int n0 = (int)((uint32_t)it - mt * nblk) * p.BN;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          w_empty += ptx::mbar_wait_timed(&empty[s], ph ^ 1);
          uint8_t* st = stage_base + (size_t)s * STAGE_BYTES;
          if constexpr (POOL) {
            // two patches: rows [0, 49) and [64, 113) of the tile (tmA's box is 49 rows)
            ptx::mbar_expect_tx(&full[s], tx - (uint32_t)Cfg::A_BYTES + 2u * TC_POOL_HW * 128u);
            ptx::tma_load_2d(st, &tmA, &full[s], kc * Cfg::KC, (int)mt * 2 * TC_POOL_HW);
            ptx::tma_load_2d(st + 64 * 128, &tmA, &full[s], kc * Cfg::KC, ((int)mt * 2 + 1) * TC_POOL_HW);
          } else if constexpr (GATED) {
            // + the SE gate of this k-chunk for the (up to four) patches the tile's rows belong to
            ptx::mbar_expect_tx(&full[s], tx + TC_GATE_ROWS * 128u);
            ptx::tma_load_2d(st, &tmA, &full[s], kc * Cfg::KC, m0);
            ptx::tma_load_2d(gate_s + s * (TC_GATE_ROWS * 128), &tmG, &full[s], kc * Cfg::KC, (int)((mt * (uint32_t)TC_BM) / (uint32_t)p.HW));
          } else {
            ptx::mbar_expect_tx(&full[s], tx);
            ptx::tma_load_2d(st, &tmA, &full[s], kc * Cfg::KC, m0);
          }
          if (!w_res) {
            if (Cfg::TF32) {
              ptx::tma_load_2d(st + Cfg::A_BYTES, &tmW, &full[s], kc * Cfg::KC, n0);
              if (!skip_wlo) ptx::tma_load_2d(st + Cfg::A_BYTES + W_BYTES, &tmWlo, &full[s], kc * Cfg::KC, n0);
            } else {
              ptx::tma_load_2d(st + Cfg::A_BYTES, &tmW, &full[s], kc * Cfg::KC, n0);
            }
          }
          if (++s == S) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      if (MC_TC_TIMING && p.dbg && blockIdx.x == 0) {
        p.dbg[0] = w_empty;
        p.dbg[1] = ptx::tc_clock() - t_begin;
      }
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer ==================================
    if (ptx::elect_one()) {
      const uint32_t idesc = (1u << 4) | (Cfg::FMT << 7) | (Cfg::FMT << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      constexpr uint64_t DESC_HI64 = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;   // SBO = 1024 >> 4, version 1, SWIZZLE_128B
      constexpr uint32_t DESC_LBO = 1u << 16;
      const uint32_t stage_lo = ptx::smem_u32(stage_base) >> 4, w_lo_base = ptx::smem_u32(w_base) >> 4;
      const uint32_t stage_step = (uint32_t)STAGE_BYTES >> 4, w_step = (uint32_t)W_BYTES >> 4;
      const uint32_t lo_ring_lo = ptx::smem_u32(lo_base) >> 4;
      uint32_t cc = 0;   // k-chunk counter of this CTA: lo slot cc & 1
      int s = 0;
      uint32_t ph = 0;
      int li = 0;
      long long w_tempty = 0, w_full = 0;
      [[maybe_unused]] long long t_commit = 0, t_issue = 0;   // MC_TC_TIMING: cycles the issuing thread spends in tcgen05.commit / tcgen05.mma
      const long long t_begin = ptx::tc_clock();
      if (w_res) {
        ptx::mbar_wait(wbar, 0);
        ptx::tc_fence_after();
      }
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++li) {
        const int gj = li / n_groups;
        const int as = li % n_groups + n_groups * (gj % ACC_DEPTH);
        const uint32_t use = (uint32_t)(gj / ACC_DEPTH);
        const int nb_i = w_res_nb ? 0 : (int)((uint32_t)it - ((uint32_t)it / nblk) * nblk);   // index into the resident weights
        w_tempty += ptx::mbar_wait_timed(&tempty[as], (use & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * acc_cols);
        for (int kc = 0; kc < p.k_chunks; ++kc, ++cc) {
          w_full += ptx::mbar_wait_timed(transform ? &ready[s] : &full[s], ph);
          ptx::tc_fence_after();
          // low descriptor words ((address >> 4) | LBO); every operand tile lives below 256 KB, so the 14-bit field never wraps
          const uint32_t a_lo = (stage_lo + (uint32_t)s * stage_step) | DESC_LBO;
          [[maybe_unused]] const uint32_t l_lo = (lo_ring_lo + (cc & 1u) * (uint32_t)(Cfg::A_BYTES >> 4)) | DESC_LBO;   // TF32 lo operand of A
          const uint32_t w_lo = w_res ? ((w_lo_base + (uint32_t)((nb_i * p.k_chunks + kc) * Cfg::NW) * w_step) | DESC_LBO)
                                      : a_lo + (uint32_t)(Cfg::A_BYTES >> 4);
          const int krem = p.K - kc * Cfg::KC;
          const int ksteps = (p.exp_flags & 16) ? 0 : (min(krem, Cfg::KC) + Cfg::UK - 1) / Cfg::UK;
          const long long t_i0 = ptx::tc_clock();
#pragma unroll
          for (int ks = 0; ks < Cfg::KC / Cfg::UK; ++ks) {
            if (ks < ksteps) {
              const uint32_t acc = (kc | ks) != 0;
              const uint32_t ko = (uint32_t)ks * 2u;   // UK elements = 32 bytes = 2 descriptor units
              if constexpr (TS && !Cfg::TF32) {
                const uint32_t a_tm = tmem_base + TS_A_COL0 + (cc & (TS_SLOTS - 1u)) * TS_SLOT_COLS + (uint32_t)ks * 8u;
                ptx::mma_ts_f16(d_tmem, a_tm, DESC_HI64 | (w_lo + ko), idesc, acc);
              } else if constexpr (TS) {
                const uint32_t a_tm = tmem_base + TS_A_COL0 + (cc & (TS_SLOTS - 1u)) * TS_SLOT_COLS + (uint32_t)ks * 8u;   // hi; lo 32 columns on
                const uint64_t whi = DESC_HI64 | (w_lo + ko), wlo = DESC_HI64 | (w_lo + w_step + ko);
                if (p.exp_flags & 256) {   // experiment: single-pass TF32 (hi x hi only)
                  ptx::mma_ts_tf32(d_tmem, a_tm, whi, idesc, acc);
                } else {
                  ptx::mma_ts_tf32(d_tmem, a_tm + 32u, whi, idesc, acc);
                  ptx::mma_ts_tf32(d_tmem, a_tm, wlo, idesc, 1u);
                  ptx::mma_ts_tf32(d_tmem, a_tm, whi, idesc, 1u);
                }
              } else if (Cfg::TF32) {
                const uint64_t ahi = DESC_HI64 | (a_lo + ko), alo = DESC_HI64 | (l_lo + ko);
                const uint64_t whi = DESC_HI64 | (w_lo + ko), wlo = DESC_HI64 | (w_lo + w_step + ko);
                ptx::mma_ss<true>(d_tmem, alo, whi, idesc, acc);
                ptx::mma_ss<true>(d_tmem, ahi, wlo, idesc, 1u);
                ptx::mma_ss<true>(d_tmem, ahi, whi, idesc, 1u);
              } else {
                ptx::mma_ss<false>(d_tmem, DESC_HI64 | (a_lo + ko), DESC_HI64 | (w_lo + ko), idesc, acc);
              }
            }
          }
          const long long t_c0 = ptx::tc_clock();
          t_issue += t_c0 - t_i0;
          ptx::mma_commit(&empty[s]);
          if (TS) ptx::mma_commit(&lo_empty[cc & (TS_SLOTS - 1u)]);
          else if (Cfg::TF32) ptx::mma_commit(&lo_empty[cc & 1u]);
          t_commit += ptx::tc_clock() - t_c0;
          if (++s == S) {
            s = 0;
            ph ^= 1;
          }
        }
        const long long t_c1 = ptx::tc_clock();
        ptx::mma_commit(&tfull[as]);
        t_commit += ptx::tc_clock() - t_c1;
      }
      if (MC_TC_TIMING && p.dbg && blockIdx.x == 0) {
        p.dbg[2] = w_tempty;
        p.dbg[3] = w_full;
        p.dbg[4] = ptx::tc_clock() - t_begin;
        p.dbg[5] = li;
        p.dbg[24] = t_commit;
        p.dbg[25] = t_issue;
      }
    }
  } else if (warp < ntw) {
    // ============================ operand transform ================================
    // Gated (project) layers: 8 warps, two threads per tile row, four 16-byte chunks each, and the gate
    // chunks of the NEXT k-chunk are fetched (L2, one round trip) while the current one is processed.
    // Ungated fp32 layers: 4 warps, one thread per row, only the TF32 lo operand is produced.
    if constexpr (GATED) {
      const int r2 = threadIdx.x;
      const int r = r2 & 127;              // tile row
      const int j0 = (r2 >> 7) * 4;        // first of this thread's four 16-byte chunks
      const uint32_t row_off = (uint32_t)r * 128u;
      const uint32_t xr = (uint32_t)(r & 7);
      constexpr int EPCH = 16 / (int)sizeof(T);    // elements per 16-byte chunk
      int s = 0;
      uint32_t ph = 0;
      // The gate chunk of the stage's k-chunk arrives with the A tile (TMA box of four gate rows: the patches the tile's
      // rows can belong to), so the transform reads it from shared memory: fetched from L2 by the threads themselves, one
      // step ahead, its ~700-cycle round trip was not hidden behind a ~300-cycle step (ncu: transform warps busy 92 %, the
      // MMA issuer starved, tensor pipe 41 % on b15.project).  All shared-memory loads of a step are issued before the first
      // dependent instruction (the inline-asm accesses keep program order).
      uint32_t cc = 0;   // k-chunk counter of this CTA: lo slot cc & 1
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const uint32_t mt = (uint32_t)it / nblk;
        const uint32_t m = mt * TC_BM + (uint32_t)r;
        const uint32_t prow = min((uint32_t)(TC_GATE_ROWS - 1), m / (uint32_t)p.HW - (mt * TC_BM) / (uint32_t)p.HW);   // rows past M: any patch (A is zero)
        for (int kc = 0; kc < p.k_chunks; ++kc, ++cc) {
          const int k0 = kc * Cfg::KC;
          const int nch = min(8, (p.K - k0) / EPCH);  // chunks that hold real data (the rest is TMA zero fill)
          ptx::mbar_wait(&full[s], ph);
          if (TS) ptx::mbar_wait(&lo_empty[cc & (TS_SLOTS - 1u)], ((cc >> TS_SLOT_SHIFT) & 1u) ^ 1u);   // the MMAs of four chunks ago are done with the TMEM slot
          else if (Cfg::TF32) ptx::mbar_wait(&lo_empty[cc & 1u], ((cc >> 1) & 1u) ^ 1u);   // the MMAs of two chunks ago are done with the slot
          const uint32_t a_hi = ptx::smem_u32(stage_base + (size_t)s * STAGE_BYTES) + row_off;
          [[maybe_unused]] const uint32_t a_lo_slot = ptx::smem_u32(lo_base + (cc & 1u) * Cfg::A_BYTES) + row_off;
          const uint32_t g_u32 = ptx::smem_u32(gate_s + s * (TC_GATE_ROWS * 128)) + prow * 128u + (uint32_t)j0 * 16u;
          uint4 raw[4], gq[4];
          if constexpr (TS && !Cfg::TF32) {
            // bf16: the thread's four 16-byte chunks = 32 gated values = 16 TMEM columns (element pairs) of its lane
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              raw[jj] = ptx::lds128(a_hi + (((uint32_t)(j0 + jj) ^ xr) << 4));
              gq[jj] = ptx::lds128(g_u32 + (uint32_t)jj * 16u);
            }
            uint32_t pk[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[jj]);
              const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gq[jj]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 m2 = __hmul2(h[e], gh[e]);
                pk[4 * jj + e] = *reinterpret_cast<const uint32_t*>(&m2);
              }
            }
            const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TS_A_COL0 + (cc & (TS_SLOTS - 1u)) * TS_SLOT_COLS +
                                   (uint32_t)j0 * 4u;
            ptx::tc_fence_after();
            ptx::tmem_st32x32b_x16(t_row, pk);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&ready[s]);
            if (++s == S) {
              s = 0;
              ph ^= 1;
            }
            continue;
          } else if constexpr (TS) {
            // row r = TMEM lane r (warp w reaches lanes 32 (w % 4) ..): this thread's 16 gated values as TF32 hi and lo into
            // columns 4 j0 .. 4 j0 + 15 of the slot's hi / lo blocks.  K tail: the TMA zero fill makes both zero.
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              raw[jj] = ptx::lds128(a_hi + (((uint32_t)(j0 + jj) ^ xr) << 4));
              gq[jj] = ptx::lds128(g_u32 + (uint32_t)jj * 16u);
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint32_t* rp = &raw[jj].x;
              const uint32_t* gp = &gq[jj].x;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t vb = __float_as_uint(__uint_as_float(rp[e]) * __uint_as_float(gp[e]));
                hi[4 * jj + e] = vb & 0xFFFFE000u;
                lo[4 * jj + e] = tf32_lo_bits(vb);
              }
            }
            const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TS_A_COL0 + (cc & (TS_SLOTS - 1u)) * TS_SLOT_COLS +
                                   (uint32_t)j0 * 4u;
            ptx::tc_fence_after();
            ptx::tmem_st32x32b_x16(t_row, hi);
            ptx::tmem_st32x32b_x16(t_row + 32u, lo);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&ready[s]);
            if (++s == S) {
              s = 0;
              ph ^= 1;
            }
            continue;
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            raw[jj] = make_uint4(0u, 0u, 0u, 0u);
            gq[jj] = raw[jj];
            if (j0 + jj < nch && !(p.exp_flags & 4)) {
              raw[jj] = ptx::lds128(a_hi + (((uint32_t)(j0 + jj) ^ xr) << 4));
              gq[jj] = ptx::lds128(g_u32 + (uint32_t)jj * 16u);
            }
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            const uint32_t phys = a_hi + (((uint32_t)j ^ xr) << 4);
            if (j < nch && !(p.exp_flags & 4)) {
              if (Cfg::TF32) {
                float v[4] = {__uint_as_float(raw[jj].x) * __uint_as_float(gq[jj].x), __uint_as_float(raw[jj].y) * __uint_as_float(gq[jj].y),
                              __uint_as_float(raw[jj].z) * __uint_as_float(gq[jj].z), __uint_as_float(raw[jj].w) * __uint_as_float(gq[jj].w)};
                uint4 hi, lo;
                uint32_t* hp = &hi.x;
                uint32_t* lp = &lo.x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  hp[e] = __float_as_uint(v[e]) & 0xFFFFE000u;   // exact TF32 value
                  lp[e] = tf32_lo_bits(__float_as_uint(v[e]));     // remainder, rounded to TF32
                }
                ptx::sts128(phys, hi);
                ptx::sts128(a_lo_slot + (((uint32_t)j ^ xr) << 4), lo);
              } else {
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw[jj]);
                const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gq[jj]);
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __hmul2(h[e], gh[e]);
                ptx::sts128(phys, raw[jj]);
              }
            } else if (Cfg::TF32 && j >= nch) {
              // zero-filled K tail: the lo copy must be zero too
              ptx::sts128(a_lo_slot + (((uint32_t)j ^ xr) << 4), make_uint4(0u, 0u, 0u, 0u));
            }
          }
          ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async) proxy
          ptx::mbar_arrive(&ready[s]);
          if (++s == S) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    } else if constexpr (Cfg::TF32) {
      // ungated fp32: one thread per tile row; the raw fp32 tile stays in place as the hi operand
      // (kind::tf32 reads only the upper 19 bits), only the lo operand a - tf32(a) is written.
      const int r = threadIdx.x;
      const uint32_t row_off = (uint32_t)r * 128u;
      const uint32_t xr = (uint32_t)(r & 7);
      int s = 0;
      uint32_t ph = 0;
      uint32_t cc = 0;   // k-chunk counter of this CTA: lo slot cc & 1
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        for (int kc = 0; kc < p.k_chunks; ++kc, ++cc) {
          const int nch = min(8, (p.K - kc * Cfg::KC) / 4);
          ptx::mbar_wait(&full[s], ph);
          ptx::mbar_wait(&lo_empty[cc & 1u], ((cc >> 1) & 1u) ^ 1u);   // the MMAs of two chunks ago are done with the slot
          const uint32_t a_hi = ptx::smem_u32(stage_base + (size_t)s * STAGE_BYTES) + row_off;
          if constexpr (TS) {
            // row r = TMEM lane r: the row's 32 values as TF32 hi / lo into the slot's two 32-column blocks, 16 at a time
            // (register budget).  The K tail is zero by the TMA fill.
            const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TS_A_COL0 + (cc & 1u) * TS_SLOT_COLS;
            ptx::tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint4 raw4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) raw4[j] = ptx::lds128(a_hi + (((uint32_t)(4 * half + j) ^ xr) << 4));
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t* rp = &raw4[j].x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  hi[4 * j + e] = rp[e] & 0xFFFFE000u;
                  lo[4 * j + e] = tf32_lo_bits(rp[e]);
                }
              }
              ptx::tmem_st32x32b_x16(t_row + 16u * half, hi);
              ptx::tmem_st32x32b_x16(t_row + 32u + 16u * half, lo);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&ready[s]);
            if (++s == S) {
              s = 0;
              ph ^= 1;
            }
            continue;
          }
          const uint32_t a_lo_slot = ptx::smem_u32(lo_base + (cc & 1u) * Cfg::A_BYTES) + row_off;
          // all loads before the first dependent instruction (the inline-asm shared-memory accesses keep program order)
          uint4 raw[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            raw[j] = make_uint4(0u, 0u, 0u, 0u);
            if (j < nch && !(p.exp_flags & 4)) raw[j] = ptx::lds128(a_hi + (((uint32_t)j ^ xr) << 4));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 lo = make_uint4(0u, 0u, 0u, 0u);   // zero-filled K tail: the lo copy must be zero too
            if (j < nch && !(p.exp_flags & 4)) {
              lo.x = tf32_lo_bits(raw[j].x);
              lo.y = tf32_lo_bits(raw[j].y);
              lo.z = tf32_lo_bits(raw[j].z);
              lo.w = tf32_lo_bits(raw[j].w);
              if (p.exp_flags & 8)   // experiment: write the truncated hi operand explicitly instead of leaving the raw fp32 in place
                ptx::sts128(a_hi + (((uint32_t)j ^ xr) << 4),
                            make_uint4(raw[j].x & 0xFFFFE000u, raw[j].y & 0xFFFFE000u, raw[j].z & 0xFFFFE000u, raw[j].w & 0xFFFFE000u));
            }
            ptx::sts128(a_lo_slot + (((uint32_t)j ^ xr) << 4), lo);
          }
          ptx::fence_proxy_async();
          ptx::mbar_arrive(&ready[s]);
          if (++s == S) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ================================== epilogue ====================================
    const int eg = (warp - ntw) >> 2;   // epilogue group: handles work items li with li % n_groups == eg
    const int quarter = warp & 3;     // TMEM lanes 32*quarter .. +31 are visible to this warp
    int li = 0;
    long long w_tfull = 0, t_work = 0;
    [[maybe_unused]] long long t_ld = 0;
    const long long t_begin = ptx::tc_clock();
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++li) {
      if (li % n_groups != eg) continue;           // not this group's item
      const int gj = li / n_groups;
      const int as = eg + n_groups * (gj % ACC_DEPTH);
      const uint32_t use = (uint32_t)(gj / ACC_DEPTH);
      const uint32_t mt = (uint32_t)it / nblk;
      const int n0 = (int)((uint32_t)it - mt * nblk) * p.BN;
      w_tfull += ptx::mbar_wait_timed(&tfull[as], use & 1);
      const long long t_item = ptx::tc_clock();
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * acc_cols);
      const int ncols = min(p.BN, p.N - n0);
      const int64_t m_warp = (int64_t)mt * TC_BM + quarter * 32;   // first row of this warp
      // Accumulator -> registers in the mma fragment layout (tcgen05.ld 16x256b), one exchange with the
      // neighbouring lane so every thread owns FOUR consecutive columns of a row, then BN-fold, swish,
      // residual and a 16-byte (fp32) / 8-byte (bf16) store straight from registers: a quad writes 64
      // (32) contiguous bytes of a row, whole 32-byte sectors, and nothing is staged through shared
      // memory (the smem pipe is the scarce resource of this kernel: TMA fills, UMMA operand reads and
      // the transform warps already share its 128 B/clk).
      const int lr = lane >> 2, q = lane & 3;
      const bool odd = (q & 1) != 0;
      const int rows_valid = (p.M - m_warp) < 32 ? (int)(p.M - m_warp) : 32;   // <= 0 when the warp is past M
      // 64-bit tile bases once, 32-bit offsets inside the tile
      // restrict: the residual (block input) never aliases the output, so its loads may be hoisted above earlier stores
      T* __restrict__ out_t = (T*)p.out + m_warp * (int64_t)p.N + n0;
      const T* __restrict__ res_t = p.res != nullptr ? (const T*)p.res + m_warp * (int64_t)p.N + n0 : nullptr;
      if constexpr (POOL) {
        // rows of this warp: patch 2 mt + quarter / 2, rows (quarter & 1) * 32 .. + 31 of its 49
        const int patch = (int)mt * 2 + (quarter >> 1);
        const int local0 = (quarter & 1) * 32;
        const bool patch_ok = patch < p.pool_nb;
        float* my_pool = pool_s + (eg * 4 + quarter) * 128;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[2][16];
          ptx::tmem_ld16x256b_x4(taddr + (uint32_t)c0, v[0]);
          ptx::tmem_ld16x256b_x4(taddr + (16u << 16) + (uint32_t)c0, v[1]);
          ptx::tmem_ld_wait();
          // thread (lr, q): columns 8 i + 2 q + {0, 1} of the group for i = 0..3, rows lr, lr + 8, lr + 16, lr + 24
          float cs[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 s2 = *reinterpret_cast<const float2*>(sc_s + n0 + c0 + 8 * i + 2 * q);
            const float2 b2 = *reinterpret_cast<const float2*>(bi_s + n0 + c0 + 8 * i + 2 * q);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
              for (int rh = 0; rh < 2; ++rh) {
                const int row = 16 * h2 + 8 * rh + lr;
                float y0 = fmaf(__uint_as_float(v[h2][4 * i + 2 * rh]), s2.x, b2.x);
                float y1 = fmaf(__uint_as_float(v[h2][4 * i + 2 * rh + 1]), s2.y, b2.y);
                if (Cfg::TF32 || !MC_BF16_TANH) {
                  y0 = __fdividef(y0, 1.f + __expf(-y0));
                  y1 = __fdividef(y1, 1.f + __expf(-y1));
                } else {
                  y0 = fmaf(y0, ptx::tanh_approx(y0), y0);
                  y1 = fmaf(y1, ptx::tanh_approx(y1), y1);
                }
                if (local0 + row < TC_POOL_HW) {   // rows past the patch's 49 hold stale shared memory: never summed
                  a0 += y0;
                  a1 += y1;
                }
              }
            cs[2 * i] = a0;
            cs[2 * i + 1] = a1;
          }
          // fixed-order tree over the eight row groups lr (lane bits 2..4)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 4);
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
            cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
          }
          if (lr == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<float2*>(my_pool + c0 + 8 * i + 2 * q) = make_float2(cs[2 * i], cs[2 * i + 1]);
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[as]);
        // the two warps of a patch combine; thread t of the group owns column t of both patches
        asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");
        {
          const int t = quarter * 32 + lane;
          if (t < ncols) {
            const float* gp = pool_s + eg * 4 * 128 + t;
            const float inv = 1.f / (float)TC_POOL_HW;
            float* feats = (float*)p.out;
            const int pa = (int)mt * 2;
            if (pa < p.pool_nb) feats[(int64_t)pa * p.N + n0 + t] = (gp[0] + gp[128]) * inv;
            if (pa + 1 < p.pool_nb) feats[(int64_t)(pa + 1) * p.N + n0 + t] = (gp[256] + gp[384]) * inv;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(2 + eg) : "memory");
        (void)patch_ok;
        t_work += ptx::tc_clock() - t_item;
        continue;
      }
      if constexpr (Cfg::TF32) {
        // fp32: stores straight from the fragment layout.  A thread holds column PAIRS (8 bytes), a quad writes the 32 contiguous
        // bytes of one sector of a row, eight rows per instruction -- the same sectors per instruction as the 16-byte form below,
        // without its lane exchange (16 SHFL + 32 SEL per 32-column group).  N is a multiple of 4 (rows stay 8-byte aligned).  The body is branch-free apart from warp-uniform
        // tests: with a (divergent-looking) bounds branch around each 4-element block ptxas fenced every block with BSSY / BSYNC
        // and the EX2 -> ADD -> RCP -> MUL chains of the eight blocks ran one after the other (23 instructions per element,
        // ncu: issue 63 %, the epilogue warps busy 80 % of the time on the front expand layers).
        const bool do_act = p.act == 1 && !(p.exp_flags & 1);
        const uint32_t sc_u32 = ptx::smem_u32(sc_s + n0 + 2 * q), bi_u32 = ptx::smem_u32(bi_s + n0 + 2 * q);
        const int rowN = lr * p.N + 2 * q, rs8 = 8 * p.N;
        float* __restrict__ orow = (float*)out_t + rowN;                       // + (16 h2 + 8 rh) * N + c0 + 8 i
        const float* __restrict__ rrow = res_t != nullptr ? (const float*)res_t + rowN : nullptr;
        bool rok[2][2];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) rok[h2][rh] = 16 * h2 + 8 * rh + lr < rows_valid;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          // valid columns of this group: the conv widths are multiples of 8, the MLP head's of 4 -- a thread's column pair
          // never straddles the end
          bool cok[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) cok[i] = c0 + 8 * i + 2 * q < ncols;
          [[maybe_unused]] float2 rpre[2][2][4];
          if constexpr (GATED) {
            if (rrow != nullptr) {
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    rpre[h2][rh][i] = make_float2(0.f, 0.f);
                    if (rok[h2][rh] && cok[i]) rpre[h2][rh][i] = *reinterpret_cast<const float2*>(rrow + (2 * h2 + rh) * rs8 + c0 + 8 * i);
                  }
            }
          }
          uint32_t v[2][16];
          if (!(p.exp_flags & 32)) {
            ptx::tmem_ld16x256b_x4(taddr + (uint32_t)c0, v[0]);
            ptx::tmem_ld16x256b_x4(taddr + (16u << 16) + (uint32_t)c0, v[1]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[0][e] = v[1][e] = 0u;
          }
          uint2 sc2[4], bi2[4];   // columns c0 + 8 i + 2 q + {0, 1} (entries past N are never stored)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            sc2[i] = ptx::lds64(sc_u32 + (uint32_t)(c0 + 8 * i) * 4u);
            bi2[i] = ptx::lds64(bi_u32 + (uint32_t)(c0 + 8 * i) * 4u);
          }
          if (!(p.exp_flags & 32)) ptx::tmem_ld_wait();
          float y[2][2][4][2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                y[h2][rh][i][0] = fmaf(__uint_as_float(v[h2][4 * i + 2 * rh]), __uint_as_float(sc2[i].x), __uint_as_float(bi2[i].x));
                y[h2][rh][i][1] = fmaf(__uint_as_float(v[h2][4 * i + 2 * rh + 1]), __uint_as_float(sc2[i].y), __uint_as_float(bi2[i].y));
              }
          if (do_act) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
              for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#if MC_SWISH_NR
                  const float2 sw = silu2_nr(make_float2(y[h2][rh][i][0], y[h2][rh][i][1]));
                  y[h2][rh][i][0] = sw.x;
                  y[h2][rh][i][1] = sw.y;
#else
#pragma unroll
                  for (int e = 0; e < 2; ++e) y[h2][rh][i][e] = __fdividef(y[h2][rh][i][e], 1.f + __expf(-y[h2][rh][i][e]));
#endif
                }
          }
          if (RELU) {
            if (p.act == 2) {
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                  for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < 2; ++e) y[h2][rh][i][e] = fmaxf(y[h2][rh][i][e], 0.f);
            }
          }
          if constexpr (GATED) {
            if (rrow != nullptr) {
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    y[h2][rh][i][0] += rpre[h2][rh][i].x;
                    y[h2][rh][i][1] += rpre[h2][rh][i].y;
                  }
            }
          } else if (rrow != nullptr) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
              for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (rok[h2][rh] && cok[i]) {
                    const float2 r2 = *reinterpret_cast<const float2*>(rrow + (2 * h2 + rh) * rs8 + c0 + 8 * i);
                    y[h2][rh][i][0] += r2.x;
                    y[h2][rh][i][1] += r2.y;
                  }
          }
          if (!(p.exp_flags & 2)) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
              for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (rok[h2][rh] && cok[i])
                    *reinterpret_cast<float2*>(orow + (2 * h2 + rh) * rs8 + c0 + 8 * i) = make_float2(y[h2][rh][i][0], y[h2][rh][i][1]);
          }
        }
      } else
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        // residual (skip connection) values of this column group: all eight loads in flight BEFORE the accumulator is
        // read.  Issued where they are consumed they sit between the lane exchanges and serialise -- eight exposed L2 / HBM
        // round trips per group (ncu on b2.project: 41 % of the epilogue's time).
        [[maybe_unused]] typename ResVec<T>::type rpre[2][2][2];
        if constexpr (GATED) {
          if (res_t != nullptr) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
              for (int pr = 0; pr < 2; ++pr)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                  const int cb = odd ? 16 * pr + 8 + 2 * (q - 1) : 16 * pr + 2 * q;
                  const int row = 16 * h2 + 8 * rh + lr;
                  rpre[h2][pr][rh] = ResVec<T>::zero();
                  if (row < rows_valid && c0 + cb < ncols) rpre[h2][pr][rh] = ResVec<T>::load(res_t + row * p.N + c0 + cb);
                }
          }
        }
        // both 16-lane halves in flight before the single wait: the tcgen05.ld round trip is the longest
        // latency of the epilogue (serialising the halves cost 40 % on the expand layers)
        uint32_t v[2][16];
        if (!(p.exp_flags & 32)) {
          ptx::tmem_ld16x256b_x4(taddr + (uint32_t)c0, v[0]);                      // lanes  0..15 of this warp's quarter
          ptx::tmem_ld16x256b_x4(taddr + (16u << 16) + (uint32_t)c0, v[1]);        // lanes 16..31
          ptx::tmem_ld_wait();
          t_ld += ptx::tc_clock() - t_item;
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[0][e] = v[1][e] = 0u;
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {        // column-block pair (8-column blocks 2pr, 2pr+1)
            // this thread ends up with columns cb .. cb+3 of the 32-column group
            const int cb = odd ? 16 * pr + 8 + 2 * (q - 1) : 16 * pr + 2 * q;
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {       // rows lr and lr + 8 of the 16-lane half
              const uint32_t a0 = v[h2][8 * pr + 2 * rh], a1 = v[h2][8 * pr + 2 * rh + 1];          // block 2pr
              const uint32_t b0 = v[h2][8 * pr + 4 + 2 * rh], b1 = v[h2][8 * pr + 4 + 2 * rh + 1];  // block 2pr+1
              const uint32_t s0 = odd ? a0 : b0, s1 = odd ? a1 : b1;
              const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
              float y[4];
              y[0] = __uint_as_float(odd ? r0 : a0);
              y[1] = __uint_as_float(odd ? r1 : a1);
              y[2] = __uint_as_float(odd ? b0 : r0);
              y[3] = __uint_as_float(odd ? b1 : r1);
              const int row = 16 * h2 + 8 * rh + lr;      // row inside this warp's 32
              if (row < rows_valid && c0 + cb < ncols) {
                const float4 s4 = *reinterpret_cast<const float4*>(sc_s + n0 + c0 + cb);
                const float4 b4 = *reinterpret_cast<const float4*>(bi_s + n0 + c0 + cb);
                y[0] = fmaf(y[0], s4.x, b4.x);
                y[1] = fmaf(y[1], s4.y, b4.y);
                y[2] = fmaf(y[2], s4.z, b4.z);
                y[3] = fmaf(y[3], s4.w, b4.w);
                if (p.act == 1 && !(p.exp_flags & 1)) {
                  if (Cfg::TF32 || !MC_BF16_TANH) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) y[e] = __fdividef(y[e], 1.f + __expf(-y[e]));
                  } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) y[e] = fmaf(y[e], ptx::tanh_approx(y[e]), y[e]);
                  }
                }
                if (RELU) {   // MLP head Linear layers (hidden layers; the last layer runs with act 0 but shares the instantiation)
                  if (p.act == 2) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) y[e] = fmaxf(y[e], 0.f);
                  }
                }
                const int off = row * p.N + c0 + cb;
                if (res_t != nullptr) {
                  float r[4];
                  if constexpr (GATED) ResVec<T>::unpack(rpre[h2][pr][rh], r);
                  else load4<T>(res_t + off, r);
#pragma unroll
                  for (int e = 0; e < 4; ++e) y[e] += r[e];
                }
                if (!(p.exp_flags & 2)) store4<T>(out_t + off, y);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[as]);
      t_work += ptx::tc_clock() - t_item;
    }
    if (MC_TC_TIMING && p.dbg && blockIdx.x == 0 && lane == 0 && (warp & 3) == 2) {
      p.dbg[8 + 3 * eg] = w_tfull;
      p.dbg[9 + 3 * eg] = t_work;
      p.dbg[10 + 3 * eg] = ptx::tc_clock() - t_begin;
      p.dbg[26 + eg] = t_ld;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
// host side: per-layer plans (tensor maps + arguments)
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2-D K-major map: dim0 = K (contiguous), dim1 = rows; box = (128 bytes of K) x box_rows; 128B swizzle; OOB -> 0.
inline int make_map(CUtensorMap* map, bool f32, const void* base, int64_t rows, int K, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = f32 ? 4 : 2;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return MC_OK;
}

// SE gate map: dim0 = K (contiguous), dim1 = patches; box = 128 bytes of K x TC_GATE_ROWS patches; plain layout; OOB -> 0.
inline int make_gate_map(CUtensorMap* map, bool f32, const void* base, int64_t rows, int K) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = f32 ? 4 : 2;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)TC_GATE_ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled (gate) failed with CUresult " + std::to_string((int)r));
  return MC_OK;
}

struct PwTcLayer {
  bool present = false;
  int N = 0, K = 0, BN = 0, n_blocks = 0, k_chunks = 0, act = 0;
  int w_mode = 0;         // PwTcArgs::w_res: 0 streamed, 1 whole weight resident, 2 one n-block resident per CTA
  bool ts = false;        // gated fp32: A operands through tensor memory (pw_tc_kernel TS instantiation)
  bool gated = false;
  void* d_w = nullptr;    // bf16 weights, or fp32 hi part
  void* d_wlo = nullptr;  // fp32 lo part
  const float *scale = nullptr, *bias = nullptr;  // into the extractor's parameter blob
  CUtensorMap tmW, tmWlo;
  // activation maps are cached per source pointer (the ping-pong buffers alternate)
  const void* a_ptr[2] = {nullptr, nullptr};
  int64_t a_rows[2] = {0, 0};
  CUtensorMap tmA[2];
  // gated launches: map over the SE gate [patches][K] with a {128 bytes of K, 4 patches} box, no swizzle
  const void* g_ptr = nullptr;
  CUtensorMap tmG;
  // POOL launch (head conv): map with a 49-row box over the same buffer
  const void* apool_ptr = nullptr;
  CUtensorMap tmApool;
};

struct PwTcPlan {
  int mode = 0, device = 0, num_sms = 148, max_batch = 0;
  bool relu_variant = false;   // launch the RELU instantiation (MLP head)
  long long* dbg_buf = nullptr;   // MC_TC_DBG role counters (device memory of this plan's device)
  std::vector<PwTcLayer> layers;  // 2*b = expand of block b, 2*b+1 = project, 32 = head conv
};

inline bool pw_tc_has(const PwTcPlan* p, int id) { return p && id < (int)p->layers.size() && p->layers[id].present; }

inline void pw_tc_free(PwTcPlan* p) {
  if (!p) return;
  for (auto& l : p->layers) {
    if (l.d_w) cudaFree(l.d_w);
    if (l.d_wlo) cudaFree(l.d_wlo);
  }
  if (p->dbg_buf) cudaFree(p->dbg_buf);
  delete p;
}

inline int pick_bn(int N, int bn_max) {
  // smallest number of equal blocks with BN a multiple of 16 and <= bn_max
  const int n16 = (N + 15) / 16 * 16;
  int blocks = (n16 + bn_max - 1) / bn_max;
  int bn = ((n16 / 16 + blocks - 1) / blocks) * 16;
  return bn;
}

inline int pw_tc_add(PwTcPlan* plan, int id, const float* w_host, const float* d_scale, const float* d_bias, int N, int K,
                     int act, bool gated) {
  PwTcLayer& l = plan->layers[id];
  const bool f32 = plan->mode == MC_MODE_FP32;
  l.N = N;
  l.K = K;
  l.act = act;
  l.gated = gated;
  l.scale = d_scale;
  l.bias = d_bias;
  {
    // MC_TC_TS_MASK=<hex>: fp32 layers (bit = plan layer id) that take A through tensor memory; default all of them
    static const unsigned long long ts_mask = getenv("MC_TC_TS_MASK") ? strtoull(getenv("MC_TC_TS_MASK"), nullptr, 16) : ~0ull;
    l.ts = (f32 || gated) && id < 64 && ((ts_mask >> id) & 1ull) != 0;
    // measured (B200, per-layer CUDA events, same run): every gated layer gains (-6 % at K = 32 ... -26 % at K = 1152); ungated
    // layers gain from K = 80 on (b6-8.expand -10 %, b9-11 -17 %, b12-15 -42 %) and lose 3 % below (one k-chunk or two: the
    // tcgen05.st round trip is not amortised); the head conv (N = 1280) loses 7 % to the narrower blocks (14 x 96 for 10 x 128)
    if (!gated && (K <= 64 || (N == 1280 && !plan->relu_variant))) l.ts = false;
    // bf16 gated layers: -20..25 % on b11-b15.project (K >= 672, two n-blocks), +5..10 % on the single-block K = 480 / 672 ones
    // (one accumulator per epilogue group instead of two), a wash on the front layers
    if (!f32 && !(K >= 672 && N >= 192)) l.ts = false;
  }
  // ungated TS layers: four 96-column accumulators + the A ring share the 512 TMEM columns
  const int bn_max = !f32 ? TcCfg<__nv_bfloat16>::BN_MAX : ((l.ts && !gated) ? 96 : TcCfg<float>::BN_MAX);
  l.BN = pick_bn(N, bn_max);
  l.n_blocks = (N + l.BN - 1) / l.BN;
  const int kc = f32 ? TcCfg<float>::KC : TcCfg<__nv_bfloat16>::KC;
  l.k_chunks = (K + kc - 1) / kc;
  // fp32 layers whose weights do not stay resident stream W hi + lo tiles through the TMA ring: cap the block width so that FOUR
  // stages fit (ungated: 96 columns -> 40 KB stages; gated: 112 -> 44 KB).  Their k-chunk period is (TMA latency + transform +
  // MMA) / stages, so the fourth stage is worth more than the wider block.
  // Measured: project layers -5..7 %, the 14x14 expand layers -9 %; the head conv (1280 = 13.3 x 96: padded blocks) +6 %, so
  // the cap only applies where 96 divides N.
  if (f32 && tc_w_res_bytes<float>(l.BN, l.n_blocks, l.k_chunks) > TC_W_RES_MAX && (gated || N % 96 == 0)) {
    l.BN = pick_bn(N, gated ? 112 : 96);
    l.n_blocks = (N + l.BN - 1) / l.BN;
  }
  // Weight residency.  Whole layer when it fits; else (expand-shaped layers: K small, N large) one n-block per CTA, with the
  // widest block that still leaves FOUR activation stages (MC_TC_BN="<id>:<bn>,..." overrides the width, MC_TC_NO_WRES_NB
  // disables the mode).
  {
    auto wbytes = [&](int bn, int nblocks) { return f32 ? tc_w_res_bytes<float>(bn, nblocks, l.k_chunks) : tc_w_res_bytes<__nv_bfloat16>(bn, nblocks, l.k_chunks); };
    auto stages_res = [&](int wb) { return f32 ? tc_num_stages_res<float>(wb, l.ts) : tc_num_stages_res<__nv_bfloat16>(wb); };
    static const bool no_w_res = getenv("MC_TC_NO_WRES") != nullptr, no_nb = getenv("MC_TC_NO_WRES_NB") != nullptr;
    l.w_mode = 0;
    if (!no_w_res && wbytes(l.BN, l.n_blocks) <= TC_W_RES_MAX) {
      l.w_mode = 1;
    } else if (!no_w_res && !no_nb && !gated) {
      int bn_forced = 0;
      if (const char* e = getenv("MC_TC_BN")) {
        for (const char* q = e; q && *q;) {
          int lid = -1, bn = 0;
          if (sscanf(q, "%d:%d", &lid, &bn) == 2 && lid == id) bn_forced = bn;
          q = strchr(q, ',');
          if (q) ++q;
        }
      }
      for (int cap = bn_forced ? bn_forced : bn_max; cap >= 64; cap -= 16) {
        const int bn = pick_bn(N, cap), nblocks = (N + bn - 1) / bn;
        if (nblocks > plan->num_sms) break;
        if (stages_res(wbytes(bn, 1)) >= (bn_forced ? 2 : 4)) {
          l.BN = bn;
          l.n_blocks = nblocks;
          l.w_mode = 2;
          break;
        }
        if (bn_forced) break;
      }
    }
  }
  const size_t n = (size_t)N * K;
  int rc;
  if (f32) {
    std::vector<float> hi(n), lo(n);
    for (size_t i = 0; i < n; ++i) {
      uint32_t bits;
      memcpy(&bits, &w_host[i], 4);
      const uint32_t lb = tf32_lo_bits(bits);
      bits &= 0xFFFFE000u;
      memcpy(&hi[i], &bits, 4);
      memcpy(&lo[i], &lb, 4);
    }
    MC_CUDA(cudaMalloc(&l.d_w, n * 4));
    MC_CUDA(cudaMalloc(&l.d_wlo, n * 4));
    MC_CUDA(cudaMemcpy(l.d_w, hi.data(), n * 4, cudaMemcpyHostToDevice));
    MC_CUDA(cudaMemcpy(l.d_wlo, lo.data(), n * 4, cudaMemcpyHostToDevice));
    if ((rc = make_map(&l.tmW, true, l.d_w, N, K, l.BN)) || (rc = make_map(&l.tmWlo, true, l.d_wlo, N, K, l.BN))) return rc;
  } else {
    std::vector<__nv_bfloat16> wb(n);
    for (size_t i = 0; i < n; ++i) wb[i] = __float2bfloat16_rn(w_host[i]);
    MC_CUDA(cudaMalloc(&l.d_w, n * 2));
    MC_CUDA(cudaMemcpy(l.d_w, wb.data(), n * 2, cudaMemcpyHostToDevice));
    if ((rc = make_map(&l.tmW, false, l.d_w, N, K, l.BN))) return rc;
    l.tmWlo = l.tmW;
  }
  l.present = true;
  return MC_OK;
}

// `params_host` is the packed blob (csrc/layers.h); scale/bias pointers index the device copy.
inline int pw_tc_build(PwTcPlan** out, const NetCfg& net, const float* params_host, const float* d_params, int mode,
                       int max_batch, int device, unsigned layer_mask_lo, unsigned layer_mask_hi) {
  *out = nullptr;
  PwTcPlan* plan = new PwTcPlan();
  plan->mode = mode;
  plan->device = device;
  cudaDeviceProp prop;
  MC_CUDA(cudaGetDeviceProperties(&prop, device));
  plan->num_sms = prop.multiProcessorCount;
  plan->layers.resize(33);
  const bool f32 = mode == MC_MODE_FP32;
  if (f32) {
    MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
  } else {
    MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
  }
  if (f32) MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
  else MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<__nv_bfloat16, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
  auto enabled = [&](int id) { return id < 32 ? ((layer_mask_lo >> id) & 1u) != 0 : ((layer_mask_hi >> (id - 32)) & 1u) != 0; };
  int rc = MC_OK;
  for (size_t bi = 0; bi < net.blocks.size() && rc == MC_OK; ++bi) {
    const BlockCfg& b = net.blocks[bi];
    if (b.expand != 1 && enabled((int)bi * 2))
      rc = pw_tc_add(plan, (int)bi * 2, params_host + b.w_exp, d_params + b.s_exp, d_params + b.b_exp, b.c_mid, b.c_in, 1, false);
    if (rc == MC_OK && enabled((int)bi * 2 + 1))
      rc = pw_tc_add(plan, (int)bi * 2 + 1, params_host + b.w_proj, d_params + b.s_proj, d_params + b.b_proj, b.c_out, b.c_mid,
                     0, true);
  }
  if (rc == MC_OK && enabled(32))
    rc = pw_tc_add(plan, 32, params_host + net.w_head, d_params + net.s_head, d_params + net.b_head, 1280, 320, 1, false);
  if (rc != MC_OK) {
    pw_tc_free(plan);
    return rc;
  }
  plan->max_batch = max_batch;
  *out = plan;
  return MC_OK;
}

// `A` is the BASE of the activation buffer (its tensor map is cached per layer); the launch covers rows
// [a_row_off, a_row_off + M) of it.  gate / res / outp are already offset to the launch's first row.
// map_rows >= 0: the activation buffer is caller-owned and holds exactly that many rows (the map must not reach past it).
inline int pw_tc_run(PwTcPlan* plan, int id, const void* A, int64_t a_row_off, const void* gate, const void* res, void* outp,
                     int64_t M, int HW, cudaStream_t st, int64_t map_rows = -1) {
  PwTcLayer& l = plan->layers[id];
  const bool f32 = plan->mode == MC_MODE_FP32;
  // tensor map of the activation source (cached: each layer only ever sees the ping-pong buffers)
  const int64_t rows = map_rows >= 0 ? map_rows : (int64_t)plan->max_batch * HW;
  int slot = -1;
  for (int i = 0; i < 2; ++i)
    if (l.a_ptr[i] == A && l.a_rows[i] == rows) slot = i;
  if (slot < 0) {
    slot = l.a_ptr[0] == nullptr ? 0 : 1;
    // rows = the buffer's capacity for this layer: rows past it are zero-filled by TMA, never read
    int rc = make_map(&l.tmA[slot], f32, A, rows, l.K, TC_BM);
    if (rc) return rc;
    l.a_ptr[slot] = A;
    l.a_rows[slot] = rows;
  }
  PwTcArgs a;
  a.scale = l.scale;
  a.bias = l.bias;
  a.gate = l.gated ? gate : nullptr;
  a.res = res;
  a.out = outp;
  a.M = M;
  a.a_row_off = (int)a_row_off;
  a.pool_nb = 0;
  const char* exp_env = getenv("MC_TC_EXP");   // read per launch: experiments switch it inside one process
  const int exp_flags = exp_env ? atoi(exp_env) : 0;
  a.exp_flags = exp_flags;
  // MC_TC_DBG=<layer id>: after that layer's launch, print CTA 0's per-role wait/total cycles (synchronises)
  static const int dbg_layer = getenv("MC_TC_DBG") ? atoi(getenv("MC_TC_DBG")) : -1;
  a.dbg = nullptr;
  if (dbg_layer == id) {
    if (!plan->dbg_buf) cudaMalloc((void**)&plan->dbg_buf, 32 * sizeof(long long));
    cudaMemsetAsync(plan->dbg_buf, 0, 32 * sizeof(long long), st);
    a.dbg = plan->dbg_buf;
  }
  a.N = l.N;
  a.K = l.K;
  a.HW = HW;
  a.BN = l.BN;
  a.n_blocks = l.n_blocks;
  a.k_chunks = l.k_chunks;
  a.act = l.act;
  a.m_tiles = (M + TC_BM - 1) / TC_BM;
  const int64_t items = a.m_tiles * a.n_blocks;
  // mode 2 needs a grid that is a multiple of n_blocks (CTA c <-> n-block c % n_blocks); tiny launches stream instead
  const int grid_nb = plan->num_sms / l.n_blocks * l.n_blocks;
  a.w_res = l.w_mode;
  if (a.w_res == 2 && items < grid_nb) a.w_res = 0;
  const int w_bytes = f32 ? tc_w_res_bytes<float>(l.BN, a.w_res == 2 ? 1 : l.n_blocks, l.k_chunks)
                          : tc_w_res_bytes<__nv_bfloat16>(l.BN, a.w_res == 2 ? 1 : l.n_blocks, l.k_chunks);
  size_t smem;
  const bool ts = l.ts && (a.gate != nullptr) == l.gated;
  const size_t lo_ring = (f32 && !ts) ? tc_lo_ring_bytes<float>() : 0;
  if (a.w_res) {
    a.stages = f32 ? tc_num_stages_res<float>(w_bytes, ts) : tc_num_stages_res<__nv_bfloat16>(w_bytes);
    smem = TC_FIXED_BYTES + (size_t)w_bytes + lo_ring + (size_t)a.stages * TC_BM * 128;
  } else {
    a.stages = f32 ? tc_num_stages<float>(l.BN, 0, ts) : tc_num_stages<__nv_bfloat16>(l.BN);
    smem = TC_FIXED_BYTES + lo_ring + (size_t)a.stages * (f32 ? tc_stage_bytes<float>(l.BN) : tc_stage_bytes<__nv_bfloat16>(l.BN));
  }
  const int grid = a.w_res == 2 ? grid_nb : (int)std::min<int64_t>(items, plan->num_sms);
  const bool gated = a.gate != nullptr;
  if (gated && l.g_ptr != a.gate) {
    int rc = make_gate_map(&l.tmG, f32, a.gate, plan->max_batch, l.K);
    if (rc) return rc;
    l.g_ptr = a.gate;
  }
  const CUtensorMap& tmG = gated ? l.tmG : l.tmW;   // ungated kernels never touch it
  if (plan->relu_variant && ts) {   // MLP head (fp32, ungated)
    static std::atomic<unsigned long long> attr_mask{0};
    if (first_use_on_device(attr_mask))
      MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, false, true, false, true>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  } else if (plan->relu_variant) {
    static std::atomic<unsigned long long> attr_mask{0};
    if (first_use_on_device(attr_mask))
      MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, false, true>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  } else if (f32 && !gated && ts) {
    static std::atomic<unsigned long long> attr_mask{0};
    if (first_use_on_device(attr_mask))
      MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, false, false, false, true>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  } else if (f32 && gated && ts) {
    static std::atomic<unsigned long long> attr_mask{0};
    if (first_use_on_device(attr_mask))
      MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, true, false, false, true>, dim3(grid), dim3(tc_threads<true>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  } else if (!f32 && gated && ts) {
    static std::atomic<unsigned long long> attr_mask{0};
    if (first_use_on_device(attr_mask))
      MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<__nv_bfloat16, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<__nv_bfloat16, true, false, false, true>, dim3(grid), dim3(tc_threads<true>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  } else if (f32 && gated)
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, true>, dim3(grid), dim3(tc_threads<true>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  else if (f32)
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, false>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  else if (gated)
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<__nv_bfloat16, true>, dim3(grid), dim3(tc_threads<true>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  else
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<__nv_bfloat16, false>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmA[slot], l.tmW, l.tmWlo, tmG, a));
  MC_CHECK_LAUNCH();
  if (a.dbg) {
    long long d[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(d, plan->dbg_buf, sizeof(d), cudaMemcpyDeviceToHost);
    static int printed = 0;
    if (printed++ < 3) {
      const double it = (double)std::max<long long>(d[5], 1);
      fprintf(stderr, "[MC_TC_DBG layer %d] M=%lld N=%d K=%d BN=%d stages=%d items/CTA=%lld | per item (cycles): producer wait_empty %.0f total %.0f | "
                      "mma wait_tempty %.0f wait_full %.0f commits %.0f issue %.0f epi0_ld %.0f total %.0f | transform wait_full %.0f total %.0f |",
              id, (long long)M, l.N, l.K, l.BN, a.stages, d[5], d[0] / it, d[1] / it, d[2] / it, d[3] / it, d[24] / it, d[25] / it, d[26] / it, d[4] / it, d[6] / it, d[7] / it);
      for (int g = 0; g < (gated ? 2 : 4); ++g)
        fprintf(stderr, " epi%d wait_tfull %.0f work %.0f total %.0f |", g, d[8 + 3 * g] / it, d[9 + 3 * g] / it, d[10 + 3 * g] / it);
      fprintf(stderr, "\n");
    }
  }
  return MC_OK;
}

// Head conv + global average pool in one launch (the POOL instantiation): A = block output [nb * 49][K], feats = [nb][N] fp32.
inline int pw_tc_run_pool(PwTcPlan* plan, int id, const void* A, int nb, float* feats, cudaStream_t st) {
  PwTcLayer& l = plan->layers[id];
  const bool f32 = plan->mode == MC_MODE_FP32;
  if (l.apool_ptr != A) {
    int rc = make_map(&l.tmApool, f32, A, (int64_t)plan->max_batch * TC_POOL_HW, l.K, TC_POOL_HW);
    if (rc) return rc;
    l.apool_ptr = A;
  }
  PwTcArgs a;
  a.scale = l.scale;
  a.bias = l.bias;
  a.gate = nullptr;
  a.res = nullptr;
  a.out = feats;
  a.M = (int64_t)nb * TC_POOL_HW;
  a.a_row_off = 0;
  { const char* e_ = getenv("MC_TC_EXP"); a.exp_flags = e_ ? atoi(e_) : 0; }
  a.dbg = nullptr;
  a.N = l.N;
  a.K = l.K;
  a.HW = TC_POOL_HW;
  a.BN = l.BN;
  a.n_blocks = l.n_blocks;
  a.k_chunks = l.k_chunks;
  a.act = l.act;
  a.w_res = 0;
  a.pool_nb = nb;
  const int stage_bytes = f32 ? tc_stage_bytes<float>(l.BN) : tc_stage_bytes<__nv_bfloat16>(l.BN);
  const bool ts = f32 && l.ts;
  a.stages = f32 ? tc_num_stages<float>(l.BN, TC_POOL_BYTES, ts) : tc_num_stages<__nv_bfloat16>(l.BN, TC_POOL_BYTES);
  const size_t smem = TC_FIXED_BYTES + TC_POOL_BYTES + ((f32 && !ts) ? tc_lo_ring_bytes<float>() : 0) + (size_t)a.stages * stage_bytes;
  a.m_tiles = (nb + 1) / 2;
  const int64_t items = a.m_tiles * a.n_blocks;
  const int grid = (int)std::min<int64_t>(items, plan->num_sms);
  if (ts) {
    static std::atomic<unsigned long long> attr_mask{0};
    if (first_use_on_device(attr_mask))
      MC_CUDA(cudaFuncSetAttribute(pw_tc_kernel<float, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET));
    MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, false, false, true, true>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmApool, l.tmW, l.tmWlo, l.tmW, a));
  } else if (f32) MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<float, false, false, true>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmApool, l.tmW, l.tmWlo, l.tmW, a));
  else MC_CUDA(launch_pdl(PDL_GEMM, pw_tc_kernel<__nv_bfloat16, false, false, true>, dim3(grid), dim3(tc_threads<false>()), smem, st, l.tmApool, l.tmW, l.tmWlo, l.tmW, a));
  MC_CHECK_LAUNCH();
  return MC_OK;
}

}  // namespace mc
