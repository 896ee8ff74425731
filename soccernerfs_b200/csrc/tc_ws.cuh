// Warp-specialised tcgen05 pipelines for the decoder layers (sm_100a): producer warps / one MMA-issuing warp / epilogue
// warps, coupled through mbarriers only -- no block-wide barrier inside the main loop.
//
// Why (ncu, profiles/ncu_r2_decoder_cfg3_before.txt): the lock-step kernels these replace (tc_pipe_kernel / tc_wgrad_kernel,
// removed; tc_rowtile_kernel in tc_linear.cu is what is left of that design, for layers wider than the pipelines take) ran
// every phase -- global loads, hi/lo split + swizzled stores, MMA issue, TMEM unload -- on the same 8 warps one after the
// other: 12.5 % active warps, 77 M warp instructions for the 192->128 layer (16.7 % of them an emulated cvt.rna.tf32),
// tensor pipe 5-15 % busy.  Here
//   * 8 producer warps do nothing but  LDG.128 -> (hi = x & 0xffffe000, lo = x - hi) -> 2 x STS.128  into an NST-stage
//     ring, with the next step's loads already in flight in registers;
//   * 1 warp issues the tcgen05.mma's of a stage as soon as its "full" mbarrier flips and hands the stage back through
//     tcgen05.commit -> "empty";
//   * 4 epilogue warps (one per TMEM lane quadrant) unload accumulator j & 1 while the MMAs of tile j + 1 run into the
//     other accumulator.
// hi is the TF32 the tensor core reads from an fp32 word anyway (low 13 mantissa bits ignored), so one LOP3 replaces the
// rounding conversion; lo = x - hi is exact and the tensor core's truncation of lo costs <= 2^-21 relative.
#pragma once
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time (no -lcuda)

#include "tc_common.cuh"

namespace kp {

constexpr int kWsEpiWarps = 4, kWsProdWarps = 8;
constexpr int kWsThreads = 32 * (kWsEpiWarps + 1 + kWsProdWarps);  // 416
constexpr int kWsProdThreads = 32 * kWsProdWarps;                  // 256
constexpr int kWsProdTid0 = 32 * (kWsEpiWarps + 1);                // first producer thread
constexpr int kWsMaxStages = 6;

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// try_wait with a suspend-time hint: the waiting warp sleeps in hardware instead of burning issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  const uint32_t b = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(b), "r"(parity), "r"(200000u)
        : "memory");
  }
}
// latency-critical hand-offs of the pipeline (stage full / empty): plain try_wait spin, no suspend hint -- the hinted wait
// wakes up late, and every step crosses three such hand-offs
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t b = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(b), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ float4 hi4(float4 x) {
  return make_float4(__uint_as_float(__float_as_uint(x.x) & 0xffffe000u), __uint_as_float(__float_as_uint(x.y) & 0xffffe000u),
                     __uint_as_float(__float_as_uint(x.z) & 0xffffe000u), __uint_as_float(__float_as_uint(x.w) & 0xffffe000u));
}

// ---- producer side: a [ROWS x COLS] fp32 sub-tile per step, 256 producer threads --------------------------------------
// thread t owns 16-byte chunk (t % CPR) of rows (t / CPR) + u * (256 / CPR), u < U = ROWS * CPR / 256.
template <int ROWS, int COLS>
struct WsRegs {
  static constexpr int CPR = COLS / 4, ROWSTEP = kWsProdThreads / CPR, U = ROWS / ROWSTEP;
  float4 v[U];
};
// FULL: every row and column of the sub-tile exists and rows are 16-byte aligned (no predicates at all)
template <int ROWS, int COLS, bool FULL>
__device__ __forceinline__ void ws_load(WsRegs<ROWS, COLS>& t, const float* __restrict__ src, int64_t ld, int pt, int rows_valid,
                                        int cols_valid) {
  constexpr int CPR = COLS / 4, ROWSTEP = kWsProdThreads / CPR, U = ROWS / ROWSTEP;
  const int r0 = pt / CPR, ch = pt % CPR;
  const float* p = src + (int64_t)r0 * ld + ch * 4;
  if (FULL) {
#pragma unroll
    for (int u = 0; u < U; ++u) t.v[u] = __ldg(reinterpret_cast<const float4*>(p + (int64_t)u * ROWSTEP * ld));
  } else {
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (ch * 4 + 4 <= cols_valid);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      t.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + u * ROWSTEP < rows_valid) {
        const float* q = p + (int64_t)u * ROWSTEP * ld;
        if (vec) {
          t.v[u] = __ldg(reinterpret_cast<const float4*>(q));
        } else {
          if (ch * 4 + 0 < cols_valid) t.v[u].x = __ldg(q + 0);
          if (ch * 4 + 1 < cols_valid) t.v[u].y = __ldg(q + 1);
          if (ch * 4 + 2 < cols_valid) t.v[u].z = __ldg(q + 2);
          if (ch * 4 + 3 < cols_valid) t.v[u].w = __ldg(q + 3);
        }
      }
    }
  }
}
// Offset (in floats) of this thread's first chunk inside a staged [ROWS x COLS] operand: COLS/32 blocks of [ROWS x 32],
// K-major view: 8-row atoms, 16-byte chunk c of row r at c ^ (r & 7);  MN-major view (128B_BASE32B): 4-row atoms,
// 32-byte chunk q of row r at q ^ (r & 3).  ROWSTEP is a multiple of 8, so the swizzle phase is the same for every u.
template <int ROWS, int COLS, int MN_VIEW>
__device__ __forceinline__ int ws_store_offset(int pt) {
  constexpr int CPR = COLS / 4;
  const int r0 = pt / CPR, ch = pt % CPR;
  const int kb = ch >> 3, c = ch & 7;
  const int cs = MN_VIEW ? ((((c >> 1) ^ (r0 & 3)) << 1) | (c & 1)) : (c ^ (r0 & 7));
  return kb * (ROWS * 32) + r0 * 32 + (cs << 2);
}
template <int ROWS, int COLS>
__device__ __forceinline__ void ws_store(const WsRegs<ROWS, COLS>& t, float* __restrict__ s_hi, float* __restrict__ s_lo, int off0) {
  constexpr int CPR = COLS / 4, ROWSTEP = kWsProdThreads / CPR, U = ROWS / ROWSTEP;
  static_assert(ROWSTEP % 8 == 0, "swizzle phase must not depend on u");
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float4 x = t.v[u];
    const float4 hi = hi4(x);
    const float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    *reinterpret_cast<float4*>(s_hi + off0 + u * ROWSTEP * 32) = hi;
    *reinterpret_cast<float4*>(s_lo + off0 + u * ROWSTEP * 32) = lo;
  }
}

// ---- TMA-fed variant --------------------------------------------------------------------------------------------------
// The activation sub-tile is brought in by the TMA engine (cp.async.bulk.tensor.2d, one [ROWS x 32-column] box per 32-col
// block, written straight into the swizzled UMMA layout: SWIZZLE_128B for K-major operands, SWIZZLE_128B_ATOM_32B for the
// MN-major views) and completes on an mbarrier: no registers, no scoreboards, any number of stages in flight, rows /
// columns past the matrix zero-filled by the hardware.  The raw fp32 words ARE the hi operand (kind::tf32 ignores the low
// 13 mantissa bits); the producer warps only derive lo = x - (x & 0xffffe000) from shared memory, linearly (lo has the
// same layout as hi, so the swizzle never appears in SIMT code).
constexpr int kTmaThreads = kWsThreads + 32;     // + the TMA-issuing warp
constexpr int kTmaWarp = kWsEpiWarps + 1;        // warp 5
constexpr int kTmaProdTid0 = 32 * (kWsEpiWarps + 2);

__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c_inner, int c_outer, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// lo = x - trunc_tf32(x) for N4 float4 of a staged operand (256 producer threads, linear, conflict-free)
template <int N4>
__device__ __forceinline__ void ws_derive_lo(const float* __restrict__ s_hi, float* __restrict__ s_lo, int pt) {
  static_assert(N4 % kWsProdThreads == 0, "operand size");
  const float4* h = reinterpret_cast<const float4*>(s_hi);
  float4* l = reinterpret_cast<float4*>(s_lo);
#pragma unroll
  for (int u = 0; u < N4 / kWsProdThreads; ++u) {
    const float4 x = h[pt + u * kWsProdThreads];
    const float4 hi = hi4(x);
    l[pt + u * kWsProdThreads] = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
  }
}

// Host: tensor map of a row-major fp32 matrix [rows x cols] (leading dimension ld floats) read in boxes of
// [box_rows x 32 columns].  Returns false when the matrix cannot be described (alignment) or the driver lacks the entry point.
static inline bool make_tmap_2d(CUtensorMap* out, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                                bool mn_view) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<Fn>(p);
  }();
  if (fn == nullptr || rows < 1 || cols < 1 || box_rows < 1 || box_rows > 256) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 4) & 15) != 0 || rows >= (int64_t(1) << 31)) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            mn_view ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace kp
