// Shared device/host helpers for libkplanes_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/kplanes_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libkplanes_b200 targets sm_100a (B200) only"
#endif

namespace kp {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;  // kernels launched through the C-ABI (kp_launch_count); the autograd
                                            // thread and the main thread both launch

#define KP_CHECK(cond, ...)          \
  do {                               \
    if (!(cond)) {                   \
      kp::set_error(__VA_ARGS__);    \
      return 1;                      \
    }                                \
  } while (0)

#define KP_LAUNCH_CHECK(name)                                                      \
  do {                                                                             \
    kp::g_launches += 1;                                                           \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      kp::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 256-bit read-only load (sm_100: LDG.E.ENL2.256): a whole 8-channel texel in one request instead of two
struct __align__(32) float8 { float4 a, b; };
__device__ __forceinline__ float8 ldg8(const float* p) {
  float8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void stg8(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
               "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}

// Vector reduction (no return value) into global memory: one 16-byte L2 atomic instead of four.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 scale4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 fma4(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2: one issue slot for two IEEE fp32 operations -- the 3-register scalar FFMA issues
// every other cycle per scheduler, so instruction-bound fp32 code doubles its FMA rate; each half rounds exactly like the
// scalar instruction, so results are bit-identical).
__device__ __forceinline__ float2 dup2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 lo2(float4 a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ float2 hi2(float4 a) { return make_float2(a.z, a.w); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive warp scan (double): used to reproduce torch's CPU cumsum, which accumulates fp32 in double
__device__ __forceinline__ double warp_incl_scan_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive warp scan (double).  NOT "inclusive - own value": with an infinite element that would be inf - inf = NaN
// for the lane holding it, while torch's cumsum gives the finite prefix (e.g. an infinite density in get_weights).
__device__ __forceinline__ double warp_excl_scan_d(double v, int lane) {
  const double incl = warp_incl_scan_d(v, lane);
  const double up = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane == 0 ? 0.0 : up;
}

// torch.nan_to_num defaults: nan -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX
__device__ __forceinline__ float nan_to_num(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return x;
}

}  // namespace kp
