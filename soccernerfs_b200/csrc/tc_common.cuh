// tcgen05 / TMEM / UMMA-descriptor helpers shared by the tensor-core decoder kernels (tc_linear.cu, decoder_fused.cu).
#pragma once
#include "common.cuh"

namespace kp {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// UMMA shared-memory descriptor (sm_100 version 1), SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);         // start address, 16-byte units
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;                           // descriptor version
  d |= (uint64_t)layout_type << 61;                 // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// K-major view of a staged tile with `rows` rows: 32-col block kb, 8-col step ks.  LBO unused (canonical 1).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int rows, int kb, int ks) {
  return umma_desc(base + kb * rows * 128 + ks * 32, 16, 1024, 2);
}
// MN-major view: K-step r8 = rows [8*r8, 8*r8+8) = two 4-row atoms 512 bytes apart (SBO); MN blocks of 32 cols are
// rows*128 bytes apart (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int rows, int r8) {
  return umma_desc(base + r8 * 1024, rows * 128, 512, 1);
}

// kind::tf32 instruction descriptor: D=F32, A=B=TF32, M x N, operand majors (0 = K-major, 1 = MN-major).
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t b = smem_u32(bar);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// Stage a [ROWS x cols_pad] tile from a row-major fp32 matrix (rows >= rows_valid / cols >= cols_valid zero-filled)
// as hi / lo TF32 operands in the blocked swizzled layout described above.
template <int MN_VIEW>
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int64_t ld, int rows, int rows_valid,
                                           int cols_valid, int cols_pad, float* __restrict__ s_hi, float* __restrict__ s_lo) {
  constexpr int U = 8;                      // independent 16-byte loads in flight per thread
  const int chunks_per_row = cols_pad / 4;  // 16-byte chunks
  const int total = rows * chunks_per_row;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (int base = threadIdx.x; base < total; base += blockDim.x * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * blockDim.x;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total) {
        const int row = idx / chunks_per_row, ch = idx % chunks_per_row;
        if (row < rows_valid) {
          const float* p = src + (int64_t)row * ld + ch * 4;
          if (vec && ch * 4 + 4 <= cols_valid) {
            v[u] = __ldg(reinterpret_cast<const float4*>(p));
          } else {
            if (ch * 4 + 0 < cols_valid) v[u].x = __ldg(p + 0);
            if (ch * 4 + 1 < cols_valid) v[u].y = __ldg(p + 1);
            if (ch * 4 + 2 < cols_valid) v[u].z = __ldg(p + 2);
            if (ch * 4 + 3 < cols_valid) v[u].w = __ldg(p + 3);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * blockDim.x;
      if (idx >= total) break;
      const int row = idx / chunks_per_row, ch = idx % chunks_per_row;
      const int kb = ch >> 3, c = ch & 7;
      float4 hi, lo;
      hi.x = to_tf32(v[u].x); hi.y = to_tf32(v[u].y); hi.z = to_tf32(v[u].z); hi.w = to_tf32(v[u].w);
      lo.x = to_tf32(v[u].x - hi.x); lo.y = to_tf32(v[u].y - hi.y); lo.z = to_tf32(v[u].z - hi.z); lo.w = to_tf32(v[u].w - hi.w);
      const int cs = MN_VIEW ? ((((c >> 1) ^ (row & 3)) << 1) | (c & 1)) : (c ^ (row & 7));  // swizzled 16-byte chunk
      const int off = kb * (rows * 32) + row * 32 + (cs << 2);  // in floats
      *reinterpret_cast<float4*>(s_hi + off) = hi;
      *reinterpret_cast<float4*>(s_lo + off) = lo;
    }
  }
}

// Two-phase staging of a 128-row activation tile so that the global loads of tile i+1 can be issued BEFORE the MMA /
// epilogue of tile i and their latency hides behind them: tile_load keeps up to 16 x 16 bytes per thread in registers
// (128 rows x 128 cols with 256 threads), tile_store converts to hi/lo TF32 and writes the swizzled operand.
// COLS_PAD (32 | 64 | 128) is a template parameter so that all index arithmetic folds to constants: thread t always
// owns 16-byte chunk (t % CPR) of rows (t / CPR) + u * (256 / CPR); since 256/CPR is a multiple of 8 the swizzle
// phase (row & 7) is the same for every u.  hi = x with the 13 low mantissa bits cleared (what the tensor core reads
// of an fp32 word anyway), lo = x - hi (exact); the tensor core's own truncation of lo costs 2^-22 relative.
template <int COLS_PAD>
struct TileRegs {
  static constexpr int CPR = COLS_PAD / 4;  // chunks per row
  static constexpr int U = CPR / 2;         // chunks per thread (128 rows, 256 threads)
  float4 v[U];
};
template <int COLS_PAD>
__device__ __forceinline__ void tile_load(TileRegs<COLS_PAD>& t, const float* __restrict__ src, int64_t ld, int rows_valid,
                                          int cols_valid) {
  constexpr int CPR = COLS_PAD / 4, U = CPR / 2, ROWSTEP = 256 / CPR;
  const int r0 = threadIdx.x / CPR, ch = threadIdx.x % CPR;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (ch * 4 + 4 <= cols_valid);
  const float* p = src + (int64_t)r0 * ld + ch * 4;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int row = r0 + u * ROWSTEP;
    t.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows_valid) {
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
// hi part of the split: the low 13 mantissa bits cleared -- exactly what kind::tf32 reads of an fp32 word, in ONE LOP3
// (cvt.rna.tf32 is emulated with ~5 instructions per element on sm_100a and was 16.7 % of the staging kernels' instructions)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
template <int MN_VIEW, int COLS_PAD>
__device__ __forceinline__ void tile_store(const TileRegs<COLS_PAD>& t, float* __restrict__ s_hi, float* __restrict__ s_lo) {
  constexpr int CPR = COLS_PAD / 4, U = CPR / 2, ROWSTEP = 256 / CPR;
  const int r0 = threadIdx.x / CPR, ch = threadIdx.x % CPR;
  const int kb = ch >> 3, c = ch & 7;
  const int cs = MN_VIEW ? ((((c >> 1) ^ (r0 & 3)) << 1) | (c & 1)) : (c ^ (r0 & 7));  // (row & 7) == (r0 & 7) for all u
  const int off0 = kb * (128 * 32) + r0 * 32 + (cs << 2);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float4 x = t.v[u];
    float4 hi, lo;
    hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
    lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
    const int off = off0 + u * ROWSTEP * 32;
    *reinterpret_cast<float4*>(s_hi + off) = hi;
    *reinterpret_cast<float4*>(s_lo + off) = lo;
  }
}

__device__ __forceinline__ float* align1024(uint8_t* p) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, int warp) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(128));
}
__device__ __forceinline__ void publish_smem_and_sync() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async (tensor core) proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

}  // namespace kp
