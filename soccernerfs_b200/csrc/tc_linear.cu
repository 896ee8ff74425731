// Dense layer on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate via 3xTF32.
//
//   Y[M,N] = act( X[M,K] * W[N,K]^T )        X, W, Y fp32 row-major, bias-free (tcnn FullyFusedMLP semantics)
//
// The decoders' parity bar is 1e-4 relative in fp32, which a single TF32 pass (10-bit mantissa) cannot meet, so each
// operand is split x = hi + lo (both exactly representable in TF32, cvt.rna) and the product is accumulated as
// hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator (error ~2^-21, the dropped lo*lo term).  The MLP is tiny in
// FLOPs, so the 3x MMA count is irrelevant; what matters is that no FFMA / LDS issue slots are spent on it.
//
// One CTA (4 warps) per 128-row tile:
//   1. all threads stage the X tile and W into shared memory as K-major SWIZZLE_128B UMMA operands
//      (blocks of [rows x 32 tf32]; 8-row x 128-byte atoms, 16-byte chunk c of row r stored at chunk c ^ (r & 7));
//   2. one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N, K=8 per instruction) into TMEM and
//      commits to an mbarrier;
//   3. each warp reads its 32 TMEM lanes (= 32 rows) with tcgen05.ld.32x32b, applies the activation, stores Y.
#include "common.cuh"

namespace kp {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, dense 8-row atoms (SBO = 1024 B), version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major; canonical 1)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

// kind::tf32 instruction descriptor: D=F32, A=B=TF32, both K-major, M x N.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

// Stage a [rows x K] fp32 row-major matrix (leading dim ld, rows >= n_valid and cols >= k_valid zero-filled) as hi/lo
// TF32 operands: K/32 blocks of [ROWS x 32], each block ROWS/8 atoms of 1024 B.
template <int ROWS>
__device__ __forceinline__ void stage_operand(const float* __restrict__ src, int64_t ld, int n_valid, int k_valid, int K,
                                              float* __restrict__ s_hi, float* __restrict__ s_lo) {
  const int chunks_per_row = K / 4;  // 16-byte chunks
  for (int idx = threadIdx.x; idx < ROWS * chunks_per_row; idx += blockDim.x) {
    const int row = idx / chunks_per_row, ch = idx % chunks_per_row;
    const int kb = ch / 8, c = ch % 8;  // 32-column block, chunk within the 128-byte row
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = ch * 4 + e;
      v[e] = (row < n_valid && k < k_valid) ? __ldg(src + (int64_t)row * ld + k) : 0.f;
    }
    float4 hi, lo;
    hi.x = to_tf32(v[0]); hi.y = to_tf32(v[1]); hi.z = to_tf32(v[2]); hi.w = to_tf32(v[3]);
    lo.x = to_tf32(v[0] - hi.x); lo.y = to_tf32(v[1] - hi.y); lo.z = to_tf32(v[2] - hi.z); lo.w = to_tf32(v[3] - hi.w);
    const int off = kb * (ROWS * 32) + (row >> 3) * 256 + (row & 7) * 32 + ((c ^ (row & 7)) << 2);  // in floats
    *reinterpret_cast<float4*>(s_hi + off) = hi;
    *reinterpret_cast<float4*>(s_lo + off) = lo;
  }
}

template <int N_PAD>  // N padded to a multiple of 16 (UMMA N for M=128), <= 64
__global__ void __launch_bounds__(128) tc_linear_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W,
                                                        int64_t ldw, float* __restrict__ Y, int64_t ldy, int64_t M, int N,
                                                        int K_valid, int K, int act) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve (1024-byte aligned blocks): A_hi, A_lo [128 x K], B_hi, B_lo [N_PAD x K]
  float* a_hi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* a_lo = a_hi + 128 * K;
  float* b_hi = a_lo + 128 * K;
  float* b_lo = b_hi + N_PAD * K;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * 128;
  const int rows_valid = (int)min((int64_t)128, M - row0);

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM allocation: one warp, power-of-two columns >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  stage_operand<128>(X + row0 * ldx, ldx, rows_valid, K_valid, K, a_hi, a_lo);
  stage_operand<N_PAD>(W, ldw, N, K_valid, K, b_hi, b_lo);
  // generic-proxy smem writes -> visible to the tensor core's async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, N_PAD);
    const uint32_t a_addr[2] = {smem_u32(a_hi), smem_u32(a_lo)};
    const uint32_t b_addr[2] = {smem_u32(b_hi), smem_u32(b_lo)};
    const int term_a[3] = {0, 1, 0}, term_b[3] = {0, 0, 1};  // hi*hi + lo*hi + hi*lo
    uint32_t accumulate = 0;
    for (int t = 0; t < 3; ++t) {
      for (int kb = 0; kb < K / 32; ++kb) {
        const uint64_t ad = umma_desc_k_sw128(a_addr[term_a[t]] + kb * (128 * 32 * 4));
        const uint64_t bd = umma_desc_k_sw128(b_addr[term_b[t]] + kb * (N_PAD * 32 * 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle row
          umma_tf32(tmem_d, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, accumulate);
          accumulate = 1;
        }
      }
    }
    // arrive on the mbarrier when all MMAs above have completed (implicit tcgen05.fence::before_thread_sync)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar)) : "memory");
  }
  // wait for the accumulator
  {
    uint32_t done = 0;
    const uint32_t bar = smem_u32(&mma_bar);
    while (!done) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(bar), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows row0 + 32w + lane
  uint32_t acc[N_PAD];
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int c0 = 0; c0 < N_PAD; c0 += 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(acc[c0 + 0]), "=r"(acc[c0 + 1]), "=r"(acc[c0 + 2]), "=r"(acc[c0 + 3]), "=r"(acc[c0 + 4]), "=r"(acc[c0 + 5]),
          "=r"(acc[c0 + 6]), "=r"(acc[c0 + 7]), "=r"(acc[c0 + 8]), "=r"(acc[c0 + 9]), "=r"(acc[c0 + 10]), "=r"(acc[c0 + 11]),
          "=r"(acc[c0 + 12]), "=r"(acc[c0 + 13]), "=r"(acc[c0 + 14]), "=r"(acc[c0 + 15])
        : "r"(taddr + (uint32_t)c0));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int row = warp * 32 + lane;
  if (row < rows_valid) {
    float* dst = Y + (row0 + row) * ldy;
#pragma unroll
    for (int j = 0; j < N_PAD; ++j) {
      if (j < N) {
        float v = __uint_as_float(acc[j]);
        if (act == ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
        dst[j] = v;
      }
    }
  }
  // release TMEM
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(64));
}

}  // namespace kp

using namespace kp;

// Y[M,N] = act(X[M,K] W[N,K]^T) on tcgen05.  Supported: K <= 128 (padded up to a multiple of 32), N <= 64.
extern "C" int kp_tc_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, float* Y, int64_t ldy, int64_t M,
                                int N, int K, int act, void* stream) {
  if (M == 0) return 0;
  KP_CHECK(X && W && Y && N >= 1 && N <= 64 && K >= 1 && K <= 128, "tc_linear_fwd: unsupported shape N=%d K=%d", N, K);
  KP_CHECK(act >= 0 && act <= 2, "tc_linear_fwd: act=%d", act);
  const int Kp = (K + 31) / 32 * 32;
  const int Np = N <= 16 ? 16 : (N <= 32 ? 32 : 64);
  const size_t smem = (size_t)(2 * 128 * Kp + 2 * Np * Kp) * sizeof(float) + 1024;
  const unsigned grid = (unsigned)ceil_div(M, 128);
  cudaStream_t st = as_stream(stream);
#define KP_TC_LAUNCH(NP)                                                                                      \
  do {                                                                                                        \
    cudaFuncSetAttribute(tc_linear_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    tc_linear_kernel<NP><<<grid, 128, smem, st>>>(X, ldx, W, ldw, Y, ldy, M, N, K, Kp, act);                   \
  } while (0)
  if (Np == 16) KP_TC_LAUNCH(16);
  else if (Np == 32) KP_TC_LAUNCH(32);
  else KP_TC_LAUNCH(64);
#undef KP_TC_LAUNCH
  KP_LAUNCH_CHECK("tc_linear_fwd");
  return 0;
}
