// Dense layers of the decoders on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate via 3xTF32.
//
//   forward        Y[M,N]   = act( X[M,K] W[N,K]^T )
//   backward data  dX[M,K]  = ( dY[M,N] W[N,K] ) * (aux > 0)
//   backward wgt   dW[N,K] += dY[M,N]^T X[M,K]
//
// all fp32 row-major, bias-free (tcnn FullyFusedMLP semantics, NS/fields/kplanes_field.py:249-273).  The decoders'
// parity bar is 1e-4 relative in fp32, which a single TF32 pass (10-bit mantissa) cannot meet, so each operand is
// split x = hi + lo (hi = cvt.rna.tf32(x), lo = x - hi exactly) and all four partial products
// lo*lo + lo*hi + hi*lo + hi*hi are accumulated in the fp32 TMEM accumulator; the only loss is the tensor core's
// truncation of lo to 11 bits (~2^-23 relative), i.e. fp32-class accuracy -- which matters because a ReLU whose
// pre-activation is rounded across zero flips a sample's whole gradient.  The MLPs are tiny in FLOPs, so the 4x MMA
// count is irrelevant; what matters is that no FFMA / LDS issue slots are spent on them.
//
// Shared-memory operand tiles: a [ROWS x COLS] fp32 tile is stored as COLS/32 blocks of [ROWS x 32]; each block is
// ROWS/8 atoms of 8 rows x 128 bytes with the 128-byte swizzle (16-byte chunk c of row r at chunk c ^ (r & 7)).
// Two UMMA views of such a tile are used, so no transposed copy of any operand is ever made:
//   K-major  view (SWIZZLE_128B):          MN = tile rows, K = tile cols  (forward A = X tile, B = W)
//   MN-major view (SWIZZLE_128B_BASE32B):  MN = tile cols, K = tile rows  (backward-data B = W "transposed";
//                  weight-gradient A = X tile, B = dY tile with K = the 128 samples of the tile)
// For 32-bit operands the tensor core only accepts the 32-byte-base swizzle in the MN-major view (atoms of 4 rows x
// 128 bytes, 32-byte chunk q of row r at chunk q ^ (r & 3)), so a tile is staged in the swizzle of the view it is
// consumed in.
#include <stdlib.h>

#include "tc_common.cuh"

namespace kp {

// ---------------------------------------------------------------------------------------------------------------
// Row-tile GEMM:  OUT[128 rows, N] = A[128, R] * B      (R = reduction depth, padded to 32; N padded to 32 for MN-B)
//   B_MN == 0 (forward):        B = W [N x R]  K-major            ; epilogue act
//   B_MN == 1 (backward data):  B = W [R x N]  MN-major view      ; epilogue (aux > 0) mask
// ---------------------------------------------------------------------------------------------------------------
template <int B_MN, int R_pad>
__global__ void __launch_bounds__(256) tc_rowtile_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W,
                                                         int64_t ldw, float* __restrict__ OUT, int64_t ldo, int64_t M, int N,
                                                         int N_pad, int R, int act, const float* __restrict__ aux,
                                                         int64_t ldaux) {
  extern __shared__ uint8_t smem_raw[];
  const int b_rows = B_MN ? R_pad : N_pad, b_cols = B_MN ? N_pad : R_pad;
  float* a_hi = align1024(smem_raw);
  float* a_lo = a_hi + 128 * R_pad;
  float* b_hi = a_lo + 128 * R_pad;
  float* b_lo = b_hi + b_rows * b_cols;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tmem_alloc(&tmem_slot, warp);
  // weights are staged ONCE per (persistent) CTA
  if (B_MN) stage_tile<1>(W, ldw, b_rows, R, N, b_cols, b_hi, b_lo);
  else stage_tile<0>(W, ldw, b_rows, N, R, b_cols, b_hi, b_lo);
  const uint32_t idesc = umma_idesc_tf32(128, N_pad, 0, B_MN);
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const int64_t n_tiles = (M + 127) / 128;
  uint32_t parity = 0;
  TileRegs<R_pad> pre;
  if ((int64_t)blockIdx.x < n_tiles)
    tile_load<R_pad>(pre, A + (int64_t)blockIdx.x * 128 * lda, lda, (int)min((int64_t)128, M - (int64_t)blockIdx.x * 128), R);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    const int rows_valid = (int)min((int64_t)128, M - row0);
    tile_store<0, R_pad>(pre, a_hi, a_lo);
    publish_smem_and_sync();
    {  // issue the next tile's global loads now; they complete while the MMAs and the epilogue below run
      const int64_t nxt = tile + gridDim.x;
      if (nxt < n_tiles) tile_load<R_pad>(pre, A + nxt * 128 * lda, lda, (int)min((int64_t)128, M - nxt * 128), R);
    }
    const uint32_t tmem_d = tmem_slot;
    if (threadIdx.x == 0) {
      // One thread issues every MMA: keep its instruction stream short -- base descriptors once, then only the
      // 14-bit start-address field advances by compile-time constants (fully unrolled).
      const uint64_t a_d[2] = {desc_kmajor(smem_u32(a_hi), 128, 0, 0), desc_kmajor(smem_u32(a_lo), 128, 0, 0)};
      const uint64_t b_d[2] = {B_MN ? desc_mnmajor(smem_u32(b_hi), b_rows, 0) : desc_kmajor(smem_u32(b_hi), b_rows, 0, 0),
                               B_MN ? desc_mnmajor(smem_u32(b_lo), b_rows, 0) : desc_kmajor(smem_u32(b_lo), b_rows, 0, 0)};
      const uint32_t b_blk = (uint32_t)(b_rows * 128) >> 4;  // K-major B: 16-byte units between 32-col blocks
#pragma unroll
      for (int t = 0; t < 4; ++t) {  // lo*lo + lo*hi + hi*lo + hi*hi (small terms first)
        const uint64_t ad0 = a_d[t <= 1], bd0 = b_d[t == 0 || t == 2];
#pragma unroll
        for (int k8 = 0; k8 < R_pad / 8; ++k8) {
          const uint32_t a_off = (uint32_t)(((k8 >> 2) * 128 * 128 + (k8 & 3) * 32) >> 4);
          const uint32_t b_off = B_MN ? (uint32_t)((k8 * 1024) >> 4) : (uint32_t)(k8 >> 2) * b_blk + (uint32_t)(((k8 & 3) * 32) >> 4);
          umma_tf32(tmem_d, ad0 + a_off, bd0 + b_off, idesc, (t | k8) != 0);
        }
      }
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warps w and w+4 share TMEM lane quadrant (w & 3) and split the columns
    const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16);
    for (int c0 = half * 16; c0 < N_pad; c0 += 32) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < rows_valid && c0 < N) {
        float* dst = OUT + (row0 + row) * ldo + c0;
        const float* ax = aux ? aux + (row0 + row) * ldaux + c0 : nullptr;
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          x[j] = __uint_as_float(v[j]);
          if (!B_MN) {
            if (act == ACT_RELU) x[j] = fmaxf(x[j], 0.f);
            else if (act == ACT_SIGMOID) x[j] = 1.f / (1.f + expf(-x[j]));
          }
        }
        const bool full = (c0 + 16 <= N) && ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(OUT) & 15) == 0);
        if (full && (ax == nullptr || ((ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0))) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o4 = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            if (B_MN && ax != nullptr) {
              const float4 m4 = __ldg(reinterpret_cast<const float4*>(ax + 4 * q));
              if (!(m4.x > 0.f)) o4.x = 0.f;
              if (!(m4.y > 0.f)) o4.y = 0.f;
              if (!(m4.z > 0.f)) o4.z = 0.f;
              if (!(m4.w > 0.f)) o4.w = 0.f;
            }
            *reinterpret_cast<float4*>(dst + 4 * q) = o4;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c0 + j < N) {
              float o1 = x[j];
              if (B_MN && ax != nullptr && !(ax[j] > 0.f)) o1 = 0.f;
              dst[j] = o1;
            }
          }
        }
      }
    }
    // all TMEM reads of this tile done before the next tile's MMAs overwrite the accumulator
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  tmem_free(tmem_slot, warp);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient, persistent: each CTA walks its 128-sample tiles, accumulating
//   D[k, n] += sum_{s in tile} X[s, k] * dY[s, n]         (D = dW^T, M = 128 >= K_in, N = N_out padded to 32)
// in ONE TMEM accumulator (both operands are MN-major views with K = samples), then adds D into dW once.
// ---------------------------------------------------------------------------------------------------------------
template <int K_pad, int N_pad>
__global__ void __launch_bounds__(256) tc_wgrad_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY,
                                                       int64_t lddy, float* __restrict__ dW, int64_t lddw, int64_t M,
                                                       int K_in, int N_out) {
  extern __shared__ uint8_t smem_raw[];
  float* x_hi = align1024(smem_raw);
  float* x_lo = x_hi + 128 * K_pad;
  float* y_hi = x_lo + 128 * K_pad;
  float* y_lo = y_hi + 128 * N_pad;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tmem_alloc(&tmem_slot, warp);
  const int64_t n_tiles = (M + 127) / 128;
  uint32_t parity = 0, accumulate = 0;
  uint32_t tmem_d = 0;
  TileRegs<K_pad> px;
  TileRegs<N_pad> py;
  if ((int64_t)blockIdx.x < n_tiles) {
    const int rv = (int)min((int64_t)128, M - (int64_t)blockIdx.x * 128);
    tile_load<K_pad>(px, X + (int64_t)blockIdx.x * 128 * ldx, ldx, rv, K_in);
    tile_load<N_pad>(py, dY + (int64_t)blockIdx.x * 128 * lddy, lddy, rv, N_out);
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    tile_store<1, K_pad>(px, x_hi, x_lo);
    tile_store<1, N_pad>(py, y_hi, y_lo);
    publish_smem_and_sync();
    {
      const int64_t nxt = tile + gridDim.x;
      if (nxt < n_tiles) {
        const int rv = (int)min((int64_t)128, M - nxt * 128);
        tile_load<K_pad>(px, X + nxt * 128 * ldx, ldx, rv, K_in);
        tile_load<N_pad>(py, dY + nxt * 128 * lddy, lddy, rv, N_out);
      }
    }
    tmem_d = tmem_slot;
    if (threadIdx.x == 0) {
      // M = 128 rows of D; when K_pad < 128 the MN blocks beyond the X tile read whatever follows in shared memory:
      // those D rows (k >= K_in) are never read back.
      const uint32_t idesc = umma_idesc_tf32(128, N_pad, 1, 1);
      const uint64_t a_d[2] = {desc_mnmajor(smem_u32(x_hi), 128, 0), desc_mnmajor(smem_u32(x_lo), 128, 0)};
      const uint64_t b_d[2] = {desc_mnmajor(smem_u32(y_hi), 128, 0), desc_mnmajor(smem_u32(y_lo), 128, 0)};
#pragma unroll
      for (int t = 0; t < 4; ++t) {  // lo*lo + lo*hi + hi*lo + hi*hi
        const uint64_t ad0 = a_d[t <= 1], bd0 = b_d[t == 0 || t == 2];
#pragma unroll
        for (int r8 = 0; r8 < 16; ++r8) {  // 128 samples = 16 K-steps of 8 rows (1024 bytes each)
          umma_tf32(tmem_d, ad0 + (uint32_t)(r8 * 64), bd0 + (uint32_t)(r8 * 64), idesc, accumulate | (uint32_t)((t | r8) != 0));
        }
      }
      accumulate = 1;
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, parity);  // operands consumed: the staging buffers may be overwritten
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (blockIdx.x < n_tiles) {
    const int quad = warp & 3, half = warp >> 2;
    const int k = quad * 32 + lane;  // D row = input feature index
    const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16);
    for (int c0 = half * 16; c0 < N_pad; c0 += 32) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (k < K_in) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < N_out) red_add_f32(dW + (int64_t)(c0 + j) * lddw + k, __uint_as_float(v[j]));
      }
    }
  }
  tmem_free(tmem_slot, warp);
}

static int pad_dim(int x) { return x <= 32 ? 32 : (x <= 64 ? 64 : 128); }  // operand tile widths: 32 | 64 | 128

template <typename Kern>
static unsigned persistent_grid(Kern kern, int64_t n_tiles, size_t smem) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // CTAs per SM from the kernel's own resources.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for these
  // kernels whatever their footprint -- measured on B200, driver 580 -- so it is not used.)
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {
    int smem_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * 256;
    const size_t smem_cta = smem + fa.sharedSizeBytes + 1024;  // + the per-CTA reservation
    per_sm = std::min<int>(65536 / std::max(1, regs_cta), (int)((size_t)smem_sm / smem_cta));
  }
  if (per_sm < 1) per_sm = 1;
  if (getenv("KP_TC_MAX_CTAS") != nullptr) per_sm = std::min(per_sm, atoi(getenv("KP_TC_MAX_CTAS")));
  if (per_sm > 4) per_sm = 4;  // 4 x 128 TMEM columns
  if (getenv("KP_TC_DEBUG") != nullptr)
    fprintf(stderr, "[kp tc] smem=%zu B -> %d CTA/SM, grid %lld\n", smem, per_sm, (long long)std::min<int64_t>(n_tiles, (int64_t)sms * per_sm));
  return (unsigned)std::min<int64_t>(n_tiles, (int64_t)sms * per_sm);
}

template <int B_MN, int R_PAD>
static void launch_rowtile(const float* A, int64_t lda, const float* W, int64_t ldw, float* OUT, int64_t ldo, int64_t M, int N,
                           int N_pad, int R, int act, const float* aux, int64_t ldaux, cudaStream_t st) {
  const size_t smem = (size_t)(2 * 128 * R_PAD + 2 * N_pad * R_PAD) * sizeof(float) + 1024;
  auto kern = tc_rowtile_kernel<B_MN, R_PAD>;
  kern<<<persistent_grid(kern, ceil_div(M, 128), smem), 256, smem, st>>>(A, lda, W, ldw, OUT, ldo, M, N, N_pad, R, act, aux, ldaux);
}
template <int B_MN>
static void dispatch_rowtile(int R_pad, const float* A, int64_t lda, const float* W, int64_t ldw, float* OUT, int64_t ldo,
                             int64_t M, int N, int N_pad, int R, int act, const float* aux, int64_t ldaux, cudaStream_t st) {
  if (R_pad == 32) launch_rowtile<B_MN, 32>(A, lda, W, ldw, OUT, ldo, M, N, N_pad, R, act, aux, ldaux, st);
  else if (R_pad == 64) launch_rowtile<B_MN, 64>(A, lda, W, ldw, OUT, ldo, M, N, N_pad, R, act, aux, ldaux, st);
  else launch_rowtile<B_MN, 128>(A, lda, W, ldw, OUT, ldo, M, N, N_pad, R, act, aux, ldaux, st);
}

template <int K_PAD, int N_PAD>
static void launch_wgrad(const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dW, int64_t lddw, int64_t M, int K,
                         int N, cudaStream_t st) {
  // the A operand always spans 4 MN blocks (M = 128): keep 4 blocks of slack after x_lo inside the allocation
  size_t smem = (size_t)(2 * 128 * K_PAD + 2 * 128 * N_PAD) * sizeof(float);
  const size_t need = (size_t)(128 * K_PAD + 128 * 128) * sizeof(float);
  if (smem < need) smem = need;
  smem += 1024;
  auto kern = tc_wgrad_kernel<K_PAD, N_PAD>;
  kern<<<persistent_grid(kern, ceil_div(M, 128), smem), 256, smem, st>>>(X, ldx, dY, lddy, dW, lddw, M, K, N);
}
template <int K_PAD>
static void dispatch_wgrad(int N_pad, const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dW, int64_t lddw,
                           int64_t M, int K, int N, cudaStream_t st) {
  if (N_pad == 32) launch_wgrad<K_PAD, 32>(X, ldx, dY, lddy, dW, lddw, M, K, N, st);
  else if (N_pad == 64) launch_wgrad<K_PAD, 64>(X, ldx, dY, lddy, dW, lddw, M, K, N, st);
  else launch_wgrad<K_PAD, 128>(X, ldx, dY, lddy, dW, lddw, M, K, N, st);
}

}  // namespace kp

using namespace kp;

// Which shapes the tensor-core path covers (others use the SIMT SGEMM in mlp.cu): operand tiles must fit shared memory.
extern "C" int kp_tc_supported(int N, int K) {
  return (N >= 1 && N <= 128 && K >= 1 && K <= 128 && pad_dim(N) * pad_dim(K) <= 128 * 64) ? 1 : 0;
}

extern "C" int kp_tc_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, float* Y, int64_t ldy, int64_t M,
                                int N, int K, int act, void* stream) {
  if (M == 0) return 0;
  KP_CHECK(X && W && Y && kp_tc_supported(N, K), "tc_linear_fwd: unsupported shape N=%d K=%d", N, K);
  KP_CHECK(act >= 0 && act <= 2, "tc_linear_fwd: act=%d", act);
  dispatch_rowtile<0>(pad_dim(K), X, ldx, W, ldw, Y, ldy, M, N, pad_dim(N), K, act, nullptr, 0, as_stream(stream));
  KP_LAUNCH_CHECK("tc_linear_fwd");
  return 0;
}

// dX[M,K] = (dY[M,N] W[N,K]) masked by (aux[M,K] > 0) when aux != NULL.
extern "C" int kp_tc_linear_bwd_data(const float* dY, int64_t lddy, const float* W, int64_t ldw, float* dX, int64_t lddx,
                                     int64_t M, int N, int K, const float* aux, int64_t ldaux, void* stream) {
  if (M == 0) return 0;
  KP_CHECK(dY && W && dX && kp_tc_supported(N, K), "tc_linear_bwd_data: unsupported shape N=%d K=%d", N, K);
  // reduction over N_out (R), output width K_in
  dispatch_rowtile<1>(pad_dim(N), dY, lddy, W, ldw, dX, lddx, M, K, pad_dim(K), N, 0, aux, ldaux, as_stream(stream));
  KP_LAUNCH_CHECK("tc_linear_bwd_data");
  return 0;
}

// dW[N,K] += dY[M,N]^T X[M,K]
extern "C" int kp_tc_linear_bwd_weight(const float* dY, int64_t lddy, const float* X, int64_t ldx, float* dW, int64_t lddw,
                                       int64_t M, int N, int K, void* stream) {
  if (M == 0) return 0;
  KP_CHECK(dY && X && dW && N >= 1 && N <= 128 && K >= 1 && K <= 128 && pad_dim(N) + pad_dim(K) <= 192,
           "tc_linear_bwd_weight: unsupported shape N=%d K=%d", N, K);
  const int Kp = pad_dim(K), Np = pad_dim(N);
  cudaStream_t st = as_stream(stream);
  if (Kp == 32) dispatch_wgrad<32>(Np, X, ldx, dY, lddy, dW, lddw, M, K, N, st);
  else if (Kp == 64) dispatch_wgrad<64>(Np, X, ldx, dY, lddy, dW, lddw, M, K, N, st);
  else dispatch_wgrad<128>(Np, X, ldx, dY, lddy, dW, lddw, M, K, N, st);
  KP_LAUNCH_CHECK("tc_linear_bwd_weight");
  return 0;
}
