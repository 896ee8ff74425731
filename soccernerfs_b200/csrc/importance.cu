// Importance-sampling weight maps on the device (sm_100a): IST = temporal-difference map of
// DynamicDataset.compute_ist (NS/data/datasets/dynamic_dataset.py:328-470).
// For image i: max over its temporal neighbours j (same camera, 0.01 < |t_j - t_i| <= ist_range; the lists are built
// on the host from the camera ids / times) of |img_i - img_j| per channel, mean over the 3 channels (sum left to
// right, then a true division by 3 like torch's CPU mean), values <= alpha (0.15: camera shake / noise) zeroed, fp16.
// An image without neighbours gets a uniform map of ones.  One thread per (image, pixel); bit-identical to the
// reference's torch ops.
#include <cuda_fp16.h>

#include "common.cuh"

namespace kp {

__global__ void __launch_bounds__(256) ist_map_kernel(const float* __restrict__ images, int B, int64_t HW,
                                                     const int32_t* __restrict__ nbr_offsets,
                                                     const int32_t* __restrict__ nbrs, float alpha,
                                                     __half* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * HW) return;
  const int i = (int)(idx / HW);
  const int64_t pix = idx % HW;
  const int begin = nbr_offsets[i], end = nbr_offsets[i + 1];
  if (begin == end) {
    out[idx] = __float2half_rn(1.0f);
    return;
  }
  const float* cur = images + ((int64_t)i * HW + pix) * 3;
  const float c0 = cur[0], c1 = cur[1], c2 = cur[2];
  float m0 = 0.f, m1 = 0.f, m2 = 0.f;
  for (int k = begin; k < end; ++k) {
    const float* other = images + ((int64_t)nbrs[k] * HW + pix) * 3;
    m0 = fmaxf(m0, fabsf(__fsub_rn(c0, other[0])));
    m1 = fmaxf(m1, fabsf(__fsub_rn(c1, other[1])));
    m2 = fmaxf(m2, fabsf(__fsub_rn(c2, other[2])));
  }
  const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(m0, m1), m2), 3.0f);
  out[idx] = __float2half_rn(mean > alpha ? mean : 0.0f);
}

}  // namespace kp

using namespace kp;

extern "C" int kp_ist_map(const float* images, int B, int64_t HW, const int32_t* nbr_offsets, const int32_t* nbrs,
                          float alpha, void* out_fp16, void* stream) {
  if (B == 0 || HW == 0) return 0;
  KP_CHECK(images != nullptr && nbr_offsets != nullptr && out_fp16 != nullptr, "ist_map: NULL argument");
  KP_CHECK(B > 0 && HW > 0, "ist_map: B=%d HW=%lld", B, (long long)HW);
  const int64_t total = (int64_t)B * HW;
  ist_map_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(images, B, HW, nbr_offsets, nbrs, alpha,
                                                                                reinterpret_cast<__half*>(out_fp16));
  KP_LAUNCH_CHECK("ist_map");
  return 0;
}
