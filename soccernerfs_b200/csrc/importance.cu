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


// ISG = DynamicDataset.compute_isg (NS/data/datasets/dynamic_dataset.py:215-326): per camera the per-pixel, per-channel
// MEDIAN over that camera's frames (torch.median: the lower of the two middle values for an even count), then per image
// the Geman-McClure residual  psi = d^2 / (d^2 + gamma^2),  d = image - median,  map = (1/3) * (psi_r + psi_g + psi_b), fp16.
// Kernel 1: one thread per (camera, pixel, channel) insertion-sorts the camera's <= kIsgMaxFrames values in local memory.
// Kernel 2: one thread per (image, pixel).  Same IEEE operations in the same order as the reference's torch ops.
constexpr int kIsgMaxFrames = 256;

__global__ void __launch_bounds__(128) isg_median_kernel(const float* __restrict__ images, int64_t HW3, const int32_t* __restrict__ cam_offsets,
                                                        const int32_t* __restrict__ cam_images, int n_cams, float* __restrict__ median) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n_cams * HW3) return;
  const int cam = (int)(idx / HW3);
  const int64_t e = idx % HW3;
  const int begin = cam_offsets[cam], n = cam_offsets[cam + 1] - begin;
  float v[kIsgMaxFrames];
  for (int i = 0; i < n; ++i) {  // insertion sort, ascending
    const float x = images[(int64_t)cam_images[begin + i] * HW3 + e];
    int j = i;
    while (j > 0 && v[j - 1] > x) {
      v[j] = v[j - 1];
      --j;
    }
    v[j] = x;
  }
  median[idx] = n > 0 ? v[(n - 1) / 2] : 0.f;
}

__global__ void __launch_bounds__(256) isg_map_kernel(const float* __restrict__ images, int B, int64_t HW, const int32_t* __restrict__ image_cam,
                                                     const float* __restrict__ median, float gamma_sq, __half* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * HW) return;
  const int i = (int)(idx / HW);
  const int64_t pix = idx % HW;
  const float* cur = images + ((int64_t)i * HW + pix) * 3;
  const float* med = median + ((int64_t)image_cam[i] * HW + pix) * 3;
  float psi[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float d = __fsub_rn(cur[c], med[c]);
    const float sq = __fmul_rn(d, d);
    psi[c] = __fdiv_rn(sq, __fadd_rn(sq, gamma_sq));
  }
  const float sum = __fadd_rn(__fadd_rn(psi[0], psi[1]), psi[2]);
  out[idx] = __float2half_rn(__fmul_rn((float)(1.0 / 3), sum));
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

extern "C" int kp_isg_map(const float* images, int B, int64_t HW, const int32_t* cam_offsets, const int32_t* cam_images,
                          const int32_t* image_cam, int n_cams, int max_frames_per_cam, float gamma_sq, float* median_scratch,
                          void* out_fp16, void* stream) {
  if (B == 0 || HW == 0) return 0;
  KP_CHECK(images && cam_offsets && cam_images && image_cam && median_scratch && out_fp16, "isg_map: NULL argument");
  KP_CHECK(B > 0 && HW > 0 && n_cams > 0, "isg_map: B=%d HW=%lld n_cams=%d", B, (long long)HW, n_cams);
  KP_CHECK(max_frames_per_cam >= 1 && max_frames_per_cam <= kIsgMaxFrames, "isg_map: %d frames of one camera (at most %d supported)",
           max_frames_per_cam, kIsgMaxFrames);
  cudaStream_t st = as_stream(stream);
  const int64_t n_med = (int64_t)n_cams * HW * 3;
  isg_median_kernel<<<(unsigned)ceil_div(n_med, 128), 128, 0, st>>>(images, HW * 3, cam_offsets, cam_images, n_cams, median_scratch);
  kp::g_launches += 1;
  isg_map_kernel<<<(unsigned)ceil_div((int64_t)B * HW, 256), 256, 0, st>>>(images, B, HW, image_cam, median_scratch, gamma_sq,
                                                                          reinterpret_cast<__half*>(out_fp16));
  KP_LAUNCH_CHECK("isg_map");
  return 0;
}
