// Loss kernels (sm_100a).
//   kp_distortion_*  : lossfun_distortion, NS/model_components/losses.py:125-136 (O(S^2) per ray, one warp per ray)
//   kp_interlevel_*  : outer / lossfun_outer, losses.py:46-95 (two searchsorted(side="right") + cumsum differences)
//   kp_plane_reg_*   : compute_plane_tv / compute_plane_smoothness / |1 - t|, losses.py:356-452, as one streaming
//                      pass over a channel-last plane (float4 per thread), analytic gradients in the backward.
//   kp_adam_step     : torch.optim.Adam update over a flat buffer (NS/engine/optimizers.py:74-160).
#include "common.cuh"

namespace kp {

constexpr int kRaysPerBlock = 4;

__global__ void __launch_bounds__(32 * kRaysPerBlock) distortion_kernel(const float* __restrict__ sdist,
                                                                        const float* __restrict__ w, int64_t N, int S,
                                                                        const float* __restrict__ grad_per_ray,
                                                                        float* __restrict__ loss, float* __restrict__ gw) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kRaysPerBlock + warp;
  if (n >= N) return;
  float* s_ut = smem + (size_t)warp * 2 * S;
  float* s_w = s_ut + S;
  const float* t = sdist + n * (S + 1);
  for (int i = lane; i < S; i += 32) {
    s_ut[i] = (t[i + 1] + t[i]) * 0.5f;
    s_w[i] = w[n * S + i];
  }
  __syncwarp();
  float acc = 0.f;
  const float gn = grad_per_ray != nullptr ? grad_per_ray[n] : 0.f;
  for (int i = lane; i < S; i += 32) {
    const float ui = s_ut[i], wi = s_w[i];
    float inner = 0.f;
    for (int j = 0; j < S; ++j) inner = fmaf(s_w[j], fabsf(ui - s_ut[j]), inner);
    const float width = t[i + 1] - t[i];
    if (gw != nullptr) gw[n * S + i] = gn * (2.f * inner + 2.f * wi * width / 3.f);
    acc += wi * inner + wi * wi * width / 3.f;
  }
  acc = warp_sum(acc);
  if (loss != nullptr && lane == 0) loss[n] = acc;
}

// upper_bound over a sorted smem row: number of entries <= v  (torch.searchsorted side="right")
__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <bool BWD>
__global__ void __launch_bounds__(32 * kRaysPerBlock) interlevel_kernel(const float* __restrict__ c,
                                                                        const float* __restrict__ w,
                                                                        const float* __restrict__ cp,
                                                                        const float* __restrict__ wp,
                                                                        const float* __restrict__ gloss, int64_t N, int S,
                                                                        int Sp, float* __restrict__ loss,
                                                                        float* __restrict__ gwp) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kRaysPerBlock + warp;
  if (n >= N) return;
  float* s_cp = smem + (size_t)warp * 3 * (Sp + 1);  // [Sp+1]
  float* s_cy = s_cp + (Sp + 1);                     // [Sp+1] = [0, cumsum(wp)]
  float* s_e = s_cy + (Sp + 1);                      // [Sp+1] scatter buffer (bwd)
  for (int i = lane; i <= Sp; i += 32) {
    s_cp[i] = cp[n * (Sp + 1) + i];
    s_e[i] = 0.f;
  }
  // cumsum(wp): double accumulate, fp32 per element (losses.py:64)
  const int epl = (Sp + 31) / 32;
  const int i0 = lane * epl, i1 = min(Sp, i0 + epl);
  double part = 0.0;
  for (int i = i0; i < i1; ++i) part += (double)wp[n * Sp + i];
  double run = warp_excl_scan_d(part, lane);
  for (int i = i0; i < i1; ++i) {
    run += (double)wp[n * Sp + i];
    s_cy[i + 1] = (float)run;
  }
  if (lane == 0) s_cy[0] = 0.f;
  __syncwarp();
  for (int i = lane; i < S; i += 32) {
    const float c0 = c[n * (S + 1) + i], c1 = c[n * (S + 1) + i + 1], wi = w[n * S + i];
    int lo = upper_bound(s_cp, Sp, c0) - 1;      // t1_starts = cp[:-1]
    lo = min(max(lo, 0), Sp - 1);
    int hi = upper_bound(s_cp + 1, Sp, c1);      // t1_ends = cp[1:]
    hi = min(max(hi, 0), Sp - 1);
    const float w_outer = s_cy[hi + 1] - s_cy[lo];
    const float diff = fmaxf(wi - w_outer, 0.f);
    if (!BWD) {
      loss[n * S + i] = diff * diff / (wi + 1e-7f);
    } else {
      // d loss / d w_outer = -2 diff / (w + eps); d w_outer / d wp_k = [k <= hi] - [k < lo]
      const float G = gloss[n * S + i] * (-2.f * diff / (wi + 1e-7f));
      if (G != 0.f) {
        atomicAdd(&s_e[hi + 1], G);
        atomicAdd(&s_e[lo], -G);
      }
    }
  }
  if (BWD) {
    __syncwarp();
    // grad_wp[k] = sum_{j > k} E[j]: suffix scan over Sp+1 entries
    double p2 = 0.0;
    for (int i = i0; i < i1; ++i) p2 += (double)s_e[i + 1];
    const double incl = warp_incl_scan_d(p2, lane);
    const double total = __shfl_sync(0xffffffffu, incl, 31);
    double suffix = total - (incl - p2);  // sum over j >= i0+1
    for (int i = i0; i < i1; ++i) {
      gwp[n * Sp + i] = (float)suffix;
      suffix -= (double)s_e[i + 1];
    }
  }
}

// ---- plane regularisers -----------------------------------------------------------------------------
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float sq4(float4 a) { return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; }

__global__ void __launch_bounds__(256) plane_reg_fwd_kernel(const float* __restrict__ t, int H, int W, int C4,
                                                            uint32_t terms, double* __restrict__ sums) {
  const int64_t total = (int64_t)H * W * C4;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int64_t row = (int64_t)W * C4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(idx / row);
    const int wq = (int)(idx % row);
    const int wcol = wq / C4;
    const float4 v = ldg4(t + idx * 4);
    if ((terms & 1u) && h + 1 < H) s0 += sq4(sub4(ldg4(t + (idx + row) * 4), v));
    if ((terms & 2u) && wcol + 1 < W) s1 += sq4(sub4(ldg4(t + (idx + C4) * 4), v));
    if ((terms & 4u) && h + 2 < H) {
      const float4 v1 = ldg4(t + (idx + row) * 4), v2 = ldg4(t + (idx + 2 * row) * 4);
      s2 += sq4(sub4(sub4(v2, v1), sub4(v1, v)));
    }
    if (terms & 8u) s3 += fabsf(1.f - v.x) + fabsf(1.f - v.y) + fabsf(1.f - v.z) + fabsf(1.f - v.w);
  }
  __shared__ double red[4][8];
  double d[4] = {warp_sum_d((double)s0), warp_sum_d((double)s1), warp_sum_d((double)s2), warp_sum_d((double)s3)};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) red[k][warp] = d[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) a += red[threadIdx.x][wv];
    if ((terms >> threadIdx.x) & 1u) atomicAdd(&sums[threadIdx.x], a);
  }
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(256) plane_reg_bwd_kernel(const float* __restrict__ t, int H, int W, int C4,
                                                            const float* __restrict__ coef, uint32_t terms, int accumulate,
                                                            float* __restrict__ grad) {
  const float k0 = (terms & 1u) ? coef[0] : 0.f, k1 = (terms & 2u) ? coef[1] : 0.f;
  const float k2 = (terms & 4u) ? coef[2] : 0.f, k3 = (terms & 8u) ? coef[3] : 0.f;
  const int64_t total = (int64_t)H * W * C4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t row = (int64_t)W * C4;
  const int h = (int)(idx / row);
  const int wcol = (int)(idx % row) / C4;
  const float4 v = ldg4(t + idx * 4);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 zero = g;
  const bool need_h = (k0 != 0.f) || (k2 != 0.f);
  float4 up1 = zero, up2 = zero, dn1 = zero, dn2 = zero;  // t[h+1], t[h+2], t[h-1], t[h-2]
  if (need_h) {
    if (h + 1 < H) up1 = ldg4(t + (idx + row) * 4);
    if (h >= 1) dn1 = ldg4(t + (idx - row) * 4);
  }
  if (k0 != 0.f) {
    if (h >= 1) g = fma4(sub4(v, dn1), 2.f * k0, g);
    if (h + 1 < H) g = fma4(sub4(up1, v), -2.f * k0, g);
  }
  if (k1 != 0.f) {
    if (wcol >= 1) g = fma4(sub4(v, ldg4(t + (idx - C4) * 4)), 2.f * k1, g);
    if (wcol + 1 < W) g = fma4(sub4(ldg4(t + (idx + C4) * 4), v), -2.f * k1, g);
  }
  if (k2 != 0.f) {
    if (h + 2 < H) up2 = ldg4(t + (idx + 2 * row) * 4);
    if (h >= 2) dn2 = ldg4(t + (idx - 2 * row) * 4);
    // s_k = (t[k+2]-t[k+1]) - (t[k+1]-t[k]);  d/dt[h] = 2 (s_{h-2} - 2 s_{h-1} + s_h)
    if (h >= 2) g = fma4(sub4(sub4(v, dn1), sub4(dn1, dn2)), 2.f * k2, g);                  // s_{h-2}
    if (h >= 1 && h + 1 < H) g = fma4(sub4(sub4(up1, v), sub4(v, dn1)), -4.f * k2, g);     // s_{h-1}
    if (h + 2 < H) g = fma4(sub4(sub4(up2, up1), sub4(up1, v)), 2.f * k2, g);               // s_h
  }
  if (k3 != 0.f) {
    g.x -= k3 * sgn(1.f - v.x); g.y -= k3 * sgn(1.f - v.y); g.z -= k3 * sgn(1.f - v.z); g.w -= k3 * sgn(1.f - v.w);
  }
  float4* gp = reinterpret_cast<float4*>(grad + idx * 4);
  if (accumulate) {
    const float4 cur = *gp;
    g.x += cur.x; g.y += cur.y; g.z += cur.z; g.w += cur.w;
  }
  *gp = g;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr_over_bc1,
                                                   float beta1, float beta2, float eps, float wd, float inv_sqrt_bc2,
                                                   float grad_scale) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n) {
    float4 pp = *reinterpret_cast<float4*>(p + i4);
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4);
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = ga[k] * grad_scale + wd * pa[k];
      ma[k] = ma[k] + (gr - ma[k]) * (1.f - beta1);
      va[k] = va[k] * beta2 + (1.f - beta2) * gr * gr;
      pa[k] -= lr_over_bc1 * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + i4) = pp;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (int64_t i = i4; i < n; ++i) {
      float gr = g[i] * grad_scale + wd * p[i];
      m[i] = m[i] + (gr - m[i]) * (1.f - beta1);
      v[i] = v[i] * beta2 + (1.f - beta2) * gr * gr;
      p[i] -= lr_over_bc1 * m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps);
    }
  }
}

__global__ void repack_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W, int to_hwc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)C * H * W;
  if (idx >= total) return;
  // idx enumerates the destination
  if (to_hwc) {
    const int c = (int)(idx % C);
    const int64_t hw = idx / C;
    dst[idx] = src[(int64_t)c * H * W + hw];
  } else {
    const int64_t hw = idx % ((int64_t)H * W);
    const int c = (int)(idx / ((int64_t)H * W));
    dst[idx] = src[hw * C + c];
  }
}

}  // namespace kp

using namespace kp;

extern "C" int kp_distortion_fwd(const float* sdist, const float* w, int64_t N, int S, float* loss_per_ray, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(sdist && w && loss_per_ray && S >= 1, "distortion_fwd: bad arguments");
  const size_t smem = (size_t)kRaysPerBlock * 2 * S * sizeof(float);
  KP_CHECK(smem <= 48 * 1024, "distortion_fwd: S=%d too large", S);
  distortion_kernel<<<(unsigned)ceil_div(N, kRaysPerBlock), 32 * kRaysPerBlock, smem, as_stream(stream)>>>(
      sdist, w, N, S, nullptr, loss_per_ray, nullptr);
  KP_LAUNCH_CHECK("distortion_fwd");
  return 0;
}

extern "C" int kp_distortion_bwd(const float* sdist, const float* w, const float* grad_per_ray, int64_t N, int S,
                                 float* grad_w, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(sdist && w && grad_per_ray && grad_w && S >= 1, "distortion_bwd: bad arguments");
  const size_t smem = (size_t)kRaysPerBlock * 2 * S * sizeof(float);
  KP_CHECK(smem <= 48 * 1024, "distortion_bwd: S=%d too large", S);
  distortion_kernel<<<(unsigned)ceil_div(N, kRaysPerBlock), 32 * kRaysPerBlock, smem, as_stream(stream)>>>(
      sdist, w, N, S, grad_per_ray, nullptr, grad_w);
  KP_LAUNCH_CHECK("distortion_bwd");
  return 0;
}

extern "C" int kp_interlevel_fwd(const float* c, const float* w, const float* cp, const float* wp, int64_t N, int S,
                                 int Sp, float* loss, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(c && w && cp && wp && loss && S >= 1 && Sp >= 1, "interlevel_fwd: bad arguments");
  const size_t smem = (size_t)kRaysPerBlock * 3 * (Sp + 1) * sizeof(float);
  KP_CHECK(smem <= 48 * 1024, "interlevel_fwd: Sp=%d too large", Sp);
  interlevel_kernel<false><<<(unsigned)ceil_div(N, kRaysPerBlock), 32 * kRaysPerBlock, smem, as_stream(stream)>>>(
      c, w, cp, wp, nullptr, N, S, Sp, loss, nullptr);
  KP_LAUNCH_CHECK("interlevel_fwd");
  return 0;
}

extern "C" int kp_interlevel_bwd(const float* c, const float* w, const float* cp, const float* wp, const float* grad_loss,
                                 int64_t N, int S, int Sp, float* grad_wp, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(c && w && cp && wp && grad_loss && grad_wp && S >= 1 && Sp >= 1, "interlevel_bwd: bad arguments");
  const size_t smem = (size_t)kRaysPerBlock * 3 * (Sp + 1) * sizeof(float);
  KP_CHECK(smem <= 48 * 1024, "interlevel_bwd: Sp=%d too large", Sp);
  interlevel_kernel<true><<<(unsigned)ceil_div(N, kRaysPerBlock), 32 * kRaysPerBlock, smem, as_stream(stream)>>>(
      c, w, cp, wp, grad_loss, N, S, Sp, nullptr, grad_wp);
  KP_LAUNCH_CHECK("interlevel_bwd");
  return 0;
}

extern "C" int kp_plane_reg_fwd(const float* plane, int H, int W, int C, uint32_t terms, double* sums4, void* stream) {
  KP_CHECK(plane && sums4 && H >= 1 && W >= 1 && C >= 4 && C % 4 == 0, "plane_reg_fwd: bad arguments (C must be a multiple of 4)");
  const int64_t total = (int64_t)H * W * (C / 4);
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), 148 * 8);
  plane_reg_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(plane, H, W, C / 4, terms, sums4);
  KP_LAUNCH_CHECK("plane_reg_fwd");
  return 0;
}

extern "C" int kp_plane_reg_bwd(const float* plane, int H, int W, int C, const float* coef_dev4, uint32_t terms,
                                int accumulate, float* grad, void* stream) {
  KP_CHECK(plane && grad && coef_dev4 && H >= 1 && W >= 1 && C >= 4 && C % 4 == 0, "plane_reg_bwd: bad arguments");
  if ((terms & 15u) == 0 && accumulate) return 0;
  const int64_t total = (int64_t)H * W * (C / 4);
  plane_reg_bwd_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(plane, H, W, C / 4, coef_dev4, terms,
                                                                                      accumulate, grad);
  KP_LAUNCH_CHECK("plane_reg_bwd");
  return 0;
}

extern "C" int kp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                            void* stream) {
  if (n == 0) return 0;
  KP_CHECK(param && grad && exp_avg && exp_avg_sq && step >= 1, "adam_step: bad arguments");
  KP_CHECK((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
           "adam_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, as_stream(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), beta1, beta2, eps, weight_decay, (float)(1.0 / sqrt(bc2)),
      grad_scale);
  KP_LAUNCH_CHECK("adam_step");
  return 0;
}

extern "C" int kp_repack_nchw_to_hwc(const float* src, float* dst, int C, int H, int W, void* stream) {
  KP_CHECK(src && dst, "repack: NULL argument");
  const int64_t total = (int64_t)C * H * W;
  repack_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(src, dst, C, H, W, 1);
  KP_LAUNCH_CHECK("repack_nchw_to_hwc");
  return 0;
}

extern "C" int kp_repack_hwc_to_nchw(const float* src, float* dst, int C, int H, int W, void* stream) {
  KP_CHECK(src && dst, "repack: NULL argument");
  const int64_t total = (int64_t)C * H * W;
  repack_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(src, dst, C, H, W, 0);
  KP_LAUNCH_CHECK("repack_hwc_to_nchw");
  return 0;
}
