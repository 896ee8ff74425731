// Alpha compositing kernels (sm_100a): one warp per ray, shuffle scans.
//   kp_weights_fwd/bwd : RaySamples.get_weights, NS/cameras/rays.py:127-149.  The reference computes
//       transmittance as exp(-cumsum(delta*sigma)), NOT as a product of (1-alpha); we reproduce that with a
//       prefix SUM (double accumulate, rounded to fp32 per element like torch's CPU cumsum) and one exp.
//   kp_render_fwd/bwd  : RGBRenderer / AccumulationRenderer / DepthRenderer / MedianRGBRenderer,
//       NS/model_components/renderers.py:58-140, 197-223, 226-287, 290-362.
#include "common.cuh"

namespace kp {

constexpr int kRaysPerBlock = 4;

__global__ void __launch_bounds__(32 * kRaysPerBlock) weights_fwd_kernel(const float* __restrict__ deltas,
                                                                         const float* __restrict__ sigma, int64_t N,
                                                                         int S, float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
  if (n >= N) return;
  const int epl = (S + 31) / 32;
  const int i0 = lane * epl, i1 = min(S, i0 + epl);
  const float* d = deltas + n * S;
  const float* s = sigma + n * S;
  double part = 0.0;
  for (int i = i0; i < i1; ++i) part += (double)__fmul_rn(d[i], s[i]);
  double run = warp_excl_scan_d(part, lane);
  for (int i = i0; i < i1; ++i) {
    const float x = __fmul_rn(d[i], s[i]);
    const float alpha = 1.f - expf(-x);
    const float trans = expf(-(float)run);
    weights[n * S + i] = nan_to_num(alpha * trans);
    run += (double)x;
  }
}

// dL/dsigma_i = delta_i * ( g_i * T_i * exp(-x_i) - sum_{k>i} g_k * w_k ),  g masked where w was non-finite.
__global__ void __launch_bounds__(32 * kRaysPerBlock) weights_bwd_kernel(const float* __restrict__ deltas,
                                                                         const float* __restrict__ sigma,
                                                                         const float* __restrict__ gw, int64_t N, int S,
                                                                         float* __restrict__ gsigma) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
  if (n >= N) return;
  const int epl = (S + 31) / 32;
  const int i0 = lane * epl, i1 = min(S, i0 + epl);
  const float* d = deltas + n * S;
  const float* s = sigma + n * S;
  const float* g = gw + n * S;
  // pass 1: prefix of x (for T_i) and per-lane totals of g*w
  double part = 0.0;
  for (int i = i0; i < i1; ++i) part += (double)__fmul_rn(d[i], s[i]);
  const double run0 = warp_excl_scan_d(part, lane);
  double run = run0, gw_part = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float x = __fmul_rn(d[i], s[i]);
    const float raw = (1.f - expf(-x)) * expf(-(float)run);
    if (isfinite(raw)) gw_part += (double)(g[i] * raw);
    run += (double)x;
  }
  // suffix sums: total - inclusive prefix
  const double incl = warp_incl_scan_d(gw_part, lane);
  const double total = __shfl_sync(0xffffffffu, incl, 31);
  double suffix_after_lane = total - incl;  // sum over lanes > this lane
  // pass 2 (reverse within the lane's chunk)
  run = run0;
  // recompute per-element values forward, store contributions, then walk backwards
  double tail = suffix_after_lane;
  // need T_i for each i: recompute running prefix forward into a small local walk
  // (epl <= 8 for S<=256; for larger S this loop is still correct, just longer)
  for (int i = i1 - 1; i >= i0; --i) {
    // prefix for element i = run0 + sum_{j in [i0,i)} x_j
    double pre = run0;
    for (int j = i0; j < i; ++j) pre += (double)__fmul_rn(d[j], s[j]);
    const float x = __fmul_rn(d[i], s[i]);
    const float ex = expf(-x), trans = expf(-(float)pre);
    const float raw = (1.f - ex) * trans;
    const float gi = isfinite(raw) ? g[i] : 0.f;
    const float dx = gi * trans * ex - (float)tail;
    gsigma[n * S + i] = d[i] * dx;
    tail += (double)(gi * raw);
  }
}

__global__ void __launch_bounds__(32 * kRaysPerBlock) render_fwd_kernel(
    const float* __restrict__ weights, const float* __restrict__ rgb, const float* __restrict__ steps,
    const float* __restrict__ bg, int bg_mode, int nan_rgb, int64_t N, int S, float* __restrict__ comp,
    float* __restrict__ acc_out, int64_t* __restrict__ median, float* __restrict__ exp_depth,
    const float* __restrict__ starts, const float* __restrict__ ends, float* __restrict__ median_depth) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
  if (n >= N) return;
  const int epl = (S + 31) / 32;
  const int i0 = lane * epl, i1 = min(S, i0 + epl);
  const float* w = weights + n * S;
  float r = 0.f, g = 0.f, b = 0.f, a = 0.f, dsum = 0.f;
  double part = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float wi = w[i];
    a += wi;
    part += (double)wi;
    if (rgb != nullptr) {
      float cr = rgb[(n * S + i) * 3 + 0], cg = rgb[(n * S + i) * 3 + 1], cb = rgb[(n * S + i) * 3 + 2];
      if (nan_rgb) { cr = nan_to_num(cr); cg = nan_to_num(cg); cb = nan_to_num(cb); }
      r = fmaf(wi, cr, r); g = fmaf(wi, cg, g); b = fmaf(wi, cb, b);
    }
    if (steps != nullptr) dsum = fmaf(wi, steps[n * S + i], dsum);
  }
  // median index: first i with cumsum(w)[i] >= 0.5  == count(cumsum < 0.5), clamped (renderers.py:260-263)
  if (median != nullptr || median_depth != nullptr) {
    double run = warp_excl_scan_d(part, lane);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) {
      run += (double)w[i];
      cnt += ((float)run < 0.5f) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
      const int mi = min(cnt, S - 1);
      if (median != nullptr) median[n] = (int64_t)mi;
      // DepthRenderer "median": steps = (starts + ends) / 2 gathered at the median index (renderers.py:256-264)
      if (median_depth != nullptr) median_depth[n] = __fdiv_rn(__fadd_rn(starts[n * S + mi], ends[n * S + mi]), 2.f);
    }
  }
  a = warp_sum(a);
  if (rgb != nullptr) { r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); }
  if (steps != nullptr) dsum = warp_sum(dsum);
  if (lane == 0) {
    if (acc_out != nullptr) acc_out[n] = a;
    if (exp_depth != nullptr) exp_depth[n] = dsum / (a + 1e-10f);
    if (comp != nullptr && rgb != nullptr) {
      float br, bgc, bb;
      if (bg_mode == 1) {
        br = rgb[(n * S + S - 1) * 3 + 0]; bgc = rgb[(n * S + S - 1) * 3 + 1]; bb = rgb[(n * S + S - 1) * 3 + 2];
        if (nan_rgb) { br = nan_to_num(br); bgc = nan_to_num(bgc); bb = nan_to_num(bb); }
      } else {
        br = bg[n * 3 + 0]; bgc = bg[n * 3 + 1]; bb = bg[n * 3 + 2];
      }
      const float rem = 1.0f - a;
      comp[n * 3 + 0] = r + br * rem;
      comp[n * 3 + 1] = g + bgc * rem;
      comp[n * 3 + 2] = b + bb * rem;
    }
  }
}

__global__ void render_bwd_kernel(const float* __restrict__ weights, const float* __restrict__ rgb,
                                  const float* __restrict__ bg, int bg_mode, int64_t N, int S,
                                  const float* __restrict__ gcomp, const float* __restrict__ gacc,
                                  float* __restrict__ gw, float* __restrict__ grgb) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * S) return;
  const int64_t n = idx / S;
  const int i = (int)(idx % S);
  float gwi = gacc != nullptr ? gacc[n] : 0.f;
  if (gcomp != nullptr) {
    const float g0 = gcomp[n * 3 + 0], g1 = gcomp[n * 3 + 1], g2 = gcomp[n * 3 + 2];
    float b0, b1, b2;
    if (bg_mode == 1) {
      b0 = rgb[(n * S + S - 1) * 3 + 0]; b1 = rgb[(n * S + S - 1) * 3 + 1]; b2 = rgb[(n * S + S - 1) * 3 + 2];
    } else {
      b0 = bg[n * 3 + 0]; b1 = bg[n * 3 + 1]; b2 = bg[n * 3 + 2];
    }
    const float c0 = rgb[idx * 3 + 0], c1 = rgb[idx * 3 + 1], c2 = rgb[idx * 3 + 2];
    gwi += g0 * (c0 - b0) + g1 * (c1 - b1) + g2 * (c2 - b2);
    if (grgb != nullptr) {
      const float wi = weights[idx];
      grgb[idx * 3 + 0] = wi * g0; grgb[idx * 3 + 1] = wi * g1; grgb[idx * 3 + 2] = wi * g2;
    }
  } else if (grgb != nullptr) {
    grgb[idx * 3 + 0] = 0.f; grgb[idx * 3 + 1] = 0.f; grgb[idx * 3 + 2] = 0.f;
  }
  gw[idx] = gwi;
}

// last_sample background: d comp / d rgb[S-1] gets an extra (1 - acc) term (one thread per ray)
__global__ void render_bwd_last_sample_kernel(const float* __restrict__ weights, int64_t N, int S,
                                              const float* __restrict__ gcomp, float* __restrict__ grgb) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a = 0.f;
  for (int i = 0; i < S; ++i) a += weights[n * S + i];
  const float rem = 1.f - a;
  for (int c = 0; c < 3; ++c) grgb[(n * S + S - 1) * 3 + c] += gcomp[n * 3 + c] * rem;
}

}  // namespace kp

using namespace kp;

extern "C" int kp_weights_fwd(const float* deltas, const float* densities, int64_t N, int S, float* weights,
                              void* stream) {
  if (N == 0) return 0;
  KP_CHECK(deltas && densities && weights && S >= 1, "weights_fwd: bad arguments");
  weights_fwd_kernel<<<(unsigned)ceil_div(N, kRaysPerBlock), 32 * kRaysPerBlock, 0, as_stream(stream)>>>(deltas, densities,
                                                                                                         N, S, weights);
  KP_LAUNCH_CHECK("weights_fwd");
  return 0;
}

extern "C" int kp_weights_bwd(const float* deltas, const float* densities, const float* grad_weights, int64_t N, int S,
                              float* grad_densities, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(deltas && densities && grad_weights && grad_densities && S >= 1, "weights_bwd: bad arguments");
  weights_bwd_kernel<<<(unsigned)ceil_div(N, kRaysPerBlock), 32 * kRaysPerBlock, 0, as_stream(stream)>>>(
      deltas, densities, grad_weights, N, S, grad_densities);
  KP_LAUNCH_CHECK("weights_bwd");
  return 0;
}

extern "C" int kp_render_fwd(const float* weights, const float* rgb, const float* steps, const float* bg, int bg_mode,
                             int nan_to_num_rgb, int64_t N, int S, float* comp_rgb, float* accumulation,
                             int64_t* median_index, float* expected_depth, const float* starts, const float* ends,
                             float* median_depth, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(weights && S >= 1, "render_fwd: bad arguments");
  KP_CHECK(comp_rgb == nullptr || rgb != nullptr, "render_fwd: comp_rgb needs rgb");
  KP_CHECK(comp_rgb == nullptr || bg_mode == 1 || bg != nullptr, "render_fwd: tensor background is NULL");
  KP_CHECK(expected_depth == nullptr || steps != nullptr, "render_fwd: expected_depth needs steps");
  KP_CHECK(median_depth == nullptr || (starts != nullptr && ends != nullptr), "render_fwd: median_depth needs starts/ends");
  render_fwd_kernel<<<(unsigned)ceil_div(N, kRaysPerBlock), 32 * kRaysPerBlock, 0, as_stream(stream)>>>(
      weights, rgb, steps, bg, bg_mode, nan_to_num_rgb, N, S, comp_rgb, accumulation, median_index, expected_depth, starts,
      ends, median_depth);
  KP_LAUNCH_CHECK("render_fwd");
  return 0;
}

extern "C" int kp_render_bwd(const float* weights, const float* rgb, const float* bg, int bg_mode, int64_t N, int S,
                             const float* grad_comp, const float* grad_acc, float* grad_weights, float* grad_rgb,
                             void* stream) {
  if (N == 0) return 0;
  KP_CHECK(weights && grad_weights && S >= 1, "render_bwd: bad arguments");
  KP_CHECK(grad_comp == nullptr || rgb != nullptr, "render_bwd: grad_comp needs rgb");
  KP_CHECK(grad_comp == nullptr || bg_mode == 1 || bg != nullptr, "render_bwd: tensor background is NULL");
  render_bwd_kernel<<<(unsigned)ceil_div(N * S, 256), 256, 0, as_stream(stream)>>>(weights, rgb, bg, bg_mode, N, S, grad_comp,
                                                                                   grad_acc, grad_weights, grad_rgb);
  KP_LAUNCH_CHECK("render_bwd");
  if (bg_mode == 1 && grad_comp != nullptr && grad_rgb != nullptr) {
    render_bwd_last_sample_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(weights, N, S, grad_comp, grad_rgb);
    KP_LAUNCH_CHECK("render_bwd_last_sample");
  }
  return 0;
}
