// Loss head of the K-Planes training step (sm_100a): everything KPlanesModel.get_loss_dict / get_metrics_dict do
// AFTER the per-ray / per-sample loss kernels, in one launch per direction instead of ~30 tiny torch kernels:
//   rgb_loss        = coef_rgb  * mean((image - rgb)^2)                      NS/models/kplanes.py:416, MSELoss
//   distortion_loss = coef_dist * mean(per-ray distortion)                   NS/model_components/losses.py:139-144
//   interlevel_loss = coef_il   * sum_l mean(per-sample outer loss, level l) losses.py:106-121
//   total           = the three + sum(extra)    (extra: already-scaled regulariser terms, kplanes.py:430-452;
//                                                 Trainer: sum(loss_dict.values()), NS/engine/trainer.py:398-400)
//   psnr            = -10 log10(mean((image - rgb)^2))                       kplanes.py:392-398 (data range 1)
// Sums are accumulated in double (block partials + one atomic per block); the last block to finish writes the
// results and resets the workspace, so the launch is self-contained and CUDA-graph replayable.
#include "common.cuh"

namespace kp {

constexpr int kHeadThreads = 256;
constexpr int kHeadMaxLevels = KP_LOSS_HEAD_MAX_LEVELS;

struct HeadArgs {
  const float* pred;   // [N,3]
  const float* image;  // [N,3]
  const float* dist;   // [N] or null
  const float* il[kHeadMaxLevels];
  int64_t il_count[kHeadMaxLevels];  // N * S_l
  int n_levels;
  int64_t N;
  float coef_rgb, coef_dist, coef_il;
  const float* extra;  // [n_extra] or null
  int n_extra;
};

__device__ __forceinline__ double block_sum_d(double v, double* smem) {
  v = warp_sum_d(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double t = (threadIdx.x < kHeadThreads / 32) ? smem[threadIdx.x] : 0.0;
  if (wid == 0) t = warp_sum_d(t);
  return t;  // valid in thread 0
}

// workspace: double acc[3], then unsigned counter (as the 4th double's low word)
__global__ void __launch_bounds__(kHeadThreads) loss_head_fwd_kernel(const __grid_constant__ HeadArgs a, double* ws,
                                                                     float* vals3, float* total, float* psnr) {
  __shared__ double smem[kHeadThreads / 32];
  __shared__ bool s_last;
  const int64_t tid = (int64_t)blockIdx.x * kHeadThreads + threadIdx.x, stride = (int64_t)gridDim.x * kHeadThreads;
  double sq = 0.0, ds = 0.0, il = 0.0;
  for (int64_t i = tid; i < 3 * a.N; i += stride) {
    const float d = a.image[i] - a.pred[i];
    sq += (double)(d * d);
  }
  if (a.dist != nullptr)
    for (int64_t i = tid; i < a.N; i += stride) ds += (double)a.dist[i];
  for (int l = 0; l < a.n_levels; ++l) {
    double s = 0.0;
    for (int64_t i = tid; i < a.il_count[l]; i += stride) s += (double)a.il[l][i];
    il += s / (double)a.il_count[l];
  }
  sq = block_sum_d(sq, smem);
  ds = block_sum_d(ds, smem);
  il = block_sum_d(il, smem);
  unsigned* counter = reinterpret_cast<unsigned*>(ws + 3);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 0, sq);
    atomicAdd(ws + 1, ds);
    atomicAdd(ws + 2, il);
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    volatile double* v = ws;
    const double mse = v[0] / (double)(3 * a.N);
    const float l_rgb = a.coef_rgb * (float)mse;
    const float l_dist = a.dist != nullptr ? a.coef_dist * (float)(v[1] / (double)a.N) : 0.f;
    const float l_il = a.coef_il * (float)v[2];
    vals3[0] = l_rgb;
    vals3[1] = l_dist;
    vals3[2] = l_il;
    float t = l_rgb + l_dist + l_il;
    for (int r = 0; r < a.n_extra; ++r) t += a.extra[r];
    *total = t;
    *psnr = -10.0f * log10f((float)mse);
    v[0] = 0.0; v[1] = 0.0; v[2] = 0.0;
    *counter = 0u;
  }
}

struct HeadGrad {
  float* g_pred;  // [N,3] or null
  float* g_dist;  // [N] or null
  float* g_il[kHeadMaxLevels];
};

// upstream: g_total (scalar, may be null = 0) and g_vals3 (may be null = 0)
__global__ void __launch_bounds__(kHeadThreads) loss_head_bwd_kernel(const __grid_constant__ HeadArgs a,
                                                                     const __grid_constant__ HeadGrad g,
                                                                     const float* g_total, const float* g_vals3) {
  const float gt = g_total != nullptr ? *g_total : 0.f;
  const float u_rgb = gt + (g_vals3 != nullptr ? g_vals3[0] : 0.f);
  const float u_dist = gt + (g_vals3 != nullptr ? g_vals3[1] : 0.f);
  const float u_il = gt + (g_vals3 != nullptr ? g_vals3[2] : 0.f);
  const int64_t tid = (int64_t)blockIdx.x * kHeadThreads + threadIdx.x, stride = (int64_t)gridDim.x * kHeadThreads;
  if (g.g_pred != nullptr) {
    const float k = u_rgb * a.coef_rgb * 2.0f / (float)(3 * a.N);
    for (int64_t i = tid; i < 3 * a.N; i += stride) g.g_pred[i] = k * (a.pred[i] - a.image[i]);
  }
  if (g.g_dist != nullptr) {
    const float k = u_dist * a.coef_dist / (float)a.N;
    for (int64_t i = tid; i < a.N; i += stride) g.g_dist[i] = k;
  }
  for (int l = 0; l < a.n_levels; ++l) {
    if (g.g_il[l] == nullptr) continue;
    const float k = u_il * a.coef_il / (float)a.il_count[l];
    for (int64_t i = tid; i < a.il_count[l]; i += stride) g.g_il[l][i] = k;
  }
}

static int fill_head(HeadArgs& a, const float* pred, const float* image, int64_t N, const float* dist,
                     const float* const* il_ptrs, const int64_t* il_counts, int n_levels, float coef_rgb, float coef_dist,
                     float coef_il, const float* extra, int n_extra) {
  KP_CHECK(pred != nullptr && image != nullptr && N > 0, "loss_head: pred/image NULL or N=0");
  KP_CHECK(n_levels >= 0 && n_levels <= kHeadMaxLevels, "loss_head: n_levels=%d out of range [0,%d]", n_levels, kHeadMaxLevels);
  KP_CHECK(n_extra == 0 || extra != nullptr, "loss_head: extra is NULL");
  a.pred = pred; a.image = image; a.dist = dist; a.N = N; a.n_levels = n_levels;
  for (int l = 0; l < kHeadMaxLevels; ++l) { a.il[l] = nullptr; a.il_count[l] = 0; }
  for (int l = 0; l < n_levels; ++l) {
    KP_CHECK(il_ptrs[l] != nullptr && il_counts[l] > 0, "loss_head: interlevel level %d invalid", l);
    a.il[l] = il_ptrs[l];
    a.il_count[l] = il_counts[l];
  }
  a.coef_rgb = coef_rgb; a.coef_dist = coef_dist; a.coef_il = coef_il;
  a.extra = extra; a.n_extra = n_extra;
  return 0;
}

static unsigned head_grid(const HeadArgs& a) {
  int64_t work = 3 * a.N;
  for (int l = 0; l < a.n_levels; ++l) work = work > a.il_count[l] ? work : a.il_count[l];
  const int64_t blocks = ceil_div(work, (int64_t)kHeadThreads * 4);
  return (unsigned)(blocks < 1 ? 1 : (blocks > 148 ? 148 : blocks));
}

}  // namespace kp

using namespace kp;

extern "C" int kp_loss_head_fwd(const float* pred, const float* image, int64_t N, const float* dist_per_ray,
                                const float* const* il_ptrs, const int64_t* il_counts, int n_levels, float coef_rgb,
                                float coef_dist, float coef_il, const float* extra, int n_extra, double* workspace4,
                                float* vals3, float* total, float* psnr, void* stream) {
  HeadArgs a;
  if (fill_head(a, pred, image, N, dist_per_ray, il_ptrs, il_counts, n_levels, coef_rgb, coef_dist, coef_il, extra, n_extra))
    return 1;
  KP_CHECK(workspace4 && vals3 && total && psnr, "loss_head_fwd: NULL output");
  loss_head_fwd_kernel<<<head_grid(a), kHeadThreads, 0, as_stream(stream)>>>(a, workspace4, vals3, total, psnr);
  KP_LAUNCH_CHECK("loss_head_fwd");
  return 0;
}

extern "C" int kp_loss_head_bwd(const float* pred, const float* image, int64_t N, const int64_t* il_counts, int n_levels,
                                float coef_rgb, float coef_dist, float coef_il, const float* grad_total,
                                const float* grad_vals3, float* grad_pred, float* grad_dist, float* const* grad_il_ptrs,
                                void* stream) {
  HeadArgs a;
  KP_CHECK(n_levels >= 0 && n_levels <= kHeadMaxLevels, "loss_head_bwd: n_levels=%d", n_levels);
  const float* il_fake[kHeadMaxLevels];
  for (int l = 0; l < n_levels; ++l) il_fake[l] = pred;  // forward inputs are not read by the backward
  if (fill_head(a, pred, image, N, nullptr, il_fake, il_counts, n_levels, coef_rgb, coef_dist, coef_il, nullptr, 0)) return 1;
  HeadGrad g;
  g.g_pred = grad_pred;
  g.g_dist = grad_dist;
  for (int l = 0; l < kHeadMaxLevels; ++l) g.g_il[l] = (l < n_levels && grad_il_ptrs != nullptr) ? grad_il_ptrs[l] : nullptr;
  loss_head_bwd_kernel<<<head_grid(a), kHeadThreads, 0, as_stream(stream)>>>(a, g, grad_total, grad_vals3);
  KP_LAUNCH_CHECK("loss_head_bwd");
  return 0;
}
