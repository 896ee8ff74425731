// Ray setup and proposal resampling kernels (sm_100a).
//   kp_aabb_intersect : NS/model_components/scene_colliders.py:57-95
//   kp_uniform_bins   : NS/model_components/ray_samplers.py:79-126 (SpacedSampler / UniformSampler /
//                       UniformLinDispPiecewiseSampler :221-246)
//   kp_pdf_resample   : NS/model_components/ray_samplers.py:274-369 (PDFSampler, include_original=False)
// Everything that decides a sample *position or index* is computed with explicitly rounded IEEE ops
// (__fadd_rn/__fmul_rn/__fdiv_rn: never contracted into FMA) in the reference's operation order, so bins
// and searchsorted indices reproduce the torch CPU path bit for bit given the same cdf.  Prefix sums are
// accumulated in double and rounded to float per element, which is what torch's CPU cumsum does for fp32.
#include "common.cuh"

namespace kp {

__global__ void aabb_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t N, float ax0, float ay0,
                            float az0, float ax1, float ay1, float az1, float near_plane, float* __restrict__ nears,
                            float* __restrict__ fars) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float amin[3] = {ax0, ay0, az0}, amax[3] = {ax1, ay1, az1};
  float tn = -INFINITY, tf = INFINITY;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float frac = __fdiv_rn(1.0f, __fadd_rn(d[i * 3 + k], 1e-6f));
    const float ta = __fmul_rn(__fsub_rn(amin[k], o[i * 3 + k]), frac);
    const float tb = __fmul_rn(__fsub_rn(amax[k], o[i * 3 + k]), frac);
    tn = fmaxf(tn, fminf(ta, tb));
    tf = fminf(tf, fmaxf(ta, tb));
  }
  tn = fmaxf(tn, near_plane);
  tf = fmaxf(tf, __fadd_rn(tn, 1e-6f));
  nears[i] = tn;
  fars[i] = tf;
}

// torch.minimum / torch.maximum semantics: a NaN operand gives NaN (fminf / fmaxf would drop it)
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? NAN : fminf(a, b); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }

// _intersect_aabb (NS/utils/math.py:201-238, max_bound = invalid_value = 1e10): the slab test Cameras.generate_rays
// applies for aabb_box (cameras.py:478-497; nerfacc's ray_aabb_intersect is its first choice, :260-270, and is not in
// the image).  Plain division by the direction (no epsilon), entry / exit clamped to [0, 1e10], a miss gives 1e10 twice.
__global__ void __launch_bounds__(256) ray_box_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t N,
                                                      float a0, float a1, float a2, float b0, float b1, float b2,
                                                      float* __restrict__ t_min, float* __restrict__ t_max) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float amin[3] = {a0, a1, a2}, amax[3] = {b0, b1, b2};
  float lo[3], hi[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float ta = __fdiv_rn(__fsub_rn(amin[k], o[i * 3 + k]), d[i * 3 + k]);
    const float tb = __fdiv_rn(__fsub_rn(amax[k], o[i * 3 + k]), d[i * 3 + k]);
    lo[k] = min_nan(ta, tb);
    hi[k] = max_nan(ta, tb);
  }
  float tn = max_nan(max_nan(lo[0], lo[1]), lo[2]);
  float tf = min_nan(min_nan(hi[0], hi[1]), hi[2]);
  tn = min_nan(max_nan(tn, 0.0f), 1e10f);  // torch.clamp(min=0, max=max_bound), NaN kept
  tf = min_nan(max_nan(tf, 0.0f), 1e10f);
  const bool miss = tf <= tn;
  t_min[i] = miss ? 1e10f : tn;
  t_max[i] = miss ? 1e10f : tf;
}

// spacing functions (ray_samplers.py:129-150 uniform; :236-246 piecewise uniform / linear-in-disparity)
__device__ __forceinline__ float spacing_fn(float x, int mode) {
  if (mode == 0) return x;
  return x < 1.f ? __fdiv_rn(x, 2.f) : __fsub_rn(1.f, __fdiv_rn(1.f, __fmul_rn(2.f, x)));
}
__device__ __forceinline__ float spacing_fn_inv(float x, int mode) {
  if (mode == 0) return x;
  return x < 0.5f ? __fmul_rn(2.f, x) : __fdiv_rn(1.f, __fsub_rn(2.f, __fmul_rn(2.f, x)));
}
// spacing_to_euclidean_fn = fn_inv(x * s_far + (1 - x) * s_near)   (ray_samplers.py:115)
__device__ __forceinline__ float to_euclid(float x, float s_near, float s_far, int mode) {
  const float v = __fadd_rn(__fmul_rn(x, s_far), __fmul_rn(__fsub_rn(1.f, x), s_near));
  return spacing_fn_inv(v, mode);
}

__device__ __forceinline__ float uniform_bin(const float* __restrict__ lin, const float* __restrict__ t_rand, int rand_stride,
                                             int64_t n, int j, int S) {
  float b = lin[j];
  if (t_rand != nullptr) {
    const float upper = j < S ? __fdiv_rn(__fadd_rn(lin[j + 1], lin[j]), 2.f) : lin[S];
    const float lower = j > 0 ? __fdiv_rn(__fadd_rn(lin[j], lin[j - 1]), 2.f) : lin[0];
    const float r = rand_stride ? t_rand[n * rand_stride + j] : t_rand[n];
    b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), r));
  }
  return b;
}

__global__ void uniform_bins_kernel(const float* __restrict__ lin, const float* __restrict__ t_rand, int rand_stride,
                                    const float* __restrict__ nears, const float* __restrict__ fars, int64_t N, int S,
                                    int mode, float* __restrict__ sbins, float* __restrict__ ebins,
                                    float* __restrict__ starts, float* __restrict__ ends, float* __restrict__ deltas) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = S + 1;
  if (idx >= N * nb) return;
  const int64_t n = idx / nb;
  const int j = (int)(idx % nb);
  const float s_near = spacing_fn(nears[n], mode), s_far = spacing_fn(fars[n], mode);
  const float b = uniform_bin(lin, t_rand, rand_stride, n, j, S);
  const float e = to_euclid(b, s_near, s_far, mode);
  sbins[idx] = b;
  ebins[idx] = e;
  if (starts != nullptr && j < S) {  // frustum starts / ends / deltas (rays.py:254-262) written contiguously
    const float e1 = to_euclid(uniform_bin(lin, t_rand, rand_stride, n, j + 1, S), s_near, s_far, mode);
    starts[n * S + j] = e;
    ends[n * S + j] = e1;
    deltas[n * S + j] = __fsub_rn(e1, e);
  }
}

// sum_i (x[i] + pad) over one fp32 row, reproducing BIT FOR BIT what torch.sum(dim=-1) returns on CPU for a
// contiguous fp32 row (ATen SumKernel.cpp: vectorized_inner_sum -> row_sum -> multi_row_sum with 8-lane vectors,
// ILP factor 4 and a 4-level cascade; verified against torch 2.11 for S = 40..4096, DESIGN.md "bit-exactness").
// The cdf that decides the searchsorted indices is normalised by this sum, so its rounding must match.
// Lane t of the warp plays vector lane (t & 7) of ILP accumulator (t >> 3).  All lanes return the result.
__device__ __forceinline__ float torch_cpu_row_sum(const float* __restrict__ x, int S, float pad, int lane) {
  constexpr int V = 8, ILP = 4, LEVELS = 4;
  const int l = lane & 7, k = lane >> 3;
  const int vec_size = S / V, size_ilp = vec_size / ILP;
  int clog = 0;
  while ((1 << clog) < size_ilp) ++clog;  // CeilLog2
  const int level_power = max(4, clog / LEVELS);
  const int level_step = 1 << level_power, level_mask = level_step - 1;
  float acc[LEVELS] = {0.f, 0.f, 0.f, 0.f};
  int i = 0;
  for (; i + level_step <= size_ilp;) {
    for (int j = 0; j < level_step; ++j, ++i) acc[0] = __fadd_rn(acc[0], __fadd_rn(x[(i * ILP + k) * V + l], pad));
#pragma unroll
    for (int j = 1; j < LEVELS; ++j) {
      acc[j] = __fadd_rn(acc[j], acc[j - 1]);
      acc[j - 1] = 0.f;
      if ((i & (level_mask << (j * level_power))) != 0) break;
    }
  }
  for (; i < size_ilp; ++i) acc[0] = __fadd_rn(acc[0], __fadd_rn(x[(i * ILP + k) * V + l], pad));
#pragma unroll
  for (int j = 1; j < LEVELS; ++j) acc[0] = __fadd_rn(acc[0], acc[j]);
  float p = acc[0];
  // leftover whole vectors go to ILP accumulator 0, then accumulators 1..3 are folded into 0 in order
  if (k == 0)
    for (int v = size_ilp * ILP; v < vec_size; ++v) p = __fadd_rn(p, __fadd_rn(x[v * V + l], pad));
  const float p1 = __shfl_sync(0xffffffffu, p, l + 8), p2 = __shfl_sync(0xffffffffu, p, l + 16),
              p3 = __shfl_sync(0xffffffffu, p, l + 24);
  const float q = __fadd_rn(__fadd_rn(__fadd_rn(p, p1), p2), p3);  // valid on lanes 0..7
  float fin = 0.f;
  for (int t = vec_size * V; t < S; ++t) fin = __fadd_rn(fin, __fadd_rn(x[t], pad));  // scalar tail first
#pragma unroll
  for (int v = 0; v < V; ++v) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, q, v));
  return fin;
}

// One warp per ray.  Dynamic smem per warp: cdf[S_in+1] + existing bins[S_in+1].
__global__ void __launch_bounds__(128) pdf_resample_kernel(
    const float* __restrict__ weights, const float* __restrict__ existing, int S_in, const float* __restrict__ u_base,
    const float* __restrict__ rand, int rand_stride, float eval_offset, const float* __restrict__ nears,
    const float* __restrict__ fars, int64_t N, int S_out, float hist_pad, float eps, int mode,
    float* __restrict__ cdf_out, float* __restrict__ sbins, float* __restrict__ ebins, int64_t* __restrict__ inds_out,
    const float* __restrict__ anneal_dev, float anneal_host, float* __restrict__ starts, float* __restrict__ ends,
    float* __restrict__ deltas) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= N) return;
  const int nb_in = S_in + 1;
  float* s_cdf = smem + (size_t)warp * 3 * nb_in;
  float* s_bins = s_cdf + nb_in;
  float* s_w = s_bins + nb_in;
  // proposal-weight annealing, torch.pow(weights, anneal) (ray_samplers.py:584), folded in; anneal == 1 is a no-op
  const float anneal = anneal_dev != nullptr ? *anneal_dev : anneal_host;
  for (int i = lane; i < S_in; i += 32) {
    const float wv = weights[n * S_in + i];
    s_w[i] = (anneal == 1.0f) ? wv : powf(wv, anneal);
  }
  __syncwarp();
  const float* w_row = s_w;
  const int epl = (S_in + 31) / 32;  // contiguous elements per lane
  const int i0 = lane * epl, i1 = min(S_in, i0 + epl);
  // weights + histogram_padding (ray_samplers.py:302), row sum in torch's CPU summation order (see below)
  float w_sum = torch_cpu_row_sum(w_row, S_in, hist_pad, lane);
  double part = 0.0;
  const float padding = fmaxf(__fsub_rn(eps, w_sum), 0.f);  // relu(eps - sum)
  const float pad_each = __fdiv_rn(padding, (float)S_in);
  w_sum = __fadd_rn(w_sum, padding);
  // pdf = w / sum; cdf = min(1, cumsum(pdf)); prepend 0   (:309-312)
  part = 0.0;
  for (int i = i0; i < i1; ++i)
    part += (double)__fdiv_rn(__fadd_rn(__fadd_rn(w_row[i], hist_pad), pad_each), w_sum);
  double run = warp_excl_scan_d(part, lane);  // exclusive prefix of lane totals
  for (int i = i0; i < i1; ++i) {
    run += (double)__fdiv_rn(__fadd_rn(__fadd_rn(w_row[i], hist_pad), pad_each), w_sum);
    s_cdf[i + 1] = fminf(1.f, (float)run);
  }
  if (lane == 0) s_cdf[0] = 0.f;
  for (int i = lane; i < nb_in; i += 32) s_bins[i] = existing[n * nb_in + i];
  __syncwarp();
  if (cdf_out != nullptr)
    for (int i = lane; i < nb_in; i += 32) cdf_out[n * nb_in + i] = s_cdf[i];
  const int nb_out = S_out + 1;
  const float s_near = spacing_fn(nears[n], mode), s_far = spacing_fn(fars[n], mode);
  for (int j = lane; j < nb_out; j += 32) {
    float u;
    if (rand != nullptr) {
      const float r = rand_stride ? rand[n * rand_stride + j] : rand[n];
      u = __fadd_rn(u_base[j], __fdiv_rn(r, (float)nb_out));  // :318-322
    } else {
      u = __fadd_rn(u_base[j], eval_offset);  // :325-326
    }
    // searchsorted(cdf, u, side="right"): number of entries <= u   (:341)
    int lo = 0, hi = nb_in;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = min(max(lo - 1, 0), S_in), above = min(max(lo, 0), S_in);
    const float c0 = s_cdf[below], c1 = s_cdf[above], b0 = s_bins[below], b1 = s_bins[above];
    float t = nan_to_num(__fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0)));
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float b = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));  // :350-351
    sbins[n * nb_out + j] = b;
    ebins[n * nb_out + j] = to_euclid(b, s_near, s_far, mode);
    if (inds_out != nullptr) inds_out[n * nb_out + j] = (int64_t)lo;
  }
  if (starts != nullptr) {
    __syncwarp();  // the warp's euclidean bins are visible to all its lanes
    for (int j = lane; j < S_out; j += 32) {
      const float e0 = ebins[n * nb_out + j], e1 = ebins[n * nb_out + j + 1];
      starts[n * S_out + j] = e0;
      ends[n * S_out + j] = e1;
      deltas[n * S_out + j] = __fsub_rn(e1, e0);
    }
  }
}

}  // namespace kp

using namespace kp;

extern "C" int kp_aabb_intersect(const float* origins, const float* directions, int64_t N, const float* aabb,
                                 float near_plane, float* nears, float* fars, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(origins && directions && aabb && nears && fars, "aabb_intersect: NULL argument");
  aabb_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(origins, directions, N, aabb[0], aabb[1], aabb[2],
                                                                        aabb[3], aabb[4], aabb[5], near_plane, nears, fars);
  KP_LAUNCH_CHECK("aabb_intersect");
  return 0;
}

extern "C" int kp_intersect_aabb(const float* origins, const float* directions, int64_t N, const float* aabb,
                                 float* t_min, float* t_max, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(origins && directions && aabb && t_min && t_max, "intersect_aabb: NULL argument");
  ray_box_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(origins, directions, N, aabb[0], aabb[1], aabb[2],
                                                                           aabb[3], aabb[4], aabb[5], t_min, t_max);
  KP_LAUNCH_CHECK("intersect_aabb");
  return 0;
}

extern "C" int kp_uniform_bins(const float* lin_bins, const float* t_rand, int rand_stride, const float* nears,
                               const float* fars, int64_t N, int S, int spacing, float* spacing_bins, float* euclid_bins,
                               float* starts, float* ends, float* deltas, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(lin_bins && nears && fars && spacing_bins && euclid_bins, "uniform_bins: NULL argument");
  KP_CHECK(S >= 1, "uniform_bins: S=%d", S);
  KP_CHECK(spacing == 0 || spacing == 1, "uniform_bins: spacing=%d unsupported", spacing);
  KP_CHECK(t_rand == nullptr || rand_stride == 0 || rand_stride == S + 1, "uniform_bins: t_rand must be [N,S+1] or [N,1]");
  KP_CHECK((starts == nullptr) == (ends == nullptr) && (starts == nullptr) == (deltas == nullptr),
           "uniform_bins: starts/ends/deltas must be given together");
  const int64_t total = N * (S + 1);
  uniform_bins_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      lin_bins, t_rand, rand_stride, nears, fars, N, S, spacing, spacing_bins, euclid_bins, starts, ends, deltas);
  KP_LAUNCH_CHECK("uniform_bins");
  return 0;
}

extern "C" int kp_pdf_resample(const float* weights, const float* existing_bins, int S_in, const float* u_base,
                               const float* rand, int rand_stride, const float* nears, const float* fars, int64_t N,
                               int S_out, float histogram_padding, float eps, int spacing, float* cdf_out,
                               float* spacing_bins, float* euclid_bins, int64_t* inds, const float* anneal_dev,
                               float anneal_host, float* starts, float* ends, float* deltas, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(weights && existing_bins && u_base && nears && fars && spacing_bins && euclid_bins, "pdf_resample: NULL argument");
  KP_CHECK(S_in >= 1 && S_out >= 1, "pdf_resample: S_in=%d S_out=%d", S_in, S_out);
  KP_CHECK(spacing == 0 || spacing == 1, "pdf_resample: spacing=%d unsupported", spacing);
  KP_CHECK(rand == nullptr || rand_stride == 0 || rand_stride == S_out + 1, "pdf_resample: rand must be [N,S_out+1] or [N,1]");
  const size_t smem = (size_t)4 * 3 * (S_in + 1) * sizeof(float);
  KP_CHECK(smem <= 200 * 1024, "pdf_resample: S_in=%d too large", S_in);
  if (smem > 48 * 1024) cudaFuncSetAttribute(pdf_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const float eval_offset = (float)(1.0 / (2.0 * (S_out + 1)));
  pdf_resample_kernel<<<(unsigned)ceil_div(N, 4), 128, smem, as_stream(stream)>>>(
      weights, existing_bins, S_in, u_base, rand, rand_stride, eval_offset, nears, fars, N, S_out, histogram_padding, eps,
      spacing, cdf_out, spacing_bins, euclid_bins, inds, anneal_dev, anneal_host, starts, ends, deltas);
  KP_LAUNCH_CHECK("pdf_resample");
  return 0;
}
