// Importance pixel sampling on the device (sm_100a): the per-image torch.multinomial calls of
// DynamicBasedPixelSampler.sample_method (NS/data/pixel_samplers.py:340-426) for all images of a step in six launches.
//
// The reference draws, for every image the step visits, k pixels proportionally to that image's fp16 IST / ISG weight map
// -- without replacement when the map has at least k non-zero pixels, with replacement otherwise (:396-398) -- one
// torch.multinomial over H*W categories per image on the host: ~200 ms per 4096-ray step at 540x960, two orders of
// magnitude more than the training step itself.  Here every selected image is one grid row: the exponential race of
// pixel_sampler_math.cuh gives each non-zero pixel a key that any thread can recompute, the k largest keys are found by
// an MSB-first radix select (four histogram passes over the map, which stays in L2 / HBM -- nothing but 4 KB of
// histograms per image is written), a fifth pass collects the winners and a one-block-per-image kernel orders them by
// key (torch.multinomial returns them in that order too) and writes (image, row, col) triplets.  Maps with fewer than k
// non-zero pixels collect all of them and draw k times with replacement from that short list.
//
// The result has the distribution of the reference's sampler, not its random stream (the reference consumes torch's CPU
// Mersenne twister pixel by pixel); data/pixel_samplers.py keeps the host path for bit-exact reproduction.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "pixel_sampler_math.cuh"

namespace kp {

struct SamplerArgs {
  const __half* weights;  // [B, HW]
  int B;
  int64_t HW;
  int width;
  const int32_t* sel;  // [n_sel, 3] = image, k, first output row
  int n_sel, k_max;
  uint32_t seed_lo, seed_hi;
  uint32_t* hist;       // [n_sel, 4, 256]
  uint32_t* counters;   // [n_sel, 2] = keys above the threshold collected, keys equal to it collected
  uint32_t* cand_bits;  // [n_sel, k_max]
  int32_t* cand_pix;    // [n_sel, k_max]
  float* cand_w;        // [n_sel, k_max]
  int64_t* out;         // [sum k, 3]
};

struct Selected {
  int img, k, first;
};

__device__ __forceinline__ Selected load_sel(const SamplerArgs& a, int j) {
  Selected s{a.sel[j * 3 + 0], a.sel[j * 3 + 1], a.sel[j * 3 + 2]};
  if (s.img < 0 || s.img >= a.B || s.k < 0) s.k = 0;
  s.k = min(s.k, a.k_max);
  return s;
}

// Calls f(pixel, weight, key bits) for every non-zero pixel of image `img` this block owns (blocks of a grid row share the
// image in chunks of 8 pixels, one 16-byte load each when the rows are 16-byte aligned).
template <typename F>
__device__ __forceinline__ void for_each_nonzero(const SamplerArgs& a, int img, F&& f) {
  const __half* row = a.weights + (int64_t)img * a.HW;
  const bool vec = (a.HW % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.weights) & 15) == 0);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
  for (int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; c < a.HW; c += stride) {
    __align__(16) __half h[8];
    if (vec) {
      *reinterpret_cast<uint4*>(h) = __ldg(reinterpret_cast<const uint4*>(row + c));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) h[q] = (c + q < a.HW) ? row[c + q] : __float2half_rn(0.f);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float w[4];
      bool any = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        w[q] = __half2float(h[half * 4 + q]);
        any |= w[q] > 0.f;
      }
      if (!any) continue;
      const int64_t g = c / 4 + half;  // c is a multiple of 8
      const Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)img, kSamplerStreamRace, (uint32_t)(g >> 32), a.seed_lo, a.seed_hi);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (w[q] > 0.f) f(g * 4 + q, w[q], race_key_bits(w[q], r.v[q]));
    }
  }
}

// state of the radix select after `passes` histogram passes, computed redundantly by every block from the histograms
struct WalkState {
  uint32_t prefix, nnz;
  int need, take_all;
};

// select_walk (pixel_sampler_math.cuh) by the whole 256-thread block: thread b holds bin b of a pass's histogram, a block
// suffix scan gives the number of keys in higher bins, and the one bin where that count crosses `need` announces itself.
// (A single thread walking the 4 x 256 bins took 15 us per block -- more than the sweep over the map it precedes.)
struct WalkScratch {
  uint32_t warp_sum[8];
  uint32_t bin, above;
};

__device__ __forceinline__ WalkState block_walk(const SamplerArgs& a, int j, int passes, int k, WalkScratch* s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WalkState st;
  st.prefix = 0, st.nnz = 0, st.need = k, st.take_all = 0;
  for (int p = 0; p < passes; ++p) {
    const uint32_t h = a.hist[((int64_t)j * 4 + p) * 256 + tid];
    uint32_t incl = h;  // keys in bins tid .. end of this warp's 32 bins
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    if (lane == 0) s->warp_sum[warp] = incl;
    __syncthreads();
    uint32_t total = 0, higher_warps = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t t = s->warp_sum[w];
      total += t;
      higher_warps += w > warp ? t : 0u;
    }
    const uint32_t above = incl - h + higher_warps;  // keys in bins > tid
    if (p == 0) {
      st.nnz = total;
      st.take_all = total <= (uint32_t)k ? 1 : 0;
    }
    if (st.take_all) break;  // uniform over the block
    const uint32_t need = (uint32_t)st.need;
    // bin 0 is also where the serial walk stops when the count never reaches `need`
    const bool mine = tid == 0 ? above < need : (above < need && above + h >= need);
    if (mine) s->bin = (uint32_t)tid, s->above = above;
    __syncthreads();
    st.prefix = (st.prefix << 8) | s->bin;
    st.need -= (int)s->above;
    __syncthreads();  // warp_sum / bin / above are rewritten by the next pass
  }
  return st;
}

template <int PASS>
__global__ void __launch_bounds__(256) sampler_hist_kernel(const __grid_constant__ SamplerArgs a) {
  __shared__ uint32_t s_hist[256];
  __shared__ WalkScratch s_walk;
  const int j = blockIdx.y;
  const Selected s = load_sel(a, j);
  if (s.k == 0) return;
  s_hist[threadIdx.x] = 0;
  uint32_t prefix = 0;
  if (PASS > 0) {
    const WalkState st = block_walk(a, j, PASS, s.k, &s_walk);
    if (st.take_all) return;  // at most k non-zero pixels: no threshold to find
    prefix = st.prefix;
  }
  __syncthreads();
  constexpr int shift = 24 - 8 * PASS;
  for_each_nonzero(a, s.img, [&](int64_t, float, uint32_t bits) {
    bool mine = true;
    if constexpr (PASS > 0) mine = (bits >> (shift + 8)) == prefix;
    if (mine) atomicAdd(&s_hist[(bits >> shift) & 255u], 1u);
  });
  __syncthreads();
  const uint32_t n = s_hist[threadIdx.x];
  if (n) atomicAdd(&a.hist[((int64_t)j * 4 + PASS) * 256 + threadIdx.x], n);
}

__global__ void __launch_bounds__(256) sampler_collect_kernel(const __grid_constant__ SamplerArgs a) {
  __shared__ WalkScratch s_walk;
  const int j = blockIdx.y;
  const Selected s = load_sel(a, j);
  if (s.k == 0) return;
  const WalkState st = block_walk(a, j, 4, s.k, &s_walk);
  const int64_t base = (int64_t)j * a.k_max;
  const int n_above = s.k - st.need;  // keys strictly above the threshold (select_walk)
  for_each_nonzero(a, s.img, [&](int64_t pix, float w, uint32_t bits) {
    int slot = -1;
    if (st.take_all || bits > st.prefix) {
      slot = (int)atomicAdd(&a.counters[j * 2 + 0], 1u);
      if (slot >= (st.take_all ? s.k : n_above)) slot = -1;  // cannot happen: the histograms counted these keys
    } else if (bits == st.prefix) {
      const int t = (int)atomicAdd(&a.counters[j * 2 + 1], 1u);
      if (t < st.need) slot = n_above + t;
    }
    if (slot >= 0) {
      a.cand_bits[base + slot] = bits;
      a.cand_pix[base + slot] = (int32_t)pix;
      a.cand_w[base + slot] = w;
    }
  });
}

__device__ __forceinline__ void write_triplet(const SamplerArgs& a, int64_t row, int img, int32_t pix) {
  a.out[row * 3 + 0] = img;
  a.out[row * 3 + 1] = pix / a.width;
  a.out[row * 3 + 2] = pix % a.width;
}

constexpr int kSortedInSmem = 2048;

// One block per selected image: order the collected pixels and write the (image, row, col) triplets.
__global__ void __launch_bounds__(256) sampler_finalize_kernel(const __grid_constant__ SamplerArgs a) {
  __shared__ WalkScratch s_walk;
  __shared__ int32_t s_pix[kSortedInSmem];
  __shared__ float s_w[kSortedInSmem];
  const int j = blockIdx.x;
  const Selected s = load_sel(a, j);
  if (s.k == 0) return;
  const WalkState st = block_walk(a, j, 1, s.k, &s_walk);
  const int64_t base = (int64_t)j * a.k_max;
  const int n_c = st.take_all ? (int)st.nnz : s.k;
  if (st.nnz == 0) {  // an all-zero map has nothing to draw from (the reference's loop skips it, :391-393): rows of -1
    for (int t = threadIdx.x; t < s.k; t += blockDim.x)
      for (int c = 0; c < 3; ++c) a.out[(int64_t)(s.first + t) * 3 + c] = -1;
    return;
  }
  if ((int)st.nnz >= s.k) {
    // without replacement: the k largest keys, largest first (ties of equal keys by pixel index)
    for (int t = threadIdx.x; t < n_c; t += blockDim.x) {
      const uint32_t bt = a.cand_bits[base + t];
      const int32_t pt = a.cand_pix[base + t];
      int rank = 0;
      for (int u = 0; u < n_c; ++u) {
        const uint32_t bu = a.cand_bits[base + u];
        rank += (bu > bt || (bu == bt && a.cand_pix[base + u] < pt)) ? 1 : 0;
      }
      write_triplet(a, s.first + rank, s.img, pt);
    }
    return;
  }
  // fewer non-zero pixels than draws: k independent draws from the n_c collected pixels, proportionally to their
  // weights (torch.multinomial(..., replacement=True)).  The list is put in pixel order first so that the result does
  // not depend on the order the collect pass's atomics happened to take.
  const bool sorted = n_c <= kSortedInSmem;
  if (sorted) {
    for (int t = threadIdx.x; t < n_c; t += blockDim.x) {
      const int32_t pt = a.cand_pix[base + t];
      int rank = 0;
      for (int u = 0; u < n_c; ++u) rank += a.cand_pix[base + u] < pt ? 1 : 0;
      s_pix[rank] = pt;
      s_w[rank] = a.cand_w[base + t];
    }
    __syncthreads();
  }
  const int32_t* pix = sorted ? s_pix : a.cand_pix + base;
  const float* w = sorted ? s_w : a.cand_w + base;
  float total = 0.f;
  for (int u = 0; u < n_c; ++u) total += w[u];
  for (int t = threadIdx.x; t < s.k; t += blockDim.x) {
    const Philox4 r = philox4x32_10((uint32_t)t, (uint32_t)s.img, kSamplerStreamReplacement, 0u, a.seed_lo, a.seed_hi);
    const float target = unit_open(r.v[0]) * total;
    float cum = 0.f;
    int pick = n_c - 1;
    for (int u = 0; u < n_c; ++u) {
      cum += w[u];
      if (cum >= target) {
        pick = u;
        break;
      }
    }
    write_triplet(a, s.first + t, s.img, pix[pick]);
  }
}

static inline int64_t align16(int64_t n) { return (n + 15) / 16 * 16; }

}  // namespace kp

using namespace kp;

extern "C" int64_t kp_importance_pixels_scratch_bytes(int n_sel, int k_max) {
  if (n_sel <= 0 || k_max <= 0) return 0;
  return align16((int64_t)n_sel * (4 * 256 + 2) * 4) + 3 * align16((int64_t)n_sel * k_max * 4);
}

extern "C" int kp_importance_pixels(const void* weights_fp16, int B, int64_t HW, int width, const int32_t* sel, int n_sel,
                                    int k_max, uint64_t seed, void* scratch, int64_t* out, void* stream) {
  if (n_sel == 0) return 0;
  KP_CHECK(weights_fp16 && sel && scratch && out, "importance_pixels: NULL argument");
  KP_CHECK(B > 0 && HW > 0 && HW < (int64_t)1 << 31 && width > 0 && n_sel > 0 && n_sel <= 65535 && k_max > 0,
           "importance_pixels: B=%d HW=%lld width=%d n_sel=%d k_max=%d", B, (long long)HW, width, n_sel, k_max);
  cudaStream_t st = as_stream(stream);
  char* p = static_cast<char*>(scratch);
  const int64_t head = align16((int64_t)n_sel * (4 * 256 + 2) * 4), cand = align16((int64_t)n_sel * k_max * 4);
  SamplerArgs a;
  a.weights = static_cast<const __half*>(weights_fp16);
  a.B = B, a.HW = HW, a.width = width, a.sel = sel, a.n_sel = n_sel, a.k_max = k_max;
  a.seed_lo = (uint32_t)seed, a.seed_hi = (uint32_t)(seed >> 32);
  a.hist = reinterpret_cast<uint32_t*>(p);
  a.counters = a.hist + (int64_t)n_sel * 4 * 256;
  a.cand_bits = reinterpret_cast<uint32_t*>(p + head);
  a.cand_pix = reinterpret_cast<int32_t*>(p + head + cand);
  a.cand_w = reinterpret_cast<float*>(p + head + 2 * cand);
  a.out = out;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)head, st);
  KP_CHECK(e == cudaSuccess, "importance_pixels: memset failed: %s", cudaGetErrorString(e));
  // blocks per image: a chunk loop of 2048 pixels per block iteration, at least ~8 blocks per SM over the whole grid
  const int64_t per_image = ceil_div(HW, 2048);
  const int64_t want = ceil_div(148 * 8, n_sel);
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(per_image, std::max<int64_t>(want, 8)));
  const dim3 grid(gx, (unsigned)n_sel);
  sampler_hist_kernel<0><<<grid, 256, 0, st>>>(a);
  sampler_hist_kernel<1><<<grid, 256, 0, st>>>(a);
  sampler_hist_kernel<2><<<grid, 256, 0, st>>>(a);
  sampler_hist_kernel<3><<<grid, 256, 0, st>>>(a);
  sampler_collect_kernel<<<grid, 256, 0, st>>>(a);
  kp::g_launches += 5;
  sampler_finalize_kernel<<<(unsigned)n_sel, 256, 0, st>>>(a);
  KP_LAUNCH_CHECK("importance_pixels");
  return 0;
}
