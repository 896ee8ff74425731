// Memory-hierarchy probes (sm_100a): the achievable rate of the field kernels' OWN access pattern at one level of the
// hierarchy, measured on the box the bench runs on -- the denominators bench.py reports the gather / scatter against
// (MEASURED_PEAKS.json only holds a streaming-copy HBM figure; there is no L2 figure and no random-line figure).
//
// Pattern = what hexplane_fwd / hexplane_bwd do to memory and nothing else: 8 adjacent lanes own one 128-byte line
// (one float4 each) of a pseudo-random texel of a buffer of `n_lines` lines; `unroll` independent lines are in flight
// per lane.  mode 0: ld.global.nc.v4 (gather), mode 1: red.global.add.v4.f32 (scatter).  The working-set size decides
// the level: a buffer well inside the 126 MB L2 gives the L2 rate, a multi-GB buffer gives the random-line HBM rate.
#include "common.cuh"

namespace kp {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // lowbias32: a cheap, well-mixed permutation-like hash
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) line_probe_kernel(float* __restrict__ buf, uint32_t n_lines, int iters, uint32_t seed,
                                                         float* __restrict__ sink) {
  constexpr int U = 8;  // independent lines in flight per lane group
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < iters; it += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t line = mix32(seed + group * (uint32_t)iters + (uint32_t)(it + u)) % n_lines;
      float* p = buf + (size_t)line * 32 + sub * 4;
      if (MODE == 0) v[u] = ldg4(p);
      else red_add_v4(p, make_float4(1.f, 1.f, 1.f, 1.f));
    }
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  if (MODE == 0 && acc.x + acc.y + acc.z + acc.w == 1.2345e30f) sink[0] = acc.x;  // keeps the loads alive
}

}  // namespace kp

using namespace kp;

// One launch touching `groups * iters` pseudo-random 128-byte lines of buf[0 : n_lines*32 floats].  Returns the number
// of lines touched in *lines_touched (host).  iters is rounded up to a multiple of 8.
extern "C" int kp_line_probe(float* buf, int64_t n_lines, int mode, int blocks, int iters, uint32_t seed, float* sink,
                             int64_t* lines_touched, void* stream) {
  KP_CHECK(buf && sink && n_lines >= 1 && n_lines < (1ll << 32) && (mode == 0 || mode == 1) && blocks >= 1 && iters >= 1,
           "line_probe: bad arguments");
  iters = (iters + 7) / 8 * 8;
  if (mode == 0) line_probe_kernel<0><<<blocks, 256, 0, as_stream(stream)>>>(buf, (uint32_t)n_lines, iters, seed, sink);
  else line_probe_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(buf, (uint32_t)n_lines, iters, seed, sink);
  KP_LAUNCH_CHECK("line_probe");
  if (lines_touched) *lines_touched = (int64_t)blocks * 32 * iters;
  return 0;
}
