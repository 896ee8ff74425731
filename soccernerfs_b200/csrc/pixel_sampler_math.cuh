// Arithmetic of the on-device importance pixel sampler (csrc/pixel_sampler.cu), kept free of CUDA-only constructs so the
// same functions compile for the host (tests/tools/pixel_sampler_host.cpp checks them against oracle/device_sampler.py).
//
// Sampling k pixels of a weight map WITHOUT replacement, proportionally to the weights, is what torch.multinomial does
// for DynamicBasedPixelSampler (NS/data/pixel_samplers.py:396-398); torch's CPU kernel realises it as an exponential
// race: q_i ~ Exp(1) per category, keep the k largest w_i / q_i.  The same race is run here with a counter-based
// generator, so that any thread can (re)compute the key of any pixel: Philox4x32-10 on the counter (pixel / 4, image,
// stream, 0) under the 64-bit seed, component pixel % 4.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define KP_HD __host__ __device__ __forceinline__
#else
#define KP_HD inline
#endif

namespace kp {

struct Philox4 {
  uint32_t v[4];
};

KP_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.v[0] = c0, o.v[1] = c1, o.v[2] = c2, o.v[3] = c3;
  return o;
}

// 23 random bits -> u = (bits + 0.5) / 2^23, exactly representable and strictly inside (0, 1)
KP_HD float unit_open(uint32_t r) { return ((float)(r >> 9) + 0.5f) * (1.0f / 8388608.0f); }

KP_HD uint32_t float_bits(float x) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(x);
#else
  uint32_t b;
  memcpy(&b, &x, 4);
  return b;
#endif
}

// key of the exponential race: w / Exp(1), w > 0.  Positive and finite, so its bit pattern orders like its value.
KP_HD uint32_t race_key_bits(float w, uint32_t r) { return float_bits(w / (-logf(unit_open(r)))); }

enum { kSamplerStreamRace = 0, kSamplerStreamReplacement = 1 };

// Radix select of the k largest keys from four 256-bin histograms (most significant byte first).  hist[p] counts, over
// the pixels whose top 8p key bits equal the prefix chosen by passes 0..p-1, the next key byte; hist[0] therefore counts
// every non-zero pixel.  With `passes` passes done: *take_all = 1 when the map has at most k non-zero pixels (everything is
// taken, no threshold); else the chosen prefix (8 * passes bits) and how many keys are still needed from inside it.
// After 4 passes the prefix is the threshold key T itself: keys > T are all taken (k - need of them) plus `need` keys == T.
KP_HD void select_walk(const uint32_t* hist, int passes, int k, uint32_t* prefix, int* need, int* take_all, uint32_t* nnz) {
  uint32_t total = 0;
  for (int b = 0; b < 256; ++b) total += hist[b];
  *nnz = total;
  *prefix = 0;
  *need = k;
  *take_all = (total <= (uint32_t)k) ? 1 : 0;
  if (*take_all) return;
  uint32_t pre = 0;
  int nd = k;
  for (int p = 0; p < passes; ++p) {
    const uint32_t* h = hist + p * 256;
    int b = 255;
    uint32_t above = 0;
    while (b > 0 && above + h[b] < (uint32_t)nd) {
      above += h[b];
      --b;
    }
    nd -= (int)above;
    pre = (pre << 8) | (uint32_t)b;
  }
  *prefix = pre;
  *need = nd;
}

}  // namespace kp
