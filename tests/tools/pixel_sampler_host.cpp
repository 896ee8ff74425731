// Host build of soccernerfs_b200/csrc/pixel_sampler_math.cuh (the device sampler's arithmetic is plain C++):
//   g++ -O1 -ffp-contract=off -I <repo>/soccernerfs_b200/csrc pixel_sampler_host.cpp -o pixel_sampler_host
// stdin (binary): int32 n, int32 image, uint64 seed, int32 k, float32 weights[n]
//                 (image == -1: the payload is uint32 key bits instead of weights, 0 = no key: hand-made keys with ties)
// stdout (binary): uint32 key_bits[n] (0 where the weight is 0), then what a four-pass radix select over those keys
//                  arrives at: uint32 threshold, int32 need, int32 take_all, uint32 nnz
// tests/test_device_sampler_math.py compares both with oracle/device_sampler.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pixel_sampler_math.cuh"

int main() {
  int32_t n, image, k;
  uint64_t seed;
  if (fread(&n, 4, 1, stdin) != 1 || fread(&image, 4, 1, stdin) != 1 || fread(&seed, 8, 1, stdin) != 1 || fread(&k, 4, 1, stdin) != 1) return 1;
  std::vector<float> w(n);
  if (fread(w.data(), 4, n, stdin) != (size_t)n) return 1;
  std::vector<uint32_t> bits(n, 0u);
  if (image == -1) {
    memcpy(bits.data(), w.data(), 4 * (size_t)n);
    for (int i = 0; i < n; ++i) w[i] = bits[i] ? 1.f : 0.f;
  }
  for (int64_t g = 0; image != -1 && g * 4 < n; ++g) {
    const kp::Philox4 r = kp::philox4x32_10((uint32_t)g, (uint32_t)image, kp::kSamplerStreamRace, (uint32_t)(g >> 32),
                                            (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int q = 0; q < 4 && g * 4 + q < n; ++q)
      if (w[g * 4 + q] > 0.f) bits[g * 4 + q] = kp::race_key_bits(w[g * 4 + q], r.v[q]);
  }
  // the four histogram passes exactly as the kernels run them
  std::vector<uint32_t> hist(4 * 256, 0u);
  uint32_t prefix = 0, nnz = 0;
  int need = k, take_all = 0;
  for (int pass = 0; pass < 4; ++pass) {
    if (pass > 0) {
      kp::select_walk(hist.data(), pass, k, &prefix, &need, &take_all, &nnz);
      if (take_all) break;
    }
    const int shift = 24 - 8 * pass;
    for (int i = 0; i < n; ++i) {
      if (w[i] <= 0.f) continue;
      if (pass == 0 || (bits[i] >> (shift + 8)) == prefix) hist[pass * 256 + ((bits[i] >> shift) & 255u)] += 1;
    }
  }
  kp::select_walk(hist.data(), 4, k, &prefix, &need, &take_all, &nnz);
  fwrite(bits.data(), 4, n, stdout);
  fwrite(&prefix, 4, 1, stdout);
  fwrite(&need, 4, 1, stdout);
  fwrite(&take_all, 4, 1, stdout);
  fwrite(&nnz, 4, 1, stdout);
  return 0;
}
