"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the on-device importance pixel sampler's ARITHMETIC
(soccernerfs_b200/csrc/pixel_sampler_math.cuh + pixel_sampler.cu), the checker of ``kp_importance_pixels``.

What the device sampler replaces is DynamicBasedPixelSampler.sample_method's per-image ``torch.multinomial`` calls
(NS/data/pixel_samplers.py:396-398): k pixels proportional to the image's weight map, without replacement when the map
has >= k non-zero pixels, with replacement otherwise.  torch's CPU kernel realises the former as an exponential race
(q_i ~ Exp(1), keep the k largest w_i / q_i); the device sampler runs the same race on a counter-based generator
(Philox4x32-10), so this file can recompute every key.  It pins the device sampler's random stream; that the stream has
the reference's DISTRIBUTION is checked separately (chi-square tests in tests/test_gpu_sampler.py and, on the CPU,
tests/test_device_sampler_math.py against torch.multinomial's own frequencies).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_RACE, STREAM_REPLACEMENT = 0, 1


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32 with 10 rounds, the key bumped after every round (pixel_sampler_math.cuh).  Counters: uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ np.uint32(k0)
        n1 = (p1 & mask).astype(np.uint32)
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ np.uint32(k1)
        n3 = (p0 & mask).astype(np.uint32)
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def unit_open(r):
    """23 random bits -> (bits + 0.5) / 2^23 in fp32: exact, strictly inside (0, 1)."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


def race_keys(weights_row: np.ndarray, image: int, seed: int) -> np.ndarray:
    """fp32 keys w / Exp(1) of every pixel of one image (0 where the weight is 0).  weights_row: fp16 or fp32 [HW]."""
    w = np.asarray(weights_row).astype(np.float32).reshape(-1)
    hw = w.shape[0]
    groups = np.arange((hw + 3) // 4, dtype=np.uint64)
    r = philox4x32_10((groups & np.uint64(0xFFFFFFFF)).astype(np.uint32), np.uint32(image), np.uint32(STREAM_RACE),
                      (groups >> np.uint64(32)).astype(np.uint32), seed & 0xFFFFFFFF, seed >> 32)
    bits = np.stack(r, axis=-1).reshape(-1)[:hw]  # component pixel % 4 of group pixel // 4
    with np.errstate(divide="ignore"):
        keys = w / (-np.log(unit_open(bits)))
    return np.where(w > 0, keys, np.float32(0)).astype(np.float32)


def sample_image(weights_row: np.ndarray, image: int, k: int, seed: int) -> np.ndarray:
    """The k pixel indices the device sampler returns for one image, in its output order."""
    w = np.asarray(weights_row).astype(np.float32).reshape(-1)
    nz = np.nonzero(w > 0)[0]
    if len(nz) >= k:  # without replacement: the k largest keys, largest first, equal keys by pixel index
        keys = race_keys(w, image, seed)[nz]
        order = np.lexsort((nz, -keys.astype(np.float64)))
        return nz[order[:k]].astype(np.int64)
    # with replacement from the short list of non-zero pixels (pixel order), sequential fp32 sums
    cum = np.cumsum(w[nz], dtype=np.float32)
    t = np.arange(k, dtype=np.uint32)
    r0 = philox4x32_10(t, np.uint32(image), np.uint32(STREAM_REPLACEMENT), np.uint32(0), seed & 0xFFFFFFFF, seed >> 32)[0]
    target = unit_open(r0) * cum[-1]
    pick = np.minimum(np.searchsorted(cum, target, side="left"), len(nz) - 1)
    return nz[pick].astype(np.int64)


def select_threshold(keys_bits: np.ndarray, k: int):
    """What select_walk arrives at after four passes: (threshold bits T, number of keys == T still needed) for the k
    largest of ``keys_bits`` (uint32, more than k of them)."""
    s = np.sort(keys_bits)[::-1]
    t = s[k - 1]
    return int(t), int(k - np.count_nonzero(keys_bits > t))
