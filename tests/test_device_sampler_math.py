"""CPU checks of the on-device importance pixel sampler (soccernerfs_b200/csrc/pixel_sampler*.cu*): its arithmetic header
compiled for the host vs the numpy restatement (oracle/device_sampler.py), and the restatement's distribution vs what it
stands in for -- the reference's per-image torch.multinomial (NS/data/pixel_samplers.py:396-398)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

from oracle import device_sampler as ds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("sampler") / "pixel_sampler_host"
    subprocess.run([gxx, "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "soccernerfs_b200", "csrc"),
                    os.path.join(ROOT, "tests", "tools", "pixel_sampler_host.cpp"), "-o", str(out)], check=True)
    return str(out)


def _run(binary, w, image, seed, k):
    w = np.asarray(w, dtype=np.float32)
    req = struct.pack("<iiQi", len(w), image, seed, k) + w.tobytes()
    res = subprocess.run([binary], input=req, stdout=subprocess.PIPE, check=True).stdout
    bits = np.frombuffer(res[: 4 * len(w)], dtype=np.uint32)
    prefix, need, take_all, nnz = struct.unpack("<IiiI", res[4 * len(w):])
    return bits, prefix, need, take_all, nnz


def test_philox_reference_vector():
    """Philox4x32-10 known-answer test (Random123 kat_vectors: counter and key all ones -> 408f276d 41c83b0e a20bc7c6
    6d5451fd; all zeros -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8)."""
    ones = np.array([0xFFFFFFFF], dtype=np.uint32)
    out = ds.philox4x32_10(ones, ones, ones, ones, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(o[0]) for o in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    zero = np.zeros(1, dtype=np.uint32)
    out = ds.philox4x32_10(zero, zero, zero, zero, 0, 0)
    assert [int(o[0]) for o in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_host_build_of_the_sampler_math_matches_the_restatement(host_binary):
    rng = np.random.default_rng(5)
    for n, nnz_frac, k in ((4099, 0.3, 10), (1000, 1.0, 50), (64, 0.1, 10), (777, 0.02, 3), (5000, 1.0, 1)):
        w = rng.random(n).astype(np.float32).astype(np.float16).astype(np.float32)
        w[rng.random(n) > nnz_frac] = 0
        seed = int(rng.integers(0, 2**63))
        image = int(rng.integers(0, 500))
        bits, prefix, need, take_all, nnz = _run(host_binary, w, image, seed, k)
        keys = ds.race_keys(w, image, seed)
        ref_bits = keys.view(np.uint32)
        # same Philox words and the same IEEE division; logf may differ from numpy's log in the last place
        assert np.array_equal(bits == 0, w == 0)
        assert np.abs(bits.astype(np.int64) - ref_bits.astype(np.int64)).max() <= 4
        assert np.mean(bits == ref_bits) > 0.5
        if nnz_frac * n > 4 * k:  # a last-place difference does not change which pixels win (keys are far apart)
            top = np.argsort(-bits.astype(np.int64), kind="stable")[:k]
            assert set(top.tolist()) == set(ds.sample_image(w, image, k, seed).tolist())
        assert nnz == int(np.count_nonzero(w))
        if nnz <= k:
            assert take_all == 1
            continue
        assert take_all == 0
        t, nd = ds.select_threshold(bits[w > 0], k)  # the radix select on the host build's own keys
        assert (prefix, need) == (t, nd)
        chosen = np.nonzero((bits > prefix) & (w > 0))[0]
        assert len(chosen) == k - need and need >= 1 and np.count_nonzero(bits == prefix) >= need


def test_radix_select_with_tied_keys(host_binary):
    """Equal keys at the threshold: the four-pass select of the host build (histograms + select_walk exactly as the kernels
    run them) must arrive at the k-th largest key and ask for exactly the missing number of ties -- hand-made keys with
    many duplicates, differing in every byte position."""
    rng = np.random.default_rng(9)
    base = np.array([0x3F800000, 0x3F800001, 0x3F800100, 0x3F810000, 0x40000000, 0x3F7FFFFF, 0x00000001], dtype=np.uint32)
    keys = np.concatenate([np.repeat(base, rng.integers(1, 9, size=len(base))), np.zeros(5, dtype=np.uint32)])
    rng.shuffle(keys)
    n_keys = int(np.count_nonzero(keys))
    for k in range(1, n_keys):
        req = struct.pack("<iiQi", len(keys), -1, 0, k) + keys.tobytes()
        res = subprocess.run([host_binary], input=req, stdout=subprocess.PIPE, check=True).stdout
        prefix, need, take_all, nnz = struct.unpack("<IiiI", res[4 * len(keys):])
        t, nd = ds.select_threshold(keys[keys > 0], k)
        assert (prefix, need, take_all, nnz) == (t, nd, 0, n_keys), k
        assert np.count_nonzero(keys > prefix) + need == k and 1 <= need <= np.count_nonzero(keys == prefix)
    req = struct.pack("<iiQi", len(keys), -1, 0, n_keys) + keys.tobytes()  # k == number of keys: everything is taken
    res = subprocess.run([host_binary], input=req, stdout=subprocess.PIPE, check=True).stdout
    assert struct.unpack("<IiiI", res[4 * len(keys):])[2:] == (1, n_keys)


def test_race_has_the_multinomial_distribution():
    """The exponential race of the restatement draws like torch.multinomial: single draws follow w / sum(w) (chi-square
    against the exact probabilities), and k draws without replacement have torch.multinomial's inclusion frequencies."""
    w = np.array([0.0, 0.5, 0.25, 0.0, 1.0, 0.125, 2.0, 0.0, 0.75, 0.375, 0.0, 1.5], dtype=np.float16)
    p = w.astype(np.float64) / w.astype(np.float64).sum()
    n_draws = 20000
    counts = np.zeros(len(w))
    for image in range(n_draws):
        counts[ds.sample_image(w, image, 1, seed=1234)[0]] += 1
    assert counts[w == 0].sum() == 0
    nz = w > 0
    chi2 = float((((counts - n_draws * p) ** 2)[nz] / (n_draws * p[nz])).sum())
    assert chi2 < 40.0, chi2  # 7 degrees of freedom: P(chi2 > 40) ~ 1e-6
    # inclusion frequencies of 3 draws without replacement vs torch.multinomial's own
    k, n = 3, 20000
    inc = np.zeros(len(w))
    for image in range(n):
        s = ds.sample_image(w, image, k, seed=99)
        assert len(set(s.tolist())) == k
        inc[s] += 1
    g = torch.Generator().manual_seed(0)
    tw = torch.from_numpy(w.astype(np.float32))
    ref = np.zeros(len(w))
    for _ in range(n):
        ref[torch.multinomial(tw, k, replacement=False, generator=g).numpy()] += 1
    # two independent estimates of the same inclusion probabilities: binomial standard error ~ sqrt(n p (1-p)) <= 71
    assert np.abs(inc - ref).max() < 6 * np.sqrt(2) * 71, (inc, ref)
    # with replacement when the map has fewer non-zero pixels than draws
    w2 = np.zeros(40, dtype=np.float16)
    w2[[3, 17]] = [1.0, 3.0]
    c = np.zeros(40)
    for image in range(4000):
        s = ds.sample_image(w2, image, 5, seed=7)
        assert set(s.tolist()) <= {3, 17}
        np.add.at(c, s, 1)
    assert abs(c[17] / c.sum() - 0.75) < 0.02
