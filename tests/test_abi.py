"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/kplanes_b200.h declares (no compute calls: there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from soccernerfs_b200 import _lib

    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kplanes_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from soccernerfs_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported by the .so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_abi_version_and_error_channel(lib):
    assert lib.kp_abi_version() == 4
    assert isinstance(lib.kp_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    """Entry points validate arguments before touching the device: a bad call fails with a message, no crash."""
    from ctypes import c_void_p

    from soccernerfs_b200 import _lib

    pts = _lib.make_points(D=5)
    rc = lib.kp_hexplane_fwd(c_void_p(0), c_void_p(0), 1, 6, 32, pts, 0, 1, 0x3F, c_void_p(0), c_void_p(0))
    assert rc != 0 and b"points.D" in lib.kp_last_error()
    rc = lib.kp_uniform_bins(c_void_p(0), c_void_p(0), 0, c_void_p(0), c_void_p(0), 4, 8, 0, c_void_p(0), c_void_p(0), c_void_p(0),
                             c_void_p(0), c_void_p(0), c_void_p(0))
    assert rc != 0 and b"NULL" in lib.kp_last_error()


def test_no_cpu_fallback():
    """CPU tensors must be rejected loudly (the product path has no CPU/eager fallback)."""
    import torch

    from soccernerfs_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.get_weights(torch.ones(2, 3), torch.ones(2, 3))


def test_sass_is_sm100a():
    """The shipped library holds sm_100a code only."""
    import subprocess

    from soccernerfs_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
