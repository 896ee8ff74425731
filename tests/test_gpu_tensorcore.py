"""tcgen05 (3xTF32) dense layer vs fp32 torch matmul."""
import pytest
import torch

from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k,act", [(128, 64, 32, 0), (300, 64, 128, 1), (1000, 16, 64, 0), (257, 3, 64, 2), (4096, 64, 31, 1)])
def test_tc_linear_matches_fp32(m, n, k, act):
    from ctypes import c_void_p

    from soccernerfs_b200 import _lib

    gen = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=gen).cuda()
    w = (torch.randn(n, k, generator=gen) / k**0.5).cuda()
    y = torch.full((m, n), float("nan"), device="cuda")
    _lib.call("kp_tc_linear_fwd", _lib.ptr(x), k, _lib.ptr(w), k, _lib.ptr(y), n, m, n, k, act, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    ref = torch.relu(ref) if act == 1 else (torch.sigmoid(ref) if act == 2 else ref)
    err = rel_err(y.cpu(), ref.float().cpu())
    fp32_err = rel_err((x @ w.t()).cpu() if act == 0 else y.cpu(), ref.float().cpu())
    assert err < 2e-6, (err, fp32_err)
    del c_void_p
