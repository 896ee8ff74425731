"""tcgen05 dense layer (4-term TF32 split, fp32-accurate) vs an fp64 torch matmul."""
import pytest
import torch

from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k,act", [(128, 64, 32, 0), (300, 64, 128, 1), (1000, 16, 64, 0), (257, 3, 64, 2), (4096, 64, 31, 1),
                                         (1000, 128, 192, 1), (515, 128, 128, 0), (300, 96, 256, 2)])
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
    assert err < 5e-6, (err, fp32_err)
    del c_void_p


@pytest.mark.parametrize("m,n,k", [(128, 64, 32), (300, 64, 128), (1000, 16, 64), (257, 3, 64), (4096, 64, 31), (777, 128, 16),
                                     (1000, 128, 192), (515, 128, 128), (300, 200, 96)])
def test_tc_linear_backward_matches_fp32(m, n, k):
    """dX = (dY W) * mask and dW += dY^T X through the MN-major operand views."""
    from soccernerfs_b200 import _lib

    gen = torch.Generator().manual_seed(m * 7 + n + k)
    x = torch.randn(m, k, generator=gen).cuda()
    w = (torch.randn(n, k, generator=gen) / k**0.5).cuda()
    dy = torch.randn(m, n, generator=gen).cuda()
    aux = torch.randn(m, k, generator=gen).cuda()
    dx = torch.full((m, k), float("nan"), device="cuda")
    _lib.call("kp_tc_linear_bwd_data", _lib.ptr(dy), n, _lib.ptr(w), k, _lib.ptr(dx), k, m, n, k, _lib.ptr(aux), k, _lib.stream_ptr())
    ref_dx = (dy.double() @ w.double()) * (aux > 0)
    assert rel_err(dx.cpu(), ref_dx.float().cpu()) < 5e-6
    dw = torch.ones(n, k, device="cuda")  # accumulates on top of existing contents
    _lib.call("kp_tc_linear_bwd_weight", _lib.ptr(dy), n, _lib.ptr(x), k, _lib.ptr(dw), k, m, n, k, _lib.stream_ptr())
    ref_dw = 1.0 + dy.double().t() @ x.double()
    assert rel_err(dw.cpu(), ref_dw.float().cpu()) < 5e-6
