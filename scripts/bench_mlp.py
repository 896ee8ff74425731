"""Micro-benchmark of the decoder GEMM kernels (tcgen05 3xTF32) at the cfg2 shapes.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from soccernerfs_b200 import _lib

M = 4096 * 48
dev = "cuda"
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (N, K) in [(64, 128), (16, 64), (64, 32), (64, 64), (3, 64)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); y = torch.empty(M, N, device=dev)
    dy = torch.randn(M, N, device=dev); dx = torch.empty(M, K, device=dev); dw = torch.zeros(N, K, device=dev)
    s = _lib.stream_ptr()
    f = t(lambda: _lib.call("kp_tc_linear_fwd", _lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(y), N, M, N, K, 1, s))
    bd = t(lambda: _lib.call("kp_tc_linear_bwd_data", _lib.ptr(dy), N, _lib.ptr(w), K, _lib.ptr(dx), K, M, N, K, _lib.ptr(x), K, s))
    bw = t(lambda: _lib.call("kp_tc_linear_bwd_weight", _lib.ptr(dy), N, _lib.ptr(x), K, _lib.ptr(dw), K, M, N, K, s))
    mm = t(lambda: torch.relu(x @ w.t()))
    print(f"N={N:3d} K={K:3d}: tc fwd {f:7.1f} us  bwd_data {bd:7.1f} us  bwd_weight {bw:7.1f} us   | torch fp32 matmul+relu {mm:7.1f} us")
