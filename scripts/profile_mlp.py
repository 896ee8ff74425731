"""One launch of each decoder GEMM kernel between cudaProfilerStart/Stop (for ncu), at a cfg2 and a cfg3 layer shape.

    ncu --profile-from-start off --set full --clock-control none --import-source on -o /tmp/mlp python scripts/profile_mlp.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from soccernerfs_b200 import _lib

dev = "cuda"
s = _lib.stream_ptr()
cases = [(4096 * 48, 64, 128), (4096 * 64, 128, 192)]
bufs = []
for (M, N, K) in cases:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); y = torch.empty(M, N, device=dev)
    dy = torch.randn(M, N, device=dev); dx = torch.empty(M, K, device=dev); dw = torch.zeros(N, K, device=dev)
    bufs.append((M, N, K, x, w, y, dy, dx, dw))


def run():
    for (M, N, K, x, w, y, dy, dx, dw) in bufs:
        _lib.call("kp_tc_linear_fwd", _lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(y), N, M, N, K, 1, s)
        _lib.call("kp_tc_linear_bwd_data", _lib.ptr(dy), N, _lib.ptr(w), K, _lib.ptr(dx), K, M, N, K, _lib.ptr(x), K, s)
        _lib.call("kp_tc_linear_bwd_weight", _lib.ptr(dy), N, _lib.ptr(x), K, _lib.ptr(dw), K, M, N, K, s)


run()
torch.cuda.synchronize()
flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
