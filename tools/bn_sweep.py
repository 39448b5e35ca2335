"""BN sweep of the tensor-core GEMM over the UNet's linear shapes (CUDA-graph replay of 20 launches each, so the
numbers are device time without the CPU launch path).  usage: python tools/bn_sweep.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops

SHAPES = [  # (M, N, K, resid, act, label)
    (32768, 320, 320, True, 0, "L0 proj/to_out"), (32768, 1024, 320, False, 0, "L0 qk"), (32768, 512, 320, False, 0, "L0 q"),
    (32768, 2560, 320, False, 4, "L0 ff1 geglu"), (32768, 320, 1280, True, 0, "L0 ff2"), (384, 32768, 320, False, 0, "L0 v^T"),
    (8192, 640, 640, True, 0, "L1 proj/to_out"), (8192, 2048, 640, False, 0, "L1 qk"), (8192, 5120, 640, False, 4, "L1 ff1"),
    (8192, 640, 2560, True, 0, "L1 ff2"), (2048, 1280, 1280, True, 0, "L2 proj"), (2048, 3072, 1280, False, 0, "L2 qk"),
    (2048, 10240, 1280, False, 4, "L2 ff1"), (2048, 1280, 5120, True, 0, "L2 ff2"), (512, 1280, 1280, True, 0, "L3 proj"),
]


def time_graph(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1000.0


for M, N, K, resid, act, label in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda")
    n_out = N // 2 if act == 4 else N
    r = torch.randn(M, n_out, device="cuda").bfloat16() if resid else None
    out = torch.empty(M, n_out, device="cuda", dtype=torch.bfloat16)
    row = []
    for bn in ([0, 64, 128, 256] if act == 4 else [0, 64, 128, 160, 256]):
        try:
            us = time_graph(lambda: ops.gemm(a, w, bias=b, resid=r, out=out, act=act, bn=bn, ksplit=1 if bn else 0))
            row.append(f"bn{bn}:{us:7.1f}")
        except Exception as ex:  # noqa
            row.append(f"bn{bn}: err")
    print(f"{label:16s} M{M:6d} N{N:6d} K{K:5d}  " + "  ".join(row) + f"   ({2.0*M*N*K/1e6:.0f} MFLOP)", flush=True)
