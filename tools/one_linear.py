import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops
for (M, K, N, ai) in [(8, 320, 1280, 0), (8, 1280, 1280, 0), (8, 1280, 19520, 1), (2, 1280, 19520, 1), (16, 1280, 19520, 1)]:
    x = torch.randn(M, K, device="cuda"); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16(); b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for _ in range(3): ops.linear_small(x, w, b, act_in=ai, out=out)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): ops.linear_small(x, w, b, act_in=ai, out=out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f"linear_small M={M} K={K} N={N}: {ms*1e3:.1f} us  ({N*K*2/ms/1e6:.0f} GB/s of weights)")
