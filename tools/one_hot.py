"""One launch each of the step's three heaviest kernel shapes inside a cudaProfilerStart/Stop window, for
`ncu --set full --profile-from-start off`: conv3x3 64x64 320->320, conv3x3 64x64 960->320 (B=8), self-attention
N=4096 d=40 (B=8, 8 heads)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops

dev, B = "cuda", 8
cases = []
for (H, C, N) in [(64, 320, 320), (64, 960, 320)]:
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    w = (torch.randn(N, 9 * C, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(B, H, H, N, device=dev, dtype=torch.bfloat16)
    cases.append(lambda x=x, w=w, bias=bias, out=out: ops.conv3x3(x, w, bias=bias, out=out))
Hh, Nq, d = 8, 4096, 40
dk, dv = ops.attention_pads(d)
q = torch.randn(B * Nq, Hh * dk, device=dev).bfloat16()
k = torch.randn(B * Nq, Hh * dk, device=dev).bfloat16()
vt = torch.randn(Hh * dv, B * Nq, device=dev).bfloat16()
o = torch.empty(B * Nq, Hh * d, device=dev, dtype=torch.bfloat16)
cases.append(lambda: ops.attention(q, k, vt, o, B, Hh, Nq, Nq, d, kv_bstride=Nq))
x = torch.randn(B, 4096, 320, device=dev).bfloat16()
g, b = torch.randn(320, device=dev), torch.randn(320, device=dev)
y = torch.empty_like(x)
cases.append(lambda: ops.groupnorm(x, g, b, 1e-5, act=1, out=y))
for _ in range(3):
    for c in cases:
        c()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for c in cases:
    c()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
