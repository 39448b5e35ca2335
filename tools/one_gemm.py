import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops
M, N, K = [int(v) for v in sys.argv[1:4]]
resid = len(sys.argv) > 4 and sys.argv[4] == "resid"
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda").bfloat16() if resid else None
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.gemm(a, w, bias=b, resid=r, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(3): ops.gemm(a, w, bias=b, resid=r, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
# event timing of 20 back-to-back launches (amortises CPU launch cost)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): ops.gemm(a, w, bias=b, resid=r, out=out)
e.record(); torch.cuda.synchronize()
print("avg ms back-to-back", s.elapsed_time(e) / 20)
