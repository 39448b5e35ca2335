"""Per-tile role timeline of CTA 0 of one GEMM launch.  Needs a library built with -DVDB_TIMELINE (the stamps are
compiled out of the product build): make -C versatile-diffusion_b200/csrc EXTRA=-DVDB_TIMELINE"""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops
from vdb200._lib import lib
M, N, K = [int(v) for v in sys.argv[1:4]]
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
b = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.gemm(a, w, bias=b, out=out)
tl = torch.zeros(256, dtype=torch.int64, device="cuda")
lib.vdb_debug_igemm_timeline.argtypes = [ctypes.c_void_p]; lib.vdb_debug_igemm_timeline(tl.data_ptr())
torch.cuda.synchronize()
ops.gemm(a, w, bias=b, out=out)
torch.cuda.synchronize()
lib.vdb_debug_igemm_timeline(None)
t = tl.cpu().view(16, 16)
t0 = int(t[0, 0])
names = ["prod_start", "mma_want", "mma_got", "mma_issued", "epi_wait", "epi_got", "epi_done", "c0_regs", "c0_staged", "c0_k0", "c0_stored"]
for it in range(3):
    print(it, {n: (int(t[it, i]) - t0 if int(t[it, i]) else None) for i, n in enumerate(names)})
