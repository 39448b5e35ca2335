"""Phase timeline of CTA (0,0) of the single-launch GroupNorm (thread 0).  Needs a -DVDB_TIMELINE build:
    make -C versatile-diffusion_b200/csrc clean all EXTRA=-DVDB_TIMELINE      (or VDB200_LIB=<that build>)
    python tools/gn_timeline.py [HW=4096] [C=320] [B=8]"""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops
from vdb200._lib import lib

HW = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C = int(sys.argv[2]) if len(sys.argv) > 2 else 320
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
x = torch.randn(B, HW, C, device="cuda").bfloat16()
g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
out = torch.empty_like(x)
for _ in range(3):
    ops.groupnorm(x, g, b, 1e-5, act=1, out=out)
tl = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.vdb_debug_gn_timeline.argtypes = [ctypes.c_void_p]
lib.vdb_debug_gn_timeline(tl.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.groupnorm(x, g, b, 1e-5, act=1, out=out); e1.record()
torch.cuda.synchronize()
lib.vdb_debug_gn_timeline(None)
t = [int(v) for v in tl.cpu()]
if not t[0]:
    sys.exit("no stamps: rebuild the library with EXTRA=-DVDB_TIMELINE")
names = ["start", "stats loads+accumulate", "reduce+publish", "grid arrival wait", "fold partials", "scale/shift table", "normalise+store"]
print(f"GroupNorm B={B} HW={HW} C={C}: {e0.elapsed_time(e1) * 1000:.1f} us launch-to-end (events); CTA(0,0) phases in ns:")
for i in range(1, 7):
    print(f"  {names[i]:26s} {t[i] - t[i - 1]:7d}")
print(f"  {'total inside the kernel':26s} {t[6] - t[0]:7d}")
