"""Per-tile role timeline of CTA (0,0,0) of one self-attention launch (softmax warp 2 and the MMA issuer).
Needs a library built with -DVDB_TIMELINE (the stamps are compiled out of the product build):
    make -C versatile-diffusion_b200/csrc clean all EXTRA=-DVDB_TIMELINE
    [VDB_ATT_BKV=64|643|128] python tools/attention_timeline.py [N=4096] [d=40]
Prints, per kv tile, nanoseconds since the first stamp: where a softmax warp's ~1.9 us per tile actually goes."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops
from vdb200._lib import lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 40
B, H = 8, 8
dk, dv = ops.attention_pads(d)
q = torch.randn(B * N, H * dk, device="cuda").bfloat16()
k = torch.randn(B * N, H * dk, device="cuda").bfloat16()
vt = torch.randn(H * dv, B * N, device="cuda").bfloat16()
out = torch.empty(B * N, H * d, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(q, k, vt, out, B, H, N, N, d)
tl = torch.zeros(256, dtype=torch.int64, device="cuda")
lib.vdb_debug_attention_timeline.argtypes = [ctypes.c_void_p]
lib.vdb_debug_attention_timeline(tl.data_ptr())
torch.cuda.synchronize()
ops.attention(q, k, vt, out, B, H, N, N, d)
torch.cuda.synchronize()
lib.vdb_debug_attention_timeline(None)
t = tl.cpu().view(16, 16)
if not int(t[0, 0]):
    sys.exit("no stamps: rebuild the library with EXTRA=-DVDB_TIMELINE")
t0 = int(t[0, 0])
names = {0: "sm_want_S", 1: "sm_S_ready", 2: "sm_scores+max", 3: "sm_max_xchg", 4: "sm_exp+P", 5: "sm_O_settled", 6: "sm_arrived",
         8: "mma_want_P", 9: "mma_P_ready", 10: "mma_S_issued", 11: "mma_PV_issued"}
print("tile " + " ".join(f"{n:>13s}" for n in names.values()))
for j in range(16):
    print(f"{j:4d} " + " ".join(f"{(int(t[j, i]) - t0) if int(t[j, i]) else -1:13d}" for i in names))
print("\nper-tile deltas of the softmax warp (ns): wait_S, ld+max, exchange, exp+store, settle, fence+arrive, tile total")
for j in range(1, 15):
    v = [int(t[j, i]) for i in range(7)]
    nxt = int(t[j + 1, 0])
    print(f"{j:4d} " + " ".join(f"{b - a:8d}" for a, b in zip(v[:-1], v[1:])) + f" {nxt - v[0]:8d}")
