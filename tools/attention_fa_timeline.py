"""Role timeline of CTA (0,0,0) of the two-tile attention kernel (attention_fa_kernel): first warp of each softmax
warpgroup and the MMA issuer, first 16 kv tiles.  Needs a -DVDB_TIMELINE build (tools/build_timeline_lib.sh):
    VDB200_LIB=$PWD/tools/bin/libvdb200_tl.so [VDB_ATT_FA=11] python tools/attention_fa_timeline.py [N=4096] [d=40]
Slots per tile: group g: 8g+7 loop top, +0 S ready, +1 S in registers (s_free), +2 max / rescale decided, +3 PV(j-1) retired,
+4 token acquired, +5 exp2 + P written (token passed next), +6 arrived on p_full;  MMA: 16+3g P_g ready, 17+3g PV_g issued,
18+3g next S issued (S_{1-g}(j+1) after PV_0, S_0(j+2) after PV_1)."""
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
tl = torch.zeros(16 * 24, dtype=torch.int64, device="cuda")
lib.vdb_debug_attention_timeline.argtypes = [ctypes.c_void_p]
lib.vdb_debug_attention_timeline(tl.data_ptr())
torch.cuda.synchronize()
ops.attention(q, k, vt, out, B, H, N, N, d)
torch.cuda.synchronize()
lib.vdb_debug_attention_timeline(None)
t = tl.cpu().view(16, 24)
if not int(t[1, 0]):
    sys.exit("no stamps: rebuild the library with EXTRA=-DVDB_TIMELINE (and a shape served by the two-tile kernel)")
t0 = min(int(x) for x in t.flatten() if int(x))
cols = [("g0_top", 7), ("g0_S", 0), ("g0_ld", 1), ("g0_max", 2), ("g0_pv", 3), ("g0_tok", 4), ("g0_exp", 5), ("g0_arr", 6),
        ("g1_top", 15), ("g1_S", 8), ("g1_ld", 9), ("g1_max", 10), ("g1_pv", 11), ("g1_tok", 12), ("g1_exp", 13), ("g1_arr", 14),
        ("m_P0", 16), ("m_PV0", 17), ("m_S1n", 18), ("m_P1", 19), ("m_PV1", 20), ("m_S0n", 21)]
print("ns since the first stamp (env VDB_ATT_FA=%s)" % os.environ.get("VDB_ATT_FA", "default"))
print("tile " + " ".join(f"{n:>7s}" for n, _ in cols))
for j in range(16):
    print(f"{j:4d} " + " ".join(f"{(int(t[j, i]) - t0) if int(t[j, i]) else -1:7d}" for _, i in cols))
for g in (0, 1):
    print(f"\ngroup {g} per-tile phases (ns): wait_S  ld  max  wait_pv  wait_token  exp2+P  arrive | tile total")
    for j in range(2, 15):
        top, S, ld, mx, pv, tok, ex, arr = (int(t[j, 8 * g + i]) for i in (7, 0, 1, 2, 3, 4, 5, 6))
        nxt = int(t[j + 1, 8 * g + 7])
        print(f"{j:4d} {S - top:7d} {ld - S:5d} {mx - ld:5d} {pv - mx:7d} {tok - pv:9d} {ex - tok:8d} {arr - ex:7d} | {nxt - top:7d}")
