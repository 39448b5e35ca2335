"""Does putting co-resident attention CTAs in anti-phase help?  Sweeps the start delay of every second CTA wave
(vdb_debug_attention_stagger) for the 64x64-level self-attention and prints in-graph time per launch."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops
from vdb200._lib import lib
lib.vdb_debug_attention_stagger.argtypes = [ctypes.c_int]
B, H, N, d = 8, 8, 4096, 40
dk, dv = ops.attention_pads(d)
q = torch.randn(B * N, H * dk, device="cuda").bfloat16(); k = torch.randn(B * N, H * dk, device="cuda").bfloat16()
vt = torch.randn(H * dv, B * N, device="cuda").bfloat16(); out = torch.empty(B * N, H * d, device="cuda", dtype=torch.bfloat16)
ref = None
for ns in [0, 300, 600, 1000, 1500, 0]:
    lib.vdb_debug_attention_stagger(ns)
    fn = lambda: ops.attention(q, k, vt, out, B, H, N, N, d)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    if ref is None: ref = out.clone()
    same = torch.equal(ref, out)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 10 * 1000)
    print(f"stagger {ns:5d} ns: {best:8.2f} us per launch  (bit-identical to stagger 0: {same})", flush=True)
