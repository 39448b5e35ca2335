"""Kernel micro-benchmarks at the C2 shapes (B=8): prints TFLOP/s / GB/s per kernel. CUDA-event timed."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops

dev = "cuda"
def bench(fn, iters=10, warmup=3):
    for _ in range(warmup): fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

res = []
def rec(name, ms, flops=None, bytes_=None):
    d = {"name": name, "ms": round(ms, 4)}
    if flops: d["tflops"] = round(flops / ms / 1e9, 1)
    if bytes_: d["gbs"] = round(bytes_ / ms / 1e6, 1)
    print(json.dumps(d), flush=True); res.append(d)

B = 8
for (H, C, N) in [(64, 320, 320), (32, 640, 640), (16, 1280, 1280), (8, 1280, 1280), (64, 960, 320), (32, 1920, 640), (16, 2560, 1280)]:
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    w = (torch.randn(N, 9 * C, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(B, H, H, N, device=dev, dtype=torch.bfloat16)
    ms = bench(lambda: ops.conv3x3(x, w, bias=bias, out=out))
    rec(f"conv3x3 {H}x{H} {C}->{N}", ms, flops=2 * B * H * H * 9 * C * N)
for (M, N, K, act) in [(32768, 320, 320, 0), (32768, 2560, 320, 4), (32768, 320, 1280, 0), (8192, 640, 640, 0), (8192, 5120, 640, 4),
                       (2048, 1280, 1280, 0), (2048, 10240, 1280, 4), (512, 1280, 1280, 0), (32768, 1024, 320, 0)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    ms = bench(lambda: ops.gemm(a, w, bias=bias, act=act))
    rec(f"gemm {M}x{N}x{K} act{act}", ms, flops=2 * M * N * K)
for (Hh, Nq, Nk, d) in [(8, 4096, 4096, 40), (8, 1024, 1024, 80), (8, 256, 256, 160), (8, 4096, 80, 40)]:
    dk, dv = ops.attention_pads(d)
    q = torch.randn(B * Nq, Hh * dk, device=dev).bfloat16()
    k = torch.randn(B * Nk, Hh * dk, device=dev).bfloat16()
    vt = torch.randn(Hh * dv, B * Nk, device=dev).bfloat16()
    out = torch.empty(B * Nq, Hh * d, device=dev, dtype=torch.bfloat16)
    nk_valid = 77 if Nk == 80 else Nk
    ms = bench(lambda: ops.attention(q, k, vt, out, B, Hh, Nq, nk_valid, d, kv_bstride=Nk))
    rec(f"attention N{Nq} M{nk_valid} d{d}", ms, flops=4 * B * Hh * Nq * nk_valid * d)
for (HW, C) in [(4096, 320), (1024, 640), (4096, 960)]:
    x = torch.randn(B, HW, C, device=dev).bfloat16()
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.empty_like(x)
    ms = bench(lambda: ops.groupnorm(x, g, b, 1e-5, act=1, out=out))
    rec(f"groupnorm+silu HW{HW} C{C}", ms, bytes_=3 * x.numel() * 2)
x = torch.randn(32768, 320, device=dev).bfloat16(); g, b = torch.randn(320, device=dev), torch.randn(320, device=dev)
out = torch.empty_like(x)
rec("layernorm 32768x320", bench(lambda: ops.layernorm(x, g, b, out=out)), bytes_=2 * x.numel() * 2)
json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
