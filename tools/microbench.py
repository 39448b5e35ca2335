"""Kernel micro-benchmarks at the C2 shapes (B=8): prints TFLOP/s / GB/s per kernel. CUDA-event timed.
    python tools/microbench.py [conv,gemm,attention,groupnorm,layernorm] [out.json]
"ms" = one launch after an L2 flush (includes ~3 us of launch latency); "graph_us" = per launch inside a CUDA graph of
20 back-to-back launches (L2-warm, no host launch path) -- the figure that matters inside the captured DDIM step."""
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

def graph_us(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / reps * 1000.0)
    return best

ONLY = set(sys.argv[1].split(",")) if len(sys.argv) > 1 and sys.argv[1] else None
OUT = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/microbench.json"
def want(fam): return ONLY is None or fam in ONLY

res = []
def rec(name, ms, flops=None, bytes_=None, gus=None):
    d = {"name": name, "ms": round(ms, 4)}
    if gus is not None:
        d["graph_us"] = round(gus, 2)
        if flops: d["graph_tflops"] = round(flops / gus / 1e6, 1)
        if bytes_: d["graph_gbs"] = round(bytes_ / gus / 1e3, 1)
    if flops: d["tflops"] = round(flops / ms / 1e9, 1)
    if bytes_: d["gbs"] = round(bytes_ / ms / 1e6, 1)
    print(json.dumps(d), flush=True); res.append(d)

B = 8
for (H, C, N) in [] if not want("conv") else [(64, 320, 320), (32, 640, 640), (16, 1280, 1280), (8, 1280, 1280), (64, 960, 320), (32, 1920, 640), (16, 2560, 1280)]:
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    w = (torch.randn(N, 9 * C, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(B, H, H, N, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.conv3x3(x, w, bias=bias, out=out)
    rec(f"conv3x3 {H}x{H} {C}->{N}", bench(fn), flops=2 * B * H * H * 9 * C * N, gus=graph_us(fn))
for (M, N, K, act) in [] if not want("gemm") else [(32768, 320, 320, 0), (32768, 2560, 320, 4), (32768, 320, 1280, 0), (8192, 640, 640, 0), (8192, 5120, 640, 4),
                       (2048, 1280, 1280, 0), (2048, 10240, 1280, 4), (512, 1280, 1280, 0), (32768, 1024, 320, 0)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    o = torch.empty(M, N // 2 if act == 4 else N, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, bias=bias, act=act, out=o)
    rec(f"gemm {M}x{N}x{K} act{act}", bench(fn), flops=2 * M * N * K, gus=graph_us(fn))
for (Hh, Nq, Nk, d) in [] if not want("attention") else [(8, 4096, 4096, 40), (8, 1024, 1024, 80), (8, 256, 256, 160), (8, 4096, 80, 40)]:
    dk, dv = ops.attention_pads(d)
    q = torch.randn(B * Nq, Hh * dk, device=dev).bfloat16()
    k = torch.randn(B * Nk, Hh * dk, device=dev).bfloat16()
    vt = torch.randn(Hh * dv, B * Nk, device=dev).bfloat16()
    out = torch.empty(B * Nq, Hh * d, device=dev, dtype=torch.bfloat16)
    nk_valid = 77 if Nk == 80 else Nk
    fn = lambda: ops.attention(q, k, vt, out, B, Hh, Nq, nk_valid, d, kv_bstride=Nk)
    rec(f"attention N{Nq} M{nk_valid} d{d}", bench(fn), flops=4 * B * Hh * Nq * nk_valid * d, gus=graph_us(fn))
for (HW, C) in [] if not want("groupnorm") else [(4096, 320), (1024, 640), (4096, 960), (1024, 1920), (256, 1280), (256, 2560), (64, 1280), (64, 2560)]:
    x = torch.randn(B, HW, C, device=dev).bfloat16()
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.empty_like(x)
    fn = lambda: ops.groupnorm(x, g, b, 1e-5, act=1, out=out)
    rec(f"groupnorm+silu HW{HW} C{C}", bench(fn), bytes_=2 * x.numel() * 2, gus=graph_us(fn))
for (rows, C) in [] if not want("layernorm") else [(32768, 320), (8192, 640), (2048, 1280), (512, 1280)]:
    x = torch.randn(rows, C, device=dev).bfloat16(); g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.empty_like(x)
    fn = lambda: ops.layernorm(x, g, b, out=out)
    rec(f"layernorm {rows}x{C}", bench(fn), bytes_=2 * x.numel() * 2, gus=graph_us(fn))
os.makedirs(os.path.dirname(OUT) or ".", exist_ok=True)
json.dump({"env": {k: v for k, v in os.environ.items() if k.startswith("VDB_")}, "results": res}, open(OUT, "w"), indent=1)
