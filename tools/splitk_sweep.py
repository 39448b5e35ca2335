"""(BN, split-K) sweep of the small-M launches: the 8x8 / 16x16 convs of the UNet and the M = 8 weight-streaming GEMMs of the 0-D
diffuser.  Every configuration rotates through enough weight copies that a replay never finds its weights in the 126 MB L2
(as in the real step); time = CUDA-graph replay of `reps` launches / reps.   usage: python tools/splitk_sweep.py"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
from vdb200 import ops

dev = "cuda"
REPS = 12


def time_graph(fns):
    for f in fns[:2]:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / len(fns) * 1000.0)
    return best


def copies(nbytes):
    return max(2, min(REPS, int(300e6 // nbytes) + 1))


def sweep(label, make, grid):
    row = []
    for bn, ks in grid:
        try:
            row.append(f"bn{bn}/ks{ks}:{time_graph(make(bn, ks)):6.1f}")
        except Exception as ex:  # noqa
            row.append(f"bn{bn}/ks{ks}:  err")
    print(f"{label:44s} " + " ".join(row), flush=True)


GRID = [(0, 0), (64, 1), (128, 1), (256, 1), (64, 2), (128, 2), (128, 3), (128, 4), (256, 2), (256, 4), (256, 7), (160, 4)]

# ---- convs: B 8, HxW, C -> N
for H, C, N in [(8, 1280, 1280), (8, 2560, 1280), (16, 1280, 1280), (16, 2560, 1280)]:
    x = torch.randn(8, H, H, C, device=dev).bfloat16()
    nw = copies(N * 9 * C * 2)
    ws = [(torch.randn(N, 9 * C, device=dev) * 0.01).bfloat16() for _ in range(nw)]
    b = torch.randn(N, device=dev)
    out = torch.empty(8, H, H, N, device=dev, dtype=torch.bfloat16)

    def make(bn, ks, x=x, ws=ws, b=b, out=out):
        return [(lambda w=ws[i % len(ws)]: ops.conv3x3(x, w, bias=b, out=out, bn=bn, ksplit=ks)) for i in range(REPS)]
    sweep(f"conv3x3 B8 {H}x{H} C{C}->N{N} ({N*9*C*2/1e6:.0f} MB w, {nw} copies)", make, GRID)

# ---- GEMMs (M rows, weight [N, K])
for M, N, K in [(512, 1280, 1280), (2048, 1280, 5120), (2048, 1280, 1280), (8, 5120, 5120), (8, 5120, 10240), (8, 1280, 1280), (8, 2560, 5120), (32, 1280, 1280)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    nw = copies(N * K * 2)
    ws = [(torch.randn(N, K, device=dev) * 0.01).bfloat16() for _ in range(nw)]
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)

    def make(bn, ks, a=a, ws=ws, b=b, out=out):
        return [(lambda w=ws[i % len(ws)]: ops.gemm(a, w, bias=b, out=out, bn=bn, ksplit=ks)) for i in range(REPS)]
    sweep(f"gemm M{M} N{N} K{K} ({N*K*2/1e6:.0f} MB w, {nw} copies)", make, GRID)

# ---- the CUDA-core skinny kernel on the same M = 8 shapes (fp32 activations)
for M, N, K in [(8, 5120, 5120), (8, 1280, 1280), (8, 2560, 5120)]:
    if M * K * 4 > 200 * 1024:
        print(f"linear_small M{M} N{N} K{K}: x does not fit shared memory"); continue
    a = torch.randn(M, K, device=dev)
    nw = copies(N * K * 2)
    ws = [(torch.randn(N, K, device=dev) * 0.01).bfloat16() for _ in range(nw)]
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    fns = [(lambda w=ws[i % len(ws)]: ops.linear_small(a, w, b, out=out)) for i in range(REPS)]
    print(f"linear_small M{M} N{N} K{K}: {time_graph(fns):6.1f} us ({N*K*2/1e6:.0f} MB w)", flush=True)
