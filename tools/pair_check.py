"""CTA-pair (cta_group::2) igemm bring-up: a few GEMM / conv shapes against a torch fp32 reference.
Run with VDB_PAIR=1 (and VDB_PAIR=0 for the single-CTA numbers)."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200"))
import torch
import torch.nn.functional as F
from vdb200 import ops
from vdb200._lib import lib
lib.vdb_debug_pair_launches.restype = ctypes.c_longlong
torch.manual_seed(0)
dev = "cuda"


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1000


def report(name, out, ref, us, flops):
    out, ref = out.float().flatten(), ref.float().flatten()
    cos = F.cosine_similarity(out, ref, dim=0).item()
    err = (out - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    print(f"{name:34s} cos {cos:.6f} relmax {err:.4f} {us:8.1f} us {flops / us / 1e6:7.1f} TF/s  pair_launches={lib.vdb_debug_pair_launches()}",
          flush=True)
    return cos > 0.9995 and err < 2e-2


ok = True
for (M, N, K, resid, act) in [(256, 256, 64, False, 0), (512, 160, 128, True, 0), (32768, 320, 320, True, 0), (8192, 3072, 640, False, 0),
                              (32768, 2560, 320, False, 4), (32768, 320, 1280, True, 0), (2048, 1280, 5120, True, 1)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    n_out = N // 2 if act == 4 else N
    r = torch.randn(M, n_out, device=dev).bfloat16() if resid else None
    out = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, bias=b, resid=r, out=out, act=act, ksplit=1)
    fn(); torch.cuda.synchronize()
    y = a.float() @ w.float().t() + b
    if act == 4:
        # packed GEGLU tiles: per 256 columns, first 128 = value, last 128 = gate
        yv = y.view(M, N // 256, 2, 128)
        y = (yv[:, :, 0] * F.gelu(yv[:, :, 1])).reshape(M, n_out)
    elif act == 1:
        y = F.silu(y)
    if resid:
        y = y + r.float()
    ok &= report(f"gemm {M}x{N}x{K} act{act} resid{int(resid)}", out, y, timeit(fn), 2.0 * M * N * K)
for (B, H, C, N) in [(8, 64, 320, 320), (8, 64, 960, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (4, 128, 256, 256)]:
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    w = (torch.randn(N, C, 3, 3, device=dev) * 0.02).bfloat16()
    b = torch.randn(N, device=dev)
    wp = w.permute(0, 2, 3, 1).reshape(N, -1).contiguous()
    out = torch.empty(B, H, H, N, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.conv3x3(x, wp, bias=b, out=out, ksplit=1)
    fn(); torch.cuda.synchronize()
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    ok &= report(f"conv3x3 B{B} {H}x{H} {C}->{N}", out, y, timeit(fn), 2.0 * B * H * H * 9 * C * N)
print("PAIR_CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
