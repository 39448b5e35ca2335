"""In-graph cost of every kernel launch of ONE DDIM step (C2 workload: bs 4, CFG -> B = 8, 64x64 latent).

ncu's launch list serialises kernels with cold caches and per-launch events include the host launch path; neither is the
regime of the captured step.  This tool intercepts the C-ABI calls of one eager DDIM step (function + arguments), then
replays EACH launch alone, `reps` times back to back inside its own CUDA graph, and times the replay with CUDA events:
the kernel's duration as it runs inside the step graph (L2-warm, no host path).  The sum over the launches against the
measured duration of the real step graph gives the launch-gap / dependency overhead of the ~410-kernel chain.

    python tools/step_breakdown.py [reps=10] > gpurun_out/step_breakdown.txt

(Replays reuse the recorded device pointers after the step's activations have been freed to torch's caching allocator:
the memory stays mapped and only ever held activations, so the replays are safe but their OUTPUT is meaningless.)"""
import ctypes
import os
import sys
from collections import OrderedDict

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "versatile-diffusion_b200"))
TEXT = "--text" in sys.argv                      # the text-latent (i2t) step of the full-size 0-D diffuser instead of the C2 image step
if TEXT:
    os.environ["VDB_TEXT_FLOWS"] = "1"
_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
REPS = int(_pos[0]) if _pos else 10
import torch  # noqa: E402
if not TEXT:
    import bench  # noqa: E402
from lib.model_zoo.ddim import DDIMSampler  # noqa: E402
from vdb200 import _lib, ops  # noqa: E402

# torch.cuda.graph() calls torch.cuda.empty_cache() on entry: the cached blocks the recorded pointers live in would be
# unmapped before the replay (first GPU run of this tool: "illegal memory access" on every replay).  Keep them mapped.
torch.cuda.empty_cache = lambda: None
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
if TEXT:
    from lib.cfg_helper import model_cfg_bank  # noqa: E402
    from lib.model_zoo import get_model  # noqa: E402
    cfg = model_cfg_bank()('vd_four_flow_v1-0')
    cfg.args.ctx_cfg_list = []
    cfg.args.vae_cfg_list = []
    torch.manual_seed(0)
    with torch.device(dev):
        net = get_model()(cfg, verbose=False)
    gd = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for _, p_ in net.named_parameters():
            if (p_.ndim == 1 and not bool((p_ == 1).all())) or not bool(p_.any()):
                p_.normal_(0.0, 0.02, generator=gd)
    net.eval()
    net.to(dev)
    xT = torch.randn(4, 768, generator=g).to(dev)
    c = (torch.randn(4, 257, 768, generator=g) * 0.5).to(dev)
    u = torch.zeros(4, 257, 768, device=dev)
    SHAPE, XT, CT = [4, 768], "text", "image"
else:
    net = bench.build_net(dev)
    xT = torch.randn(4, 4, 64, 64, generator=g).to(dev)
    c = (torch.randn(4, 77, 768, generator=g) * 0.5).to(dev)
    u = (torch.randn(4, 77, 768, generator=g) * 0.5).to(dev)
    SHAPE, XT, CT = [4, 4, 64, 64], "image", "text"


def sample(sampler, steps):
    with torch.no_grad():
        return sampler.sample(steps=steps, shape=SHAPE, x_info={"type": XT, "xt": xT},
                              c_info={"type": CT, "conditioning": c, "unconditional_conditioning": u,
                                      "unconditional_guidance_scale": 7.5}, verbose=False, eta=0.)[0]


# ---- 1. the real thing: 50-step sample through the captured step graph
S = DDIMSampler(net)
for _ in range(2):
    sample(S, 50)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sample(S, 50); e1.record(); torch.cuda.synchronize()
step_us = e0.elapsed_time(e1) * 1000.0 / 50

# ---- 2. record the launches of one eager step (the second of a 2-step sample: no per-sample context projections)
records = []
real = {}


_PARTS = ctypes.c_int(0)      # vdb_gemm_ln_bf16 producers report their partial count through a host int*: the caller's is gone at replay


def make_wrapper(name, fn):
    def wrapper(*args):
        rc = fn(*args)
        if name == "vdb_gemm_ln_bf16" and args[22]:
            args = args[:22] + (ctypes.addressof(_PARTS),) + args[23:]
        records.append((name, fn, args))
        return rc
    return wrapper


E = DDIMSampler(net, use_cuda_graph=False)
sample(E, 2)                                    # warm: packs, workspaces
SKIP = {"vdb_version", "vdb_last_error", "vdb_launch_count", "vdb_reset_launch_count", "vdb_num_sms", "vdb_attention_dk_pad",
        "vdb_attention_dv_pad", "vdb_groupnorm_nsplit", "vdb_groupnorm_scratch_floats"}
for name in _lib.SIGNATURES:
    if name in SKIP:
        continue
    real[name] = getattr(_lib.lib, name)
    setattr(ops.lib, name, make_wrapper(name, real[name]))        # ops.lib is the same CDLL object: shadow the attribute
try:
    sample(E, 2)
finally:
    for name, fn in real.items():
        setattr(ops.lib, name, fn)
torch.cuda.synchronize()
# keep the launches of the LAST step: everything after the last vdb_ddim_cfg_step of step 0
idx = [i for i, r in enumerate(records) if r[0] == "vdb_ddim_cfg_step"]
step = records[idx[0] + 1: idx[1] + 1] if len(idx) >= 2 else records
step = [r for r in step if r[0] != "vdb_permute_f32"]              # boundary conversions are outside the captured step


def label(name, a):
    """family + the shape arguments that identify the launch (positions follow include/vdb200.h)."""
    if name == "vdb_gemm_bf16":
        return f"gemm M{a[1]} N{a[8]} K{a[2] + a[5]} act{a[18]}{' +res' if a[13] else ''}"
    if name == "vdb_gemm_ln_bf16":
        return f"gemm M{a[1]} N{a[5]} K{a[2]} act{a[12]}{' +res' if a[8] else ''}{' ln-in' if a[13] else ''}{'-cols' if a[19] else ''}{' stats-out' if a[21] else ''}"
    if name == "vdb_gemm_skinny_bf16":
        return f"gemm_skinny S{a[1]} R{a[8]} K{a[2] + a[5]}{' +res' if a[12] else ''}{' transposed' if a[16] else ''}"
    if name == "vdb_conv3x3_bf16":
        return f"conv3x3 B{a[1]} {a[2]}x{a[3]} C{a[4]}+{a[10]}+{a[12]} -> N{a[7]} mode{a[5]}"
    if name == "vdb_attention_bf16":
        return f"attention B{a[10]} H{a[11]} Nq{a[12]} Nk{a[13]} d{a[16]}"
    if name == "vdb_groupnorm_nhwc":
        return f"groupnorm B{a[4]} HW{a[5]} C{a[1] + a[3]} act{a[10]}"
    if name == "vdb_layernorm":
        return f"layernorm rows{a[1]} C{a[2]}"
    return name.replace("vdb_", "")


def time_one(fn, args):
    stream = torch.cuda.current_stream().cuda_stream
    a = list(args)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        cs = torch.cuda.current_stream().cuda_stream
        a[-1] = cs                                                   # every entry point takes the stream last
        for _ in range(REPS):
            rc = fn(*a)
            if rc:
                raise RuntimeError(_lib.lib.vdb_last_error())
    gr.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); gr.replay(); s1.record(); torch.cuda.synchronize()
        best = min(best, s0.elapsed_time(s1) * 1000.0 / REPS)
    del gr
    return best


def time_family(items):
    """all launches of one family, recorded order, ONE pass per replay: weights arrive from HBM as in the real step (a launch
    replayed alone keeps its weights in the 126 MB L2, which flatters the weight-streaming GEMMs of the 0-D diffuser)"""
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        cs = torch.cuda.current_stream().cuda_stream
        for _, fn, args in items:
            a = list(args); a[-1] = cs
            if fn(*a):
                raise RuntimeError(_lib.lib.vdb_last_error())
    gr.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); gr.replay(); s1.record(); torch.cuda.synchronize()
        best = min(best, s0.elapsed_time(s1) * 1000.0)
    del gr
    return best


agg = OrderedDict()
total = 0.0
for name, fn, args in step:
    try:
        us = time_one(fn, args)
    except Exception as ex:  # noqa
        print(f"# could not replay {name}: {str(ex)[:100]}")
        continue
    total += us
    k = label(name, args)
    d = agg.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += us

print(f"one DDIM step inside the captured graph: {step_us:9.1f} us  ({len(step)} launches recorded)")
print(f"sum of the launches, each replayed alone in-graph: {total:9.1f} us  -> chain overhead (gaps, tails, cold L2) {step_us - total:8.1f} us "
      f"= {100.0 * (step_us - total) / step_us:.1f} % of the step")
print(f"{'launch (shape)':64s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:64s} {n:4d} {us:10.1f} {us / n:9.2f} {100.0 * us / total:6.1f}%")
fam = OrderedDict()
for k, (n, us) in agg.items():
    f = k.split()[0]
    d = fam.setdefault(f, [0, 0.0])
    d[0] += n
    d[1] += us
print("\nby family:")
for f, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"  {f:20s} n={n:4d} {us:10.1f} us {100.0 * us / total:6.1f}%")

print("\nby family, one pass over the family's launches per replay (cold weights, as in the step):")
byf = OrderedDict()
for r in step:
    byf.setdefault(label(r[0], r[2]).split()[0], []).append(r)
tot_f = 0.0
for f, items in byf.items():
    try:
        us = time_family(items)
    except Exception as ex:  # noqa
        print(f"  {f}: could not replay ({str(ex)[:80]})")
        continue
    tot_f += us
    print(f"  {f:20s} n={len(items):4d} {us:10.1f} us  ({us / len(items):7.2f} us per launch)")
print(f"  sum {tot_f:10.1f} us of the {step_us:.1f} us step")
try:
    print(f"  whole step's launches in one graph, recorded order: {time_family(step):10.1f} us")
except Exception as ex:  # noqa
    print(f"  whole step: could not replay ({str(ex)[:80]})")
