"""bench.py — headline benchmark of the B200-native Versatile-Diffusion sampling hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch: DDIMSampler.sample (50 DDIM steps, CFG 7.5, eta 0)
+ VD_v2_0.vae_decode for bs=4 512x512 images per GPU (BASELINE.json configs[1]: text-to-image single flow,
bf16).  Weights are random-init at the full architecture size (no checkpoints offline), contexts are
synthetic [bs,77,768] tensors (CLIP encoding is per-prompt and amortised), x_T is seeded noise.

Prints ONE JSON line (rank 0): value = whole-job images/s with inputs resident in HBM; e2e = the same
through the public API with pinned-host inputs (H2D of contexts + x_T, D2H of the uint8-able images) inside
the timed region; roofline = dominant kernel family timed with CUDA events inside this run; cpu_baseline =
the oracle port timed on the host cores on a bounded sample.  `--impl reference` times the reference's CPU
path (oracle port of lib/model_zoo, all host threads) on the same config with bounded samples.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "versatile-diffusion_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "512x512 images/sec @ 50-step DDIM (bs=4/GPU)"
UNIT = "images/s"
BS, LAT, DDIM_STEPS, SCALE = 4, 64, 50, 7.5
SEED = 100
# BASELINE.json configs[1..3]; FLOPs per 512^2 image = 100 UNet rows (50 steps x CFG pair) + VAE decode
CONFIGS = {
    "c2": {"workload": "t2i single-flow 512x512, 50-step DDIM, CFG 7.5, eta 0, bs 4/GPU + VAE decode (configs[1])",
           "ctx": [("text", 77, 1.0)], "flop_per_image": 80.33e12 + 2.5145e12},
    "c3": {"workload": "image-variation flow (CLIP-image context, 257 tokens) 512x512, 50-step DDIM, CFG 7.5, bs 4/GPU + VAE decode (configs[2])",
           "ctx": [("image", 257, 1.0)], "flop_per_image": 81.85e12 + 2.5145e12},
    "c4": {"workload": "dual-context (text 0.7 + image 0.3) guided generation 512x512, 50-step DDIM, CFG 7.5, bs 4/GPU + VAE decode (configs[3])",
           "ctx": [("text", 77, 0.7), ("image", 257, 0.3)], "flop_per_image": 120.0e12 + 2.5145e12},
}
FLOP_PER_IMAGE = CONFIGS["c2"]["flop_per_image"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_net(device):
    """Full-size VD (image VAE + 2D diffuser + text-context blocks), random init on the GPU, bf16 compute."""
    import torch
    from lib.cfg_helper import model_cfg_bank
    from lib.model_zoo import get_model
    cfg = model_cfg_bank()('vd_four_flow_v1-0')
    cfg.args.ctx_cfg_list = []            # contexts are synthetic here (CLIP is timed separately)
    torch.manual_seed(0)
    with torch.device(device):
        net = get_model()(cfg, verbose=False)
    g = torch.Generator(device=device).manual_seed(1)
    with torch.no_grad():                  # zero_module() tensors and biases -> N(0, 0.02) (SURVEY §8c pitfall)
        for _, p in net.named_parameters():
            if p.ndim == 1 or not bool(p.any()):
                if p.ndim == 1 and p.shape[0] > 0 and bool((p == 1).all()):
                    continue               # norm scales stay at 1
                p.normal_(0.0, 0.02, generator=g)
    net.eval()
    net.to(device)
    return net


def run_product(args):
    import torch
    import torch.distributed as dist
    from lib.model_zoo.ddim import DDIMSampler
    from vdb200 import ops, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg = CONFIGS[args.config]
    net = build_net(device)
    sampler = DDIMSampler(net)

    # contexts: rank 0 "encodes" (synthetic) and broadcasts over NCCL — the only collective of the path
    # (vdb200.parallel.broadcast_context); uncond of the image context is zeros as in app.py:345
    g = torch.Generator().manual_seed(2)
    ctx_h = []
    for ctype, L, ratio in cfg["ctx"]:
        c = (torch.randn(1, L, 768, generator=g) * 0.5).pin_memory()
        u = (torch.zeros(1, L, 768) if ctype == "image" else torch.randn(1, L, 768, generator=g) * 0.5).pin_memory()
        ctx_h.append((ctype, ratio, c, u))
    ctx_d = [(t, r, c.to(device), u.to(device)) for t, r, c, u in ctx_h]
    if rank != 0:
        for _, _, c, u in ctx_d:
            c.zero_(); u.zero_()                      # only rank 0 holds the encoded contexts before the broadcast
    parallel.broadcast_context([x for _, _, c, u in ctx_d for x in (c, u)])
    # this rank's rows of the GLOBAL batch, x_T drawn per global row index: an N-rank run reproduces the 1-rank rows
    rows = parallel.shard_rows(BS * world, rank, world)
    xT_h = parallel.seeded_latents(rows, (4, LAT, LAT), seed=SEED).pin_memory()
    xT_d = xT_h.to(device)
    ctx_rep = [(t, r, c.repeat(BS, 1, 1).contiguous(), u.repeat(BS, 1, 1).contiguous()) for t, r, c, u in ctx_d]
    img_h = torch.empty(BS, 3, 8 * LAT, 8 * LAT, dtype=torch.float32).pin_memory()

    def sample_with(smp, x0, ctxs, steps=DDIM_STEPS):
        xi = {"type": "image", "xt": x0}
        if len(ctxs) == 1:
            t, _, c, u = ctxs[0]
            return smp.sample(steps=steps, shape=[BS, 4, LAT, LAT], x_info=xi,
                              c_info={"type": t, "conditioning": c, "unconditional_conditioning": u,
                                      "unconditional_guidance_scale": SCALE}, verbose=False, eta=0.)[0]
        return smp.sample_multicontext(steps=steps, shape=[BS, 4, LAT, LAT], x_info=xi,
                                       c_info_list=[{"type": t, "conditioning": c, "unconditional_conditioning": u,
                                                     "unconditional_guidance_scale": SCALE, "ratio": r} for t, r, c, u in ctxs],
                                       verbose=False, eta=0.)[0]

    def one_pass(host_io):
        if host_io:
            x0 = xT_h.to(device, non_blocking=True)
            ctxs = [(t, r, c.to(device, non_blocking=True).repeat(BS, 1, 1), u.to(device, non_blocking=True).repeat(BS, 1, 1))
                    for t, r, c, u in ctx_h]
        else:
            x0, ctxs = xT_d, ctx_rep
        x = sample_with(sampler, x0, ctxs)
        im = net.vae_decode(x, "image")
        if host_io:
            img_h.copy_(im, non_blocking=True)
        return x, im

    def timed(n, host_io):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            one_pass(host_io)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            one_pass(False)
        one_pass(True)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ops.reset_launch_count()
        ms = timed(args.steps, False)
        # launches: graph replays do not pass through the C ABI, so count one DDIM step and scale
        per_step = getattr(sampler, "last_step_launches", 0)
        # ---- output check of the TIMED path (VERDICT r1 #1b): the graph-replayed sampler must reproduce an eager run of
        # the same kernels bit for bit, and every image must be finite and inside [0, 1]
        x_g, im_g = one_pass(False)
        x_e = sample_with(DDIMSampler(net, use_cuda_graph=False), xT_d, ctx_rep)
        check = {"graph_equals_eager_bitwise": bool(torch.equal(x_g, x_e)), "finite": bool(torch.isfinite(im_g).all()),
                 "image_min": round(float(im_g.min()), 4), "image_max": round(float(im_g.max()), 4),
                 "latent_std": round(float(x_g.float().std()), 4)}
        if world > 1:   # sharded rows: rank r's x_T rows are rows [4r, 4r+4) of the 1-rank draw (vdb200.parallel.seeded_latents)
            check["rows"] = list(rows)
        del x_e
        decode_launches = 0
        c0 = ops.launch_count()
        net.vae_decode(xT_d, "image")
        decode_launches = ops.launch_count() - c0
        launches = args.steps * (per_step * DDIM_STEPS + decode_launches + 4)
        ms_e2e = timed(args.steps, True)
        clk = clocks.stop() if rank == 0 else None

        # ---- roofline leg: per-family CUDA-event timing of one eager DDIM step + decode
        roof, fam = None, None
        if rank == 0:
            eager = DDIMSampler(net, use_cuda_graph=False)
            ops.profile_start()
            sample_with(eager, xT_d, ctx_rep, steps=2)
            fam = ops.profile_stop()
            peaks = measured_peaks()
            # dominant kernel of the step = igemm_kernel (tcgen05 implicit-GEMM mainloop): it serves both the conv3x3 and
            # the gemm families (49 % of the step in the ncu launch list, profiles/)
            ig_ms = sum(fam[k]["ms"] for k in ("conv3x3", "gemm") if k in fam)
            ig_fl = sum(fam[k]["flops"] for k in ("conv3x3", "gemm") if k in fam)
            ig_by = sum(fam[k]["bytes"] for k in ("conv3x3", "gemm") if k in fam)
            ig_n = sum(fam[k]["launches"] for k in ("conv3x3", "gemm") if k in fam)
            tflops = ig_fl / ig_ms / 1e9 if ig_ms > 0 else 0.0
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "igemm_traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            roof = {"bound": "tensor", "kernel": "igemm_kernel (conv3x3 + gemm families)", "achieved": round(tflops, 1),
                    "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": round(tflops / peaks["tflops_sustained"], 4),
                    "traffic": traffic, "algorithmic_bytes_per_launch": round(ig_by / max(ig_n, 1)),
                    "algorithmic_flops_per_launch": round(ig_fl / max(ig_n, 1)),
                    "peak_source": peaks["source"] + " (sustained: timed inside a long step)", "launches": ig_n,
                    "avg_launch_ms": round(ig_ms / max(ig_n, 1), 4),
                    "how": "CUDA events around every launch of one eager DDIM step pair on the launching stream "
                           "(small launches include host launch latency, so this under-states the kernel)"}
            roof["families"] = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                    "tflops": round(v["flops"] / v["ms"] / 1e9, 1) if v["ms"] > 0 and v["flops"] else None,
                                    "gbs": round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] > 0 else None}
                                for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}

    attn_frac = None
    if roof is not None and fam and "attention" in fam and fam["attention"]["ms"] > 0:
        # the metric's "attn TC-util %": attention-core FLOPs (4 B h Nq Nk d, unpadded) / event time / measured sustained peak
        attn_frac = round(fam["attention"]["flops"] / fam["attention"]["ms"] / 1e9 / measured_peaks()["tflops_sustained"], 4)
    images = BS * world * args.steps
    value = images / (ms / 1e3)
    e2e_value = images / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_baseline_sample(net, tuple(cfg["ctx"])) if world == 1 and not args.no_cpu_baseline else None
    peaks = measured_peaks()
    # ---- roofline refinement (last GPU work of the run, single GPU only): the SAME igemm launches of two eager DDIM steps,
    # re-issued back to back inside one CUDA graph and timed with CUDA events -> the kernel's launch duration without the host
    # launch path that the per-launch events above include.  Any failure leaves the eager figures in place.
    if world == 1 and roof is not None and not args.no_graph_roofline:
        try:
            with torch.no_grad():
                eager = DDIMSampler(net, use_cuda_graph=False)
                ops.record_start()
                sample_with(eager, xT_d, ctx_rep, steps=2)
                recs = ops.record_stop()
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    ops.replay(recs)
                gr.replay()
                torch.cuda.synchronize()
                best = None
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    gr.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    t = e0.elapsed_time(e1)
                    best = t if best is None else min(best, t)
            g_fl, g_n = sum(r[3] for r in recs), len(recs)
            eager_ms = roof["avg_launch_ms"] * roof["launches"]
            if g_n == roof["launches"] and best and 0.2 * eager_ms < best <= 1.05 * eager_ms:
                g_tflops = g_fl / best / 1e9
                roof["achieved_eager_events"] = roof["achieved"]
                roof["frac_eager_events"] = roof["frac"]
                roof["avg_launch_ms_eager_events"] = roof["avg_launch_ms"]
                roof["achieved"] = round(g_tflops, 1)
                roof["frac"] = round(g_tflops / peaks["tflops_sustained"], 4)
                roof["avg_launch_ms"] = round(best / g_n, 4)
                roof["how"] = ("the %d igemm launches of two eager DDIM steps re-issued back to back inside one CUDA graph, CUDA events "
                               "around the replay on the launching stream (best of 3); '*_eager_events' = per-launch events in the eager "
                               "step, which include the host launch latency" % g_n)
            else:
                roof["graph_replay"] = "discarded (launches %s vs %s, %.3f ms vs eager %.3f ms)" % (g_n, roof["launches"], best or -1.0, eager_ms)
            del recs, gr
        except BaseException as ex:  # noqa: the bench line must still be printed
            roof["graph_replay"] = "failed: %s" % (str(ex)[:200],)
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights, synthetic context)",
        "config": {"workload": cfg["workload"], "name": args.config,
                   "global_batch": BS * world, "latent": [4, LAT, LAT], "parallelism": f"dp{world} (batch shards, one NCCL context broadcast)",
                   "l2": "working set (3.3 GB weights + activations) exceeds the 126 MB L2 every step; no explicit flush",
                   "tensor_frac_of_step": round(value / world * cfg["flop_per_image"] / (peaks["tflops_sustained"] * 1e12), 4),
                   "attn_tensor_frac": attn_frac},
        "e2e": {"value": round(e2e_value, 4), "unit": UNIT,
                "h2d_bytes_per_step": int(xT_h.numel() * 4 + sum((c.numel() + u.numel()) * 4 for _, _, c, u in ctx_h)),
                "d2h_bytes_per_step": int(img_h.numel() * 4)},
        "gpu_launches": int(launches), "clocks": clk, "check": check, "roofline": roof, "cpu_baseline": cpu,
    }
    if not (check["graph_equals_eager_bitwise"] and check["finite"]):
        line["invalid"] = "output check failed: the timed path does not reproduce the eager kernels / non-finite images"
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_state_dict(net=None):
    """fp32 CPU weights for the oracle port: copied from the product net when given, else synthesised."""
    import torch
    if net is not None:
        return {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
    from lib.cfg_helper import model_cfg_bank
    from lib.model_zoo import get_model
    cfg = model_cfg_bank()('vd_four_flow_v1-0')
    cfg.args.ctx_cfg_list = []
    torch.manual_seed(0)
    m = get_model()(cfg, verbose=False)
    with torch.no_grad():
        for p in m.parameters():
            if not bool(p.any()):
                p.normal_(0.0, 0.02)
    return {k: v.detach().float() for k, v in m.state_dict().items()}


def host_info():
    """(physical cores, CPU model string) of this host; torchrun exports OMP_NUM_THREADS=1, so the CPU legs set the
    thread count explicitly instead of inheriting it (64 threads vs 1 under torchrun)."""
    cores, model = set(), "unknown"
    try:
        phys = core = None
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and model == "unknown":
                model = line.split(":", 1)[1].strip()
            elif line.startswith("physical id"):
                phys = line.split(":", 1)[1].strip()
            elif line.startswith("core id"):
                core = line.split(":", 1)[1].strip()
            elif not line.strip():
                if phys is not None and core is not None:
                    cores.add((phys, core))
                phys = core = None
    except OSError:
        pass
    n = len(cores) or (os.cpu_count() or 1)
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    return max(n, 1), model


def _cpu_time_step(sd, reps, rows=1, ctx=(("text", 77, 1.0),)):
    """One CFG UNet evaluation for `rows` images (B = 2 * rows, latent 64x64) + one K4-equivalent update, oracle port."""
    import torch
    from oracle import vd_oracle as O
    g = torch.Generator().manual_seed(7)
    x = torch.randn(rows, 4, LAT, LAT, generator=g)
    cs = [torch.randn(rows, L, 768, generator=g) * 0.5 for _, L, _ in ctx]
    us = [torch.randn(rows, L, 768, generator=g) * 0.5 for _, L, _ in ctx]
    kw = {} if len(ctx) == 1 and ctx[0][0] == "text" else {"c_types": tuple(t for t, _, _ in ctx)}
    if len(ctx) > 1:
        kw["ratios"] = [r for _, _, r in ctx]
    sched = O.ddim_schedule(O.ddpm_schedule()["alphas_cumprod"], DDIM_STEPS)
    ts = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            O.p_sample_ddim(sd, x, cs, us, torch.tensor([981] * rows), DDIM_STEPS - 1, sched, SCALE, **kw)
            ts.append(time.perf_counter() - t0)
    return ts


def cpu_baseline_sample(net=None, ctx=(("text", 77, 1.0),)):
    import torch
    from oracle import vd_oracle as O
    threads, model = host_info()
    torch.set_num_threads(threads)
    sd = _cpu_state_dict(net)
    ts = _cpu_time_step(sd, 2, ctx=ctx)
    with torch.no_grad():
        t0 = time.perf_counter()
        O.vae_decode(sd, torch.randn(1, 4, LAT, LAT))
        t_dec = time.perf_counter() - t0
    per_image = DDIM_STEPS * min(ts) + t_dec
    return {"value": round(1.0 / per_image, 6), "unit": UNIT, "cores": threads, "cpu": model, "kind": "port",
            "sample": f"oracle/vd_oracle.py (fp32 torch CPU port of lib/model_zoo, {threads} threads): 2 CFG UNet steps of one image "
                      f"(B=2, latent 64x64) at {min(ts):.2f} s/step + 1 VAE decode at {t_dec:.2f} s, "
                      f"extrapolated to 50 steps", "s_per_ddim_step": round(min(ts), 3), "s_vae_decode": round(t_dec, 3)}


def run_reference(args):
    """--impl reference: the reference's own CPU path for this config, timed on the host cores.
    /root/reference does not exist on the GPU box, so this is the oracle PORT (kind 'port').  Every timed step is a
    bounded sample (one CFG DDIM step of ONE image, B = 2 rows); in addition ONE CFG step at the product's own batch
    (bs 4 -> B = 8 rows) is timed so that the per-forward comparison is like for like (`same_config_per_forward`)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from oracle import vd_oracle as O
    threads, model = host_info()
    torch.set_num_threads(threads)
    cfg = CONFIGS[args.config]
    ctx = tuple(cfg["ctx"])
    sd = _cpu_state_dict(None)
    _cpu_time_step(sd, args.warmup if args.warmup > 0 else 1, ctx=ctx)
    ts = _cpu_time_step(sd, args.steps, ctx=ctx)
    t_b8 = _cpu_time_step(sd, 1, rows=BS, ctx=ctx)[0]
    with torch.no_grad():
        t0 = time.perf_counter()
        O.vae_decode(sd, torch.randn(1, 4, LAT, LAT))
        t_dec = time.perf_counter() - t0
    ts_sorted = sorted(ts)
    step_s = ts_sorted[len(ts_sorted) // 2]                      # median: robust against a noisy neighbour on the host
    per_image = DDIM_STEPS * step_s + t_dec
    value = 1.0 / per_image
    value_b8 = BS / (DDIM_STEPS * t_b8 + BS * t_dec)             # images/s from the B = 8 step (the product's own batch)
    sample = (f"each step = one CFG DDIM step of one 512x512 image (UNet B=2, latent 64x64) on CPU fp32 with {threads} threads "
              f"({model}), median {step_s:.2f} s (min {ts_sorted[0]:.2f}, max {ts_sorted[-1]:.2f}); images/s extrapolated as "
              f"1/(50*step + vae_decode {t_dec:.2f} s); one CFG step at the product batch (bs 4, B=8): {t_b8:.2f} s")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_s * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init weights, synthetic context)",
            "config": {"workload": cfg["workload"] + " — bounded CPU sample", "name": args.config, "global_batch": 1,
                       "same_config_per_forward": True, "s_per_cfg_step_bs4": round(t_b8, 3),
                       "images_per_s_from_bs4_step": round(value_b8, 6), "threads": threads, "cpu": model},
            "cpu_baseline": {"value": round(value, 6), "unit": UNIT, "cores": threads, "cpu": model, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configs[1..3]; the driver line is c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph-roofline", action="store_true", help="keep the per-launch eager event timing of the roofline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
