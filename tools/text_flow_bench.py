"""Text-latent diffusion (i2t: 768-d text latent, CLIP-image context) at FULL size on one GPU: time of one CFG UNet evaluation and of
a 50-step DDIM walk for bs 4 (B = 8 rows), against the weight-streaming bound (every FCBlock / Linear_MultiDim weight is read once
per evaluation: M = 8 rows cannot amortise it).  Random-init weights, synthetic context.
    VDB_TEXT_FLOWS=1 python tools/text_flow_bench.py"""
import json
import os
import sys
import time

os.environ["VDB_TEXT_FLOWS"] = "1"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "versatile-diffusion_b200"))
import torch  # noqa: E402
from lib.cfg_helper import model_cfg_bank  # noqa: E402
from lib.model_zoo import get_model  # noqa: E402
from lib.model_zoo.ddim import DDIMSampler  # noqa: E402

dev = torch.device("cuda", 0)
cfg = model_cfg_bank()('vd_four_flow_v1-0')
cfg.args.ctx_cfg_list = []
cfg.args.vae_cfg_list = []
torch.manual_seed(0)
t0 = time.time()
with torch.device(dev):
    net = get_model()(cfg, verbose=False)
g = torch.Generator(device=dev).manual_seed(1)
with torch.no_grad():
    for _, p in net.named_parameters():
        if p.ndim == 1 or not bool(p.any()):
            if p.ndim == 1 and p.shape[0] > 0 and bool((p == 1).all()):
                continue
            p.normal_(0.0, 0.02, generator=g)
net.eval()
net.to(dev)
text = net.diffuser["text"]
wbytes = sum(p.numel() for p in text.data_blocks.parameters()) * 2          # bf16 packed copies streamed per evaluation
wbytes += sum(p.numel() for p in net.diffuser["image"].context_blocks.parameters()) * 2
print(f"built in {time.time() - t0:.1f} s; data + context weights read per evaluation: {wbytes / 1e9:.2f} GB (bf16)")
bs = 4
gq = torch.Generator().manual_seed(3)
xT = torch.randn(bs, 768, generator=gq).to(dev)
c = (torch.randn(bs, 257, 768, generator=gq) * 0.5).to(dev)
u = torch.zeros(bs, 257, 768, device=dev)
x2, t2, c2 = torch.cat([xT] * 2), torch.full((2 * bs,), 981, device=dev), torch.cat([u, c])
with torch.no_grad():
    for _ in range(3):
        net.apply_model({"type": "text", "x": x2}, t2, {"type": "image", "c": c2})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = net.apply_model({"type": "text", "x": x2}, t2, {"type": "image", "c": c2})
    e1.record()
    torch.cuda.synchronize()
    ms_fwd = e0.elapsed_time(e1) / 10
    S = DDIMSampler(net)
    kw = dict(steps=50, shape=[bs, 768], x_info={"type": "text", "xt": xT},
              c_info={"type": "image", "conditioning": c, "unconditional_conditioning": u, "unconditional_guidance_scale": 7.5},
              verbose=False, eta=0.)
    for _ in range(2):
        S.sample(**kw)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        z, _ = S.sample(**kw)
    e1.record()
    torch.cuda.synchronize()
    ms_samp = e0.elapsed_time(e1) / 3
peak = 6571.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"workload": "i2t diffusion, text latent [4, 768], CFG 7.5, CLIP-image context, bf16, full-size 0-D diffuser",
                  "eager_forward_ms_B8": round(ms_fwd, 3), "forward_weight_gbs": round(wbytes / ms_fwd / 1e6, 1),
                  "ddim50_ms": round(ms_samp, 2), "ms_per_ddim_step": round(ms_samp / 50, 3),
                  "step_weight_gbs": round(wbytes / (ms_samp / 50) / 1e6, 1), "hbm_peak_gbs": peak,
                  "frac_of_hbm_peak_in_graph": round(wbytes / (ms_samp / 50) / 1e6 / peak, 3),
                  "latents_per_s": round(bs / (ms_samp / 1e3), 2), "finite": bool(torch.isfinite(z).all())}))
