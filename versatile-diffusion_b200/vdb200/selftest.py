"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle.

Builds a reduced-width VD (model_channels 64) with seeded synthetic weights, runs one CFG DDIM step
(UNet forward over the doubled batch through every kernel family: tcgen05 GEMM / implicit conv / flash
attention, GroupNorm, LayerNorm, K4) plus a VAE decode, and compares with oracle/vd_oracle.py."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def smoke():
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device: the vdb200 path has no CPU fallback")
    for p in (ROOT, os.path.join(ROOT, "versatile-diffusion_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from lib.cfg_helper import model_cfg_bank
    from lib.model_zoo import get_model
    from lib.model_zoo.ddim import DDIMSampler
    from oracle import vd_oracle as O, weights
    from vdb200 import ops

    torch.cuda.set_device(0)
    cfg = model_cfg_bank()('vd_four_flow_v1-0')
    cfg.args.ctx_cfg_list = []
    for _, d in cfg.args.diffuser_cfg_list:
        d.args.update(dict(model_channels=64))
    cfg.args.vae_cfg_list[0][1].args.ddconfig.update(dict(ch=64))
    net = get_model()(cfg, verbose=False)
    sd = weights.synth_state_dict(weights.param_shapes(net), seed=1)
    net.load_state_dict(sd, strict=False)
    net.eval()
    net.to("cuda:0")

    g = torch.Generator().manual_seed(0)
    xT = torch.randn(1, 4, 32, 32, generator=g)
    c, u = torch.randn(1, 77, 768, generator=g) * 0.5, torch.randn(1, 77, 768, generator=g) * 0.5
    ops.reset_launch_count()
    with torch.no_grad():
        x, inter = DDIMSampler(net).sample(
            steps=2, shape=[1, 4, 32, 32], x_info={"type": "image", "xt": xT},
            c_info={"type": "text", "conditioning": c.cuda(), "unconditional_conditioning": u.cuda(),
                    "unconditional_guidance_scale": 7.5}, verbose=False, eta=0.)
        img = net.vae_decode(x, "image")
        ref = O.ddim_sample(sd, xT, [c], [u], 2, 7.5, model_channels=64)
        ref_img = O.vae_decode(sd, ref)
    torch.cuda.synchronize()
    cos = F.cosine_similarity(x.float().cpu().flatten(), ref.flatten(), dim=0).item()
    err = (img.float().cpu() - ref_img).abs().max().item()
    n = ops.launch_count()
    print(f"[smoke] 2-step DDIM latent cosine vs oracle {cos:.6f}; decoded image max|err| {err:.4f}; "
          f"{n} vdb200 kernel launches")
    # bf16 vs fp32, 2 CFG DDIM steps from pure noise (CFG 7.5 amplifies the per-forward bf16 error ~7.5x): latent cosine >= 0.999
    # (observed 0.99905), decoded [0,1] image within 0.1 (observed 0.088)
    if not (cos >= 0.999 and err <= 0.1 and n > 0):
        raise AssertionError("smoke: CUDA path deviates from the CPU oracle")


if __name__ == "__main__":
    smoke()
