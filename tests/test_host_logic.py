"""Host-side logic of the drop-in surface (CPU only): config bank, registry, checkpoint key ABI against the
reference-dumped fixtures, layer orders, DDIMSampler.make_schedule against reference tables, error behaviour."""
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def build(mini, device="cpu"):
    from lib.cfg_helper import model_cfg_bank
    from lib.model_zoo import get_model
    cfg = model_cfg_bank()('vd_four_flow_v1-0')
    cfg.args.ctx_cfg_list = []
    if mini:
        for _, d in cfg.args.diffuser_cfg_list:
            d.args.update(dict(model_channels=64))
        cfg.args.vae_cfg_list[0][1].args.ddconfig.update(dict(ch=64))
    with torch.device(device):
        return get_model()(cfg, verbose=False)


@pytest.mark.parametrize("mini,fixture", [(True, "keys_mini.json"), (False, "keys_full.json")])
def test_checkpoint_keys_and_shapes_match_reference(mini, fixture):
    ref = {k: tuple(v) for k, v in json.load(open(os.path.join(GOLD, fixture))).items()}
    net = build(mini, device="cpu" if mini else "meta")     # full size: shapes only
    ours = {k: tuple(v.shape) for k, v in net.named_parameters()}
    assert set(ours) == set(ref), sorted(set(ours) ^ set(ref))[:10]
    assert all(ours[k] == ref[k] for k in ref)
    # the 12 schedule buffers live at the top level like the reference's
    for b in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
              "posterior_variance", "posterior_mean_coef1", "posterior_mean_coef2"):
        assert b in dict(net.named_buffers())


def test_layer_orders_match_the_reference_walk():
    from oracle.vd_oracle import unet_layout
    net = build(True)
    d2, d0 = net.diffuser["image"], net.diffuser["text"]
    assert d2.layer_order == d0.layer_order and net.check_diffuser()
    assert d2.layer_order.count('d') == 30 and d2.layer_order.count('c') == 16
    assert d2.layer_order.count('save_hidden_feature') == 12 == d2.layer_order.count('load_hidden_feature')
    kinds = {"conv_in": "d", "res": "d", "down": "d", "up": "d", "out": "d", "ctx": "c", "save": "save_hidden_feature",
             "load": "load_hidden_feature"}
    assert [kinds[k] for k, _, _ in unet_layout(model_channels=64)] == d2.layer_order
    assert len(d2.data_blocks) == 30 and len(d2.context_blocks) == 16 and len(d0.context_blocks) == 16
    assert not hasattr(d0, "data_blocks") and set(d2.parameter_group) == {"global", "data", "context"}


def test_ddim_make_schedule_matches_reference_tables():
    from lib.model_zoo.ddim import DDIMSampler
    net = build(True)
    sched = json.load(open(os.path.join(GOLD, "schedule.json")))
    for steps in (50, 10):
        S = DDIMSampler(net)
        S.make_schedule(ddim_num_steps=steps, ddim_eta=0., verbose=False)
        g = sched[str(steps)]
        assert [int(v) for v in S.ddim_timesteps] == g["timesteps"]
        np.testing.assert_array_equal(np.asarray(S.ddim_alphas, dtype=np.float32), np.asarray(g["alphas"], dtype=np.float32))
        np.testing.assert_array_equal(np.asarray(S.ddim_alphas_prev, dtype=np.float32), np.asarray(g["alphas_prev"], dtype=np.float32))
        np.testing.assert_array_equal(np.asarray(S.ddim_sqrt_one_minus_alphas, dtype=np.float32),
                                      np.asarray(g["sqrt_one_minus_alphas"], dtype=np.float32))
        assert not np.any(np.asarray(S.ddim_sigmas))
    assert S.ddpm_num_timesteps == 1000


def test_config_bank_and_registry_surface():
    from lib.cfg_helper import model_cfg_bank
    from lib.model_zoo import get_model
    bank = model_cfg_bank()
    u = bank("openai_unet_2d_v1")
    assert u.type == "openai_unet_2d_next" and u.args.model_channels == 320 and u.args.channel_mult == [1, 2, 4, 4]
    assert bank("openai_unet_0d_v1_c").args.parts == ["context"]
    assert bank("autokl_v1").args.ddconfig.ch_mult == [1, 2, 4, 4]
    vd = bank("vd_four_flow_v1-0")
    assert vd.args.latent_scale_factor["image"] == 0.18215 and vd.args.global_layer_ptr == "image"
    with pytest.raises(KeyError):
        bank("optimus_v1")                               # text-latent flows are out of the hot-path build
    assert get_model() is get_model()                    # singleton like the reference
    # the 0-D diffuser's data blocks (text-latent flows, round 2) build from the reference's config keys; every FCBlock gets a
    # column slot of the diffuser-level fused embedding projection
    cfg0 = bank("openai_unet_0d_v1")
    assert cfg0.args.parts == ["global", "data", "context"] or "data" in cfg0.args.parts
    cfg0.args.model_channels, cfg0.args.input_channels, cfg0.args.output_channels = 32, 24, 24
    cfg0.args.context_dim, cfg0.args.num_heads = 16, 2
    u0 = get_model()(cfg0)
    fcs = u0._fc_blocks()
    assert len(fcs) > 0 and fcs[0].emb_slot == (0, None) and u0._emb_total == sum(f.out_channels for f in fcs)
    assert [f.emb_slot[0] for f in fcs] == sorted(f.emb_slot[0] for f in fcs)


def test_to_returns_none_and_no_cpu_path():
    from lib.model_zoo.ddim import DDIMSampler
    net = build(True)
    assert net.to("cpu") is None and net.device == "cpu"             # reference semantics (vd.py:114-116)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.apply_model({"type": "image", "x": torch.zeros(1, 4, 8, 8)}, torch.zeros(1, dtype=torch.long),
                        {"type": "text", "c": torch.zeros(1, 77, 768)})
    with pytest.raises(RuntimeError, match="no CPU path"):
        DDIMSampler(net).sample(steps=2, shape=[1, 4, 8, 8], x_info={"type": "image"},
                                c_info={"type": "text", "conditioning": torch.zeros(1, 77, 768),
                                        "unconditional_conditioning": torch.zeros(1, 77, 768),
                                        "unconditional_guidance_scale": 7.5}, verbose=False)


def test_every_entry_point_refuses_cpu_tensors():
    """No CPU fallback anywhere in the product: q_sample, timestep_embedding, the VAE and the raw ops all raise."""
    from lib.model_zoo.diffusion_utils import timestep_embedding
    from vdb200 import ops
    net = build(True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.q_sample(torch.zeros(1, 4, 8, 8), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        timestep_embedding(torch.tensor([1, 21]), 320)
    with pytest.raises((RuntimeError, ValueError)):
        net.vae_decode(torch.zeros(1, 4, 8, 8), "image")
    with pytest.raises((RuntimeError, ValueError)):
        net.vae_encode(torch.zeros(1, 3, 64, 64), "image")
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.layernorm(torch.zeros(4, 320, dtype=torch.bfloat16), torch.ones(320), torch.zeros(320))
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))


def test_missing_library_fails_at_import(tmp_path):
    """The binding must fail loudly when libvdb200.so is absent (no silent library / eager fallback)."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "versatile-diffusion_b200")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "try:\n    import vdb200\n    print('IMPORTED')\n"
            "except (ImportError, OSError) as e:\n    print('REFUSED', type(e).__name__)\n") % pkg
    env = dict(os.environ, VDB200_LIB=str(tmp_path / "nope" / "libvdb200.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "REFUSED" in out.stdout and "IMPORTED" not in out.stdout, out.stdout + out.stderr[-500:]


def test_packed_weights_invalidate_on_load_and_cast():
    net = build(True)
    rb = net.diffuser["image"].data_blocks[1][0]
    rb._packed = {"stale": True}
    net.load_state_dict(net.state_dict())
    assert rb._packed is None
    rb._packed = {"stale": True}
    net.half()
    assert rb._packed is None and rb.in_layers[2].weight.dtype == torch.float16


def test_layernorm_fold_algebra_matches_linear_of_layernorm():
    """What vdb_gemm_ln_bf16's consumer epilogue evaluates — r * (x W'^T - mu * s) + c with the weights packed by fold_layernorm and
    (mu, r) rebuilt from per-chunk partial (sum, sum of squares) — is Linear(LayerNorm(x)) (attention.py:206-218); checked in
    fp64 on the CPU with the bf16-rounded folded weights the kernel would multiply with."""
    import torch
    import torch.nn.functional as F
    from lib.model_zoo.attention import fold_layernorm
    g = torch.Generator().manual_seed(3)
    M, C, N, eps = 37, 320, 96, 1e-5
    x = (torch.randn(M, C, generator=g) * 1.7 + 0.9).to(torch.bfloat16).double()         # rows with a mean, as the residual stream has
    w, b = torch.randn(N, C, generator=g) * C ** -0.5, torch.randn(N, generator=g)
    gamma, beta = 1.0 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    wg, s, c = fold_layernorm(w, b, gamma, beta)
    assert wg.dtype == torch.bfloat16 and torch.equal(s, wg.float().sum(1))
    # statistics exactly as a producer GEMM writes them: partial sums over column ranges (here 4 partials of 80 columns)
    parts = torch.stack([torch.stack([xc.sum(1), (xc * xc).sum(1)], -1) for xc in x.split(80, dim=1)])     # [4, M, 2]
    su, sq = parts[..., 0].sum(0), parts[..., 1].sum(0)
    mu = su / C
    r = torch.rsqrt((sq / C - mu * mu).clamp_min(0) + eps)
    out = r[:, None] * (x @ wg.double().t() - mu[:, None] * s.double()[None, :]) + c.double()[None, :]
    # reference with the SAME rounded weights: LayerNorm(x) * gamma folded == (x_hat * gamma) W^T with W' = bf16(W * gamma)
    xhat = (x - x.mean(1, keepdim=True)) * torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + eps)
    ref = xhat @ wg.double().t() + (w.double() @ beta.double() + b.double())[None, :]
    assert (out - ref).abs().max().item() < 1e-5          # (s and c are fp32 tables)
    # and against torch's own LayerNorm + Linear with unrounded weights: only the bf16 rounding of W * gamma separates them
    full = F.linear(F.layer_norm(x, (C,), gamma.double(), beta.double(), eps), w.double(), b.double())
    assert (out - full).abs().max().item() < 5e-2 and F.cosine_similarity(out.flatten(), full.flatten(), dim=0).item() > 0.99999
