"""Pins the CPU oracle (oracle/vd_oracle.py) against the golden fixtures that oracle/make_golden.py produced by
running the UNMODIFIED reference, against the known-answer constants of SURVEY.md §8c, and — when
/root/reference is present — against the live reference.  fp32 vs fp32: tolerance is op-reordering round-off."""
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _close(a, b, rtol=2e-4, what=""):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item() + 1e-12
    assert err <= rtol * scale, f"{what}: {err:.3g} vs scale {scale:.3g}"


@pytest.fixture(scope="module")
def mini():
    from oracle import weights
    from oracle.make_golden import golden_inputs, WEIGHT_SEED
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(GOLD, "keys_mini.json"))).items()}
    sd = weights.synth_state_dict(shapes, seed=WEIGHT_SEED)
    return sd, golden_inputs("mini"), dict(np.load(os.path.join(GOLD, "mini.npz")))


def test_schedule_known_answers():
    from oracle import vd_oracle as O
    sched = json.load(open(os.path.join(GOLD, "schedule.json")))
    ddpm = O.ddpm_schedule()
    assert abs(float(ddpm["betas"][0]) - 0.00085) < 1e-9 and abs(float(ddpm["betas"][999]) - 0.012) < 1e-8
    assert abs(float(ddpm["alphas_cumprod"][0]) - 0.99915) < 1e-6
    assert abs(float(ddpm["alphas_cumprod"][999]) - 0.004660098513) < 1e-9
    np.testing.assert_allclose([float(ddpm["betas"][0]), float(ddpm["betas"][999])], sched["betas_0_999"], rtol=0, atol=0)
    # SURVEY §8c KATs (probed from the reference)
    kat = {50: dict(a=[0.998296022, 0.980380774, 0.00728172716, 0.00577550009], p=[0.999149978, 0.998296022, 0.00911730994, 0.00728172716],
                    s=[0.0412792638, 0.14006865, 0.996352494, 0.997108042]),
           10: dict(a=[0.998296022, 0.892980516, 0.0365465246, 0.0140048983], p=[0.999149978, 0.998296022, 0.0819167122, 0.0365465246],
                    s=[0.0412792638, 0.327138335, 0.981556654, 0.992972851])}
    for steps in (50, 10):
        o = O.ddim_schedule(ddpm["alphas_cumprod"], steps)
        g = sched[str(steps)]
        assert list(o["timesteps"]) == g["timesteps"]
        assert g["timesteps"][:3] == ([1, 21, 41] if steps == 50 else [1, 101, 201])
        for k in ("alphas", "alphas_prev", "sigmas", "sqrt_one_minus_alphas"):
            np.testing.assert_array_equal(np.asarray(o[k], dtype=np.float32), np.asarray(g[k], dtype=np.float32))
        idx = [0, 1, -2, -1]
        np.testing.assert_allclose(o["alphas"][idx], kat[steps]["a"], rtol=2e-7)
        np.testing.assert_allclose(o["alphas_prev"][idx], kat[steps]["p"], rtol=2e-7)
        np.testing.assert_allclose(o["sqrt_one_minus_alphas"][idx], kat[steps]["s"], rtol=2e-7)
        assert not np.any(o["sigmas"])


def test_unet_forward_vs_reference_golden(mini):
    from oracle import vd_oracle as O
    sd, gi, gold = mini
    with torch.no_grad():
        _close(O.apply_model(sd, gi["x"], gi["t"], [gi["c_text"]], model_channels=64), gold["eps_text"], what="text ctx")
        _close(O.apply_model(sd, gi["x"], gi["t"], [gi["c_img"]], c_types=("image",), model_channels=64), gold["eps_image"], what="image ctx")
        _close(O.apply_model(sd, gi["x"], gi["t"], [gi["c_text"], gi["c_img"]], ratios=[0.7, 0.3], c_types=("text", "image"),
                             model_channels=64), gold["eps_dual"], what="dual ctx mix")
        _close(O.timestep_embedding(torch.tensor([1, 21, 501, 981]), 320), gold["t_emb"], rtol=1e-6, what="t_emb")


def test_ddim_trajectory_vs_reference_golden(mini):
    from oracle import vd_oracle as O
    sd, gi, gold = mini
    with torch.no_grad():
        x, trace = O.ddim_sample(sd, gi["xT"], [gi["c"]], [gi["u"]], 5, 7.5, collect=True, model_channels=64)
    _close(x, gold["ddim5_final"], rtol=1e-3, what="final latent")
    _close(torch.stack([t["pred_x0"] for t in trace]), gold["ddim5_pred_x0"], rtol=1e-3, what="pred_x0 trace")


def test_vae_vs_reference_golden(mini):
    from oracle import vd_oracle as O
    sd, gi, gold = mini
    with torch.no_grad():
        _close(O.vae_decode(sd, gi["z"]), gold["vae_decode"], what="vae_decode")
        mean = O.vae_encode(sd, gi["img"], noise=None) / 0.18215
    _close(mean, gold["vae_moments"][:, :4], what="vae_encode mean")


def test_c1_full_size_vs_reference_golden():
    """BASELINE config 1 end to end on the oracle at full size (about a minute of CPU)."""
    from oracle import vd_oracle as O, weights
    from oracle.make_golden import golden_inputs, WEIGHT_SEED
    gold = dict(np.load(os.path.join(GOLD, "c1_full.npz")))
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(GOLD, "keys_full.json"))).items()}
    sd = weights.synth_state_dict(shapes, seed=WEIGHT_SEED)
    gi = golden_inputs("c1")
    with torch.no_grad():
        eps0 = O.apply_model(sd, torch.cat([gi["xT"]] * 2), torch.tensor([901, 901]), [torch.cat([gi["u"], gi["c"]])])
        _close(eps0, gold["eps0"], what="C1 eps0")
        img = O.vae_decode(sd, torch.as_tensor(gold["final"]))
    assert (img - torch.as_tensor(gold["image"].astype(np.float32))).abs().max().item() <= 2e-3   # fp16-stored fixture


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib/model_zoo"), reason="reference tree not on this machine")
def test_live_reference_matches_oracle_on_fresh_inputs():
    """Runs in a subprocess: the reference's package is also called `lib`."""
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from oracle import ref_shims, weights, vd_oracle as O
net = ref_shims.build_vd(unet_overrides=dict(model_channels=64), vae_overrides=dict(ch=64))
sd = weights.synth_state_dict(weights.param_shapes(net), seed=5)
net.load_state_dict(sd, strict=False)
g = torch.Generator().manual_seed(77)
x, t, c = torch.randn(2, 4, 24, 16, generator=g), torch.tensor([3, 777]), torch.randn(2, 33, 768, generator=g)
with torch.no_grad():
    a = net.apply_model({'type': 'image', 'x': x}, t, {'type': 'image', 'c': c})
    b = O.apply_model(sd, x, t, [c], c_types=('image',), model_channels=64)
    z = torch.randn(1, 4, 8, 12, generator=g)
    da, db = net.vae_decode(z, 'image'), O.vae_decode(sd, z)
assert (a - b).abs().max() <= 2e-4 * a.abs().max(), (a - b).abs().max()
assert (da - db).abs().max() <= 1e-4
# img2img start (ddim.py:97-103): the reference draws q_sample's noise with randn_like -> same seed on both sides
import lib.model_zoo.ddim as rd
S = rd.DDIMSampler(net)
x0 = torch.randn(1, 4, 16, 16, generator=g) * 0.8
cc, uu = torch.randn(1, 33, 768, generator=g) * 0.5, torch.randn(1, 33, 768, generator=g) * 0.5
torch.manual_seed(123)
with torch.no_grad():
    xa, inter = S.sample(steps=8, shape=[1, 4, 16, 16], x_info={'type': 'image', 'x0': x0, 'x0_forward_timesteps': 5},
                         c_info={'type': 'image', 'conditioning': cc, 'unconditional_conditioning': uu,
                                 'unconditional_guidance_scale': 7.5}, verbose=False, eta=0.)
    torch.manual_seed(123)
    nz = torch.randn_like(x0)
    xb = O.ddim_sample(sd, None, [cc], [uu], 8, 7.5, c_types=('image',), model_channels=64, x0=x0, x0_forward_timesteps=5,
                       x0_noise=nz)
assert (xa - xb).abs().max() <= 5e-4 * xa.abs().max(), (xa - xb).abs().max()
# dual-context sampler (BASELINE config 4's entry point, ddim.py:173-298): text 0.7 + image 0.3, x_T injected through randn
ct, ut = torch.randn(1, 20, 768, generator=g) * 0.5, torch.randn(1, 20, 768, generator=g) * 0.5
ci, ui = torch.randn(1, 33, 768, generator=g) * 0.5, torch.zeros(1, 33, 768)
xT = torch.randn(1, 4, 16, 16, generator=g)
orig = torch.randn
torch.randn = lambda *a, **k: xT.clone() if (len(a) > 0 and list(a[0]) == list(xT.shape)) else orig(*a, **k)
try:
    with torch.no_grad():
        xm, _ = S.sample_multicontext(steps=4, shape=[1, 4, 16, 16], x_info={'type': 'image'},
                                      c_info_list=[{'type': 'text', 'conditioning': ct, 'unconditional_conditioning': ut,
                                                    'unconditional_guidance_scale': 7.5, 'ratio': 0.7},
                                                   {'type': 'image', 'conditioning': ci, 'unconditional_conditioning': ui,
                                                    'unconditional_guidance_scale': 7.5, 'ratio': 0.3}], verbose=False, eta=0.)
finally:
    torch.randn = orig
with torch.no_grad():
    xo = O.ddim_sample(sd, xT, [ct, ci], [ut, ui], 4, 7.5, c_types=('text', 'image'), ratios=[0.7, 0.3], model_channels=64)
assert (xm - xo).abs().max() <= 5e-4 * xm.abs().max(), (xm - xo).abs().max()
# text-latent flows (SURVEY 8f rank 4): the 0-D diffuser with its data blocks, apply_model + the 4-step CFG DDIM walk on [n, 768]
net_t = ref_shims.build_vd(unet_overrides=dict(model_channels=64), with_vae=False, text_parts='dc')
sd_t = weights.synth_state_dict(weights.param_shapes(net_t), seed=6)
net_t.load_state_dict(sd_t, strict=False)
xt = torch.randn(2, 768, generator=g)
ci2, ui2 = torch.randn(2, 40, 768, generator=g) * 0.5, torch.zeros(2, 40, 768)
with torch.no_grad():
    ea = net_t.apply_model({'type': 'text', 'x': xt}, torch.tensor([5, 900]), {'type': 'image', 'c': ci2})
    eb = O.apply_model_text(sd_t, xt, torch.tensor([5, 900]), [ci2], c_types=('image',), model_channels=64)
assert (ea - eb).abs().max() <= 2e-4 * ea.abs().max(), (ea - eb).abs().max()
St = rd.DDIMSampler(net_t)
orig = torch.randn
torch.randn = lambda *a, **k: xt.clone() if (len(a) > 0 and list(a[0]) == [2, 768]) else orig(*a, **k)
try:
    with torch.no_grad():
        xs, _ = St.sample(steps=4, shape=[2, 768], x_info={'type': 'text'},
                          c_info={'type': 'image', 'conditioning': ci2, 'unconditional_conditioning': ui2,
                                  'unconditional_guidance_scale': 7.5}, verbose=False, eta=0.)
finally:
    torch.randn = orig
with torch.no_grad():
    xr = O.ddim_sample_text(sd_t, xt, [ci2], [ui2], 4, 7.5, c_types=('image',), model_channels=64)
assert (xs - xr).abs().max() <= 5e-4 * xs.abs().max(), (xs - xr).abs().max()
print('LIVE-OK')
""" % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stderr[-2000:]


def test_oracle_text_latent_diffuser_vs_reference_golden():
    """the 0-D diffuser restatement (Linear_MultiDim / FCBlock_MultiDim walk) against goldens produced by the
    unmodified reference's apply_model on a [B, 768] text latent (tests/golden/mini_text.npz)."""
    import json
    import numpy as np
    from oracle import vd_oracle as O, weights
    from oracle.make_golden import golden_inputs, WEIGHT_SEED
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(gold_dir, "keys_mini_text.json"))).items()}
    sd = weights.synth_state_dict(shapes, seed=WEIGHT_SEED)
    gold = dict(np.load(os.path.join(gold_dir, "mini_text.npz")))
    gt = golden_inputs("text")
    with torch.no_grad():
        t2t = O.apply_model_text(sd, gt["x"], gt["t"], [gt["c_text"]], c_types=("text",), model_channels=64)
        i2t = O.apply_model_text(sd, gt["x"], gt["t"], [gt["c_img"]], c_types=("image",), model_channels=64)
    for out, key in ((t2t, "eps_t2t"), (i2t, "eps_i2t")):
        ref = torch.as_tensor(gold[key])
        assert (out - ref).abs().max().item() <= 2e-4 * ref.abs().max().item(), key
