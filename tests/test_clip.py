"""CLIP context encoders: oracle pinned against transformers.CLIPModel (CPU), host preprocessing pinned against
CLIPImageProcessor (CPU), and the CUDA path against the oracle (gpu)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tokens(n=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    tok = torch.full((n, 77), 49407, dtype=torch.long)
    tok[:, 0] = 49406
    for i in range(n):
        k = 5 + 7 * i
        tok[i, 1:1 + k] = torch.randint(1000, 40000, (k,), generator=g)
    return tok


def _synth(model, seed):
    from oracle import weights
    shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
    sd = weights.synth_state_dict(shapes, seed=seed)
    model.load_state_dict(sd, strict=False)
    return sd


def test_oracle_clip_matches_hf_clipmodel():
    """Third-party arithmetic pin: oracle restatement == transformers.CLIPModel (this image: 5.5.0) on random weights."""
    from transformers import CLIPModel
    from lib.model_zoo.clip import vit_l14_config
    from oracle import vd_oracle as O
    torch.manual_seed(0)
    m = CLIPModel(vit_l14_config()).eval()
    sd = _synth(m, 3)
    tok = _tokens()
    g = torch.Generator().manual_seed(1)
    px = torch.randn(1, 3, 224, 224, generator=g)
    with torch.no_grad():
        out = m.text_model(input_ids=tok)
        z = m.text_projection(out.last_hidden_state)
        ref_t = z / torch.norm(m.text_projection(out.pooler_output).unsqueeze(1), dim=-1, keepdim=True)
        mine_t = O.clip_text_encode({"ctx.text.model." + k: v for k, v in sd.items()}, tok)
        o = m.vision_model(pixel_values=px)
        zi = m.visual_projection(m.vision_model.post_layernorm(o.last_hidden_state))
        ref_i = zi / torch.norm(zi[:, 0:1], dim=-1, keepdim=True)
        mine_i = O.clip_image_encode({"ctx.image.model." + k: v for k, v in sd.items()}, px)
    assert (ref_t - mine_t).abs().max().item() <= 2e-6
    assert (ref_i - mine_i).abs().max().item() <= 2e-6


def test_preprocess_matches_hf_image_processor():
    from transformers import CLIPImageProcessor
    from PIL import Image
    from lib.model_zoo.clip import CLIPImageContextEncoder
    g = np.random.RandomState(0)
    imgs = [Image.fromarray(g.randint(0, 255, (300, 420, 3), dtype=np.uint8)), Image.fromarray(g.randint(0, 255, (512, 512, 3), dtype=np.uint8))]
    ref = CLIPImageProcessor()(images=imgs, return_tensors="pt")["pixel_values"]
    mine = CLIPImageContextEncoder.preprocess(imgs)
    assert mine.shape == ref.shape == (2, 3, 224, 224)
    # transformers 5.5 resizes with torchvision, the reference's pinned 4.24 (and this code) with PIL: results differ by
    # at most ONE 8-bit level (1/255/std = 0.0146..0.0150) on a few pixels of a noise image, and not at all on smooth ones
    d = (mine - ref).abs()
    assert d.max().item() <= 0.0151 and d.mean().item() <= 2e-4
    t = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(0))
    import torchvision.transforms as tvtrans
    ref_t = CLIPImageProcessor()(images=[tvtrans.ToPILImage()(t[0])], return_tensors="pt")["pixel_values"]
    d = (CLIPImageContextEncoder.preprocess(t) - ref_t).abs()
    assert d.max().item() <= 0.0151 and d.mean().item() <= 2e-4
    x = np.linspace(0, 1, 400)[None, :, None] * np.ones((300, 1, 3))
    smooth = [Image.fromarray((x * 255).astype(np.uint8))]
    assert (CLIPImageContextEncoder.preprocess(smooth) - CLIPImageProcessor()(images=smooth, return_tensors="pt")["pixel_values"]).abs().max().item() <= 1e-6


@pytest.mark.gpu
def test_clip_text_and_image_encoders_vs_oracle():
    from lib.cfg_helper import model_cfg_bank
    from lib.model_zoo import get_model
    from oracle import vd_oracle as O
    torch.manual_seed(0)
    tenc = get_model()(model_cfg_bank()("clip_text_context_encoder"), verbose=False)
    sd = _synth(tenc.model, 3)
    tenc.to("cuda")
    tok = _tokens(3, seed=5)
    with torch.no_grad():
        out = tenc.encode_tokens(tok)
        ref = O.clip_text_encode({"ctx.text.model." + k: v for k, v in sd.items()}, tok)
    assert out.shape == (3, 77, 768)
    cos = F.cosine_similarity(out.float().cpu().flatten(), ref.flatten(), dim=0).item()
    err = (out.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    print(f"[parity] CLIP text encode: cos {cos:.6f} rel max err {err:.4f}")
    assert cos >= 0.999 and err <= 0.05
    del tenc

    ienc = get_model()(model_cfg_bank()("clip_image_context_encoder"), verbose=False)
    sd = _synth(ienc.model, 3)
    ienc.to("cuda")
    g = torch.Generator().manual_seed(2)
    imgs = torch.rand(2, 3, 300, 360, generator=g)
    px = ienc.preprocess(imgs)
    with torch.no_grad():
        out = ienc.encode(imgs.cuda())
        ref = O.clip_image_encode({"ctx.image.model." + k: v for k, v in sd.items()}, px)
    assert out.shape == (2, 257, 768)
    cos = F.cosine_similarity(out.float().cpu().flatten(), ref.flatten(), dim=0).item()
    err = (out.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    print(f"[parity] CLIP image encode: cos {cos:.6f} rel max err {err:.4f}")
    assert cos >= 0.999 and err <= 0.05
    # masked variant: all-ones mask == unmasked (clip.py:110-111); half mask scales tokens
    with torch.no_grad():
        same = ienc.encode(imgs.cuda(), masks=torch.ones(2, 1, 64, 64))
        half = ienc.encode(imgs.cuda(), masks=torch.cat([torch.ones(2, 1, 64, 32), torch.zeros(2, 1, 64, 32)], -1))
    assert torch.equal(same, out)
    assert torch.isfinite(half).all() and half[:, 1:].abs().sum() < out[:, 1:].abs().sum()
    # masked variant against the oracle restatement of clip.py:103-143 (soft-edged, non-patch-aligned mask)
    g = torch.Generator().manual_seed(6)
    m = torch.zeros(2, 1, 96, 80)
    m[0, :, 10:70, 5:60] = 1.0
    m[1] = torch.rand(1, 96, 80, generator=g)
    with torch.no_grad():
        outm = ienc.encode(imgs.cuda(), masks=m)
        refm = O.clip_image_encode_wmask({"ctx.image.model." + k: v for k, v in sd.items()}, px, m)
    cos = F.cosine_similarity(outm.float().cpu().flatten(), refm.flatten(), dim=0).item()
    err = (outm.float().cpu() - refm).abs().max().item() / refm.abs().max().item()
    print(f"[parity] CLIP masked image encode: cos {cos:.6f} rel max err {err:.4f}")
    assert cos >= 0.999 and err <= 0.05


def test_oracle_masked_clip_reduces_to_unmasked_and_scales_tokens():
    """CPU: the masked restatement (clip.py:103-143) equals the unmasked encode for an all-ones mask, scales every output token
    by its mask factor, and matches transformers' own vision tower when the factors are applied to its embeddings by hand."""
    from transformers import CLIPModel
    from lib.model_zoo.clip import vit_l14_config
    from oracle import vd_oracle as O
    cfg = vit_l14_config()
    cfg.vision_config.num_hidden_layers = 2
    cfg.text_config.num_hidden_layers = 1
    torch.manual_seed(0)
    m = CLIPModel(cfg).eval()
    sd = {"ctx.image.model." + k: v for k, v in _synth(m, 4).items()}
    px = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    kw = dict(layers=2)
    ones = O.clip_image_encode_wmask(sd, px, torch.ones(2, 1, 50, 70), **kw)
    assert torch.equal(ones, O.clip_image_encode(sd, px, **kw))
    mask = torch.ones(2, 1, 224, 224)
    mask[:, :, :14, :14] = 0.0                       # patch token 1 (first patch) fully masked
    mask[1, :, 100:, :] = 0.3
    out = O.clip_image_encode_wmask(sd, px, mask, **kw)
    assert out[:, 1].abs().max().item() == 0.0        # a zero factor zeroes the token on the way out (clip.py:141)
    # independent check through HF's modules: scale the embeddings, run the encoder, post_layernorm + projection + norm + scale
    with torch.no_grad():
        emb = m.vision_model.embeddings(px)
        gs = mask.mean(dim=[1, 2, 3], keepdim=True).flatten(2)
        vt = F.avg_pool2d(mask, 14, stride=14).flatten(2).transpose(1, 2)
        tm = torch.cat([gs, vt], 1)
        h = m.vision_model.pre_layrnorm(emb * tm)
        h = m.vision_model.encoder(inputs_embeds=h).last_hidden_state
        z = m.visual_projection(m.vision_model.post_layernorm(h))
        ref = z / torch.norm(z[:, 0:1], dim=-1, keepdim=True) * tm
    assert (ref - out).abs().max().item() <= 5e-6
