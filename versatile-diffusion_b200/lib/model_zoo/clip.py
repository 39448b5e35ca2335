"""Frozen CLIP ViT-L/14 context encoders on vdb200 kernels — reference lib/model_zoo/clip.py:30-149.

Same registry names, constructor arguments and `encode` semantics:
  CLIPTextContextEncoder.encode(text)   -> [n, 77, 768]   text_projection on ALL tokens / ||proj(pooled)||   (:53-62)
  CLIPImageContextEncoder.encode(images, masks=None) -> [n, 257, 768]
        post_layernorm on ALL tokens -> visual_projection -> / ||token 0||   (:88-101); masked variant (:103-143)
`self.model` is a `transformers.CLIPModel` used ONLY as the parameter container (so checkpoints keep the
`ctx.{text,image}.model.*` keys); its forward is never called — the encoder layers run on the vdb200 GEMM,
LayerNorm and flash-attention kernels (pre-LN blocks, quick_gelu, causal mask for text).
Token streams are kept as bf16 [n, Lp, C] with Lp = 80 / 264 (77 / 257 rounded up to 8, TMA alignment).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from lib.model_zoo.common.get_model import register
from .diffusion_utils import PackedModule, bf16, f32, require_cuda

symbol = 'clip'
VERSION = "openai/clip-vit-large-patch14"
IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)     # CLIPImageProcessor defaults for this checkpoint
IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)


def _ops():
    from vdb200 import ops
    return ops


def vit_l14_config():
    from transformers import CLIPConfig
    return CLIPConfig(
        text_config=dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                         max_position_embeddings=77, vocab_size=49408, hidden_act='quick_gelu', projection_dim=768),
        vision_config=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                           image_size=224, patch_size=14, hidden_act='quick_gelu', projection_dim=768),
        projection_dim=768)


def _load_clip_model(version):
    """from_pretrained when the files are cached locally, else the same architecture with random init
    (weights then come from the VD checkpoint's ctx.* keys)."""
    from transformers import CLIPModel
    try:
        m = CLIPModel.from_pretrained(version, local_files_only=True)
        m._vdb_random_init = False
        return m
    except Exception as ex:   # no local files: the reference would fail here (it always downloads the pretrained CLIP)
        import warnings
        warnings.warn(f"CLIP '{version}' is not cached locally ({type(ex).__name__}): the encoder starts from RANDOM weights and "
                      f"is only valid once a checkpoint with ctx.* keys has been loaded (encode() raises otherwise)", RuntimeWarning)
        m = CLIPModel(vit_l14_config())
        m._vdb_random_init = True
        m.register_load_state_dict_post_hook(_clip_model_loaded_hook)
        return m


def _clip_model_loaded_hook(module, incompatible_keys):
    """the same for weights loaded straight into encoder.model (tests, tools)"""
    if not incompatible_keys.missing_keys:
        module._vdb_random_init = False


def _clip_weights_loaded_hook(module, incompatible_keys):
    """load_state_dict post hook of the context encoders: the random-init fallback is only cleared when the checkpoint really
    carried this encoder's CLIP weights (no missing model.* key)."""
    missing = [k for k in incompatible_keys.missing_keys if k.startswith("model.") or ".model." in k]
    if not missing:
        module.model._vdb_random_init = False


def _require_clip_weights(enc):
    if getattr(enc.model, "_vdb_random_init", False) and not getattr(enc, "allow_random_init", False):
        raise RuntimeError("CLIP context encoder still has its random-init fallback weights: load a checkpoint with the ctx.* keys "
                           "(load_state_dict) or set encoder.allow_random_init = True for synthetic-weight tests")


class AbstractEncoder(PackedModule):
    def encode(self, *args, **kwargs):
        raise NotImplementedError


def disabled_train(self, mode=True):
    return self


class _ClipTower(object):
    """Kernel-side packing + forward of one HF CLIP transformer tower (text_model / vision_model)."""

    @staticmethod
    def pack_layers(layers):
        out = []
        for lyr in layers:
            a = lyr.self_attn
            wo, bo = a.out_proj.weight.detach().float(), a.out_proj.bias.detach().float()
            out.append({
                "ln1": (f32(lyr.layer_norm1.weight), f32(lyr.layer_norm1.bias)),
                "ln2": (f32(lyr.layer_norm2.weight), f32(lyr.layer_norm2.bias)),
                "wqk": bf16(torch.cat([a.q_proj.weight.detach(), a.k_proj.weight.detach()], 0)),
                "bqk": f32(torch.cat([a.q_proj.bias.detach(), a.k_proj.bias.detach()], 0)),
                "wv": bf16(a.v_proj.weight),
                # softmax rows sum to 1, so P(V + 1 b_v^T) = PV + b_v: the V bias moves through out_proj
                "wo": bf16(wo), "bo": (bo + wo @ a.v_proj.bias.detach().float()).contiguous(),
                "w1": bf16(lyr.mlp.fc1.weight), "b1": f32(lyr.mlp.fc1.bias),
                "w2": bf16(lyr.mlp.fc2.weight), "b2": f32(lyr.mlp.fc2.bias),
                "eps": lyr.layer_norm1.eps})
        return out

    @staticmethod
    def run_layers(x, layers, B, L, Lp, heads, causal):
        """x: bf16 [B*Lp, C] (pad rows finite). Pre-LN attention + quick_gelu MLP per layer."""
        ops = _ops()
        C = x.shape[1]
        d = C // heads
        o = torch.zeros(B * Lp, C, dtype=torch.bfloat16, device=x.device)   # pad rows must stay finite
        for p in layers:
            h = ops.layernorm(x, *p["ln1"], eps=p["eps"])
            qk = ops.gemm(h, p["wqk"], bias=p["bqk"])                       # [B*Lp, 2C]: q | k
            vt = ops.gemm(p["wv"], h)                                       # [C, B*Lp] = V^T (bias folded into bo)
            ops.attention(qk, qk, vt, o, B, heads, L, L, d, scale=d ** -0.5, q_col0=0, k_col0=C, causal=causal,
                          q_bstride=Lp, kv_bstride=Lp)
            x = ops.gemm(o, p["wo"], bias=p["bo"], resid=x)
            h = ops.layernorm(x, *p["ln2"], eps=p["eps"])
            h = ops.gemm(h, p["w1"], bias=p["b1"], act=ops.ACT_QUICK_GELU)
            x = ops.gemm(h, p["w2"], bias=p["b2"], resid=x)
        return x


@register('clip_text_context_encoder')
class CLIPTextContextEncoder(AbstractEncoder):
    def __init__(self, version=VERSION, max_length=77, fp16=False):
        super().__init__()
        self.version = version
        self.tokenizer = None            # loaded lazily: needs the vocabulary files (not shipped offline)
        self.model = _load_clip_model(version)
        self.max_length = max_length
        self.fp16 = fp16
        self.register_load_state_dict_post_hook(_clip_weights_loaded_hook)
        self.freeze()

    def get_device(self):
        return self.model.text_projection.weight.device

    def freeze(self):
        self.model = self.model.eval()
        self.train = disabled_train
        for param in self.parameters():
            param.requires_grad = False

    def _pack(self):
        tm = self.model.text_model
        return {"tok": f32(tm.embeddings.token_embedding.weight), "pos": f32(tm.embeddings.position_embedding.weight),
                "layers": _ClipTower.pack_layers(tm.encoder.layers),
                "lnf": (f32(tm.final_layer_norm.weight), f32(tm.final_layer_norm.bias)), "eps": tm.final_layer_norm.eps,
                "proj": bf16(self.model.text_projection.weight),
                "heads": tm.encoder.layers[0].self_attn.num_heads}

    def tokenize(self, text):
        if self.tokenizer is None:
            from transformers import CLIPTokenizer
            try:
                self.tokenizer = CLIPTokenizer.from_pretrained(self.version, local_files_only=True)
            except Exception as e:
                raise RuntimeError(f"CLIP tokenizer files for '{self.version}' are not available locally; "
                                   "call encode_tokens(token_ids) with pre-tokenised input") from e
        be = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                            return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return be["input_ids"]

    @torch.no_grad()
    def encode_tokens(self, tokens):
        """tokens: int64 [n, 77] (BOS 49406, EOS/pad 49407) -> fp32 [n, 77, 768]."""
        ops = _ops()
        dev = self.get_device()
        tokens = tokens.to(dev).long().contiguous()
        require_cuda(tokens, "CLIPTextContextEncoder")
        p = self.packed()
        B, L = tokens.shape
        Lp = (L + 7) // 8 * 8
        x = ops.clip_text_embed(tokens, p["tok"], p["pos"], Lp).view(B * Lp, -1)
        x = _ClipTower.run_layers(x, p["layers"], B, L, Lp, p["heads"], causal=True)
        x = ops.layernorm(x, *p["lnf"], eps=p["eps"])
        z = ops.gemm(x, p["proj"])                                  # text_projection on every token
        eos = tokens.argmax(dim=-1).to(torch.int32).contiguous()    # pooled token = EOS position (highest id)
        return ops.scale_by_row_norm(z.view(B, Lp, -1), L, idx=eos)

    def encode(self, text):
        _require_clip_weights(self)
        z = self.encode_tokens(self.tokenize(text))
        return z.half() if self.fp16 else z


@register('clip_image_context_encoder')
class CLIPImageContextEncoder(AbstractEncoder):
    def __init__(self, version=VERSION, fp16=False):
        super().__init__()
        self.version = version
        self.model = _load_clip_model(version)
        self.fp16 = fp16
        self.register_load_state_dict_post_hook(_clip_weights_loaded_hook)
        self.freeze()

    def get_device(self):
        return self.model.text_projection.weight.device

    def freeze(self):
        self.model = self.model.eval()
        self.train = disabled_train
        for param in self.parameters():
            param.requires_grad = False

    def _pack(self):
        vm = self.model.vision_model
        w = vm.embeddings.patch_embedding.weight.detach()            # [1024, 3, 14, 14], no bias
        k = w[0].numel()
        kpad = (k + 63) // 64 * 64
        wp = torch.zeros(w.shape[0], kpad, dtype=torch.bfloat16, device=w.device)
        wp[:, :k] = w.reshape(w.shape[0], -1).to(torch.bfloat16)
        return {"wpatch": wp, "kpad": kpad, "patch": w.shape[-1], "cls": f32(vm.embeddings.class_embedding),
                "pos": f32(vm.embeddings.position_embedding.weight),
                "pre": (f32(vm.pre_layrnorm.weight), f32(vm.pre_layrnorm.bias)),
                "post": (f32(vm.post_layernorm.weight), f32(vm.post_layernorm.bias)), "eps": vm.post_layernorm.eps,
                "layers": _ClipTower.pack_layers(vm.encoder.layers), "proj": bf16(self.model.visual_projection.weight),
                "heads": vm.encoder.layers[0].self_attn.num_heads}

    @staticmethod
    def preprocess(images, size=224):
        """CLIPProcessor equivalent (clip.py:89-93): tensors in [0,1] go through PIL (uint8) like the reference,
        resize shortest side to 224 (bicubic), centre crop, rescale, normalise. Returns fp32 [n,3,224,224] (CPU)."""
        from PIL import Image
        if isinstance(images, torch.Tensor):
            arr = (images.detach().float().cpu().clamp(0, 1) * 255).byte().permute(0, 2, 3, 1).numpy()   # ToPILImage: mul(255).byte()
            images = [Image.fromarray(a) for a in arr]
        out = []
        for im in images:
            im = im.convert("RGB")
            w, h = im.size
            # transformers 4.24 get_resize_output_image_size: new_long = int(size * long / short) (truncation)
            nw, nh = (size, int(size * h / w)) if w <= h else (int(size * w / h), size)
            im = im.resize((nw, nh), resample=Image.BICUBIC)
            left, top = (nw - size) // 2, (nh - size) // 2
            im = im.crop((left, top, left + size, top + size))
            a = np.asarray(im, dtype=np.float32) / 255.0
            a = (a - np.asarray(IMAGE_MEAN, dtype=np.float32)) / np.asarray(IMAGE_STD, dtype=np.float32)
            out.append(torch.from_numpy(a).permute(2, 0, 1))
        return torch.stack(out).contiguous()

    @staticmethod
    def pil_bicubic_coeffs(in_size, out_size):
        """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BICUBIC filter (Resample.c): per output coordinate the
        first tap, the tap count and the int32 weights with 22 fractional bits, computed in double precision in Pillow's
        operation order.  -> (bounds int32 [out, 2], kk int32 [out, ksize])."""
        import math
        scale = filterscale = in_size / out_size
        if filterscale < 1.0:
            filterscale = 1.0
        support = 2.0 * filterscale
        ksize = int(math.ceil(support)) * 2 + 1
        inv = 1.0 / filterscale
        bounds = np.zeros((out_size, 2), dtype=np.int32)
        kk = np.zeros((out_size, ksize), dtype=np.int32)
        for xx in range(out_size):
            center = (xx + 0.5) * scale
            xmin = max(int(center - support + 0.5), 0)
            cnt = min(int(center + support + 0.5), in_size) - xmin
            ws, total = [], 0.0
            for x in range(cnt):
                t = abs((x + xmin - center + 0.5) * inv)
                if t < 1.0:
                    w = ((-0.5 + 2.0) * t - (-0.5 + 3.0)) * t * t + 1
                elif t < 2.0:
                    w = (((t - 5) * t + 8) * t - 4) * -0.5
                else:
                    w = 0.0
                ws.append(w)
                total += w
            for x, w in enumerate(ws):
                if total != 0.0:
                    w = w / total
                kk[xx, x] = int(-0.5 + w * (1 << 22)) if w < 0 else int(0.5 + w * (1 << 22))
            bounds[xx] = (xmin, cnt)
        return bounds, kk

    _pre_tables = {}

    @classmethod
    def preprocess_device(cls, images, size=224):
        """The preprocessing of `preprocess` without leaving the GPU (opt-in: VDB_CLIP_PRE_DEVICE=1 and a CUDA tensor input):
        same integers as ToPILImage + Pillow (vdb200 kernels, coefficient tables cached per input size)."""
        ops = _ops()
        require_cuda(images, "CLIPImageContextEncoder.preprocess_device")
        n, _, h, w = images.shape
        key = (h, w, size, str(images.device))
        tab = cls._pre_tables.get(key)
        if tab is None:
            nw, nh = (size, int(size * h / w)) if w <= h else (int(size * w / h), size)
            dev = images.device
            to = lambda pair: tuple(torch.from_numpy(a).to(dev).contiguous() for a in pair)
            tab = {"nw": nw, "nh": nh, "h": to(cls.pil_bicubic_coeffs(w, nw)) if nw != w else None,
                   "v": to(cls.pil_bicubic_coeffs(h, nh)) if nh != h else None}
            cls._pre_tables[key] = tab
        return ops.clip_preprocess_device(images.float().contiguous(), tab, size, IMAGE_MEAN, IMAGE_STD)

    @torch.no_grad()
    def encode_pixels(self, pixels, tok_scale=None):
        """pixels: fp32 [n,3,224,224] already preprocessed -> fp32 [n, 257, 768]. tok_scale: [n,257] (masked variant)."""
        ops = _ops()
        dev = self.get_device()
        pixels = pixels.to(dev).float().contiguous()
        require_cuda(pixels, "CLIPImageContextEncoder")
        p = self.packed()
        B = pixels.shape[0]
        g = pixels.shape[-1] // p["patch"]
        L = g * g + 1
        Lp = (L + 7) // 8 * 8
        col = ops.patchify(pixels, p["patch"], p["kpad"])
        pe = ops.gemm(col, p["wpatch"])                              # patch_embedding conv (stride = kernel) as a GEMM
        ts = None if tok_scale is None else tok_scale.to(dev).float().contiguous()
        x = ops.vit_assemble(pe, p["cls"], p["pos"], B, L, Lp, tok_scale=ts).view(B * Lp, -1)
        x = ops.layernorm(x, *p["pre"], eps=p["eps"])
        x = _ClipTower.run_layers(x, p["layers"], B, L, Lp, p["heads"], causal=False)
        x = ops.layernorm(x, *p["post"], eps=p["eps"])              # on ALL tokens (clip.py:97-98)
        z = ops.gemm(x, p["proj"])
        return ops.scale_by_row_norm(z.view(B, Lp, -1), L, idx=None, row_scale=ts)

    def _preprocess_any(self, images):
        import os
        if os.environ.get("VDB_CLIP_PRE_DEVICE") == "1" and isinstance(images, torch.Tensor) and images.is_cuda:
            return self.preprocess_device(images)
        return self.preprocess(images)

    def _encode(self, images):
        z = self.encode_pixels(self._preprocess_any(images))
        return z.half() if self.fp16 else z

    @torch.no_grad()
    def _encode_wmask(self, images, masks):
        """clip.py:103-143: per-token mask = mean of the (bilinear 224x224) mask over each patch, global token = mask
        mean; embeddings and the final tokens are multiplied by it."""
        assert isinstance(masks, torch.Tensor)
        assert (len(masks.shape) == 4) and (masks.shape[1] == 1)
        masks = torch.clamp(masks, 0, 1).float()
        masks = F.interpolate(masks, [224, 224], mode='bilinear')
        if masks.sum() == masks.numel():
            return self._encode(images)
        patch = self.model.vision_model.embeddings.patch_embedding.kernel_size[0]
        gscale = masks.mean(axis=[1, 2, 3], keepdim=True).flatten(2)
        vtoken = F.avg_pool2d(masks, patch, stride=patch).flatten(2).transpose(1, 2)          # conv with ones / P^2
        vtoken_mask = torch.cat([gscale, vtoken], dim=1).squeeze(-1)                            # [n, 257]
        z = self.encode_pixels(self._preprocess_any(images), tok_scale=vtoken_mask)
        return z.half() if self.fp16 else z

    def encode(self, images, masks=None):
        _require_clip_weights(self)
        return self._encode(images) if masks is None else self._encode_wmask(images, masks)
