"""AutoencoderKL (kl-f8) on vdb200 kernels — reference lib/model_zoo/autokl.py:14-49.
encode(x, out_posterior=False): x in [0,1] NCHW -> x*2-1 -> Encoder -> quant_conv -> Gaussian sample.
decode(z): post_quant_conv -> Decoder -> clamp((d+1)/2, 0, 1).
"""
import torch
import torch.nn as nn

from lib.model_zoo.common.get_model import register
from .autokl_modules import Encoder, Decoder
from .diffusion_utils import PackedModule, f32, require_cuda
from .distributions import DiagonalGaussianDistribution


def _ops():
    from vdb200 import ops
    return ops


@register('autoencoderkl')
class AutoencoderKL(PackedModule):
    def __init__(self, ddconfig, lossconfig, embed_dim):
        super().__init__()
        if lossconfig is not None:
            raise NotImplementedError("LPIPS/discriminator loss is training-only")
        ddconfig = dict(ddconfig)
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        assert ddconfig["double_z"]
        self.quant_conv = torch.nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = torch.nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim

    def _pack(self):
        return {"wq": f32(self.quant_conv.weight).view(self.quant_conv.out_channels, -1).contiguous(),
                "bq": f32(self.quant_conv.bias),
                "wpq": f32(self.post_quant_conv.weight).view(self.post_quant_conv.out_channels, -1).contiguous(),
                "bpq": f32(self.post_quant_conv.bias)}

    def _moments_nhwc(self, x):
        require_cuda(x, "AutoencoderKL.encode")
        ops = _ops()
        p = self.packed()
        xh = ops.nchw_to_nhwc(x.float().contiguous())
        h = self.encoder(xh, in_scale=2.0, in_shift=-1.0)          # x*2-1 folded into the im2col (autokl.py:34)
        return ops.pointwise_small(h, p["wq"], p["bq"])            # quant_conv (1x1, 8 -> 8)

    @torch.no_grad()
    def encode(self, x, out_posterior=False, noise=None, post_scale=1.0):
        """noise: optional standard-normal draw [B,embed_dim,h,w] (default: CPU generator like the reference)."""
        ops = _ops()
        m = self._moments_nhwc(x)
        if out_posterior:
            return DiagonalGaussianDistribution(ops.nhwc_to_nchw(m).to(x.dtype))
        B, h, w, _ = m.shape
        if noise is None:
            noise = torch.randn(B, self.embed_dim, h, w)            # distributions.py:36 (CPU draw, then move)
        nz = ops.nchw_to_nhwc(noise.to(m.device).float().contiguous())
        z = ops.gaussian_sample(m, nz, post_mul=post_scale)
        return ops.nhwc_to_nchw(z).to(x.dtype)

    def _post_quant_nhwc(self, z, pre_scale=1.0):
        """(pre_scale * z) -> post_quant_conv, NCHW in, fp32 NHWC out."""
        ops = _ops()
        p = self.packed()
        zh = ops.nchw_to_nhwc(z.float().contiguous())
        return ops.pointwise_small(zh, p["wpq"], p["bpq"], pre_mul=pre_scale)

    @torch.no_grad()
    def decode(self, z, pre_scale=1.0):
        require_cuda(z, "AutoencoderKL.decode")
        ops = _ops()
        zq = self._post_quant_nhwc(z, pre_scale)                              # (1/scale)*z then post_quant_conv
        dec = self.decoder(zq)                                                # fp32 NHWC [B,H,W,3]
        return ops.nhwc_to_nchw(dec, mul=0.5, add=0.5, clamp01=True).to(z.dtype)   # clamp((dec+1)/2, 0, 1)
