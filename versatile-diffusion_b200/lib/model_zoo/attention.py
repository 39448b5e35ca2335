"""SpatialTransformer / BasicTransformerBlock / CrossAttention / GEGLU feed-forward on vdb200 kernels
(reference lib/model_zoo/attention.py:37-64, 152-266).  Same constructor signatures and parameter
names; activations are NHWC bf16 (the reference's 'b c h w -> b (h w) c' rearrange is a no-op here).

Kernel schedule of one SpatialTransformer (x: [B,H,W,C] bf16):
  GN32(eps 1e-6) -> proj_in GEMM -> [LN -> fused q|k GEMM + V^T GEMM -> flash attention -> to_out GEMM(+resid)]
  -> [LN -> q GEMM (K, V^T of the context cached across DDIM steps) -> flash attention -> to_out GEMM(+resid)]
  -> LN -> GEGLU GEMM -> FF-out GEMM(+resid) -> proj_out GEMM (+ x_in, * mixing ratio)
"""
import torch
from torch import nn

from .diffusion_utils import PackedModule, bf16, f32, require_cuda, zero_module, pack_conv1x1


def _ops():
    from vdb200 import ops
    return ops


class PaddedContext(object):
    """A context batch already in kernel layout: bf16 [B, Lp, C] with Lp = L rounded up to 8 and zero pad rows.
    DDIMSampler keeps one per context in a persistent buffer so the captured CUDA graph (and the cached K / V^T
    of every cross-attention layer) stay valid when a new prompt is copied in."""

    def __init__(self, data, length):
        self.data, self.length = data, length


def Normalize(in_channels):
    return torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class GEGLU(PackedModule):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def _pack(self):
        ops = _ops()
        w, b = self.proj.weight, self.proj.bias
        n2 = w.shape[0] // 2
        half = 128  # per 256-column tile: 128 value rows then their 128 gate rows
        if n2 % half:
            raise ValueError("GEGLU width must be a multiple of 128 for the fused epilogue")
        idx = []
        for t in range(n2 // half):
            idx += list(range(t * half, (t + 1) * half)) + list(range(n2 + t * half, n2 + (t + 1) * half))
        idx = torch.tensor(idx, device=w.device)
        return {"w": bf16(w.detach()[idx]), "b": f32(b.detach()[idx]), "act": ops.ACT_GEGLU}

    def forward(self, x):  # x: [rows, C] bf16
        p = self.packed()
        return _ops().gemm(x, p["w"], bias=p["b"], act=p["act"])


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("only the gated (GEGLU) feed-forward is on the VD hot path")
        inner_dim = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))


class CrossAttention(PackedModule):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        self.is_self = context_dim is None
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        self._kv_cache = None

    def _pad_heads(self, w, dpad):
        """[H*d, K] -> [H*dpad, K] with zero rows after each head's d rows"""
        h, d = self.heads, self.dim_head
        out = torch.zeros(h, dpad, w.shape[1], dtype=torch.bfloat16, device=w.device)
        out[:, :d] = w.detach().view(h, d, -1).to(torch.bfloat16)
        return out.view(h * dpad, -1).contiguous()

    def _pack(self):
        dk, dv = _ops().attention_pads(self.dim_head)
        wq, wk = self._pad_heads(self.to_q.weight, dk), self._pad_heads(self.to_k.weight, dk)
        self._kv_cache = None
        return {"dk": dk, "dv": dv, "wq": wq, "wk": wk, "wqk": torch.cat([wq, wk], 0).contiguous() if self.is_self else None,
                "wv": self._pad_heads(self.to_v.weight, dv), "wo": bf16(self.to_out[0].weight),
                "bo": f32(self.to_out[0].bias)}

    _KV_CACHE_SLOTS = 4

    def _kv_get(self, key):
        for ent in (self._kv_cache or ()):
            if ent[0] == key:
                return ent
        return None

    def _kv_put(self, ent):
        """small per-layer cache keyed by the context buffer: two contexts of the same type (sample_multicontext with two
        'image' c_infos) or two samplers sharing one model keep their own K / V^T buffers instead of evicting each other
        (which silently re-projected the context inside every step and forced a graph re-capture per sample() call)"""
        cache = [e for e in (self._kv_cache or ()) if e[0] != ent[0]]
        cache.append(ent)
        self._kv_cache = cache[-self._KV_CACHE_SLOTS:]

    def context_kv(self, context):
        """K [B*Lp, H*dk] and V^T [H*dvp, B*Lp] of a context; cached while the same tensor (and version) is passed
        again — the context is constant over the DDIM loop, so this runs once per sample() call, not per step."""
        p = self.packed()
        ops = _ops()
        if isinstance(context, PaddedContext):
            data, L = context.data, context.length
            B, Lp, Cc = data.shape
            key = ("padded", data.data_ptr(), (B, Lp, Cc), L)
            cflat = data.view(B * Lp, Cc)
            ent = self._kv_get(key)
            if ent is not None:
                k, vt, ver = ent[1]
                if ver != data._version:          # new prompt copied into the same buffer: refresh IN PLACE
                    ops.gemm(cflat, p["wk"], out=k)
                    ops.gemm(p["wv"], cflat, out=vt)
                    self._kv_put((key, (k, vt, data._version), data))
                return k, vt, L, Lp
            k = ops.gemm(cflat, p["wk"])
            vt = ops.gemm(p["wv"], cflat)
            self._kv_put((key, (k, vt, data._version), data))
            return k, vt, L, Lp
        # the cache entry HOLDS the context tensor: while it is alive no other allocation can reuse its address, so
        # (pointer, version, shape) identifies the contents (a freed-and-reallocated buffer would otherwise alias it)
        key = (context.data_ptr(), context._version, tuple(context.shape), context.dtype)
        ent = self._kv_get(key)
        if ent is not None:
            return ent[1]
        B, L, Cc = context.shape
        Lp = (L + 7) // 8 * 8   # kv stride per batch item must be a multiple of 8 (TMA alignment)
        cpad = torch.zeros(B, Lp, Cc, dtype=torch.bfloat16, device=context.device)
        cpad[:, :L] = context.to(torch.bfloat16)
        cflat = cpad.view(B * Lp, Cc)
        val = (ops.gemm(cflat, p["wk"]), ops.gemm(p["wv"], cflat), L, Lp)
        self._kv_put((key, val, context))
        return val

    def forward(self, x, context=None, resid=None, B=1):
        """x: [B*N, C] bf16 (already layer-normed); returns to_out(attn) + resid as [B*N, C] bf16."""
        ops = _ops()
        p = self.packed()
        H, d, dk = self.heads, self.dim_head, p["dk"]
        N = x.shape[0] // B
        o = torch.empty(x.shape[0], H * d, dtype=torch.bfloat16, device=x.device)
        if context is None and N % 8 == 0:
            qk = ops.gemm(x, p["wqk"])                    # [B*N, 2*H*dk]: q | k
            vt = ops.gemm(p["wv"], x)                     # [H*dvp, B*N]
            ops.attention(qk, qk, vt, o, B, H, N, N, d, scale=self.scale, q_col0=0, k_col0=H * dk)
        elif context is None:
            # ragged token count (latent sides not a multiple of 8 at this level): keys/values from a copy of
            # the tokens padded to a multiple of 8 per batch item (TMA alignment of the V^T columns)
            Np = (N + 7) // 8 * 8
            xp = torch.zeros(B, Np, x.shape[1], dtype=torch.bfloat16, device=x.device)
            xp[:, :N] = x.view(B, N, -1)
            xp = xp.view(B * Np, -1)
            q = ops.gemm(x, p["wq"])
            k = ops.gemm(xp, p["wk"])
            vt = ops.gemm(p["wv"], xp)
            ops.attention(q, k, vt, o, B, H, N, N, d, scale=self.scale, kv_bstride=Np)
        else:
            q = ops.gemm(x, p["wq"])
            k, vt, L, Lp = self.context_kv(context)
            ops.attention(q, k, vt, o, B, H, N, L, d, scale=self.scale, kv_bstride=Lp)
        return ops.gemm(o, p["wo"], bias=p["bo"], resid=resid)


class BasicTransformerBlock(PackedModule):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False):
        super().__init__()
        if disable_self_attn:
            raise NotImplementedError("disable_self_attn=True is not used by the VD configs")
        self.disable_self_attn = disable_self_attn
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint

    def _pack(self):
        return {n: (f32(getattr(self, n).weight), f32(getattr(self, n).bias)) for n in ("norm1", "norm2", "norm3")} | \
            {"w2": bf16(self.ff.net[2].weight), "b2": f32(self.ff.net[2].bias)}

    def forward(self, x, context=None, B=1):
        """x: [B*N, C] bf16 tokens (reference _forward, attention.py:214-218)."""
        ops = _ops()
        p = self.packed()
        x = self.attn1(ops.layernorm(x, *p["norm1"], eps=self.norm1.eps), None, resid=x, B=B)
        x = self.attn2(ops.layernorm(x, *p["norm2"], eps=self.norm2.eps), context, resid=x, B=B)
        h = self.ff.net[0](ops.layernorm(x, *p["norm3"], eps=self.norm3.eps))
        return ops.gemm(h, p["w2"], bias=p["b2"], resid=x)


class SpatialTransformer(PackedModule):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, disable_self_attn=False):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim,
                                   disable_self_attn=disable_self_attn) for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))

    def _pack(self):
        return {"g": f32(self.norm.weight), "b": f32(self.norm.bias), "win": pack_conv1x1(self.proj_in.weight),
                "bin": f32(self.proj_in.bias), "wout": pack_conv1x1(self.proj_out.weight), "bout": f32(self.proj_out.bias)}

    def forward(self, x, context=None, ratio=1.0, acc=None):
        """x: NHWC bf16 [B,H,W,C].  Returns ratio * proj_out(blocks(...)) + (acc if given else x):
        with acc=None, ratio=1 this is the reference forward (x + x_in, attention.py:255-266); with the
        running `acc` it is one term of VD_v2_0.context_mixing (vd.py:391-396) accumulated in the epilogue."""
        require_cuda(x, "SpatialTransformer")
        ops = _ops()
        p = self.packed()
        B, H, W, C = x.shape
        xn = ops.groupnorm(x, p["g"], p["b"], self.norm.eps)
        t = ops.gemm(xn.view(B * H * W, C), p["win"], bias=p["bin"])
        for blk in self.transformer_blocks:
            t = blk(t, context, B=B)
        base = x if acc is None else acc
        out = ops.gemm(t, p["wout"], bias=p["bout"], resid=base.view(B * H * W, -1), alpha=float(ratio))
        return out.view(B, H, W, -1)
