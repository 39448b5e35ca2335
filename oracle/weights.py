"""TEST INFRASTRUCTURE ONLY. Deterministic, construction-order-independent synthetic weights.

Every tensor is drawn from its own torch CPU generator seeded by crc32(key) ^ seed, so the reference
(in the build container), the oracle restatement and the CUDA path (on the GPU box) all see identical
weights without shipping checkpoints.  Unlike the reference's own init, no tensor is left at zero
(zero_module() zeroes 55 tensors, which would make every residual branch a no-op — SURVEY.md §8c).
"""
import zlib

import torch


def tensor_for(key, shape, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if len(shape) >= 2:                                   # conv / linear / embedding weights
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        std = fan_in ** -0.5
        if "embedding" in key:
            std = 0.02
        t = torch.randn(shape, generator=g) * std
    elif leaf == "weight":                                # norm scales
        t = 1.0 + 0.1 * torch.randn(shape, generator=g)
    elif leaf in ("class_embedding",):
        t = 0.02 * torch.randn(shape, generator=g)
    elif leaf == "logit_scale":
        t = torch.full(shape, 2.6592)
    else:                                                 # biases
        t = 0.02 * torch.randn(shape, generator=g)
    return t.to(dtype)


SKIP_SUFFIXES = ("position_ids",)


def synth_state_dict(shapes, seed=0, keep=None):
    """shapes: {key: shape}. Schedule buffers etc. (1-D non-parameter keys listed in `keep`) are skipped."""
    sd = {}
    for k, shp in shapes.items():
        if k.endswith(SKIP_SUFFIXES) or (keep is not None and k not in keep):
            continue
        sd[k] = tensor_for(k, shp, seed)
    return sd


def param_shapes(module):
    """{key: shape} of the learnable tensors of an nn.Module (parameters only, reference key names)."""
    return {k: tuple(v.shape) for k, v in module.named_parameters()}
