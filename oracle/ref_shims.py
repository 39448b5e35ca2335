"""TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference from /root/reference on CPU.

Only usable where /root/reference exists (the build container): used by oracle/make_golden.py to
generate tests/golden/* and by tests/test_oracle.py to pin oracle/vd_oracle.py against the reference.
Shims: stub matplotlib + easydict, torch.cuda.device_count -> 1
(lib/sync.py:31-35 divides by it), DDIMSampler.register_buffer without the forced .to('cuda')
(ddim.py:17-21), cwd=/root/reference while configs are parsed (lib/cfg_helper.py:104).
"""
import contextlib
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "lib", "model_zoo"))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        d = dict(d or {}); d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) and not isinstance(x, _EasyDict) else x for x in v)
        elif isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        dict.__setattr__(self, k, v)
        dict.__setitem__(self, k, v)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {}); d.update(f)
        for k in d:
            setattr(self, k, d[k])

    def pop(self, k, *a):
        if hasattr(self, k):
            delattr(self, k)
        return dict.pop(self, k, *a)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_loaded = None


def load():
    """Returns a namespace with the reference's get_model, model_cfg_bank, DDIMSampler, modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    import torch
    # our own drop-in package is also called `lib`: make sure the reference's wins inside this process
    for name in [m for m in sys.modules if m == "lib" or m.startswith("lib.")]:
        del sys.modules[name]
    for n in ("matplotlib", "matplotlib.pyplot"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict"); m.EasyDict = _EasyDict; sys.modules["easydict"] = m
    if REF in sys.path:
        sys.path.remove(REF)
    sys.path.insert(0, REF)
    if not torch.cuda.is_available():
        torch.cuda.device_count = lambda: 1
    with _cwd(REF):
        from lib.model_zoo.common.get_model import get_model
        from lib.cfg_helper import model_cfg_bank
        from lib.model_zoo.ddim import DDIMSampler
        import lib.model_zoo.vd as vd
        import lib.model_zoo.openaimodel as openaimodel
        import lib.model_zoo.attention as attention
        import lib.model_zoo.autokl as autokl
        import lib.model_zoo.diffusion_utils as diffusion_utils
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)
    ns = types.SimpleNamespace(get_model=get_model, model_cfg_bank=model_cfg_bank, DDIMSampler=DDIMSampler, vd=vd,
                               openaimodel=openaimodel, attention=attention, autokl=autokl,
                               diffusion_utils=diffusion_utils)
    _loaded = ns
    return ns


def build_vd(unet_overrides=None, vae_overrides=None, with_text_ctx=True, with_vae=True, text_parts="c"):
    """The reference VD_v2_0 with diffuser.image (global+data+context), diffuser.text (context blocks
    only, configs/model/openai_unet.yaml:78-81) and vae.image — no CLIP, no Optimus, no 0D data blocks."""
    ns = load()
    with _cwd(REF):
        bank = ns.model_cfg_bank()
        ak = bank("autokl_v1"); ak.pop("pth")
        if vae_overrides:
            ak.args.ddconfig.update(vae_overrides)
        u2 = bank("openai_unet_2d_v1")
        u0 = bank("openai_unet_0d_v1_" + text_parts)     # "c": context blocks only (image sampling); "dc": + the text-latent data blocks
        if unet_overrides:
            u2.args.update(unet_overrides)
            u0.args.update({k: v for k, v in unet_overrides.items() if k in ("model_channels", "channel_mult", "num_heads", "context_dim")})
        vdc = bank("vd_base")
        vdc.args.vae_cfg_list = [["image", ak]] if with_vae else []
        vdc.args.ctx_cfg_list = []
        dl = [["image", u2]]
        if with_text_ctx:
            dl.append(["text", u0])
        vdc.args.diffuser_cfg_list = dl
        vdc.args.global_layer_ptr = "image"
        vdc.args.latent_scale_factor = {"image": 0.18215}
        net = ns.get_model()(vdc, verbose=False)
    net.eval()
    return net
