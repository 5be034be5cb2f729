"""TEST INFRASTRUCTURE — not part of the product.

Imports the *unmodified* reference modules from /root/reference (read-only) on CPU, through the shim set of
SURVEY.md §8(c): stub packages for dependencies that are absent in this image (pytorch_lightning, omegaconf,
matplotlib, taming, mmcv, xformers, torchvision.transforms.functional_tensor) and a bypass of the glob-importing
``basicsr`` package ``__init__``s.  ``xformers.ops.memory_efficient_attention`` is provided as exact softmax
attention (``F.scaled_dot_product_attention``), which is the published semantics of the un-vendored dependency
(requirements.txt:20, unpinned).

Oracle patch D1 (documented deviation): the rearrange helpers of ldm/modules/diffusionmodules/util.py:271-288 are
called with ``t=None`` by SpatialTemporalConv.forward / TemporalAttention.forward (util.py:301-307,
attention.py:135-141), which crashes; the intent (``b = bt // num_frames`` one line above) is ``t = num_frames``, so
the helpers are wrapped to infer ``t`` from ``b``.

Only used (a) here, in the build container, to validate oracle/torch_ref.py and to generate tests/golden/*, and
(b) by CPU tests that skip when /root/reference is absent (it does not exist on the GPU box).
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("MGLD_REFERENCE_ROOT", "/root/reference")
_installed = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "ldm"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def install():
    """Install the shims and put the reference root on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    # (1) basicsr packages without their glob-importing __init__
    for name in ["basicsr", "basicsr.archs", "basicsr.data", "basicsr.ops", "basicsr.utils_pkg_placeholder"]:
        if name.endswith("placeholder"):
            continue
        _pkg(name, os.path.join(REF_ROOT, *name.split(".")))

    # (2) pytorch_lightning
    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    def seed_everything(seed):
        import random

        import numpy as np
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        return seed

    pl = _mod("pytorch_lightning", LightningModule=LightningModule, seed_everything=seed_everything)
    _mod("pytorch_lightning.utilities")
    _mod("pytorch_lightning.utilities.distributed", rank_zero_only=lambda f: f)
    pl.utilities = sys.modules["pytorch_lightning.utilities"]

    # (3) omegaconf
    class ListConfig(list):
        pass

    class DictConfig(dict):
        pass

    _mod("omegaconf", ListConfig=ListConfig, DictConfig=DictConfig, OmegaConf=None)
    _mod("omegaconf.listconfig", ListConfig=ListConfig)
    # (3b) skimage (imported by scripts/util_image.py:14 for I/O helpers the hot path never calls)
    _mod("skimage", img_as_ubyte=lambda a: a, img_as_float32=lambda a: a)
    # (4) matplotlib
    _mod("matplotlib")
    _mod("matplotlib.pyplot")
    # (5) taming
    _mod("taming"); _mod("taming.modules"); _mod("taming.modules.vqvae")
    _mod("taming.modules.vqvae.quantize", VectorQuantizer2=type("VectorQuantizer2", (nn.Module,), {}))
    # (6) mmcv
    _mod("mmcv"); _mod("mmcv.ops", Correlation=type("Correlation", (nn.Module,), {}))
    # (7) torchvision.transforms.functional_tensor (removed upstream)
    try:
        import torchvision.transforms.functional as TF
        _mod("torchvision.transforms.functional_tensor", rgb_to_grayscale=TF.rgb_to_grayscale)
    except Exception:
        pass

    # (8) xformers: exact softmax attention
    def memory_efficient_attention(q, k, v, attn_bias=None, op=None, scale=None):
        assert attn_bias is None
        if q.dim() == 3:
            return F.scaled_dot_product_attention(q, k, v, scale=scale)
        # (B, M, H, K) layout
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=scale)
        return o.transpose(1, 2)

    xf = _mod("xformers")
    xf.ops = _mod("xformers.ops", memory_efficient_attention=memory_efficient_attention)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if os.path.join(REF_ROOT, "scripts") not in sys.path:   # the scripts import their siblings by bare name
        sys.path.append(os.path.join(REF_ROOT, "scripts"))

    # oracle patch D1: infer t from b in the 4d<->5d / 4d<->3d helpers
    util = importlib.import_module("ldm.modules.diffusionmodules.util")
    from einops import rearrange

    def from_4d_to_5d(inp, b, c, t, h, w):
        t = inp.shape[0] // b if t is None else t
        return rearrange(inp, "(b t) c h w -> b c t h w", b=b, c=c, t=t, h=h, w=w)

    def from_5d_to_4d(inp, b, c, t, h, w):
        return rearrange(inp, "b c t h w -> (b t) c h w")

    def from_4d_to_3d(inp, b, c, t, h, w):
        t = inp.shape[0] // b if t is None else t
        return rearrange(inp, "(b t) c h w -> (b h w) t c", b=b, c=c, t=t, h=h, w=w)

    def from_3d_to_4d(inp, b, c, t, h, w):
        t = inp.shape[1] if t is None else t
        return rearrange(inp, "(b h w) t c -> (b t) c h w", b=b, c=c, t=t, h=h, w=w)

    util.from_4d_to_5d, util.from_5d_to_4d = from_4d_to_5d, from_5d_to_4d
    util.from_4d_to_3d, util.from_3d_to_4d = from_4d_to_3d, from_3d_to_4d
    _installed = True


def ref(name):
    """Import a reference module by dotted name (after install())."""
    install()
    return importlib.import_module(name)
