"""Shared test helpers: deterministic weights keyed by parameter name, tiny configs, synthetic inputs."""
import contextlib
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# tiny-but-structurally-complete configs (all kernel-visible channel counts stay multiples of 64)
TINY_UNET = dict(num_frames=2, image_size=32, in_channels=4, out_channels=4, model_channels=64,
                 attention_resolutions=[2, 1], num_res_blocks=1, channel_mult=[1, 2], num_head_channels=64,
                 use_spatial_transformer=True, use_linear_in_transformer=True, transformer_depth=1, context_dim=128,
                 use_checkpoint=False, legacy=False, semb_channels=64)
TINY_STRUCT = dict(num_frames=2, image_size=96, in_channels=4, model_channels=64, out_channels=64, num_res_blocks=1,
                   attention_resolutions=[2, 1], dropout=0, channel_mult=[1, 2], conv_resample=True, dims=2,
                   use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                   use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False)
TINY_DD = dict(double_z=True, num_frames=2, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64,
               ch_mult=[1, 2, 2, 2], num_res_blocks=1, attn_resolutions=[], dropout=0.0)


def det_tensor(key, shape, scale=None):
    """Deterministic pseudo-random tensor that depends only on (key, shape): CPU generator seeded by a hash of the key."""
    seed = int.from_bytes(hashlib.sha256(key.encode()).digest()[:4], "little")
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if scale is not None:
        t = t * scale
    return t


def det_state_dict(shapes):
    """shapes: {key: shape}.  Weights ~ N(0, 1/fan_in)-ish so activations stay O(1) through the network; norm scales
    near 1; temporal_alpha = 0.5 (the reference leaves it uninitialised); no zero-init modules."""
    sd = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        if k.endswith("temporal_alpha"):
            sd[k] = torch.full(shp, 0.5)
        elif k.endswith(".weight") and len(shp) == 1:      # norm gains
            sd[k] = 1.0 + 0.1 * det_tensor(k, shp)
        elif k.endswith(".bias"):
            sd[k] = 0.05 * det_tensor(k, shp)
        elif len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            sd[k] = det_tensor(k, shp, scale=fan_in ** -0.5)
        else:
            sd[k] = det_tensor(k, shp)
    return sd


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def psnr(a, b):
    """calculate_psnr_pt, basicsr/metrics/psnr_ssim.py:52-80, on [0,1] (n,3,h,w) tensors (crop_border 0, RGB): per-image
    10 log10(1 / (mse + 1e-8)), averaged over the frames"""
    mse = ((a.double() - b.double()) ** 2).mean(dim=[1, 2, 3])
    return float((10.0 * torch.log10(1.0 / (mse + 1e-8))).mean())


def close_frac(a, b, rtol=3e-3, atol=1e-4):
    """fraction of elements inside the north-star tolerance |a-b| <= atol + rtol*|b| (torch.allclose's criterion)"""
    a, b = a.float(), b.float()
    return ((a - b).abs() <= atol + rtol * b.abs()).float().mean().item()


class cpu_rng:
    """Context manager: every `torch.randn` / `torch.randn_like` inside draws from the CPU generator and is then moved to
    the requested device.  The reference run on the CPU (golden fixtures, oracle pipeline) and the product run on the GPU
    then consume ONE identical noise stream — on a CUDA device the reference itself mixes a CPU draw (posterior sample,
    distributions.py:36) with CUDA-generator draws, which no other machine could reproduce."""

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like

        def randn(*size, device=None, dtype=None, generator=None, **kw):
            if len(size) == 1 and not isinstance(size[0], int):
                size = tuple(size[0])
            out = self._randn(*size, dtype=dtype, generator=generator, **kw)
            return out.to(device) if device is not None else out

        def randn_like(x, **kw):
            return self._randn(tuple(x.shape), dtype=x.dtype).to(x.device)
        torch.randn, torch.randn_like = randn, randn_like
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like
        return False


@contextlib.contextmanager
def fp32_reference_math():
    """TF32 off for cuDNN convs and cuBLAS matmuls: the oracle on the GPU must be a true fp32 reference (a TF32 conv has a
    10-bit mantissa — coarser than the fp16 path under test)"""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def raft_state_dict(shapes):
    """deterministic RAFT weights: positive BatchNorm running_var, the duplicated norm3 / downsample.1 entries of the
    reference's state_dict kept identical (they are one module there), and a small flow head so that ten random-init
    GRU iterations stay bounded."""
    sd = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        if k.endswith("running_var"):
            sd[k] = 0.5 + det_tensor(k, shp).abs()
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * det_tensor(k, shp)
        else:
            sd[k] = det_state_dict({k: shp})[k]
    for k in list(sd):
        if ".downsample.1." in k:
            sd[k.replace(".downsample.1.", ".norm3.")] = sd[k]
        if k.startswith("update_block.flow_head.conv2"):
            sd[k] = sd[k] * 0.02
    return sd
