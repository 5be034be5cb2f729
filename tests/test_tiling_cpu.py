"""Tile geometry of BASELINE.json's 720p / 1080p configurations, host logic only: the pad
rule of the script (quirk D13), the VAE tile starts of ImageSpliterTh (util_image.py:686) with the latent stride 750 // 8
(quirk D10) and the UNet tile grid of p_mean_variance_canvas (ddpm.py:4203-4231)."""
import torch

from mgld_vsr_b200.ddpm import LatentDiffusionVSRTextWT
from mgld_vsr_b200.pipeline import ImageSpliterTh, VSRPipeline


class _StubModel:
    ops = None

    def respace(self, steps):
        return None, None


def _units(h, w, T=2):
    pipe = VSRPipeline(_StubModel(), None, ddpm_steps=50, n_frames=T)
    pipe.upsample_scale = pipe.upscale
    meta, units = pipe._segment_units(torch.zeros(T, 3, h, w), use_guidance=False)
    return meta, units


def test_1080p_pad_and_vae_tiles():
    meta, units = _units(1080, 1920)
    assert tuple(meta["im"].shape[-2:]) == (1088, 1952)                # D13: both dims grow, 1920 -> 1952
    assert len(units) == 6 and all(tuple(u[0].shape[-2:]) == (960, 960) for u in units)
    starts = sorted({(i[0], i[2]) for i in meta["infos"]})
    assert starts == [(0, 0), (0, 750), (0, 992), (128, 0), (128, 750), (128, 992)]
    lat = ImageSpliterTh(torch.zeros(1, 1, 1088 // 8, 1952 // 8), 960 // 8, 750 // 8)
    assert lat.height_starts_list == [0, 16] and lat.width_starts_list == [0, 93, 124]      # D10: 93, not 93.75


def test_720p_pad_and_vae_tiles():
    meta, units = _units(720, 1280)
    assert tuple(meta["im"].shape[-2:]) == (736, 1312)
    assert len(units) == 2 and all(tuple(u[0].shape[-2:]) == (736, 960) for u in units)
    assert sorted(i[2] for i in meta["infos"]) == [0, 352]


def test_512_is_one_unit_and_unpadded():
    meta, units = _units(512, 512)
    assert not meta["flag_pad"] and len(units) == 1 and meta["sp"] is None


def test_unet_tile_grids():
    off = LatentDiffusionVSRTextWT._tile_offsets
    assert off(64, 64, 64, 32) == [(0, 0)]                               # 512^2: one tile per step
    g = off(120, 120, 64, 32)                                            # one 960^2 VAE tile of the 1080p config
    assert len(g) == 9 and sorted({o[0] for o in g}) == [0, 32, 56] and sorted({o[1] for o in g}) == [0, 32, 56]
    g = off(92, 120, 64, 32)                                             # 736 x 960 VAE tile of the 720p config
    assert len(g) == 6 and sorted({o[0] for o in g}) == [0, 32, 56] and sorted({o[1] for o in g}) == [0, 28]
