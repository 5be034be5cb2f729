"""YAML configs and the ``target:`` / ``params:`` plugin mechanism of the reference (ldm/util.py:78-103), accepted
unchanged.  OmegaConf is not a dependency: configs are plain dicts with attribute access.

``instantiate_from_config`` resolves the reference's dotted target strings to this package's classes through
``TARGET_ALIASES``; unknown ``torch.nn.*`` targets (e.g. ``torch.nn.Identity`` loss configs) resolve
to the real object, training-only targets resolve to ``None``.
"""
import importlib

import yaml


class Config(dict):
    """dict with attribute access (the subset of OmegaConf's DictConfig the reference scripts rely on)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(o):
    if isinstance(o, dict):
        return Config({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_wrap(v) for v in o]
    return o


def load_config(path):
    """OmegaConf.load(path) replacement."""
    with open(path) as f:
        return _wrap(yaml.safe_load(f))


TARGET_ALIASES = {
    "ldm.models.diffusion.ddpm.LatentDiffusionVSRTextWT": "mgld_vsr_b200.ddpm.LatentDiffusionVSRTextWT",
    "ldm.modules.diffusionmodules.openaimodel.InflatedUNetModelDualcondV2":
        "mgld_vsr_b200.unet.InflatedUNetModelDualcondV2",
    "ldm.modules.diffusionmodules.openaimodel.InflatedEncoderUNetModelWT":
        "mgld_vsr_b200.unet.InflatedEncoderUNetModelWT",
    "ldm.models.autoencoder.AutoencoderKL": "mgld_vsr_b200.autoencoder.AutoencoderKL",
    "ldm.models.autoencoder.VideoAutoencoderKLResi": "mgld_vsr_b200.autoencoder.VideoAutoencoderKLResi",
    "basicsr.archs.raft_arch.RAFT_SR": "mgld_vsr_b200.raft.RAFT_SR",
    "ldm.modules.encoders.modules.FrozenOpenCLIPEmbedder": "mgld_vsr_b200.ddpm.FrozenOpenCLIPEmbedder",
}
# training-only plugins the inference path never calls (loss functions, data modules)
IGNORED_PREFIXES = ("ldm.modules.losses.", "main.", "basicsr.data.", "taming.")


def get_obj_from_str(string):
    string = TARGET_ALIASES.get(string, string)
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config, **extra):
    """ldm/util.py:78-88"""
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    target = config["target"]
    if target.startswith(IGNORED_PREFIXES):
        return None
    params = dict(config.get("params", dict()) or {})
    params.update(extra)
    return get_obj_from_str(target)(**params)
