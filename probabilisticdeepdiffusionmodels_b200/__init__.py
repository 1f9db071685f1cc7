"""pddm-b200: B200-native (sm_100a) implementation of the improved-diffusion hot path of
ArturPrzybysz/ProbabilisticDeepDiffusionModels -- the UNet eps/variance network and the DDPM math around it --
behind the reference's own Python API.

    from probabilisticdeepdiffusionmodels_b200 import Engine, get_model

Importing the package does not touch CUDA; the kernels live in ``libpddm_b200.so`` (C ABI, ``include/pddm.h``),
built in-tree by ``python -m probabilisticdeepdiffusionmodels_b200.build``.  There is no CPU / eager fallback.
"""
__version__ = "0.1.0"

_LAZY = {
    "Engine": ("engine", "Engine"),
    "get_model": ("modules", "get_model"),
    "get_unet": ("modules", "get_unet"),
    "UNetModel": ("unet", "UNetModel"),
    "get_betas": ("schedules", "get_betas"),
    "UniformSampler": ("timesteps", "UniformSampler"),
    "ImportanceSampler": ("timesteps", "ImportanceSampler"),
    "StepwiseLog": ("timesteps", "StepwiseLog"),
    "Ema": ("weight_average", "Ema"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
