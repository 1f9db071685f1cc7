from probabilisticdeepdiffusionmodels_b200.modules import get_model, get_unet  # noqa: F401
from probabilisticdeepdiffusionmodels_b200.unet import UNetModel  # noqa: F401
