"""Configuration presets of the K-Planes path (see ``method_configs``)."""
from .method_configs import (  # noqa: F401
    KPLANES_DATAMANAGER,
    KPLANES_MODEL,
    KPLANES_OPTIMIZERS,
    KPLANES_TRAINER,
    kplanes_model_config,
)
