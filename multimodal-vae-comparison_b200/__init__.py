"""mmvae_b200 -- B200-native latent + objective hot path of multimodal-vae-comparison.

Plugin registry mirrors reference models/__init__.py:1-8: the host loop looks a model class up by
``cfg.mixing.lower()`` (reference trainer.py:109), so ``getattr(mmvae_b200, "moe")`` etc. resolve to the drop-ins.
"""
__version__ = "0.1.0"

from .mmvae_base import TorchMMVAE  # noqa: E402,F401
from .mmvae_models import DMVAE as dmvae  # noqa: E402,F401
from .mmvae_models import MOE as moe  # noqa: E402,F401
from .mmvae_models import POE as poe  # noqa: E402,F401
from .mmvae_models import MoPOE as mopoe  # noqa: E402,F401
from .objectives import MultimodalObjective, ReconLoss, UnimodalObjective, unimodal_objective  # noqa: E402,F401
from .output_storage import VAEOutput  # noqa: E402,F401
from . import utils  # noqa: E402,F401
from .graphed import GraphedObjective  # noqa: E402,F401

MODEL_REGISTRY = {"moe": moe, "poe": poe, "mopoe": mopoe, "dmvae": dmvae}
