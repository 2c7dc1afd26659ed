"""``algorithm.nn_models.layers`` of the plugin surface (reference: algorithm/nn_models/layers/__init__.py:1-3)."""
from .image_layers import *  # noqa: F401,F403
from .linear_layers import *  # noqa: F401,F403
from .seq_layers import *  # noqa: F401,F403
