"""Plugin ``nn`` surface: what ``envs/*/nn.py`` imports as ``algorithm.nn_models``
(reference: algorithm/nn_models/__init__.py:1-6 — the same six star-imports).  ``ModelQ`` / ``ModelPolicy`` /
``ModelSimpleRep`` / the stock ``GRU`` are lowered onto the CUDA kernels by the learner
(``asac_b200/lowering.py``); the remaining classes are the reference's plugin vocabulary as plain torch
modules, so that every plugin file imports unchanged."""
from .exploration import *  # noqa: F401,F403
from .layers import *  # noqa: F401,F403
from .policy import *  # noqa: F401,F403
from .predictions import *  # noqa: F401,F403
from .q import *  # noqa: F401,F403
from .representation import *  # noqa: F401,F403
