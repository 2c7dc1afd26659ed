"""Plugin ``nn`` surface: what ``envs/*/nn.py`` imports as ``algorithm.nn_models``
(reference: algorithm/nn_models/__init__.py).  Only the classes on the B200 hot path exist;
see DESIGN.md §7 for what is out of scope."""
from .layers import GRU, LinearLayers, ResBlock
from .policy import ModelBasePolicy, ModelPolicy
from .q import ModelBaseQ, ModelQ
from .representation import ModelBaseAttentionRep, ModelBaseRep, ModelSimpleRep

__all__ = ['GRU', 'LinearLayers', 'ResBlock', 'ModelBasePolicy', 'ModelPolicy', 'ModelBaseQ', 'ModelQ',
           'ModelBaseAttentionRep', 'ModelBaseRep', 'ModelSimpleRep']
