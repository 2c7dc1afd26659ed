"""Configuration enumerations callers pass to ``SAC_Base`` (the reference keeps them in
algorithm/utils/enums.py).  Only the sequence-encoder switch exists on the B200 hot path; ``SAC_Base``
also accepts any object whose ``.name`` is ``'RNN'`` (e.g. the reference's own enum member)."""
from enum import Enum

SEQ_ENCODER = Enum('SEQ_ENCODER', {'RNN': 1, 'ATTN': 2})
