"""Configuration enumerations and their yaml converters (reference: algorithm/utils/enums.py:1-44).
``SAC_Base`` also accepts any object whose ``.name`` matches (e.g. the reference's own enum members)."""
from enum import Enum

__all__ = ['SEQ_ENCODER', 'SIAMESE', 'CURIOSITY', 'convert_config_to_enum', 'convert_config_to_string']

SEQ_ENCODER = Enum('SEQ_ENCODER', {'RNN': 1, 'ATTN': 2})
SIAMESE = Enum('SIAMESE', {'ATC': 1, 'BYOL': 2})
CURIOSITY = Enum('CURIOSITY', {'FORWARD': 1, 'INVERSE': 2})

_KEYS = {'seq_encoder': SEQ_ENCODER, 'option_seq_encoder': SEQ_ENCODER, 'siamese': SIAMESE, 'curiosity': CURIOSITY}


def convert_config_to_enum(config: dict) -> None:
    """In place: the yaml strings of a ``sac_config`` -> enum members (None stays None)."""
    for key, enum in _KEYS.items():
        if config.get(key) is not None:
            config[key] = enum[config[key]]


def convert_config_to_string(config: dict) -> None:
    """In place: the inverse of :func:`convert_config_to_enum`."""
    for key in _KEYS:
        if config.get(key) is not None:
            config[key] = config[key].name
