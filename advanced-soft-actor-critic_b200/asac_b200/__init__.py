"""asac_b200 — Blackwell-native SAC learner step + prioritized replay (the hot path of
BlueFisher/Advanced-Soft-Actor-Critic), behind the reference's own Python API.

    from asac_b200 import SAC_Base, PrioritizedReplayBuffer       # or, unchanged reference imports:
    from algorithm.sac_base import SAC_Base                        # (alias package next to this one)
    import algorithm.nn_models as m

Importing this package does not need a GPU; constructing a learner or a buffer does, and
raises when ``libasac_b200.so`` or a CUDA device is missing (no CPU fallback).
"""
from . import nn_models  # noqa: F401

__all__ = ['SAC_Base', 'PrioritizedReplayBuffer', 'nn_models']


def __getattr__(name):
    if name == 'SAC_Base':
        from .sac_base import SAC_Base
        return SAC_Base
    if name == 'PrioritizedReplayBuffer':
        from .replay_buffer import PrioritizedReplayBuffer
        return PrioritizedReplayBuffer
    raise AttributeError(name)
