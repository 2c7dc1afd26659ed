"""Distribution helpers of the reference's ``algorithm/utils/operators.py:12-59`` as torch functions, for
plugin / actor-side code.  The learner's kernels implement the same formulas (``csrc/sac.cu``: summed
squash Jacobian subtracted from EVERY action dimension, 1e-2 floor, +inf masking)."""
from __future__ import annotations

import numpy as np
import torch

__all__ = ['squash_correction_log_prob', 'squash_correction_prob', 'sum_log_prob', 'prod_prob', 'sum_entropy',
           'gen_n_pre_actions']


def _squash_floor(x: torch.Tensor) -> torch.Tensor:
    return torch.clamp_min(1 - torch.tanh(x) ** 2, 1e-2)


def squash_correction_log_prob(dist: torch.distributions.Distribution, x: torch.Tensor) -> torch.Tensor:
    """log N(x) per dimension minus the SUMMED log-Jacobian of tanh (keepdim, so it lands on every dimension)."""
    return dist.log_prob(x) - torch.log(_squash_floor(x)).sum(dim=-1, keepdim=True)


def squash_correction_prob(dist: torch.distributions.Distribution, x: torch.Tensor) -> torch.Tensor:
    return torch.exp(dist.log_prob(x)) / _squash_floor(x).prod(dim=-1, keepdim=True)


def sum_log_prob(log_prob: torch.Tensor, keepdim=False) -> torch.Tensor:
    """In place: +inf entries (padded action dimensions) count as 0."""
    log_prob[log_prob == torch.inf] = 0.
    return log_prob.sum(-1, keepdim=keepdim)


def prod_prob(prob: torch.Tensor, keepdim=False) -> torch.Tensor:
    """In place: infinite factors count as 1; an infinite or NaN product becomes 1."""
    prob[torch.isinf(prob)] = 1.
    out = prob.prod(-1, keepdim=keepdim)
    out[~torch.isfinite(out)] = 1.
    return out


def sum_entropy(entropy: torch.Tensor) -> torch.Tensor:
    entropy[entropy == torch.inf] = 0.
    return entropy.sum(-1)


def gen_n_pre_actions(n_actions, keep_last_action=False):
    """[batch, n, A] actions -> the action taken BEFORE each step (zeros first); with ``keep_last_action``
    one step longer."""
    xp_zeros, xp_cat = (torch.zeros_like, torch.cat) if isinstance(n_actions, torch.Tensor) else \
        (np.zeros_like, np.concatenate)
    if n_actions.shape[1] == 0 and keep_last_action:
        shape = (n_actions.shape[0], 1, *n_actions.shape[2:])
        if isinstance(n_actions, torch.Tensor):
            return torch.zeros(shape, dtype=n_actions.dtype, device=n_actions.device)
        return np.zeros(shape, dtype=n_actions.dtype)
    body = n_actions if keep_last_action else n_actions[:, :-1]
    return xp_cat([xp_zeros(n_actions[:, 0:1]), body], 1)
