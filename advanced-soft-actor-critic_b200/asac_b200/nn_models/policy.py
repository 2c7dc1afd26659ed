"""Policy plugin classes (reference: algorithm/nn_models/policy.py — NormalWithPadding :10-43,
JointOneHotCategorical :46-84, ModelBasePolicy / ModelPolicy :87-174, ModelTermination :177-194)."""
from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

from torch.distributions.utils import _standard_normal

from .layers import LinearLayers

__all__ = ['NormalWithPadding', 'JointOneHotCategorical', 'ModelBasePolicy', 'ModelPolicy', 'ModelTermination']


class NormalWithPadding(torch.distributions.Normal):
    """Normal whose padded action dimensions (``padding_mask`` over the last axis) sample 0 and report
    ``+inf`` log-probability / entropy, which ``sum_log_prob`` / ``sum_entropy`` then drop
    (policy.py:10-43, utils/operators.py:22-36)."""

    def __init__(self, loc, scale, padding_mask, validate_args=None):
        super().__init__(loc, scale, validate_args)
        self.padding_mask = padding_mask

    def sample(self, sample_shape=torch.Size()):
        draw = super().sample(sample_shape)
        draw[..., self.padding_mask] = 0.
        return draw

    def rsample(self, sample_shape=torch.Size()):
        eps = _standard_normal(self._extended_shape(sample_shape), dtype=self.loc.dtype, device=self.loc.device)
        keep = ~self.padding_mask
        return self.loc * keep + eps * (self.scale * keep)

    def log_prob(self, value):
        lp = super().log_prob(value)
        lp[self.padding_mask] = torch.inf
        return lp

    def entropy(self):
        ent = super().entropy()
        ent[self.padding_mask] = torch.inf
        return ent


class JointOneHotCategorical(torch.distributions.Distribution):
    """Independent one-hot categorical branches side by side: samples / probs / logits are the branches
    concatenated on the last axis, ``log_prob`` / ``entropy`` are stacked per branch (policy.py:46-84)."""

    def __init__(self, dists: list[torch.distributions.OneHotCategorical]):
        self._dists = dists
        self.logits_size_list = [d.logits.shape[-1] for d in dists]

    probs = property(lambda self: torch.cat([d.probs for d in self._dists], dim=-1))
    logits = property(lambda self: torch.cat([d.logits for d in self._dists], dim=-1))
    dists = property(lambda self: self._dists)

    def sample(self, sample_shape=torch.Size()) -> torch.Tensor:
        return torch.cat([d.sample(sample_shape) for d in self._dists], dim=-1)

    def sample_deter(self) -> torch.Tensor:
        """argmax of every branch as a one-hot."""
        picks = [nn.functional.one_hot(d.logits.argmax(dim=-1), size) for d, size in zip(self._dists,
                                                                                         self.logits_size_list)]
        return torch.cat(picks, dim=-1)

    def log_prob(self, value) -> torch.Tensor:
        parts = value.split(self.logits_size_list, dim=-1)
        return torch.stack([d.log_prob(v) for d, v in zip(self._dists, parts)], dim=-1)

    def entropy(self) -> torch.Tensor:
        return torch.stack([d.entropy() for d in self._dists], dim=-1)


class ModelBasePolicy(nn.Module):
    """``ModelPolicy(state_size, d_action_sizes, c_action_size, model_abs_dir, **nn_config['policy'])``
    (sac_base.py:416-419)."""

    def __init__(self, state_size: int, d_action_sizes: list[int], c_action_size: int,
                 model_abs_dir: Path | None = None, **kwargs):
        super().__init__()
        self.state_size = state_size
        self.d_action_sizes = d_action_sizes
        self.c_action_size = c_action_size
        self.model_abs_dir = model_abs_dir
        self._build_model(**kwargs)

    def _build_model(self, **kwargs):
        pass

    def forward(self, state, obs_list):
        raise Exception('ModelPolicy not implemented')


class ModelPolicy(ModelBasePolicy):
    """state -> dense -> c_dense -> (mean_dense, logstd_dense) -> Normal(5 tanh(mean/5),
    exp(clamp(logstd, -20, 0.5)))  (policy.py:165-170)."""

    def _build_model(self, dense_n=64, dense_depth=0, d_dense_n=64, d_dense_depth=3,
                     c_dense_n=64, c_dense_depth=3, mean_n=64, mean_depth=0,
                     logstd_n=64, logstd_depth=0, dropout=0.):
        self.dense = LinearLayers(self.state_size, dense_n, dense_depth, dropout=dropout)
        if self.d_action_sizes:
            self.d_dense_list = nn.ModuleList(
                LinearLayers(self.dense.output_size, d_dense_n, d_dense_depth, size, dropout=dropout)
                for size in self.d_action_sizes)
        if self.c_action_size:
            self.c_dense = LinearLayers(self.dense.output_size, c_dense_n, c_dense_depth, dropout=dropout)
            self.mean_dense = LinearLayers(self.c_dense.output_size, mean_n, mean_depth, self.c_action_size,
                                           dropout=dropout)
            self.logstd_dense = LinearLayers(self.c_dense.output_size, logstd_n, logstd_depth,
                                             self.c_action_size, dropout=dropout)

    def forward(self, state, obs_list):
        trunk = self.dense(state)
        d_policy = c_policy = None
        if self.d_action_sizes:
            d_policy = JointOneHotCategorical([torch.distributions.OneHotCategorical(logits=head(trunk),
                                                                                    validate_args=False)
                                               for head in self.d_dense_list])
        if self.c_action_size:
            hidden = self.c_dense(trunk)
            loc = 5. * torch.tanh(self.mean_dense(hidden) / 5.)
            scale = torch.exp(torch.clamp(self.logstd_dense(hidden), -20, 0.5))
            c_policy = torch.distributions.Normal(loc, scale, validate_args=False)
        return d_policy, c_policy


class ModelTermination(nn.Module):
    """Option termination probability: sigmoid(clamp(dense(state), -3, 3))  (policy.py:177-194)."""

    def __init__(self, state_size):
        super().__init__()
        self.state_size = state_size
        self._build_model()

    def _build_model(self, dense_n=64, dense_depth=2, dropout=0.):
        self.dense = LinearLayers(self.state_size, dense_n, dense_depth, output_size=1, dropout=dropout)

    def forward(self, state: torch.Tensor, obs_list: list[torch.Tensor]) -> torch.Tensor:
        return torch.sigmoid(self.dense(state).clamp(-3., 3.))
