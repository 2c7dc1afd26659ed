"""Policy plugin classes (reference: algorithm/nn_models/policy.py:87-174)."""
from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

from .layers import LinearLayers


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
            raise NotImplementedError('discrete action branches are outside the B200 hot path (SURVEY.md §8f)')
        if self.c_action_size:
            hidden = self.c_dense(trunk)
            loc = 5. * torch.tanh(self.mean_dense(hidden) / 5.)
            scale = torch.exp(torch.clamp(self.logstd_dense(hidden), -20, 0.5))
            c_policy = torch.distributions.Normal(loc, scale, validate_args=False)
        return d_policy, c_policy
