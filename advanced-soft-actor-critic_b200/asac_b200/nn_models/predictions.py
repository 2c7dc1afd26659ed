"""Prediction heads of the plugin ``nn`` surface (reference: algorithm/nn_models/predictions.py —
transition :7-55, reward :58-80, observation :83-105).  Plain torch modules; the B200 learner raises
for ``use_prediction`` (outside the hot path), the classes exist so that plugin files
import unchanged."""
from __future__ import annotations

import torch
from torch import nn

from .layers import LinearLayers

__all__ = ['ModelBaseTransition', 'ModelTransition', 'ModelBaseReward', 'ModelReward', 'ModelBaseObservation']


class ModelBaseTransition(nn.Module):
    """(obs_list, state, action) -> Normal over the next state."""

    def __init__(self, state_size, d_action_size, c_action_size, use_extra_data):
        super().__init__()
        self.state_size = state_size
        self.d_action_size = d_action_size
        self.c_action_size = c_action_size
        self.use_extra_data = use_extra_data
        self.action_size = d_action_size + c_action_size
        self._build_model()

    def _build_model(self):
        pass

    def forward(self, obs_list, state, action):
        raise Exception('ModelBaseTransition not implemented')

    def extra_obs(self, obs_list):
        raise Exception('ModelBaseTransition.extra_obs not implemented')


class ModelTransition(ModelBaseTransition):
    def _build_model(self, dense_n=64, dense_depth=0, extra_size=0):
        if self.use_extra_data and extra_size == 0:
            raise Exception('use_extra_data is True but extra_size is zero')
        in_dim = self.state_size + self.action_size + (extra_size if self.use_extra_data else 0)
        self.dense = LinearLayers(in_dim, dense_n, dense_depth, self.state_size * 2)

    def forward(self, obs_list, state, action):
        parts = [state, self.extra_obs(obs_list), action] if self.use_extra_data else [state, action]
        mean, logstd = self.dense(torch.cat(parts, dim=-1)).chunk(2, dim=-1)
        return torch.distributions.Normal(mean, torch.exp(logstd).clamp(0.1, 1.0), validate_args=False)


class ModelBaseReward(nn.Module):
    """state -> predicted reward."""

    def __init__(self, state_size):
        super().__init__()
        self.state_size = state_size
        self._build_model()

    def _build_model(self):
        pass

    def forward(self, state):
        raise Exception('ModelBaseReward not implemented')


class ModelReward(ModelBaseReward):
    def _build_model(self, dense_n=64, dense_depth=0):
        self.dense = LinearLayers(self.state_size, dense_n, dense_depth, 1)

    def forward(self, state):
        return self.dense(state)


class ModelBaseObservation(nn.Module):
    """state -> reconstructed observation(s); subclasses also supply ``get_loss(state, obs_list)``."""

    def __init__(self, state_size, obs_shapes, use_extra_data):
        super().__init__()
        self.state_size = state_size
        self.obs_shapes = obs_shapes
        self.use_extra_data = use_extra_data
        self._build_model()

    def _build_model(self):
        pass

    def forward(self, state):
        raise Exception('ModelBaseObservation not implemented')

    def get_loss(self, state, obs_list) -> torch.Tensor:
        raise Exception('ModelBaseObservation.get_loss not implemented')
