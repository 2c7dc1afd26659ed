"""Exploration heads of the plugin ``nn`` surface (reference: algorithm/nn_models/exploration.py —
ModelRND :7-60, ModelOptionSelectorRND :63-85, forward / inverse dynamics :88-138).  Plain torch
modules with the reference's attribute names; the B200 learner raises for ``use_rnd`` / ``curiosity``
(outside the hot path), the classes exist so that plugin files import unchanged."""
from __future__ import annotations

import torch
from torch import nn

from .layers import LinearLayers

__all__ = ['ModelRND', 'ModelOptionSelectorRND', 'ModelBaseForwardDynamic', 'ModelForwardDynamic',
           'ModelBaseInverseDynamic', 'ModelInverseDynamic']


def _stack_heads(heads, state: torch.Tensor) -> torch.Tensor:
    """[*batch, n_heads, f]: every head evaluated on the same state."""
    return torch.stack([head(state) for head in heads], dim=-2)


class ModelRND(nn.Module):
    """Random-network-distillation features: one net on the state, one per discrete action on the state,
    one on ``cat[state, c_action]``."""

    def __init__(self, state_size: int, d_action_summed_size: int, c_action_size: int):
        super().__init__()
        self.state_size = state_size
        self.d_action_summed_size = d_action_summed_size
        self.c_action_size = c_action_size
        self._build_model()

    def _build_model(self, dense_n=64, dense_depth=2, output_size=None):
        net = lambda in_dim: LinearLayers(in_dim, dense_n, dense_depth, output_size)
        self.s_dense = net(self.state_size)
        if self.d_action_summed_size:
            self.d_dense_list = nn.ModuleList(net(self.state_size) for _ in range(self.d_action_summed_size))
        if self.c_action_size:
            self.c_dense = net(self.state_size + self.c_action_size)

    def cal_s_rnd(self, state) -> torch.Tensor:
        return self.s_dense(state)                                   # [*batch, f]

    def cal_d_rnd(self, state) -> torch.Tensor:
        return _stack_heads(self.d_dense_list, state)                # [*batch, d_action_summed_size, f]

    def cal_c_rnd(self, state, c_action) -> torch.Tensor:
        return self.c_dense(torch.cat([state, c_action], dim=-1))    # [*batch, f]


class ModelOptionSelectorRND(nn.Module):
    def __init__(self, state_size, num_options: int):
        super().__init__()
        self.state_size = state_size
        self.num_options = num_options
        self._build_model()

    def _build_model(self, dense_n=64, dense_depth=2, output_size=None):
        self.dense_list = nn.ModuleList(LinearLayers(self.state_size, dense_n, dense_depth, output_size)
                                        for _ in range(self.num_options))

    def cal_rnd(self, state) -> torch.Tensor:
        return _stack_heads(self.dense_list, state)                  # [*batch, num_options, f]


class _DynamicBase(nn.Module):
    _what = ''

    def __init__(self, state_size, action_size):
        super().__init__()
        self.state_size = state_size
        self.action_size = action_size
        self._build_model()

    def _build_model(self):
        pass

    def forward(self, *args):
        raise Exception(f'{self._what} not implemented')


class ModelBaseForwardDynamic(_DynamicBase):
    """(state, action) -> predicted next state."""
    _what = 'ModelBaseForwardDynamic'


class ModelForwardDynamic(ModelBaseForwardDynamic):
    def _build_model(self, dense_n=64, dense_depth=2):
        self.dense = LinearLayers(self.state_size + self.action_size, dense_n, dense_depth, self.state_size)

    def forward(self, state: torch.Tensor, action: torch.Tensor):
        return self.dense(torch.cat([state, action], dim=-1))


class ModelBaseInverseDynamic(_DynamicBase):
    """(state_from, state_to) -> predicted action."""
    _what = 'ModelBaseInverseDynamic'


class ModelInverseDynamic(ModelBaseInverseDynamic):
    def _build_model(self, dense_n=64, dense_depth=2):
        self.dense = LinearLayers(2 * self.state_size, dense_n, dense_depth, self.action_size)

    def forward(self, state_from: torch.Tensor, state_to: torch.Tensor):
        return self.dense(torch.cat([state_from, state_to], dim=-1))
