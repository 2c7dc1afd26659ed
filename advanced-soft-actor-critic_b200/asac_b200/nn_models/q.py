"""Critic plugin classes (reference: algorithm/nn_models/q.py:9-91)."""
from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

from .layers import LinearLayers

__all__ = ['ModelBaseQ', 'ModelQ']


class ModelBaseQ(nn.Module):
    """``ModelQ(state_size, d_action_sizes, c_action_size, is_target, model_abs_dir)``
    (sac_base.py:396-413)."""

    def __init__(self, state_size: int, d_action_sizes: list[int], c_action_size: int, is_target: bool,
                 model_abs_dir: Path | None = None):
        super().__init__()
        vars(self).update(state_size=state_size, d_action_sizes=d_action_sizes, c_action_size=c_action_size,
                          is_target=is_target, model_abs_dir=model_abs_dir)
        self._build_model()

    def _build_model(self):
        """Subclasses create their layers here (called once from the constructor)."""

    def forward(self, state, action, obs_list):
        raise Exception('ModelQ not implemented')


class ModelQ(ModelBaseQ):
    """state -> dense -> { one head per discrete branch, c_dense(cat[c_state_dense(.), c_action_dense(action)]) -> 1 }.
    Attribute names and registration order are the reference's (q.py:34-72): they are the ``state_dict`` keys
    and the parameter indices of the optimizer state in a checkpoint."""

    def _build_model(self, dense_n=64, dense_depth=0, d_dense_n=64, d_dense_depth=3,
                     c_state_n=64, c_state_depth=0, c_action_n=64, c_action_depth=0,
                     c_dense_n=64, c_dense_depth=3, dropout=0.):
        def mlp(in_dim, width, depth, out=None):
            return LinearLayers(in_dim, width, depth, out, dropout=dropout)

        self.dense = mlp(self.state_size, dense_n, dense_depth)
        trunk_width = self.dense.output_size
        if self.d_action_sizes:
            self.d_dense_list = nn.ModuleList(mlp(trunk_width, d_dense_n, d_dense_depth, k) for k in self.d_action_sizes)
        if self.c_action_size:
            self.c_state_dense = mlp(trunk_width, c_state_n, c_state_depth)
            self.c_action_dense = mlp(self.c_action_size, c_action_n, c_action_depth)
            self.c_dense = mlp(self.c_state_dense.output_size + self.c_action_dense.output_size, c_dense_n,
                               c_dense_depth, 1)

    def forward(self, state, c_action, obs_list):
        trunk = self.dense(state)
        d_qs = c_q = None
        if self.d_action_sizes:
            d_qs = torch.cat([head(trunk) for head in self.d_dense_list], dim=-1)
        if self.c_action_size:
            joint = torch.cat([self.c_state_dense(trunk), self.c_action_dense(c_action)], dim=-1)
            c_q = self.c_dense(joint)
        return d_qs, c_q
