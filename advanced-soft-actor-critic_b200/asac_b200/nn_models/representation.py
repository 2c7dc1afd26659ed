"""Representation-model base classes of the plugin surface
(reference: algorithm/nn_models/representation.py:9-139)."""
from __future__ import annotations

from pathlib import Path

import torch
from torch import nn


class ModelBaseRep(nn.Module):
    """Constructor contract: ``ModelRep(obs_names, obs_shapes, d_action_sizes, c_action_size,
    is_target, model_abs_dir, **nn_config['rep'])`` (sac_base.py:339-361)."""

    def __init__(self, obs_names: list[str], obs_shapes: list[tuple], d_action_sizes: list[int],
                 c_action_size: int, is_target: bool, model_abs_dir: Path | None = None, **kwargs):
        super().__init__()
        self.obs_names = obs_names
        self.obs_shapes = obs_shapes
        self.d_action_sizes = d_action_sizes
        self.c_action_size = c_action_size
        self.is_target = is_target
        self.model_abs_dir = model_abs_dir
        self._build_model(**kwargs)

    def _build_model(self, **kwargs):
        pass

    def forward(self, obs_list, pre_action, pre_seq_hidden_state, padding_mask=None):
        raise Exception('ModelRep not implemented')

    def _get_empty_seq_hidden_state(self, state: torch.Tensor) -> torch.Tensor:
        return state.new_zeros((*state.shape[:-1], 0))

    def get_augmented_encoders(self, obs_list):
        raise Exception('get_augmented_encoders not implemented')

    def get_state_from_encoders(self, encoders, obs_list, pre_action, pre_seq_hidden_state, padding_mask=None):
        raise Exception('get_state_from_encoders not implemented')


class ModelSimpleRep(ModelBaseRep):
    """state = concat of the 1-D observations, empty hidden state (representation.py:74-83).
    The learner recognises this class and lets the replay gather write the concatenated state
    directly, so no kernel runs for it."""

    def forward(self, obs_list, pre_action, pre_seq_hidden_state, padding_mask=None):
        vectors = [o for o, shape in zip(obs_list, self.obs_shapes) if len(shape) == 1]
        state = torch.cat(vectors, dim=-1)
        return state, self._get_empty_seq_hidden_state(state)


class ModelBaseAttentionRep(ModelBaseRep):
    """Attention representation contract (representation.py:86-139); subclasses supply forward."""

    def forward(self, seq_q_len, index, obs_list, pre_action, pre_seq_hidden_state,
                is_prev_hidden_state=False, query_only_attend_to_rest_key=False, padding_mask=None):
        raise Exception('ModelAttentionRep not implemented')

    def __call__(self, seq_q_len, index, obs_list, pre_action, pre_seq_hidden_state,
                 is_prev_hidden_state=False, query_only_attend_to_rest_key=False, padding_mask=None):
        return nn.Module.__call__(self, seq_q_len, index, obs_list, pre_action, pre_seq_hidden_state,
                                  is_prev_hidden_state, query_only_attend_to_rest_key, padding_mask)
