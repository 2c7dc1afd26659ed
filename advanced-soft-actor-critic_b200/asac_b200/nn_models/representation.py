"""Representation-model classes of the plugin surface (reference: algorithm/nn_models/representation.py —
base / simple / attention :9-139, option selector :145-243, siamese projection / prediction heads :246-306)."""
from __future__ import annotations

from pathlib import Path

import torch
from torch import nn

from .layers import LinearLayers

__all__ = ['ModelBaseRep', 'ModelSimpleRep', 'ModelBaseAttentionRep', 'ModelBaseOptionSelectorRep',
           'ModelBaseOptionSelectorAttentionRep', 'ModelVOverOptions', 'ModelBaseRepProjection', 'ModelRepProjection',
           'ModelBaseRepPrediction', 'ModelRepPrediction']


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

    def get_state_from_encoders(self, encoders, seq_q_len, index, obs_list, pre_action, pre_seq_hidden_state,
                                is_prev_hidden_state=False, query_only_attend_to_rest_key=False, padding_mask=None):
        raise Exception('get_state_from_encoders not implemented')


# ---------------------------------------------------------------- option selector (representation.py:145-243)
class ModelBaseOptionSelectorRep(ModelBaseRep):
    """Representation of the option selector: takes ``use_dilation`` before ``model_abs_dir`` and a
    ``pre_termination_mask`` in ``forward``.  Option-critic training is outside the B200 hot path; the class
    exists so that plugin files import unchanged."""

    def __init__(self, obs_names, obs_shapes, d_action_sizes, c_action_size, is_target, use_dilation,
                 model_abs_dir: Path | None = None, **kwargs):
        self.use_dilation = use_dilation  # plain attribute: set before nn.Module.__init__ on purpose
        super().__init__(obs_names, obs_shapes, d_action_sizes, c_action_size, is_target, model_abs_dir, **kwargs)

    def forward(self, obs_list, pre_action, pre_seq_hidden_state, pre_termination_mask=None, padding_mask=None):
        raise Exception('ModelOptionSelectorRep not implemented')

    def __call__(self, obs_list, pre_action, pre_seq_hidden_state, pre_termination_mask=None, padding_mask=None):
        return nn.Module.__call__(self, obs_list, pre_action, pre_seq_hidden_state, pre_termination_mask,
                                  padding_mask)


class ModelBaseOptionSelectorAttentionRep(ModelBaseOptionSelectorRep, ModelBaseAttentionRep):
    def forward(self, seq_q_len, index, obs_list, pre_action, pre_seq_hidden_state, pre_termination_mask=None,
                is_prev_hidden_state=False, query_only_attend_to_rest_key=False, padding_mask=None):
        raise Exception('ModelOptionSelectorAttentionRep not implemented')

    def __call__(self, seq_q_len, index, obs_list, pre_action, pre_seq_hidden_state, pre_termination_mask=None,
                 is_prev_hidden_state=False, query_only_attend_to_rest_key=False, padding_mask=None):
        return nn.Module.__call__(self, seq_q_len, index, obs_list, pre_action, pre_seq_hidden_state,
                                  pre_termination_mask, is_prev_hidden_state, query_only_attend_to_rest_key,
                                  padding_mask)


class ModelVOverOptions(nn.Module):
    """state -> one value per option."""

    def __init__(self, state_size: int, num_options: int, is_target: bool):
        super().__init__()
        self.state_size = state_size
        self.num_options = num_options
        self.is_target = is_target
        self._build_model()

    def _build_model(self, dense_n=64, dense_depth=2):
        self.dense = LinearLayers(self.state_size, dense_n, dense_depth, self.num_options)

    def forward(self, state: torch.Tensor) -> torch.Tensor:
        return self.dense(state)


# ---------------------------------------------------------------- siamese heads (representation.py:246-306)
class _EncoderHead(nn.Module):
    _what = ''

    def __init__(self, encoder_size):
        super().__init__()
        self.encoder_size = encoder_size
        self._build_model()

    def _build_model(self):
        pass

    def forward(self, encoder: torch.Tensor):
        raise Exception(f'{self._what} not implemented')


class ModelBaseRepProjection(_EncoderHead):
    _what = 'ModelBaseRepProjection'


class ModelRepProjection(ModelBaseRepProjection):
    def _build_model(self, dense_n=None, dense_depth=1, projection_size=None):
        width = self.encoder_size if dense_n is None else dense_n
        self.dense = LinearLayers(self.encoder_size, width, dense_depth,
                                  width - 2 if projection_size is None else projection_size)

    def forward(self, encoder):
        return self.dense(encoder)


class ModelBaseRepPrediction(_EncoderHead):
    _what = 'ModelBaseRepPrediction'


class ModelRepPrediction(ModelBaseRepPrediction):
    def _build_model(self, dense_n=None, dense_depth=1):
        width = self.encoder_size if dense_n is None else dense_n
        self.dense = LinearLayers(self.encoder_size, width, dense_depth, self.encoder_size)

    def forward(self, encoder):
        return self.dense(encoder)
