"""Dense building blocks of the plugin ``nn`` surface.

Same constructor arguments, attribute names and ``state_dict`` keys as the reference's
``algorithm/nn_models/layers/linear_layers.py:24-119`` (``ResBlock.linear`` / ``LinearLayers.dense``)
so that every ``envs/*/nn.py`` and every ``.pth`` checkpoint loads unchanged.  These modules are
the *parameter containers* and the actor-side torch path; the learner's update runs on the
flat copies of the same storage inside the CUDA kernels (see ``asac_b200/lowering.py``).
"""
from __future__ import annotations

import torch
from torch import nn

__all__ = ['ResBlock', 'LinearLayers']


def _kaiming_zero_bias(linear: nn.Linear) -> nn.Linear:
    # linear_layers.py:40-42,105-108: kaiming_uniform_ weight (a=0), zero bias
    nn.init.kaiming_uniform_(linear.weight.data)
    nn.init.zeros_(linear.bias.data)
    return linear


class ResBlock(nn.Module):
    """y = act(Wx + b) (+ x when input and output widths agree)."""

    def __init__(self, input_size: int, output_size: int | None = None,
                 activation: type[nn.Module] | None = None, residual: bool = True):
        super().__init__()
        output_size = input_size if output_size is None else output_size
        self.residual = bool(residual) and input_size == output_size
        self.linear = _kaiming_zero_bias(nn.Linear(input_size, output_size))
        self.act = (nn.GELU if activation is None else activation)()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[-1] != self.linear.in_features:
            raise AssertionError(f'ResBlock expects {self.linear.in_features} features, got {x.shape[-1]}')
        y = self.act(self.linear(x))
        return y + x if self.residual else y


class LinearLayers(nn.Module):
    """``dense_depth`` ResBlocks (each followed by Dropout) and an optional Linear head."""

    def __init__(self, input_size: int, dense_n: int | list[int] = 64, dense_depth: int = 0,
                 output_size: int | None = None, activation: type[nn.Module] | None = None,
                 residual: bool = True, dropout: float = 0.):
        super().__init__()
        widths = list(dense_n) if isinstance(dense_n, (list, tuple)) else [dense_n] * dense_depth
        self.input_size = input_size
        self.output_size = input_size
        blocks: list[nn.Module] = []
        for width in widths:
            blocks += [ResBlock(self.output_size, width, activation=activation, residual=residual),
                       nn.Dropout(dropout)]
            self.output_size = width
        if output_size:
            blocks.append(_kaiming_zero_bias(nn.Linear(self.output_size, output_size)))
            self.output_size = output_size
        self.dense = nn.Sequential(*blocks)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[-1] != self.input_size:
            raise AssertionError(f'LinearLayers expects {self.input_size} features, got {x.shape[-1]}')
        return self.dense(x)
