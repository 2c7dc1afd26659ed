"""Image / ray encoders of the plugin ``nn`` surface (reference:
algorithm/nn_models/layers/image_layers.py — shape helpers :12-91, Conv1dLayers :94-143, stock conv
stacks :146-185, ConvLayers :188-229, ConvTransposeLayers :232-258, VisionTransformer :261-354,
Transform :357-378).  Same constructor arguments, attribute names (``state_dict`` keys) and results;
bodies restated.  ``ConvLayers(h, w, c, 'simple')`` is the encoder of BASELINE configs[4]
(tests/nn_conv_attn.py:11).
"""
from __future__ import annotations

import math
from functools import partial
from typing import Callable

import torch
from torch import nn

from .linear_layers import LinearLayers

__all__ = ['conv1d_output_size', 'conv2d_output_shape', 'pool_out_shape', 'convtranspose_output_shape',
           'default_conv1d', 'Conv1dLayers', 'small_visual', 'simple_visual', 'nature_visual', 'ConvLayers',
           'ConvTransposeLayers', 'VisionTransformer', 'Transform']


def _pair(v) -> tuple[int, int]:
    return v if isinstance(v, tuple) else (int(v), int(v))


def conv1d_output_size(l: int, kernel_size: int = 1, stride: int = 1, padding: int = 0, dilation: int = 1) -> int:
    """Output length of ``nn.Conv1d`` (torch's formula)."""
    return math.floor((l + 2 * padding - dilation * (kernel_size - 1) - 1) / stride + 1)


def conv2d_output_shape(h_w: tuple[int, int], kernel_size: int | tuple[int, int] = 1, stride: int = 1,
                        padding: int = 0, dilation: int = 1) -> tuple[int, int]:
    """Output (height, width) of ``nn.Conv2d``."""
    kh, kw = _pair(kernel_size)
    return (conv1d_output_size(h_w[0], kh, stride, padding, dilation),
            conv1d_output_size(h_w[1], kw, stride, padding, dilation))


def pool_out_shape(h_w: tuple[int, int], kernel_size: int, stride: int) -> tuple[int, int]:
    """Output (height, width) of ``nn.MaxPool2d`` without padding."""
    return (h_w[0] - kernel_size) // stride + 1, (h_w[1] - kernel_size) // stride + 1


def convtranspose_output_shape(h_w: tuple[int, int], kernel_size: int | tuple[int, int] = 1, stride: int = 1,
                               padding: int = 0, output_padding: int = 0, dilation: int = 1) -> tuple[int, int]:
    """Output (height, width) of ``nn.ConvTranspose2d``."""
    kh, kw = _pair(kernel_size)
    side = lambda n, k: (n - 1) * stride - 2 * padding + dilation * (k - 1) + output_padding + 1
    return side(h_w[0], kh), side(h_w[1], kw)


def default_conv1d(l, channels) -> tuple[nn.Module, int, int]:
    """Conv1d(c->16, k8 s4) LeakyReLU Conv1d(16->32, k4 s2) LeakyReLU -> (module, out length, 32)."""
    out_l = conv1d_output_size(conv1d_output_size(l, 8, 4), 4, 2)
    stack = nn.Sequential(nn.Conv1d(channels, 16, 8, 4), nn.LeakyReLU(), nn.Conv1d(16, 32, 4, 2), nn.LeakyReLU())
    return stack, out_l, 32


def _resolve(conv, table: dict, args: tuple, kind: str):
    if isinstance(conv, str):
        if conv not in table:
            raise RuntimeError(f'No pre-defined {conv} convolutional layer')
        return table[conv](*args)
    if isinstance(conv, tuple):
        return conv
    raise RuntimeError(f'Argument conv should a {kind}')


class Conv1dLayers(nn.Module):
    """Ray / 1-D signal encoder: ``[..., l, channels]`` -> conv stack over l -> ``LinearLayers``."""

    def __init__(self, in_l: int, in_channels: int, conv: str | tuple[nn.Module, int, int],
                 out_dense_n: int = 64, out_dense_depth: int = 0, output_size: int = None):
        super().__init__()
        self.conv_layers, l, out_c = _resolve(conv, {'default': default_conv1d}, (in_l, in_channels),
                                              'tuple[nn.Module, tuple[int, int], int]')
        self.conv_output_size = l * out_c
        self.dense = LinearLayers(self.conv_output_size, out_dense_n, out_dense_depth, output_size)
        self.output_size = self.dense.output_size

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() < 3:
            raise AssertionError('The dimension of input should be greater than or equal to 3')
        lead = x.shape[:-2]
        feat = self.conv_layers(x.reshape(-1, *x.shape[-2:]).permute(0, 2, 1))
        return self.dense(feat.reshape(*lead, self.conv_output_size))


def small_visual(height, width, channels) -> tuple[nn.Module, tuple[int, int], int]:
    hw = (height, width)
    for _ in range(2):
        hw = pool_out_shape(conv2d_output_shape(hw, 3, 1), 2, 2)
    stack = nn.Sequential(nn.Conv2d(channels, 35, [3, 3], [1, 1]), nn.LeakyReLU(), nn.MaxPool2d(2, 2),
                          nn.Conv2d(35, 144, [3, 3], [1, 1]), nn.LeakyReLU(), nn.MaxPool2d(2, 2))
    return stack, hw, 144


def simple_visual(height, width, channels) -> tuple[nn.Module, tuple[int, int], int]:
    """Conv2d(c->16, 8x8 s4) GELU Conv2d(16->32, 4x4 s2) GELU (image_layers.py:160-170)."""
    hw = conv2d_output_shape(conv2d_output_shape((height, width), 8, 4), 4, 2)
    stack = nn.Sequential(nn.Conv2d(channels, 16, [8, 8], [4, 4]), nn.GELU(),
                          nn.Conv2d(16, 32, [4, 4], [2, 2]), nn.GELU())
    return stack, hw, 32


def nature_visual(height, width, channels) -> tuple[nn.Module, tuple[int, int], int]:
    hw = (height, width)
    for k, s in ((8, 4), (4, 2), (3, 1)):
        hw = conv2d_output_shape(hw, k, s)
    stack = nn.Sequential(nn.Conv2d(channels, 32, [8, 8], [4, 4]), nn.LeakyReLU(),
                          nn.Conv2d(32, 64, [4, 4], [2, 2]), nn.LeakyReLU(),
                          nn.Conv2d(64, 64, [3, 3], [1, 1]), nn.LeakyReLU())
    return stack, hw, 64


class ConvLayers(nn.Module):
    """Image encoder: ``[..., C, H, W]`` -> conv stack ('small' | 'simple' | 'nature' | a
    ``(module, (h, w), channels)`` tuple) -> flatten -> ``LinearLayers`` (image_layers.py:188-229)."""

    def __init__(self, in_height: int, in_width: int, in_channels: int,
                 conv: str | tuple[nn.Module, tuple[int, int], int],
                 out_dense_n: int | list[int] = 64, out_dense_depth: int = 0, output_size: int = None):
        super().__init__()
        table = {'small': small_visual, 'simple': simple_visual, 'nature': nature_visual}
        self.conv_layers, (h, w), out_c = _resolve(conv, table, (in_height, in_width, in_channels),
                                                   'tuple[nn.Module, tuple[int, int], int]')
        self.conv_output_size = h * w * out_c
        self.dense = LinearLayers(self.conv_output_size, out_dense_n, out_dense_depth, output_size)
        self.output_size = self.dense.output_size

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() < 4:
            raise AssertionError('The dimension of input should be greater than or equal to 4')
        lead = x.shape[:-3]
        feat = self.conv_layers(x.reshape(-1, *x.shape[-3:]))
        return self.dense(feat.reshape(*lead, self.conv_output_size))


class ConvTransposeLayers(nn.Module):
    """Decoder: ``LinearLayers`` -> ``[channels, height, width]`` -> the given transposed-conv module."""

    def __init__(self, input_size: int, in_dense_n: int, in_dense_depth: int,
                 height: int, width: int, channels: int, conv_transpose: nn.Module):
        super().__init__()
        self._height, self._width, self._channels = height, width, channels
        self.dense = LinearLayers(input_size, in_dense_n, in_dense_depth, height * width * channels)
        self.conv_transpose = conv_transpose

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() < 2:
            raise AssertionError('The dimension of input should be greater than or equal to 2')
        flat = self.dense(x)
        vis = self.conv_transpose(flat.reshape(-1, self._channels, self._height, self._width))
        return vis.reshape(*flat.shape[:-1], *vis.shape[1:])


class VisionTransformer(nn.Module):
    """Patchify conv + class token + torchvision's transformer ``Encoder``; returns the class token
    (image_layers.py:261-354)."""

    def __init__(self, image_size: int, in_channels: int, patch_size: int, num_layers: int, num_heads: int,
                 hidden_dim: int, mlp_dim: int, dropout: float = 0.0, attention_dropout: float = 0.0,
                 norm_layer: Callable[..., torch.nn.Module] = partial(nn.LayerNorm, eps=1e-6)):
        super().__init__()
        from torchvision.models.vision_transformer import Encoder
        torch._assert(image_size % patch_size == 0, 'Input shape indivisible by patch size!')
        self.image_size, self.patch_size = image_size, patch_size
        self.hidden_dim, self.mlp_dim = hidden_dim, mlp_dim
        self.attention_dropout, self.dropout, self.norm_layer = attention_dropout, dropout, norm_layer
        self.conv_proj = nn.Conv2d(in_channels=in_channels, out_channels=hidden_dim, kernel_size=patch_size,
                                   stride=patch_size)
        self.class_token = nn.Parameter(torch.zeros(1, 1, hidden_dim))
        self.seq_length = (image_size // patch_size) ** 2 + 1
        self.encoder = Encoder(self.seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout,
                               attention_dropout, norm_layer)
        fan_in = in_channels * patch_size * patch_size
        nn.init.trunc_normal_(self.conv_proj.weight, std=math.sqrt(1 / fan_in))
        if self.conv_proj.bias is not None:
            nn.init.zeros_(self.conv_proj.bias)

    def _process_input(self, x: torch.Tensor) -> torch.Tensor:
        n, _, h, w = x.shape
        torch._assert(h == self.image_size, f'Wrong image height! Expected {self.image_size} but got {h}!')
        torch._assert(w == self.image_size, f'Wrong image width! Expected {self.image_size} but got {w}!')
        return self.conv_proj(x).reshape(n, self.hidden_dim, -1).permute(0, 2, 1)  # [n, patches, hidden]

    def forward(self, x: torch.Tensor):
        if x.dim() < 4:
            raise AssertionError('The dimension of input should be greater than or equal to 4')
        lead = x.shape[:-3]
        tokens = self._process_input(x.reshape(-1, *x.shape[-3:]))
        tokens = torch.cat([self.class_token.expand(tokens.shape[0], -1, -1), tokens], dim=1)
        return self.encoder(tokens)[:, 0].reshape(*lead, self.hidden_dim)


class Transform(nn.Module):
    """Applies an image transform to ``[..., C, H, W]`` by flattening the leading dimensions."""

    def __init__(self, transform: Callable[[torch.Tensor], torch.Tensor] | None = None):
        super().__init__()
        self.transform = transform

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.transform is None:
            return x
        if x.dim() < 4:
            raise AssertionError('The dimension of input should be greater than or equal to 4')
        out = self.transform(x.reshape(-1, *x.shape[-3:]))
        return out.reshape(*x.shape[:-3], *out.shape[1:])
