"""Sequence layers of the plugin ``nn`` surface: the stock ``GRU`` wrapper and the episode attention
stack (reference: algorithm/nn_models/layers/seq_layers.py — GRU :14-114, enums :117-128,
MultiheadAttention :131-294, gates :297-345, EpisodeMultiheadAttentionBlock :348-547,
EpisodeMultiheadAttention :550-731, positional encodings :734-851).

Constructor arguments, attribute names (= ``state_dict`` keys) and results are the reference's, so
plugin files and checkpoints load unchanged; the bodies are restated.  These torch modules are the
parameter containers, the actor-side path and the probe the learner checks its kernels against
(``csrc/rep_gru.cu`` runs the stock GRU on the same storage).
"""
from __future__ import annotations

import math
from enum import Enum

import torch
from torch import nn
from torch.nn.utils import rnn as rnn_utils

from .linear_layers import LinearLayers

__all__ = ['GRU', 'POSITIONAL_ENCODING', 'GATE', 'MultiheadAttention', 'GatedResidualLayer', 'GatedOutputLayer',
           'GatedRecurrentLayer', 'GatedCatLayer', 'EpisodeMultiheadAttentionBlock', 'EpisodeMultiheadAttention',
           'AbsolutePositionalEncoding', 'RotaryPositionalEncoding', 'RotaryPositionalEncoding2']


class GRU(nn.Module):
    """``num_layers`` single-layer batch-first ``nn.GRU`` modules under ``_grus``;
    ``forward(x [B, L, in], h0 [B, layers, H], padding_mask [B, L]) -> (output [B, L, H],
    hn [B, L, layers, H])`` — every layer's output at every step (seq_layers.py:41-114).

    With a padding mask a sequence may be padded on BOTH sides (burn-in rows before an episode start,
    rows after its end): the valid run is shifted to the front, the layers run on packed sequences of the
    valid lengths, and every layer's output is shifted back with the padded steps zeroed
    (seq_layers.py:60-103)."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1, bias: bool = True,
                 dropout: float = 0.0, device=None, dtype=None) -> None:
        super().__init__()
        self.num_layers = num_layers
        self._grus = nn.ModuleList(
            nn.GRU(input_size=hidden_size if i else input_size, hidden_size=hidden_size, num_layers=1, bias=bias,
                   batch_first=True, dropout=dropout, device=device, dtype=dtype)
            for i in range(num_layers))

    def forward(self, x: torch.Tensor, h0: torch.Tensor | None = None, padding_mask: torch.Tensor | None = None):
        starts = None if h0 is None else h0.transpose(0, 1).contiguous()  # [layers, B, H]
        first = lambda i: None if starts is None else starts[i:i + 1]
        per_layer = []
        if padding_mask is None:
            for i, cell in enumerate(self._grus):
                x, _ = cell(x, first(i))
                per_layer.append(x)
            return x, torch.stack(per_layer, dim=2)

        B, L, _ = x.shape
        steps = torch.arange(L, device=x.device).expand(B, L)
        lead = padding_mask.long().argmin(dim=1, keepdim=True)       # padded steps before the first valid one
        lengths = (~padding_mask).sum(dim=1).cpu().clamp(min=1)
        src = (steps + lead).clamp(max=L - 1)                        # valid run moved to the front
        dst = (steps - lead).clamp(min=0)                            # ... and back
        stream = rnn_utils.pack_padded_sequence(x.gather(1, src.unsqueeze(-1).expand(-1, -1, x.shape[-1])), lengths,
                                                batch_first=True, enforce_sorted=False)
        out = None
        for i, cell in enumerate(self._grus):
            stream, _ = cell(stream, first(i))
            front, _ = rnn_utils.pad_packed_sequence(stream, batch_first=True, total_length=L)
            out = front.gather(1, dst.unsqueeze(-1).expand(-1, -1, front.shape[-1]))
            out = out.masked_fill(padding_mask.unsqueeze(-1), 0.0)
            per_layer.append(out)
        return out, torch.stack(per_layer, dim=2)


class POSITIONAL_ENCODING(Enum):
    ABSOLUTE = 1
    ABSOLUTE_CAT = 2
    ROPE = 3
    ROPE2 = 4


class GATE(Enum):
    RESIDUAL = 1
    OUTPUT = 2
    RECURRENT = 3
    CAT = 4


class AbsolutePositionalEncoding(nn.Module):
    """Sinusoid table indexed by step: column 2k holds sin(pos / 10000^(4k/d)), column 2k+1
    cos(pos / 10000^((4k+2)/d)) — the reference's exponent ``2 * column / d`` (seq_layers.py:734-749)."""

    def __init__(self, d_model: int, max_seq_len: int = 5000):
        super().__init__()
        self.d_model = d_model
        pos = torch.arange(max_seq_len, dtype=torch.float64).unsqueeze(1)
        col = torch.arange(d_model, dtype=torch.float64).unsqueeze(0)
        angle = pos / torch.pow(torch.tensor(10000., dtype=torch.float64), 2. * col / d_model)
        table = torch.where((torch.arange(d_model) % 2 == 0).unsqueeze(0), torch.sin(angle), torch.cos(angle))
        self.register_buffer('pe', table.to(torch.float32))

    @torch.no_grad()
    def forward(self, indexes):
        return self.pe[indexes.type(torch.int64)]


class RotaryPositionalEncoding(nn.Module):
    """Rotation of consecutive feature pairs by ``pos * theta^(-2k/d)`` as a complex product
    (seq_layers.py:753-794)."""

    def __init__(self, d_model: int, max_seq_len: int = 5000, theta: float = 10000.0):
        super().__init__()
        inv = 1.0 / (theta ** (torch.arange(0, d_model, 2)[:d_model // 2] / d_model))
        angles = torch.outer(torch.arange(max_seq_len), inv)
        self.register_buffer('freqs_cis', torch.polar(torch.ones_like(angles), angles))

    def _rotate(self, x, indexes):
        pairs = torch.view_as_complex(x.reshape(*x.shape[:-1], -1, 2))
        return torch.view_as_real(pairs * self.freqs_cis[indexes.type(torch.int64)]).flatten(2).type_as(x)

    def forward(self, xq_indexes, xk_indexes, xq, xk):
        return self._rotate(xq, xq_indexes), self._rotate(xk, xk_indexes)


class RotaryPositionalEncoding2(nn.Module):
    """Half-split rotary encoding: ``x * cos + [-x_hi, x_lo] * sin`` on the first ``d_model`` features
    (seq_layers.py:797-851)."""

    def __init__(self, d_model: int, max_seq_len: int = 5000, base: int = 10_000):
        super().__init__()
        self.d_model = d_model
        inv = 1. / (base ** (torch.arange(0, d_model, 2).float() / d_model))
        angles = torch.outer(torch.arange(max_seq_len).float(), inv).repeat(1, 2)
        self.register_buffer('cos_cached', angles.cos())
        self.register_buffer('sin_cached', angles.sin())

    def _rotate(self, x, indexes):
        d, half = self.d_model, self.d_model // 2
        head, rest = x[..., :d], x[..., d:]
        swapped = torch.cat([-head[:, :, half:], head[:, :, :half]], dim=-1)
        head = head * self.cos_cached[indexes] + swapped * self.sin_cached[indexes]
        return torch.cat((head, rest), dim=-1)

    def forward(self, xq_indexes, xk_indexes, xq, xk):
        return self._rotate(xq, xq_indexes), self._rotate(xk, xk_indexes)


class MultiheadAttention(nn.Module):
    """softmax(q k^T / sqrt(head_dim) + mask) v with ``LinearLayers`` projections
    (seq_layers.py:131-294).  Returns the attention output and the head-averaged weights.
    A query row whose keys are ALL masked gets a zero mask (uniform weights) during the softmax
    and zero output / zero weights afterwards (:257-259, :286-289)."""

    def __init__(self, embed_dim: int, num_heads: int = 1, pe: POSITIONAL_ENCODING | None = None,
                 qkv_dense_depth: int = 0, out_dense_depth: int = 0, out_size: int | None = None,
                 dropout: float = 0.) -> None:
        super().__init__()
        if embed_dim % num_heads:
            raise AssertionError('embed_dim must be divisible by num_heads')
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.pe, self.dropout = pe, dropout
        in_dim = embed_dim
        if pe in (POSITIONAL_ENCODING.ABSOLUTE, POSITIONAL_ENCODING.ABSOLUTE_CAT):
            self.abpe = AbsolutePositionalEncoding(embed_dim)
            in_dim = embed_dim * (2 if pe == POSITIONAL_ENCODING.ABSOLUTE_CAT else 1)
        elif pe == POSITIONAL_ENCODING.ROPE:
            self.rope = RotaryPositionalEncoding(embed_dim)
        elif pe == POSITIONAL_ENCODING.ROPE2:
            self.rope = RotaryPositionalEncoding2(embed_dim)
        proj = lambda: LinearLayers(in_dim, dense_n=embed_dim, dense_depth=qkv_dense_depth, output_size=embed_dim,
                                    dropout=dropout)
        self.q_proj, self.k_proj, self.v_proj = proj(), proj(), proj()
        self.out_proj = LinearLayers(embed_dim, dense_n=embed_dim, dense_depth=out_dense_depth, output_size=out_size,
                                     dropout=dropout)

    def forward(self, query, key, value, query_index=None, key_index=None, key_padding_mask=None, attn_mask=None):
        lead = query.shape[:-2]
        query, key, value = (t.reshape(-1, *t.shape[-2:]) for t in (query, key, value))
        bsz, Lq, Lk, h = query.shape[0], query.shape[1], key.shape[1], self.num_heads
        if attn_mask is not None and attn_mask.dim() not in (2, 3):
            raise AssertionError('attn_mask is [batch, seq_q_len, seq_k_len] or [seq_q_len, seq_k_len]')

        if self.pe is not None:
            if query_index is None:
                query_index = torch.arange(Lq, device=query.device).expand(bsz, Lq)
            if key_index is None:
                key_index = torch.arange(Lk, device=key.device).expand(bsz, Lk)
        if self.pe == POSITIONAL_ENCODING.ABSOLUTE:
            q_pe, k_pe = self.abpe(query_index), self.abpe(key_index)
            query, key, value = q_pe + query, k_pe + key, k_pe + value
        elif self.pe == POSITIONAL_ENCODING.ABSOLUTE_CAT:
            q_pe, k_pe = self.abpe(query_index), self.abpe(key_index)
            query, key, value = (torch.cat([query, q_pe], dim=-1), torch.cat([key, k_pe], dim=-1),
                                 torch.cat([value, k_pe], dim=-1))

        q, k, v = self.q_proj(query), self.k_proj(key), self.v_proj(value)
        if self.pe in (POSITIONAL_ENCODING.ROPE, POSITIONAL_ENCODING.ROPE2):
            q, k = self.rope(query_index, key_index, q, k)
        heads = lambda t: t.reshape(bsz, t.shape[1], h, self.head_dim).transpose(1, 2)  # [bsz, h, L, head_dim]
        scores = torch.matmul(heads(q) / math.sqrt(self.head_dim), heads(k).transpose(-2, -1))

        blocked = None if attn_mask is None else (attn_mask if attn_mask.dim() == 3 else attn_mask.expand(bsz, Lq, Lk))
        if key_padding_mask is not None:
            pad = key_padding_mask.reshape(-1, Lk).unsqueeze(1)
            blocked = pad.expand(bsz, Lq, Lk) if blocked is None else torch.logical_or(blocked, pad)
        dead = None
        if blocked is not None:
            dead = blocked.all(dim=-1)                                   # [bsz, Lq]: nothing to attend to
            bias = torch.zeros(bsz, Lq, Lk, dtype=query.dtype, device=query.device)
            bias.masked_fill_(torch.logical_and(blocked, ~dead.unsqueeze(-1)), float('-inf'))
            scores = scores + bias.unsqueeze(1)
        weights = torch.softmax(scores, dim=-1)
        if self.training and self.dropout > 0.:
            weights = nn.functional.dropout(weights, p=self.dropout)
        out = torch.matmul(weights, heads(v)).transpose(1, 2).reshape(bsz, Lq, self.embed_dim)
        weights = weights.mean(1) if h > 1 else weights[:, 0]
        out = self.out_proj(out)
        if dead is not None:
            live = ~dead.unsqueeze(-1)
            out, weights = out * live, weights * live
        return out.reshape(*lead, *out.shape[1:]), weights.reshape(*lead, *weights.shape[1:])


class GatedResidualLayer(nn.Module):
    def forward(self, x, y):
        return x + y


def _kaiming_linear(embed_dim: int, bias: bool) -> nn.Linear:
    lin = nn.Linear(embed_dim, embed_dim, bias=bias)
    nn.init.kaiming_uniform_(lin.weight.data)
    return lin


class GatedOutputLayer(nn.Module):
    """x + sigmoid(W x * y)  (seq_layers.py:302-311)."""

    def __init__(self, embed_dim: int):
        super().__init__()
        self.embed_dim = embed_dim
        self.dense = _kaiming_linear(embed_dim, False)

    def forward(self, x, y):
        return x + torch.sigmoid(self.dense(x) * y)


class GatedRecurrentLayer(nn.Module):
    """GRU-style gate between the block input x and the attention output y (seq_layers.py:314-340)."""

    def __init__(self, embed_dim: int):
        super().__init__()
        self.embed_dim = embed_dim
        for name, bias in (('dense_x_r', False), ('dense_y_r', False), ('dense_x_z', True), ('dense_y_z', False),
                           ('dense_x_g', False), ('dense_y_g', False)):
            setattr(self, name, _kaiming_linear(embed_dim, bias))

    def forward(self, x, y):
        reset = torch.sigmoid(self.dense_x_r(x) + self.dense_y_r(y))
        update = torch.sigmoid(self.dense_x_z(x) + self.dense_y_z(y))
        cand = torch.tanh(self.dense_x_g(reset * x) + self.dense_y_g(y))
        return (1 - update) * x + update * cand


class GatedCatLayer(nn.Module):
    def forward(self, x, y):
        return torch.cat([x, y], dim=-1)


class EpisodeMultiheadAttentionBlock(nn.Module):
    """One causal self-attention layer over an episode window (seq_layers.py:348-547): optional
    LayerNorm on the keys, ``MultiheadAttention`` with the mask of :meth:`get_attn_mask`, an optional gate
    with the un-normalised query, and zeroed outputs on padded steps."""

    def __init__(self, embed_dim: int, num_heads: int, pe: POSITIONAL_ENCODING | None = None,
                 qkv_dense_depth: int = 0, out_dense_depth: int = 1, dropout: float = 0.,
                 gate: GATE | None = None, use_layer_norm: bool = False):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.gate, self.use_layer_norm = gate, use_layer_norm
        self.output_dim = embed_dim * (2 if gate == GATE.CAT else 1)
        if use_layer_norm:
            self.layer_norm = nn.LayerNorm(embed_dim)
        self.attn = MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads, pe=pe,
                                       qkv_dense_depth=qkv_dense_depth, out_dense_depth=out_dense_depth,
                                       dropout=dropout)
        gates = {GATE.RESIDUAL: GatedResidualLayer, GATE.CAT: GatedCatLayer,
                 GATE.OUTPUT: lambda: GatedOutputLayer(embed_dim), GATE.RECURRENT: lambda: GatedRecurrentLayer(embed_dim)}
        if gate in gates:
            self.gatedlayer = gates[gate]()

    def get_attn_mask(self, seq_k_len: int, seq_q_len_only_attend_to_rest_key: int | None = None,
                      key_index: torch.Tensor | None = None, key_padding_mask: torch.Tensor | None = None,
                      device='cpu') -> torch.Tensor:
        """True = blocked.  Default: strictly causal.  With ``seq_q_len_only_attend_to_rest_key = q`` (option
        critic) the window is ``k - q`` rest keys followed by q queries.  Kept as the reference computes it:
        the top-left block is ``eye`` (a rest row sees every rest key but itself, :416-418); the query block
        stays fully blocked (:422-426 OR the complement of the identity into a block of ones, a no-op); given
        ``key_index`` a query row sees the rest keys whose index is not larger than its own (:428-441).
        ``key_padding_mask`` columns are blocked for every row (:446-453)."""
        k = seq_k_len
        if seq_q_len_only_attend_to_rest_key is None:
            mask = torch.ones(k, k, dtype=torch.bool, device=device).triu(diagonal=1)
        else:
            q = seq_q_len_only_attend_to_rest_key
            r = k - q
            mask = torch.ones(k, k, dtype=torch.bool, device=device)
            mask[:r, :r] = torch.eye(r, r, dtype=torch.bool, device=device)
            if key_index is not None:
                mask = mask.repeat(key_index.shape[0], 1, 1)
                mask[:, r:, :r] = key_index[:, r:].unsqueeze(-1) < key_index[:, :r].unsqueeze(1)
        if key_padding_mask is not None:
            if mask.dim() == 2:
                mask = mask.repeat(key_padding_mask.shape[0], 1, 1)
            mask = torch.logical_or(mask, key_padding_mask.unsqueeze(1))
        return mask

    def forward(self, key: torch.Tensor, seq_q_len: int, cut_query: bool = True,
                query_only_attend_to_rest_key: bool = False, key_index: torch.Tensor | None = None,
                key_padding_mask: torch.Tensor | None = None):
        """key [batch, seq_k_len, embed]; ``key_index`` / ``key_padding_mask`` may be shorter than seq_k_len
        (stored hidden states were prepended): they are left-extended with index -1 / the first mask value.
        -> (output [batch, seq_q_len | seq_k_len, output_dim], attn_weights)."""
        k_len = key.shape[1]
        raw_query = key[:, -seq_q_len:] if cut_query else key
        if self.use_layer_norm:
            key = self.layer_norm(key)
        if key_index is not None:
            short = k_len - key_index.shape[1]
            assert short >= 0
            key_index = torch.cat([key_index.new_full((key_index.shape[0], short), -1), key_index], dim=1)
        if key_padding_mask is not None:
            short = k_len - key_padding_mask.shape[1]
            assert short >= 0
            key_padding_mask = torch.cat([key_padding_mask[:, :1].repeat(1, short), key_padding_mask], dim=1)
        mask = self.get_attn_mask(k_len, seq_q_len if query_only_attend_to_rest_key else None, key_index,
                                  key_padding_mask, device=key.device)
        query, query_index = key, key_index
        if cut_query:
            query = key[:, -seq_q_len:]
            query_index = None if key_index is None else key_index[:, -seq_q_len:]
            mask = mask[..., -seq_q_len:, :]
        out, weights = self.attn(query, key, key, query_index=query_index, key_index=key_index, attn_mask=mask)
        if self.gate is not None:
            out = self.gatedlayer(raw_query, out)
        if key_padding_mask is not None:
            out = out * (~key_padding_mask[:, -out.shape[1]:]).to(out.dtype).unsqueeze(-1)
        return out, weights


class EpisodeMultiheadAttention(nn.Module):
    """Stack of :class:`EpisodeMultiheadAttentionBlock` with per-layer stored hidden states
    (seq_layers.py:550-731).  Every per-layer argument may be a scalar or a list of ``num_layers``.
    The hidden state of a step is the concatenation of the outputs of all layers but the last at
    that step, so that later calls can attend to earlier steps without re-encoding them:

    * ``hidden_state is None``: the whole window goes through every layer (learner, first pass);
    * ``hidden_state`` given, ``is_prev_hidden_state=False``: stored states of steps BEFORE the window are
      prepended to the keys of layers 1.. (learner with burn-in states, :664-687);
    * ``is_prev_hidden_state=True``: the same, the stored states being those of the previous steps as the
      actor collects them (:689-726).
    -> (encoded query, next hidden state [batch, seq_q_len, sum(dims[:-1])], list of attention weights)."""

    def __init__(self, embed_dim: int, num_layers: int = 2, num_heads: int | list[int] = 1,
                 pe=False, qkv_dense_depth: int | list[int] = 0, out_dense_depth: int | list[int] = 1,
                 dropout: float | list[float] = 0., gate=None, use_layer_norm: bool | list[bool] = False):
        super().__init__()
        self.num_layers = num_layers

        def per_layer(value):
            values = value if isinstance(value, list) else [value] * num_layers
            assert len(values) == num_layers
            return values

        columns = [per_layer(v) for v in (num_heads, pe, qkv_dense_depth, out_dense_depth, dropout, gate,
                                          use_layer_norm)]
        self._attn_list = nn.ModuleList()
        width = embed_dim
        for heads_i, pe_i, qkv_i, out_i, drop_i, gate_i, norm_i in zip(*columns):
            block = EpisodeMultiheadAttentionBlock(width, heads_i, pe=pe_i, qkv_dense_depth=qkv_i,
                                                   out_dense_depth=out_i, dropout=drop_i, gate=gate_i,
                                                   use_layer_norm=norm_i)
            self._attn_list.append(block)
            width = block.output_dim
        self._output_dim_list = [block.output_dim for block in self._attn_list]
        self.output_dim = width
        self.output_hidden_state_dim = sum(self._output_dim_list[:-1]) if num_layers > 1 else 1

    def forward(self, key: torch.Tensor, seq_q_len: int = 1, cut_query: bool = True,
                hidden_state: torch.Tensor | None = None, is_prev_hidden_state: bool = False,
                query_only_attend_to_rest_key: bool = False, key_index: torch.Tensor | None = None,
                key_padding_mask: torch.Tensor | None = None):
        k_len, n = key.shape[1], self.num_layers
        assert seq_q_len <= k_len
        common = dict(query_only_attend_to_rest_key=query_only_attend_to_rest_key, key_index=key_index,
                      key_padding_mask=key_padding_mask)
        blocks = list(self._attn_list)
        kept, weights = [], []   # per-layer outputs on the query steps (all layers but the last); attention maps
        stored = None
        if hidden_state is not None and n > 1:
            stored = hidden_state.split(self._output_dim_list[:-1], dim=-1)

        # a single layer fed with the actor's previous states runs un-cut and is sliced afterwards (:690-701)
        slice_single = n == 1 and hidden_state is not None and is_prev_hidden_state
        x = key
        for i, block in enumerate(blocks):
            last = i == n - 1
            if i > 0 and stored is not None:
                if is_prev_hidden_state:
                    x = x[:, -k_len:]
                x = torch.cat([stored[i - 1], x], dim=1)
            x, w = block(x, seq_q_len, cut_query=(cut_query and not slice_single) if last else False, **common)
            weights.append(w)
            if not last:
                kept.append(x[:, -seq_q_len:])
        if n == 1:
            if slice_single and cut_query:
                x = x[:, -seq_q_len:]
            return x, torch.zeros(key.shape[0], seq_q_len, 1, device=key.device), weights
        return x, torch.cat(kept, dim=-1), weights
