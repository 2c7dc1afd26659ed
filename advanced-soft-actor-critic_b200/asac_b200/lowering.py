"""Lowering of stock plugin networks onto the flat parameter layout of the CUDA kernels.

A plugin ``ModelQ`` / ``ModelPolicy`` is *stock* when it is the reference's default topology
(algorithm/nn_models/q.py:34-91, policy.py:116-174): identity ``dense`` / ``c_state_dense`` /
``c_action_dense`` and a ``c_dense`` made of equal-width ``ResBlock(Linear + GELU)`` blocks.
Every ``ModelQ`` / ``ModelPolicy`` under the reference's ``envs/`` is of that form (SURVEY.md §2
row 4).  For such nets the learner keeps ONE flat fp32 buffer per net family and re-points the
``nn.Parameter`` storage of the modules at views of it, so the torch modules (actor side,
``state_dict`` / checkpoints) and the kernels (learner side) share the same HBM bytes.

Flat layout per net (include/asac_b200.h):
    W0[H, in] b0[H]  W1[H, H] b1[H] ...  Whead[O, H] bhead[O]
For the policy the head is [mean rows; logstd rows].
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from . import nn_models as m


class NotStockNetwork(NotImplementedError):
    """Raised when a plugin network cannot be lowered onto the fused kernels."""


@dataclass(frozen=True)
class NetShape:
    in_dim: int
    hidden: int
    depth: int
    out_dim: int

    @property
    def count(self) -> int:
        n, k = 0, self.in_dim
        for _ in range(self.depth):
            n += self.hidden * k + self.hidden
            k = self.hidden
        return n + self.out_dim * self.hidden + self.out_dim

    @property
    def stride(self) -> int:
        return (self.count + 3) // 4 * 4


def _no_params(mod: nn.Module) -> bool:
    return sum(p.numel() for p in mod.parameters()) == 0


def _trunk_blocks(layers: m.LinearLayers, what: str) -> tuple[list[m.ResBlock], nn.Linear | None]:
    blocks, head = [], None
    for mod in layers.dense:
        if isinstance(mod, m.ResBlock):
            if head is not None:
                raise NotStockNetwork(f'{what}: block after the output layer')
            act = mod.act
            if not isinstance(act, nn.GELU) or getattr(act, 'approximate', 'none') != 'none':
                raise NotStockNetwork(f'{what}: activation {type(act).__name__} is not exact GELU')
            lin = mod.linear
            if mod.residual != (lin.in_features == lin.out_features):
                raise NotStockNetwork(f'{what}: residual flag does not follow in == out')
            blocks.append(mod)
        elif isinstance(mod, nn.Dropout):
            if mod.p != 0:
                raise NotStockNetwork(f'{what}: dropout p={mod.p} != 0')
        elif isinstance(mod, nn.Linear):
            head = mod
        else:
            raise NotStockNetwork(f'{what}: unexpected module {type(mod).__name__}')
    if not blocks:
        raise NotStockNetwork(f'{what}: needs at least one dense block')
    widths = {b.linear.out_features for b in blocks}
    if len(widths) != 1:
        raise NotStockNetwork(f'{what}: unequal block widths {sorted(widths)}')
    return blocks, head


def analyze_d_heads(mod: nn.Module, what: str) -> tuple[list[NetShape], list[list[nn.Parameter]]]:
    """The discrete-action heads of a stock ModelQ / ModelPolicy (q.py:60-64, policy.py:143-147): one
    ``LinearLayers(state -> d_dense_n x d_dense_depth -> d_action_size_k)`` per branch.
    -> (shape per branch, parameters per branch in flat order)."""
    shapes, params = [], []
    for k, layers in enumerate(mod.d_dense_list):
        blocks, head = _trunk_blocks(layers, f'{what}.d_dense_list[{k}]')
        if head is None or head.out_features != mod.d_action_sizes[k]:
            raise NotStockNetwork(f'{what}.d_dense_list[{k}] has no output layer of size {mod.d_action_sizes[k]}')
        shapes.append(NetShape(blocks[0].linear.in_features, blocks[0].linear.out_features, len(blocks), head.out_features))
        params.append([p for b in blocks for p in (b.linear.weight, b.linear.bias)] + [head.weight, head.bias])
    if len({(s.in_dim, s.hidden, s.depth) for s in shapes}) != 1:
        raise NotStockNetwork(f'{what}: discrete heads of unequal width / depth')
    return shapes, params


def analyze_q(q: nn.Module) -> tuple[NetShape | None, list[nn.Parameter]]:
    """-> (shape, parameters in flat order) of the continuous critic (None, [] without continuous actions).
    Discrete heads are lowered separately (``analyze_d_heads``).  Raises NotStockNetwork otherwise."""
    if not isinstance(q, m.ModelQ) or type(q).forward is not m.ModelQ.forward:
        raise NotStockNetwork('ModelQ overrides forward')
    n_d = sum(p.numel() for p in q.d_dense_list.parameters()) if q.d_action_sizes else 0
    if not _no_params(q.dense):
        raise NotStockNetwork('ModelQ with a shared dense trunk')
    if not q.c_action_size:
        if sum(p.numel() for p in q.parameters()) != n_d:
            raise NotStockNetwork('ModelQ has parameters outside d_dense_list')
        return None, []
    if not (_no_params(q.c_state_dense) and _no_params(q.c_action_dense)):
        raise NotStockNetwork('ModelQ with dense / c_state_dense / c_action_dense layers')
    blocks, head = _trunk_blocks(q.c_dense, 'ModelQ.c_dense')
    if head is None or head.out_features != 1:
        raise NotStockNetwork('ModelQ.c_dense has no scalar output layer')
    shape = NetShape(blocks[0].linear.in_features, blocks[0].linear.out_features, len(blocks), 1)
    params = [p for b in blocks for p in (b.linear.weight, b.linear.bias)] + [head.weight, head.bias]
    if sum(p.numel() for p in q.parameters()) != shape.count + n_d:
        raise NotStockNetwork('ModelQ has parameters outside c_dense / d_dense_list')
    return shape, params


def analyze_policy(pi: nn.Module) -> tuple[NetShape | None, list[nn.Parameter]]:
    if not isinstance(pi, m.ModelPolicy) or type(pi).forward is not m.ModelPolicy.forward:
        raise NotStockNetwork('ModelPolicy overrides forward')
    n_d = sum(p.numel() for p in pi.d_dense_list.parameters()) if pi.d_action_sizes else 0
    if not _no_params(pi.dense):
        raise NotStockNetwork('ModelPolicy with a shared dense trunk')
    if not pi.c_action_size:
        if sum(p.numel() for p in pi.parameters()) != n_d:
            raise NotStockNetwork('ModelPolicy has parameters outside d_dense_list')
        return None, []
    blocks, head = _trunk_blocks(pi.c_dense, 'ModelPolicy.c_dense')
    if head is not None:
        raise NotStockNetwork('ModelPolicy.c_dense has an output layer')
    heads = []
    for name in ('mean_dense', 'logstd_dense'):
        mods = list(getattr(pi, name).dense)
        if len(mods) != 1 or not isinstance(mods[0], nn.Linear):
            raise NotStockNetwork(f'ModelPolicy.{name} is not a single Linear')
        heads.append(mods[0])
    A = pi.c_action_size
    shape = NetShape(blocks[0].linear.in_features, blocks[0].linear.out_features, len(blocks), 2 * A)
    # flat order: trunk, then head weight rows [mean; logstd], then head bias [mean; logstd]
    params = [p for b in blocks for p in (b.linear.weight, b.linear.bias)]
    params += [heads[0].weight, heads[1].weight, heads[0].bias, heads[1].bias]
    if sum(p.numel() for p in pi.parameters()) != shape.count + n_d:
        raise NotStockNetwork('ModelPolicy has parameters outside c_dense / mean_dense / logstd_dense / d_dense_list')
    return shape, params


def bind_parameters(params: list[nn.Parameter], flat: torch.Tensor) -> None:
    """Copies each parameter into consecutive slices of ``flat`` and makes the parameter a view
    of that slice (shared storage from then on)."""
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            view = flat[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            off += n
    if off > flat.numel():
        raise AssertionError('flat buffer too small')


def flat_from_state_dict(shape: NetShape, state: dict, policy: bool, prefix: str = '') -> torch.Tensor:
    """Builds the flat fp32 vector of one stock net from a reference-format ``state_dict``
    (keys ``c_dense.dense.{2l}.linear.{weight,bias}``, ...).  CPU helper for tests / import."""
    parts = []
    for layer in range(shape.depth):
        parts += [state[f'{prefix}c_dense.dense.{2 * layer}.linear.weight'],
                  state[f'{prefix}c_dense.dense.{2 * layer}.linear.bias']]
    if policy:
        parts += [state[f'{prefix}mean_dense.dense.0.weight'], state[f'{prefix}logstd_dense.dense.0.weight'],
                  state[f'{prefix}mean_dense.dense.0.bias'], state[f'{prefix}logstd_dense.dense.0.bias']]
    else:
        parts += [state[f'{prefix}c_dense.dense.{2 * shape.depth}.weight'],
                  state[f'{prefix}c_dense.dense.{2 * shape.depth}.bias']]
    flat = torch.cat([torch.as_tensor(p, dtype=torch.float32).reshape(-1) for p in parts])
    if flat.numel() != shape.count:
        raise ValueError(f'state_dict holds {flat.numel()} values, the shape needs {shape.count}')
    out = torch.zeros(shape.stride, dtype=torch.float32)
    out[:shape.count] = flat
    return out


def state_dict_from_flat(shape: NetShape, flat: torch.Tensor, policy: bool) -> dict[str, torch.Tensor]:
    """Inverse of :func:`flat_from_state_dict` (returns views on ``flat``'s device)."""
    out, off, k = {}, 0, shape.in_dim
    H = shape.hidden

    def take(n, *view):
        nonlocal off
        t = flat[off:off + n].view(*view)
        off += n
        return t

    for layer in range(shape.depth):
        out[f'c_dense.dense.{2 * layer}.linear.weight'] = take(H * k, H, k)
        out[f'c_dense.dense.{2 * layer}.linear.bias'] = take(H, H)
        k = H
    if policy:
        A = shape.out_dim // 2
        out['mean_dense.dense.0.weight'] = take(A * H, A, H)
        out['logstd_dense.dense.0.weight'] = take(A * H, A, H)
        out['mean_dense.dense.0.bias'] = take(A, A)
        out['logstd_dense.dense.0.bias'] = take(A, A)
    else:
        out[f'c_dense.dense.{2 * shape.depth}.weight'] = take(H, 1, H)
        out[f'c_dense.dense.{2 * shape.depth}.bias'] = take(1, 1)
    return out


# --------------------------------------------------------------------------------------- representation
@dataclass(frozen=True)
class GruShape:
    """Stock multi-layer GRU over cat[obs, pre_action] (include/asac_b200.h, AsacGruShape)."""
    obs_size: int
    action_size: int
    hidden: int
    layers: int

    def in_dim(self, layer: int) -> int:
        return self.obs_size + self.action_size if layer == 0 else self.hidden

    @property
    def count(self) -> int:
        return sum(3 * self.hidden * (self.in_dim(l) + self.hidden + 2) for l in range(self.layers))

    @property
    def stride(self) -> int:
        return (self.count + 3) // 4 * 4


_GRU_TENSORS = ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0')


def gru_flat_from_state_dict(shape: GruShape, state: dict, prefix: str = 'rnn.') -> torch.Tensor:
    """Flat fp32 vector of a stock ``GRU`` wrapper from its ``state_dict`` (keys
    ``<prefix>_grus.<layer>.{weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0}``,
    seq_layers.py:27-39).  CPU helper for tests / import."""
    parts = [torch.as_tensor(state[f'{prefix}_grus.{l}.{name}'], dtype=torch.float32).reshape(-1)
             for l in range(shape.layers) for name in _GRU_TENSORS]
    flat = torch.cat(parts)
    if flat.numel() != shape.count:
        raise ValueError(f'state_dict holds {flat.numel()} values, the shape needs {shape.count}')
    out = torch.zeros(shape.stride, dtype=torch.float32)
    out[:shape.count] = flat
    return out


def gru_state_dict_from_flat(shape: GruShape, flat: torch.Tensor, prefix: str = 'rnn.') -> dict[str, torch.Tensor]:
    out, off, H = {}, 0, shape.hidden
    for l in range(shape.layers):
        k = shape.in_dim(l)
        for name, view in (('weight_ih_l0', (3 * H, k)), ('weight_hh_l0', (3 * H, H)), ('bias_ih_l0', (3 * H,)),
                           ('bias_hh_l0', (3 * H,))):
            n = 1
            for d in view:
                n *= d
            out[f'{prefix}_grus.{l}.{name}'] = flat[off:off + n].view(*view)
            off += n
    return out


def analyze_rep(rep: nn.Module, obs_shapes: list[tuple], action_size: int) -> tuple[GruShape, list[nn.Parameter]] | None:
    """-> None for a parameter-free ``ModelSimpleRep``; (shape, parameters in flat order) for a
    representation that is ONE stock ``GRU`` wrapper applied to ``cat[obs_list[0], pre_action]``
    with ``pre_seq_hidden_state[:, 0]`` as initial state (envs/test/nn_rnn.py:6-21).  The structure
    is checked here; that ``forward`` really computes that function is checked by the learner with a
    probe against the kernels (``SAC_Base._probe_rep``).  Raises NotStockNetwork otherwise."""
    n_params = sum(p.numel() for p in rep.parameters())
    if n_params == 0:
        if type(rep).forward is not m.ModelSimpleRep.forward:
            raise NotStockNetwork('parameter-free representation that is not ModelSimpleRep')
        return None
    grus = [mod for mod in rep.modules() if isinstance(mod, m.GRU)]
    if len(grus) != 1:
        raise NotStockNetwork('representation is not a single stock GRU (encoder representations beyond the '
                              'GRU-over-vector-observation form are outside the fused path)')
    gru = grus[0]
    cells = list(gru._grus)
    H = cells[0].hidden_size
    if len(obs_shapes) < 1 or len(obs_shapes[0]) != 1:
        raise NotStockNetwork('GRU representation needs a vector observation first')
    shape = GruShape(obs_shapes[0][0], action_size, H, len(cells))
    params = []
    for l, cell in enumerate(cells):
        if (cell.hidden_size != H or cell.input_size != shape.in_dim(l) or cell.num_layers != 1 or not cell.bias
                or cell.bidirectional or not cell.batch_first or cell.dropout != 0):
            raise NotStockNetwork(f'GRU layer {l} is not a plain single-layer batch-first GRU of width {H}')
        params += [cell.weight_ih_l0, cell.weight_hh_l0, cell.bias_ih_l0, cell.bias_hh_l0]
    if n_params != shape.count:
        raise NotStockNetwork('representation has parameters outside its GRU')
    return shape, params
