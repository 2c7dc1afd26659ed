"""torch-CPU fp32 restatement of the reference's SAC update with a recurrent representation
(TEST INFRASTRUCTURE; see ``oracle/sac_oracle.py`` for the rules).

Covers the ``seq_encoder=SEQ_ENCODER.RNN`` flow of ``/root/reference/algorithm/sac_base.py`` for a
plugin representation of the form of ``envs/test/nn_rnn.py:6-21``: ``state, hn = GRU(cat[obs,
pre_action], h0)`` with the stock multi-layer ``GRU`` wrapper (nn_models/layers/seq_layers.py:14-114,
no padding mask handed to it):

* ``gru_forward``             <- ``GRU.forward`` / ``torch.nn.GRU`` cell (seq_layers.py:41-113)
* ``SacRepOracle.rep``        <- ``get_l_states`` + ``get_bnx_data``        (sac_base.py:1090-1146)
* ``SacRepOracle.step``       <- ``_train`` (:2057-2116) + tail of ``train`` (:2558-2605):
    online and target representation of the whole window, critic loss back-propagated through
    the online representation (through every burn-in step), Adam on critics and representation,
    representation re-evaluated with the new weights, policy / alpha on the new state,
    ``get_l_probs`` on the new states, ``_get_td_error`` with the new state and the TARGET
    states, next hidden states for the ``pre_seq_hidden_state`` write-back.

Pinned by ``tests/golden/sac_rnn*.npz`` (generated from the real reference by
``oracle/gen_golden.py``), see ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import dataclasses
import math
from dataclasses import dataclass

import torch

from .sac_oracle import SacBatch, SacHyper, SacNoise, SacOracle, q_forward


def gru_param_names(layers: int, prefix: str = 'rnn.'):
    names = []
    for layer in range(layers):
        names += [f'{prefix}_grus.{layer}.weight_ih_l0', f'{prefix}_grus.{layer}.weight_hh_l0',
                  f'{prefix}_grus.{layer}.bias_ih_l0', f'{prefix}_grus.{layer}.bias_hh_l0']
    return names


def init_gru(in_dim: int, hidden: int, layers: int, gen: torch.Generator, prefix: str = 'rnn.'):
    """torch.nn.GRU.reset_parameters: every tensor U(-1/sqrt(H), 1/sqrt(H))."""
    bound = 1.0 / math.sqrt(hidden)
    p, k = {}, in_dim
    for layer in range(layers):
        for name, shape in (('weight_ih_l0', (3 * hidden, k)), ('weight_hh_l0', (3 * hidden, hidden)),
                            ('bias_ih_l0', (3 * hidden,)), ('bias_hh_l0', (3 * hidden,))):
            p[f'{prefix}_grus.{layer}.{name}'] = ((torch.rand(*shape, generator=gen) * 2 - 1) * bound).float()
        k = hidden
    return p


def gru_forward(p: dict, layers: int, x: torch.Tensor, h0: torch.Tensor | None, prefix: str = 'rnn.'):
    """x [B, L, in], h0 [B, layers, H] -> (output of the top layer [B, L, H], every layer's output
    [B, L, layers, H]).  Gate order r, z, n; h' = (h - n) * z + n as ATen's GRU cell evaluates it."""
    per_layer = []
    for layer in range(layers):
        w_ih, w_hh = p[f'{prefix}_grus.{layer}.weight_ih_l0'], p[f'{prefix}_grus.{layer}.weight_hh_l0']
        b_ih, b_hh = p[f'{prefix}_grus.{layer}.bias_ih_l0'], p[f'{prefix}_grus.{layer}.bias_hh_l0']
        H = w_hh.shape[1]
        h = x.new_zeros(x.shape[0], H) if h0 is None else h0[:, layer]
        outs = []
        for t in range(x.shape[1]):
            gi = torch.nn.functional.linear(x[:, t], w_ih, b_ih)
            gh = torch.nn.functional.linear(h, w_hh, b_hh)
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (h - n) * z + n
            outs.append(h)
        x = torch.stack(outs, dim=1)
        per_layer.append(x)
    return x, torch.stack(per_layer, dim=2)


@dataclass
class SacRepBatch:
    """``_train``'s arguments for a recurrent run (sac_base.py:2027-2050)."""
    obs: torch.Tensor             # [B, L, So]            bnx_obses_list[0]
    hidden0: torch.Tensor         # [B, layers, H]        bnx_pre_seq_hidden_states[:, 0]
    actions: torch.Tensor         # [B, L-1, A]
    rewards: torch.Tensor         # [B, L-1]
    dones: torch.Tensor           # [B, L-1] bool
    mu_probs: torch.Tensor        # [B, L-1, A]
    last_masks: torch.Tensor      # [B, L-1] bool
    padding_masks: torch.Tensor   # [B, L-1] bool
    priority_is: torch.Tensor | None = None

    def with_states(self, states: torch.Tensor) -> SacBatch:
        return SacBatch(states=states, actions=self.actions, rewards=self.rewards, dones=self.dones,
                        mu_probs=self.mu_probs, last_masks=self.last_masks, padding_masks=self.padding_masks,
                        priority_is=self.priority_is)

    def to(self, dtype: torch.dtype) -> 'SacRepBatch':
        f = lambda t: None if t is None else (t.to(dtype) if t.is_floating_point() else t)
        return SacRepBatch(**{k: f(getattr(self, k)) for k in self.__dataclass_fields__})


class SacRepOracle(SacOracle):
    def __init__(self, hp: SacHyper, obs_size: int, rep_layers: int, seed: int = 0,
                 dtype: torch.dtype = torch.float32):
        """``hp.state_size`` is the GRU width (the state IS the top layer's output)."""
        self.obs_size, self.rep_layers = obs_size, rep_layers
        gen = torch.Generator().manual_seed(seed + 7919)
        in_dim = obs_size + hp.action_size
        self.rep = {k: v.to(dtype) for k, v in init_gru(in_dim, hp.state_size, rep_layers, gen).items()}
        self.rep_target = {k: v.clone() for k, v in self.rep.items()}
        super().__init__(hp, seed, dtype)

    def _wire(self):
        super()._wire()
        for t in self.rep.values():
            t.requires_grad_(True)
        self.opt_rep = torch.optim.Adam(list(self.rep.values()), lr=self.hp.learning_rate)

    def load_rep(self, rep, rep_target):
        to_t = lambda d: {k: torch.as_tensor(v).detach().to(self.dtype).clone() for k, v in d.items()}
        self.rep, self.rep_target = to_t(rep), to_t(rep_target)
        self._wire()

    def copy_state_from(self, other: 'SacRepOracle', adam: bool = False) -> None:
        self.rep = {k: v.detach().to(self.dtype).clone() for k, v in other.rep.items()}
        self.rep_target = {k: v.detach().to(self.dtype).clone() for k, v in other.rep_target.items()}
        super().copy_state_from(other, adam)
        if adam:
            self._copy_adam(self.opt_rep, self.rep, other.opt_rep, other.rep)

    @torch.no_grad()
    def polyak(self, tau: float):  # sac_base.py:745-764: representation first, then the critics
        for k in self.rep_target:
            self.rep_target[k].copy_(self.rep_target[k] * (1. - tau) + self.rep[k] * tau)
        super().polyak(tau)

    def encode(self, b: SacRepBatch, target: bool = False):
        """get_bnx_data + get_l_states (sac_base.py:1090-1146): pre_action[t] = action[t-1], zeros first."""
        pre = torch.cat([torch.zeros_like(b.actions[:, :1]), b.actions], dim=1)
        x = torch.cat([b.obs, pre], dim=-1)
        return gru_forward(self.rep_target if target else self.rep, self.rep_layers, x, b.hidden0)

    def step(self, b: SacRepBatch, noise: SacNoise) -> dict:
        hp = self.hp; s0 = hp.burn_in_step
        if self.global_step % hp.update_target_per_step == 0:
            self.polyak(hp.tau)
        states, _ = self.encode(b)
        with torch.no_grad():
            target_states, _ = self.encode(b, target=True)
        self.opt_rep.zero_grad()
        out = self.train_q(b.with_states(states), noise.eps_y)  # backward reaches self.rep
        out['grad_rep'] = {k: t.grad.clone() for k, t in self.rep.items()}
        self.opt_rep.step()
        with torch.no_grad():
            states2, hn2 = self.encode(b)
        b2 = b.with_states(states2)
        out.update(self.train_policy(b2, noise.eps_pi))
        if hp.use_auto_alpha:
            out.update(self.train_alpha(b2, noise.eps_alpha))
        pi_probs = None
        if hp.use_n_step_is:
            pi_probs = self.l_probs(states2[:, :-1], b.actions)
            out['pi_probs'] = pi_probs
        if hp.use_priority:
            with torch.no_grad():
                q_vals = [q_forward(q, hp.q_depth, states2[:, s0], b.actions[:, s0]) for q in self.q]
                y = self.get_y(b.last_masks[:, s0:], b.padding_masks[:, s0:], target_states[:, s0:],
                               b.actions[:, s0:], b.rewards[:, s0:], b.dones[:, s0:],
                               pi_probs[:, s0:] if pi_probs is not None else None, noise.eps_td)
                err = torch.cat([torch.abs(qv - y) for qv in q_vals], dim=-1)
            out['td_error'], out['y_td'] = torch.mean(err, dim=-1, keepdim=True), y
        out.update(states=states.detach(), target_states=target_states, states_post=states2,
                   next_hidden=hn2[:, :-1])
        self.global_step += 1
        return out

    def snapshot(self) -> dict:
        d = super().snapshot()
        for k, t in self.rep.items():
            d[f'rep.{k}'] = t.detach().clone().numpy()
        for k, t in self.rep_target.items():
            d[f'rept.{k}'] = t.detach().clone().numpy()
        return d


__all__ = ['gru_forward', 'gru_param_names', 'init_gru', 'SacRepBatch', 'SacRepOracle', 'dataclasses']
