"""torch-CPU fp32 restatement of the reference's SAC update with DISCRETE (and hybrid) action branches
(TEST INFRASTRUCTURE; see ``oracle/sac_oracle.py`` for the rules).  Groundwork for SURVEY §8f rank 4:
no CUDA path exists for it yet — ``SAC_Base`` raises for ``d_action_sizes`` — but the oracle is pinned
to the reference (``tests/golden/sac_disc*.npz``, ``tests/test_oracle_golden.py``) so that the kernels
can be written against it.

Covers ``/root/reference/algorithm/sac_base.py`` with ``d_action_sizes`` non-empty, ``discrete_dqn_like``
False, stock ``ModelQ`` / ``ModelPolicy`` (identity ``dense``, one ``LinearLayers`` per action branch):

* ``d_heads_forward``          <- ``ModelQ.forward`` / ``ModelPolicy.forward`` discrete heads (q.py:74-80,
                                  policy.py:152-160)
* ``JointCategorical``         <- ``JointOneHotCategorical``                (policy.py:47-84)
* ``SacHybridOracle.get_y``    <- ``_get_y`` discrete branch + continuous   (sac_base.py:1356-1464),
                                  ``get_dqn_like_d_y`` with ``discrete_dqn_like`` (sac_base.py:1193-1242)
* ``.train_q``                 <- ``_train_rep_q``                          (sac_base.py:1516-1603)
* ``.train_policy``            <- ``_train_policy`` incl. entropy penalty   (sac_base.py:1858-1911)
* ``.train_alpha``             <- ``_train_alpha`` (one Adam over both log alphas, :472, 1913-1949)
* ``.l_probs`` / ``.td_error`` <- ``get_l_probs`` (:1159-1189) / ``_get_td_error`` (:2182-2245)

Actions are the concatenation [one-hot per discrete branch ..., continuous ...]; ``mu_probs`` likewise
(per-branch probabilities, then per-dimension densities).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
from torch.nn import functional as F

from .sac_oracle import (SacBatch, SacHyper, SacNoise, SacOracle, init_linear, policy_forward, prod_prob,
                         q_forward, squash_log_prob, squash_prob, sum_log_prob, trunk_forward)


@dataclass
class HybridHyper(SacHyper):
    d_action_sizes: list = field(default_factory=list)
    # per-column target: ratio * log(size of the column's branch), sac_base.py:466 (a [d_sum] tensor)
    target_d_alpha: object = 0.98
    d_policy_entropy_penalty: float = 0.5
    d_depth: int = 3
    discrete_dqn_like: bool = False  # sac_base.py:1193-1242: double-DQN target, no discrete policy / alpha loss

    @property
    def d_sum(self) -> int:
        return sum(self.d_action_sizes)

    @property
    def branches(self) -> int:
        return len(self.d_action_sizes)


def d_head_names(branch: int, depth: int):
    pre = f'd_dense_list.{branch}.dense.'
    names = []
    for layer in range(depth):
        names += [f'{pre}{2 * layer}.linear.weight', f'{pre}{2 * layer}.linear.bias']
    return names + [f'{pre}{2 * depth}.weight', f'{pre}{2 * depth}.bias']


def init_d_heads(state_size: int, hp: HybridHyper, gen: torch.Generator) -> dict:
    p = {}
    for k, size in enumerate(hp.d_action_sizes):
        names, d_in = d_head_names(k, hp.d_depth), state_size
        for layer in range(hp.d_depth):
            p[names[2 * layer]], p[names[2 * layer + 1]] = init_linear(hp.hidden, d_in, gen)
            d_in = hp.hidden
        p[names[-2]], p[names[-1]] = init_linear(size, d_in, gen)
    return p


def d_heads_forward(p: dict, hp: HybridHyper, state: torch.Tensor) -> torch.Tensor:
    """One ``LinearLayers(state -> hidden x d_depth -> d_action_size_k)`` per branch, concatenated."""
    outs = []
    for k in range(hp.branches):
        pre = f'd_dense_list.{k}.dense.'
        x = state
        for layer in range(hp.d_depth):
            w, b = p[f'{pre}{2 * layer}.linear.weight'], p[f'{pre}{2 * layer}.linear.bias']
            h = F.gelu(F.linear(x, w, b))
            x = h + x if w.shape[0] == w.shape[1] else h
        outs.append(F.linear(x, p[f'{pre}{2 * hp.d_depth}.weight'], p[f'{pre}{2 * hp.d_depth}.bias']))
    return torch.cat(outs, dim=-1)


class JointCategorical:
    """policy.py:47-84 over the concatenated logits of the branches, with torch.distributions'
    own arithmetic: normalised logits l - logsumexp(l), probs = softmax of those, log_prob = the
    normalised logit at argmax(value) (OneHotCategorical.log_prob), entropy = -sum(clamp(l) * p)."""

    def __init__(self, logits: torch.Tensor, sizes: list):
        self.sizes = list(sizes)
        self.logits_n = [l - l.logsumexp(dim=-1, keepdim=True) for l in logits.split(self.sizes, dim=-1)]
        self.p = [torch.softmax(l, dim=-1) for l in self.logits_n]

    @property
    def probs(self) -> torch.Tensor:
        return torch.cat(self.p, dim=-1)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:  # [..., branches]
        out = []
        for v, l in zip(value.split(self.sizes, dim=-1), self.logits_n):
            idx = v.max(-1)[1]
            out.append(l.gather(-1, idx.unsqueeze(-1)).squeeze(-1))
        return torch.stack(out, dim=-1)

    def entropy(self) -> torch.Tensor:  # [..., branches]
        out = []
        for l, p in zip(self.logits_n, self.p):
            out.append(-(torch.clamp(l, min=torch.finfo(l.dtype).min) * p).sum(-1))
        return torch.stack(out, dim=-1)


class SacHybridOracle(SacOracle):
    def __init__(self, hp: HybridHyper, seed: int = 0, dtype: torch.dtype = torch.float32):
        gen = torch.Generator().manual_seed(seed + 104729)
        self.log_d_alpha = torch.tensor(hp.init_log_alpha, dtype=dtype)
        cast = lambda d: {k: v.to(dtype) for k, v in d.items()}
        self._dq = [cast(init_d_heads(hp.state_size, hp, gen)) for _ in range(hp.ensemble_q_num)]
        self._dqt = [cast(init_d_heads(hp.state_size, hp, gen)) for _ in range(hp.ensemble_q_num)]
        self._dpi = cast(init_d_heads(hp.state_size, hp, gen))
        super().__init__(hp, seed, dtype)

    # ---- parameters: one dict per net holding both kinds of heads, keyed like the reference's state_dict
    def _wire(self):
        hp = self.hp
        if hasattr(self, '_dq'):
            for q, dq in zip(self.q, self._dq):
                q.update(dq)
            for q, dq in zip(self.q_target, self._dqt):
                q.update(dq)
            self.policy.update(self._dpi)
            del self._dq, self._dqt, self._dpi
        if hp.action_size == 0:  # drop the continuous heads of a discrete-only run
            for net in self.q + self.q_target:
                for k in [k for k in net if k.startswith('c_dense')]:
                    del net[k]
            for k in [k for k in self.policy if not k.startswith('d_dense_list')]:
                del self.policy[k]
        super()._wire()
        self.log_d_alpha = self.log_d_alpha.detach().clone().requires_grad_(True)
        # sac_base.py:472: ONE Adam over [log_d_alpha, log_c_alpha]; a parameter without a gradient is skipped
        self.opt_alpha = torch.optim.Adam([self.log_d_alpha, self.log_c_alpha], lr=hp.learning_rate)

    def load_params(self, q, q_target, policy, log_c_alpha, log_d_alpha=None):
        if log_d_alpha is not None:
            self.log_d_alpha = torch.as_tensor(log_d_alpha).detach().to(self.dtype).reshape(()).clone()
        super().load_params(q, q_target, policy, log_c_alpha)

    # ---- forward helpers
    def _q(self, p, state, c_action):
        hp = self.hp
        d = d_heads_forward(p, hp, state) if hp.branches else None
        c = q_forward(p, hp.q_depth, state, c_action) if hp.action_size else None
        return d, c

    def _pi(self, state):
        hp = self.hp
        d = JointCategorical(d_heads_forward(self.policy, hp, state), hp.d_action_sizes) if hp.branches else None
        c = policy_forward(self.policy, hp.policy_depth, state) if hp.action_size else None
        return d, c

    # ---- sac_base.py:1297-1466
    @torch.no_grad()
    def get_y(self, last_masks, padding_masks, nx_states, n_actions, rewards, dones, mu_probs, eps, perms=None):
        """``perms`` (DQN-like only): the two ``torch.randperm(E)`` draws of sac_base.py:1366 and :1377.  The
        reference permutes the TARGET stack and the ONLINE stack independently before pairing member i's
        argmax with member i's target value, so with E > 1 its result depends on those draws; the
        fixtures record them, identity permutations otherwise."""
        hp = self.hp
        D = hp.d_sum
        nx_actions = torch.cat([n_actions, torch.zeros_like(n_actions[:, :1])], dim=1)
        d_pi, c_pi = self._pi(nx_states)
        sampled = None
        if hp.action_size:
            loc, scale = c_pi
            sampled = loc + eps * scale
        qs = [self._q(q, nx_states, torch.tanh(sampled) if sampled is not None else None) for q in self.q_target]
        d_y = c_y = None
        if hp.branches and hp.discrete_dqn_like:
            # double DQN on the last solid step: argmax of the ONLINE critics, value of the TARGET critics,
            # min over the ensemble, n-step discounted rewards in front (sac_base.py:1193-1242, 1368-1383)
            solid = torch.logical_or(last_masks, padding_masks)
            last_idx = solid.shape[1] - torch.flip(solid.to(torch.uint8), dims=[1]).argmin(1) - 1   # operators.py:7-9
            rows = torch.arange(solid.shape[0])
            nxt_states = nx_states[:, 1:]
            nxt_c = torch.tanh(sampled[:, 1:]) if sampled is not None else None
            eval_q = torch.stack([self._q(q, nxt_states, nxt_c)[0] for q in self.q])[:, rows, last_idx]
            tgt_q = torch.stack([q[0][:, 1:] for q in qs])[:, rows, last_idx]
            if perms is not None:
                tgt_q, eval_q = tgt_q[torch.as_tensor(perms[0])], eval_q[torch.as_tensor(perms[1])]
            masks = [F.one_hot(part.argmax(dim=-1), size) for part, size in
                     zip(eval_q.split(hp.d_action_sizes, dim=-1), hp.d_action_sizes)]
            chosen = torch.sum(tgt_q * torch.cat(masks, dim=-1), dim=-1, keepdim=True) / hp.branches
            next_q = chosen.min(dim=0)[0]
            done = dones[rows, last_idx].unsqueeze(-1)
            g = torch.sum(self.gamma_ratio * rewards, dim=-1, keepdim=True)
            d_y = g + torch.pow(torch.tensor(hp.gamma, dtype=self.dtype), last_idx.unsqueeze(-1) + 1) * next_q * ~done
        elif hp.branches:
            d_alpha = torch.exp(self.log_d_alpha)
            mean_q = torch.stack([q[0] for q in qs]).mean(dim=0)          # mean over the ensemble (:1384-1385)
            probs = d_pi.probs
            v = torch.sum(probs * (mean_q - d_alpha * torch.log(probs.clamp(min=1e-8))), dim=-1) / hp.branches
            pi = mu = None
            if hp.use_n_step_is:
                mu = mu_probs[..., :D] * n_actions[..., :D]
                mu = torch.where(mu == 0., torch.ones_like(mu), mu).prod(-1)
                pi = torch.exp(d_pi.log_prob(nx_actions[..., :D]).sum(-1))[:, :-1]
            d_y = self.v_trace(last_masks, padding_masks, rewards, dones, mu, pi, v[:, :-1], v[:, 1:])
        if hp.action_size:
            alpha = torch.exp(self.log_c_alpha)
            loc, scale = c_pi
            logp = sum_log_prob(squash_log_prob(loc, scale, sampled))
            min_q = torch.stack([q[1] for q in qs]).min(dim=0)[0].squeeze(-1)
            v = min_q - alpha * logp
            pi = mu = None
            if hp.use_n_step_is:
                stored = torch.atanh(torch.clamp(nx_actions[..., D:], -0.999, 0.999))
                pi = prod_prob(squash_prob(loc, scale, stored)[:, :-1])
                mu = prod_prob(mu_probs[..., D:])
            c_y = self.v_trace(last_masks, padding_masks, rewards, dones, mu, pi, v[:, :-1], v[:, 1:])
        return d_y, c_y

    # ---- sac_base.py:1516-1603
    def train_q(self, b: SacBatch, eps_y, perms=None):
        hp = self.hp; s0 = hp.burn_in_step; D = hp.d_sum
        state, action = b.states[:, s0], b.actions[:, s0]
        q_vals = [self._q(q, state, action[..., D:]) for q in self.q]
        d_y, c_y = self.get_y(b.last_masks[:, s0:], b.padding_masks[:, s0:], b.states[:, s0:], b.actions[:, s0:],
                              b.rewards[:, s0:], b.dones[:, s0:], b.mu_probs[:, s0:], eps_y, perms)
        losses = []
        for i, (d_q, c_q) in enumerate(q_vals):
            loss = torch.zeros((state.shape[0], 1), dtype=self.dtype)
            if hp.branches:
                q_single = torch.sum(action[..., :D] * d_q, dim=-1, keepdim=True) / hp.branches
                loss = loss + (q_single - d_y) ** 2
            if hp.action_size:
                if hp.clip_epsilon > 0:
                    with torch.no_grad():
                        tq = self._q(self.q_target[i], state, action[..., D:])[1]
                    clipped = tq + torch.clamp(c_q - tq, -hp.clip_epsilon, hp.clip_epsilon)
                    loss = loss + torch.maximum((clipped - c_y) ** 2, (c_q - c_y) ** 2)
                else:
                    loss = loss + loss + (c_q - c_y) ** 2  # `+= loss + mse` in the reference (:1562)
            if b.priority_is is not None:
                loss = loss * b.priority_is
            losses.append(torch.mean(loss))
        for opt in self.opt_q:
            opt.zero_grad()
        torch.stack(losses).sum().backward()
        grads = [{k: t.grad.clone() for k, t in q.items()} for q in self.q]
        for opt in self.opt_q:
            opt.step()
        return dict(d_y=d_y, y=c_y, loss_q=[l.detach() for l in losses], grad_q=grads)

    # ---- sac_base.py:1841-1911
    def train_policy(self, b: SacBatch, eps_pi):
        hp = self.hp; s0 = hp.burn_in_step; D = hp.d_sum
        state, action = b.states[:, s0], b.actions[:, s0]
        d_pi, c_pi = self._pi(state)
        loss = torch.zeros((state.shape[0], 1), dtype=self.dtype)
        with torch.no_grad():
            d_alpha, c_alpha = torch.exp(self.log_d_alpha), torch.exp(self.log_c_alpha)
        if hp.branches and not hp.discrete_dqn_like:
            probs = d_pi.probs
            with torch.no_grad():
                mean_q = torch.stack([self._q(q, state, action[..., D:])[0] for q in self.q]).mean(dim=0)
            inner = d_alpha * torch.log(probs.clamp(min=1e-8)) - mean_q
            loss = loss + torch.sum(probs * inner, dim=1, keepdim=True) / hp.branches
            mu = b.mu_probs[:, s0, :D]
            mu_entropy = -torch.sum(mu * torch.log(mu.clamp(min=1e-8)), dim=-1) / hp.branches
            pi_entropy = d_pi.entropy().sum(-1) / hp.branches
            loss = loss + hp.d_policy_entropy_penalty * (torch.pow(mu_entropy - pi_entropy, 2.) / 2.).unsqueeze(-1)
        if hp.action_size:
            loc, scale = c_pi
            sampled = loc + eps_pi * scale
            qs = [self._q(q, state, torch.tanh(sampled))[1] for q in self.q]
            logp = sum_log_prob(squash_log_prob(loc, scale, sampled), keepdim=True)
            loss = loss + c_alpha * logp - torch.stack(qs).min(dim=0)[0]
        total = torch.mean(loss)
        out = dict(loss_policy=total.detach())
        if (hp.branches and not hp.discrete_dqn_like) or hp.action_size:  # sac_base.py:1904-1907
            self.opt_policy.zero_grad()
            total.backward(inputs=list(self.policy.values()))
            out['grad_policy'] = {k: (t.grad.clone() if t.grad is not None else torch.zeros_like(t))
                                  for k, t in self.policy.items()}
            self.opt_policy.step()
        if hp.branches:
            out['d_entropy'] = torch.mean(d_pi.entropy().sum(-1) / hp.branches).detach()
        return out

    # ---- sac_base.py:1913-1949
    def train_alpha(self, b: SacBatch, eps_alpha):
        hp = self.hp; s0 = hp.burn_in_step
        state = b.states[:, s0]
        with torch.no_grad():
            d_pi, c_pi = self._pi(state)
        loss = torch.zeros((state.shape[0], 1), dtype=self.dtype)
        d_on = hp.branches and not hp.discrete_dqn_like
        if d_on:
            probs = d_pi.probs
            inner = self.log_d_alpha * (-torch.log(probs.clamp(min=1e-8)) - hp.target_d_alpha)
            loss = loss + torch.sum(probs * inner, dim=1, keepdim=True) / hp.branches
        if hp.action_size:
            loc, scale = c_pi
            sampled = loc + eps_alpha * scale
            lp = squash_log_prob(loc, scale, sampled)
            valid = torch.sum(lp != torch.inf, dim=-1, keepdim=True)
            loss = loss + self.log_c_alpha * (-sum_log_prob(lp, keepdim=True) - hp.target_c_alpha * -valid)
        total = torch.mean(loss)
        self.opt_alpha.zero_grad()
        used = ([self.log_d_alpha] if d_on else []) + ([self.log_c_alpha] if hp.action_size else [])
        total.backward(inputs=used)
        out = dict(loss_alpha=total.detach())
        if d_on:
            out['grad_log_d_alpha'] = self.log_d_alpha.grad.clone()
        if hp.action_size:
            out['grad_log_alpha'] = self.log_c_alpha.grad.clone()
        self.opt_alpha.step()
        return out

    # ---- sac_base.py:1159-1189
    @torch.no_grad()
    def l_probs(self, states, actions):
        hp = self.hp; D = hp.d_sum
        d_pi, c_pi = self._pi(states)
        parts = []
        if hp.branches:
            parts.append(d_pi.probs)
        if hp.action_size:
            loc, scale = c_pi
            parts.append(squash_prob(loc, scale, torch.atanh(torch.clamp(actions[..., D:], -0.999, 0.999))))
        return torch.cat(parts, dim=-1)

    # ---- sac_base.py:2182-2245
    @torch.no_grad()
    def td_error(self, b: SacBatch, pi_probs, eps_td, perms=None):
        hp = self.hp; s0 = hp.burn_in_step; D = hp.d_sum
        state, action = b.states[:, s0], b.actions[:, s0]
        q_vals = [self._q(q, state, action[..., D:]) for q in self.q]
        d_y, c_y = self.get_y(b.last_masks[:, s0:], b.padding_masks[:, s0:], b.states[:, s0:], b.actions[:, s0:],
                              b.rewards[:, s0:], b.dones[:, s0:],
                              pi_probs[:, s0:] if pi_probs is not None else None, eps_td, perms)
        errs = []
        for d_q, c_q in q_vals:
            e = torch.zeros((state.shape[0], 1), dtype=self.dtype)
            if hp.branches:
                e = e + torch.abs(torch.sum(action[..., :D] * d_q, dim=-1, keepdim=True) / hp.branches - d_y)
            if hp.action_size:
                e = e + torch.abs(c_q - c_y)
            errs.append(e)
        return torch.mean(torch.cat(errs, dim=-1), dim=-1, keepdim=True), (d_y, c_y)

    def step(self, b: SacBatch, noise: SacNoise, perms=None) -> dict:
        """``perms``: [(target, online) for _train_rep_q's _get_y, (target, online) for _get_td_error's] — the
        reference's randperm draws of a DQN-like discrete-only step, in call order."""
        hp = self.hp
        if self.global_step % hp.update_target_per_step == 0:
            self.polyak(hp.tau)
        out = self.train_q(b, noise.eps_y, None if perms is None else perms[0])
        out.update(self.train_policy(b, noise.eps_pi))
        if hp.use_auto_alpha and ((hp.branches and not hp.discrete_dqn_like) or hp.action_size):  # sac_base.py:2115
            out.update(self.train_alpha(b, noise.eps_alpha))
        pi_probs = None
        if hp.use_n_step_is:
            pi_probs = self.l_probs(b.states[:, :-1], b.actions)
            out['pi_probs'] = pi_probs
        if hp.use_priority:
            out['td_error'], (out['d_y_td'], out['y_td']) = self.td_error(b, pi_probs, noise.eps_td,
                                                                          None if perms is None else perms[1])
        self.global_step += 1
        return out

    def snapshot(self) -> dict:
        d = super().snapshot()
        d['log_d_alpha'] = self.log_d_alpha.detach().clone().numpy()
        return d
