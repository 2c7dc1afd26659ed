"""torch-CPU fp32 restatement of the reference's SAC update (TEST INFRASTRUCTURE).

Continuous-action branch of ``/root/reference/algorithm/sac_base.py`` for stock
``ModelQ`` / ``ModelPolicy`` networks on already-encoded states (the vector-obs
``ModelSimpleRep`` case, representation.py:74-83, where state == concat(obs)):

* ``mlp_forward``            <- ``LinearLayers`` / ``ResBlock``   (linear_layers.py:24-119)
* ``q_forward``              <- ``ModelQ.forward``                (q.py:74-91)
* ``policy_forward``         <- ``ModelPolicy.forward``           (policy.py:152-174)
* ``squash_log_prob/_prob``  <- utils/operators.py:12-31
* ``SacOracle.v_trace``      <- ``SAC_Base._v_trace``             (sac_base.py:1244-1295)
* ``SacOracle.get_y``        <- ``SAC_Base._get_y`` cont. branch  (sac_base.py:1345-1351,1423-1464)
* ``SacOracle.train_q``      <- ``_train_rep_q``                  (sac_base.py:1516-1603)
* ``SacOracle.train_policy`` <- ``_train_policy``                 (sac_base.py:1882-1908)
* ``SacOracle.train_alpha``  <- ``_train_alpha``                  (sac_base.py:1930-1949)
* ``SacOracle.l_probs``      <- ``get_l_probs``                   (sac_base.py:1159-1189)
* ``SacOracle.td_error``     <- ``_get_td_error``                 (sac_base.py:2182-2245)
* ``SacOracle.polyak``       <- ``_update_target_variables``      (sac_base.py:745-764)
* ``SacOracle.step``         <- ``_train`` + tail of ``train``    (sac_base.py:2057-2126,2558-2583)

It uses autograd and ``torch.optim.Adam`` exactly like the reference does (the
third-party arithmetic is PyTorch itself, pinned only by the environment:
torch 2.11.0).  Gaussian draws are injected (``eps_*`` arguments) at the points
where the reference calls ``Normal.rsample`` / ``Normal.sample``.

Parity is pinned by ``tests/golden/sac_*.npz`` (generated from the real
reference by ``oracle/gen_golden.py``); see ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch.nn import functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


# --------------------------------------------------------------------------- nets
def q_param_names(depth: int):
    names = []
    for layer in range(depth):
        names += [f'c_dense.dense.{2 * layer}.linear.weight', f'c_dense.dense.{2 * layer}.linear.bias']
    names += [f'c_dense.dense.{2 * depth}.weight', f'c_dense.dense.{2 * depth}.bias']
    return names


def policy_param_names(depth: int):
    names = []
    for layer in range(depth):
        names += [f'c_dense.dense.{2 * layer}.linear.weight', f'c_dense.dense.{2 * layer}.linear.bias']
    names += ['mean_dense.dense.0.weight', 'mean_dense.dense.0.bias',
              'logstd_dense.dense.0.weight', 'logstd_dense.dense.0.bias']
    return names


def init_linear(out_dim: int, in_dim: int, gen: torch.Generator):
    """kaiming_uniform_(a=0) weight, zero bias (linear_layers.py:40-42,105-108)."""
    bound = math.sqrt(6.0 / in_dim)
    w = (torch.rand(out_dim, in_dim, generator=gen) * 2 - 1) * bound
    return w.float(), torch.zeros(out_dim)


def init_q(state_size, action_size, hidden, depth, gen):
    p, d_in = {}, state_size + action_size
    names = q_param_names(depth)
    for layer in range(depth):
        w, b = init_linear(hidden, d_in, gen)
        p[names[2 * layer]], p[names[2 * layer + 1]] = w, b
        d_in = hidden
    p[names[-2]], p[names[-1]] = init_linear(1, d_in, gen)
    return p


def init_policy(state_size, action_size, hidden, depth, gen):
    p, d_in = {}, state_size
    names = policy_param_names(depth)
    for layer in range(depth):
        w, b = init_linear(hidden, d_in, gen)
        p[names[2 * layer]], p[names[2 * layer + 1]] = w, b
        d_in = hidden
    p['mean_dense.dense.0.weight'], p['mean_dense.dense.0.bias'] = init_linear(action_size, d_in, gen)
    p['logstd_dense.dense.0.weight'], p['logstd_dense.dense.0.bias'] = init_linear(action_size, d_in, gen)
    return p


def trunk_forward(p: dict, depth: int, x: torch.Tensor) -> torch.Tensor:
    """``depth`` ResBlocks: GELU_erf(Wx+b), plus x when in==out (linear_layers.py:46-56)."""
    for layer in range(depth):
        w = p[f'c_dense.dense.{2 * layer}.linear.weight']
        b = p[f'c_dense.dense.{2 * layer}.linear.bias']
        h = F.gelu(F.linear(x, w, b))
        x = h + x if w.shape[0] == w.shape[1] else h
    return x


def q_forward(p: dict, depth: int, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """q.py:85-89 with identity ``dense`` / ``c_state_dense`` / ``c_action_dense``."""
    h = trunk_forward(p, depth, torch.cat([state, action], dim=-1))
    return F.linear(h, p[f'c_dense.dense.{2 * depth}.weight'], p[f'c_dense.dense.{2 * depth}.bias'])


def policy_forward(p: dict, depth: int, state: torch.Tensor):
    """policy.py:165-170 -> (loc, scale) of the pre-squash Normal."""
    h = trunk_forward(p, depth, state)
    mean = F.linear(h, p['mean_dense.dense.0.weight'], p['mean_dense.dense.0.bias'])
    logstd = F.linear(h, p['logstd_dense.dense.0.weight'], p['logstd_dense.dense.0.bias'])
    return torch.tanh(mean / 5.) * 5., torch.exp(torch.clamp(logstd, -20, 0.5))


def normal_log_prob(loc, scale, x):
    """torch.distributions.Normal.log_prob."""
    var = scale ** 2
    return -((x - loc) ** 2) / (2 * var) - scale.log() - LOG_SQRT_2PI


def squash_floor(x):
    return torch.maximum(1 - torch.square(torch.tanh(x)), torch.tensor(1e-2))


def squash_log_prob(loc, scale, x):
    """operators.py:12-14: the SUMMED Jacobian term is subtracted from EVERY action dim."""
    return normal_log_prob(loc, scale, x) - torch.sum(torch.log(squash_floor(x)), dim=-1, keepdim=True)


def squash_prob(loc, scale, x):
    """operators.py:17-19."""
    return torch.exp(normal_log_prob(loc, scale, x)) / torch.prod(squash_floor(x), dim=-1, keepdim=True)


def sum_log_prob(lp, keepdim=False):  # operators.py:22-24
    lp = lp.clone()
    lp[lp == torch.inf] = 0.
    return lp.sum(-1, keepdim=keepdim)


def prod_prob(prob):  # operators.py:27-31
    prob = prob.clone()
    prob[torch.isinf(prob)] = 1.
    out = prob.prod(-1)
    out[torch.logical_or(torch.isinf(out), torch.isnan(out))] = 1.
    return out


# --------------------------------------------------------------------------- learner
@dataclass
class SacHyper:
    state_size: int
    action_size: int
    ensemble_q_num: int = 2
    hidden: int = 64
    q_depth: int = 3
    policy_depth: int = 3
    burn_in_step: int = 0
    n_step: int = 1
    tau: float = 0.005
    update_target_per_step: int = 1
    init_log_alpha: float = -2.3
    use_auto_alpha: bool = True
    target_c_alpha: float = 1.
    learning_rate: float = 3e-4
    gamma: float = 0.99
    v_lambda: float = 1.
    v_rho: float = 1.
    v_c: float = 1.
    clip_epsilon: float = 0.2
    use_n_step_is: bool = True
    use_priority: bool = True
    ensemble_q_sample: int = 0  # 0 = every critic; < ensemble_q_num: min over a random subset (sac_base.py:1434-1436, 1887)


@dataclass
class SacBatch:
    """What ``_sample_from_replay_buffer`` hands to ``_train`` (sac_base.py:2466-2494)
    for a vector-obs run; ``states`` is concat(obs) == ModelSimpleRep output."""
    states: torch.Tensor          # [B, L, S]      L = b + n + 1
    actions: torch.Tensor         # [B, L-1, A]
    rewards: torch.Tensor         # [B, L-1]
    dones: torch.Tensor           # [B, L-1] bool
    mu_probs: torch.Tensor        # [B, L-1, A]
    last_masks: torch.Tensor      # [B, L-1] bool
    padding_masks: torch.Tensor   # [B, L-1] bool
    priority_is: torch.Tensor | None = None  # [B, 1]

    def to(self, dtype: torch.dtype) -> 'SacBatch':
        f = lambda t: None if t is None else (t.to(dtype) if t.is_floating_point() else t)
        return SacBatch(**{k: f(getattr(self, k)) for k in self.__dataclass_fields__})


@dataclass
class SacNoise:
    eps_y: torch.Tensor       # [B, n+1, A]  rsample in _get_y            (sac_base.py:1346)
    eps_pi: torch.Tensor      # [B, A]       rsample in _train_policy     (:1883)
    eps_alpha: torch.Tensor   # [B, A]       sample  in _train_alpha      (:1932)
    eps_td: torch.Tensor      # [B, n+1, A]  rsample in _get_td_error's _get_y (:2223)

    def to(self, dtype: torch.dtype) -> 'SacNoise':
        return SacNoise(**{k: getattr(self, k).to(dtype) for k in self.__dataclass_fields__})


class SacOracle:
    def __init__(self, hp: SacHyper, seed: int = 0, dtype: torch.dtype = torch.float32):
        """dtype=float64 evaluates the same formulas in double precision: used by the GPU tests to
        measure how much of a difference is the REFERENCE's own fp32 rounding (ill-conditioned rows,
        e.g. Normal.log_prob(rsample) with a tiny scale) rather than an error of the kernels."""
        self.hp = hp
        self.dtype = dtype
        gen = torch.Generator().manual_seed(seed)
        S, A, H = hp.state_size, hp.action_size, hp.hidden
        self.q = [init_q(S, A, H, hp.q_depth, gen) for _ in range(hp.ensemble_q_num)]
        self.q_target = [init_q(S, A, H, hp.q_depth, gen) for _ in range(hp.ensemble_q_num)]
        self.policy = init_policy(S, A, H, hp.policy_depth, gen)
        self.log_c_alpha = torch.tensor(hp.init_log_alpha, dtype=torch.float32)
        self.global_step = 0
        if dtype != torch.float32:
            cast = lambda d: {k: v.to(dtype) for k, v in d.items()}
            self.q, self.q_target = [cast(x) for x in self.q], [cast(x) for x in self.q_target]
            self.policy, self.log_c_alpha = cast(self.policy), self.log_c_alpha.to(dtype)
        self._wire()
        self.polyak(1.)  # fresh start: hard copy (sac_base.py:629)

    def _wire(self):
        hp = self.hp
        for net in self.q + [self.policy]:
            for t in net.values():
                t.requires_grad_(True)
        self.log_c_alpha.requires_grad_(True)
        self.opt_q = [torch.optim.Adam(list(q.values()), lr=hp.learning_rate) for q in self.q]
        self.opt_policy = torch.optim.Adam(list(self.policy.values()), lr=hp.learning_rate)
        # the reference hands [log_d_alpha, log_c_alpha] to one Adam; log_d_alpha never
        # receives a grad in continuous-only runs, so Adam skips it (sac_base.py:472,1946-1948)
        self.opt_alpha = torch.optim.Adam([self.log_c_alpha], lr=hp.learning_rate)
        n = hp.n_step
        self.gamma_ratio = torch.logspace(0, n - 1, n, hp.gamma).to(self.dtype)      # sac_base.py:285
        self.lambda_ratio = torch.logspace(0, n - 1, n, hp.v_lambda).to(self.dtype)  # :286

    def load_params(self, q, q_target, policy, log_c_alpha):
        to_t = lambda d: {k: torch.as_tensor(v).detach().to(self.dtype).clone() for k, v in d.items()}
        self.q = [to_t(x) for x in q]
        self.q_target = [to_t(x) for x in q_target]
        self.policy = to_t(policy)
        self.log_c_alpha = torch.as_tensor(log_c_alpha).detach().to(self.dtype).reshape(()).clone()
        self._wire()

    def copy_state_from(self, other: 'SacOracle', adam: bool = False) -> None:
        """Parameters and step counter of ``other`` (cast to this oracle's dtype); fresh Adam state unless
        ``adam`` (then the moments and step counts travel too, so a float64 twin can follow a float32 oracle
        over several steps)."""
        self.load_params(other.q, other.q_target, other.policy, other.log_c_alpha)
        self.global_step = other.global_step
        if adam:
            for i in range(len(self.q)):
                self._copy_adam(self.opt_q[i], self.q[i], other.opt_q[i], other.q[i])
            self._copy_adam(self.opt_policy, self.policy, other.opt_policy, other.policy)
            self._copy_adam(self.opt_alpha, {'a': self.log_c_alpha}, other.opt_alpha, {'a': other.log_c_alpha})

    def _copy_adam(self, opt, params: dict, src_opt, src_params: dict) -> None:
        for k, p in params.items():
            st = src_opt.state.get(src_params[k])
            if st:
                opt.state[p] = {'step': st['step'].clone(), 'exp_avg': st['exp_avg'].to(self.dtype).clone(),
                                'exp_avg_sq': st['exp_avg_sq'].to(self.dtype).clone()}

    # ---- sac_base.py:745-764
    @torch.no_grad()
    def polyak(self, tau: float):
        for tgt, src in zip(self.q_target, self.q):
            for k in tgt:
                tgt[k].copy_(tgt[k] * (1. - tau) + src[k] * tau)

    # ---- sac_base.py:1244-1295
    @torch.no_grad()
    def v_trace(self, last_masks, padding_masks, rewards, dones, mu_probs, pi_probs, vs, next_vs):
        hp = self.hp
        td = rewards + hp.gamma * ~dones * next_vs - vs
        td = self.gamma_ratio * td
        if hp.use_n_step_is:
            td = self.lambda_ratio * td
            ratio = pi_probs / mu_probs.clamp(min=1e-8)
            rho = torch.minimum(ratio, torch.tensor(hp.v_rho))
            c = torch.minimum(ratio, torch.tensor(hp.v_c))
            c = torch.cat([torch.ones((ratio.shape[0], 1), dtype=ratio.dtype), c[..., :-1]], dim=-1)
            c = torch.cumprod(c, dim=1)
            td = c * rho * td
        td = td * ~(torch.logical_or(last_masks, padding_masks))
        return vs[:, 0:1] + torch.sum(td, dim=1, keepdim=True)

    # ---- sac_base.py:1297-1466 (continuous branch)
    @torch.no_grad()
    def get_y(self, last_masks, padding_masks, nx_states, n_actions, rewards, dones, mu_probs, eps, perms=None):
        """``perms``: the two ``torch.randperm(E)`` draws of sac_base.py:1434 (current rows) and :1436 (next rows)
        when ``ensemble_q_sample < ensemble_q_num``."""
        hp = self.hp
        alpha = torch.exp(self.log_c_alpha)
        nx_actions = torch.cat([n_actions, torch.zeros_like(n_actions[:, :1])], dim=1)
        loc, scale = policy_forward(self.policy, hp.policy_depth, nx_states)
        sampled = loc + eps * scale  # Normal.rsample
        squashed = torch.tanh(sampled)
        qs = [q_forward(q, hp.q_depth, nx_states, squashed) for q in self.q_target]
        logp = sum_log_prob(squash_log_prob(loc, scale, sampled))  # [B, n+1]
        stacked = torch.stack(qs)
        Es = hp.ensemble_q_sample if 0 < hp.ensemble_q_sample < hp.ensemble_q_num else hp.ensemble_q_num
        sel_cur = torch.as_tensor(perms[0])[:Es] if perms is not None else slice(None)
        sel_next = torch.as_tensor(perms[1])[:Es] if perms is not None else slice(None)
        v_cur = stacked[sel_cur].min(dim=0)[0].squeeze(-1) - alpha * logp     # [B, n+1]; rows [:, :-1] are used
        v_next = stacked[sel_next].min(dim=0)[0].squeeze(-1) - alpha * logp   # rows [:, 1:] are used
        pi = mu = None
        if hp.use_n_step_is:
            stored = torch.atanh(torch.clamp(nx_actions, -0.999, 0.999))
            pi = prod_prob(squash_prob(loc, scale, stored)[:, :-1])
            mu = prod_prob(mu_probs)
        return self.v_trace(last_masks, padding_masks, rewards, dones, mu, pi, v_cur[:, :-1], v_next[:, 1:])

    # ---- sac_base.py:1516-1603
    def train_q(self, b: SacBatch, eps_y, perms=None):
        hp = self.hp; s0 = hp.burn_in_step
        state, action = b.states[:, s0], b.actions[:, s0]
        q_vals = [q_forward(q, hp.q_depth, state, action) for q in self.q]
        y = self.get_y(b.last_masks[:, s0:], b.padding_masks[:, s0:], b.states[:, s0:], b.actions[:, s0:],
                       b.rewards[:, s0:], b.dones[:, s0:], b.mu_probs[:, s0:], eps_y, perms)
        losses = []
        for i, q_val in enumerate(q_vals):
            if hp.clip_epsilon > 0:
                with torch.no_grad():
                    tq = q_forward(self.q_target[i], hp.q_depth, state, action)
                clipped = tq + torch.clamp(q_val - tq, -hp.clip_epsilon, hp.clip_epsilon)
                loss = torch.maximum((clipped - y) ** 2, (q_val - y) ** 2)
            else:
                loss = (q_val - y) ** 2
            if b.priority_is is not None:
                loss = loss * b.priority_is
            losses.append(torch.mean(loss))
        for opt in self.opt_q:
            opt.zero_grad()
        torch.stack(losses).sum().backward()
        grads = [{k: t.grad.clone() for k, t in q.items()} for q in self.q]
        for opt in self.opt_q:
            opt.step()
        return dict(y=y, q=[v.detach() for v in q_vals], loss_q=[l.detach() for l in losses], grad_q=grads)

    # ---- sac_base.py:1882-1911
    def train_policy(self, b: SacBatch, eps_pi, perm=None):
        hp = self.hp; s0 = hp.burn_in_step
        state = b.states[:, s0]
        loc, scale = policy_forward(self.policy, hp.policy_depth, state)
        with torch.no_grad():
            alpha = torch.exp(self.log_c_alpha)
        sampled = loc + eps_pi * scale
        qs = [q_forward(q, hp.q_depth, state, torch.tanh(sampled)) for q in self.q]
        logp = sum_log_prob(squash_log_prob(loc, scale, sampled), keepdim=True)
        stacked = torch.stack(qs)
        if perm is not None:  # sac_base.py:1887
            Es = hp.ensemble_q_sample if 0 < hp.ensemble_q_sample < hp.ensemble_q_num else hp.ensemble_q_num
            stacked = stacked[torch.as_tensor(perm)[:Es]]
        min_q = stacked.min(dim=0)[0]
        loss = torch.mean(alpha * logp - min_q)
        self.opt_policy.zero_grad()
        loss.backward(inputs=list(self.policy.values()))
        grads = {k: t.grad.clone() for k, t in self.policy.items()}
        self.opt_policy.step()
        entropy = torch.mean((0.5 + 0.5 * math.log(2 * math.pi) + torch.log(scale)).sum(-1)).detach()
        return dict(loss_policy=loss.detach(), grad_policy=grads, entropy=entropy)

    # ---- sac_base.py:1913-1949
    def train_alpha(self, b: SacBatch, eps_alpha):
        hp = self.hp; s0 = hp.burn_in_step
        with torch.no_grad():
            loc, scale = policy_forward(self.policy, hp.policy_depth, b.states[:, s0])
            sampled = loc + eps_alpha * scale  # Normal.sample == torch.normal(loc, scale)
            lp = squash_log_prob(loc, scale, sampled)
            valid = torch.sum(lp != torch.inf, dim=-1, keepdim=True)
            lp = sum_log_prob(lp, keepdim=True)
            target = hp.target_c_alpha * -valid
        loss = torch.mean(self.log_c_alpha * (-lp - target))
        self.opt_alpha.zero_grad()
        loss.backward(inputs=[self.log_c_alpha])
        grad = self.log_c_alpha.grad.clone()
        self.opt_alpha.step()
        return dict(loss_alpha=loss.detach(), grad_log_alpha=grad)

    # ---- sac_base.py:1159-1189
    @torch.no_grad()
    def l_probs(self, states, actions):
        loc, scale = policy_forward(self.policy, self.hp.policy_depth, states)
        return squash_prob(loc, scale, torch.atanh(torch.clamp(actions, -0.999, 0.999)))

    # ---- sac_base.py:2182-2245
    @torch.no_grad()
    def td_error(self, b: SacBatch, pi_probs, eps_td, perms=None):
        hp = self.hp; s0 = hp.burn_in_step
        state, action = b.states[:, s0], b.actions[:, s0]
        q_vals = [q_forward(q, hp.q_depth, state, action) for q in self.q]
        y = self.get_y(b.last_masks[:, s0:], b.padding_masks[:, s0:], b.states[:, s0:], b.actions[:, s0:],
                       b.rewards[:, s0:], b.dones[:, s0:],
                       pi_probs[:, s0:] if pi_probs is not None else None, eps_td, perms)
        err = torch.cat([torch.abs(qv - y) for qv in q_vals], dim=-1)
        return torch.mean(err, dim=-1, keepdim=True), y

    # ---- sac_base.py:2057-2126 + 2558-2583
    def step(self, b: SacBatch, noise: SacNoise, perms=None) -> dict:
        """``perms`` [5, E] (ensemble_q_sample < ensemble_q_num): the reference's randperm draws in call order —
        _get_y current / next rows, _train_policy, _get_y of _get_td_error current / next rows."""
        hp = self.hp
        if self.global_step % hp.update_target_per_step == 0:
            self.polyak(hp.tau)
        out = self.train_q(b, noise.eps_y, None if perms is None else perms[0:2])
        out.update(self.train_policy(b, noise.eps_pi, None if perms is None else perms[2]))
        if hp.use_auto_alpha:
            out.update(self.train_alpha(b, noise.eps_alpha))
        pi_probs = None
        if hp.use_n_step_is:
            pi_probs = self.l_probs(b.states[:, :-1], b.actions)
            out['pi_probs'] = pi_probs
        if hp.use_priority:
            out['td_error'], out['y_td'] = self.td_error(b, pi_probs, noise.eps_td, None if perms is None else perms[3:5])
        self.global_step += 1
        return out

    def snapshot(self) -> dict:
        d = {}
        for i, (q, qt) in enumerate(zip(self.q, self.q_target)):
            for k, t in q.items():
                d[f'q{i}.{k}'] = t.detach().clone().numpy()
            for k, t in qt.items():
                d[f'qt{i}.{k}'] = t.detach().clone().numpy()
        for k, t in self.policy.items():
            d[f'pi.{k}'] = t.detach().clone().numpy()
        d['log_c_alpha'] = self.log_c_alpha.detach().clone().numpy()
        return d
