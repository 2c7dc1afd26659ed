"""Discrete (and the discrete half of hybrid) action branches of the learner: flat parameter buffers of the
per-branch heads of every ModelQ and of ModelPolicy, and the launches of ``csrc/sac_discrete.cuh`` in the
reference's order (sac_base.py:1356-1421, 1543-1570, 1858-1880, 1924-1929, 1176-1178, 2226-2230).

The discrete nets share nothing with the continuous ones except the per-sample SUM of the losses (and the
optimizers: one Adam per ModelQ / ModelPolicy / [log_d_alpha, log_c_alpha], hence shared step counters), so
the learner runs them as their own stages next to the continuous kernels:

    d_y          policy logits + target critics on every row of the window -> expectation, V-trace    (stage 'target')
    critics      online critics on s_b -> loss gradient w.r.t. their outputs -> backward -> Adam         ('q')
    policy       online critics (after their step) on s_b, policy logits -> gradient -> backward -> Adam ('pi')
    alpha        policy logits after its step -> d loss / d log_d_alpha -> Adam                          ('post')
    l_probs      probabilities of every row -> mu-prob write-back
    td error     d_y again (mu := pi), |q_single - d_y| added to the continuous td error
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, lowering
from ._lib import check, ptr


class DiscreteBranch:
    def __init__(self, sac, q_list, target_q_list, policy):
        self.sac = sac
        lib = self.lib = sac._lib
        dev = sac.device
        f32 = dict(dtype=torch.float32, device=dev)
        E, B, L = sac.ensemble_q_num, sac.batch_size, sac.burn_in_step + sac.n_step + 1
        self.sizes = list(sac.d_action_sizes)
        self.K, self.D = len(self.sizes), sum(self.sizes)
        A = sac.c_action_size
        self.AF = self.D + A
        lowered = [lowering.analyze_d_heads(q, 'ModelQ') for q in list(q_list) + list(target_q_list)]
        pi_shapes, pi_params = lowering.analyze_d_heads(policy, 'ModelPolicy')
        shapes = lowered[0][0]
        if any(s != shapes for s, _ in lowered) or [(s.in_dim, s.hidden, s.depth) for s in pi_shapes] != \
                [(s.in_dim, s.hidden, s.depth) for s in shapes]:
            raise lowering.NotStockNetwork('discrete heads of critics / targets / policy differ in shape')
        if shapes[0].in_dim != sac.state_size:
            raise lowering.NotStockNetwork('discrete heads do not read the state directly')
        self.shapes = shapes
        cfg = self.cfg = _lib.AsacDiscreteConfig()
        cfg.branches, cfg.hidden, cfg.depth, cfg.state_size = self.K, shapes[0].hidden, shapes[0].depth, sac.state_size
        for k, size in enumerate(self.sizes):
            cfg.sizes[k] = size
        cfg.target_d_alpha = float(sac.target_d_alpha)
        cfg.entropy_penalty = float(sac.d_policy_entropy_penalty)
        P = self.P = int(lib.asac_dnets_member_floats(C.byref(cfg)))
        self.branch_off = np.cumsum([0] + [s.stride for s in shapes])[:-1]
        # flat buffers: online critics [E, P], targets [E, P], policy [P]; Adam moments alike
        self.q, self.qt, self.pi = torch.zeros(E, P, **f32), torch.zeros(E, P, **f32), torch.zeros(P, **f32)
        self.q_m, self.q_v = torch.zeros(E, P, **f32), torch.zeros(E, P, **f32)
        self.pi_m, self.pi_v = torch.zeros(P, **f32), torch.zeros(P, **f32)
        for i in range(E):
            for k in range(self.K):
                o = int(self.branch_off[k])
                lowering.bind_parameters(lowered[i][1][k], self.q[i, o:o + shapes[k].stride])
                lowering.bind_parameters(lowered[E + i][1][k], self.qt[i, o:o + shapes[k].stride])
        for k in range(self.K):
            o = int(self.branch_off[k])
            lowering.bind_parameters(pi_params[k], self.pi[o:o + shapes[k].stride])
        self.dqn_like = bool(sac.discrete_dqn_like)
        self.log_alpha = torch.full((1,), float(sac._init_log_alpha), **f32)
        self.alpha_m, self.alpha_v = torch.zeros(1, **f32), torch.zeros(1, **f32)
        # per-step buffers
        R = B * L
        self.tiles = int(lib.asac_dnets_tiles(B))
        self.wk = {
            'pi_logits': torch.zeros(R, self.D, **f32), 'tq': torch.zeros(E, R, self.D, **f32),
            'q_b': torch.zeros(E, B, self.D, **f32), 'd_y': torch.zeros(B, **f32), 'd_y_td': torch.zeros(B, **f32),
            'dq_out': torch.zeros(E, B, self.D, **f32), 'loss_q': torch.zeros(E, B, **f32),
            'q_single': torch.zeros(E, B, **f32), 'dpi_out': torch.zeros(1, B, self.D, **f32),
            'loss_pi': torch.zeros(B, **f32), 'entropy': torch.zeros(B, **f32),
            'grad_q_part': torch.zeros(self.tiles, E, P, **f32), 'grad_q': torch.zeros(E, P, **f32),
            'grad_pi_part': torch.zeros(self.tiles, 1, P, **f32), 'grad_pi': torch.zeros(P, **f32),
            'pi_probs_d': torch.zeros(B, L - 1, self.D, **f32), 'grad_alpha': torch.zeros(1, **f32),
            'loss_alpha': torch.zeros(1, **f32),
        }
        if self.dqn_like:
            self.wk['eq'] = torch.zeros(E, R, self.D, **f32)  # online critics on every row (the arg-max side)

    # ------------------------------------------------------------------ helpers
    def _forward(self, params, member_stride, members, x, x_row_stride, rows, out):
        # (x may be a strided view — states[:, b] — hence its data pointer and an explicit row stride)
        check(self.lib.asac_dnets_forward(C.byref(self.cfg), ptr(params), member_stride, members, x.data_ptr(), x_row_stride,
                                          rows, ptr(out), _lib.current_stream()), 'dnets_forward')

    def _backward(self, params, member_stride, members, x, x_row_stride, rows, d_out, grad_part, d_x=None):
        check(self.lib.asac_dnets_backward(C.byref(self.cfg), ptr(params), member_stride, members, x.data_ptr(), x_row_stride,
                                           rows, ptr(d_out), ptr(grad_part), ptr(d_x), _lib.current_stream()),
              'dnets_backward')

    def _adam(self, param, m, v, part, tiles, count, grad, counter_index):
        sac = self.sac
        if sac._world > 1:  # reduce first, average over ranks, then Adam on the averaged gradient
            torch.sum(part.view(tiles, -1), dim=0, out=grad.view(-1))
            from . import dist as adist
            adist.all_reduce_sum_(grad)
            grad.mul_(1.0 / sac._world)
            part, tiles = grad, 1
        check(self.lib.asac_flat_reduce_adam(ptr(param), ptr(m), ptr(v), ptr(part), tiles, count, count, ptr(grad),
                                             ptr(sac._counters[counter_index:]), float(sac.learning_rate),
                                             _lib.current_stream()), 'flat_reduce_adam')

    def polyak(self, force_tau: float | None = None) -> None:
        sac = self.sac
        tau = float(sac._cfg.tau) if force_tau is None else float(force_tau)
        check(self.lib.asac_flat_polyak(ptr(self.qt), ptr(self.q), self.qt.numel(), ptr(sac._counters),
                                        int(sac.update_target_per_step), tau, float(np.float32(1. - tau)),
                                        0 if force_tau is None else 1, _lib.current_stream()), 'flat_polyak')

    # ------------------------------------------------------------------ stages
    def stage_target(self, st, states_pi, states_v, post: bool) -> None:
        """d_y (train) / d_y of the td-error pass (post): sac_base.py:1384-1412."""
        sac, wk, bt = self.sac, self.wk, st['bt']
        B, L, S = sac.batch_size, sac._cfg.seq_len, sac.state_size
        E, P = sac.ensemble_q_num, self.P
        if post and states_v.data_ptr() != states_pi.data_ptr() and not self.dqn_like:
            # trained representation: _get_td_error's _get_y runs the policy on the TARGET representation's states
            # (sac_base.py:2226-2233), the probabilities of get_l_probs (the mu side) on the re-encoded ones
            if 'pi_logits_v' not in wk:
                wk['pi_logits_v'] = torch.zeros_like(wk['pi_logits'])
            self._forward(self.pi, 0, 1, states_v, S, B * L, wk['pi_logits_v'])
            logits_v = wk['pi_logits_v']
        else:
            logits_v = wk['pi_logits']
        if self.dqn_like:  # get_dqn_like_d_y: no policy in the target (sac_base.py:1193-1242, 1363-1383)
            self._forward(self.qt, P, E, states_v, S, B * L, wk['tq'])
            self._forward(self.q, P, E, states_v, S, B * L, wk['eq'])
            perms = sac._ens_perms[7:9] if post else sac._ens_perms[5:7]   # (target, online) of this _get_y call
            check(self.lib.asac_d_target_dqn(C.byref(sac._cfg_d), C.byref(self.cfg), ptr(wk['eq']), ptr(wk['tq']),
                                             ptr(perms[0]), ptr(perms[1]), ptr(bt['rewards']), ptr(bt['dones']),
                                             ptr(bt['last_masks']), ptr(bt['padding_masks']),
                                             ptr(wk['d_y_td'] if post else wk['d_y']), _lib.current_stream()),
                  'd_target_dqn')
            return
        if not post:  # (post: the logits of the policy after its step are already there, see stage_alpha)
            self._forward(self.pi, 0, 1, states_pi, S, B * L, wk['pi_logits'])
        self._forward(self.qt, P, E, states_v, S, B * L, wk['tq'])
        check(self.lib.asac_d_target(C.byref(sac._cfg_d), C.byref(self.cfg), ptr(logits_v), ptr(wk['tq']),
                                     ptr(bt['actions_full']), ptr(bt['mu_full']),
                                     ptr(wk['pi_probs_d']) if post else None, ptr(bt['rewards']), ptr(bt['dones']),
                                     ptr(bt['last_masks']), ptr(bt['padding_masks']), ptr(self.log_alpha),
                                     ptr(wk['d_y_td'] if post else wk['d_y']), _lib.current_stream()), 'd_target')

    def stage_q(self, st, states, scale: float, want_dx: bool = False) -> None:
        """Critic loss of the discrete part, backward, partial gradients (no Adam yet).  ``want_dx``: also
        d loss_i / d state[:, b] of every member and branch into wk['d_x'] [E, K, B, S] (trained representation)."""
        sac, wk, bt = self.sac, self.wk, st['bt']
        B, L, S, b = sac.batch_size, sac._cfg.seq_len, sac.state_size, sac.burn_in_step
        E, P = sac.ensemble_q_num, self.P
        x = states[:, b]  # [B, S] view: row stride L * S
        self._forward(self.q, P, E, x, L * S, B, wk['q_b'])
        w = st['smp']['w'] if (sac.use_priority and sac.use_replay_buffer) else None
        check(self.lib.asac_d_q_grad(C.byref(sac._cfg_d), C.byref(self.cfg), ptr(wk['q_b']), ptr(bt['actions_full']),
                                     ptr(wk['d_y']), ptr(w), float(scale), ptr(wk['dq_out']), ptr(wk['loss_q']),
                                     ptr(wk['q_single']), _lib.current_stream()), 'd_q_grad')
        if want_dx and 'd_x' not in wk:
            wk['d_x'] = torch.zeros(E, self.K, B, S, dtype=torch.float32, device=sac.device)
        self._backward(self.q, P, E, x, L * S, B, wk['dq_out'], wk['grad_q_part'], wk['d_x'] if want_dx else None)

    def adam_q(self) -> None:
        wk, E, P = self.wk, self.sac.ensemble_q_num, self.P
        self._adam(self.q, self.q_m, self.q_v, wk['grad_q_part'], self.tiles, E * P, wk['grad_q'], 1)

    def stage_pi(self, st, states) -> None:
        """Policy loss of the discrete part (sac_base.py:1858-1880): critics after their step, no gradient to them."""
        sac, wk, bt = self.sac, self.wk, st['bt']
        B, L, S, b = sac.batch_size, sac._cfg.seq_len, sac.state_size, sac.burn_in_step
        E, P = sac.ensemble_q_num, self.P
        x = states[:, b]
        self._forward(self.q, P, E, x, L * S, B, wk['q_b'])
        self._forward(self.pi, 0, 1, x, L * S, B, wk['dpi_out'])  # logits of s_b (scratch: overwritten by the gradient)
        logits = wk['dpi_out'].clone()
        self._logits_b = logits
        check(self.lib.asac_d_pi_grad(C.byref(sac._cfg_d), C.byref(self.cfg), ptr(logits), 1, 0, ptr(wk['q_b']),
                                      ptr(bt['mu_full']), ptr(self.log_alpha), ptr(wk['dpi_out']), ptr(wk['loss_pi']),
                                      ptr(wk['entropy']), _lib.current_stream()), 'd_pi_grad')
        self._backward(self.pi, 0, 1, x, L * S, B, wk['dpi_out'], wk['grad_pi_part'])

    def adam_pi(self) -> None:
        wk = self.wk
        self._adam(self.pi, self.pi_m, self.pi_v, wk['grad_pi_part'], self.tiles, self.P, wk['grad_pi'], 2)

    def stage_alpha(self, st, states_pi) -> None:
        """Policy logits after the policy step on every row; _train_alpha's discrete part + Adam on log_d_alpha
        (reads the alpha optimizer's step counter: call BEFORE that counter is advanced)."""
        sac, wk = self.sac, self.wk
        B, L, S = sac.batch_size, sac._cfg.seq_len, sac.state_size
        self._forward(self.pi, 0, 1, states_pi, S, B * L, wk['pi_logits'])
        if sac.use_auto_alpha and not self.dqn_like:  # (DQN-like: no discrete alpha loss, sac_base.py:1924)
            check(self.lib.asac_d_alpha(C.byref(sac._cfg_d), C.byref(self.cfg), ptr(wk['pi_logits']), ptr(self.log_alpha),
                                        ptr(self.alpha_m), ptr(self.alpha_v), ptr(sac._counters[3:]), 1.0,
                                        ptr(wk['grad_alpha']), ptr(wk['loss_alpha']), _lib.current_stream()), 'd_alpha')

    def stage_probs_td(self, st, states_pi, states_v, states_b, accumulate_td: bool) -> None:
        """get_l_probs (discrete columns) and the discrete part of the td error, after stage_alpha."""
        sac, wk, bt = self.sac, self.wk, st['bt']
        B, L, S, b = sac.batch_size, sac._cfg.seq_len, sac.state_size, sac.burn_in_step
        E, P = sac.ensemble_q_num, self.P
        s = _lib.current_stream()
        full = sac._wk.get('pi_probs_full')
        check(self.lib.asac_d_probs(C.byref(sac._cfg_d), C.byref(self.cfg), ptr(wk['pi_logits']), ptr(wk['pi_probs_d']),
                                    ptr(full), s), 'd_probs')
        if sac.use_priority:
            self.stage_target(st, states_pi, states_v, post=True)
            x = states_b[:, b]
            self._forward(self.q, P, E, x, L * S, B, wk['q_b'])
            check(self.lib.asac_d_td(C.byref(sac._cfg_d), C.byref(self.cfg), ptr(wk['q_b']), ptr(bt['actions_full']),
                                     ptr(wk['d_y_td']), ptr(sac._wk['td_error']), 1 if accumulate_td else 0, s), 'd_td')
