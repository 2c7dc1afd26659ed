"""Drives the SAC kernels of libasac_b200.so through the C ABI on explicit tensors, stage by
stage, so that every intermediate can be compared with the CPU oracle (tests only)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from asac_b200 import _lib, lowering
from asac_b200._lib import check, ptr


class SacCuda:
    """Holds the flat parameters / Adam state / work buffers for one SacHyper and exposes the
    staged entry points of include/asac_b200.h."""

    def __init__(self, hp, batch_size: int, device='cuda:0', rep_kind: int = 0):
        self.lib = _lib.load()
        self.hp, self.B = hp, batch_size
        self.dev = torch.device(device)
        S, A, E = hp.state_size, hp.action_size, hp.ensemble_q_num
        self.q_shape = lowering.NetShape(S + A, hp.hidden, hp.q_depth, 1)
        self.pi_shape = lowering.NetShape(S, hp.hidden, hp.policy_depth, 2 * A)
        assert self.q_shape.count == self.lib.asac_mlp_param_count(S + A, hp.hidden, hp.q_depth, 1)
        assert self.q_shape.stride == self.lib.asac_mlp_param_stride(S + A, hp.hidden, hp.q_depth, 1)
        assert self.pi_shape.count == self.lib.asac_mlp_param_count(S, hp.hidden, hp.policy_depth, 2 * A)
        n, L = hp.n_step, hp.burn_in_step + hp.n_step + 1
        self.L = L
        cfg = _lib.AsacSacConfig()
        cfg.learning_rate = float(hp.learning_rate)
        cfg.batch, cfg.seq_len, cfg.burn_in, cfg.n_step = batch_size, L, hp.burn_in_step, n
        cfg.state_size, cfg.action_size, cfg.ensemble = S, A, E
        cfg.q_hidden, cfg.q_depth, cfg.pi_hidden, cfg.pi_depth = hp.hidden, hp.q_depth, hp.hidden, hp.policy_depth
        cfg.use_n_step_is, cfg.use_priority = int(hp.use_n_step_is), int(hp.use_priority)
        cfg.use_auto_alpha = int(hp.use_auto_alpha)
        cfg.update_target_per_step = int(hp.update_target_per_step)
        cfg.bn_stride = L - 1  # golden batches carry [B, L-1, ...] arrays
        cfg.tau, cfg.one_minus_tau = float(hp.tau), float(np.float32(1. - hp.tau))
        cfg.gamma, cfg.v_rho, cfg.v_c = float(hp.gamma), float(hp.v_rho), float(hp.v_c)
        cfg.clip_epsilon, cfg.target_c_alpha = float(hp.clip_epsilon), float(hp.target_c_alpha)
        cfg.td_error_min, cfg.td_error_max, cfg.per_alpha = 0.01, 1.0, 0.9
        cfg.rep_kind = rep_kind
        cfg.ensemble_sample = int(getattr(hp, 'ensemble_q_sample', 0) or 0)
        gr = torch.logspace(0, n - 1, n, hp.gamma)
        lr = torch.logspace(0, n - 1, n, hp.v_lambda)
        for k in range(n):
            cfg.gamma_ratio[k], cfg.lambda_ratio[k] = float(gr[k]), float(lr[k])
        self.cfg = cfg
        tile = self.lib.asac_sac_tile_batch(C.byref(cfg))
        assert tile >= 1, self.lib.asac_last_error()
        self.tile = tile
        T = self.n_tiles = (batch_size + tile - 1) // tile

        f32 = dict(dtype=torch.float32, device=self.dev)
        Pq, Ppi = self.q_shape.stride, self.pi_shape.stride
        self.q = torch.zeros(E, Pq, **f32)
        self.qt = torch.zeros(E, Pq, **f32)
        self.pi = torch.zeros(Ppi, **f32)
        self.log_alpha = torch.zeros(1, **f32)
        self.q_m, self.q_v = torch.zeros_like(self.q), torch.zeros_like(self.q)
        self.pi_m, self.pi_v = torch.zeros_like(self.pi), torch.zeros_like(self.pi)
        self.alpha_m, self.alpha_v = torch.zeros(1, **f32), torch.zeros(1, **f32)
        self.counters = torch.zeros(8, dtype=torch.int64, device=self.dev)
        prm = _lib.AsacSacParams()
        prm.q, prm.q_target, prm.pi, prm.log_alpha = ptr(self.q), ptr(self.qt), ptr(self.pi), ptr(self.log_alpha)
        prm.q_m, prm.q_v, prm.pi_m, prm.pi_v = ptr(self.q_m), ptr(self.q_v), ptr(self.pi_m), ptr(self.pi_v)
        prm.alpha_m, prm.alpha_v, prm.counters = ptr(self.alpha_m), ptr(self.alpha_v), ptr(self.counters)
        self.prm = prm

        self.wk = {
            'y': torch.zeros(batch_size, **f32), 'tq': torch.zeros(E, batch_size, **f32),
            'q_val': torch.zeros(E, batch_size, **f32), 'loss_q': torch.zeros(T, E, **f32),
            'grad_q_part': torch.zeros(T, E, Pq, **f32), 'grad_q': torch.zeros(E, Pq, **f32),
            'grad_pi_part': torch.zeros(T, Ppi, **f32), 'grad_pi': torch.zeros(Ppi, **f32),
            'stats_pi': torch.zeros(T, 2, **f32), 'grad_alpha_part': torch.zeros(T, 2, **f32),
            'grad_alpha': torch.zeros(1, **f32), 'pi_probs': torch.zeros(batch_size, L - 1, A, **f32),
            'post_parts': torch.zeros(batch_size, 2 + E, **f32),
            'y_td': torch.zeros(batch_size, **f32), 'td_error': torch.zeros(batch_size, **f32),
        }
        work = _lib.AsacSacWork()
        work.n_tiles = T
        for k, t in self.wk.items():
            setattr(work, k, ptr(t))
        self.work = work
        self._keep = []

    # ---- parameters
    def load_params(self, q, q_target, policy, log_c_alpha):
        for i, sd in enumerate(q):
            self.q[i].copy_(lowering.flat_from_state_dict(self.q_shape, sd, policy=False))
        for i, sd in enumerate(q_target):
            self.qt[i].copy_(lowering.flat_from_state_dict(self.q_shape, sd, policy=False))
        self.pi.copy_(lowering.flat_from_state_dict(self.pi_shape, policy, policy=True))
        self.log_alpha.fill_(float(log_c_alpha))

    def sync_from_oracle(self, oracle, what=('q', 'qt', 'pi', 'alpha')) -> None:
        """Copies the oracle's parameters AND torch.optim.Adam state (exp_avg, exp_avg_sq, step) into
        the CUDA buffers, so that a stage can be compared from an identical starting point."""
        def moments(opt, params: dict, key: str):
            out = {}
            for name, p in params.items():
                st = opt.state.get(p, {})
                out[name] = st[key].detach() if key in st else torch.zeros_like(p)
            return out

        def steps(opt, params: dict) -> int:
            vals = {int(float(opt.state[p]['step'])) for p in params.values() if p in opt.state}
            assert len(vals) <= 1
            return vals.pop() if vals else 0

        E = self.hp.ensemble_q_num
        if 'q' in what:
            for i in range(E):
                det = {k: v.detach() for k, v in oracle.q[i].items()}
                self.q[i].copy_(lowering.flat_from_state_dict(self.q_shape, det, False))
                self.q_m[i].copy_(lowering.flat_from_state_dict(self.q_shape, moments(oracle.opt_q[i], oracle.q[i], 'exp_avg'), False))
                self.q_v[i].copy_(lowering.flat_from_state_dict(self.q_shape, moments(oracle.opt_q[i], oracle.q[i], 'exp_avg_sq'), False))
            self.counters[1] = steps(oracle.opt_q[0], oracle.q[0])
        if 'qt' in what:
            for i in range(E):
                self.qt[i].copy_(lowering.flat_from_state_dict(self.q_shape, oracle.q_target[i], False))
        if 'pi' in what:
            det = {k: v.detach() for k, v in oracle.policy.items()}
            self.pi.copy_(lowering.flat_from_state_dict(self.pi_shape, det, True))
            self.pi_m.copy_(lowering.flat_from_state_dict(self.pi_shape, moments(oracle.opt_policy, oracle.policy, 'exp_avg'), True))
            self.pi_v.copy_(lowering.flat_from_state_dict(self.pi_shape, moments(oracle.opt_policy, oracle.policy, 'exp_avg_sq'), True))
            self.counters[2] = steps(oracle.opt_policy, oracle.policy)
        if 'alpha' in what:
            self.log_alpha.fill_(float(oracle.log_c_alpha.detach()))
            st = oracle.opt_alpha.state.get(oracle.log_c_alpha, {})
            self.alpha_m.fill_(float(st['exp_avg']) if 'exp_avg' in st else 0.)
            self.alpha_v.fill_(float(st['exp_avg_sq']) if 'exp_avg_sq' in st else 0.)
            self.counters[3] = int(float(st['step'])) if 'step' in st else 0
        self.counters[0] = oracle.global_step

    def td_error(self):
        check(self.lib.asac_sac_td_error(C.byref(self.cfg), C.byref(self.prm), C.byref(self.work), self._s()),
              'td_error')

    def snapshot(self) -> dict:
        d = {}
        for i in range(self.hp.ensemble_q_num):
            for k, t in lowering.state_dict_from_flat(self.q_shape, self.q[i], False).items():
                d[f'q{i}.{k}'] = t.cpu().numpy()
            for k, t in lowering.state_dict_from_flat(self.q_shape, self.qt[i], False).items():
                d[f'qt{i}.{k}'] = t.cpu().numpy()
        for k, t in lowering.state_dict_from_flat(self.pi_shape, self.pi, True).items():
            d[f'pi.{k}'] = t.cpu().numpy()
        d['log_c_alpha'] = self.log_alpha[0].cpu().numpy()
        return d

    def grad_q_dict(self, i: int) -> dict:
        return {k: t.cpu().numpy() for k, t in
                lowering.state_dict_from_flat(self.q_shape, self.wk['grad_q'][i], False).items()}

    def grad_pi_dict(self) -> dict:
        return {k: t.cpu().numpy() for k, t in
                lowering.state_dict_from_flat(self.pi_shape, self.wk['grad_pi'], True).items()}

    # ---- batch
    def make_batch(self, b, noise, perms=None) -> _lib.AsacSacBatch:
        """b: oracle SacBatch (CPU tensors, [B, L-1, ...] layout), noise: SacNoise; perms [5, E]: the reference's
        randperm draws of the step (ensemble_q_sample < ensemble_q_num)."""
        dev = self.dev
        t = {
            'states': b.states.float(), 'actions': b.actions.float(), 'rewards': b.rewards.float(),
            'dones': b.dones.to(torch.uint8), 'last_masks': b.last_masks.to(torch.uint8),
            'padding_masks': b.padding_masks.to(torch.uint8), 'mu_probs': b.mu_probs.float(),
            'eps_y': noise.eps_y.float(), 'eps_pi': noise.eps_pi.float(), 'eps_alpha': noise.eps_alpha.float(),
            'eps_td': noise.eps_td.float(),
        }
        t = {k: v.contiguous().to(dev) for k, v in t.items()}
        if b.priority_is is not None:
            t['priority_is'] = b.priority_is.float().contiguous().view(-1).to(dev)
        if perms is not None:
            t['ensemble_perms'] = torch.as_tensor(np.asarray(perms), dtype=torch.int32).contiguous().to(dev)
        self._keep = [t]
        batch = _lib.AsacSacBatch()
        for k, v in t.items():
            setattr(batch, k, ptr(v))
        return batch

    # ---- staged calls
    def _s(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def polyak(self, force_tau=-1.0):
        check(self.lib.asac_sac_polyak(C.byref(self.cfg), C.byref(self.prm), float(force_tau), self._s()), 'polyak')

    def target_y(self, batch):
        check(self.lib.asac_sac_target_y(C.byref(self.cfg), C.byref(self.prm), C.byref(batch), C.byref(self.work),
                                         self._s()), 'target_y')

    def q_backward(self, batch):
        check(self.lib.asac_sac_q_backward(C.byref(self.cfg), C.byref(self.prm), C.byref(batch), C.byref(self.work),
                                           self._s()), 'q_backward')

    def policy_backward(self, batch):
        check(self.lib.asac_sac_policy_backward(C.byref(self.cfg), C.byref(self.prm), C.byref(batch),
                                                C.byref(self.work), self._s()), 'policy_backward')

    def post(self, batch):
        check(self.lib.asac_sac_post(C.byref(self.cfg), C.byref(self.prm), C.byref(batch), C.byref(self.work),
                                     self._s()), 'post')

    def reduce_grads(self, which):
        check(self.lib.asac_sac_reduce_grads(C.byref(self.cfg), C.byref(self.work), which, self._s()), 'reduce')

    def adam(self, which, scale=1.0):
        check(self.lib.asac_sac_adam(C.byref(self.cfg), C.byref(self.prm), C.byref(self.work), which, float(scale),
                                     self._s()), 'adam')

    def reduce_adam(self, which):
        check(self.lib.asac_sac_reduce_adam(C.byref(self.cfg), C.byref(self.prm), C.byref(self.work), which,
                                            self._s()), 'reduce_adam')

    def advance(self):
        check(self.lib.asac_sac_advance_step(C.byref(self.prm), self._s()), 'advance')

    def step(self, batch):
        check(self.lib.asac_sac_step(C.byref(self.cfg), C.byref(self.prm), C.byref(batch), C.byref(self.work),
                                     self._s()), 'sac_step')

    def step_networks(self, batch, with_polyak=1):
        check(self.lib.asac_sac_step_networks(C.byref(self.cfg), C.byref(self.prm), C.byref(batch),
                                              C.byref(self.work), int(with_polyak), None, self._s()), 'sac_step_networks')

    def finish_step(self, nodes, capacity, store_ids, data_ids, per_state):
        check(self.lib.asac_sac_finish_step(C.byref(self.cfg), C.byref(self.prm), C.byref(self.work), ptr(nodes),
                                            capacity, ptr(store_ids), ptr(data_ids), ptr(per_state), None, self._s()),
              'sac_finish_step')

    def staged_step(self, batch) -> dict:
        """The sequence of asac_sac_step, one entry point at a time, returning every intermediate."""
        hp, out = self.hp, {}
        self.polyak()
        self.target_y(batch)
        out['y'] = self.wk['y'].cpu().numpy()
        self.q_backward(batch)
        self.reduce_grads(0)
        out['q'] = self.wk['q_val'].cpu().numpy()
        out['loss_q'] = (self.wk['loss_q'].sum(0) / self.B).cpu().numpy()
        out['grad_q'] = [self.grad_q_dict(i) for i in range(hp.ensemble_q_num)]
        self.adam(0)
        self.policy_backward(batch)
        self.reduce_grads(1)
        out['grad_policy'] = self.grad_pi_dict()
        out['loss_policy'] = float(self.wk['stats_pi'][:, 0].sum().item()) / self.B
        out['entropy'] = float(self.wk['stats_pi'][:, 1].sum().item()) / self.B
        self.adam(1)
        need_post = hp.use_auto_alpha or hp.use_n_step_is or hp.use_priority
        if need_post:
            self.post(batch)
            if hp.use_n_step_is:
                out['pi_probs'] = self.wk['pi_probs'].cpu().numpy()
        if hp.use_auto_alpha:
            self.reduce_grads(2)
            out['grad_log_alpha'] = self.wk['grad_alpha'].cpu().numpy()
            self.adam(2)
        if need_post:  # _get_td_error sees the updated alpha (sac_base.py:2115-2116 precede :2571)
            check(self.lib.asac_sac_td_error(C.byref(self.cfg), C.byref(self.prm), C.byref(self.work), self._s()),
                  'td_error')
            out['td_error'] = self.wk['td_error'].cpu().numpy()
            out['y_td'] = self.wk['y_td'].cpu().numpy()
        self.advance()
        return out


# ------------------------------------------------------------------------------------------ GRU representation
def gru_forward(shape: lowering.GruShape, params_list, obs, actions=None, pre_actions=None, h0=None,
                save=False, want_hn=True):
    """asac_gru_forward on explicit CUDA tensors -> list of dicts (states, hn, save) per parameter set."""
    lib = _lib.load()
    B, L = obs.shape[:2]
    H, NL = shape.hidden, shape.layers
    f32 = dict(dtype=torch.float32, device=obs.device)
    cs = _lib.AsacGruShape(shape.obs_size, shape.action_size, H, NL)
    nets = (_lib.AsacGruNet * len(params_list))()
    outs = []
    for i, prm in enumerate(params_list):
        o = {'states': torch.zeros(B, L, H, **f32), 'hn': torch.zeros(B, L, NL, H, **f32) if want_hn else None,
             'save': torch.zeros(B, L, NL, 4 * H, **f32) if save else None}
        nets[i].params, nets[i].states, nets[i].hn, nets[i].save = ptr(prm), ptr(o['states']), ptr(o['hn']), ptr(o['save'])
        outs.append(o)
    bn_stride = 0 if actions is None else actions.shape[1]
    check(lib.asac_gru_forward(C.byref(cs), nets, len(params_list), ptr(obs), ptr(actions), bn_stride,
                               ptr(pre_actions), ptr(h0), 0 if h0 is None else h0[0].numel(), B, L,
                               torch.cuda.current_stream(obs.device).cuda_stream), 'gru_forward')
    return outs


def gru_backward(shape: lowering.GruShape, params, obs, actions, h0, t_grad, grad_state, fwd):
    """asac_gru_backward -> reduced gradient [P] (the partial tiles summed in tile order, as
    asac_flat_reduce_adam does)."""
    lib = _lib.load()
    B, L = obs.shape[:2]
    cs = _lib.AsacGruShape(shape.obs_size, shape.action_size, shape.hidden, shape.layers)
    tile = lib.asac_gru_backward_tile(C.byref(cs), t_grad)
    assert tile >= 1, lib.asac_last_error()
    tiles = B  # one partial gradient per sequence
    part = torch.full((tiles, shape.stride), float('nan'), dtype=torch.float32, device=obs.device)
    check(lib.asac_gru_backward(C.byref(cs), ptr(params), ptr(obs), ptr(actions), actions.shape[1], None, ptr(h0),
                                0 if h0 is None else h0[0].numel(), B, L, t_grad, ptr(grad_state),
                                grad_state.shape[0], ptr(fwd['hn']), ptr(fwd['save']), ptr(part),
                                torch.cuda.current_stream(obs.device).cuda_stream), 'gru_backward')
    total = torch.zeros(shape.stride, dtype=torch.float32, device=obs.device)
    for t in range(tiles):
        total += part[t]
    return total[:shape.count], part


class SacRepCuda(SacCuda):
    """SacCuda + a trained GRU representation (asac_sac_step_networks_rep)."""

    def __init__(self, hp, batch_size: int, obs_size: int, rep_layers: int, device='cuda:0'):
        super().__init__(hp, batch_size, device, rep_kind=1)
        B, L, H, A, E = batch_size, self.L, hp.state_size, hp.action_size, hp.ensemble_q_num
        self.gshape = lowering.GruShape(obs_size, A, H, rep_layers)
        cs = _lib.AsacGruShape(obs_size, A, H, rep_layers)
        assert self.lib.asac_gru_param_count(C.byref(cs)) == self.gshape.count
        f32 = dict(dtype=torch.float32, device=self.dev)
        P = self.gshape.stride
        self.rep_p, self.rep_t = torch.zeros(P, **f32), torch.zeros(P, **f32)
        self.rep_m, self.rep_v = torch.zeros(P, **f32), torch.zeros(P, **f32)
        rtile = self.lib.asac_gru_backward_tile(C.byref(cs), hp.burn_in_step)
        assert rtile >= 1, self.lib.asac_last_error()
        rt = B
        self.rb = {'obs': torch.zeros(B, L, obs_size, **f32), 'h0': torch.zeros(B, rep_layers, H, **f32),
                   'states': torch.zeros(B, L, H, **f32), 'states_post': torch.zeros(B, L, H, **f32),
                   'target_states': torch.zeros(B, L, H, **f32), 'hn': torch.zeros(B, L, rep_layers, H, **f32),
                   'hn_post': torch.zeros(B, L, rep_layers, H, **f32),
                   'save': torch.zeros(B, L, rep_layers, 4 * H, **f32), 'grad_part': torch.zeros(rt, P, **f32),
                   'grad': torch.zeros(P, **f32)}
        self.wk['grad_state'] = torch.zeros(E, B, H, **f32)
        self.work.grad_state = ptr(self.wk['grad_state'])
        rep = _lib.AsacGruRep()
        rep.shape = cs
        rep.params, rep.params_target, rep.m, rep.v = ptr(self.rep_p), ptr(self.rep_t), ptr(self.rep_m), ptr(self.rep_v)
        rep.obs, rep.h0, rep.h0_b_stride = ptr(self.rb['obs']), ptr(self.rb['h0']), rep_layers * H
        for k in ('states', 'states_post', 'target_states', 'hn', 'hn_post', 'save', 'grad_part', 'grad'):
            setattr(rep, k, ptr(self.rb[k]))
        rep.rep_tiles = rt
        self.rep = rep

    def load_rep(self, rep_sd, rep_target_sd):
        self.rep_p.copy_(lowering.gru_flat_from_state_dict(self.gshape, rep_sd))
        self.rep_t.copy_(lowering.gru_flat_from_state_dict(self.gshape, rep_target_sd))

    def make_rep_batch(self, b, noise) -> _lib.AsacSacBatch:
        """b: oracle SacRepBatch."""
        from oracle.sac_oracle import SacBatch
        self.rb['obs'].copy_(b.obs.float())
        self.rb['h0'].copy_(b.hidden0.float())
        fake = SacBatch(states=torch.zeros(1), actions=b.actions, rewards=b.rewards, dones=b.dones,
                        mu_probs=b.mu_probs, last_masks=b.last_masks, padding_masks=b.padding_masks,
                        priority_is=b.priority_is)
        batch = self.make_batch(fake, noise)
        batch.states, batch.states_post = ptr(self.rb['states']), ptr(self.rb['states_post'])
        batch.target_states = ptr(self.rb['target_states'])
        return batch

    def step_rep(self, batch) -> dict:
        check(self.lib.asac_sac_step_networks_rep(C.byref(self.cfg), C.byref(self.prm), C.byref(batch),
                                                  C.byref(self.work), C.byref(self.rep), 1, None, self._s()),
              'sac_step_networks_rep')
        check(self.lib.asac_sac_staged_tail(C.byref(self.cfg), C.byref(self.prm), C.byref(self.work), self._s()),
              'sac_staged_tail')
        hp = self.hp
        out = {'y': self.wk['y'].cpu().numpy(), 'grad_q': [self.grad_q_dict(i) for i in range(hp.ensemble_q_num)],
               'grad_policy': self.grad_pi_dict(),
               'grad_rep': {k: t.cpu().numpy() for k, t in
                            lowering.gru_state_dict_from_flat(self.gshape, self.rb['grad']).items()},
               'states': self.rb['states'].cpu().numpy(), 'states_post': self.rb['states_post'].cpu().numpy(),
               'target_states': self.rb['target_states'].cpu().numpy(),
               'next_hidden': self.rb['hn_post'][:, :-1].cpu().numpy(),
               'grad_state': self.wk['grad_state'].cpu().numpy()}
        if hp.use_n_step_is:
            out['pi_probs'] = self.wk['pi_probs'].cpu().numpy()
        if hp.use_auto_alpha:
            out['grad_log_alpha'] = self.wk['grad_alpha'].cpu().numpy()
        if hp.use_priority:
            out['td_error'] = self.wk['td_error'].cpu().numpy()
            out['y_td'] = self.wk['y_td'].cpu().numpy()
        return out

    def sync_from_oracle(self, oracle, what=('q', 'qt', 'pi', 'alpha', 'rep')) -> None:
        """SacCuda.sync_from_oracle + the representation's parameters, target copy and Adam state."""
        super().sync_from_oracle(oracle, tuple(w for w in what if w != 'rep'))
        if 'rep' in what:
            flat = lambda d: lowering.gru_flat_from_state_dict(self.gshape, {k: v.detach().float() for k, v in d.items()})
            self.rep_p.copy_(flat(oracle.rep))
            self.rep_t.copy_(flat(oracle.rep_target))
            st = oracle.opt_rep.state
            mom = lambda key: {k: (st[p][key].detach() if p in st else torch.zeros_like(p)) for k, p in oracle.rep.items()}
            self.rep_m.copy_(flat(mom('exp_avg')))
            self.rep_v.copy_(flat(mom('exp_avg_sq')))
            steps = {int(float(st[p]['step'])) for p in oracle.rep.values() if p in st}
            self.counters[4] = steps.pop() if steps else 0

    def snapshot(self) -> dict:
        d = super().snapshot()
        for k, t in lowering.gru_state_dict_from_flat(self.gshape, self.rep_p).items():
            d[f'rep.{k}'] = t.cpu().numpy()
        for k, t in lowering.gru_state_dict_from_flat(self.gshape, self.rep_t).items():
            d[f'rept.{k}'] = t.cpu().numpy()
        return d
