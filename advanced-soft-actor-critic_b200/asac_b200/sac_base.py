"""B200-native SAC learner with the reference's ``SAC_Base`` API.

Mirrors ``algorithm/sac_base.py`` of the reference for the hot path named in BASELINE.json:
same constructor keywords (sac_base.py:22-93), ``train() -> int`` (:2496-2609),
``put_episode`` (:2303-2349), ``choose_action`` (:968-1019), checkpoint key names (:493-566).
The per-batch update is NOT a torch op graph: sampling, window gather + padding, noise,
Polyak, ``_get_y``, critic / policy / alpha losses with hand-derived backward passes, Adam,
``get_l_probs``, ``_get_td_error``, the priority update and the mu-prob write-back are kernels of
``libasac_b200.so`` enqueued on one stream and (by default) replayed as a single CUDA graph.

Scope: continuous actions, stock ``ModelQ`` / ``ModelPolicy`` topologies, and a
representation that is either ``ModelSimpleRep`` (vector observations) or ONE stock ``GRU`` over
``cat[obs, pre_action]`` (``seq_encoder=SEQ_ENCODER.RNN``, envs/test/nn_rnn.py) trained through the
critic loss by ``csrc/rep_gru.cu``.  Anything else raises ``NotImplementedError`` at construction —
there is no silent torch or CPU fallback for the update path.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import random
from pathlib import Path

import numpy as np
import torch

from . import _lib, lowering
from . import dist as adist
from . import nn_models as m
from ._lib import check, ptr
from .replay_buffer import PrioritizedReplayBuffer


class _FlatAdam:
    """``torch.optim.Adam``-shaped view (``state_dict`` / ``load_state_dict``) of flat moment
    buffers that the CUDA Adam kernel updates (sac_base.py:296-300)."""

    def __init__(self, params: list[torch.nn.Parameter], flat_param, m_flat: torch.Tensor | None,
                 v_flat: torch.Tensor | None, counters: torch.Tensor, counter_index: int, lr: float):
        """``flat_param``: one flat buffer (with ``m_flat`` / ``v_flat``) or a list of (flat, m, v) triples — a module
        whose parameters live in two buffers (continuous net + discrete heads) still has ONE optimizer."""
        self._params, self._lr = params, lr
        self._counters, self._ci = counters, counter_index
        triples = flat_param if isinstance(flat_param, list) else [(flat_param, m_flat, v_flat)]
        self._views = []
        for p in params:
            for flat, mf, vf in triples:
                off = (p.data_ptr() - flat.data_ptr()) // 4
                if 0 <= off and off + p.numel() <= flat.numel() and flat.numel() > 0:
                    mf, vf = mf.reshape(-1), vf.reshape(-1)
                    self._views.append((mf[off:off + p.numel()].view(p.shape), vf[off:off + p.numel()].view(p.shape)))
                    break
            else:
                raise AssertionError('parameter outside the flat buffers of its optimizer')

    def state_dict(self) -> dict:
        step = float(self._counters[self._ci].item())
        state = {}
        if step > 0:
            for i, (mm, vv) in enumerate(self._views):
                state[i] = {'step': torch.tensor(step), 'exp_avg': mm.clone(), 'exp_avg_sq': vv.clone()}
        group = dict(lr=self._lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False,
                     foreach=None, capturable=False, differentiable=False, fused=None,
                     params=list(range(len(self._views))))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd: dict) -> None:
        steps = set()
        for i, (mm, vv) in enumerate(self._views):
            st = sd['state'].get(i)
            if st is None:
                mm.zero_(); vv.zero_()
                continue
            mm.copy_(st['exp_avg']); vv.copy_(st['exp_avg_sq'])
            steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise RuntimeError('per-parameter Adam step counts differ; not representable')
        self._counters[self._ci] = steps.pop() if steps else 0


class SAC_Base:
    _closed = False

    def __init__(self,
                 obs_names: list[str],
                 obs_shapes: list[tuple[int]],
                 d_action_sizes: list[int],
                 c_action_size: int,
                 model_abs_dir: Path | None,
                 nn,

                 device: str | None = None,
                 ma_name: str | None = None,
                 summary_path: str | None = 'log',
                 train_mode: bool = True,
                 last_ckpt: str | None = None,

                 nn_config: dict | None = None,

                 seed: float | None = None,
                 write_summary_per_step: float = 1e3,
                 save_model_per_step: float = 1e5,

                 use_replay_buffer: bool = True,
                 use_priority: bool = True,

                 ensemble_q_num: int = 2,
                 ensemble_q_sample: int = 2,

                 burn_in_step: int = 0,
                 n_step: int = 1,
                 seq_encoder=None,

                 batch_size: int = 256,
                 tau: float = 0.005,
                 update_target_per_step: int = 1,
                 init_log_alpha: float = -2.3,
                 use_auto_alpha: bool = True,
                 target_d_alpha: float = 0.98,
                 target_c_alpha: float = 1.,
                 d_policy_entropy_penalty: float = 0.5,

                 learning_rate: float = 3e-4,

                 gamma: float = 0.99,
                 v_lambda: float = 1.,
                 v_rho: float = 1.,
                 v_c: float = 1.,
                 clip_epsilon: float = 0.2,

                 discrete_dqn_like: bool = False,
                 discrete_dqn_epsilon: float = 0.2,
                 use_n_step_is: bool = True,

                 siamese=None,
                 siamese_use_q: bool = False,
                 siamese_use_adaptive: bool = False,

                 use_prediction: bool = False,
                 transition_kl: float = 0.8,
                 use_extra_data: bool = True,

                 curiosity=None,
                 curiosity_strength: float = 1.,
                 use_rnd: bool = False,
                 rnd_n_sample: int = 10,

                 use_normalization: bool = False,

                 offline_enabled: bool = False,
                 offline_loss: bool = False,

                 action_noise: list[float] | None = None,

                 replay_config: dict | None = None,

                 use_cuda_graph: bool = True):
        self._lib = _lib.load()  # fails loudly when libasac_b200.so is missing
        if not torch.cuda.is_available():
            raise _lib.AsacError('asac_b200.SAC_Base needs a CUDA device: the update path has no CPU fallback')

        self.obs_names = obs_names
        self.obs_shapes = obs_shapes
        self.d_action_sizes = d_action_sizes
        self.d_action_summed_size = sum(d_action_sizes)
        self.d_action_branch_size = len(d_action_sizes)
        self.c_action_size = c_action_size
        self.model_abs_dir = model_abs_dir
        self.ma_name = ma_name
        self.train_mode = train_mode

        self.use_replay_buffer = use_replay_buffer
        self.use_priority = use_priority
        self.ensemble_q_num = ensemble_q_num
        self.ensemble_q_sample = ensemble_q_sample
        self.burn_in_step = burn_in_step
        self.n_step = n_step
        self.seq_encoder = seq_encoder

        self.write_summary_per_step = int(write_summary_per_step)
        self.save_model_per_step = int(save_model_per_step)
        self.batch_size = batch_size
        self.tau = tau
        self.update_target_per_step = update_target_per_step
        self.use_auto_alpha = use_auto_alpha
        self.target_d_alpha = target_d_alpha
        self.target_c_alpha = target_c_alpha
        self.d_policy_entropy_penalty = d_policy_entropy_penalty
        self.learning_rate = learning_rate
        self.gamma = gamma
        self.v_lambda = v_lambda
        self.v_rho = v_rho
        self.v_c = v_c
        self.clip_epsilon = clip_epsilon
        self.use_n_step_is = use_n_step_is
        self.action_noise = action_noise
        self.use_cuda_graph = use_cuda_graph
        self.discrete_dqn_like = bool(discrete_dqn_like) and bool(d_action_sizes)
        self.discrete_dqn_epsilon = discrete_dqn_epsilon
        self._init_log_alpha = init_log_alpha

        unsupported = {
            'neither discrete nor continuous actions': not d_action_sizes and not c_action_size,
            'siamese': siamese is not None,
            'use_prediction': use_prediction,
            'curiosity': curiosity is not None,
            'use_rnd': use_rnd,
            'use_normalization': use_normalization,
            'offline_enabled': offline_enabled,
            'ensemble_q_sample > ensemble_q_num': ensemble_q_sample > ensemble_q_num,
            'ensemble_q_sample < ensemble_q_num with discrete action branches': bool(d_action_sizes) and
            ensemble_q_sample != ensemble_q_num,
            'action_noise': action_noise is not None,
        }
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError('outside the B200 hot path (SURVEY.md §8): ' + ', '.join(bad))

        if device is None:
            device = f'cuda:{torch.cuda.current_device()}'
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.AsacError(f"device '{device}': the B200 learner runs on CUDA only")
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())

        self._logger = logging.getLogger('sac.base' if ma_name is None else f'sac.base.{ma_name}')
        self._seed = seed
        if seed is not None:
            torch.manual_seed(int(seed))  # sac_base.py:246-248 seeds torch only
        torch.distributions.Distribution.set_default_validate_args(False)

        self.summary_writer = None
        if model_abs_dir is not None and summary_path is not None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                path = Path(model_abs_dir).joinpath(summary_path)
                self.summary_writer = SummaryWriter(str(path if ma_name is None else path.joinpath(ma_name)))
            except Exception as e:  # tensorboard is optional
                self._logger.warning(f'no summary writer: {e}')
        self.summary_available = False

        self._rank, self._world = adist.world()
        with torch.cuda.device(self.device):
            self._build_model(nn, nn_config, init_log_alpha)
            self._build_ckpt()
            self._init_replay_buffer(replay_config)
            self._build_step_buffers()
            self._init_or_restore(int(last_ckpt) if last_ckpt is not None else None)
        self._graphs = [None, None]
        self._bridge_graph, self._bridge_eager_steps = None, 0
        self._graph_columns_key = None
        self._steps_since_check = 0
        # NCCL all-reduces are captured into the step's CUDA graph (ASAC_GRAPH_COLLECTIVES=0 keeps them eager)
        self._graph_collectives = os.environ.get('ASAC_GRAPH_COLLECTIVES', '1') != '0'
        if self._world > 1:
            # replicated weights: every rank starts from rank 0's networks and optimizer state
            adist.broadcast_([self._q_flat, self._qt_flat, self._pi_flat, self._log_alpha_buf, self._q_m, self._q_v,
                              self._pi_m, self._pi_v, self._alpha_m, self._alpha_v, self._counters])
            if self._bridge is not None:
                adist.broadcast_([self._rep_flat, self._rept_flat, self._rep_m, self._rep_v])
            if self._gru is not None:
                if self._peer_table is None or not (self.use_priority and self.batch_size <= 1024):
                    raise NotImplementedError('data-parallel learner with a trained representation needs the '
                                              'peer-memory gradient exchange and the fused tail (PER, batch <= 1024)')
                adist.broadcast_([self._rep_flat, self._rept_flat, self._rep_m, self._rep_v])
            self._noise_seed ^= 0x9E3779B97F4A7C15 * (self._rank + 1) & 0x3FFFFFFFFFFFFFFF
            self.replay_buffer._seed ^= 0xD1B54A32D192ED03 * (self._rank + 1) & 0x3FFFFFFFFFFFFFFF

    # ------------------------------------------------------------------ construction
    def _build_model(self, nn, nn_config: dict | None, init_log_alpha: float) -> None:
        nn_config = {} if nn_config is None else nn_config
        nn_config = {k: ({} if nn_config.get(k) is None else nn_config[k]) for k in ('rep', 'policy')}
        dev = self.device
        A = self.c_action_size
        D = self.d_action_summed_size
        # Discrete-only runs (A == 0) keep the continuous structures of the step at width 1 — buffers and config only,
        # no continuous kernel is ever launched for them (self._c_enabled) — so that one code path sizes everything.
        self._c_enabled = A > 0
        Ae = max(A, 1)

        self._gamma_ratio = torch.logspace(0, self.n_step - 1, self.n_step, self.gamma)      # sac_base.py:285
        self._lambda_ratio = torch.logspace(0, self.n_step - 1, self.n_step, self.v_lambda)  # sac_base.py:286
        # sac_base.py:291-294: first option of every discrete branch, zeros for the continuous part
        self._np_padding_action = np.concatenate([np.eye(k, dtype=np.float32)[0] for k in self.d_action_sizes] +
                                                 [np.zeros(A, dtype=np.float32)])
        self._padding_action_full = torch.from_numpy(self._np_padding_action).to(dev)
        self._padding_action = torch.zeros(Ae, dtype=torch.float32, device=dev)

        self.model_rep = nn.ModelRep(self.obs_names, self.obs_shapes, self.d_action_sizes, A, False,
                                     self.model_abs_dir, **nn_config['rep']).to(dev)
        self.model_target_rep = nn.ModelRep(self.obs_names, self.obs_shapes, self.d_action_sizes, A, True,
                                            self.model_abs_dir, **nn_config['rep']).to(dev)
        for p in self.model_target_rep.parameters():
            p.requires_grad = False
        f32 = dict(dtype=torch.float32, device=dev)
        # counters: global_step, Adam steps of critics / policy / alpha / representation
        self._counters = torch.zeros(8, dtype=torch.int64, device=dev)
        self._is_attn = self.seq_encoder is not None and getattr(self.seq_encoder, 'name', str(self.seq_encoder)) == 'ATTN'
        self.optimizer_rep = None
        self._gru = None
        self._bridge = None
        self._vector_obs = [(name, shape) for name, shape in zip(self.obs_names, self.obs_shapes) if len(shape) == 1]
        try:
            if self._is_attn:
                raise lowering.NotStockNetwork('attention representation')
            lowered = lowering.analyze_rep(self.model_rep, self.obs_shapes, A)
            if lowered is not None and os.environ.get('ASAC_REP_BRIDGE', '1') == '2':  # tests: stock GRU through torch
                raise lowering.NotStockNetwork('ASAC_REP_BRIDGE=2')
            if lowered is not None and self.d_action_sizes:  # the fused GRU step has no discrete stages
                raise lowering.NotStockNetwork('stock GRU with discrete action branches')
        except lowering.NotStockNetwork as e:
            if os.environ.get('ASAC_REP_BRIDGE', '1') == '0':
                raise
            # any other ModelRep: the plugin's torch modules produce the states, the kernels do the rest
            from .rep_bridge import TorchRepBridge
            self._logger.info(f'representation runs as the plugin\'s torch module ({e})')
            lowered = None
            self._bridge = TorchRepBridge(self.model_rep, self.model_target_rep, self._is_attn, self.burn_in_step, dev)
        if self._bridge is not None:
            br = self._bridge
            self.state_size, self.seq_hidden_state_shape = br.probe(self.obs_shapes, A)
            self._rep_flat, self._rept_flat, self._rep_m, self._rep_v = br.flat, br.flat_target, br.m, br.v
            self.optimizer_rep = _FlatAdam(br.params, br.flat, br.m, br.v, self._counters, 4, self.learning_rate)
        elif lowered is None:
            if self.seq_encoder is not None:
                raise NotImplementedError('seq_encoder is set but ModelRep is ModelSimpleRep')
            self.state_size = sum(shape[0] for _, shape in self._vector_obs)
            if self.state_size == 0:
                raise NotImplementedError('ModelSimpleRep needs at least one vector observation')
            self.seq_hidden_state_shape = (0,)
        else:
            if self.seq_encoder is None:
                raise NotImplementedError('a recurrent ModelRep needs seq_encoder=SEQ_ENCODER.RNN')
            self._gru, rep_params = lowered
            target = lowering.analyze_rep(self.model_target_rep, self.obs_shapes, A)
            if target is None or target[0] != self._gru:
                raise lowering.NotStockNetwork('target representation differs from the online one')
            P = self._gru.stride
            self._rep_flat, self._rept_flat = torch.zeros(P, **f32), torch.zeros(P, **f32)
            self._rep_m, self._rep_v = torch.zeros(P, **f32), torch.zeros(P, **f32)
            lowering.bind_parameters(rep_params, self._rep_flat)
            lowering.bind_parameters(target[1], self._rept_flat)
            self.state_size = self._gru.hidden
            self.seq_hidden_state_shape = (self._gru.layers, self._gru.hidden)
            self._gru_c = _lib.AsacGruShape(self._gru.obs_size, A, self._gru.hidden, self._gru.layers)
            self._probe_rep()
            self.optimizer_rep = _FlatAdam(rep_params, self._rep_flat, self._rep_m, self._rep_v, self._counters, 4,
                                           self.learning_rate)

        E = self.ensemble_q_num
        self.model_q_list = [nn.ModelQ(self.state_size, self.d_action_sizes, A, False, self.model_abs_dir).to(dev)
                             for _ in range(E)]
        self.model_target_q_list = [nn.ModelQ(self.state_size, self.d_action_sizes, A, True,
                                              self.model_abs_dir).to(dev) for _ in range(E)]
        self.model_policy = nn.ModelPolicy(self.state_size, self.d_action_sizes, A, self.model_abs_dir,
                                           **nn_config['policy']).to(dev)
        for q in self.model_target_q_list:
            for p in q.parameters():
                p.requires_grad = False

        # ---- lower the stock nets onto flat buffers shared with the kernels
        q_shapes = [lowering.analyze_q(q) for q in self.model_q_list + self.model_target_q_list]
        self._pi_shape, pi_params = lowering.analyze_policy(self.model_policy)
        self._q_shape = q_shapes[0][0]
        if any(s != self._q_shape for s, _ in q_shapes):
            raise lowering.NotStockNetwork('ensemble members differ in shape')
        if not self._c_enabled:  # placeholders for the sizes below; never launched
            self._q_shape = lowering.NetShape(self.state_size + 1, 64, 1, 1)
            self._pi_shape = lowering.NetShape(self.state_size, 64, 1, 2)
        Pq, Ppi = self._q_shape.stride, self._pi_shape.stride
        self._q_flat = torch.zeros(E, Pq, **f32)
        self._qt_flat = torch.zeros(E, Pq, **f32)
        self._pi_flat = torch.zeros(Ppi, **f32)
        for i in range(E):
            lowering.bind_parameters(q_shapes[i][1], self._q_flat[i])
            lowering.bind_parameters(q_shapes[E + i][1], self._qt_flat[i])
        lowering.bind_parameters(pi_params, self._pi_flat)
        self._q_m, self._q_v = torch.zeros_like(self._q_flat), torch.zeros_like(self._q_flat)
        self._pi_m, self._pi_v = torch.zeros_like(self._pi_flat), torch.zeros_like(self._pi_flat)
        self._disc = None
        if self.d_action_sizes:
            if self._gru is not None:
                raise AssertionError('a representation with discrete branches runs as the torch module')
            if self._world > 1:
                raise NotImplementedError('discrete action branches in the data-parallel learner')
            from .discrete import DiscreteBranch
            self._disc = DiscreteBranch(self, self.model_q_list, self.model_target_q_list, self.model_policy)

        self.log_d_alpha = self._disc.log_alpha[0] if self._disc is not None else \
            torch.tensor(init_log_alpha, dtype=torch.float32, device=dev)
        self.log_c_alpha = torch.full((1,), init_log_alpha, **f32)[0]  # 0-dim view of a 1-element buffer
        self._log_alpha_buf = self.log_c_alpha.view(1)
        self._alpha_m, self._alpha_v = torch.zeros(1, **f32), torch.zeros(1, **f32)
        self.global_step = torch.tensor(0, dtype=torch.int64)
        self._host_step = 0

        lr = self.learning_rate
        dq = self._disc
        flats_q = lambda i: [(self._q_flat[i], self._q_m[i], self._q_v[i])] + \
            ([(dq.q[i], dq.q_m[i], dq.q_v[i])] if dq is not None else [])
        self.optimizer_q_list = [_FlatAdam(list(q.parameters()), flats_q(i), None, None, self._counters, 1, lr)
                                 for i, q in enumerate(self.model_q_list)]
        self.optimizer_policy = _FlatAdam(list(self.model_policy.parameters()),
                                          [(self._pi_flat, self._pi_m, self._pi_v)] +
                                          ([(dq.pi, dq.pi_m, dq.pi_v)] if dq is not None else []), None, None,
                                          self._counters, 2, lr)
        self.optimizer_alpha = None
        if self.use_auto_alpha:
            self.optimizer_alpha = _AlphaAdam(self._alpha_m, self._alpha_v, self._counters, lr,
                                              d_state=(dq.alpha_m, dq.alpha_v) if (dq is not None and not self.discrete_dqn_like) else None,
                                              c_enabled=self._c_enabled)

        # ---- C structs
        cfg = _lib.AsacSacConfig()
        cfg.learning_rate = float(lr)
        cfg.batch, cfg.burn_in, cfg.n_step = self.batch_size, self.burn_in_step, self.n_step
        cfg.seq_len = self.burn_in_step + self.n_step + 1
        cfg.state_size, cfg.action_size, cfg.ensemble = self.state_size, Ae, E
        cfg.q_hidden, cfg.q_depth = self._q_shape.hidden, self._q_shape.depth
        cfg.pi_hidden, cfg.pi_depth = self._pi_shape.hidden, self._pi_shape.depth
        cfg.use_n_step_is, cfg.use_priority = int(self.use_n_step_is), int(self.use_priority)
        cfg.use_auto_alpha = int(self.use_auto_alpha)
        cfg.update_target_per_step = int(self.update_target_per_step)
        cfg.bn_stride = cfg.seq_len
        cfg.rep_kind = 0 if (self._gru is None and self._bridge is None) else 1
        cfg.ensemble_sample = self.ensemble_q_sample if 0 < self.ensemble_q_sample < E else 0
        cfg.rep_param_stride = 0 if self._gru is None else self._gru.stride
        cfg.tau, cfg.one_minus_tau = float(self.tau), float(np.float32(1. - self.tau))
        cfg.gamma, cfg.v_rho, cfg.v_c = float(self.gamma), float(self.v_rho), float(self.v_c)
        cfg.clip_epsilon, cfg.target_c_alpha = float(self.clip_epsilon), float(self.target_c_alpha)
        if self.n_step > _lib.MAX_NSTEP:
            raise NotImplementedError(f'n_step > {_lib.MAX_NSTEP}')
        for k in range(self.n_step):
            cfg.gamma_ratio[k] = float(self._gamma_ratio[k])
            cfg.lambda_ratio[k] = float(self._lambda_ratio[k])
        self._cfg = cfg
        self._cfg_d = _lib.AsacSacConfig.from_buffer_copy(cfg)  # what the discrete kernels see: the true action width
        self._cfg_d.action_size = A
        tile = self._lib.asac_sac_tile_batch(C.byref(cfg))
        if tile < 1:
            check(tile, 'asac_sac_tile_batch')
        self._n_tiles = (self.batch_size + tile - 1) // tile

        prm = _lib.AsacSacParams()
        prm.q, prm.q_target, prm.pi = ptr(self._q_flat), ptr(self._qt_flat), ptr(self._pi_flat)
        prm.log_alpha = ptr(self._log_alpha_buf)
        prm.q_m, prm.q_v, prm.pi_m, prm.pi_v = ptr(self._q_m), ptr(self._q_v), ptr(self._pi_m), ptr(self._pi_v)
        prm.alpha_m, prm.alpha_v = ptr(self._alpha_m), ptr(self._alpha_v)
        prm.counters = ptr(self._counters)
        self._prm = prm

    def _gru_forward(self, obs0: torch.Tensor, pre_action: torch.Tensor, h0: torch.Tensor | None):
        """asac_gru_forward with the online representation: obs0 [rows, l, So], pre_action [rows, l, A],
        h0 [rows, layers, H] -> (state [rows, l, H], hn [rows, l, layers, H])."""
        g = self._gru
        rows, l = int(obs0.shape[0]), int(obs0.shape[1])
        f32 = dict(dtype=torch.float32, device=self.device)
        state, hn = torch.empty(rows, l, g.hidden, **f32), torch.empty(rows, l, g.layers, g.hidden, **f32)
        net = _lib.AsacGruNet(ptr(self._rep_flat), ptr(state), ptr(hn), None)
        check(self._lib.asac_gru_forward(C.byref(self._gru_c), C.byref(net), 1, ptr(obs0.contiguous()), None, 0,
                                         ptr(pre_action.contiguous()), ptr(None if h0 is None else h0.contiguous()),
                                         g.layers * g.hidden, rows, l, _lib.current_stream()), 'gru_forward')
        return state, hn

    @torch.no_grad()
    def _probe_rep(self) -> None:
        """The structural lowering (lowering.analyze_rep) only proves that the plugin's ModelRep OWNS one
        stock GRU; that its forward IS ``GRU(cat[obs_list[0], pre_action], pre_seq_hidden_state[:, 0])``
        is checked here by running the plugin's torch forward and the kernel on the same random
        inputs (with an initial state, and with None as at sac_base.py:353-357)."""
        g, dev = self._gru, self.device
        gen = torch.Generator(device='cpu').manual_seed(1234)
        rows, l = 5, 3
        obs_list = [torch.randn(rows, l, *shape, generator=gen).to(dev) for shape in self.obs_shapes]
        pre_action = torch.rand(rows, l, self.c_action_size, generator=gen).to(dev)
        hidden = (torch.randn(rows, l, g.layers, g.hidden, generator=gen) * 0.5).to(dev)
        import warnings
        for h in (hidden, None):
            try:
                # cuDNN notes that the weights are views of one flat buffer; its RNNs default to TF32
                with warnings.catch_warnings(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    warnings.simplefilter('ignore')
                    want_state, want_hn = self.model_rep(obs_list, pre_action, h)
            except Exception as e:  # noqa: BLE001
                raise lowering.NotStockNetwork(f'ModelRep.forward failed on the probe inputs: {e}') from e
            state, hn = self._gru_forward(obs_list[0], pre_action, None if h is None else h[:, 0])
            if want_state.shape != state.shape or want_hn.shape != hn.shape or \
                    not torch.allclose(want_state, state, atol=2e-5, rtol=0) or \
                    not torch.allclose(want_hn, hn, atol=2e-5, rtol=0):
                raise lowering.NotStockNetwork(
                    'ModelRep.forward is not GRU(cat[obs_list[0], pre_action], pre_seq_hidden_state[:, 0]) '
                    '(envs/test/nn_rnn.py form); other recurrent representations are outside the fused path')

    def _build_ckpt(self) -> None:
        """Same key names as sac_base.py:493-566 so .pth files interchange."""
        ck = {'global_step': self.global_step}
        if self.optimizer_rep is not None:  # sac_base.py:506-509
            ck['model_rep'] = self.model_rep
            ck['model_target_rep'] = self.model_target_rep
            ck['optimizer_rep'] = self.optimizer_rep
        for i in range(self.ensemble_q_num):
            ck[f'model_q_{i}'] = self.model_q_list[i]
            ck[f'model_target_q_{i}'] = self.model_target_q_list[i]
            ck[f'optimizer_q_{i}'] = self.optimizer_q_list[i]
        ck['model_policy'] = self.model_policy
        ck['optimizer_policy'] = self.optimizer_policy
        ck['log_d_alpha'] = self.log_d_alpha
        ck['log_c_alpha'] = self.log_c_alpha
        if self.use_auto_alpha:
            ck['optimizer_alpha'] = self.optimizer_alpha
        self.ckpt_dict = ck

    def _init_replay_buffer(self, replay_config: dict | None) -> None:
        if not self.train_mode:
            return
        if not self.use_replay_buffer:  # on-policy: every window once, in shuffled batches (sac_base.py:642-646)
            if self._world > 1:
                raise NotImplementedError('data-parallel learner without a replay buffer')
            from .batch_buffer import BatchBuffer
            self.batch_buffer = BatchBuffer(burn_in_step=self.burn_in_step, n_step=self.n_step,
                                            padding_action=self._np_padding_action, batch_size=self.batch_size,
                                            device=self.device)
            self._cfg.use_priority = 0  # priority_is is None without a replay buffer (sac_base.py:2553)
            return
        replay_config = {} if replay_config is None else dict(replay_config)
        if self._seed is not None:
            replay_config.setdefault('seed', int(self._seed))
        # Data-parallel learner: `capacity` is the GLOBAL capacity; every rank owns capacity / world ring slots with
        # its own segment tree (asac_b200/dist.py).  `episode_routing`: 'local' — every rank stores the episodes its
        # own actors hand to put_episode (the usual one-process-per-GPU deployment); 'round_robin' — every rank is
        # handed EVERY episode (one replicated actor stream) and keeps the ones dist.episode_owner assigns to it.
        self._episode_routing = replay_config.pop('episode_routing', os.environ.get('ASAC_EPISODE_ROUTING', 'local'))
        if self._episode_routing not in ('local', 'round_robin'):
            raise ValueError(f"episode_routing '{self._episode_routing}' (local | round_robin)")
        self._episode_counter = 0
        self.global_replay_capacity = None
        if self._world > 1 and replay_config.pop('shard', True):
            self.global_replay_capacity = int(replay_config.get('capacity', 524288))
            replay_config['capacity'] = adist.shard_capacity(self.global_replay_capacity, self._world)
        self.replay_buffer = PrioritizedReplayBuffer(batch_size=self.batch_size, sample_prev_n=self.burn_in_step,
                                                     sample_post_n=self.n_step, device=self.device,
                                                     logger_parent_name=self._logger.name, **replay_config)
        self.replay_buffer._flush_hook = self.flush_priority_update
        self._cfg.td_error_min = float(self.replay_buffer.td_error_min)
        self._cfg.td_error_max = float(self.replay_buffer.td_error_max)
        self._cfg.per_alpha = float(self.replay_buffer.alpha)

    def _build_step_buffers(self) -> None:
        """Persistent device buffers of one train() call (static addresses -> CUDA graph)."""
        B, L, S, A, E = self.batch_size, self._cfg.seq_len, self.state_size, self._cfg.action_size, self.ensemble_q_num
        n, T = self.n_step, self._n_tiles
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        Pq, Ppi = self._q_shape.stride, self._pi_shape.stride
        wk = self._wk = {
            'y': torch.zeros(B, **f32), 'tq': torch.zeros(E, B, **f32), 'q_val': torch.zeros(E, B, **f32),
            'loss_q': torch.zeros(T, E, **f32), 'grad_q_part': torch.zeros(T, E, Pq, **f32),
            'grad_q': torch.zeros(E, Pq, **f32), 'grad_pi_part': torch.zeros(T, Ppi, **f32),
            'grad_pi': torch.zeros(Ppi, **f32), 'stats_pi': torch.zeros(T, 2, **f32),
            'grad_alpha_part': torch.zeros(T, 2, **f32), 'grad_alpha': torch.zeros(1, **f32),
            'pi_probs': torch.zeros(B, L - 1, A, **f32), 'post_parts': torch.zeros(B, 2 + E, **f32), 'y_td': torch.zeros(B, **f32),
            'td_error': torch.zeros(B, **f32),
        }
        if self._gru is not None or self._bridge is not None:
            wk['grad_state'] = torch.zeros(E, B, S, **f32)
        if self._disc is not None:
            wk['pi_probs_full'] = torch.ones(B, L - 1, self._disc.AF, **f32)
        work = _lib.AsacSacWork()
        work.n_tiles = T
        for k, t in wk.items():
            setattr(work, k, ptr(t))
        self._work = work
        self._rw = None
        if self._gru is not None:
            g = self._gru
            rtile = self._lib.asac_gru_backward_tile(C.byref(self._gru_c), self.burn_in_step)
            if rtile < 1:
                check(rtile, 'asac_gru_backward_tile')
            self._rw = {'hn': torch.zeros(B, L, g.layers, g.hidden, **f32),
                        'hn_post': torch.zeros(B, L, g.layers * g.hidden, **f32),
                        'save': torch.zeros(B, L, g.layers, 4 * g.hidden, **f32),
                        'grad_part': torch.zeros(B, g.stride, **f32),  # one partial gradient per sequence
                        'grad': torch.zeros(g.stride, **f32)}
        # The sampled batch lives in a "batch set" (sample outputs, gathered windows, the C structs pointing
        # at them).  Two sets: while the networks train on one, the NEXT step's sample + gather fill the
        # other on a parallel branch (ASAC_SAMPLE_AHEAD=0: one set, sample and gather on the critical path).
        # the step's torch.randperm(E) draws (rows 0-4: ensemble_q_sample < ensemble_q_num, see AsacSacBatch;
        # rows 5-8: (target, online) shuffles of the DQN-like target in _train_rep_q and in _get_td_error)
        self._ens_perms = torch.zeros(9, E, dtype=torch.int32, device=dev)
        self._sample_ahead = self.use_replay_buffer and os.environ.get('ASAC_SAMPLE_AHEAD', '1') != '0' \
            and self._bridge is None and self._disc is None  # (those steps run in program order on one stream)
        self._sets = [self._make_batch_set() for _ in range(2 if self._sample_ahead else 1)]
        self._cur, self._primed = 0, False
        # With sampling one step ahead the tree update of step N is only needed by the sample of step N + 2:
        # the fused tail then writes the td errors and the update itself runs first thing on the NEXT
        # step's ahead branch, off the critical path (same priorities seen by every batch as before;
        # ASAC_DEFER_TREE=0 keeps it inside the tail).  `_pending` = a step's tree update has not run yet.
        self._noise_ahead = self._sample_ahead and os.environ.get('ASAC_NOISE_AHEAD', '1') != '0'
        self._defer_tree = self._sample_ahead and os.environ.get('ASAC_DEFER_TREE', '1') != '0'
        self._pending = self._defer_active = False
        self._prefetch_stream = torch.cuda.Stream(device=dev)
        self._noise_seed = (int(self._seed) if self._seed is not None else random.getrandbits(62)) ^ 0x5AC5AC
        self._side_stream = torch.cuda.Stream(device=dev)
        # Optional (ASAC_L2_PREFETCH=1): pull the buffers the critical-path kernels touch first and nothing
        # else of the step has brought into L2 (policy parameters, Adam moments; the Polyak kernel reads the
        # critics) on the side branch.  Measured on B200, same box, L2 flushed between steps: 5352 vs 5348
        # steps/s, warm 6527 vs 6540 — the weight pipe already has every slot in flight behind ONE
        # HBM round trip per kernel, so there is nothing left to hide; off by default.
        self._prefetch = None
        if os.environ.get('ASAC_L2_PREFETCH', '0') == '1':
            regions = [self._pi_flat, self._q_m, self._q_v, self._pi_m, self._pi_v]
            if self._gru is not None:
                regions += [self._rep_flat, self._rep_m, self._rep_v]
            self._prefetch = ((C.c_void_p * len(regions))(*[t.data_ptr() for t in regions]),
                              (C.c_int64 * len(regions))(*[t.numel() * 4 for t in regions]), regions)
        self._min_prob = torch.zeros(1, dtype=torch.float64, device=dev)    # sharded replay: global min sampling probability
        self._act_counter = torch.zeros(1, dtype=torch.int64, device=dev)  # Philox counter of choose_action
        self._actor_bufs = {}
        self.actor_tensor_cores = False
        # data-parallel learner: gradient exchange inside the reduce+Adam kernels over NVLink peer memory
        # (ASAC_PEER_EXCHANGE=0, or a failed mapping, keeps the NCCL all-reduce between the kernels)
        self._peers = self._peer_table = None
        if self._world > 1 and os.environ.get('ASAC_PEER_EXCHANGE', '1') != '0':
            try:
                n_recv = self._lib.asac_peer_recv_words(C.byref(self._cfg), self._world)
                self._peers = adist.PeerGradientExchange(n_recv, dev)
                self._peer_table = self._peers.table()
            except Exception as e:  # noqa: BLE001 - any failure of the mapping falls back to NCCL
                self._logger.warning(f'peer-memory gradient exchange unavailable ({e}); using NCCL all-reduce')
                self._peers = self._peer_table = None
            # every rank must run the SAME exchange: one rank on NCCL while the others poll peer memory would
            # spin until the exchange's timeout
            if not adist.all_agree(self._peer_table is not None, dev):
                if self._peer_table is not None:
                    self._logger.warning('a peer rank could not map the exchange buffers; all ranks use NCCL')
                self._peers = self._peer_table = None
        self._ready_version = None  # replay version for which every rank was seen ready to step

    def _make_batch_set(self) -> dict:
        B, L, S, A = self.batch_size, self._cfg.seq_len, self.state_size, self._cfg.action_size
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        bt = {
            'index': torch.zeros(B, L, dtype=torch.int32, device=dev),
            'states': torch.zeros(B, L, S, **f32), 'actions': torch.zeros(B, L, A, **f32),
            'rewards': torch.zeros(B, L, **f32), 'dones': torch.zeros(B, L, **u8),
            'last_masks': torch.zeros(B, L, **u8), 'padding_masks': torch.zeros(B, L, **u8),
            'mu_probs': torch.zeros(B, L, A, **f32),
        }
        if self._disc is not None:
            bt['actions_full'] = torch.zeros(B, L, self._disc.AF, **f32)
            bt['mu_full'] = torch.zeros(B, L, self._disc.AF, **f32)
        smp = {'slots': torch.zeros(B, dtype=torch.int32, device=dev), 'ids': torch.zeros(B, dtype=torch.int64, device=dev),
               'p': torch.zeros(B, **f32), 'w': torch.zeros(B, **f32)}
        batch = _lib.AsacSacBatch()
        batch.states, batch.actions, batch.rewards = ptr(bt['states']), ptr(bt['actions']), ptr(bt['rewards'])
        batch.dones, batch.last_masks = ptr(bt['dones']), ptr(bt['last_masks'])
        batch.padding_masks, batch.mu_probs = ptr(bt['padding_masks']), ptr(bt['mu_probs'])
        batch.priority_is = ptr(smp['w']) if (self.use_priority and self.use_replay_buffer) else None
        # one noise buffer per batch set, four views: eps_y, eps_pi, eps_alpha, eps_td
        n = self.n_step
        sizes = [B * (n + 1) * A, B * A, B * A, B * (n + 1) * A]
        noise = torch.zeros(sum(sizes), **f32)
        offs = np.cumsum([0] + sizes)
        batch.eps_y, batch.eps_pi, batch.eps_alpha, batch.eps_td = [ptr(noise[offs[i]:offs[i + 1]]) for i in range(4)]
        if self._cfg.ensemble_sample:  # the step's torch.randperm draws, refreshed on the device every step
            batch.ensemble_perms = ptr(self._ens_perms)
        rep = None
        if self._gru is not None:
            g = self._gru
            NL, H = g.layers, g.hidden
            bt['obs'] = torch.zeros(B, L, g.obs_size, **f32)
            bt['hidden'] = torch.zeros(B, L, NL * H, **f32)
            bt['states_post'], bt['target_states'] = torch.zeros(B, L, S, **f32), torch.zeros(B, L, S, **f32)
            rep = _lib.AsacGruRep()
            rep.shape = self._gru_c
            rep.params, rep.params_target = ptr(self._rep_flat), ptr(self._rept_flat)
            rep.m, rep.v = ptr(self._rep_m), ptr(self._rep_v)
            rep.obs, rep.h0, rep.h0_b_stride = ptr(bt['obs']), ptr(bt['hidden']), L * NL * H
            rep.states, rep.states_post, rep.target_states = ptr(bt['states']), ptr(bt['states_post']), \
                ptr(bt['target_states'])
            for k, t in self._rw.items():
                setattr(rep, k, ptr(t))
            rep.rep_tiles = B
            batch.states_post, batch.target_states = ptr(bt['states_post']), ptr(bt['target_states'])
        if self._bridge is not None:
            # observation windows are allocated with the stored dtypes once the replay's columns exist (_gather_specs)
            hid = int(np.prod(self.seq_hidden_state_shape))
            bt['hidden'] = torch.zeros(B, L, max(hid, 0), **f32)
            bt['states_post'], bt['target_states'] = torch.zeros(B, L, S, **f32), torch.zeros(B, L, S, **f32)
            batch.states_post, batch.target_states = ptr(bt['states_post']), ptr(bt['target_states'])
        return {'bt': bt, 'smp': smp, 'batch': batch, 'rep': rep, 'specs': None, 'noise': noise}

    # the batch set of the most recent train() call (tests, bench and summaries read these)
    _bt = property(lambda self: self._sets[self._cur]['bt'])
    _smp = property(lambda self: self._sets[self._cur]['smp'])
    _batch = property(lambda self: self._sets[self._cur]['batch'])
    _rep = property(lambda self: self._sets[self._cur]['rep'])
    _specs = property(lambda self: self._sets[self._cur]['specs'])
    _noise = property(lambda self: self._sets[self._cur]['noise'])

    @property
    def _graph(self):
        return next((g for g in self._graphs if g is not None), None)

    def _init_or_restore(self, last_ckpt: int | None) -> None:
        """sac_base.py:568-629."""
        self.ckpt_dir = None
        fresh = True
        if self.model_abs_dir:
            self.ckpt_dir = ckpt_dir = Path(self.model_abs_dir).joinpath('model')
            ckpts = sorted(int(p.stem) for p in ckpt_dir.glob('*.pth')) if ckpt_dir.exists() else []
            ckpt_dir.mkdir(parents=True, exist_ok=True)
            if ckpts:
                fresh = False
                if last_ckpt is None or last_ckpt not in ckpts:
                    if last_ckpt is not None:
                        self._logger.warning(f'{last_ckpt} NOT IN {ckpts}, using {ckpts[-1]}')
                    last_ckpt = ckpts[-1]
                path = ckpt_dir.joinpath(f'{last_ckpt}.pth')
                restored = torch.load(path, map_location=self.device, weights_only=True)
                failed = False
                for name, obj in self.ckpt_dict.items():
                    if name not in restored:
                        self._logger.warning(f'{name} not in {last_ckpt}.pth')
                        continue
                    if isinstance(obj, torch.Tensor):
                        obj.copy_(restored[name].to(obj.device))  # in place: the kernels hold the address
                    else:
                        if failed and name.startswith('optimizer'):
                            continue
                        try:
                            obj.load_state_dict(restored[name])
                        except RuntimeError as e:
                            failed = True
                            self._logger.error(e)
                self._host_step = int(self.global_step.item())
                self._counters[0] = self._host_step
                self._logger.info(f'Restored from {path}')
                if self.train_mode and self.use_replay_buffer:
                    self.replay_buffer.load(ckpt_dir, last_ckpt)
        if fresh:
            self._logger.info('Initializing from scratch')
            self._update_target_variables()
        self.set_train_mode(self.train_mode)

    # ------------------------------------------------------------------ small API
    def _update_target_variables(self, tau=1.) -> None:
        """sac_base.py:745-764 (hard copy by default)."""
        with torch.cuda.device(self.device):
            check(self._lib.asac_sac_polyak(C.byref(self._cfg), C.byref(self._prm), float(tau),
                                            _lib.current_stream()), 'sac_polyak')
            if self._disc is not None:
                self._disc.polyak(force_tau=float(tau))
            if self._bridge is not None:
                check(self._lib.asac_flat_polyak(ptr(self._rept_flat), ptr(self._rep_flat), self._bridge.count, None, 1,
                                                 float(tau), float(np.float32(1. - tau)), 1, _lib.current_stream()),
                      'flat_polyak')
            if self._gru is not None:
                check(self._lib.asac_flat_polyak(ptr(self._rept_flat), ptr(self._rep_flat), self._gru.count, None, 1,
                                                 float(tau), float(np.float32(1. - tau)), 1, _lib.current_stream()),
                      'flat_polyak')

    def set_train_mode(self, train_mode=True):
        self.train_mode = train_mode
        for mod in self.ckpt_dict.values():
            if isinstance(mod, torch.nn.Module):
                mod.train(mode=train_mode)

    def get_global_step(self) -> int:
        return self._host_step

    def increase_global_step(self) -> int:
        self._host_step += 1
        self.global_step.fill_(self._host_step)
        return self._host_step

    def get_initial_action(self, batch_size, get_numpy=True):
        if get_numpy:
            return np.repeat(self._np_padding_action[np.newaxis, :], batch_size, axis=0)
        return self._padding_action_full.repeat(batch_size, 1)

    def get_initial_seq_hidden_state(self, batch_size, get_numpy=True):
        if get_numpy:
            return np.zeros([batch_size, *self.seq_hidden_state_shape], dtype=np.float32)
        return torch.zeros([batch_size, *self.seq_hidden_state_shape], device=self.device)

    def save_model(self, save_replay_buffer=False) -> None:
        if self.ckpt_dir is None:
            return
        self.flush_priority_update()
        step = self.get_global_step()
        path = self.ckpt_dir.joinpath(f'{step}.pth')
        torch.save({k: (v.detach().clone() if isinstance(v, torch.Tensor) else v.state_dict())
                    for k, v in self.ckpt_dict.items()}, path)
        self._logger.info(f'Model saved at {path}')
        if save_replay_buffer and self.use_replay_buffer:
            self.replay_buffer.save(self.ckpt_dir, step)

    def write_constant_summaries(self, constant_summaries: list[dict], iteration=None) -> None:
        if self.summary_writer is None:
            return
        for s in constant_summaries:
            self.summary_writer.add_scalar(s['tag'], s['simple_value'],
                                           self.get_global_step() if iteration is None else iteration)
        self.summary_writer.flush()

    # ------------------------------------------------------------------ actor side
    def _actor_io(self, rows: int) -> dict:
        """Per batch size: ONE pinned host staging block + its device mirror for everything choose_action
        sends (vector observations, pre_action, hidden state, offline action / injected draws) and ONE
        device block + pinned mirror for everything it returns (action, prob, next hidden state) — one
        H2D and one D2H copy per call instead of one per tensor."""
        io = self._actor_bufs.get(rows)
        if io is not None:
            return io
        A, S = self.c_action_size, self.state_size
        So = self._gru.obs_size if self._gru is not None else sum(shape[0] for _, shape in self._vector_obs)
        NLH = 0 if self._gru is None else self._gru.layers * self._gru.hidden
        f32 = dict(dtype=torch.float32, device=self.device)
        n_in = rows * (So + A + NLH + A)           # obs | pre_action | hidden | offline action or eps
        n_out = rows * (2 * A + NLH)
        io = {'h_in': torch.empty(n_in, dtype=torch.float32).pin_memory(), 'd_in': torch.empty(n_in, **f32),
              'h_out': torch.empty(n_out, dtype=torch.float32).pin_memory(), 'd_out': torch.empty(n_out, **f32),
              'scratch': torch.empty(rows, 2 * A, **f32), 'state': torch.empty(rows, 1, S, **f32)}
        cuts = np.cumsum([0, rows * So, rows * A, rows * NLH, rows * A])
        io['in_cuts'] = [int(c) for c in cuts]
        cuts = np.cumsum([0, rows * A, rows * A, rows * NLH])
        io['out_cuts'] = [int(c) for c in cuts]
        if len(self._actor_bufs) > 8:
            self._actor_bufs.clear()
        self._actor_bufs[rows] = io
        return io

    def _policy_act(self, state: torch.Tensor, action: torch.Tensor, prob: torch.Tensor, scratch: torch.Tensor,
                    offline: torch.Tensor | None, eps: torch.Tensor | None, disable_sample: bool) -> None:
        """asac_policy_act on ``state [rows, S]`` into ``action`` / ``prob [rows, A]`` (device, contiguous)."""
        rows, sh = int(state.shape[0]), self._pi_shape
        check(self._lib.asac_policy_act(ptr(self._pi_flat), self.state_size, sh.hidden, sh.depth, self.c_action_size,
                                        ptr(state), rows, ptr(eps), ptr(offline), int(bool(disable_sample)),
                                        self._noise_seed, ptr(self._act_counter), ptr(scratch), ptr(action), ptr(prob),
                                        1 if (self.actor_tensor_cores and rows >= 2048) else 0,
                                        _lib.current_stream()), 'policy_act')
        self._act_counter += 1

    @torch.no_grad()
    def choose_action(self, obs_list, pre_action, pre_seq_hidden_state, offline_action=None,
                      disable_sample: bool = False, force_rnd_if_available: bool = False, eps=None):
        """sac_base.py:968-1019 (continuous branch of _choose_action :882-966) on the device kernels:
        the concatenated vector observations (through one GRU step when the representation is recurrent,
        :1003) go through the policy's flat parameters (asac_policy_act: exact-fp32 row-tile forward; with
        ``self.actor_tensor_cores = True`` the tcgen05 3xTF32 forward from 2048 rows on) and one elementwise
        kernel samples, squashes and evaluates the per-dimension probability.  The tensor-core forward is
        opt-in because the probability amplifies the 1.5e-6 error of the pre-activations by |x - mu| / sigma^2.
        ``eps`` ([batch, A] N(0,1) draws) replaces the on-device Philox draws (tests).
        Host traffic: one pinned staging block in, one out (``_actor_io``)."""
        if self.action_noise is not None or self._bridge is not None or self._disc is not None:
            return self._choose_action_torch(obs_list, pre_action, pre_seq_hidden_state, offline_action,
                                             disable_sample)
        with torch.cuda.device(self.device):
            A, S = self.c_action_size, self.state_size
            rows = int(np.shape(obs_list[0])[0])
            io = self._actor_io(rows)
            c = io['in_cuts']
            h_in = io['h_in'].numpy()
            # ModelSimpleRep: concat of the vector observations (representation.py:74-83); the recurrent
            # representation reads obs_list[0] only (envs/test/nn_rnn.py)
            vec = [np.asarray(o, dtype=np.float32).reshape(rows, -1)
                   for shape, o in zip(self.obs_shapes, obs_list) if len(shape) == 1]
            vec = vec[:1] if self._gru is not None else vec
            h_in[c[0]:c[1]] = (vec[0] if len(vec) == 1 else np.concatenate(vec, axis=-1)).reshape(-1)
            if self._gru is not None:
                h_in[c[1]:c[2]] = np.asarray(pre_action, dtype=np.float32).reshape(-1)
                h_in[c[2]:c[3]] = np.asarray(pre_seq_hidden_state, dtype=np.float32).reshape(-1)
            extra = offline_action if offline_action is not None else eps
            if extra is not None:
                h_in[c[3]:c[4]] = (extra.detach().cpu().numpy() if isinstance(extra, torch.Tensor)
                                   else np.asarray(extra, dtype=np.float32)).reshape(-1)
            used = c[4] if extra is not None else (c[3] if self._gru is not None else c[1])
            d_in, d_out = io['d_in'], io['d_out']
            d_in[:used].copy_(io['h_in'][:used], non_blocking=True)
            oc = io['out_cuts']
            action, prob = d_out[oc[0]:oc[1]].view(rows, A), d_out[oc[1]:oc[2]].view(rows, A)
            if self._gru is not None:  # one GRU step from the caller's hidden state (sac_base.py:1003)
                g = self._gru
                hn = d_out[oc[2]:oc[3]]
                net = _lib.AsacGruNet(ptr(self._rep_flat), ptr(io['state']), ptr(hn), None)
                check(self._lib.asac_gru_forward(C.byref(self._gru_c), C.byref(net), 1, ptr(d_in[c[0]:c[1]]), None, 0,
                                                 ptr(d_in[c[1]:c[2]]), ptr(d_in[c[2]:c[3]]), g.layers * g.hidden,
                                                 rows, 1, _lib.current_stream()), 'gru_forward')
                state = io['state'].view(rows, S)
            else:
                state = d_in[c[0]:c[1]].view(rows, S)
            extra_dev = None if extra is None else d_in[c[3]:c[4]].view(rows, A)
            self._policy_act(state, action, prob, io['scratch'],
                             extra_dev if offline_action is not None else None,
                             extra_dev if (offline_action is None and eps is not None) else None, disable_sample)
            io['h_out'].copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            out = io['h_out'].numpy()
            hidden = out[oc[2]:oc[3]].reshape(rows, *self.seq_hidden_state_shape).copy() if self._gru is not None \
                else np.zeros((rows, *self.seq_hidden_state_shape), dtype=np.float32)
            return out[oc[0]:oc[1]].reshape(rows, A).copy(), out[oc[1]:oc[2]].reshape(rows, A).copy(), hidden

    @torch.no_grad()
    def choose_attn_action(self, ep_indexes, ep_padding_masks, ep_obses_list, ep_pre_actions, ep_pre_attn_states,
                           offline_action=None, disable_sample: bool = False, force_rnd_if_available: bool = False,
                           eps=None):
        """sac_base.py:1022-1086: the last ``burn_in_step`` steps of the running episode go through the
        attention representation with ONE query step (``is_prev_hidden_state=False``); the policy kernels
        act on the resulting state.  -> (action [batch, A], prob [batch, A], attn_state [batch, *shape]).
        The representation runs as the plugin's torch module on the parameters it shares with the learner."""
        if not self._is_attn:
            raise NotImplementedError('choose_attn_action needs seq_encoder=SEQ_ENCODER.ATTN (sac_base.py:1022)')
        with torch.cuda.device(self.device):
            b = self.burn_in_step
            tail = lambda x: torch.from_numpy(np.ascontiguousarray(x[:, -b:])).to(self.device)
            obs = self._process_torch_obs_list([tail(o) for o in ep_obses_list])
            state, attn_state, _ = self.model_rep(1, tail(ep_indexes), obs, tail(ep_pre_actions),
                                                  pre_seq_hidden_state=tail(ep_pre_attn_states),
                                                  is_prev_hidden_state=False, padding_mask=tail(ep_padding_masks))
            state = state.squeeze(1).contiguous()
            if self.action_noise is not None or self._disc is not None:  # discrete heads: the plugin's modules
                action, prob = self._act_from_state_torch(state, [o[:, -1] for o in obs], offline_action,
                                                          disable_sample, eps)
                return action.cpu().numpy(), prob.cpu().numpy(), attn_state.squeeze(1).cpu().numpy()
            rows, A = int(state.shape[0]), self.c_action_size
            f32 = dict(dtype=torch.float32, device=self.device)
            out, scratch = torch.empty(2, rows, A, **f32), torch.empty(rows, 2 * A, **f32)
            to_dev = lambda x: None if x is None else torch.as_tensor(x, dtype=torch.float32).to(self.device).contiguous()
            self._policy_act(state, out[0], out[1], scratch, to_dev(offline_action), to_dev(eps), disable_sample)
            host = out.cpu().numpy()
            return host[0], host[1], attn_state.squeeze(1).cpu().numpy()

    def _process_torch_obs_list(self, obs_list: list[torch.Tensor]) -> list[torch.Tensor]:
        """uint8 images -> [0, 1] floats, bool -> float (sac_base.py:1088-1099 convention)."""
        for i, o in enumerate(obs_list):
            if o.dtype == torch.uint8:
                obs_list[i] = o.float() / 255.
            elif o.dtype == torch.bool:
                obs_list[i] = o.float()
        return obs_list

    def log_episode(self, force: bool = False, **episode_trans) -> None:
        """sac_base.py:2247-2300: called by AgentManager after every episode (agent.py:688).  With an
        attention representation and a summary writer, the attention maps of the episode go to
        TensorBoard (needs matplotlib); in every case the pending-summary flag is cleared."""
        if not force and (self.summary_writer is None or not self.summary_available):
            return
        if self.summary_writer is not None and self._is_attn:
            try:
                from matplotlib.figure import Figure
                with torch.no_grad(), torch.cuda.device(self.device):
                    idx = torch.from_numpy(episode_trans['ep_indexes']).to(self.device)
                    obs = self._process_torch_obs_list([torch.from_numpy(o).to(self.device)
                                                        for o in episode_trans['ep_obses_list']])
                    from .utils.operators import gen_n_pre_actions
                    pre = torch.from_numpy(gen_n_pre_actions(episode_trans['ep_actions'])).to(self.device)
                    *_, maps = self.model_rep(idx.shape[1], idx, obs, pre, None)
                for i, w in enumerate(maps):
                    fig = Figure()
                    fig.subplots().imshow(w[0].cpu().numpy())
                    self.summary_writer.add_figure(f'attn_weight/{i}', fig, self.get_global_step())
            except ImportError:
                self._logger.warning('matplotlib is not installed: attention maps are not logged')
        self.summary_available = False

    @torch.no_grad()
    def _choose_action_torch(self, obs_list, pre_action, pre_seq_hidden_state, offline_action=None,
                             disable_sample: bool = False, eps=None):
        """The same computation with the plugin's torch modules on the shared parameter storage
        (kept for action_noise runs and as the in-process cross-check of the kernels)."""
        obs = [torch.from_numpy(np.asarray(o)).to(self.device) for o in obs_list]
        for i, o in enumerate(obs):
            if o.dtype == torch.uint8:
                obs[i] = o.float() / 255.
            elif o.dtype == torch.bool:
                obs[i] = o.float()
        if self._gru is not None or self._bridge is not None:
            to_dev = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.device)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):  # cuDNN RNNs default to TF32
                state, hidden = self.model_rep([o.unsqueeze(1) for o in obs], to_dev(pre_action).unsqueeze(1),
                                               to_dev(pre_seq_hidden_state).unsqueeze(1))
        else:
            state, hidden = self.model_rep([o.unsqueeze(1) for o in obs], None, None)
        state, hidden = state.squeeze(1), hidden.squeeze(1)
        action, prob = self._act_from_state_torch(state, obs, offline_action, disable_sample, eps)
        return action.cpu().numpy(), prob.cpu().numpy(), hidden.cpu().numpy()

    def _act_from_state_torch(self, state, obs, offline_action=None, disable_sample: bool = False, eps=None):
        """_choose_action (sac_base.py:882-966) from an encoded state with the plugin's policy / critic modules:
        -> (action, prob) [batch, D + A] on the device."""
        d_policy, c_policy = self.model_policy(state, obs)
        D = self.d_action_summed_size
        offline = None if offline_action is None else torch.as_tensor(offline_action, dtype=torch.float32).to(self.device)
        actions, probs = [], []
        if self.d_action_sizes:  # sac_base.py:901-936, 954-956 (policy branch; the DQN-like one is not supported)
            if offline is not None:
                d_action = offline[..., :D]
            elif self.discrete_dqn_like:  # arg-max of the first critic, epsilon-greedy while training (:903-925)
                c_in = c_policy.sample() if self.c_action_size else None
                d_qs, _ = self.model_q_list[0](state, c_in, obs)
                d_action = torch.cat([torch.nn.functional.one_hot(q.argmax(dim=-1), k).float()
                                      for q, k in zip(d_qs.split(self.d_action_sizes, dim=-1), self.d_action_sizes)], dim=-1)
                if self.train_mode:
                    batch = d_action.shape[0]
                    mask = (torch.rand(batch) < self.discrete_dqn_epsilon).to(self.device)
                    rnd = torch.cat([torch.nn.functional.one_hot(torch.randint(0, k, (batch,), device=self.device), k).float()
                                     for k in self.d_action_sizes], dim=-1)
                    d_action[mask] = rnd[mask]
            elif disable_sample:
                d_action = d_policy.sample_deter().float()
            else:
                d_action = d_policy.sample()
            actions.append(d_action)
            # prob stays 1 for the discrete columns of a DQN-like run (sac_base.py:951-956)
            probs.append(torch.ones_like(d_policy.probs) if self.discrete_dqn_like else d_policy.probs)
        if self.c_action_size:
            if offline is not None:
                c_action = offline[..., D:]
            elif disable_sample:
                c_action = torch.tanh(c_policy.mean)
            elif eps is not None:
                c_action = torch.tanh(c_policy.mean + c_policy.stddev * torch.as_tensor(eps, device=self.device))
            else:
                c_action = torch.tanh(c_policy.sample())
            if self.action_noise is not None:  # sac_base.py:859-880
                batch = c_action.shape[0]
                noise = torch.linspace(*self.action_noise, steps=batch, device=self.device)
                c_action = torch.tanh(torch.atanh(c_action) + torch.randn(batch, self.c_action_size, device=self.device)
                                      * noise.unsqueeze(1))
            x = torch.atanh(torch.clamp(c_action, -0.999, 0.999))
            floor = torch.clamp_min(1 - torch.tanh(x) ** 2, 1e-2)
            actions.append(c_action)
            probs.append(torch.exp(c_policy.log_prob(x)) / floor.prod(-1, keepdim=True))  # operators.py:17-19
        return torch.cat(actions, dim=-1), torch.cat(probs, dim=-1)

    # ------------------------------------------------------------------ ingest
    def put_episode(self, ep_indexes, ep_obses_list, ep_actions, ep_rewards, ep_dones, ep_probs,
                    ep_pre_seq_hidden_states) -> None:
        """sac_base.py:2303-2396."""
        if ep_indexes.shape[1] < self.n_step:
            return
        assert ep_indexes.dtype == np.int32
        last = np.zeros_like(ep_indexes, dtype=bool)
        last[:, -1] = True
        last[ep_indexes == -1] = True
        if not self.use_replay_buffer:  # sac_base.py:2341-2349
            self.batch_buffer.put_episode(ep_indexes=ep_indexes, ep_last_masks=last, ep_obses_list=ep_obses_list,
                                          ep_actions=ep_actions, ep_rewards=ep_rewards, ep_dones=ep_dones,
                                          ep_probs=ep_probs, ep_pre_seq_hidden_states=ep_pre_seq_hidden_states)
            return
        if self._world > 1 and self._episode_routing == 'round_robin':
            owner = adist.episode_owner(self._episode_counter, self._world)
            self._episode_counter += 1
            if owner != self._rank:
                return
        storage = {'index': ep_indexes.squeeze(0), 'last_mask': last.squeeze(0)}
        for name, o in zip(self.obs_names, ep_obses_list):
            storage[f'obs_{name}'] = o.squeeze(0)
        storage.update(action=ep_actions.squeeze(0), reward=ep_rewards.squeeze(0), done=ep_dones.squeeze(0),
                       mu_prob=ep_probs.squeeze(0), pre_seq_hidden_state=ep_pre_seq_hidden_states.squeeze(0))
        self.replay_buffer.add(storage, ignore_size=1)

    # ------------------------------------------------------------------ the step
    def _gather_specs(self, bt: dict):
        rb = self.replay_buffer
        S, A = self.state_size, self.c_action_size
        if self._disc is not None:  # stored rows are [one-hot per branch ..., continuous ...]: gathered whole, split after
            A = self._disc.AF
        specs = [('index', bt['index'], 4, 0, _lib.ROLE_INDEX),
                 ('last_mask', bt['last_masks'], 1, 0, _lib.ROLE_COPY),
                 ('action', bt['actions_full' if self._disc is not None else 'actions'], 4 * A, 0, _lib.ROLE_ACTION),
                 ('reward', bt['rewards'], 4, 0, _lib.ROLE_REWARD),
                 ('done', bt['dones'], 1, 0, _lib.ROLE_DONE),
                 ('mu_prob', bt['mu_full' if self._disc is not None else 'mu_probs'], 4 * A, 0, _lib.ROLE_MU_PROB)]
        off = 0
        if self._gru is not None:  # the representation reads obs_list[0] and the stored hidden state of the first row
            name, g = self.obs_names[0], self._gru
            if rb._columns[f'obs_{name}'].dtype != torch.float32:
                raise NotImplementedError(f'vector observation {name} is not stored as float32')
            specs.append((f'obs_{name}', bt['obs'], 4 * g.obs_size, 0, _lib.ROLE_COPY))
            specs.append(('pre_seq_hidden_state', bt['hidden'], 4 * g.layers * g.hidden, 0, _lib.ROLE_HIDDEN))
            if rb._row_bytes('pre_seq_hidden_state') != 4 * g.layers * g.hidden:
                raise ValueError('stored pre_seq_hidden_state rows do not have the shape (layers, hidden)')
        if self._bridge is not None:  # every observation window as stored + the stored hidden states of the window
            bt['obs_list'] = []
            for name, shape in zip(self.obs_names, self.obs_shapes):
                col = rb._columns[f'obs_{name}']
                if tuple(col.shape[1:]) != tuple(shape):
                    raise ValueError(f'stored obs_{name} rows have shape {tuple(col.shape[1:])}, expected {tuple(shape)}')
                win = torch.zeros(self.batch_size, self._cfg.seq_len, *shape, dtype=col.dtype, device=self.device)
                bt['obs_list'].append(win)
                specs.append((f'obs_{name}', win, rb._row_bytes(f'obs_{name}'), 0, _lib.ROLE_COPY))
            hb = rb._row_bytes('pre_seq_hidden_state')
            if hb != 4 * bt['hidden'].shape[-1]:
                raise ValueError('stored pre_seq_hidden_state rows do not have seq_hidden_state_shape')
            if hb:
                specs.append(('pre_seq_hidden_state', bt['hidden'], hb, 0, _lib.ROLE_HIDDEN))
        for name, shape in ([] if (self._gru is not None or self._bridge is not None) else self._vector_obs):
            col = rb._columns[f'obs_{name}']
            if col.dtype != torch.float32:
                raise NotImplementedError(f'vector observation {name} is stored as {col.dtype}; float32 expected')
            specs.append((f'obs_{name}', bt['states'], 4 * S, 4 * off, _lib.ROLE_COPY))
            off += shape[0]
        for key, _, nbytes, _, _ in specs[:6]:
            if rb._row_bytes(key) != nbytes:
                raise ValueError(f'stored column {key} has {rb._row_bytes(key)} bytes per row, expected {nbytes}')
        return specs

    def _enqueue_ensemble_perms(self) -> None:
        """The five torch.randperm(E) draws of a step (sac_base.py:1434, 1436, 1887 and the two of _get_td_error's
        _get_y), keyed by the global step; every rank of a data-parallel learner draws the same ones."""
        if self._cfg.ensemble_sample or self.discrete_dqn_like:
            check(self._lib.asac_ensemble_perms(ptr(self._ens_perms), 9, self.ensemble_q_num,
                                                (int(self._seed) if self._seed is not None else 0) ^ 0x5EED,
                                                ptr(self._counters), _lib.current_stream()), 'ensemble_perms')

    def _enqueue_noise(self, st: dict, stream_id: int) -> None:
        """The four Gaussian draws of a step (sac_base.py:1346, 1883, 1932, 2223) into batch set `st`:
        Philox keyed by (seed, global step at execution, stream_id)."""
        check(self._lib.asac_fill_normal(ptr(st['noise']), st['noise'].numel(), self._noise_seed, ptr(self._counters),
                                         stream_id, _lib.current_stream()), 'fill_normal')

    def _enqueue_sample(self, st: dict) -> None:
        """Prioritized sample + IS weights (replay_buffer.py:347-354) and the window gather fused with the
        padding rule (replay_buffer.py:356-362, sac_base.py:2435-2453) into batch set `st`, on the current stream."""
        lib, rb, smp = self._lib, self.replay_buffer, st['smp']
        check(lib.asac_per_sample(ptr(rb._nodes), rb.capacity, ptr(rb._store_ids), self.batch_size, None, rb._seed,
                                  ptr(rb._draw_counter), ptr(rb._per_state), ptr(smp['slots']), ptr(smp['ids']),
                                  ptr(smp['p']), ptr(smp['w']), _lib.current_stream()), 'per_sample')
        if self._world > 1 and self.use_priority:
            # IS weights against the smallest sampling probability over ALL shards' batches (replay_buffer.py:352-354
            # on the union of the draws): one 8-byte MIN all-reduce, off the critical path with sampling one step ahead
            self._min_prob.copy_(rb._per_state[2:3])
            torch.distributed.all_reduce(self._min_prob, op=torch.distributed.ReduceOp.MIN)
            check(lib.asac_per_shard_weights(ptr(rb._nodes), self.batch_size, ptr(smp['p']), ptr(rb._per_state),
                                             ptr(self._min_prob), ptr(smp['w']), _lib.current_stream()),
                  'per_shard_weights')
        if self._disc is not None:
            rb._gather(smp['ids'], st['specs'], self._padding_action_full, st['bt']['padding_masks'])
            if self._c_enabled:  # the continuous kernels read their own columns
                D = self._disc.D
                st['bt']['actions'].copy_(st['bt']['actions_full'][..., D:])
                st['bt']['mu_probs'].copy_(st['bt']['mu_full'][..., D:])
            return
        rb._gather(smp['ids'], st['specs'], self._padding_action, st['bt']['padding_masks'])

    def _enqueue_tree_update(self, st: dict) -> None:
        """PrioritizedReplayBuffer.update (replay_buffer.py:412-427) for the batch of set `st` from the td
        errors of the last step, on the current stream."""
        rb = self.replay_buffer
        check(self._lib.asac_per_update(ptr(rb._nodes), rb.capacity, ptr(rb._store_ids), ptr(st['smp']['ids']),
                                        ptr(self._wk['td_error']), self.batch_size, float(rb.td_error_min),
                                        float(rb.td_error_max), float(rb.alpha), 0, ptr(rb._per_state),
                                        _lib.current_stream()), 'per_update')

    def flush_priority_update(self) -> None:
        """Applies a deferred tree update now (before the replay is saved, closed or read from outside).
        A later step re-applying the same td errors is harmless: same ids, same values, same guard."""
        if self._pending:
            with torch.cuda.device(self.device):
                self._enqueue_tree_update(self._sets[self._cur])
            self._pending = False

    def _enqueue_step(self) -> None:
        """One train() on the device, eagerly: picks the batch sets, primes the first batch when needed."""
        if self._bridge is not None or self._disc is not None:
            return self._enqueue_step_staged()
        cur = (1 - self._cur) if (self._sample_ahead and self._primed) else self._cur
        if self._sample_ahead and not self._primed:
            self._enqueue_sample(self._sets[cur])
            if self._noise_ahead:
                self._enqueue_noise(self._sets[cur], 1)  # later steps get theirs one step ahead (stream 0)
            self._primed = True
        self._enqueue_step_sets(cur, apply_pending=self._pending)
        self._cur = cur

    def _enqueue_step_sets(self, cur: int, apply_pending: bool = True) -> None:
        """Everything one train() does on the device (graph-capturable) for batch set `cur`.  Critical
        path: value pass -> critics -> Adam -> policy -> Adam -> post pass -> fused tail.  Parallel
        branches of the graph: the Polyak update and the Gaussian draws; the mu-prob / hidden-state
        write-backs; and — like the reference's prefetch thread, which samples while `_train` runs
        (replay_buffer.py:339-375) — the sample + gather of the NEXT step into the other batch set.
        That branch reads the tree before this step's priority update and the rings before this
        step's write-backs (both ordered after it), so the schedule is deterministic: batch N + 1
        sees the priorities as of update N - 1."""
        lib, rb = self._lib, self.replay_buffer
        st = self._sets[cur]
        nxt = self._sets[1 - cur] if self._sample_ahead else None
        smp, cfg, prm, batch, work = st['smp'], self._cfg, self._prm, st['batch'], self._work
        bt, rep_c = st['bt'], st['rep']
        B = self.batch_size
        main = torch.cuda.current_stream(self.device)
        side = self._side_stream
        ahead = self._prefetch_stream
        peers = C.byref(self._peer_table) if self._peer_table is not None else None
        fused = self._world == 1 or peers is not None  # one code path for 1 GPU and for NVLink peers
        fast_tail = fused and self.use_priority and B <= 1024
        # side branch: _update_target_variables (sac_base.py:2057-2058) + the four Gaussian draws
        side.wait_stream(main)
        with torch.cuda.stream(side):
            s2 = side.cuda_stream
            if self._prefetch is not None:  # policy parameters and Adam moments -> L2 while sample / gather run
                check(lib.asac_l2_prefetch(self._prefetch[0], self._prefetch[1], len(self._prefetch[2]), s2),
                      'l2_prefetch')
            if fast_tail or rep_c is not None:
                check(lib.asac_sac_polyak(C.byref(cfg), C.byref(prm), -1.0, s2), 'sac_polyak')
            if rep_c is not None:
                check(lib.asac_flat_polyak(ptr(self._rept_flat), ptr(self._rep_flat), self._gru.count,
                                           ptr(self._counters), int(self.update_target_per_step), cfg.tau,
                                           cfg.one_minus_tau, 0, s2), 'flat_polyak')
            if nxt is None or not self._noise_ahead:  # this step's draws, next to the Polyak update
                self._enqueue_noise(st, 0)
            self._enqueue_ensemble_perms()
        stream = main.cuda_stream
        # 1 + 2. sample and gather: of the next step on its own branch, or (single batch set) of this one here
        defer = self._defer_active = self._defer_tree and fast_tail
        if nxt is not None:
            ahead.wait_stream(main)
            with torch.cuda.stream(ahead):
                if defer and apply_pending:  # the previous step trained on `nxt`'s buffers: its tree update, deferred
                    self._enqueue_tree_update(nxt)
                self._enqueue_sample(nxt)
                if self._noise_ahead:
                    self._enqueue_noise(nxt, 0)  # the next step's draws too (the global step is read before this
                #                              step's tail advances it: the branch is joined ahead of the tail)
        else:
            self._enqueue_sample(st)
        main.wait_stream(side)
        # 3. _train + get_l_probs + _get_td_error
        if rep_c is not None:  # trained GRU representation (sac_base.py:2066-2116)
            check(lib.asac_sac_step_networks_rep(C.byref(cfg), C.byref(prm), C.byref(batch), C.byref(work),
                                                 C.byref(rep_c), 0, peers, stream), 'sac_step_networks_rep')
            if not fast_tail:
                check(lib.asac_sac_staged_tail(C.byref(cfg), C.byref(prm), C.byref(work), stream), 'sac_staged_tail')
        elif not fused:
            self._enqueue_sac_step_data_parallel(stream, batch)
        elif fast_tail:
            check(lib.asac_sac_step_networks(C.byref(cfg), C.byref(prm), C.byref(batch), C.byref(work), 0, peers,
                                             stream), 'sac_step_networks')
        elif peers is not None:  # non-prioritized data-parallel run: NCCL path keeps the staged tail
            self._enqueue_sac_step_data_parallel(stream, batch)
        else:
            check(lib.asac_sac_step(C.byref(cfg), C.byref(prm), C.byref(batch), C.byref(work), stream), 'sac_step')
        # 4. mu-prob write-back (sac_base.py:2598-2605): needs the post pass only -> side branch
        if self.use_n_step_is or rep_c is not None:
            side.wait_stream(main)
            if nxt is not None:
                side.wait_stream(ahead)  # the next batch was gathered from the rings as they were before these writes
            with torch.cuda.stream(side):
                if rep_c is not None:  # next hidden states -> pre_seq_hidden_state of the following rows (:2589-2596)
                    rb.write_back(smp['ids'], 'pre_seq_hidden_state', self._rw['hn_post'], 1 - self.burn_in_step,
                                  bt['padding_masks'], n_rows=cfg.seq_len - 1)
                if self.use_n_step_is:
                    rb.write_back(smp['ids'], 'mu_prob', self._wk['pi_probs'], -self.burn_in_step,
                                  bt['padding_masks'])
        if nxt is not None:
            main.wait_stream(ahead)  # ... and sampled from the tree as it was before this step's priority update
        # 5. alpha step, td error, priority update (sac_base.py:2115-2116, 2571-2584), step counters
        if self.use_priority:
            if fast_tail:
                check(lib.asac_sac_finish_step(C.byref(cfg), C.byref(prm), C.byref(work),
                                               None if defer else ptr(rb._nodes), rb.capacity,
                                               ptr(rb._store_ids), ptr(smp['ids']), ptr(rb._per_state), peers,
                                               stream), 'sac_finish_step')
                self._pending = defer
            else:
                check(lib.asac_per_update(ptr(rb._nodes), rb.capacity, ptr(rb._store_ids), ptr(smp['ids']),
                                          ptr(self._wk['td_error']), B, float(rb.td_error_min),
                                          float(rb.td_error_max), float(rb.alpha), 0, ptr(rb._per_state), stream),
                      'per_update')
        if self.use_n_step_is or rep_c is not None:
            main.wait_stream(side)

    def _enqueue_step_staged(self) -> None:
        """One train() of a learner whose step is not one of the fused graphs: a representation that runs as the
        plugin's torch module (rep_bridge.py) and / or discrete action branches (discrete.py).  Sample + gather,
        the staged networks step, priority update, write-backs (sac_base.py:2556-2605) — one stream, program order."""
        rb, st = self.replay_buffer, self._sets[self._cur]
        bt, smp = st['bt'], st['smp']
        b, L = self.burn_in_step, self._cfg.seq_len
        self._enqueue_sample(st)
        self._enqueue_noise(st, 0)
        self._enqueue_ensemble_perms()
        hidden_post = self._staged_step_networks(st)
        if self.use_priority:
            self._enqueue_tree_update(st)
        if hidden_post is not None and int(np.prod(self.seq_hidden_state_shape)) != 0:
            hp = hidden_post.reshape(self.batch_size, L, -1).contiguous()
            rb.write_back(smp['ids'], 'pre_seq_hidden_state', hp, 1 - b, bt['padding_masks'], n_rows=L - 1)
        if self.use_n_step_is:  # [discrete probabilities, continuous densities] per row
            rows = self._wk['pi_probs_full'] if self._disc is not None else self._wk['pi_probs']
            rb.write_back(smp['ids'], 'mu_prob', rows, -b, bt['padding_masks'])

    _enqueue_step_bridge = _enqueue_step_discrete = _enqueue_step_staged

    def _staged_step_networks(self, st: dict):
        """_train + get_l_probs + _get_td_error on the batch held by set `st` (sac_base.py:2057-2116, 2556-2583) with
        the staged kernels, in the reference's order:

            _update_target_variables
            [representation: states (autograd graph kept), target_states]
            _train_rep_q:   y of the discrete and the continuous part, ONE loss per critic, one Adam step per ModelQ,
                            [the representation's share of loss.backward() + its Adam step, re-encoded states]
            _train_policy:  both parts, one Adam step        _train_alpha: both alphas, one Adam
            get_l_probs, _get_td_error

        -> next_bnx_seq_hidden_states (None without a torch-module representation)."""
        lib, br, dq, c = self._lib, self._bridge, self._disc, self._c_enabled
        bt, wk = st['bt'], self._wk
        cfg, prm, batch, work = C.byref(self._cfg), C.byref(self._prm), C.byref(st['batch']), C.byref(self._work)
        stream = _lib.current_stream()
        scale = 1.0 / self._world
        b, L = self.burn_in_step, self._cfg.seq_len
        bump = lambda mask: check(lib.asac_bump_counters(ptr(self._counters), mask, stream), 'bump_counters')
        dqn = self.discrete_dqn_like
        # ---- _update_target_variables
        if c:
            check(lib.asac_sac_polyak(cfg, prm, -1.0, stream), 'sac_polyak')
        if dq is not None:
            dq.polyak()
        states_t = post_t = target_t = bt['states']
        hidden_post = None
        if br is not None:
            check(lib.asac_flat_polyak(ptr(br.flat_target), ptr(br.flat), br.count, ptr(self._counters),
                                       int(self.update_target_per_step), self._cfg.tau, self._cfg.one_minus_tau, 0,
                                       stream), 'flat_polyak')
            # get_bnx_data (sac_base.py:1090-1116) from the b + n stored rows
            bn_index = bt['index'][:, :L - 1]
            index = torch.cat([bn_index, bn_index[:, -1:] + (bn_index[:, -1:] != -1)], dim=1)
            bn_pad = bt['padding_masks'][:, :L - 1].bool()
            padding = torch.cat([bn_pad, bn_pad[:, -1:]], dim=1)
            actions = (bt['actions_full'] if dq is not None else bt['actions'])[:, :L - 1]
            pre_action = torch.cat([torch.zeros_like(actions[:, :1]), actions], dim=1)
            obs_list = self._process_torch_obs_list(list(bt['obs_list']))
            hidden = bt['hidden'].view(self.batch_size, L, *self.seq_hidden_state_shape)
            states, _ = br.l_states(index, padding, obs_list, pre_action, hidden, target=False)
            with torch.no_grad():
                target_states, _ = br.l_states(index, padding, obs_list, pre_action, hidden, target=True)
                bt['states'].copy_(states)
                bt['target_states'].copy_(target_states)
            post_t, target_t = bt['states_post'], bt['target_states']
        # ---- _train_rep_q
        if dq is not None:
            dq.stage_target(st, states_t, states_t, post=False)
        if c:
            check(lib.asac_sac_target_y(cfg, prm, batch, work, stream), 'target_y')
            check(lib.asac_sac_q_backward(cfg, prm, batch, work, stream), 'q_backward')
            check(lib.asac_sac_reduce_grads(cfg, work, 0, stream), 'reduce_grads')
            adist.all_reduce_sum_(wk['grad_q'])
        if dq is not None:
            # `loss += loss + mse` (sac_base.py:1562): without the clipped loss the reference counts the discrete part twice
            dq.stage_q(st, states_t, 2.0 if (c and self.clip_epsilon <= 0) else 1.0, want_dx=br is not None)
            dq.adam_q()  # same optimizer as the continuous part: reads the step counter before it advances
        if c:
            check(lib.asac_sac_adam(cfg, prm, work, 0, scale, stream), 'adam')
        else:
            bump(2)
        if br is not None:
            # the representation's share of loss.backward() and optimizer_rep.step() (sac_base.py:1573-1601)
            g = wk['grad_state'].sum(dim=0) if c else torch.zeros_like(wk['grad_state'][0])
            if dq is not None:
                g = g + dq.wk['d_x'].sum(dim=(0, 1))
            br.backward(states, g)
            adist.all_reduce_sum_(br.grad)
            if self._world > 1:
                br.grad.mul_(scale)
            check(lib.asac_flat_reduce_adam(ptr(br.flat), ptr(br.m), ptr(br.v), ptr(br.grad), 1, br.stride, br.count,
                                            ptr(br.grad_out), ptr(self._counters[4:]), float(self.learning_rate), stream),
                  'flat_reduce_adam')
            bump(16)
            with torch.no_grad():  # get_l_states with the new weights (sac_base.py:2099-2105)
                states_post, hidden_post = br.l_states(index, padding, obs_list, pre_action, hidden, target=False)
                bt['states_post'].copy_(states_post)
            self._last_bridge = {'states': states.detach(), 'hidden_post': hidden_post}
        # ---- _train_policy
        if c:
            check(lib.asac_sac_policy_backward(cfg, prm, batch, work, stream), 'policy_backward')
            check(lib.asac_sac_reduce_grads(cfg, work, 1, stream), 'reduce_grads')
            adist.all_reduce_sum_(wk['grad_pi'])
        if dq is not None and not dqn:  # (DQN-like: no discrete policy / alpha loss, sac_base.py:1858, 1924)
            dq.stage_pi(st, post_t)
            dq.adam_pi()
        if c:
            check(lib.asac_sac_adam(cfg, prm, work, 1, scale, stream), 'adam')
        elif not dqn:
            bump(4)
        # ---- _train_alpha, get_l_probs, _get_td_error
        need_post = self.use_auto_alpha or self.use_n_step_is or self.use_priority
        if c and need_post:
            check(lib.asac_sac_post(cfg, prm, batch, work, stream), 'post')
        if dq is not None:
            dq.stage_alpha(st, post_t)
        if self.use_auto_alpha:
            if c:
                check(lib.asac_sac_reduce_grads(cfg, work, 2, stream), 'reduce_grads')
                adist.all_reduce_sum_(wk['grad_alpha'])
                check(lib.asac_sac_adam(cfg, prm, work, 2, scale, stream), 'adam')
            elif not dqn:
                bump(8)
        if c and need_post:
            check(lib.asac_sac_td_error(cfg, prm, work, stream), 'td_error')
            if dq is not None and self.use_n_step_is:
                wk['pi_probs_full'][..., dq.D:].copy_(wk['pi_probs'])
        if dq is not None:
            dq.stage_probs_td(st, post_t, target_t, post_t, accumulate_td=c)
        check(lib.asac_sac_advance_step(prm, stream), 'advance_step')
        return hidden_post

    _bridge_step_networks = _discrete_step_networks = _staged_step_networks

    def _enqueue_sac_step_data_parallel(self, stream, batch_struct) -> None:
        """asac_sac_step with a SUM all-reduce of each reduced gradient buffer between the backward
        pass and its Adam kernel (grad_scale = 1/world): the only collective of the step."""
        lib, cfg, prm, batch, work = self._lib, C.byref(self._cfg), C.byref(self._prm), C.byref(batch_struct), \
            C.byref(self._work)
        scale = 1.0 / self._world
        check(lib.asac_sac_polyak(cfg, prm, -1.0, stream), 'polyak')
        check(lib.asac_sac_target_y(cfg, prm, batch, work, stream), 'target_y')
        check(lib.asac_sac_q_backward(cfg, prm, batch, work, stream), 'q_backward')
        check(lib.asac_sac_reduce_grads(cfg, work, 0, stream), 'reduce_grads')
        adist.all_reduce_sum_(self._wk['grad_q'])
        check(lib.asac_sac_adam(cfg, prm, work, 0, scale, stream), 'adam')
        check(lib.asac_sac_policy_backward(cfg, prm, batch, work, stream), 'policy_backward')
        check(lib.asac_sac_reduce_grads(cfg, work, 1, stream), 'reduce_grads')
        adist.all_reduce_sum_(self._wk['grad_pi'])
        check(lib.asac_sac_adam(cfg, prm, work, 1, scale, stream), 'adam')
        if self.use_auto_alpha or self.use_n_step_is or self.use_priority:
            check(lib.asac_sac_post(cfg, prm, batch, work, stream), 'post')
        if self.use_auto_alpha:
            check(lib.asac_sac_reduce_grads(cfg, work, 2, stream), 'reduce_grads')
            adist.all_reduce_sum_(self._wk['grad_alpha'])
            check(lib.asac_sac_adam(cfg, prm, work, 2, scale, stream), 'adam')
        if self.use_auto_alpha or self.use_n_step_is or self.use_priority:
            check(lib.asac_sac_td_error(cfg, prm, work, stream), 'td_error')
        check(lib.asac_sac_advance_step(prm, stream), 'advance_step')

    def _train_on_policy(self, step: int) -> int:
        """train() without a replay buffer (sac_base.py:2509-2515, 2553): the oldest shuffled batch of windows
        from the BatchBuffer goes into the step's device buffers and the same kernels run — Polyak, _get_y,
        critic / [representation /] policy / alpha steps — with no IS weights, priorities or write-backs."""
        batch = self.batch_buffer.get_batch()
        if batch is None:
            return step
        (bn_indexes, bn_last_masks, bn_padding_masks, bnx_obses_list, bn_actions, bn_rewards, bn_dones, bn_mu_probs,
         bnx_hidden) = batch
        st = self._sets[0]
        bt, bn = st['bt'], self.burn_in_step + self.n_step
        with torch.cuda.device(self.device):
            bt['index'][:, :bn].copy_(bn_indexes)
            bt['last_masks'][:, :bn].copy_(bn_last_masks)
            bt['padding_masks'][:, :bn].copy_(bn_padding_masks)
            D = self.d_action_summed_size
            if self._disc is not None:
                bt['actions_full'][:, :bn].copy_(bn_actions)
                bt['mu_full'][:, :bn].copy_(bn_mu_probs)
            if self._c_enabled:
                bt['actions'][:, :bn].copy_(bn_actions[..., D:])
                bt['mu_probs'][:, :bn].copy_(bn_mu_probs[..., D:])
            bt['rewards'][:, :bn].copy_(bn_rewards)
            bt['dones'][:, :bn].copy_(bn_dones)
            if self._bridge is not None:
                bt['obs_list'] = [o.to(self.device) for o in bnx_obses_list]
                bt['hidden'].copy_(bnx_hidden.reshape(bt['hidden'].shape))
            elif self._gru is not None:
                bt['obs'].copy_(bnx_obses_list[0])
                bt['hidden'].copy_(bnx_hidden.reshape(bt['hidden'].shape))
            else:
                off = 0
                for (name, shape), o in zip(zip(self.obs_names, self.obs_shapes), bnx_obses_list):
                    if len(shape) == 1:  # ModelSimpleRep: state = concat of the vector observations
                        bt['states'][:, :, off:off + shape[0]].copy_(o)
                        off += shape[0]
            lib, cfg, prm, work = self._lib, C.byref(self._cfg), C.byref(self._prm), C.byref(self._work)
            stream = _lib.current_stream()
            self._enqueue_noise(st, 0)
            self._enqueue_ensemble_perms()
            if self._bridge is not None or self._disc is not None:
                self._staged_step_networks(st)
            elif st['rep'] is not None:
                check(lib.asac_flat_polyak(ptr(self._rept_flat), ptr(self._rep_flat), self._gru.count,
                                           ptr(self._counters), int(self.update_target_per_step), self._cfg.tau,
                                           self._cfg.one_minus_tau, 0, stream), 'flat_polyak')
                check(lib.asac_sac_polyak(cfg, prm, -1.0, stream), 'sac_polyak')
                check(lib.asac_sac_step_networks_rep(cfg, prm, C.byref(st['batch']), work, C.byref(st['rep']), 0, None,
                                                     stream), 'sac_step_networks_rep')
                check(lib.asac_sac_staged_tail(cfg, prm, work, stream), 'sac_staged_tail')
            else:
                check(lib.asac_sac_step(cfg, prm, C.byref(st['batch']), work, stream), 'sac_step')
        if self.save_model_per_step and step % self.save_model_per_step == 0:
            self.save_model()
        return self.increase_global_step()

    def train(self) -> int:
        step = self.get_global_step()
        if not self.use_replay_buffer:
            return self._train_on_policy(step)
        rb = self.replay_buffer
        if not self._ready_to_step():
            return step
        # the replay's host mutex: an actor thread's add() must not interleave its bookkeeping / staging
        # copies with the enqueue (or the capture) of a step
        with rb._mutex, torch.cuda.device(self.device):
            key = rb._columns_version
            if self._graph_columns_key != key:  # first step, or the storage was re-allocated (load / clear)
                for st in self._sets:
                    st['specs'] = self._gather_specs(st['bt'])
                self._graphs, self._graph_columns_key, self._primed = [None, None], key, False
                self._bridge_graph, self._bridge_eager_steps = (None if os.environ.get('ASAC_BRIDGE_GRAPH', '1') != '0'
                                                                else False), 0
                self._pending = False  # the storage was re-allocated: a deferred update has nothing to apply to
                self._enqueue_step()  # eager warm-up (also sets the kernels' shared-memory attributes)
            elif (self._bridge is not None or self._disc is not None) and self.use_cuda_graph and \
                    self._bridge_graph is not False and (self._world == 1 or self._graph_collectives):
                # (also the discrete / hybrid step: one stream, program order, captured whole)
                # the plugin's module inside the step's CUDA graph (forward x3, autograd backward): tried after a few
                # eager steps (cuDNN plans, lazy initialisation); a module that synchronises or branches on device
                # data cannot be captured — the step then stays eager for good
                self._bridge_eager_steps += 1
                if self._bridge_graph is None and self._bridge_eager_steps >= 3:
                    try:
                        torch.cuda.synchronize(self.device)
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                            self._enqueue_step()
                        self._bridge_graph = graph
                        self._graphs[0] = graph
                    except Exception as e:  # noqa: BLE001
                        self._logger.warning(f'representation module is not CUDA-graph capturable ({type(e).__name__}: '
                                             f'{str(e).splitlines()[0] if str(e) else ""}); the step runs eagerly')
                        self._bridge_graph = False
                        torch.cuda.synchronize(self.device)
                        if self._bridge is not None:
                            self._bridge.zero_grad()
                        self._enqueue_step()
                elif self._bridge_graph is None:
                    self._enqueue_step()
                if self._bridge_graph:
                    self._bridge_graph.replay()
            elif self._bridge is not None or self._disc is not None or not self.use_cuda_graph or \
                    (self._world > 1 and not self._graph_collectives):
                self._enqueue_step()
            else:  # one captured graph per batch set (they alternate)
                cur = (1 - self._cur) if self._sample_ahead else self._cur
                if self._graphs[cur] is None:
                    graph = torch.cuda.CUDAGraph()
                    # thread_local: an actor thread allocating / synchronising elsewhere must not invalidate the capture
                    with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                        self._enqueue_step_sets(cur)
                    self._graphs[cur] = graph
                self._graphs[cur].replay()
                self._cur = cur
                self._pending = self._defer_active  # (a flush in between cleared it; the replay deferred again)
        self._steps_since_check += 1
        if self._world > 1 and self._peer_table is not None and self._steps_since_check >= 256:
            self.check_peer_exchange()
        if self.save_model_per_step and step % self.save_model_per_step == 0:
            self.save_model()
        if self.summary_writer is not None and step % self.write_summary_per_step == 0:
            self._write_summaries(step)
        return self.increase_global_step()

    def _ready_to_step(self) -> bool:
        """sac_base.py:2503-2506 (train() is a no-op until the buffer holds more than one batch) — agreed on
        ACROSS ranks: shards fill unevenly, and a rank that stepped alone would wait inside the gradient
        exchange for peers that never enqueued the step.  One MIN all-reduce per call until every shard is
        ready; after that the check is free."""
        rb = self.replay_buffer
        if self._ready_version == rb._columns_version and rb.is_lg_batch_size:
            return True
        ready = rb.is_lg_batch_size
        if self._world > 1:
            ready = adist.all_agree(ready, self.device)
        if ready:
            self._ready_version = rb._columns_version
        return ready

    def check_peer_exchange(self) -> None:
        """Raises when a kernel gave up waiting for a peer's gradients (a crashed or diverged rank):
        the in-kernel exchange polls with a time limit and reports here instead of hanging the GPU."""
        self._steps_since_check = 0
        n = int(self._lib.asac_peer_timeouts(1))
        if n:
            raise _lib.AsacError(f'gradient exchange: {n} wait(s) for a peer rank timed out — a rank crashed or '
                                 'the ranks called train() a different number of times')

    def _write_summaries(self, step: int) -> None:
        wk, B = self._wk, self.batch_size
        w = self.summary_writer
        w.add_scalar('loss/q', float(wk['loss_q'].sum().item()) / (B * self.ensemble_q_num), step)  # ensemble mean
        w.add_scalar('loss/c_entropy', float(wk['stats_pi'][:, 1].sum().item()) / B, step)
        w.add_scalar('loss/c_alpha', float(torch.exp(self.log_c_alpha).item()), step)
        w.add_scalar('metric/replay_id', self.replay_buffer.get_curr_id(), step)
        w.flush()
        self.summary_available = True

    def last_step_stats(self) -> dict[str, float]:
        """Scalars of the most recent step (synchronises): what sac_base.py:2128-2178 logs."""
        wk, B = self._wk, self.batch_size
        return {'loss_q': float(wk['loss_q'].sum().item()) / (B * self.ensemble_q_num),  # mean over the ensemble
                'loss_policy': float(wk['stats_pi'][:, 0].sum().item()) / B,
                'c_entropy': float(wk['stats_pi'][:, 1].sum().item()) / B,
                'c_alpha': float(torch.exp(self.log_c_alpha).item()),
                'td_error_mean': float(wk['td_error'].mean().item())}

    def close(self):
        self._closed = True
        if getattr(self, '_peer_table', None) is not None:
            torch.cuda.synchronize(self.device)
            self.check_peer_exchange()
        if hasattr(self, 'replay_buffer') and getattr(self, '_pending', False):
            self.flush_priority_update()
        self._graphs = [None, None]
        if hasattr(self, 'replay_buffer'):
            self.replay_buffer.check_nan()
            self.replay_buffer.close()
        if self.summary_writer is not None:
            self.summary_writer.close()


class _AlphaAdam:
    """Adam state of ``[log_d_alpha, log_c_alpha]`` (sac_base.py:472): entry 0 exists when discrete branches train
    their alpha, entry 1 when continuous actions do (a parameter that never gets a gradient has no state)."""

    def __init__(self, m_buf, v_buf, counters, lr, d_state=None, c_enabled=True):
        self._m, self._v, self._counters, self._lr = m_buf, v_buf, counters, lr
        self._d, self._c = d_state, c_enabled

    def state_dict(self) -> dict:
        step = float(self._counters[3].item())
        state = {}
        if step > 0:
            if self._d is not None:
                state[0] = {'step': torch.tensor(step), 'exp_avg': self._d[0][0].clone(), 'exp_avg_sq': self._d[1][0].clone()}
            if self._c:
                state[1] = {'step': torch.tensor(step), 'exp_avg': self._m[0].clone(), 'exp_avg_sq': self._v[0].clone()}
        group = dict(lr=self._lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False,
                     foreach=None, capturable=False, differentiable=False, fused=None, params=[0, 1])
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd: dict) -> None:
        steps = []
        for key, (m, v) in ((0, self._d if self._d is not None else (None, None)), (1, (self._m, self._v))):
            if m is None:
                continue
            st = sd['state'].get(key)
            if st is None:
                m.zero_(); v.zero_()
                continue
            m[0] = st['exp_avg']; v[0] = st['exp_avg_sq']
            steps.append(int(float(st['step'])))
        self._counters[3] = max(steps) if steps else 0
