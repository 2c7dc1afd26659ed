#!/usr/bin/env python
"""Benchmark of the SAC learner step + prioritized replay (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (configs[1] of BASELINE.json): env_type=TEST vector obs (6,), continuous A=2, stock
envs/test/nn.py nets, SAC + PER (alpha=0.9, capacity=524288, full), batch_size=256,
ensemble_q_num=2, n_step=1.  One step = one SAC_Base.train(): prioritized sample, window gather,
the whole update (critics, policy, alpha), td-error, priority update, mu-prob write-back.

Prints ONE JSON line (see the keys in main()).  `value` is train() steps/s with the replay
resident in HBM, each timed step bracketed by CUDA events with an L2 flush in between;
`e2e` adds, per step, a put_episode() of host transitions (pinned memory -> H2D) and a D2H
read of the step's td-errors through the public API.  `--impl reference` times the CPU port of
the reference's own path (oracle/, NumPy sumtree + torch-CPU update) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CFG = dict(name='c2', obs_shape=(6,), A=2, B=256, n_step=1, burn_in=0, E=2, hidden=64, depth=3, capacity=524288,
           per_alpha=0.9, episode_len=100, rep=None)
# BASELINE.json configs[2] (Pendulum shapes, n_step=5 + V-trace, batch 1024): `--config c3`, a secondary
# measurement — the default run is configs[1], the one the metric is quoted on
CONFIGS = {
    'c2': dict(CFG, name='c2'),
    'c3': dict(CFG, name='c3', obs_shape=(3,), A=1, B=1024, n_step=5, depth=2),
    # BASELINE.json configs[3] on ONE GPU (the 8-GPU replay sharding of that config is `--gpus 8` once the
    # representation's gradient joins the peer exchange): envs/test/nn_rnn.py GRU(6 + 2 -> 8, 2 layers),
    # burn-in 40, n_step 5 -> windows of 46 rows, batch 256 sequences, PER
    'c4': dict(CFG, name='c4', burn_in=40, n_step=5, rep=dict(hidden=8, layers=2)),
    # BASELINE.json configs[4]: the reference's tests/nn_conv_attn.py (ConvLayers 'simple' + EpisodeMultiheadAttention;
    # the plugin fixes the image at 3x30x30, observation shapes of tests/test_sac_params.py), seq_encoder=ATTN,
    # PER, batch 512.  The representation runs as the plugin's torch module (asac_b200/rep_bridge.py), everything
    # else on this repo's kernels; replay capacity 65536 (the float32 images of 524288 transitions would not fit
    # the stock storage layout: 5.6 GB per 524288 x 2700 floats is fine on 180 GB, but filling it takes minutes)
    'c5': dict(CFG, name='c5', B=512, burn_in=8, n_step=5, capacity=65536, bridge=dict(
        obs_names=['vector', 'image'], obs_shapes=[(10,), (3, 30, 30)], seq_encoder='ATTN')),
}
WORKLOADS = {
    'c2': 'TEST vector-obs(6,) A=2 SAC+PER alpha=0.9 capacity=524288 (full) batch=256 ensemble_q=2 n_step=1, '
          'stock envs/test/nn.py nets (H=64, depth 3)',
    'c3': 'Pendulum-v1 shapes obs(3,) A=1 SAC+PER n_step=5 V-trace (v_lambda=1, use_n_step_is) capacity=524288 (full) '
          'batch=1024 ensemble_q=2, envs/gym/pendulum/nn.py nets (H=64, depth 2), synthetic episodes',
    'c4': 'R2D2-style TEST vector-obs(6,) A=2 seq_encoder=RNN (envs/test/nn_rnn.py: GRU(8 -> 8, 2 layers) trained '
          'through the critic loss) burn_in_step=40 n_step=5 (46-row windows) SAC+PER capacity=524288 (full) '
          'batch=256 sequences ensemble_q=2, stock H=64 depth-3 Q / policy nets, single GPU',
}
WORKLOADS['c5'] = ('tests/nn_conv_attn.py (ConvLayers(30,30,3,simple) + EpisodeMultiheadAttention), obs vector(10,) + '
                   'image(3,30,30), A=2, seq_encoder=ATTN, burn_in_step=8 n_step=5 (14-row windows), SAC+PER '
                   'capacity=65536 (full) batch=512 ensemble_q=2; representation = plugin torch module (cuDNN/cuBLAS, '
                   'TF32 off), everything else on the hand-written kernels')
WORKLOAD = WORKLOADS['c2']
METRIC = 'sac_grad_steps_per_sec_batch256'
UNIT = 'steps/s'


# --------------------------------------------------------------------------- synthetic data
_HIDDEN_SHAPE = None  # set by build_learner for a bridged representation (the learner probes it)


def hidden_shape():
    if _HIDDEN_SHAPE is not None:
        return _HIDDEN_SHAPE
    r = CFG.get('rep')
    return (r['layers'], r['hidden']) if r else (0,)


def obs_shapes():
    br = CFG.get('bridge')
    return [tuple(s) for s in br['obs_shapes']] if br else [tuple(CFG['obs_shape'])]


def synth_episode(rng, T, S, A):
    """tests/get_synthesis_data.py:114-126 of the reference: obs randn, action rand, reward randn,
    done randint, probs rand (float32), hidden state randn (empty without a recurrent representation)."""
    return dict(ep_indexes=np.arange(T, dtype=np.int32)[None],
                ep_obses_list=[rng.randn(1, T, *shape).astype(np.float32) for shape in obs_shapes()],
                ep_actions=rng.rand(1, T, A).astype(np.float32),
                ep_rewards=rng.randn(1, T).astype(np.float32),
                ep_dones=rng.randint(0, 2, size=(1, T)).astype(bool),
                ep_probs=rng.rand(1, T, A).astype(np.float32),
                ep_pre_seq_hidden_states=rng.randn(1, T, *hidden_shape()).astype(np.float32))


def pin_episode(ep):
    out = {}
    for k, v in ep.items():
        if isinstance(v, list):
            out[k] = [torch.from_numpy(x).pin_memory().numpy() for x in v]
        else:
            out[k] = torch.from_numpy(v).pin_memory().numpy()
    return out


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        return self.snapshot()

    def snapshot(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == 'Active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# --------------------------------------------------------------------------- algorithmic work
def flops_per_row(S, A, H, d):
    q = 2 * ((S + A) * H + (d - 1) * H * H + H)
    pi = 2 * (S * H + (d - 1) * H * H + 2 * H * A)
    return q, pi


def stage_work(cfg):
    """ALGORITHMIC flops / bytes per launch of each kernel."""
    S, A, B, n, E, H, d = cfg['obs_shape'][0], cfg['A'], cfg['B'], cfg['n_step'], cfg['E'], cfg['hidden'], cfg['depth']
    So, r = S, cfg.get('rep')
    if r:
        S = r['hidden']  # the nets see the GRU's output
    b = cfg['burn_in']
    L = b + n + 1
    D = int(np.log2(cfg['capacity']))
    q, pi = flops_per_row(S, A, H, d)
    Pq = (S + A) * H + H + (d - 1) * (H * H + H) + H + 1
    Ppi = S * H + H + (d - 1) * (H * H + H) + 2 * A * H + 2 * A
    row = 4 + 1 + 4 * So + 4 * A + 4 + 1 + 4 * A + (4 * r['layers'] * r['hidden'] if r else 0)
    work = {
        'per_sample': dict(bound='hbm', bytes=B * (8 * D + 28)),
        'gather': dict(bound='hbm', bytes=2 * B * L * row),
        'fill_normal': dict(bound='hbm', bytes=4 * (2 * B * (n + 1) * A + 2 * B * A)),
        'polyak': dict(bound='hbm', bytes=3 * 4 * E * Pq),
        'value_pass_train': dict(bound='tensor', flops=pi * B * (n + 1) + E * q * (B * (n + 1) + B)),
        'q_backward': dict(bound='tensor', flops=3 * E * q * B),
        'adam_q': dict(bound='hbm', bytes=E * Pq * 4 * 7),
        'policy_backward': dict(bound='tensor', flops=(3 * pi + 2 * E * q) * B),
        'adam_pi': dict(bound='hbm', bytes=Ppi * 4 * 7),
        'value_pass_post': dict(bound='tensor', flops=pi * B * L + E * q * (B * (n + 1) + B)),
        'finish_step': dict(bound='hbm', bytes=B * 4 * (2 + E) + B * 8 + 64),
        'finish_step_fused_tree': dict(bound='hbm', bytes=B * (12 * D + 16) + B * 4 * (2 + E) + 64),
        'adam_alpha': dict(bound='hbm', bytes=64),
        'per_update': dict(bound='hbm', bytes=B * (12 * D + 16)),
        'write_back': dict(bound='hbm', bytes=B * (L - 1) * (A * 4 + 8)),
    }
    if r:
        Hg, NL = r['hidden'], r['layers']
        cell = sum(2 * 3 * Hg * ((So + A if l == 0 else Hg) + Hg) for l in range(NL))  # flops of one step, all layers
        Pg = sum(3 * Hg * ((So + A if l == 0 else Hg) + Hg + 2) for l in range(NL))
        work['value_pass_post'] = dict(bound='tensor', flops=pi * B * (L + n + 1) + E * q * (B * (n + 1) + B))
        work['gru_forward_pair'] = dict(bound='tensor', flops=2 * cell * B * L)
        work['gru_backward'] = dict(bound='tensor', flops=3 * cell * B * (b + 1))  # dX/dH chain + weight gradients
        work['adam_rep'] = dict(bound='hbm', bytes=Pg * 4 * 7)
        work['gru_forward_post'] = dict(bound='tensor', flops=cell * B * L)
        work['write_back_hidden'] = dict(bound='hbm', bytes=B * (L - 1) * (4 * NL * Hg + 8))
    return work


# --------------------------------------------------------------------------- GPU arm
PLUGIN_FILES = {'c2': 'envs_test_nn.py', 'c3': 'envs_gym_pendulum_nn.py', 'c4': 'envs_test_nn_rnn.py',
                'c5': 'tests_nn_conv_attn.py'}


def load_plugin(config: str):
    """The reference's own plugin file for the config, VERBATIM (tests/golden/plugins; byte-compared with the
    checkout by tests/test_plugin_surface.py), executed the way sac_main.py:353-364 does — through the
    `algorithm` alias package."""
    import importlib.util
    path = ROOT / 'tests' / 'golden' / 'plugins' / PLUGIN_FILES[config]
    spec = importlib.util.spec_from_file_location(f'bench_nn_{config}', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_learner(device, seed, capacity, fill, batch=None):
    """`capacity` is the GLOBAL replay capacity: under torch.distributed the learner shards it (SAC_Base owns the
    sharding); `fill` = transitions to store on this rank (None: the rank's whole shard)."""
    global _HIDDEN_SHAPE
    from asac_b200 import SAC_Base
    nn = load_plugin(CFG['name'])
    seq_encoder, names, shapes = None, ['vector'], [CFG['obs_shape']]
    from asac_b200.utils.enums import SEQ_ENCODER
    if CFG.get('rep'):
        seq_encoder = SEQ_ENCODER.RNN
    if CFG.get('bridge'):
        br = CFG['bridge']
        seq_encoder = SEQ_ENCODER[br['seq_encoder']] if br['seq_encoder'] else None
        names, shapes = list(br['obs_names']), obs_shapes()
    sac = SAC_Base(obs_names=names, obs_shapes=shapes, d_action_sizes=[], c_action_size=CFG['A'],
                   model_abs_dir=None, nn=nn, device=device, seed=seed, batch_size=batch or CFG['B'], n_step=CFG['n_step'],
                   burn_in_step=CFG['burn_in'], ensemble_q_num=CFG['E'], ensemble_q_sample=CFG['E'],
                   seq_encoder=seq_encoder, use_priority=True,
                   replay_config={'capacity': capacity, 'alpha': CFG['per_alpha']})
    if CFG.get('bridge'):
        _HIDDEN_SHAPE = tuple(sac.seq_hidden_state_shape)
    rng = np.random.RandomState(seed)
    S, A, T = CFG['obs_shape'][0], CFG['A'], CFG['episode_len']
    fill = sac.replay_buffer.capacity if fill is None else fill
    while sac.replay_buffer.size < fill:
        sac.put_episode(**synth_episode(rng, T, S, A))
    torch.cuda.synchronize()
    return sac, rng


def profile_stages(sac, steps):
    """Average device time of every stage of the step (plus the standalone alpha / priority-update
    entry points the fused tail replaces), each captured in its own CUDA graph."""
    import ctypes as C
    from asac_b200 import _lib
    from asac_b200._lib import check, ptr
    lib, rb = sac._lib, sac.replay_buffer
    cfg, prm, batch, work = C.byref(sac._cfg), C.byref(sac._prm), C.byref(sac._batch), C.byref(sac._work)
    smp, B = sac._smp, sac.batch_size
    s = lambda: _lib.current_stream()
    stages = [
        ('per_sample', lambda: check(lib.asac_per_sample(ptr(rb._nodes), rb.capacity, ptr(rb._store_ids), B, None,
                                                         rb._seed, ptr(rb._draw_counter), ptr(rb._per_state),
                                                         ptr(smp['slots']), ptr(smp['ids']), ptr(smp['p']),
                                                         ptr(smp['w']), s()))),
        ('gather', lambda: rb._gather(smp['ids'], sac._specs, sac._padding_action, sac._bt['padding_masks'])),
        ('fill_normal', lambda: check(lib.asac_fill_normal(ptr(sac._noise), sac._noise.numel(), sac._noise_seed,
                                                           ptr(sac._counters), 0, s()))),
        ('polyak', lambda: check(lib.asac_sac_polyak(cfg, prm, -1.0, s()))),
        ('value_pass_train', lambda: check(lib.asac_sac_target_y(cfg, prm, batch, work, s()))),
        ('q_backward', lambda: check(lib.asac_sac_q_backward(cfg, prm, batch, work, s()))),
        ('adam_q', lambda: check(lib.asac_sac_reduce_adam(cfg, prm, work, 0, s()))),
        ('policy_backward', lambda: check(lib.asac_sac_policy_backward(cfg, prm, batch, work, s()))),
        ('adam_pi', lambda: check(lib.asac_sac_reduce_adam(cfg, prm, work, 1, s()))),
        ('value_pass_post', lambda: check(lib.asac_sac_post(cfg, prm, batch, work, s()))),
        # the step's tail as the learner runs it (tree update deferred to the next step's ahead branch when
        # sampling one step ahead) and the fully fused variant (alpha + td + tree update in one CTA)
        ('finish_step', lambda: check(lib.asac_sac_finish_step(cfg, prm, work,
                                                               None if sac._defer_active else ptr(rb._nodes),
                                                               rb.capacity, ptr(rb._store_ids), ptr(smp['ids']),
                                                               ptr(rb._per_state), None, s()))),
        ('finish_step_fused_tree', lambda: check(lib.asac_sac_finish_step(cfg, prm, work, ptr(rb._nodes), rb.capacity,
                                                                          ptr(rb._store_ids), ptr(smp['ids']),
                                                                          ptr(rb._per_state), None, s()))),
        ('adam_alpha', lambda: check(lib.asac_sac_reduce_adam(cfg, prm, work, 2, s()))),
        ('per_update', lambda: check(lib.asac_per_update(ptr(rb._nodes), rb.capacity, ptr(rb._store_ids),
                                                         ptr(smp['ids']), ptr(sac._wk['td_error']), B,
                                                         float(rb.td_error_min), float(rb.td_error_max),
                                                         float(rb.alpha), 0, ptr(rb._per_state), s()))),
        ('write_back', lambda: rb.write_back(smp['ids'], 'mu_prob', sac._wk['pi_probs'], -sac.burn_in_step,
                                             sac._bt['padding_masks'])),
    ]
    if sac._rep is not None:  # trained GRU representation: its kernels, each alone
        rp, gshape = sac._rep, sac._gru_c
        L, bn = sac._cfg.seq_len, sac._cfg.bn_stride
        acts, P = ptr(sac._bt['actions']), sac._gru.count
        pair = (_lib.AsacGruNet * 2)()
        pair[0] = _lib.AsacGruNet(rp.params, rp.states, rp.hn, rp.save)
        pair[1] = _lib.AsacGruNet(rp.params_target, rp.target_states, None, None)
        again = _lib.AsacGruNet(rp.params, rp.states_post, rp.hn_post, None)
        stages += [
            ('gru_forward_pair', lambda: check(lib.asac_gru_forward(C.byref(gshape), pair, 2, rp.obs, acts, bn, None, rp.h0,
                                                                    rp.h0_b_stride, B, L, s()))),
            ('gru_backward', lambda: check(lib.asac_gru_backward(C.byref(gshape), rp.params, rp.obs, acts, bn, None, rp.h0,
                                                                 rp.h0_b_stride, B, L, sac.burn_in_step,
                                                                 ptr(sac._wk['grad_state']), sac.ensemble_q_num, rp.hn,
                                                                 rp.save, rp.grad_part, s()))),
            ('adam_rep', lambda: check(lib.asac_flat_reduce_adam(rp.params, rp.m, rp.v, rp.grad_part, rp.rep_tiles,
                                                                 sac._gru.stride, P, rp.grad, ptr(sac._counters[4:]),
                                                                 float(sac.learning_rate), s()))),
            ('gru_forward_post', lambda: check(lib.asac_gru_forward(C.byref(gshape), C.byref(again), 1, rp.obs, acts, bn,
                                                                    None, rp.h0, rp.h0_b_stride, B, L, s()))),
            ('write_back_hidden', lambda: rb.write_back(smp['ids'], 'pre_seq_hidden_state', sac._rw['hn_post'],
                                                        1 - sac.burn_in_step, sac._bt['padding_masks'], n_rows=L - 1)),
        ]
    # one tiny CUDA graph per stage, replayed back to back: device time without launch overhead
    # (warm L2; the whole-step numbers of run_gpu() are the ones taken with a flushed L2)
    out = {}
    side = torch.cuda.Stream()
    for name, fn in stages:
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) * 1e3 / steps  # us
    return out


def per_bulk_sample(rb, batch=1 << 22, iters=10):
    """SumTree.sample at a batch size where bandwidth, not the 19-level dependent chain, is the bound:
    asac_tree_sample over the learner's full tree.  Algorithmic bytes = batch * (8 D + 8) (two fp32
    children per level, slot + priority out); the 4 MiB tree is L2-resident, so the achieved figure
    is compared with the HBM peak only as a yardstick."""
    from asac_b200 import _lib
    from asac_b200._lib import check, ptr
    lib = _lib.load()
    D = int(np.log2(rb.capacity))
    slot = torch.empty(batch, dtype=torch.int32, device=rb.device)
    p = torch.empty(batch, dtype=torch.float32, device=rb.device)
    counter = torch.zeros(1, dtype=torch.int64, device=rb.device)
    s = _lib.current_stream()
    for _ in range(2):
        check(lib.asac_tree_sample(ptr(rb._nodes), rb.capacity, batch, None, 1234, ptr(counter), ptr(slot), ptr(p), s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        check(lib.asac_tree_sample(ptr(rb._nodes), rb.capacity, batch, None, 1234, ptr(counter), ptr(slot), ptr(p), s))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = batch * (8 * D + 8)
    return {'batch': batch, 'ms': round(ms, 4), 'samples_per_s': batch / (ms * 1e-3), 'algorithmic_bytes': nbytes,
            'achieved_GBps': nbytes / (ms * 1e-3) / 1e9}


def tensor_core_forward(rows=1 << 20, iters=10):
    """The stock critic forward on `rows` rows: exact-fp32 FFMA kernel vs the tcgen05 3xTF32 kernel."""
    from asac_b200 import _lib, lowering
    from asac_b200._lib import check, ptr
    lib = _lib.load()
    in_dim, H, depth, out = CFG['obs_shape'][0] + CFG['A'], CFG['hidden'], CFG['depth'], 1
    shape = lowering.NetShape(in_dim, H, depth, out)
    gen = torch.Generator(device='cuda').manual_seed(0)
    flat = (torch.rand(shape.stride, device='cuda', generator=gen) - 0.5) * 0.3
    x = torch.randn(rows, in_dim, device='cuda', generator=gen)
    y = torch.zeros(rows, out, device='cuda')
    s = _lib.current_stream()
    flops = 2 * (in_dim * H + (depth - 1) * H * H + H * out) * rows
    res = {'rows': rows, 'net': f'{in_dim}->{H}x{depth}->{out}'}
    for name, fn in (('ffma_fp32', lib.asac_mlp_forward), ('tcgen05_3xtf32', lib.asac_mlp_forward_tc)):
        for _ in range(3):
            check(fn(ptr(flat), in_dim, H, depth, out, ptr(x), rows, ptr(y), s), name)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            check(fn(ptr(flat), in_dim, H, depth, out, ptr(x), rows, ptr(y), s), name)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[name] = {'ms': round(ms, 4), 'rows_per_s': rows / (ms * 1e-3), 'fp32_equiv_TFLOPs': flops / (ms * 1e-3) / 1e12}
    res['tcgen05_3xtf32']['tensor_TFLOPs_issued'] = 3 * res['tcgen05_3xtf32']['fp32_equiv_TFLOPs'] * \
        (8 * H + (depth - 1) * H * H + H * 16) / (in_dim * H + (depth - 1) * H * H + H * out)
    return res


def _params_checksum(sac) -> torch.Tensor:
    """Bit pattern of every replicated parameter summed as int64 (an exact, order-free checksum)."""
    flats = [sac._q_flat, sac._qt_flat, sac._pi_flat, sac._log_alpha_buf]
    if getattr(sac, '_rep_flat', None) is not None:
        flats += [sac._rep_flat, sac._rept_flat]
    return torch.stack([t.reshape(-1).view(torch.int32).to(torch.int64).sum() for t in flats]).sum().reshape(1)


def measure(args, steps, warmup, clocks=None, detail=True):
    """Builds the learner of the current CFG, times `steps` train() calls (L2 flushed between them, CUDA events on
    the launching stream, max over ranks), the same steps with a warm L2, and the end-to-end loop through the public
    API.  -> the JSON line as a dict (rank 0; None on the other ranks)."""
    from asac_b200 import _lib
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    device = f'cuda:{local_rank}'
    lib = _lib.load()
    strong = args.scaling == 'strong' and world > 1
    if strong and CFG['B'] % world:
        raise SystemExit(f"--scaling strong: batch {CFG['B']} is not divisible by {world} ranks")
    batch = CFG['B'] // world if strong else CFG['B']

    t_fill = time.perf_counter()
    sac, rng = build_learner(device, seed=1 + rank, capacity=CFG['capacity'], fill=None, batch=batch)
    t_fill = time.perf_counter() - t_fill
    capacity = sac.replay_buffer.capacity
    S, A = CFG['obs_shape'][0], CFG['A']

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    # ---- warm-up (first call eager, second captures the CUDA graph)
    for _ in range(max(warmup, 3)):
        sac.train()
    barrier()

    # kernels of this library per train(): counted on one eager step (every rank takes it, so the
    # step's collectives stay matched across ranks)
    lib.asac_reset_launch_count()
    sac._enqueue_step()
    torch.cuda.synchronize()
    launches_per_step = int(lib.asac_launch_count())
    sac.increase_global_step()

    # ---- timed: K steps, one CUDA-event pair per step, L2 flushed between steps
    barrier()
    torch.cuda.profiler.start()  # `ncu --profile-from-start off` captures exactly the timed steps
    events = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sac.train()
        e1.record()
        events.append((e0, e1))
    barrier()
    torch.cuda.profiler.stop()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in events]
    total_ms = float(np.sum(step_ms))

    # ---- the same K steps back to back with a warm L2 (steady state of a running learner)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sac.train()
    e1.record()
    barrier()
    warm_ms = e0.elapsed_time(e1)

    # ---- end to end through the public API: host transitions in, td-errors out, every step
    T_in = 8
    eps = [pin_episode(synth_episode(rng, T_in, S, A)) for _ in range(32)]
    h2d = sum(v.nbytes for k, v in eps[0].items() if not isinstance(v, list)) + sum(x.nbytes for x in eps[0]['ep_obses_list'])
    h2d += T_in  # the derived last_mask column
    # Every step: put_episode() of pinned host transitions (H2D inside), train(), and a D2H copy of THAT step's
    # td-errors into a pinned host buffer which the host then reads.  The read of step i is awaited after step i + 1
    # has been enqueued (two host buffers, one CUDA event each) — the way an asynchronous learner consumes its
    # results — so the host work of step i + 1 overlaps the device work of step i; every step's result is read.
    td_hosts = [torch.empty(batch, dtype=torch.float32).pin_memory() for _ in range(2)]
    td_events = [torch.cuda.Event() for _ in range(2)]
    for i in range(5):
        sac.put_episode(**eps[i % 32]); sac.train(); td_hosts[0].copy_(sac._wk['td_error']); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc = 0.0
    for i in range(steps):
        sac.put_episode(**eps[i % 32])
        sac.train()
        td_hosts[i & 1].copy_(sac._wk['td_error'], non_blocking=True)
        td_events[i & 1].record()
        if i > 0:
            td_events[(i - 1) & 1].synchronize()
            acc += float(td_hosts[(i - 1) & 1][0])
    td_events[(steps - 1) & 1].synchronize()
    acc += float(td_hosts[(steps - 1) & 1][0])
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    td_host = td_hosts[0]
    clock_info = clocks.snapshot() if clocks is not None else None

    # ---- max over ranks; replicas must hold bit-identical parameters after all those steps
    replicas_identical = None
    if world > 1:
        t = torch.tensor([total_ms, warm_ms, e2e_ms], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms, warm_ms, e2e_ms = [float(x) for x in t.tolist()]
        ck = _params_checksum(sac)
        cks = [torch.zeros_like(ck) for _ in range(world)]
        torch.distributed.all_gather(cks, ck)
        replicas_identical = all(int(c.item()) == int(cks[0].item()) for c in cks)
        assert replicas_identical, f'data-parallel replicas diverged: parameter checksums {[int(c.item()) for c in cks]}'

    # weak scaling: every rank runs a batch-B update per step (global batch B x N), `value` counts them all;
    # strong scaling: the global batch stays B (B / N per rank), `value` = optimizer steps per second
    opt_steps_per_s = steps / (total_ms * 1e-3)
    units_per_step = 1 if strong else world
    value = units_per_step * opt_steps_per_s
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
        'ms_per_step': total_ms / steps, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': batch * world, 'replay_capacity_per_gpu': capacity,
                   'parallelism': f'dp{world}' if world > 1 else 'single',
                   'l2': 'flushed between timed steps (256 MiB memset, outside the event pairs)',
                   'cuda_graph': bool(sac._graph is not None),
                   'value_definition': (f"optimizer steps per second at a global batch of {batch * world}" if strong else
                                        f"batch-{CFG['B']} gradient steps per second summed over ranks "
                                        f"(= optimizer_steps_per_s x {world} ranks, each on its own batch)")},
        'optimizer_steps_per_s': opt_steps_per_s,
        'samples_per_s': opt_steps_per_s * batch * world,
        'value_warm_l2': units_per_step * steps / (warm_ms * 1e-3),
        'e2e': {'value': units_per_step * steps / (e2e_ms * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(td_host.numel() * 4),
                'what': f'put_episode({T_in} host transitions, pinned) + train() + D2H of td_error[{batch}] per step; the host '
                        'reads step i after enqueueing step i + 1 (double-buffered pinned results)'},
        'gpu_launches': launches_per_step * steps * world,
        'launches_per_step': launches_per_step,
        'clocks': clock_info,
        'fill_seconds': round(t_fill, 2),
    }
    if replicas_identical is not None:
        out['replicas_identical'] = replicas_identical

    if rank == 0 and world == 1 and sac._bridge is None:
        prof = profile_stages(sac, steps=200)
        work = stage_work(CFG)
        peaks = {}
        pk = ROOT / 'MEASURED_PEAKS.json'
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        tf_peak = float(peaks.get('bf16_tflops', 1590.0))
        # fp32 FFMA peak of the part: 148 SMs x 128 lanes x 2 flop x SM clock (the row-tile kernels' own ceiling)
        sm_mhz = (clock_info or {}).get('sm_max_mhz') or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)'
        tc_value_pass = {0: bool(lib.asac_sac_value_pass_on_tc(ctypes_byref(sac._cfg), 0)),
                         1: bool(lib.asac_sac_value_pass_on_tc(ctypes_byref(sac._cfg), 1))}
        kernels = {}
        for name, us in prof.items():
            w = work[name]
            if w['bound'] == 'hbm':
                ach = w['bytes'] / (us * 1e-6) / 1e9
                kernels[name] = {'us': round(us, 3), 'bound': 'hbm', 'algorithmic_bytes': w['bytes'],
                                 'achieved_GBps': round(ach, 3), 'frac': ach / hbm_peak}
            else:
                ach = w['flops'] / (us * 1e-6) / 1e12
                on_tc = (name == 'value_pass_train' and tc_value_pass[0]) or (name == 'value_pass_post' and tc_value_pass[1])
                kernels[name] = {'us': round(us, 3), 'bound': 'tensor', 'algorithmic_flops': w['flops'],
                                 'achieved_TFLOPs': round(ach, 5), 'frac': ach / tf_peak,
                                 'pipe': 'tcgen05 3xTF32 (k_value_pass_tc)' if on_tc else 'fp32 FFMA row tiles',
                                 'frac_of_fp32_ffma_peak': ach / fp32_peak}
        # standalone entry points that are not part of the step as the learner schedules it
        not_in_step = ('adam_alpha', 'finish_step_fused_tree') if sac._defer_active else \
            ('adam_alpha', 'per_update', 'finish_step_fused_tree')
        in_step = {k: v for k, v in prof.items() if k not in not_in_step}
        top = max(in_step, key=in_step.get)
        k = kernels[top]
        traffic = None
        tr = ROOT / 'profiles' / 'ncu_dram_traffic.json'  # dram__bytes_read+write per launch, `ncu --set full`
        if tr.exists():
            traffic = json.loads(tr.read_text()).get(f"{CFG['name']}:{top}", json.loads(tr.read_text()).get(top))
        out['roofline'] = {'kernel': top, 'bound': k['bound'],
                           'achieved': k.get('achieved_GBps', k.get('achieved_TFLOPs')),
                           'peak': hbm_peak if k['bound'] == 'hbm' else tf_peak,
                           'unit': 'GB/s' if k['bound'] == 'hbm' else 'TFLOP/s', 'frac': k['frac'],
                           'traffic': traffic, 'peak_source': src,
                           'share_of_step': in_step[top] / sum(in_step.values()),
                           'pipe': k.get('pipe'), 'fp32_ffma_peak_TFLOPs': round(fp32_peak, 2),
                           'frac_of_fp32_ffma_peak': k.get('frac_of_fp32_ffma_peak'),
                           'note': 'flop-bound stage of a latency-bound step: one 16-row tile per SM at B=256 '
                                   '(DESIGN.md §6); `peak` is the dense bf16 tensor peak the contract names, the '
                                   'fp32 FFMA ceiling of the kernel that actually runs is given beside it'}
        out['kernels'] = kernels
        if detail:
            out['tensor_core_forward'] = tensor_core_forward()
            out['per_bulk_sample'] = per_bulk_sample(sac.replay_buffer)
            out['per_bulk_sample']['frac_of_hbm_peak'] = out['per_bulk_sample']['achieved_GBps'] / hbm_peak
            per_us = prof['per_sample'] + prof['per_update']
            per_bytes = work['per_sample']['bytes'] + work['per_update']['bytes']
            out['per_sample_update'] = {'us': round(per_us, 3), 'algorithmic_bytes': per_bytes,
                                        'achieved_GBps': per_bytes / (per_us * 1e-6) / 1e9,
                                        'frac_of_hbm_peak': per_bytes / (per_us * 1e-6) / 1e9 / hbm_peak,
                                        'dependent_levels': int(np.log2(capacity))}
    sac.close()
    del sac, flush
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def ctypes_byref(x):
    import ctypes
    return ctypes.byref(x)


def run_gpu(args):
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=torch.device(f'cuda:{local_rank}'))
    clocks = ClockSampler(local_rank)
    clocks.start()  # from before the warm-up on: a 100 ms sampler never fires inside a few-ms timed region
    out = measure(args, args.steps, args.warmup, clocks=clocks, detail=True)
    # the other BASELINE configs, short runs of the same measurement, embedded so that one default run records them
    if world == 1 and args.config == 'c2' and not args.no_sub_results:
        global WORKLOAD, METRIC
        subs = {}
        for name, (k, w) in (('c3', (400, 10)), ('c4', (400, 10)), ('c5', (40, 5))):
            CFG.clear(); CFG.update(CONFIGS[name])
            WORKLOAD, METRIC = WORKLOADS[name], f"sac_grad_steps_per_sec_batch{CFG['B']}"
            try:
                r = measure(args, k, w, clocks=None, detail=False)
                subs[name] = {kk: r[kk] for kk in ('metric', 'value', 'unit', 'steps', 'ms_per_step', 'value_warm_l2', 'e2e',
                                                  'launches_per_step', 'roofline', 'kernels') if kk in r}
                subs[name]['workload'] = WORKLOADS[name]
            except Exception as e:  # noqa: BLE001 - a secondary measurement must not take the headline down
                subs[name] = {'error': f'{type(e).__name__}: {e}'}
        CFG.clear(); CFG.update(CONFIGS['c2'])
        WORKLOAD, METRIC = WORKLOADS['c2'], f"sac_grad_steps_per_sec_batch{CFG['B']}"
        out['other_configs'] = subs
    if rank == 0 and world == 1:
        out['cpu_baseline'] = cpu_port_baseline(budget_s=args.cpu_seconds)
    out_clocks = clocks.stop()
    if rank == 0:
        out['clocks'] = out_clocks
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# --------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_port_setup(seed=0):
    from oracle.replay_oracle import PerOracle
    from oracle.sac_oracle import SacHyper, SacOracle
    S, A, B = CFG['obs_shape'][0], CFG['A'], CFG['B']
    rng = np.random.RandomState(seed)
    per = PerOracle(batch_size=B, sample_prev_n=CFG['burn_in'], sample_post_n=CFG['n_step'],
                    capacity=CFG['capacity'], alpha=CFG['per_alpha'])
    per.vectorized = True
    C = per.capacity
    T = CFG['episode_len']
    # bulk fill (equivalent to C/T add() calls followed by priority updates): full ring, priorities from
    # |N(0,1)| td errors, episode tails at zero priority (replay_buffer.py:303-306)
    index = (np.arange(C) % T).astype(np.int32)
    per.store.columns = {
        '_id': np.arange(C, dtype=np.int64), 'index': index, 'last_mask': index == T - 1,
        'obs_vector': rng.randn(C, S).astype(np.float32), 'action': rng.rand(C, A).astype(np.float32),
        'reward': rng.randn(C).astype(np.float32), 'done': rng.randint(0, 2, size=C).astype(bool),
        'mu_prob': rng.rand(C, A).astype(np.float32),
        'pre_seq_hidden_state': rng.randn(C, *hidden_shape()).astype(np.float32)}
    per.store.size, per.store.next_id = C, C
    leaves = np.power(np.clip(np.abs(rng.randn(C)).astype(np.float32), 0.01, 1.0), np.float32(CFG['per_alpha']))
    leaves[index == T - 1] = 0
    per.tree.nodes[C - 1:] = leaves
    per.tree.rebuild()
    r = CFG.get('rep')
    hp = SacHyper(state_size=r['hidden'] if r else S, action_size=A, ensemble_q_num=CFG['E'], hidden=CFG['hidden'],
                  q_depth=CFG['depth'], policy_depth=CFG['depth'], burn_in_step=CFG['burn_in'], n_step=CFG['n_step'])
    if r:
        from oracle.rep_oracle import SacRepOracle
        return per, SacRepOracle(hp, S, r['layers'], seed=seed), rng
    return per, SacOracle(hp, seed=seed), rng


def cpu_port_step(per, sac, rng):
    """One SAC_Base.train() of the reference restated on the CPU (sac_base.py:2496-2609)."""
    from oracle.replay_oracle import pad_sampled_batch
    from oracle.sac_oracle import SacBatch, SacNoise
    B, A, n, b = CFG['B'], CFG['A'], CFG['n_step'], CFG['burn_in']
    data_ids, batch, weights, _ = per.sample(rng.random_sample(B))
    batch = pad_sampled_batch(batch, b, np.zeros(A, dtype=np.float32))
    t = torch.from_numpy
    common = dict(actions=t(batch['action'][:, :-1]), rewards=t(batch['reward'][:, :-1]),
                  dones=t(batch['done'][:, :-1]), mu_probs=t(batch['mu_prob'][:, :-1]),
                  last_masks=t(batch['last_mask'][:, :-1]), padding_masks=t(batch['padding_mask'][:, :-1]),
                  priority_is=t(weights))
    if CFG.get('rep'):
        from oracle.rep_oracle import SacRepBatch
        sb = SacRepBatch(obs=t(batch['obs_vector']), hidden0=t(batch['pre_seq_hidden_state'][:, 0]), **common)
    else:
        sb = SacBatch(states=t(batch['obs_vector']), **common)
    noise = SacNoise(eps_y=torch.randn(B, n + 1, A), eps_pi=torch.randn(B, A), eps_alpha=torch.randn(B, A),
                     eps_td=torch.randn(B, n + 1, A))
    out = sac.step(sb, noise)
    per.update(data_ids, out['td_error'].numpy())
    pad = batch['padding_mask'][:, :-1].reshape(-1)
    ptrs = (data_ids[:, None] + np.arange(-b, n)[None, :]).reshape(-1)
    per.update_transitions(ptrs[~pad], 'mu_prob', out['pi_probs'].numpy().reshape(-1, A)[~pad])
    if CFG.get('rep'):  # sac_base.py:2589-2596
        nh = out['next_hidden'].numpy()
        per.update_transitions(ptrs[~pad] + 1, 'pre_seq_hidden_state', nh.reshape(-1, *nh.shape[2:])[~pad])


def cpu_port_baseline(budget_s=12.0, steps=None, warmup=5):
    if CFG.get('bridge'):  # no port of a plugin-defined representation: only the reference itself can run it
        root = real_reference_root()
        if root is None:
            return {'unavailable': 'this configuration\'s representation is the plugin\'s own torch module; its CPU arm is '
                                   'the reference itself, which is not on this box'}
        return real_reference_baseline(root, steps or 20, warmup)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    per, sac, rng = cpu_port_setup()
    for _ in range(warmup):
        cpu_port_step(per, sac, rng)
    t0, done = time.perf_counter(), 0
    while True:
        cpu_port_step(per, sac, rng)
        done += 1
        if (steps is not None and done >= steps) or (steps is None and time.perf_counter() - t0 >= budget_s):
            break
    dt = time.perf_counter() - t0
    return {'value': done / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': f"{done} train() steps of the same workload (B={CFG['B']}, capacity {CFG['capacity']} full) in {dt:.1f} s: "
                      f'NumPy sumtree + torch-CPU fp32 update (oracle/), {threads} torch threads',
            'ms_per_step': dt / done * 1e3}


def real_reference_root():
    """A checkout / install of the UNMODIFIED reference, when one is reachable: $ASAC_REFERENCE_ROOT, baseline/_ref
    (driver-written on some boxes), /root/reference (the build container).  None on a bare GPU box."""
    for c in (os.environ.get('ASAC_REFERENCE_ROOT'), ROOT / 'baseline' / '_ref', '/root/reference'):
        if c and (Path(c) / 'algorithm' / 'sac_base.py').exists() and (Path(c) / 'algorithm' / 'replay_buffer.py').exists():
            return Path(c)
    return None


def real_reference_baseline(root: Path, steps: int, warmup: int):
    """The reference's own SAC_Base on the host cores: stock code path (its NumPy sumtree, prefetch thread, torch-CPU
    update) through its public API — put_episode to fill the replay, then train() — with the plugin file of the
    config.  Nothing of this repo is on that path except the import shims for what a GPU-less box lacks
    (oracle/ref_shims.py: matplotlib, torch.cuda.Stream, pin_memory)."""
    import importlib.util
    import oracle.ref_shims as rs
    rs.REFERENCE_ROOT = Path(root)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    SAC_Base, _, _ = rs.import_reference()
    from algorithm.utils.enums import SEQ_ENCODER  # the reference's own enum (its package is imported now)
    path = ROOT / 'tests' / 'golden' / 'plugins' / PLUGIN_FILES[CFG['name']]  # verbatim copy of the reference's file
    spec = importlib.util.spec_from_file_location(f'ref_bench_nn_{CFG["name"]}', path)
    nn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nn)
    seq_encoder, names, shapes = None, ['vector'], [tuple(CFG['obs_shape'])]
    if CFG.get('rep'):
        seq_encoder = SEQ_ENCODER.RNN
    if CFG.get('bridge'):
        br = CFG['bridge']
        seq_encoder = SEQ_ENCODER[br['seq_encoder']] if br['seq_encoder'] else None
        names, shapes = list(br['obs_names']), obs_shapes()
    sac = SAC_Base(obs_names=names, obs_shapes=shapes, d_action_sizes=[], c_action_size=CFG['A'], model_abs_dir=None,
                   nn=nn, device='cpu', seed=1, batch_size=CFG['B'], n_step=CFG['n_step'], burn_in_step=CFG['burn_in'],
                   ensemble_q_num=CFG['E'], ensemble_q_sample=CFG['E'], seq_encoder=seq_encoder, use_priority=True,
                   replay_config={'capacity': CFG['capacity'], 'alpha': CFG['per_alpha']})
    global _HIDDEN_SHAPE
    _HIDDEN_SHAPE = tuple(int(x) for x in sac.seq_hidden_state_shape)
    rng = np.random.RandomState(1)
    S, A, T = CFG['obs_shape'][0], CFG['A'], CFG['episode_len']
    t_fill = time.perf_counter()
    while not sac.replay_buffer.is_full:
        sac.put_episode(**synth_episode(rng, T, S, A))
    t_fill = time.perf_counter() - t_fill
    for _ in range(warmup):
        sac.train()
    t0 = time.perf_counter()
    for _ in range(steps):
        sac.train()
    dt = time.perf_counter() - t0
    sac.close()
    return {'value': steps / dt, 'unit': UNIT, 'cores': threads, 'kind': 'reference',
            'sample': f"{steps} SAC_Base.train() calls of the UNMODIFIED reference at {root} (B={CFG['B']}, capacity "
                      f"{CFG['capacity']} full, filled through put_episode in {t_fill:.0f} s) in {dt:.1f} s, {threads} torch threads",
            'ms_per_step': dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    base, note = None, None
    root = real_reference_root()
    if root is not None and os.environ.get('ASAC_REFERENCE_ARM', 'auto') != 'port':
        try:
            base = real_reference_baseline(root, args.steps, max(args.warmup, 3))
            note = f'the unmodified reference at {root}, stock code path, CPU'
        except Exception as e:  # noqa: BLE001 - e.g. a missing dependency of the reference on this box
            note = f'reference at {root} did not run ({type(e).__name__}: {e}); '
    if base is None:
        base = cpu_port_baseline(steps=args.steps, warmup=max(args.warmup, 3))
        note = (note or '') + ('CPU port of the reference path (oracle/): the Python reference itself is not on this box '
                               '(it cannot travel with the snapshot)')
    world = int(os.environ.get('WORLD_SIZE', 1))
    out = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': world,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': base['ms_per_step'], 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': WORKLOAD, 'note': note},
           'cpu_baseline': base,
           'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='N > 1: weak = every rank a full batch (default), strong = the global batch stays B')
    ap.add_argument('--no-sub-results', action='store_true', help='skip the short runs of the other BASELINE configs')
    args = ap.parse_args()
    global WORKLOAD, METRIC
    CFG.update(CONFIGS[args.config])
    WORKLOAD = WORKLOADS[args.config]
    METRIC = f"sac_grad_steps_per_sec_batch{CFG['B']}"
    if args.impl == 'reference':
        args.steps = 300 if args.steps is None else min(args.steps, 2000)
        args.warmup = 5 if args.warmup is None else args.warmup
        run_reference(args)
    else:
        args.steps = 2000 if args.steps is None else args.steps
        args.warmup = 20 if args.warmup is None else args.warmup
        run_gpu(args)


if __name__ == '__main__':
    main()
