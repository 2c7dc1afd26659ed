"""Mint golden vectors by running the REAL reference (build container only).

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

The reference's own tests contain no assertions or golden vectors for this path
(SURVEY.md §4), so fixtures are produced differentially: identical seeded inputs
and injected Gaussian / uniform draws go through ``/root/reference`` (imported
read-only via ``oracle/ref_shims.py``) and every intermediate the CUDA path must
reproduce is stored.  The files travel to the GPU box; the reference does not.

Fixtures
--------
``per_<case>.npz``  ``SumTree`` / ``DataStorage`` / ``PrioritizedReplayBuffer``
                    traces (replay_buffer.py:21-477) + the learner-side padding
                    (sac_base.py:2435-2453).
``sac_<case>.npz``  ``SAC_Base._train`` + ``get_l_probs`` + ``_get_td_error``
                    traces (sac_base.py:2027-2245) over several consecutive steps.
"""
from __future__ import annotations

import importlib.util
import random
import sys
import threading
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.ref_shims import REFERENCE_ROOT, import_reference, load_reference_nn  # noqa: E402

GOLDEN = ROOT / 'tests' / 'golden'


def _load_synth():
    path = REFERENCE_ROOT / 'tests' / 'get_synthesis_data.py'
    spec = importlib.util.spec_from_file_location('ref_get_synthesis_data', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _NoThread:
    """Keeps ``PrioritizedReplayBuffer.__init__`` from starting its prefetch thread
    (replay_buffer.py:290-291) so the trace is deterministic."""

    def __enter__(self):
        self._start = threading.Thread.start
        threading.Thread.start = lambda self_: None
        return self

    def __exit__(self, *exc):
        threading.Thread.start = self._start


def _one_prefetch(rb, unit_uniform):
    """Run exactly one iteration of ``_prefetch_loop`` (replay_buffer.py:339-375) with the
    uniforms injected into ``np.random.uniform`` (:194)."""
    got = {}
    orig_uniform, orig_put = np.random.uniform, rb._queue.put

    def fake_uniform(low, high):
        return low + (high - low) * unit_uniform

    def fake_put(item):
        got['item'] = item
        rb._closed = True

    np.random.uniform, rb._queue.put = fake_uniform, fake_put
    try:
        rb._prefetch_loop()
    finally:
        np.random.uniform, rb._queue.put = orig_uniform, orig_put
        rb._closed = False
    return got['item']


# ----------------------------------------------------------------------------- PER
def gen_per_case(name, *, capacity, batch_size, prev_n, post_n, alpha, episode_lens, n_rounds, seed,
                 zero_fraction=0.0):
    _, ref_rb, _ = import_reference()
    rng = np.random.RandomState(seed)
    with _NoThread():
        rb = ref_rb.PrioritizedReplayBuffer(batch_size=batch_size, sample_prev_n=prev_n, sample_post_n=post_n,
                                            device=torch.device('cpu'), capacity=capacity, alpha=alpha)
    out = {'meta': np.array([capacity, batch_size, prev_n, post_n, n_rounds, len(episode_lens)], dtype=np.int64),
           'alpha': np.float64(alpha)}
    A = 3

    def make_episode(T):
        return {
            'index': np.arange(T, dtype=np.int32),
            'last_mask': np.concatenate([np.zeros(T - 1, dtype=bool), [True]]),
            'obs_vector': rng.randn(T, 5).astype(np.float32),
            'obs_image': rng.randint(0, 256, size=(T, 2, 3), dtype=np.uint8),
            'action': rng.rand(T, A).astype(np.float32),
            'reward': rng.randn(T).astype(np.float32),
            'done': rng.randint(0, 2, size=T).astype(bool),
            'mu_prob': rng.rand(T, A).astype(np.float32),
            'pre_seq_hidden_state': rng.randn(T, 2, 2).astype(np.float32),
        }

    for e, T in enumerate(episode_lens):
        ep = make_episode(T)
        for k, v in ep.items():
            out[f'add{e}.{k}'] = v
        rb.add(ep, ignore_size=1)
        out[f'add{e}.tree'] = rb._sum_tree._tree.copy()
        out[f'add{e}.ids'] = rb._trans_storage._buffer['_id'].copy()

    for r in range(n_rounds):
        if zero_fraction > 0 and r == 0:
            # knock out random leaves so the `right == 0` branch of the descent is exercised
            idx = rng.choice(rb.capacity, size=int(rb.capacity * zero_fraction), replace=False)
            rb._sum_tree.update(idx, np.zeros(len(idx), dtype=np.float32))
            out['zeroed.idx'] = idx.astype(np.int64)
            out['zeroed.tree'] = rb._sum_tree._tree.copy()
        u = rng.random_sample(batch_size)
        beta_before = float(rb.beta)
        data_ids, transitions, is_weights = _one_prefetch(rb, u)
        out[f'r{r}.u'] = u
        out[f'r{r}.beta_before'] = np.float64(beta_before)
        out[f'r{r}.data_ids'] = np.asarray(data_ids)
        out[f'r{r}.is_weights'] = is_weights.numpy().copy()
        for k, v in transitions.items():
            out[f'r{r}.batch.{k}'] = v.numpy().copy()
        # priorities: some ids duplicated, some stale (overwritten id), some NaN-free td errors
        td = np.abs(rng.randn(batch_size, 1)).astype(np.float32) * 0.7
        upd_ids = np.asarray(data_ids).copy()
        if batch_size >= 4:
            upd_ids[1] = upd_ids[0]                 # duplicate -> last one wins
            upd_ids[2] = upd_ids[2] + 7 * rb.capacity  # id that is no longer resident
        rb.update(upd_ids, td)
        out[f'r{r}.upd_ids'] = upd_ids
        out[f'r{r}.td'] = td
        out[f'r{r}.tree_after_update'] = rb._sum_tree._tree.copy()
        # mu_prob write-back on the same ids
        new_mu = rng.rand(batch_size, A).astype(np.float32)
        rb.update_transitions(upd_ids, 'mu_prob', new_mu)
        out[f'r{r}.new_mu'] = new_mu
        out[f'r{r}.mu_after'] = rb._trans_storage._buffer['mu_prob'].copy()
        # one more episode in between rounds so ids advance / wrap
        ep = make_episode(int(rng.randint(post_n + 2, 12)))
        for k, v in ep.items():
            out[f'r{r}.ep.{k}'] = v
        rb.add(ep, ignore_size=1)
        out[f'r{r}.tree_after_add'] = rb._sum_tree._tree.copy()
    out['final.ids'] = rb._trans_storage._buffer['_id'].copy()
    out['final.size'] = np.int64(rb._trans_storage.size)
    out['final.next_id'] = np.int64(rb._trans_storage._id)
    rb.close()
    np.savez_compressed(GOLDEN / f'per_{name}.npz', **out)
    print('wrote', f'per_{name}.npz', sum(v.nbytes for v in out.values()), 'bytes raw')


def gen_padding_case(name, *, burn_in_step, n_step, batch_size, capacity, seed):
    """``_sample_from_replay_buffer`` (sac_base.py:2398-2494) on a buffer with short
    episodes so windows cross episode boundaries."""
    SAC_Base, _, _ = import_reference()
    synth = _load_synth()
    nn = load_reference_nn('envs/test/nn.py')
    np.random.seed(seed); random.seed(seed)
    with _NoThread():
        sac = SAC_Base(obs_names=['vector'], obs_shapes=[(4,)], d_action_sizes=[], c_action_size=2,
                       model_abs_dir=None, nn=nn, device='cpu', seed=seed, batch_size=batch_size,
                       burn_in_step=burn_in_step, n_step=n_step,
                       replay_config={'capacity': capacity})
    out = {'meta': np.array([burn_in_step, n_step, batch_size, capacity], dtype=np.int64)}
    for e in range(12):
        ep = synth.gen_episode_trans([(4,)], [], 2, sac.seq_hidden_state_shape,
                                     episode_len=int(np.random.randint(n_step + 1, 9)))
        sac.put_episode(**ep)
    rb = sac.replay_buffer
    u = np.random.random_sample(batch_size)
    item = _one_prefetch(rb, u)
    data_ids, transitions, is_weights = item
    out['u'] = u
    out['data_ids'] = np.asarray(data_ids)
    for k, v in transitions.items():
        out[f'raw.{k}'] = v.numpy().copy()
    rb.sample = lambda: (data_ids, {k: v.clone() for k, v in transitions.items()}, is_weights.unsqueeze(-1))
    pointers, batch = sac._sample_from_replay_buffer()
    names = ['bn_indexes', 'bn_last_masks', 'bn_padding_masks', 'bnx_obs', 'bn_actions', 'bn_rewards',
             'bn_dones', 'bn_mu_probs', 'bnx_pre_seq_hidden_states', 'priority_is']
    for nme, t in zip(names, batch):
        if isinstance(t, list):
            t = t[0]
        out[f'padded.{nme}'] = t.numpy().copy()
    sac.close()
    np.savez_compressed(GOLDEN / f'pad_{name}.npz', **out)
    print('wrote', f'pad_{name}.npz')


# ----------------------------------------------------------------------------- SAC
class _NoiseTap:
    """Feeds recorded draws into ``Normal.rsample`` (via ``_standard_normal``) and
    ``Normal.sample`` (via ``torch.normal``)."""

    def __init__(self):
        import torch.distributions.normal as tdn
        self.tdn = tdn
        self.queue: list[torch.Tensor] = []

    def __enter__(self):
        self._std, self._normal = self.tdn._standard_normal, torch.normal

        def std(shape, dtype, device):
            z = self.queue.pop(0)
            assert tuple(z.shape) == tuple(shape), (z.shape, shape)
            return z.clone()

        def normal(mean, std_, *a, **k):
            z = self.queue.pop(0)
            assert z.shape == mean.shape
            return mean + std_ * z

        self.tdn._standard_normal, torch.normal = std, normal
        return self

    def __exit__(self, *exc):
        self.tdn._standard_normal, torch.normal = self._std, self._normal


def _write_custom_nn(tmpdir: Path, hidden: int, depth: int) -> Path:
    """A plugin file in the style of envs/gym/pendulum/nn.py with other widths."""
    p = tmpdir / f'nn_h{hidden}_d{depth}.py'
    p.write_text(
        'import algorithm.nn_models as m\n\n'
        'ModelRep = m.ModelSimpleRep\n\n\n'
        'class ModelQ(m.ModelQ):\n'
        '    def _build_model(self):\n'
        f'        super()._build_model(c_dense_n={hidden}, c_dense_depth={depth})\n\n\n'
        'class ModelPolicy(m.ModelPolicy):\n'
        '    def _build_model(self):\n'
        f'        super()._build_model(c_dense_n={hidden}, c_dense_depth={depth})\n')
    return p


def gen_sac_case(name, *, S, A, E, hidden, depth, B, b, n, steps, seed, use_priority=True, nn_rel=None, Es=None,
                 **hyper):
    """``Es`` (ensemble_q_sample < E): the reference takes the min over a random SUBSET of the critics
    (``torch.randperm(E)[:Es]``, sac_base.py:1434-1436, 1887) — the draws are recorded in call order under
    ``s*.in.perms``: _get_y current rows, next rows; _train_policy; _get_y of _get_td_error current, next."""
    SAC_Base, _, _ = import_reference()
    import tempfile
    if nn_rel is not None:
        nn = load_reference_nn(nn_rel)
    else:
        with tempfile.TemporaryDirectory() as td:
            path = _write_custom_nn(Path(td), hidden, depth)
            spec = importlib.util.spec_from_file_location(path.stem, path)
            nn = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(nn)
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    with _NoThread():
        sac = SAC_Base(obs_names=['vector'], obs_shapes=[(S,)], d_action_sizes=[], c_action_size=A,
                       model_abs_dir=None, nn=nn, device='cpu', seed=seed, batch_size=B,
                       burn_in_step=b, n_step=n, ensemble_q_num=E, ensemble_q_sample=E if Es is None else Es,
                       use_priority=use_priority, replay_config={'capacity': 1024}, **hyper)
    # make targets differ from the online nets (a fresh start hard-copies them, sac_base.py:629)
    with torch.no_grad():
        for tq in sac.model_target_q_list:
            for p in tq.parameters():
                p.add_(torch.randn_like(p) * 0.02)
        # non-zero biases so bias gradients / Adam are exercised from a generic point
        for net in sac.model_q_list + [sac.model_policy]:
            for pn, p in net.named_parameters():
                if pn.endswith('bias'):
                    p.add_(torch.randn_like(p) * 0.05)
    L = b + n + 1
    out = {'meta': np.array([S, A, E, hidden, depth, B, b, n, steps, int(use_priority)], dtype=np.int64)}
    hp = dict(tau=sac.tau, update_target_per_step=sac.update_target_per_step, learning_rate=sac.learning_rate,
              gamma=sac.gamma, v_lambda=sac.v_lambda, v_rho=float(sac.v_rho), v_c=float(sac.v_c),
              clip_epsilon=sac.clip_epsilon, use_n_step_is=float(sac.use_n_step_is),
              target_c_alpha=sac.target_c_alpha, init_log_alpha=float(sac.log_c_alpha),
              use_auto_alpha=float(sac.use_auto_alpha))
    for k, v in hp.items():
        out[f'hp.{k}'] = np.float64(v)

    def dump_params(prefix):
        for i in range(E):
            for k, t in sac.model_q_list[i].state_dict().items():
                out[f'{prefix}.q{i}.{k}'] = t.detach().numpy().copy()
            for k, t in sac.model_target_q_list[i].state_dict().items():
                out[f'{prefix}.qt{i}.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_policy.state_dict().items():
            out[f'{prefix}.pi.{k}'] = t.detach().numpy().copy()
        out[f'{prefix}.log_c_alpha'] = sac.log_c_alpha.detach().numpy().copy()

    dump_params('init')

    # wrap to capture intermediates
    ys = []
    orig_get_y = sac._get_y

    def tap_get_y(**kw):
        d_y, c_y = orig_get_y(**kw)
        ys.append(c_y.clone())
        return d_y, c_y

    sac._get_y = tap_get_y
    cap = {}
    orig_q, orig_pi = sac._train_rep_q, sac._train_policy

    def tap_q(**kw):
        r = orig_q(**kw)
        cap['loss_q0'] = r[0].detach().clone()
        return r

    def tap_pi(**kw):
        r = orig_pi(**kw)
        cap['c_entropy'] = r[1].detach().clone()
        return r

    sac._train_rep_q, sac._train_policy = tap_q, tap_pi

    for s in range(steps):
        states = torch.from_numpy(rng.randn(B, L, S).astype(np.float32))
        actions = torch.from_numpy((rng.rand(B, L - 1, A) * 1.9 - 0.95).astype(np.float32))
        if s == 0:
            actions[0, :, 0] = 1.0    # exercises clamp(a, +-0.999) before atanh
            actions[1, :, -1] = -1.0
        rewards = torch.from_numpy(rng.randn(B, L - 1).astype(np.float32))
        dones = torch.from_numpy(rng.rand(B, L - 1) < 0.15)
        mu_probs = torch.from_numpy((rng.rand(B, L - 1, A) * 1.5 + 0.01).astype(np.float32))
        # trailing windows that ran past an episode end -> padding; a few last_masks
        pad = np.zeros((B, L - 1), dtype=bool)
        last = np.zeros((B, L - 1), dtype=bool)
        for r in range(B):
            if L - 1 > 1 and rng.rand() < 0.4:
                cut = rng.randint(b + 1, L)  # first invalid position after the anchor
                if cut < L - 1:
                    pad[r, cut:] = True
                last[r, cut - 1] = rng.rand() < 0.7
            if b > 0 and rng.rand() < 0.3:
                pad[r, :rng.randint(1, b + 1)] = True
        mu_probs[torch.from_numpy(pad)] = 1.
        rewards[torch.from_numpy(pad)] = 0.
        dones[torch.from_numpy(pad)] = True
        actions[torch.from_numpy(pad)] = 0.
        index = torch.arange(L - 1, dtype=torch.int32).repeat(B, 1)
        index[torch.from_numpy(pad)] = -1
        pri = torch.from_numpy((rng.rand(B, 1) * 0.9 + 0.1).astype(np.float32)) if use_priority else None
        noise = dict(eps_y=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)),
                     eps_pi=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_alpha=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_td=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)))
        pre = f's{s}'
        for k, t in dict(states=states, actions=actions, rewards=rewards, dones=dones, mu_probs=mu_probs,
                         last_masks=torch.from_numpy(last), padding_masks=torch.from_numpy(pad)).items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()
        if pri is not None:
            out[f'{pre}.in.priority_is'] = pri.numpy().copy()
        for k, t in noise.items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()

        ys.clear()
        perms, real_randperm = [], torch.randperm

        def tap_randperm(*a, **k):
            pp = real_randperm(*a, **k)
            perms.append(pp.numpy().copy())
            return pp

        if Es is not None:
            torch.randperm = tap_randperm
        with _NoiseTap() as tap:
            tap.queue = [noise['eps_y'], noise['eps_pi']]
            if sac.use_auto_alpha:
                tap.queue.append(noise['eps_alpha'])
            if use_priority:
                tap.queue.append(noise['eps_td'])
            hidden_states = torch.zeros(B, L, 0)
            bnx_states, _, bnx_target_states = sac._train(
                bn_indexes=index.clone(), bn_last_masks=torch.from_numpy(last).clone(),
                bn_padding_masks=torch.from_numpy(pad).clone(), bnx_obses_list=[states.clone()],
                bn_actions=actions.clone(), bn_rewards=rewards.clone(), bn_dones=dones.clone(),
                bn_mu_probs=mu_probs.clone(), bnx_pre_seq_hidden_states=hidden_states,
                priority_is=pri.clone() if pri is not None else None)
            out[f'{pre}.out.y'] = ys[0].numpy().copy()
            out[f'{pre}.out.loss_q0'] = cap['loss_q0'].numpy().copy()
            out[f'{pre}.out.c_entropy'] = cap['c_entropy'].numpy().copy()
            for i in range(E):
                for k, p in sac.model_q_list[i].named_parameters():
                    out[f'{pre}.grad.q{i}.{k}'] = p.grad.detach().numpy().copy()
            for k, p in sac.model_policy.named_parameters():
                out[f'{pre}.grad.pi.{k}'] = p.grad.detach().numpy().copy()
            if sac.use_auto_alpha:
                out[f'{pre}.grad.log_c_alpha'] = sac.log_c_alpha.grad.detach().numpy().copy()
            assert sac.log_d_alpha.grad is None
            pi_probs = None
            bn_states = bnx_states[:, :-1]
            if sac.use_n_step_is:
                pi_probs = sac.get_l_probs(l_obses_list=[states[:, :-1]], l_states=bn_states, l_actions=actions)
                out[f'{pre}.out.pi_probs'] = pi_probs.numpy().copy()
            if use_priority:
                td = sac._get_td_error(
                    n_last_masks=torch.from_numpy(last)[:, b:], n_padding_masks=torch.from_numpy(pad)[:, b:],
                    nx_obses_list=[states[:, b:]], state=bn_states[:, b],
                    nx_target_states=bnx_target_states[:, b:], n_actions=actions[:, b:],
                    n_rewards=rewards[:, b:].clone(), n_dones=dones[:, b:],
                    n_mu_probs=pi_probs[:, b:].clone() if sac.use_n_step_is else None)
                out[f'{pre}.out.td_error'] = td.numpy().copy()
                out[f'{pre}.out.y_td'] = ys[1].numpy().copy()
            assert not tap.queue
        torch.randperm = real_randperm
        if Es is not None:
            assert len(perms) == (5 if use_priority else 3), len(perms)
            out[f'{pre}.in.perms'] = np.stack(perms).astype(np.int64)
        sac.increase_global_step()
        dump_params(f'{pre}.after')
    if Es is not None:
        out['ensemble_q_sample'] = np.int64(Es)
    sac.close()
    np.savez_compressed(GOLDEN / f'sac_{name}.npz', **out)
    print('wrote', f'sac_{name}.npz', sum(v.nbytes for v in out.values()), 'bytes raw')


def gen_sac_rnn_case(name, *, So, A, E, B, b, n, steps, seed, use_priority=True, **hyper):
    """``_train`` + tail of ``train`` with the recurrent plugin ``envs/test/nn_rnn.py`` (GRU(So + A -> 8,
    2 layers), hidden shape (2, 8)) and ``seq_encoder=SEQ_ENCODER.RNN``: the critic loss trains the
    representation through every burn-in step (sac_base.py:2066-2116)."""
    SAC_Base, _, _ = import_reference()
    from algorithm.utils.enums import SEQ_ENCODER
    nn = load_reference_nn('envs/test/nn_rnn.py')
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    with _NoThread():
        sac = SAC_Base(obs_names=['vector'], obs_shapes=[(So,)], d_action_sizes=[], c_action_size=A,
                       model_abs_dir=None, nn=nn, device='cpu', seed=seed, batch_size=B,
                       burn_in_step=b, n_step=n, ensemble_q_num=E, ensemble_q_sample=E,
                       seq_encoder=SEQ_ENCODER.RNN, use_priority=use_priority,
                       replay_config={'capacity': 1024}, **hyper)
    layers, H = tuple(sac.seq_hidden_state_shape)
    S = sac.state_size
    assert S == H
    with torch.no_grad():
        for net in sac.model_target_q_list + [sac.model_target_rep]:
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.02)
        for net in sac.model_q_list + [sac.model_policy]:
            for pn, p in net.named_parameters():
                if pn.endswith('bias'):
                    p.add_(torch.randn_like(p) * 0.05)
    L = b + n + 1
    q_depth = len([k for k in sac.model_q_list[0].state_dict() if k.endswith('linear.weight')])
    q_hidden = sac.model_q_list[0].state_dict()['c_dense.dense.0.linear.weight'].shape[0]
    out = {'meta': np.array([So, A, E, q_hidden, q_depth, B, b, n, steps, int(use_priority), layers, H],
                            dtype=np.int64)}
    hp = dict(tau=sac.tau, update_target_per_step=sac.update_target_per_step, learning_rate=sac.learning_rate,
              gamma=sac.gamma, v_lambda=sac.v_lambda, v_rho=float(sac.v_rho), v_c=float(sac.v_c),
              clip_epsilon=sac.clip_epsilon, use_n_step_is=float(sac.use_n_step_is),
              target_c_alpha=sac.target_c_alpha, init_log_alpha=float(sac.log_c_alpha),
              use_auto_alpha=float(sac.use_auto_alpha))
    for k, v in hp.items():
        out[f'hp.{k}'] = np.float64(v)

    def dump_params(prefix):
        for i in range(E):
            for k, t in sac.model_q_list[i].state_dict().items():
                out[f'{prefix}.q{i}.{k}'] = t.detach().numpy().copy()
            for k, t in sac.model_target_q_list[i].state_dict().items():
                out[f'{prefix}.qt{i}.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_policy.state_dict().items():
            out[f'{prefix}.pi.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_rep.state_dict().items():
            out[f'{prefix}.rep.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_target_rep.state_dict().items():
            out[f'{prefix}.rept.{k}'] = t.detach().numpy().copy()
        out[f'{prefix}.log_c_alpha'] = sac.log_c_alpha.detach().numpy().copy()

    dump_params('init')
    ys = []
    orig_get_y = sac._get_y

    def tap_get_y(**kw):
        d_y, c_y = orig_get_y(**kw)
        ys.append(c_y.clone())
        return d_y, c_y

    sac._get_y = tap_get_y
    rep_grads = {}
    orig_rep_step = sac.optimizer_rep.step

    def tap_rep_step(*a, **k):  # gradients as the representation's Adam sees them
        for kk, p in sac.model_rep.named_parameters():
            rep_grads[kk] = p.grad.detach().numpy().copy()
        return orig_rep_step(*a, **k)

    sac.optimizer_rep.step = tap_rep_step

    for s in range(steps):
        obs = torch.from_numpy(rng.randn(B, L, So).astype(np.float32))
        actions = torch.from_numpy((rng.rand(B, L - 1, A) * 1.9 - 0.95).astype(np.float32))
        rewards = torch.from_numpy(rng.randn(B, L - 1).astype(np.float32))
        dones = torch.from_numpy(rng.rand(B, L - 1) < 0.1)
        mu_probs = torch.from_numpy((rng.rand(B, L - 1, A) * 1.5 + 0.01).astype(np.float32))
        hidden = torch.from_numpy((rng.randn(B, L, layers, H) * 0.5).astype(np.float32))
        pad = np.zeros((B, L - 1), dtype=bool)
        last = np.zeros((B, L - 1), dtype=bool)
        for r in range(B):
            if L - 1 > 1 and rng.rand() < 0.4:
                cut = rng.randint(b + 1, L)
                if cut < L - 1:
                    pad[r, cut:] = True
                last[r, cut - 1] = rng.rand() < 0.7
            if b > 0 and rng.rand() < 0.3:
                pad[r, :rng.randint(1, b + 1)] = True
        tpad = torch.from_numpy(pad)
        mu_probs[tpad] = 1.
        rewards[tpad] = 0.
        dones[tpad] = True
        actions[tpad] = 0.
        hidden[:, :-1][tpad] = 0.  # sac_base.py:2453
        index = torch.arange(L - 1, dtype=torch.int32).repeat(B, 1)
        index[tpad] = -1
        pri = torch.from_numpy((rng.rand(B, 1) * 0.9 + 0.1).astype(np.float32)) if use_priority else None
        noise = dict(eps_y=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)),
                     eps_pi=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_alpha=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_td=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)))
        pre = f's{s}'
        for k, t in dict(obs=obs, hidden0=hidden[:, 0], actions=actions, rewards=rewards, dones=dones,
                         mu_probs=mu_probs, last_masks=torch.from_numpy(last), padding_masks=tpad).items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()
        if pri is not None:
            out[f'{pre}.in.priority_is'] = pri.numpy().copy()
        for k, t in noise.items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()

        ys.clear()
        with _NoiseTap() as tap:
            tap.queue = [noise['eps_y'], noise['eps_pi']]
            if sac.use_auto_alpha:
                tap.queue.append(noise['eps_alpha'])
            if use_priority:
                tap.queue.append(noise['eps_td'])
            bnx_states, next_hidden, bnx_target_states = sac._train(
                bn_indexes=index.clone(), bn_last_masks=torch.from_numpy(last).clone(),
                bn_padding_masks=tpad.clone(), bnx_obses_list=[obs.clone()],
                bn_actions=actions.clone(), bn_rewards=rewards.clone(), bn_dones=dones.clone(),
                bn_mu_probs=mu_probs.clone(), bnx_pre_seq_hidden_states=hidden.clone(),
                priority_is=pri.clone() if pri is not None else None)
            out[f'{pre}.out.y'] = ys[0].numpy().copy()
            out[f'{pre}.out.states_post'] = bnx_states.detach().numpy().copy()
            out[f'{pre}.out.target_states'] = bnx_target_states.detach().numpy().copy()
            out[f'{pre}.out.next_hidden'] = next_hidden[:, :-1].detach().numpy().copy()
            for k, g in rep_grads.items():
                out[f'{pre}.grad.rep.{k}'] = g
            for i in range(E):
                for k, p in sac.model_q_list[i].named_parameters():
                    out[f'{pre}.grad.q{i}.{k}'] = p.grad.detach().numpy().copy()
            for k, p in sac.model_policy.named_parameters():
                out[f'{pre}.grad.pi.{k}'] = p.grad.detach().numpy().copy()
            if sac.use_auto_alpha:
                out[f'{pre}.grad.log_c_alpha'] = sac.log_c_alpha.grad.detach().numpy().copy()
            pi_probs = None
            bn_states = bnx_states[:, :-1]
            if sac.use_n_step_is:
                pi_probs = sac.get_l_probs(l_obses_list=[obs[:, :-1]], l_states=bn_states, l_actions=actions)
                out[f'{pre}.out.pi_probs'] = pi_probs.numpy().copy()
            if use_priority:
                td = sac._get_td_error(
                    n_last_masks=torch.from_numpy(last)[:, b:], n_padding_masks=tpad[:, b:],
                    nx_obses_list=[obs[:, b:]], state=bn_states[:, b],
                    nx_target_states=bnx_target_states[:, b:], n_actions=actions[:, b:],
                    n_rewards=rewards[:, b:].clone(), n_dones=dones[:, b:],
                    n_mu_probs=pi_probs[:, b:].clone() if sac.use_n_step_is else None)
                out[f'{pre}.out.td_error'] = td.numpy().copy()
                out[f'{pre}.out.y_td'] = ys[1].numpy().copy()
            assert not tap.queue
        sac.increase_global_step()
        dump_params(f'{pre}.after')
    sac.close()
    np.savez_compressed(GOLDEN / f'sac_{name}.npz', **out)
    print('wrote', f'sac_{name}.npz', sum(v.nbytes for v in out.values()), 'bytes raw')


def gen_sac_plugin_rep_case(name, *, nn_rel, obs_names, obs_shapes, seq_encoder, A, E, B, b, n, steps, seed,
                            use_priority=True, **hyper):
    """``_train`` + tail of ``train`` with ANY representation plugin of the reference (``nn_rel``): convolutional
    encoders, the packed GRU, the episode attention stack (``seq_encoder`` 'RNN' / 'ATTN' / None).  Same
    recording as ``gen_sac_rnn_case``; the representation's parameters are stored under their state_dict keys."""
    SAC_Base, _, _ = import_reference()
    from algorithm.utils.enums import SEQ_ENCODER
    nn = load_reference_nn(nn_rel)
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    enc = None if seq_encoder is None else SEQ_ENCODER[seq_encoder]
    with _NoThread():
        sac = SAC_Base(obs_names=list(obs_names), obs_shapes=[tuple(s) for s in obs_shapes], d_action_sizes=[],
                       c_action_size=A, model_abs_dir=None, nn=nn, device='cpu', seed=seed, batch_size=B,
                       burn_in_step=b, n_step=n, ensemble_q_num=E, ensemble_q_sample=E, seq_encoder=enc,
                       use_priority=use_priority, replay_config={'capacity': 1024}, **hyper)
    hshape = tuple(int(x) for x in sac.seq_hidden_state_shape)
    S = sac.state_size
    with torch.no_grad():
        for net in sac.model_target_q_list + [sac.model_target_rep]:
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.02)
        for net in sac.model_q_list + [sac.model_policy]:
            for pn, p in net.named_parameters():
                if pn.endswith('bias'):
                    p.add_(torch.randn_like(p) * 0.05)
    L = b + n + 1
    q_depth = len([k for k in sac.model_q_list[0].state_dict() if k.endswith('linear.weight')])
    q_hidden = sac.model_q_list[0].state_dict()['c_dense.dense.0.linear.weight'].shape[0]
    out = {'meta': np.array([S, A, E, q_hidden, q_depth, B, b, n, steps, int(use_priority)], dtype=np.int64),
           'hidden_shape': np.array(hshape, dtype=np.int64),
           'seq_encoder': np.array(seq_encoder or ''), 'nn_rel': np.array(nn_rel),
           'obs_names': np.array(list(obs_names))}
    for i, shape in enumerate(obs_shapes):
        out[f'obs_shape{i}'] = np.array(shape, dtype=np.int64)
    hp = dict(tau=sac.tau, update_target_per_step=sac.update_target_per_step, learning_rate=sac.learning_rate,
              gamma=sac.gamma, v_lambda=sac.v_lambda, v_rho=float(sac.v_rho), v_c=float(sac.v_c),
              clip_epsilon=sac.clip_epsilon, use_n_step_is=float(sac.use_n_step_is),
              target_c_alpha=sac.target_c_alpha, init_log_alpha=float(sac.log_c_alpha),
              use_auto_alpha=float(sac.use_auto_alpha))
    for k, v in hp.items():
        out[f'hp.{k}'] = np.float64(v)

    def dump_params(prefix):
        for i in range(E):
            for k, t in sac.model_q_list[i].state_dict().items():
                out[f'{prefix}.q{i}.{k}'] = t.detach().numpy().copy()
            for k, t in sac.model_target_q_list[i].state_dict().items():
                out[f'{prefix}.qt{i}.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_policy.state_dict().items():
            out[f'{prefix}.pi.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_rep.state_dict().items():
            out[f'{prefix}.rep.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_target_rep.state_dict().items():
            out[f'{prefix}.rept.{k}'] = t.detach().numpy().copy()
        out[f'{prefix}.log_c_alpha'] = sac.log_c_alpha.detach().numpy().copy()

    dump_params('init')
    ys = []
    orig_get_y = sac._get_y

    def tap_get_y(**kw):
        d_y, c_y = orig_get_y(**kw)
        ys.append(c_y.clone())
        return d_y, c_y

    sac._get_y = tap_get_y
    rep_grads = {}
    orig_rep_step = sac.optimizer_rep.step

    def tap_rep_step(*a, **k):  # gradients as the representation's Adam sees them
        for kk, p in sac.model_rep.named_parameters():
            rep_grads[kk] = (torch.zeros_like(p) if p.grad is None else p.grad).detach().numpy().copy()
        return orig_rep_step(*a, **k)

    sac.optimizer_rep.step = tap_rep_step

    for s in range(steps):
        obses = [torch.from_numpy(rng.randn(B, L, *shape).astype(np.float32)) for shape in obs_shapes]
        actions = torch.from_numpy((rng.rand(B, L - 1, A) * 1.9 - 0.95).astype(np.float32))
        rewards = torch.from_numpy(rng.randn(B, L - 1).astype(np.float32))
        dones = torch.from_numpy(rng.rand(B, L - 1) < 0.1)
        mu_probs = torch.from_numpy((rng.rand(B, L - 1, A) * 1.5 + 0.01).astype(np.float32))
        hidden = torch.from_numpy((rng.randn(B, L, *hshape) * 0.5).astype(np.float32))
        pad = np.zeros((B, L - 1), dtype=bool)
        last = np.zeros((B, L - 1), dtype=bool)
        for r in range(B):
            if L - 1 > 1 and rng.rand() < 0.4:
                cut = rng.randint(b + 1, L)
                if cut < L - 1:
                    pad[r, cut:] = True
                last[r, cut - 1] = rng.rand() < 0.7
            if b > 0 and rng.rand() < 0.3:
                pad[r, :rng.randint(1, b + 1)] = True
        tpad = torch.from_numpy(pad)
        mu_probs[tpad] = 1.
        rewards[tpad] = 0.
        dones[tpad] = True
        actions[tpad] = 0.
        hidden[:, :-1][tpad] = 0.  # sac_base.py:2453
        first = torch.from_numpy(rng.randint(0, 50, size=(B, 1)).astype(np.int32))
        index = torch.arange(L - 1, dtype=torch.int32).repeat(B, 1) + first  # episode positions of the rows
        index[tpad] = -1
        pri = torch.from_numpy((rng.rand(B, 1) * 0.9 + 0.1).astype(np.float32)) if use_priority else None
        noise = dict(eps_y=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)),
                     eps_pi=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_alpha=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_td=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)))
        pre = f's{s}'
        for i, o in enumerate(obses):
            out[f'{pre}.in.obs{i}'] = o.numpy().copy()
        for k, t in dict(hidden=hidden, index=index, actions=actions, rewards=rewards, dones=dones, mu_probs=mu_probs,
                         last_masks=torch.from_numpy(last), padding_masks=tpad).items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()
        if pri is not None:
            out[f'{pre}.in.priority_is'] = pri.numpy().copy()
        for k, t in noise.items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()

        ys.clear()
        with _NoiseTap() as tap:
            tap.queue = [noise['eps_y'], noise['eps_pi']]
            if sac.use_auto_alpha:
                tap.queue.append(noise['eps_alpha'])
            if use_priority:
                tap.queue.append(noise['eps_td'])
            bnx_states, next_hidden, bnx_target_states = sac._train(
                bn_indexes=index.clone(), bn_last_masks=torch.from_numpy(last).clone(),
                bn_padding_masks=tpad.clone(), bnx_obses_list=[o.clone() for o in obses],
                bn_actions=actions.clone(), bn_rewards=rewards.clone(), bn_dones=dones.clone(),
                bn_mu_probs=mu_probs.clone(), bnx_pre_seq_hidden_states=hidden.clone(),
                priority_is=pri.clone() if pri is not None else None)
            out[f'{pre}.out.y'] = ys[0].numpy().copy()
            out[f'{pre}.out.states_post'] = bnx_states.detach().numpy().copy()
            out[f'{pre}.out.target_states'] = bnx_target_states.detach().numpy().copy()
            out[f'{pre}.out.next_hidden'] = next_hidden[:, :-1].detach().numpy().copy()
            for k, g in rep_grads.items():
                out[f'{pre}.grad.rep.{k}'] = g
            for i in range(E):
                for k, p in sac.model_q_list[i].named_parameters():
                    out[f'{pre}.grad.q{i}.{k}'] = p.grad.detach().numpy().copy()
            for k, p in sac.model_policy.named_parameters():
                out[f'{pre}.grad.pi.{k}'] = p.grad.detach().numpy().copy()
            if sac.use_auto_alpha:
                out[f'{pre}.grad.log_c_alpha'] = sac.log_c_alpha.grad.detach().numpy().copy()
            pi_probs = None
            bn_states = bnx_states[:, :-1]
            if sac.use_n_step_is:
                pi_probs = sac.get_l_probs(l_obses_list=[o[:, :-1] for o in obses], l_states=bn_states,
                                           l_actions=actions)
                out[f'{pre}.out.pi_probs'] = pi_probs.numpy().copy()
            if use_priority:
                td = sac._get_td_error(
                    n_last_masks=torch.from_numpy(last)[:, b:], n_padding_masks=tpad[:, b:],
                    nx_obses_list=[o[:, b:] for o in obses], state=bn_states[:, b],
                    nx_target_states=bnx_target_states[:, b:], n_actions=actions[:, b:],
                    n_rewards=rewards[:, b:].clone(), n_dones=dones[:, b:],
                    n_mu_probs=pi_probs[:, b:].clone() if sac.use_n_step_is else None)
                out[f'{pre}.out.td_error'] = td.numpy().copy()
                out[f'{pre}.out.y_td'] = ys[1].numpy().copy()
            assert not tap.queue
        sac.increase_global_step()
        dump_params(f'{pre}.after')
    sac.close()
    np.savez_compressed(GOLDEN / f'sac_{name}.npz', **out)
    print('wrote', f'sac_{name}.npz', sum(v.nbytes for v in out.values()), 'bytes raw')


def gen_ckpt_case(name, *, rnn: bool, seed: int):
    """A checkpoint directory written by the REAL reference (`save_model(save_replay_buffer=True)`,
    sac_base.py:654-668; replay_buffer.py:96-111, 220-227, 436-446) after a few real `train()` steps,
    plus what the restored learner must reproduce: the deterministic action of `choose_action`
    (sac_base.py:968-1019) on recorded observations.  Interchange fixture for SURVEY §8f rank 2."""
    import shutil
    SAC_Base, _, _ = import_reference()
    synth = _load_synth()
    kw, hidden_shape = {}, (0,)
    if rnn:
        from algorithm.utils.enums import SEQ_ENCODER
        nn = load_reference_nn('envs/test/nn_rnn.py')
        kw, hidden_shape = dict(seq_encoder=SEQ_ENCODER.RNN, burn_in_step=3, n_step=2), (2, 8)
    else:
        nn = load_reference_nn('envs/test/nn.py')
    out_dir = GOLDEN / f'ckpt_{name}'
    shutil.rmtree(out_dir, ignore_errors=True)
    out_dir.mkdir(parents=True)
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    sac = SAC_Base(obs_names=['vector'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2,
                   model_abs_dir=out_dir, nn=nn, device='cpu', seed=seed, batch_size=8, summary_path=None,
                   save_model_per_step=10 ** 9, replay_config={'capacity': 64}, **kw)
    for _ in range(3):
        ep = synth.gen_episode_trans(obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2,
                                     seq_hidden_state_shape=hidden_shape, episode_len=12)
        sac.put_episode(**ep)
    import time
    done = 0
    for _ in range(200):  # the prefetch thread fills the queue asynchronously
        if sac.train() > done:
            done += 1
        if done == 4:
            break
        time.sleep(0.01)
    assert done == 4, done
    sac.save_model(save_replay_buffer=True)
    rng = np.random.RandomState(seed)
    obs = rng.randn(5, 6).astype(np.float32)
    pre_action = rng.rand(5, 2).astype(np.float32)
    pre_hidden = (rng.randn(5, *hidden_shape) * 0.3).astype(np.float32)
    action, prob, hidden = sac.choose_action([obs], pre_action, pre_hidden, disable_sample=True)
    np.savez(out_dir / 'expect.npz', obs=obs, pre_action=pre_action, pre_hidden=pre_hidden, action=action,
             prob=prob, hidden=hidden, global_step=np.int64(sac.get_global_step()),
             rb_size=np.int64(sac.replay_buffer.size))
    sac.close()
    for p in out_dir.iterdir():  # keep model/ and expect.npz only
        if p.name not in ('model', 'expect.npz'):
            shutil.rmtree(p) if p.is_dir() else p.unlink()
    (out_dir / 'model' / '0.pth').unlink()  # the step-0 save of sac_base.py:2555-2556; the fixture is step 4
    print('wrote', out_dir, sorted(x.name for x in (out_dir / 'model').iterdir()))


def gen_sac_discrete_case(name, *, S, d_action_sizes, A, E, B, b, n, steps, seed, use_priority=True, **hyper):
    """``_train`` + ``get_l_probs`` + ``_get_td_error`` with discrete (A = 0) or hybrid (A > 0) action
    branches and the stock nets of envs/test/nn.py (sac_base.py:1356-1421, 1858-1880, 1924-1929):
    fixture for oracle/discrete_oracle.py."""
    SAC_Base, _, _ = import_reference()
    nn = load_reference_nn('envs/test/nn.py')
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    with _NoThread():
        sac = SAC_Base(obs_names=['vector'], obs_shapes=[(S,)], d_action_sizes=list(d_action_sizes), c_action_size=A,
                       model_abs_dir=None, nn=nn, device='cpu', seed=seed, batch_size=B,
                       burn_in_step=b, n_step=n, ensemble_q_num=E, ensemble_q_sample=E,
                       use_priority=use_priority, replay_config={'capacity': 1024}, **hyper)
    with torch.no_grad():
        for tq in sac.model_target_q_list:
            for p in tq.parameters():
                p.add_(torch.randn_like(p) * 0.02)
        for net in sac.model_q_list + [sac.model_policy]:
            for pn, p in net.named_parameters():
                if pn.endswith('bias'):
                    p.add_(torch.randn_like(p) * 0.05)
    D, L = sum(d_action_sizes), b + n + 1
    out = {'meta': np.array([S, A, E, 64, 3, B, b, n, steps, int(use_priority)], dtype=np.int64),
           'd_action_sizes': np.array(d_action_sizes, dtype=np.int64)}
    hp = dict(tau=sac.tau, update_target_per_step=sac.update_target_per_step, learning_rate=sac.learning_rate,
              gamma=sac.gamma, v_lambda=sac.v_lambda, v_rho=float(sac.v_rho), v_c=float(sac.v_c),
              clip_epsilon=sac.clip_epsilon, use_n_step_is=float(sac.use_n_step_is),
              target_c_alpha=sac.target_c_alpha, init_log_alpha=float(sac.log_c_alpha.detach()),
              use_auto_alpha=float(sac.use_auto_alpha), target_d_alpha=sac.target_d_alpha,
              d_policy_entropy_penalty=sac.d_policy_entropy_penalty,
              discrete_dqn_like=float(sac.discrete_dqn_like))
    for k, v in hp.items():
        out[f'hp.{k}'] = np.float64(v)

    def dump_params(prefix):
        for i in range(E):
            for k, t in sac.model_q_list[i].state_dict().items():
                out[f'{prefix}.q{i}.{k}'] = t.detach().numpy().copy()
            for k, t in sac.model_target_q_list[i].state_dict().items():
                out[f'{prefix}.qt{i}.{k}'] = t.detach().numpy().copy()
        for k, t in sac.model_policy.state_dict().items():
            out[f'{prefix}.pi.{k}'] = t.detach().numpy().copy()
        out[f'{prefix}.log_c_alpha'] = sac.log_c_alpha.detach().numpy().copy()
        out[f'{prefix}.log_d_alpha'] = sac.log_d_alpha.detach().numpy().copy()

    dump_params('init')
    ys = []
    orig_get_y = sac._get_y

    def tap_get_y(**kw):
        d_y, c_y = orig_get_y(**kw)
        ys.append((d_y.clone(), None if c_y is None else c_y.clone()))
        return d_y, c_y

    sac._get_y = tap_get_y
    for s in range(steps):
        states = torch.from_numpy(rng.randn(B, L, S).astype(np.float32))
        onehots, mus = [], []
        for size in d_action_sizes:
            idx = rng.randint(0, size, size=(B, L - 1))
            onehots.append(np.eye(size, dtype=np.float32)[idx])
            m = rng.rand(B, L - 1, size).astype(np.float32) + 0.05
            mus.append(m / m.sum(-1, keepdims=True))
        act_parts, mu_parts = onehots, mus
        if A:
            act_parts = onehots + [(rng.rand(B, L - 1, A) * 1.9 - 0.95).astype(np.float32)]
            mu_parts = mus + [(rng.rand(B, L - 1, A) * 1.5 + 0.01).astype(np.float32)]
        actions = torch.from_numpy(np.concatenate(act_parts, axis=-1))
        mu_probs = torch.from_numpy(np.concatenate(mu_parts, axis=-1))
        rewards = torch.from_numpy(rng.randn(B, L - 1).astype(np.float32))
        dones = torch.from_numpy(rng.rand(B, L - 1) < 0.15)
        pad = np.zeros((B, L - 1), dtype=bool)
        last = np.zeros((B, L - 1), dtype=bool)
        for r in range(B):
            if L - 1 > 1 and rng.rand() < 0.4:
                cut = rng.randint(b + 1, L)
                if cut < L - 1:
                    pad[r, cut:] = True
                last[r, cut - 1] = rng.rand() < 0.7
            if b > 0 and rng.rand() < 0.3:
                pad[r, :rng.randint(1, b + 1)] = True
        tpad = torch.from_numpy(pad)
        mu_probs[tpad] = 1.
        rewards[tpad] = 0.
        dones[tpad] = True
        actions[tpad] = 0.  # padding_action (sac_base.py:291-294)
        index = torch.arange(L - 1, dtype=torch.int32).repeat(B, 1)
        index[tpad] = -1
        pri = torch.from_numpy((rng.rand(B, 1) * 0.9 + 0.1).astype(np.float32)) if use_priority else None
        noise = dict(eps_y=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)),
                     eps_pi=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_alpha=torch.from_numpy(rng.randn(B, A).astype(np.float32)),
                     eps_td=torch.from_numpy(rng.randn(B, n + 1, A).astype(np.float32)))
        pre = f's{s}'
        for k, t in dict(states=states, actions=actions, rewards=rewards, dones=dones, mu_probs=mu_probs,
                         last_masks=torch.from_numpy(last), padding_masks=tpad).items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()
        if pri is not None:
            out[f'{pre}.in.priority_is'] = pri.numpy().copy()
        for k, t in noise.items():
            out[f'{pre}.in.{k}'] = t.numpy().copy()
        ys.clear()
        perms, real_randperm = [], torch.randperm

        def tap_randperm(*a, **k):  # the ensemble shuffles of sac_base.py:1366, 1377, ... in call order
            p = real_randperm(*a, **k)
            perms.append(p.numpy().copy())
            return p
        torch.randperm = tap_randperm
        with _NoiseTap() as tap:
            if A:
                tap.queue = [noise['eps_y'], noise['eps_pi']]
                if sac.use_auto_alpha:
                    tap.queue.append(noise['eps_alpha'])
                if use_priority:
                    tap.queue.append(noise['eps_td'])
            bnx_states, _, bnx_target_states = sac._train(
                bn_indexes=index.clone(), bn_last_masks=torch.from_numpy(last).clone(),
                bn_padding_masks=tpad.clone(), bnx_obses_list=[states.clone()],
                bn_actions=actions.clone(), bn_rewards=rewards.clone(), bn_dones=dones.clone(),
                bn_mu_probs=mu_probs.clone(), bnx_pre_seq_hidden_states=torch.zeros(B, L, 0),
                priority_is=pri.clone() if pri is not None else None)
            out[f'{pre}.out.d_y'] = ys[0][0].numpy().copy()
            if A:
                out[f'{pre}.out.y'] = ys[0][1].numpy().copy()
            for i in range(E):
                for k, p in sac.model_q_list[i].named_parameters():
                    out[f'{pre}.grad.q{i}.{k}'] = p.grad.detach().numpy().copy()
            for k, p in sac.model_policy.named_parameters():
                if p.grad is not None:
                    out[f'{pre}.grad.pi.{k}'] = p.grad.detach().numpy().copy()
            if sac.use_auto_alpha:
                if sac.log_d_alpha.grad is not None:
                    out[f'{pre}.grad.log_d_alpha'] = sac.log_d_alpha.grad.detach().numpy().copy()
                if A:
                    out[f'{pre}.grad.log_c_alpha'] = sac.log_c_alpha.grad.detach().numpy().copy()
            pi_probs = None
            bn_states = bnx_states[:, :-1]
            if sac.use_n_step_is:
                pi_probs = sac.get_l_probs(l_obses_list=[states[:, :-1]], l_states=bn_states, l_actions=actions)
                out[f'{pre}.out.pi_probs'] = pi_probs.numpy().copy()
            if use_priority:
                td = sac._get_td_error(
                    n_last_masks=torch.from_numpy(last)[:, b:], n_padding_masks=tpad[:, b:],
                    nx_obses_list=[states[:, b:]], state=bn_states[:, b],
                    nx_target_states=bnx_target_states[:, b:], n_actions=actions[:, b:],
                    n_rewards=rewards[:, b:].clone(), n_dones=dones[:, b:],
                    n_mu_probs=pi_probs[:, b:].clone() if sac.use_n_step_is else None)
                out[f'{pre}.out.td_error'] = td.numpy().copy()
            assert not tap.queue
        torch.randperm = real_randperm
        out[f'{pre}.in.perms'] = np.stack(perms) if perms else np.zeros((0, E), dtype=np.int64)
        sac.increase_global_step()
        dump_params(f'{pre}.after')
    sac.close()
    np.savez_compressed(GOLDEN / f'sac_{name}.npz', **out)
    print('wrote', f'sac_{name}.npz', sum(v.nbytes for v in out.values()), 'bytes raw')


CASES = {
    'per_small': lambda: gen_per_case('small', capacity=64, batch_size=8, prev_n=2, post_n=3, alpha=0.9,
                                      episode_lens=[9, 5, 17, 30, 12], n_rounds=3, seed=1),
    'per_zeros': lambda: gen_per_case('zeros', capacity=256, batch_size=32, prev_n=0, post_n=1, alpha=0.6,
                                      episode_lens=[100, 100, 90, 40], n_rounds=2, seed=2, zero_fraction=0.3),
    'pad_b2n3': lambda: gen_padding_case('b2n3', burn_in_step=2, n_step=3, batch_size=16, capacity=128, seed=3),
    'pad_b0n1': lambda: gen_padding_case('b0n1', burn_in_step=0, n_step=1, batch_size=16, capacity=128, seed=4),
    # config-2 shapes (envs/test/nn.py: H=64, depth 3), small batch, 3 consecutive steps
    'sac_c2': lambda: gen_sac_case('c2', S=6, A=2, E=2, hidden=64, depth=3, B=32, b=0, n=1, steps=3, seed=10,
                                   nn_rel='envs/test/nn.py'),
    # config-3 shapes (envs/gym/pendulum/nn.py: depth 2), n=5 V-trace with IS
    'sac_c3': lambda: gen_sac_case('c3', S=3, A=1, E=2, hidden=64, depth=2, B=24, b=0, n=5, steps=2, seed=11,
                                   nn_rel='envs/gym/pendulum/nn.py', v_lambda=1.0, use_n_step_is=True),
    # odd sizes: 3 critics, burn-in rows, lambda/rho/c != 1, no PER weights, clip_epsilon<=0 branch
    'sac_odd': lambda: gen_sac_case('odd', S=5, A=3, E=3, hidden=32, depth=1, B=12, b=2, n=3, steps=2, seed=12,
                                    use_priority=False, v_lambda=0.95, v_rho=0.9, v_c=0.8, clip_epsilon=0.0, tau=0.05,
                                    update_target_per_step=2, gamma=0.97),
    'sac_nois': lambda: gen_sac_case('nois', S=4, A=2, E=2, hidden=32, depth=2, B=10, b=0, n=2, steps=2, seed=13,
                                     use_n_step_is=False, use_auto_alpha=False),
    # min over a random 2-of-3 subset of the critics (ensemble_q_sample < ensemble_q_num), recorded randperm draws
    'sac_sub': lambda: gen_sac_case('sub', S=5, A=2, E=3, hidden=32, depth=2, B=12, b=1, n=3, steps=2, seed=19, Es=2),
    # config-4 shapes (envs/test/nn_rnn.py: GRU(6 + 2 -> 8, 2 layers)), shorter burn-in, 2 steps
    'sac_rnn': lambda: gen_sac_rnn_case('rnn', So=6, A=2, E=2, B=16, b=5, n=3, steps=2, seed=14, v_lambda=0.95),
    'sac_rnn_b0': lambda: gen_sac_rnn_case('rnn_b0', So=6, A=2, E=2, B=8, b=0, n=1, steps=2, seed=15,
                                           use_n_step_is=False),
    # the BASELINE configs at their FULL shapes, one step each: configs[1] B=256; configs[2] B=1024, n=5;
    # configs[3] B=256 sequences of burn-in 40 + n_step 5 (46-row windows, padding inside the burn-in)
    'sac_c2_b256': lambda: gen_sac_case('c2_b256', S=6, A=2, E=2, hidden=64, depth=3, B=256, b=0, n=1, steps=1,
                                        seed=30, nn_rel='envs/test/nn.py'),
    'sac_c3_b1024': lambda: gen_sac_case('c3_b1024', S=3, A=1, E=2, hidden=64, depth=2, B=1024, b=0, n=5, steps=1,
                                         seed=31, nn_rel='envs/gym/pendulum/nn.py', v_lambda=1.0, use_n_step_is=True),
    'sac_rnn_c4': lambda: gen_sac_rnn_case('rnn_c4', So=6, A=2, E=2, B=256, b=40, n=5, steps=1, seed=32),
    # representations that run as the plugin's torch module (asac_b200/rep_bridge.py): the reference's own test
    # plugins with a convolutional encoder in front of a packed GRU / the episode attention stack / nothing
    'sac_conv_rnn': lambda: gen_sac_plugin_rep_case('conv_rnn', nn_rel='tests/nn_conv_rnn.py',
                                                    obs_names=['vector', 'image'], obs_shapes=[(10,), (3, 30, 30)],
                                                    seq_encoder='RNN', A=2, E=2, B=12, b=5, n=3, steps=2, seed=40),
    'sac_conv_attn': lambda: gen_sac_plugin_rep_case('conv_attn', nn_rel='tests/nn_conv_attn.py',
                                                     obs_names=['vector', 'image'], obs_shapes=[(10,), (3, 30, 30)],
                                                     seq_encoder='ATTN', A=2, E=2, B=12, b=5, n=3, steps=2, seed=41),
    'sac_conv_vanilla': lambda: gen_sac_plugin_rep_case('conv_vanilla', nn_rel='tests/nn_conv_vanilla.py',
                                                        obs_names=['vector', 'image'], obs_shapes=[(10,), (3, 30, 30)],
                                                        seq_encoder=None, A=2, E=2, B=12, b=0, n=3, steps=2, seed=42),
    # discrete / hybrid action branches
    'sac_disc': lambda: gen_sac_discrete_case('disc', S=6, d_action_sizes=[3, 4], A=0, E=2, B=10, b=0, n=3, steps=2,
                                              seed=16, v_lambda=0.9),
    'sac_hybrid': lambda: gen_sac_discrete_case('hybrid', S=6, d_action_sizes=[3], A=2, E=2, B=10, b=1, n=2, steps=2,
                                                seed=17),
    'sac_dqn': lambda: gen_sac_discrete_case('dqn', S=6, d_action_sizes=[4, 2], A=0, E=2, B=10, b=0, n=3, steps=2,
                                             seed=18, discrete_dqn_like=True),
    # checkpoint directories written by the reference itself (interchange)
    'ckpt_vector': lambda: gen_ckpt_case('vector', rnn=False, seed=21),
    'ckpt_rnn': lambda: gen_ckpt_case('rnn', rnn=True, seed=22),
}


def main():
    """python oracle/gen_golden.py [case ...]   (no arguments: every case)"""
    GOLDEN.mkdir(parents=True, exist_ok=True)
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()


if __name__ == '__main__':
    main()
