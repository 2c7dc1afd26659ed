"""The comparisons that pin the CPU oracle to the reference, as functions of a fixture dict: used on the
committed fixtures (tests/test_oracle_golden.py) and on fixtures minted on the spot from the mounted
reference with fresh seeds (tests/test_oracle_vs_reference.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle.replay_oracle import PerOracle, pad_sampled_batch
from oracle.sac_oracle import SacOracle
from tests.helpers import golden_batch, golden_params, rel_err, sac_case_meta, sac_hyper_from_golden

EP_KEYS = ['index', 'last_mask', 'obs_vector', 'obs_image', 'action', 'reward', 'done', 'mu_prob',
           'pre_seq_hidden_state']


def check_per_trace(g: dict) -> None:
    capacity, batch_size, prev_n, post_n, n_rounds, n_eps = [int(x) for x in g['meta']]
    per = PerOracle(batch_size=batch_size, sample_prev_n=prev_n, sample_post_n=post_n, capacity=capacity,
                    alpha=float(g['alpha']))
    for e in range(n_eps):
        per.add({k: g[f'add{e}.{k}'] for k in EP_KEYS}, ignore_size=1)
        assert np.array_equal(per.tree.nodes, g[f'add{e}.tree'])
        assert np.array_equal(per.store.columns['_id'], g[f'add{e}.ids'])
    for r in range(n_rounds):
        if r == 0 and 'zeroed.idx' in g:
            per.tree.update(g['zeroed.idx'], np.zeros(len(g['zeroed.idx']), dtype=np.float32))
            assert np.array_equal(per.tree.nodes, g['zeroed.tree'])
        assert per.beta == pytest.approx(float(g[f'r{r}.beta_before']))
        data_ids, batch, weights, _ = per.sample(g[f'r{r}.u'])
        assert np.array_equal(data_ids, g[f'r{r}.data_ids'])
        assert np.array_equal(weights[:, 0], g[f'r{r}.is_weights'])
        for k in EP_KEYS:
            assert np.array_equal(batch[k], g[f'r{r}.batch.{k}']), k
        per.update(g[f'r{r}.upd_ids'], g[f'r{r}.td'])
        assert np.array_equal(per.tree.nodes, g[f'r{r}.tree_after_update'])
        per.update_transitions(g[f'r{r}.upd_ids'], 'mu_prob', g[f'r{r}.new_mu'])
        assert np.array_equal(per.store.columns['mu_prob'], g[f'r{r}.mu_after'])
        per.add({k: g[f'r{r}.ep.{k}'] for k in EP_KEYS}, ignore_size=1)
        assert np.array_equal(per.tree.nodes, g[f'r{r}.tree_after_add'])
    assert np.array_equal(per.store.columns['_id'], g['final.ids'])
    assert per.store.size == int(g['final.size'])
    assert per.store.next_id == int(g['final.next_id'])


def check_sac_steps(g: dict) -> None:
    torch.set_num_threads(1)
    m = sac_case_meta(g)
    hp = sac_hyper_from_golden(g)
    oracle = SacOracle(hp)
    oracle.load_params(*golden_params(g, 'init', m['E']))
    tol = 2e-6  # two fp32 CPU evaluations of the same graph; the CUDA gate is 1e-5 (relative to scale)
    for s in range(m['steps']):
        batch, noise = golden_batch(g, s)
        pre = f's{s}.'
        out = oracle.step(batch, noise, g[pre + 'in.perms'] if pre + 'in.perms' in g else None)
        assert rel_err(out['y'], g[pre + 'out.y']) < tol
        assert rel_err(out['loss_q'][0], g[pre + 'out.loss_q0']) < tol
        assert rel_err(out['entropy'], g[pre + 'out.c_entropy']) < tol
        for i in range(m['E']):
            for k, v in out['grad_q'][i].items():
                assert rel_err(v, g[f'{pre}grad.q{i}.{k}']) < tol, (s, i, k)
        for k, v in out['grad_policy'].items():
            assert rel_err(v, g[f'{pre}grad.pi.{k}']) < tol, (s, k)
        if hp.use_auto_alpha:
            assert rel_err(out['grad_log_alpha'], g[pre + 'grad.log_c_alpha']) < tol
        if hp.use_n_step_is:
            assert rel_err(out['pi_probs'], g[pre + 'out.pi_probs']) < 1e-5
        if hp.use_priority:
            assert rel_err(out['td_error'], g[pre + 'out.td_error']) < 1e-5
            assert rel_err(out['y_td'], g[pre + 'out.y_td']) < 1e-5
        snap = oracle.snapshot()
        for k, v in snap.items():
            assert rel_err(v, g[f'{pre}after.{k}']) < 1e-5, (s, k)


def check_recurrent_sac_steps(g: dict) -> None:
    """The GRU-representation flow (envs/test/nn_rnn.py, seq_encoder=RNN) against the real reference:
    BPTT gradients of the representation, re-encoded states, next hidden states, td error on the
    target states."""
    from tests.helpers import golden_rep_batch, rep_oracle_from_golden
    torch.set_num_threads(1)
    m = sac_case_meta(g)
    oracle = rep_oracle_from_golden(g)
    hp = oracle.hp
    tol = 2e-6
    for s in range(m['steps']):
        batch, noise = golden_rep_batch(g, s)
        out = oracle.step(batch, noise)
        pre = f's{s}.'
        assert rel_err(out['y'], g[pre + 'out.y']) < tol
        assert rel_err(out['target_states'], g[pre + 'out.target_states']) < tol
        assert rel_err(out['states_post'], g[pre + 'out.states_post']) < tol
        assert rel_err(out['next_hidden'], g[pre + 'out.next_hidden']) < tol
        for k, v in out['grad_rep'].items():
            assert rel_err(v, g[f'{pre}grad.rep.{k}']) < tol, (s, k)
            assert np.abs(g[f'{pre}grad.rep.{k}']).max() > 0, 'fixture should exercise the representation gradient'
        for i in range(m['E']):
            for k, v in out['grad_q'][i].items():
                assert rel_err(v, g[f'{pre}grad.q{i}.{k}']) < tol, (s, i, k)
        for k, v in out['grad_policy'].items():
            assert rel_err(v, g[f'{pre}grad.pi.{k}']) < tol, (s, k)
        if hp.use_auto_alpha:
            assert rel_err(out['grad_log_alpha'], g[pre + 'grad.log_c_alpha']) < tol
        if hp.use_n_step_is:
            assert rel_err(out['pi_probs'], g[pre + 'out.pi_probs']) < 1e-5
        if hp.use_priority:
            assert rel_err(out['td_error'], g[pre + 'out.td_error']) < 1e-5
            assert rel_err(out['y_td'], g[pre + 'out.y_td']) < 1e-5
        for k, v in oracle.snapshot().items():
            assert rel_err(v, g[f'{pre}after.{k}']) < 1e-5, (s, k)


def check_hybrid_sac_steps(g: dict) -> None:
    """Discrete / hybrid action branches (oracle/discrete_oracle.py) against a reference fixture."""
    from oracle.discrete_oracle import HybridHyper, SacHybridOracle
    torch.set_num_threads(1)
    m = sac_case_meta(g)
    base = sac_hyper_from_golden(g)
    hpv = {k[3:]: (float(v) if np.ndim(v) == 0 else torch.from_numpy(np.asarray(v, dtype=np.float32)))
           for k, v in g.items() if k.startswith('hp.')}
    hp = HybridHyper(**{**base.__dict__, 'd_action_sizes': [int(x) for x in g['d_action_sizes']],
                        'target_d_alpha': hpv['target_d_alpha'],
                        'd_policy_entropy_penalty': hpv['d_policy_entropy_penalty'], 'd_depth': 3,
                        'discrete_dqn_like': bool(hpv.get('discrete_dqn_like', 0.0))})
    oracle = SacHybridOracle(hp)
    oracle.load_params(*golden_params(g, 'init', m['E']), log_d_alpha=g['init.log_d_alpha'])
    tol = 2e-6
    for s in range(m['steps']):
        batch, noise = golden_batch(g, s)
        pre = f's{s}.'
        perms = None
        if hp.discrete_dqn_like:  # the reference's result depends on its randperm draws there (see get_y)
            assert hp.action_size == 0 and g[pre + 'in.perms'].shape[0] == (4 if hp.use_priority else 2)
            p = g[pre + 'in.perms']
            perms = [(p[0], p[1]), (p[2], p[3])] if hp.use_priority else [(p[0], p[1]), None]
        out = oracle.step(batch, noise, perms)
        assert rel_err(out['d_y'], g[pre + 'out.d_y']) < tol
        if hp.action_size:
            assert rel_err(out['y'], g[pre + 'out.y']) < tol
        for i in range(m['E']):
            for k, v in out['grad_q'][i].items():
                assert rel_err(v, g[f'{pre}grad.q{i}.{k}']) < tol, (s, i, k)
        for k, v in out.get('grad_policy', {}).items():
            if f'{pre}grad.pi.{k}' in g:  # (dqn-like hybrid runs leave the discrete heads of the policy without a gradient)
                assert rel_err(v, g[f'{pre}grad.pi.{k}']) < tol, (s, k)
        assert ('grad_policy' in out) == any(k.startswith(f'{pre}grad.pi.') for k in g)
        if 'grad_log_d_alpha' in out:
            assert rel_err(out['grad_log_d_alpha'], g[pre + 'grad.log_d_alpha']) < tol
        else:
            assert (pre + 'grad.log_d_alpha') not in g
        if hp.use_auto_alpha and False:
            assert rel_err(out['grad_log_d_alpha'], g[pre + 'grad.log_d_alpha']) < tol
            if hp.action_size:
                assert rel_err(out['grad_log_alpha'], g[pre + 'grad.log_c_alpha']) < tol
        if hp.use_n_step_is:
            assert rel_err(out['pi_probs'], g[pre + 'out.pi_probs']) < 1e-5
        if hp.use_priority:
            assert rel_err(out['td_error'], g[pre + 'out.td_error']) < 1e-5
        for k, v in oracle.snapshot().items():
            assert rel_err(v, g[f'{pre}after.{k}']) < 1e-5, (s, k)


def check_padding(g: dict) -> None:
    """sac_base.py:2435-2453 (learner-side padding of a sampled window) against a reference fixture."""
    b = int(g['meta'][0])
    raw = {k[4:]: v for k, v in g.items() if k.startswith('raw.')}
    out = pad_sampled_batch(raw, b, np.zeros(raw['action'].shape[-1], dtype=np.float32))
    assert np.array_equal(out['index'][:, :-1], g['padded.bn_indexes'])
    assert np.array_equal(out['padding_mask'][:, :-1], g['padded.bn_padding_masks'])
    assert np.array_equal(out['last_mask'][:, :-1], g['padded.bn_last_masks'])
    assert np.array_equal(out['action'][:, :-1], g['padded.bn_actions'])
    assert np.array_equal(out['reward'][:, :-1], g['padded.bn_rewards'])
    assert np.array_equal(out['done'][:, :-1], g['padded.bn_dones'])
    assert np.array_equal(out['mu_prob'][:, :-1], g['padded.bn_mu_probs'])
    assert np.array_equal(out['obs_vector'], g['padded.bnx_obs'])
    if b > 0:  # with b == 0 the zero-priority episode tail (ignore_size=1) keeps windows inside an episode
        assert out['padding_mask'].any(), 'fixture should contain padded rows'
