"""Pins the CPU oracle (oracle/*.py) to the fixtures minted from the real reference
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle.replay_oracle import PerOracle, SumTreeOracle, pad_sampled_batch
from oracle.sac_oracle import SacOracle
from tests.helpers import (golden_batch, golden_params, load_golden, rel_err, sac_case_meta,
                           sac_hyper_from_golden)

EP_KEYS = ['index', 'last_mask', 'obs_vector', 'obs_image', 'action', 'reward', 'done', 'mu_prob',
           'pre_seq_hidden_state']


@pytest.mark.parametrize('name', ['per_small.npz', 'per_zeros.npz'])
def test_per_trace_bit_exact(name):
    g = load_golden(name)
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


def test_rebuild_equals_incremental():
    g = load_golden('per_zeros.npz')
    capacity = int(g['meta'][0])
    t = SumTreeOracle(capacity)
    ref = g['r1.tree_after_add']
    t.nodes[capacity - 1:] = ref[capacity - 1:]
    t.rebuild()
    assert np.array_equal(t.nodes, ref)


@pytest.mark.parametrize('name', ['pad_b2n3.npz', 'pad_b0n1.npz'])
def test_padding_matches_reference(name):
    g = load_golden(name)
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


@pytest.mark.parametrize('name', ['sac_c2.npz', 'sac_c3.npz', 'sac_odd.npz', 'sac_nois.npz'])
def test_sac_oracle_matches_reference(name):
    torch.set_num_threads(1)
    g = load_golden(name)
    m = sac_case_meta(g)
    hp = sac_hyper_from_golden(g)
    oracle = SacOracle(hp)
    oracle.load_params(*golden_params(g, 'init', m['E']))
    tol = 2e-6  # two fp32 CPU evaluations of the same graph; the CUDA gate is 1e-5 (relative to scale)
    for s in range(m['steps']):
        batch, noise = golden_batch(g, s)
        out = oracle.step(batch, noise)
        pre = f's{s}.'
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


@pytest.mark.parametrize('name', ['sac_rnn.npz', 'sac_rnn_b0.npz'])
def test_recurrent_sac_oracle_matches_reference(name):
    """The GRU-representation flow (envs/test/nn_rnn.py, seq_encoder=RNN) against the real reference:
    BPTT gradients of the representation, re-encoded states, next hidden states, td error on the
    target states."""
    from tests.helpers import golden_rep_batch, rep_oracle_from_golden
    torch.set_num_threads(1)
    g = load_golden(name)
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


def test_vectorized_descent_equals_scalar():
    rng = np.random.RandomState(5)
    t = SumTreeOracle(1024)
    leaves = rng.rand(1024).astype(np.float32)
    leaves[rng.rand(1024) < 0.3] = 0
    t.nodes[1023:] = leaves
    t.rebuild()
    v = t.draw(512, rng.random_sample(512))
    a, pa = t.descend(v)
    b, pb = t.descend_vectorized(v)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)
