"""GPU: the reference's own test strategy (tests/test_sac_params.py: every combination of action kinds, replay /
on-policy, priorities, sequence encoders, DQN-like, n-step IS is constructed with the conv plugins and driven through
choose_action / choose_attn_action -> put_episode -> train() until the step counter moves) on the drop-in, with the
SAME plugin files (tests/nn_conv_vanilla.py / nn_conv_rnn.py / nn_conv_attn.py, verbatim under tests/golden/plugins),
the same observation shapes, burn-in 5 and n-step 3, synthetic episodes of tests/get_synthesis_data.py's shapes.
The reference's matrix has 2^k x 3 entries and asserts nothing but "runs"; this one walks a covering subset and also
checks that the step counter advances, the losses stay finite and the parameters move.  (Out of scope and raising:
siamese, RND, normalisation — SURVEY §8.)"""
import importlib.util
import itertools
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PLUGINS = Path(__file__).resolve().parent / 'golden' / 'plugins'
OBS_NAMES = ['vector', 'image']
OBS_SHAPES = [(10,), (3, 30, 30)]
FILES = {None: 'tests_nn_conv_vanilla.py', 'RNN': 'tests_nn_conv_rnn.py', 'ATTN': 'tests_nn_conv_attn.py'}


def _plugin(enc):
    path = PLUGINS / FILES[enc]
    spec = importlib.util.spec_from_file_location('nn_matrix_' + path.stem, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _actions(rng, lead, d_sizes, A):
    parts = [np.eye(k, dtype=np.float32)[rng.randint(0, k, size=lead)] for k in d_sizes]
    parts.append(np.clip(rng.rand(*lead, A), -.99, .99).astype(np.float32))
    return np.concatenate(parts, axis=-1)


def _cases():
    out = []
    for enc, d_sizes, A in itertools.product([None, 'RNN', 'ATTN'], [[], [3, 3, 4]], [0, 4]):
        if not d_sizes and not A:
            continue
        # every (encoder, action kind) with the replay + priorities + n-step IS; the other switches on a rotating subset
        out.append(dict(enc=enc, d=d_sizes, A=A, replay=True, pri=True, dqn=False, nis=True))
    out += [dict(enc=None, d=[3, 3, 4], A=4, replay=True, pri=False, dqn=True, nis=False),
            dict(enc='RNN', d=[3, 3, 4], A=0, replay=True, pri=True, dqn=True, nis=True),
            dict(enc='RNN', d=[], A=4, replay=False, pri=False, dqn=False, nis=True),
            dict(enc='ATTN', d=[3, 3, 4], A=4, replay=False, pri=False, dqn=False, nis=False),
            dict(enc=None, d=[3, 3, 4], A=0, replay=False, pri=False, dqn=False, nis=True),
            dict(enc='ATTN', d=[], A=4, replay=True, pri=False, dqn=False, nis=False)]
    return out


@pytest.mark.parametrize('p', _cases(), ids=lambda p: '-'.join(f'{k}={v}' for k, v in p.items()).replace(' ', ''))
def test_reference_parameter_matrix_runs(p):
    from algorithm.sac_base import SAC_Base
    from algorithm.utils.enums import SEQ_ENCODER
    rng = np.random.RandomState(7)
    enc = p['enc']
    sac = SAC_Base(obs_names=OBS_NAMES, obs_shapes=OBS_SHAPES, model_abs_dir=None, nn=_plugin(enc),
                   d_action_sizes=p['d'], c_action_size=p['A'], use_replay_buffer=p['replay'], use_priority=p['pri'],
                   burn_in_step=5, n_step=3, seq_encoder=None if enc is None else SEQ_ENCODER[enc],
                   discrete_dqn_like=p['dqn'], use_n_step_is=p['nis'], batch_size=16, seed=1,
                   replay_config={'capacity': 4096})
    hshape = tuple(sac.seq_hidden_state_shape)
    AF = sum(p['d']) + p['A']
    before = [x.detach().clone() for x in list(sac.model_rep.parameters()) + list(sac.model_q_list[0].parameters())]
    step, rounds = 0, 0
    while step < 4 and rounds < 12:
        rounds += 1
        if enc == 'ATTN':
            T = int(rng.randint(1, 40))
            act, prob, hid = sac.choose_attn_action(
                ep_indexes=np.arange(T, dtype=np.int32)[None].repeat(10, 0), ep_padding_masks=np.zeros((10, T), dtype=bool),
                ep_obses_list=[rng.randn(10, T, *s).astype(np.float32) for s in OBS_SHAPES],
                ep_pre_actions=_actions(rng, (10, T), p['d'], p['A']),
                ep_pre_attn_states=rng.randn(10, T, *hshape).astype(np.float32))
        else:
            act, prob, hid = sac.choose_action([rng.randn(10, *s).astype(np.float32) for s in OBS_SHAPES],
                                               _actions(rng, (10,), p['d'], p['A']),
                                               rng.randn(10, *hshape).astype(np.float32))
        assert act.shape == (10, AF) and prob.shape == (10, AF) and hid.shape == (10, *hshape)
        T = int(rng.randint(30, 100))
        sac.put_episode(ep_indexes=np.arange(T, dtype=np.int32)[None],
                        ep_obses_list=[rng.randn(1, T, *s).astype(np.float32) for s in OBS_SHAPES],
                        ep_actions=_actions(rng, (1, T), p['d'], p['A']), ep_rewards=rng.randn(1, T).astype(np.float32),
                        ep_dones=rng.randint(0, 2, size=(1, T)).astype(bool), ep_probs=rng.rand(1, T, AF).astype(np.float32),
                        ep_pre_seq_hidden_states=rng.randn(1, T, *hshape).astype(np.float32))
        step = sac.train()
    torch.cuda.synchronize()
    assert step >= 4, step
    after = list(sac.model_rep.parameters()) + list(sac.model_q_list[0].parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, after)), 'nothing trained'
    assert all(torch.isfinite(x).all() for x in after)
    if p['pri']:
        assert torch.isfinite(sac._wk['td_error']).all()
    sac.close()
