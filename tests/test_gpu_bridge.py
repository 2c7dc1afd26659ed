"""GPU: representations that run as the plugin's torch module (asac_b200/rep_bridge.py) — the reference's own
test plugins (tests/nn_conv_vanilla.py, nn_conv_rnn.py, nn_conv_attn.py: convolutional encoder, packed GRU,
episode attention), loaded VERBATIM from tests/golden/plugins, against fixtures minted by running the reference's
``_train`` / ``get_l_probs`` / ``_get_td_error`` on the same batches (oracle/gen_golden.py:gen_sac_plugin_rep_case).

Everything but the representation runs on this repo's kernels; the representation's forward / backward is
cuDNN / cuBLAS through torch with TF32 off.  Tolerance: 1e-5 relative to scale for every stage of the first
step (2e-5 for quantities downstream of the convolution's backward, whose cuDNN algorithm sums in another
order than the reference's CPU kernels), the Adam-aware envelope for parameters after a step."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu

PLUGINS = Path(__file__).resolve().parent / 'golden' / 'plugins'
CASES = ['sac_conv_vanilla.npz', 'sac_conv_rnn.npz', 'sac_conv_attn.npz']
TOL = 1e-5


def _plugin(rel: str):
    path = PLUGINS / rel.replace('/', '_')
    spec = importlib.util.spec_from_file_location('nn_' + path.stem, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _learner(g, **extra):
    from algorithm.sac_base import SAC_Base
    from algorithm.utils.enums import SEQ_ENCODER
    S, A, E, qh, qd, B, b, n, steps, use_pri = (int(x) for x in g['meta'])
    hp = {k[3:]: float(v) for k, v in g.items() if k.startswith('hp.')}
    enc = str(g['seq_encoder'])
    obs_names = [str(x) for x in g['obs_names']]
    obs_shapes = [tuple(int(x) for x in g[f'obs_shape{i}']) for i in range(len(obs_names))]
    sac = SAC_Base(obs_names=obs_names, obs_shapes=obs_shapes, d_action_sizes=[], c_action_size=A, model_abs_dir=None,
                   nn=_plugin(str(g['nn_rel'])), seed=3, batch_size=B, burn_in_step=b, n_step=n, ensemble_q_num=E,
                   ensemble_q_sample=E, seq_encoder=SEQ_ENCODER[enc] if enc else None, use_priority=bool(use_pri),
                   tau=hp['tau'], update_target_per_step=int(hp['update_target_per_step']),
                   learning_rate=hp['learning_rate'], gamma=hp['gamma'], v_lambda=hp['v_lambda'], v_rho=hp['v_rho'],
                   v_c=hp['v_c'], clip_epsilon=hp['clip_epsilon'], use_n_step_is=bool(hp['use_n_step_is']),
                   target_c_alpha=hp['target_c_alpha'], init_log_alpha=hp['init_log_alpha'],
                   use_auto_alpha=bool(hp['use_auto_alpha']), replay_config={'capacity': 1024}, **extra)
    assert sac._bridge is not None
    assert sac.state_size == S and tuple(sac.seq_hidden_state_shape) == tuple(int(x) for x in g['hidden_shape'])
    return sac, obs_shapes


def _load(sac, g, prefix):
    def sub(tag):
        pre = f'{prefix}.{tag}.'
        return {k[len(pre):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(pre)}
    with torch.no_grad():
        for i in range(sac.ensemble_q_num):
            sac.model_q_list[i].load_state_dict(sub(f'q{i}'))
            sac.model_target_q_list[i].load_state_dict(sub(f'qt{i}'))
        sac.model_policy.load_state_dict(sub('pi'))
        sac.model_rep.load_state_dict(sub('rep'))
        sac.model_target_rep.load_state_dict(sub('rept'))
        sac.log_c_alpha.copy_(torch.from_numpy(np.asarray(g[f'{prefix}.log_c_alpha'])))


def _fill(sac, g, s, obs_shapes):
    st = sac._sets[0]
    bt, dev = st['bt'], sac.device
    pre = f's{s}.in.'
    L = sac._cfg.seq_len
    put = lambda dst, src: dst[:, :src.shape[1]].copy_(torch.from_numpy(np.ascontiguousarray(src)).to(dev))
    bt['obs_list'] = [torch.from_numpy(g[f'{pre}obs{i}']).to(dev) for i in range(len(obs_shapes))]
    put(bt['index'], g[pre + 'index'])
    put(bt['actions'], g[pre + 'actions'])
    put(bt['rewards'], g[pre + 'rewards'])
    put(bt['dones'], g[pre + 'dones'].astype(np.uint8))
    put(bt['mu_probs'], g[pre + 'mu_probs'])
    put(bt['last_masks'], g[pre + 'last_masks'].astype(np.uint8))
    put(bt['padding_masks'], g[pre + 'padding_masks'].astype(np.uint8))
    bt['hidden'].copy_(torch.from_numpy(g[pre + 'hidden']).reshape(bt['hidden'].shape).to(dev))
    if pre + 'priority_is' in g:
        st['smp']['w'].copy_(torch.from_numpy(g[pre + 'priority_is'].reshape(-1)).to(dev))
    noise = np.concatenate([g[pre + k].reshape(-1) for k in ('eps_y', 'eps_pi', 'eps_alpha', 'eps_td')])
    st['noise'].copy_(torch.from_numpy(noise).to(dev))
    return st


def _named_flat(sac, which, flat):
    from asac_b200 import lowering
    if which == 'pi':
        return lowering.state_dict_from_flat(sac._pi_shape, flat, policy=True)
    return lowering.state_dict_from_flat(sac._q_shape, flat, policy=False)


@pytest.mark.parametrize('name', CASES)
def test_plugin_representation_step_matches_reference(name):
    g = load_golden(name)
    sac, obs_shapes = _learner(g)
    E, B, b, n, steps = (int(g['meta'][i]) for i in (2, 5, 6, 7, 8))
    L = b + n + 1
    _load(sac, g, 'init')
    worst = {}
    for s in range(steps):
        st = _fill(sac, g, s, obs_shapes)
        hidden_post = sac._bridge_step_networks(st)
        torch.cuda.synchronize()
        wk, bt, br = sac._wk, st['bt'], sac._bridge
        pre = f's{s}.'
        err = {'y': rel_err(wk['y'].cpu().numpy(), g[pre + 'out.y'].reshape(-1)),
               'target_states': rel_err(bt['target_states'].cpu().numpy(), g[pre + 'out.target_states']),
               'states_post': rel_err(bt['states_post'].cpu().numpy(), g[pre + 'out.states_post'])}
        if int(np.prod(sac.seq_hidden_state_shape)):
            err['next_hidden'] = rel_err(hidden_post[:, :-1].cpu().numpy(), g[pre + 'out.next_hidden'])
        off = 0
        for k, p in sac.model_rep.named_parameters():
            got = br.grad_out[off:off + p.numel()].view(p.shape).cpu().numpy()
            err[f'grad.rep.{k}'] = rel_err(got, g[f'{pre}grad.rep.{k}'])
            off += p.numel()
        for i in range(E):
            for k, v in _named_flat(sac, 'q', wk['grad_q'][i]).items():
                err[f'grad.q{i}.{k}'] = rel_err(v.cpu().numpy(), g[f'{pre}grad.q{i}.{k}'])
        for k, v in _named_flat(sac, 'pi', wk['grad_pi']).items():
            err[f'grad.pi.{k}'] = rel_err(v.cpu().numpy(), g[f'{pre}grad.pi.{k}'])
        if pre + 'out.pi_probs' in g:
            err['pi_probs'] = rel_err(wk['pi_probs'].cpu().numpy(), g[pre + 'out.pi_probs'])
        if pre + 'out.td_error' in g:
            err['td_error'] = rel_err(wk['td_error'].cpu().numpy(), g[pre + 'out.td_error'].reshape(-1))
        print(name, 'step', s, 'worst:', sorted(err.items(), key=lambda kv: -kv[1])[:5])
        for k, v in err.items():
            worst[k] = max(worst.get(k, 0.), v)
        if s == 0:  # from identical parameters: every stage
            bad = {k: v for k, v in err.items() if not v < (2 * TOL if k.startswith(('grad.rep', 'grad.pi', 'pi_probs'))
                                                            else TOL)}
            assert not bad, bad
        # parameters after the step: Adam moves every component by <= lr per step whatever the gradient's size
        lr = float(g['hp.learning_rate'])
        for tag, mod in [('rep', sac.model_rep), ('rept', sac.model_target_rep), ('pi', sac.model_policy)] + \
                [(f'q{i}', sac.model_q_list[i]) for i in range(E)]:
            for k, t in mod.state_dict().items():
                d = float(np.max(np.abs(t.cpu().numpy() - g[f'{pre}after.{tag}.{k}'])))
                assert d <= 2.5 * lr * (s + 1), (tag, k, d)
        assert abs(float(sac.log_c_alpha) - float(g[pre + 'after.log_c_alpha'])) < 1e-5
    sac.close()


@pytest.mark.parametrize('name', CASES)
def test_plugin_representation_trains_end_to_end(name):
    """put_episode -> train() x 5 -> choose_action / choose_attn_action with the conv plugins, as the reference's own
    tests/test_sac_params.py drives them; the stored hidden states and mu-probs get written back."""
    g = load_golden(name)
    sac, obs_shapes = _learner(g)
    A = int(g['meta'][1])
    hshape = tuple(sac.seq_hidden_state_shape)
    rng = np.random.RandomState(5)
    for _ in range(6):
        T = int(rng.randint(30, 60))
        sac.put_episode(ep_indexes=np.arange(T, dtype=np.int32)[None],
                        ep_obses_list=[rng.randn(1, T, *s).astype(np.float32) for s in obs_shapes],
                        ep_actions=(rng.rand(1, T, A) * 1.8 - 0.9).astype(np.float32),
                        ep_rewards=rng.randn(1, T).astype(np.float32),
                        ep_dones=np.zeros((1, T), dtype=bool), ep_probs=rng.rand(1, T, A).astype(np.float32),
                        ep_pre_seq_hidden_states=(rng.randn(1, T, *hshape) * 0.1).astype(np.float32))
    before = [p.detach().clone() for p in sac.model_rep.parameters()]
    hid0 = sac.replay_buffer._columns['pre_seq_hidden_state'].clone()
    for i in range(5):
        assert sac.train() == i + 1
    torch.cuda.synchronize()
    assert any(not torch.equal(a, p) for a, p in zip(before, sac.model_rep.parameters())), 'representation did not train'
    if int(np.prod(hshape)):
        assert not torch.equal(hid0, sac.replay_buffer._columns['pre_seq_hidden_state'])
    stats = sac.last_step_stats()
    assert all(np.isfinite(v) for v in stats.values()), stats
    rows = 4
    if str(g['seq_encoder']) == 'ATTN':
        T = 9
        act, prob, hidden = sac.choose_attn_action(
            ep_indexes=np.arange(T, dtype=np.int32)[None].repeat(rows, 0), ep_padding_masks=np.zeros((rows, T), dtype=bool),
            ep_obses_list=[rng.randn(rows, T, *s).astype(np.float32) for s in obs_shapes],
            ep_pre_actions=rng.rand(rows, T, A).astype(np.float32),
            ep_pre_attn_states=rng.randn(rows, T, *hshape).astype(np.float32))
    else:
        act, prob, hidden = sac.choose_action([rng.randn(rows, *s).astype(np.float32) for s in obs_shapes],
                                              rng.rand(rows, A).astype(np.float32),
                                              rng.randn(rows, *hshape).astype(np.float32))
    assert act.shape == (rows, A) and prob.shape == (rows, A) and hidden.shape == (rows, *hshape)
    assert np.all(np.isfinite(act)) and np.all(np.abs(act) <= 1.0)
    sac.close()
