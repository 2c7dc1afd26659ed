"""GPU: discrete and hybrid action branches (csrc/sac_discrete.cuh, asac_b200/discrete.py) against fixtures minted
by running the reference's ``_train`` / ``get_l_probs`` / ``_get_td_error`` with ``d_action_sizes`` on the same batches
(oracle/gen_golden.py:gen_sac_discrete_case; the oracle of the same path is pinned to them in test_oracle_golden.py).
Stock nets of envs/test/nn.py (loaded verbatim), 1e-5 relative to scale for every stage of every step from the
reference's parameters; parameters after each free-running step within Adam's per-step envelope."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu

PLUGIN = Path(__file__).resolve().parent / 'golden' / 'plugins' / 'envs_test_nn.py'
TOL = 1e-5


def _plugin():
    spec = importlib.util.spec_from_file_location('nn_plugin_discrete', PLUGIN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _learner(g, **extra):
    from algorithm.sac_base import SAC_Base
    S, A, E, qh, qd, B, b, n, steps, use_pri = (int(x) for x in g['meta'])
    hp = {k[3:]: v for k, v in g.items() if k.startswith('hp.')}
    sizes = [int(x) for x in g['d_action_sizes']]
    ratio = float(np.asarray(hp['target_d_alpha']).reshape(-1)[0] / np.log(sizes[0]))
    sac = SAC_Base(obs_names=['vector'], obs_shapes=[(S,)], d_action_sizes=sizes, c_action_size=A, model_abs_dir=None,
                   nn=_plugin(), seed=3, batch_size=B, burn_in_step=b, n_step=n, ensemble_q_num=E, ensemble_q_sample=E,
                   use_priority=bool(use_pri), tau=float(hp['tau']),
                   update_target_per_step=int(hp['update_target_per_step']), learning_rate=float(hp['learning_rate']),
                   gamma=float(hp['gamma']), v_lambda=float(hp['v_lambda']), v_rho=float(hp['v_rho']), v_c=float(hp['v_c']),
                   clip_epsilon=float(hp['clip_epsilon']), use_n_step_is=bool(hp['use_n_step_is']),
                   target_c_alpha=float(hp['target_c_alpha']), target_d_alpha=ratio,
                   d_policy_entropy_penalty=float(hp['d_policy_entropy_penalty']),
                   init_log_alpha=float(hp['init_log_alpha']), use_auto_alpha=bool(hp['use_auto_alpha']),
                   discrete_dqn_like=bool(float(hp.get('discrete_dqn_like', 0.0))),
                   replay_config={'capacity': 1024}, **extra)
    assert sac._disc is not None and sac._disc.D == sum(sizes)
    return sac


def _load(sac, g, prefix):
    def sub(tag):
        pre = f'{prefix}.{tag}.'
        return {k[len(pre):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(pre)}
    with torch.no_grad():
        for i in range(sac.ensemble_q_num):
            sac.model_q_list[i].load_state_dict(sub(f'q{i}'))
            sac.model_target_q_list[i].load_state_dict(sub(f'qt{i}'))
        sac.model_policy.load_state_dict(sub('pi'))
        sac.log_c_alpha.copy_(torch.from_numpy(np.asarray(g[f'{prefix}.log_c_alpha'])))
        sac.log_d_alpha.copy_(torch.from_numpy(np.asarray(g[f'{prefix}.log_d_alpha'])))


def _fill(sac, g, s):
    st = sac._sets[0]
    bt, dev, D = st['bt'], sac.device, sac._disc.D
    pre = f's{s}.in.'
    put = lambda dst, src: dst[:, :src.shape[1]].copy_(torch.from_numpy(np.ascontiguousarray(src)).to(dev))
    bt['states'].copy_(torch.from_numpy(g[pre + 'states']).to(dev))
    put(bt['actions_full'], g[pre + 'actions'])
    put(bt['mu_full'], g[pre + 'mu_probs'])
    if sac.c_action_size:
        put(bt['actions'], g[pre + 'actions'][..., D:])
        put(bt['mu_probs'], g[pre + 'mu_probs'][..., D:])
    put(bt['rewards'], g[pre + 'rewards'])
    put(bt['dones'], g[pre + 'dones'].astype(np.uint8))
    put(bt['last_masks'], g[pre + 'last_masks'].astype(np.uint8))
    put(bt['padding_masks'], g[pre + 'padding_masks'].astype(np.uint8))
    if pre + 'priority_is' in g:
        st['smp']['w'].copy_(torch.from_numpy(g[pre + 'priority_is'].reshape(-1)).to(dev))
    if sac.c_action_size:
        noise = np.concatenate([g[pre + k].reshape(-1) for k in ('eps_y', 'eps_pi', 'eps_alpha', 'eps_td')])
        st['noise'].copy_(torch.from_numpy(noise).to(dev))
    return st


def _d_named(sac, flat):
    """Flat discrete-head buffer of one member -> tensors under the reference's state_dict names."""
    out, dq = {}, sac._disc
    for k, shape in enumerate(dq.shapes):
        off, kin, H = int(dq.branch_off[k]), shape.in_dim, shape.hidden

        def take(n, *view):
            nonlocal off
            t = flat[off:off + n].view(*view)
            off += n
            return t
        pre = f'd_dense_list.{k}.dense.'
        for layer in range(shape.depth):
            out[f'{pre}{2 * layer}.linear.weight'] = take(H * kin, H, kin)
            out[f'{pre}{2 * layer}.linear.bias'] = take(H, H)
            kin = H
        out[f'{pre}{2 * shape.depth}.weight'] = take(shape.out_dim * H, shape.out_dim, H)
        out[f'{pre}{2 * shape.depth}.bias'] = take(shape.out_dim, shape.out_dim)
    return out


def _policy_gradient_gap(g, s):
    """|fp32 - fp64| of the oracle's policy gradient on step `s` of the fixture, relative to scale, per parameter."""
    from oracle.discrete_oracle import HybridHyper, SacHybridOracle
    from tests.helpers import golden_batch, golden_params, sac_case_meta, sac_hyper_from_golden
    torch.set_num_threads(1)
    m, base = sac_case_meta(g), sac_hyper_from_golden(g)
    hp = HybridHyper(**{**base.__dict__, 'd_action_sizes': [int(x) for x in g['d_action_sizes']],
                        'target_d_alpha': torch.from_numpy(np.asarray(g['hp.target_d_alpha'], dtype=np.float32)),
                        'd_policy_entropy_penalty': float(g['hp.d_policy_entropy_penalty']), 'd_depth': 3})
    prefix = 'init' if s == 0 else f's{s - 1}.after'
    batch, noise = golden_batch(g, s)
    grads = []
    for dtype in (torch.float32, torch.float64):
        o = SacHybridOracle(hp, dtype=dtype)
        o.load_params(*golden_params(g, prefix, m['E']), log_d_alpha=g[f'{prefix}.log_d_alpha'])
        b, nz = batch.to(dtype), noise.to(dtype)
        o.polyak(hp.tau)
        o.train_q(b, nz.eps_y)
        grads.append(o.train_policy(b, nz.eps_pi)['grad_policy'])
    return {k: rel_err(grads[0][k].numpy(), grads[1][k].numpy()) for k in grads[0] if not k.startswith('d_dense_list')}


@pytest.mark.parametrize('name', ['sac_disc.npz', 'sac_hybrid.npz', 'sac_dqn.npz'])
def test_discrete_step_matches_reference(name):
    from asac_b200 import lowering
    g = load_golden(name)
    sac = _learner(g)
    E, steps = int(g['meta'][2]), int(g['meta'][8])
    dq = sac._disc
    lr = float(g['hp.learning_rate'])
    for s in range(steps):
        _load(sac, g, 'init' if s == 0 else f's{s - 1}.after')  # every step from the reference's own parameters
        if s > 0:  # ... and its optimizer state is what the previous step left: keep ours (checked below)
            pass
        st = _fill(sac, g, s)
        if f's{s}.in.perms' in g:  # DQN-like: the reference's (target, online) shuffles of _train_rep_q and _get_td_error
            sac._ens_perms[5:5 + len(g[f's{s}.in.perms'])].copy_(torch.from_numpy(g[f's{s}.in.perms'].astype(np.int32)))
        sac._discrete_step_networks(st)
        torch.cuda.synchronize()
        pre = f's{s}.'
        err = {'d_y': rel_err(dq.wk['d_y'].cpu().numpy(), g[pre + 'out.d_y'].reshape(-1))}
        if sac.c_action_size:
            err['y'] = rel_err(sac._wk['y'].cpu().numpy(), g[pre + 'out.y'].reshape(-1))
        for i in range(E):
            for k, v in _d_named(sac, dq.wk['grad_q'][i]).items():
                err[f'grad.q{i}.{k}'] = rel_err(v.cpu().numpy(), g[f'{pre}grad.q{i}.{k}'])
            if sac.c_action_size:
                for k, v in lowering.state_dict_from_flat(sac._q_shape, sac._wk['grad_q'][i], policy=False).items():
                    err[f'grad.q{i}.{k}'] = rel_err(v.cpu().numpy(), g[f'{pre}grad.q{i}.{k}'])
        for k, v in _d_named(sac, dq.wk['grad_pi']).items():
            if f'{pre}grad.pi.{k}' in g:  # (a DQN-like run leaves the policy's discrete heads without a gradient)
                err[f'grad.pi.{k}'] = rel_err(v.cpu().numpy(), g[f'{pre}grad.pi.{k}'])
        if sac.c_action_size:
            for k, v in lowering.state_dict_from_flat(sac._pi_shape, sac._wk['grad_pi'], policy=True).items():
                err[f'grad.pi.{k}'] = rel_err(v.cpu().numpy(), g[f'{pre}grad.pi.{k}'])
        if pre + 'grad.log_d_alpha' in g:
            err['grad.log_d_alpha'] = rel_err(dq.wk['grad_alpha'].cpu().numpy(), g[pre + 'grad.log_d_alpha'].reshape(-1))
        if pre + 'out.pi_probs' in g:
            err['pi_probs'] = rel_err(sac._wk['pi_probs_full'].cpu().numpy(), g[pre + 'out.pi_probs'])
        if pre + 'out.td_error' in g:
            err['td_error'] = rel_err(sac._wk['td_error'].cpu().numpy(), g[pre + 'out.td_error'].reshape(-1))
        print(name, 'step', s, 'worst:', sorted(err.items(), key=lambda kv: -kv[1])[:6])
        if s == 0:  # (later steps start from the reference's parameters but OUR Adam moments: gradients still comparable)
            pass
        # The continuous policy gradient (hybrid case) is ill-conditioned in fp32 — the torch-fp32 reference itself sits
        # a few 1e-5 from the float64 evaluation of the same formulas (tests/test_gpu_sac.py header) — so those
        # components get the rule used there: TOL + 2 x (the oracle's own fp32-vs-fp64 gap).  Everything else: TOL.
        gap = _policy_gradient_gap(g, s) if sac.c_action_size else {}
        bad = {k: v for k, v in err.items()
               if not v < TOL + 2 * gap.get(k[len('grad.pi.'):], 0.0) + (4 * TOL if k == 'pi_probs' else 0.0)}
        assert not bad, (bad, gap)
        # after the step: Adam moves every component by <= lr whatever the gradient's size
        for tag, mod in [('pi', sac.model_policy)] + [(f'q{i}', sac.model_q_list[i]) for i in range(E)] + \
                [(f'qt{i}', sac.model_target_q_list[i]) for i in range(E)]:
            for k, t in mod.state_dict().items():
                d = float(np.max(np.abs(t.cpu().numpy() - g[f'{pre}after.{tag}.{k}'])))
                assert d <= (2.5 * lr if s == 0 else 2.5 * lr), (tag, k, d)
        if s == 0:
            assert abs(float(sac.log_d_alpha) - float(g[pre + 'after.log_d_alpha'])) < 1e-5
            assert abs(float(sac.log_c_alpha) - float(g[pre + 'after.log_c_alpha'])) < 1e-5
    sac.close()


@pytest.mark.parametrize('sizes,A,dqn', [([3, 4], 0, False), ([3], 2, False), ([4, 2], 0, True), ([3], 2, True)])
def test_discrete_learner_trains_end_to_end(sizes, A, dqn):
    """put_episode -> train() x 6 (CUDA graph from the second step on) -> choose_action, discrete-only and hybrid."""
    from algorithm.sac_base import SAC_Base
    D = sum(sizes)
    sac = SAC_Base(obs_names=['vector'], obs_shapes=[(6,)], d_action_sizes=sizes, c_action_size=A, model_abs_dir=None,
                   nn=_plugin(), seed=5, batch_size=32, n_step=3, discrete_dqn_like=dqn,
                   replay_config={'capacity': 2048})
    rng = np.random.RandomState(1)
    for _ in range(6):
        T = 40
        onehot = np.concatenate([np.eye(k, dtype=np.float32)[rng.randint(0, k, size=T)] for k in sizes], axis=-1)
        act = np.concatenate([onehot, (rng.rand(T, A) * 1.8 - 0.9).astype(np.float32)], axis=-1)[None]
        sac.put_episode(ep_indexes=np.arange(T, dtype=np.int32)[None], ep_obses_list=[rng.randn(1, T, 6).astype(np.float32)],
                        ep_actions=act, ep_rewards=rng.randn(1, T).astype(np.float32),
                        ep_dones=np.zeros((1, T), dtype=bool), ep_probs=rng.rand(1, T, D + A).astype(np.float32),
                        ep_pre_seq_hidden_states=np.zeros((1, T, 0), dtype=np.float32))
    before = [p.detach().clone() for p in sac.model_policy.parameters()]
    before_q = [p.detach().clone() for p in sac.model_q_list[0].parameters()]
    mu0 = sac.replay_buffer._columns['mu_prob'].clone()
    for i in range(6):
        assert sac.train() == i + 1
    torch.cuda.synchronize()
    assert all(not torch.equal(a, p) for a, p in zip(before_q, sac.model_q_list[0].parameters()))
    if not dqn:
        assert all(not torch.equal(a, p) for a, p in zip(before, sac.model_policy.parameters()))
        assert float(sac.log_d_alpha) != pytest.approx(-2.3, abs=1e-9)
    assert not torch.equal(mu0, sac.replay_buffer._columns['mu_prob'])
    assert torch.isfinite(sac._wk['td_error']).all()
    act, prob, _ = sac.choose_action([rng.randn(5, 6).astype(np.float32)], np.zeros((5, D + A), dtype=np.float32),
                                     np.zeros((5, 0), dtype=np.float32))
    assert act.shape == (5, D + A) and prob.shape == (5, D + A)
    c = 0
    for k in sizes:  # one-hot per branch; probabilities of a branch sum to one (DQN-like: they stay 1, :951-956)
        assert np.all(act[:, c:c + k].sum(-1) == 1)
        assert np.allclose(prob[:, c:c + k].sum(-1), k if dqn else 1, atol=1e-5)
        c += k
    # checkpoint state dicts carry both halves of every optimizer
    sd = sac.optimizer_q_list[0].state_dict()
    assert len(sd['state']) == len(list(sac.model_q_list[0].parameters()))
    assert set(sac.optimizer_alpha.state_dict()['state']) == ({1} if dqn and A else set() if dqn else {0, 1} if A else {0})
    sac.close()
