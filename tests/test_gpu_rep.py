"""GPU: the recurrent-representation kernels (csrc/rep_gru.cu) and the SAC step with a trained GRU
representation (asac_sac_step_networks_rep) through the C ABI, against torch.nn.GRU / autograd and
against the CPU oracle pinned to the reference's envs/test/nn_rnn.py run (tests/golden/sac_rnn*.npz)."""
import numpy as np
import pytest
import torch

from tests.helpers import golden_rep_batch, load_golden, rel_err, rep_oracle_from_golden, sac_case_meta

pytestmark = pytest.mark.gpu

TOL = 1e-5  # relative to scale


def _random_gru(shape, seed):
    from oracle.rep_oracle import init_gru
    gen = torch.Generator().manual_seed(seed)
    return init_gru(shape.obs_size + shape.action_size, shape.hidden, shape.layers, gen)


@pytest.mark.parametrize('So,A,H,NL,B,L', [(6, 2, 8, 2, 37, 9), (3, 1, 16, 1, 8, 5), (10, 3, 40, 3, 5, 4), (6, 2, 64, 1, 3, 3),
                                           (5, 2, 16, 2, 9, 7), (7, 1, 32, 1, 6, 6), (5, 2, 10, 3, 7, 6), (4, 2, 8, 4, 5, 8)])
def test_gru_forward_matches_torch(So, A, H, NL, B, L):
    """States, every layer's hidden states and the actor-side single step against torch.nn.GRU
    (seq_layers.py:41-113 without a padding mask); two parameter sets in one launch."""
    from asac_b200 import lowering
    from asac_b200.nn_models import GRU
    from tests.cuda_harness import gru_forward
    shape = lowering.GruShape(So, A, H, NL)
    dev = torch.device('cuda:0')
    rng = np.random.RandomState(So * 100 + H)
    sds = [_random_gru(shape, 1), _random_gru(shape, 2)]
    flats = [lowering.gru_flat_from_state_dict(shape, sd).to(dev) for sd in sds]
    obs = torch.from_numpy(rng.randn(B, L, So).astype(np.float32))
    actions = torch.from_numpy(rng.rand(B, L - 1, A).astype(np.float32) * 2 - 1)
    h0 = torch.from_numpy(rng.randn(B, NL, H).astype(np.float32) * 0.5)
    outs = gru_forward(shape, flats, obs.to(dev), actions.to(dev), None, h0.to(dev), save=True)
    pre = torch.cat([torch.zeros(B, 1, A), actions], dim=1)
    for sd, out in zip(sds, outs):
        ref = GRU(So + A, H, NL)
        ref.load_state_dict({k[len('rnn.'):]: v for k, v in sd.items()})
        with torch.no_grad():
            states, hn = ref(torch.cat([obs, pre], -1), h0)
        assert rel_err(out['states'].cpu(), states) < TOL
        assert rel_err(out['hn'].cpu(), hn) < TOL
        assert torch.equal(out['hn'][:, :, -1], out['states'])
    # actor side: one step, pre_action handed over directly, zero initial state
    pa = torch.from_numpy(rng.rand(B, 1, A).astype(np.float32))
    one = gru_forward(shape, flats[:1], obs[:, :1].contiguous().to(dev), None, pa.to(dev), None)[0]
    ref = GRU(So + A, H, NL)
    ref.load_state_dict({k[len('rnn.'):]: v for k, v in sds[0].items()})
    with torch.no_grad():
        states, hn = ref(torch.cat([obs[:, :1], pa], -1), None)
    assert rel_err(one['states'].cpu(), states) < TOL and rel_err(one['hn'].cpu(), hn) < TOL


@pytest.mark.parametrize('So,A,H,NL,B,L,tg,E', [(6, 2, 8, 2, 19, 9, 5, 2), (3, 1, 16, 1, 8, 5, 0, 1),
                                                 (10, 3, 40, 3, 6, 6, 5, 3), (6, 2, 8, 2, 7, 46, 40, 2),
                                                 (5, 2, 16, 2, 9, 7, 4, 2), (7, 1, 32, 1, 6, 6, 5, 1),
                                                 (5, 2, 10, 3, 7, 6, 3, 2), (4, 2, 8, 4, 5, 8, 7, 2)])
def test_gru_backward_matches_autograd(So, A, H, NL, B, L, tg, E):
    """BPTT of a state gradient applied at step t_grad, against autograd through the oracle's
    restatement of the cell (in float64, so that the comparison measures the kernel alone)."""
    from asac_b200 import lowering
    from oracle.rep_oracle import gru_forward as oracle_gru
    from tests.cuda_harness import gru_backward, gru_forward
    shape = lowering.GruShape(So, A, H, NL)
    dev = torch.device('cuda:0')
    rng = np.random.RandomState(tg * 7 + H)
    sd = _random_gru(shape, 3)
    flat = lowering.gru_flat_from_state_dict(shape, sd).to(dev)
    obs = torch.from_numpy(rng.randn(B, L, So).astype(np.float32))
    actions = torch.from_numpy(rng.rand(B, L - 1, A).astype(np.float32) * 2 - 1)
    h0 = torch.from_numpy(rng.randn(B, NL, H).astype(np.float32) * 0.5)
    gs = torch.from_numpy(rng.randn(E, B, H).astype(np.float32))
    fwd = gru_forward(shape, [flat], obs.to(dev), actions.to(dev), None, h0.to(dev), save=True)[0]
    grad, part = gru_backward(shape, flat, obs.to(dev), actions.to(dev), h0.to(dev), tg, gs.to(dev), fwd)
    assert not torch.isnan(part[:, :shape.count]).any(), 'every partial-gradient element must be written'
    p64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    pre = torch.cat([torch.zeros(B, 1, A), actions], dim=1)
    states, _ = oracle_gru(p64, NL, torch.cat([obs, pre], -1).double(), h0.double())
    (states[:, tg] * gs.double().sum(0)).sum().backward()
    ref = lowering.gru_flat_from_state_dict(shape, {k: v.grad for k, v in p64.items()})[:shape.count]
    got = grad.cpu()
    for k, v in lowering.gru_state_dict_from_flat(shape, got).items():
        r = lowering.gru_state_dict_from_flat(shape, ref)[k]
        assert rel_err(v, r) < TOL, (k, rel_err(v, r))


def _gap_checks():
    from tests.test_gpu_sac import Checks
    return Checks()


def _compare_recurrent_step(ck, pre, got, want, want64, hp, E, golden=None):
    """Every output of one step: err(CUDA, oracle fp32) <= TOL + slack * gap, gap = the oracle's own
    fp32-vs-fp64 distance for that quantity (tests/test_gpu_sac.py::Checks).  BPTT through a long burn-in
    and exp(log_prob) are the quantities where fp32 itself is not determined to 1e-5."""
    v = lambda d: d.view(-1) if isinstance(d, torch.Tensor) else d
    for k in ('states', 'target_states', 'states_post', 'next_hidden'):
        ck.add(pre + k, got[k], want[k], want64[k])
    ck.add(pre + 'y', got['y'], v(want['y']), v(want64['y']))
    for i in range(E):
        for k, g_ in got['grad_q'][i].items():
            ck.add(f'{pre}grad.q{i}.{k}', g_, want['grad_q'][i][k], want64['grad_q'][i][k])
    for k, g_ in got['grad_rep'].items():
        ck.add(f'{pre}grad.rep.{k}', g_, want['grad_rep'][k], want64['grad_rep'][k])
    for k, g_ in got['grad_policy'].items():
        ck.add(f'{pre}grad.pi.{k}', g_, want['grad_policy'][k], want64['grad_policy'][k])
    if hp.use_auto_alpha:
        ck.add(pre + 'grad.log_alpha', got['grad_log_alpha'], want['grad_log_alpha'], want64['grad_log_alpha'])
    if hp.use_n_step_is:  # heavy-tailed max statistic (see test_gpu_sac._oracle_stage_step): slack 5
        ck.add(pre + 'pi_probs', got['pi_probs'], want['pi_probs'], want64['pi_probs'], slack=5.0)
    if hp.use_priority:
        # y' consumes pi_probs as mu: it inherits their noise through the IS ratios
        ck.add(pre + 'y_td', got['y_td'], v(want['y_td']), v(want64['y_td']), slack=5.0)
        ck.add(pre + 'td_error', got['td_error'], v(want['td_error']), v(want64['td_error']), slack=5.0)
    if golden is not None:
        g, gp = golden
        for k in ('states_post', 'next_hidden'):
            ck.add(f'{pre}golden.{k}', got[k], g[gp + 'out.' + k], want64[k])
        for k, g_ in got['grad_rep'].items():
            ck.add(f'{pre}golden.grad.rep.{k}', g_, g[f'{gp}grad.rep.{k}'], want64['grad_rep'][k])
        if hp.use_priority:
            ck.add(pre + 'golden.td_error', got['td_error'], g[gp + 'out.td_error'].reshape(-1), v(want64['td_error']),
                   slack=5.0)


@pytest.mark.parametrize('name', ['sac_rnn.npz', 'sac_rnn_b0.npz', 'sac_rnn_c4.npz'])
def test_recurrent_step_against_oracle_and_golden(name):
    """asac_sac_step_networks_rep + tail, consecutive steps from the golden initial parameters: every
    output against the oracle run on the same inputs and against the reference's own numbers.
    sac_rnn_c4 is BASELINE configs[3] at its full shape: 256 sequences, burn-in 40 + n_step 5 (46-row
    windows, BPTT through 41 steps), padding inside the burn-in, minted from the reference's
    envs/test/nn_rnn.py run."""
    from tests.cuda_harness import SacRepCuda
    from tests.helpers import golden_params, golden_rep_params
    from tests.test_gpu_sac import adam_excess
    torch.set_num_threads(4)
    g = load_golden(name)
    m = sac_case_meta(g)
    oracle = rep_oracle_from_golden(g)
    o64 = rep_oracle_from_golden(g, dtype=torch.float64)
    hp = oracle.hp
    cu = SacRepCuda(hp, m['B'], m['So'], m['rep_layers'])
    cu.load_params(*golden_params(g, 'init', m['E']))
    cu.load_rep(*golden_rep_params(g, 'init'))
    ck = _gap_checks()
    for s in range(m['steps']):
        batch, noise = golden_rep_batch(g, s)
        cb = cu.make_rep_batch(batch, noise)
        o64.copy_state_from(oracle, adam=True)   # the float64 twin starts every step from the fp32 oracle's state
        want = oracle.step(batch, noise)
        want64 = o64.step(batch.to(torch.float64), noise.to(torch.float64))
        got = cu.step_rep(cb)
        torch.cuda.synchronize()
        _compare_recurrent_step(ck, f's{s}.', got, want, want64, hp, m['E'], golden=(g, f's{s}.'))
        # post-Adam parameters: 1e-5, plus Adam's scale-invariant term where a gradient component is itself
        # rounding noise (tests/test_gpu_sac.py header); the Adam kernel is pinned on identical gradients there
        snap, ref = cu.snapshot(), oracle.snapshot()
        grads = {f'q{i}.{k}': (got['grad_q'][i][k], want['grad_q'][i][k]) for i in range(m['E']) for k in got['grad_q'][i]}
        grads.update({f'pi.{k}': (got['grad_policy'][k], want['grad_policy'][k]) for k in got['grad_policy']})
        grads.update({f'rep.{k}': (got['grad_rep'][k], want['grad_rep'][k]) for k in got['grad_rep']})
        for k, val in snap.items():
            if k in grads:
                ex, _ = adam_excess(val, ref[k], grads[k][0], np.asarray(grads[k][1]), hp.learning_rate)
                ck.raw(f's{s}.after.{k}', ex * TOL)
            else:  # target networks (Polyak), log_alpha
                ck.raw(f's{s}.after.{k}', rel_err(val, ref[k]))
        cnt = cu.counters.cpu().tolist()
        assert cnt[0] == s + 1 and cnt[1] == s + 1 and cnt[2] == s + 1 and cnt[4] == s + 1
        cu.sync_from_oracle(oracle)  # next step from identical parameters and Adam moments
    ck.dump(name.replace('.npz', ''))
    bad = ck.bad()
    print(f'{name}: {len(ck.rows)} checks; worst {ck.report()}')
    assert not bad, f'{len(bad)}/{len(ck.rows)} over {TOL} + slack*gap: {ck.report(bad)}'


def test_recurrent_step_large_batch_shares_policy_rows():
    """B = 1024 sequences of 5 rows: 16 batch elements per tile, so the post pass holds 128 policy rows
    per tile and the two cluster ranks each run half of them and exchange the head outputs over
    distributed shared memory (k_value_pass, `share`).  Random parameters and inputs against the oracle."""
    from oracle.rep_oracle import SacRepBatch, SacRepOracle
    from oracle.sac_oracle import SacHyper, SacNoise
    from tests.cuda_harness import SacRepCuda
    torch.set_num_threads(4)
    B, So, A, NL, H, b, n = 1024, 6, 2, 2, 8, 2, 2
    L = b + n + 1
    hp = SacHyper(state_size=H, action_size=A, ensemble_q_num=2, hidden=64, q_depth=3, policy_depth=3,
                  burn_in_step=b, n_step=n, clip_epsilon=0.0, v_lambda=0.9)
    oracle = SacRepOracle(hp, So, NL, seed=5)
    with torch.no_grad():  # targets differ from the online nets, biases off zero
        for net in oracle.q_target + [oracle.rep_target]:
            for t in net.values():
                t.add_(torch.randn_like(t) * 0.02)
    cu = SacRepCuda(hp, B, So, NL)
    assert cu.tile == 16
    snap = lambda d: {k: v.detach() for k, v in d.items()}
    cu.load_params([snap(q) for q in oracle.q], [snap(q) for q in oracle.q_target], snap(oracle.policy),
                   float(oracle.log_c_alpha))
    cu.load_rep(snap(oracle.rep), snap(oracle.rep_target))
    rng = np.random.RandomState(9)
    t = lambda x: torch.from_numpy(x)
    pad = np.zeros((B, L - 1), dtype=bool)
    pad[rng.rand(B) < 0.2, :1] = True
    batch = SacRepBatch(obs=t(rng.randn(B, L, So).astype(np.float32)),
                        hidden0=t((rng.randn(B, NL, H) * 0.5).astype(np.float32)),
                        actions=t((rng.rand(B, L - 1, A) * 1.8 - 0.9).astype(np.float32)),
                        rewards=t(rng.randn(B, L - 1).astype(np.float32)), dones=t(rng.rand(B, L - 1) < 0.1),
                        mu_probs=t((rng.rand(B, L - 1, A) + 0.05).astype(np.float32)),
                        last_masks=t(rng.rand(B, L - 1) < 0.05), padding_masks=t(pad),
                        priority_is=t((rng.rand(B, 1) * 0.9 + 0.1).astype(np.float32)))
    noise = SacNoise(eps_y=t(rng.randn(B, n + 1, A).astype(np.float32)), eps_pi=t(rng.randn(B, A).astype(np.float32)),
                     eps_alpha=t(rng.randn(B, A).astype(np.float32)), eps_td=t(rng.randn(B, n + 1, A).astype(np.float32)))
    o64 = SacRepOracle(hp, So, NL, seed=5, dtype=torch.float64)
    o64.copy_state_from(oracle)
    want = oracle.step(batch, noise)
    want64 = o64.step(batch.to(torch.float64), noise.to(torch.float64))
    got = cu.step_rep(cu.make_rep_batch(batch, noise))
    torch.cuda.synchronize()
    ck = _gap_checks()
    _compare_recurrent_step(ck, 'b1024.', got, want, want64, hp, 2)
    ck.dump('rep_large_b1024')
    bad = ck.bad()
    print(f'large batch: {len(ck.rows)} checks; worst {ck.report()}')
    assert not bad, f'{len(bad)}/{len(ck.rows)} over {TOL} + slack*gap: {ck.report(bad)}'
