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


@pytest.mark.parametrize('name', ['sac_rnn.npz', 'sac_rnn_b0.npz'])
def test_recurrent_step_against_oracle_and_golden(name):
    """asac_sac_step_networks_rep + tail, consecutive steps from the golden initial parameters: every
    output against the oracle run on the same inputs and against the reference's own numbers."""
    from tests.cuda_harness import SacRepCuda
    torch.set_num_threads(1)
    g = load_golden(name)
    m = sac_case_meta(g)
    oracle = rep_oracle_from_golden(g)
    hp = oracle.hp
    cu = SacRepCuda(hp, m['B'], m['So'], m['rep_layers'])
    from tests.helpers import golden_params, golden_rep_params
    cu.load_params(*golden_params(g, 'init', m['E']))
    cu.load_rep(*golden_rep_params(g, 'init'))
    for s in range(m['steps']):
        batch, noise = golden_rep_batch(g, s)
        cb = cu.make_rep_batch(batch, noise)
        want = oracle.step(batch, noise)
        got = cu.step_rep(cb)
        torch.cuda.synchronize()
        pre = f's{s}.'
        assert rel_err(got['states'], want['states']) < TOL
        assert rel_err(got['target_states'], want['target_states']) < TOL
        assert rel_err(got['y'], want['y'].view(-1)) < TOL
        for i in range(m['E']):
            for k, v in got['grad_q'][i].items():
                assert rel_err(v, want['grad_q'][i][k]) < TOL, (s, i, k)
        for k, v in got['grad_rep'].items():
            assert rel_err(v, want['grad_rep'][k]) < TOL, (s, k, rel_err(v, want['grad_rep'][k]))
            assert rel_err(v, g[f'{pre}grad.rep.{k}']) < TOL, (s, k)
        assert rel_err(got['states_post'], want['states_post']) < TOL
        assert rel_err(got['states_post'], g[pre + 'out.states_post']) < TOL
        assert rel_err(got['next_hidden'], want['next_hidden']) < TOL
        assert rel_err(got['next_hidden'], g[pre + 'out.next_hidden']) < TOL
        for k, v in got['grad_policy'].items():
            assert rel_err(v, want['grad_policy'][k]) < TOL, (s, k)
        if hp.use_auto_alpha:
            assert rel_err(got['grad_log_alpha'], want['grad_log_alpha']) < TOL
        if hp.use_n_step_is:
            assert rel_err(got['pi_probs'], want['pi_probs']) < 5e-5
        if hp.use_priority:
            assert rel_err(got['y_td'], want['y_td'].view(-1)) < 5e-5
            assert rel_err(got['td_error'], want['td_error'].view(-1)) < 5e-5
            assert rel_err(got['td_error'], g[pre + 'out.td_error'].reshape(-1)) < 5e-5
        snap, ref = cu.snapshot(), oracle.snapshot()
        for k, v in snap.items():
            assert rel_err(v, ref[k]) < 2e-5, (s, k, rel_err(v, ref[k]))
        cnt = cu.counters.cpu().tolist()
        assert cnt[0] == s + 1 and cnt[1] == s + 1 and cnt[2] == s + 1 and cnt[4] == s + 1


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
    want = oracle.step(batch, noise)
    got = cu.step_rep(cu.make_rep_batch(batch, noise))
    torch.cuda.synchronize()
    assert rel_err(got['y'], want['y'].view(-1)) < TOL
    for i in range(2):
        for k, v in got['grad_q'][i].items():
            assert rel_err(v, want['grad_q'][i][k]) < TOL, (i, k)
    for k, v in got['grad_rep'].items():
        assert rel_err(v, want['grad_rep'][k]) < TOL, (k, rel_err(v, want['grad_rep'][k]))
    for k, v in got['grad_policy'].items():
        assert rel_err(v, want['grad_policy'][k]) < TOL, k
    assert rel_err(got['states_post'], want['states_post']) < TOL
    rel = np.abs(got['pi_probs'] - want['pi_probs'].numpy()) / np.maximum(1.0, np.abs(want['pi_probs'].numpy()))
    assert np.quantile(rel, 0.999) < 5e-5 and rel.max() < 5e-3  # heavy-tailed amplification (see test_gpu_sac)
    assert np.quantile(np.abs(got['td_error'] - want['td_error'].view(-1).numpy()), 0.999) < 1e-4
