"""GPU: the SAC_Base drop-in end to end — plugin files load unchanged, train() advances, the CUDA
graph replays the eager sequence bit-for-bit, checkpoints round-trip."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PLUGINS = Path(__file__).resolve().parent / 'golden' / 'plugins'
# VERBATIM copies of the reference's plugin files (tests/test_plugin_surface.py compares them byte for byte
# with the checkout where it is mounted): envs/test/nn.py, envs/gym/pendulum/nn.py, envs/test/nn_rnn.py
PLUGIN_TEST = PLUGINS / 'envs_test_nn.py'
PLUGIN_PENDULUM = PLUGINS / 'envs_gym_pendulum_nn.py'
PLUGIN_RNN = PLUGINS / 'envs_test_nn_rnn.py'


def _plugin(tmp_path: Path, path: Path, name: str):
    """Executes the plugin file the way sac_main.py:353-364 does."""
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _episode(rng, obs_shapes, A, T, hidden_shape=(0,)):
    return dict(ep_indexes=np.arange(T, dtype=np.int32)[None],
                ep_obses_list=[rng.randn(1, T, *s).astype(np.float32) for s in obs_shapes],
                ep_actions=rng.rand(1, T, A).astype(np.float32), ep_rewards=rng.randn(1, T).astype(np.float32),
                ep_dones=rng.randint(0, 2, size=(1, T)).astype(bool), ep_probs=rng.rand(1, T, A).astype(np.float32),
                ep_pre_seq_hidden_states=(rng.randn(1, T, *hidden_shape) * 0.3).astype(np.float32))


def _make(nn, graph, obs_shapes=((6,),), A=2, hidden_shape=(0,), **kw):
    from algorithm.sac_base import SAC_Base  # the alias package: reference-style import
    kw.setdefault('batch_size', 64)
    kw.setdefault('replay_config', {'capacity': 4096})
    sac = SAC_Base(obs_names=[f'o{i}' for i in range(len(obs_shapes))], obs_shapes=list(obs_shapes),
                   d_action_sizes=[], c_action_size=A, model_abs_dir=None, nn=nn, seed=7, use_cuda_graph=graph, **kw)
    rng = np.random.RandomState(0)
    for _ in range(12):
        sac.put_episode(**_episode(rng, obs_shapes, A, 50, hidden_shape))
    return sac


@pytest.mark.parametrize('plugin,kw', [(PLUGIN_TEST, dict()),
                                       (PLUGIN_PENDULUM, dict(n_step=5, v_lambda=1.0, use_n_step_is=True)),
                                       (PLUGIN_TEST, dict(burn_in_step=2, n_step=3, use_priority=False))])
def test_graph_replay_equals_eager(tmp_path, plugin, kw):
    nn = _plugin(tmp_path, plugin, 'nn_plugin')
    a = _make(nn, graph=False, **kw)
    b = _make(nn, graph=True, **kw)
    for i in range(6):
        assert a.train() == i + 1
        assert b.train() == i + 1
    torch.cuda.synchronize()
    assert b._graph is not None and a._graph is None
    for pa, pb in zip(a.model_policy.parameters(), b.model_policy.parameters()):
        assert torch.equal(pa, pb)
    for qa, qb in zip(a.model_q_list + a.model_target_q_list, b.model_q_list + b.model_target_q_list):
        for pa, pb in zip(qa.parameters(), qb.parameters()):
            assert torch.equal(pa, pb)
    assert torch.equal(a.replay_buffer._nodes, b.replay_buffer._nodes)
    assert torch.equal(a.replay_buffer._columns['mu_prob'], b.replay_buffer._columns['mu_prob'])
    assert torch.equal(a.log_c_alpha, b.log_c_alpha)
    stats = b.last_step_stats()
    assert all(np.isfinite(v) for v in stats.values()), stats
    assert float(a.log_c_alpha) != pytest.approx(-2.3, abs=1e-9)  # alpha moved
    a.close(); b.close()


@pytest.mark.parametrize('plugin,kw', [(PLUGIN_TEST, dict()),
                                       (PLUGIN_PENDULUM, dict(n_step=5, v_lambda=1.0, use_n_step_is=True)),
                                       (PLUGIN_RNN, dict(burn_in_step=4, n_step=3, seq_encoder='RNN'))])
def test_programmatic_dependent_launch_does_not_change_results(tmp_path, plugin, kw):
    """The step's kernels start ahead of their predecessor under programmatic dependent launch and wait
    (griddepcontrol.wait) in front of their first dependent read: the same learner with the attribute off
    (asac_set_pdl(0): plain stream order) must hold bit-identical parameters, tree and write-backs."""
    from asac_b200 import _lib
    lib = _lib.load()
    nn = _plugin(tmp_path, plugin, 'nn_plugin')
    kw = dict(kw)
    extra = {}
    if kw.get('seq_encoder') == 'RNN':
        from algorithm.utils.enums import SEQ_ENCODER
        kw['seq_encoder'] = SEQ_ENCODER.RNN
        extra = dict(obs_shapes=((6,),), hidden_shape=(2, 8))
    learners = []
    before = lib.asac_set_pdl(1)
    try:
        for on in (1, 0):
            lib.asac_set_pdl(on)
            sac = _make(nn, graph=True, **extra, **kw)
            for _ in range(8):
                sac.train()
            torch.cuda.synchronize()
            learners.append(sac)
    finally:
        lib.asac_set_pdl(before)
    a, b = learners
    mods = lambda s: [s.model_rep, s.model_policy] + list(s.model_q_list) + list(s.model_target_q_list)
    for ma, mb in zip(mods(a), mods(b)):
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(pa, pb)
    assert torch.equal(a.log_c_alpha, b.log_c_alpha)
    a.flush_priority_update(); b.flush_priority_update()
    assert torch.equal(a.replay_buffer._nodes, b.replay_buffer._nodes)
    assert torch.equal(a.replay_buffer._columns['mu_prob'], b.replay_buffer._columns['mu_prob'])
    assert torch.equal(a._wk['td_error'], b._wk['td_error'])
    a.close(); b.close()


def test_strict_add_order_equals_the_undeferred_schedule(tmp_path, monkeypatch):
    """add() gives new rows the maximum leaf priority.  With the priority update deferred to the next step's parallel
    branch that maximum lags one update; ASAC_STRICT_ADD_ORDER=1 applies the pending update first.  Interleaving
    put_episode() and train() it must then retrace the schedule that never defers (ASAC_DEFER_TREE=0) bit for bit."""
    nn = _plugin(tmp_path, PLUGIN_TEST, 'nn_plugin')
    runs = []
    for env in ({'ASAC_DEFER_TREE': '0'}, {'ASAC_STRICT_ADD_ORDER': '1'}):
        for k in ('ASAC_DEFER_TREE', 'ASAC_STRICT_ADD_ORDER'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sac = _make(nn, graph=True)
        rng = np.random.RandomState(3)
        for i in range(10):
            ep = _episode(rng, ((6,),), 2, 20)
            ep['ep_rewards'] = ep['ep_rewards'] * (1.0 + i)  # growing td errors: the maximum priority keeps moving
            sac.put_episode(**ep)
            sac.train()
        sac.flush_priority_update()
        torch.cuda.synchronize()
        runs.append(sac)
    a, b = runs
    assert b._defer_tree and not a._defer_tree and b.replay_buffer._strict_add_order
    assert torch.equal(a.replay_buffer._nodes, b.replay_buffer._nodes)
    for pa, pb in zip(a.model_policy.parameters(), b.model_policy.parameters()):
        assert torch.equal(pa, pb)
    for qa, qb in zip(a.model_q_list, b.model_q_list):
        for pa, pb in zip(qa.parameters(), qb.parameters()):
            assert torch.equal(pa, pb)
    a.close(); b.close()


def test_train_returns_step_until_buffer_exceeds_batch(tmp_path):
    from algorithm.sac_base import SAC_Base
    nn = _plugin(tmp_path, PLUGIN_TEST, 'nn_plugin2')
    sac = SAC_Base(obs_names=['v'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2, model_abs_dir=None, nn=nn,
                   batch_size=32, replay_config={'capacity': 256})
    assert sac.train() == 0  # sac_base.py:2503-2506
    rng = np.random.RandomState(0)
    sac.put_episode(**_episode(rng, [(6,)], 2, 20))
    assert sac.train() == 0
    sac.put_episode(**_episode(rng, [(6,)], 2, 20))
    assert sac.train() == 1
    action, prob, hidden = sac.choose_action([rng.randn(5, 6).astype(np.float32)],
                                             np.zeros((5, 2), np.float32), np.zeros((5, 0), np.float32))
    assert action.shape == (5, 2) and prob.shape == (5, 2) and hidden.shape == (5, 0)
    assert np.all(np.abs(action) <= 1)
    sac.close()


def test_module_parameters_alias_kernel_storage(tmp_path):
    """The nn.Module parameters are views of the flat buffers the kernels update."""
    nn = _plugin(tmp_path, PLUGIN_TEST, 'nn_plugin3')
    sac = _make(nn, graph=False)
    w = sac.model_q_list[1].c_dense.dense[0].linear.weight
    before = w.detach().clone()
    sac.train()
    torch.cuda.synchronize()
    assert not torch.equal(before, w)
    flat = sac._q_flat[1, :w.numel()].view_as(w)
    assert flat.data_ptr() == w.data_ptr() and torch.equal(flat, w)
    # hard-copied targets at start, polyak afterwards: target != online after a step
    assert not torch.equal(sac._q_flat, sac._qt_flat)
    sac.close()


def test_checkpoint_roundtrip_and_reference_key_names(tmp_path):
    nn = _plugin(tmp_path, PLUGIN_TEST, 'nn_plugin4')
    from algorithm.sac_base import SAC_Base
    mk = lambda: SAC_Base(obs_names=['v'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2,
                          model_abs_dir=tmp_path / 'run', nn=nn, batch_size=32, summary_path=None,
                          replay_config={'capacity': 256}, seed=1)
    sac = mk()
    rng = np.random.RandomState(0)
    for _ in range(3):
        sac.put_episode(**_episode(rng, [(6,)], 2, 30))
    for _ in range(3):
        sac.train()
    sac.save_model(save_replay_buffer=True)
    saved = torch.load(tmp_path / 'run' / 'model' / '3.pth', weights_only=True)
    assert set(saved) == {'global_step', 'model_q_0', 'model_target_q_0', 'optimizer_q_0', 'model_q_1',
                          'model_target_q_1', 'optimizer_q_1', 'model_policy', 'optimizer_policy', 'log_d_alpha',
                          'log_c_alpha', 'optimizer_alpha'}
    assert 'c_dense.dense.0.linear.weight' in saved['model_q_0']
    assert 'mean_dense.dense.0.weight' in saved['model_policy']
    assert float(saved['optimizer_q_0']['state'][0]['step']) == 3
    other = mk()
    assert other.get_global_step() == 3
    assert torch.equal(other._q_flat, sac._q_flat) and torch.equal(other._pi_m, sac._pi_m)
    assert torch.equal(other._counters, sac._counters)
    assert torch.equal(other.replay_buffer._nodes, sac.replay_buffer._nodes)
    assert other.replay_buffer.size == sac.replay_buffer.size
    assert other.train() == 4
    sac.close(); other.close()


def test_unsupported_configurations_fail_loudly(tmp_path):
    from algorithm.sac_base import SAC_Base
    nn = _plugin(tmp_path, PLUGIN_TEST, 'nn_plugin5')
    base = dict(obs_names=['v'], obs_shapes=[(6,)], c_action_size=2, model_abs_dir=None, nn=nn)
    with pytest.raises(NotImplementedError):
        SAC_Base(d_action_sizes=[3], ensemble_q_num=3, ensemble_q_sample=2, **base)  # (discrete branches: all critics only)
    with pytest.raises(NotImplementedError):
        SAC_Base(d_action_sizes=[], use_rnd=True, **base)
    with pytest.raises(Exception):
        SAC_Base(d_action_sizes=[], device='cpu', **base)


def test_choose_action_kernels_match_torch_modules(tmp_path):
    """choose_action on asac_policy_act (FFMA forward for small batches, tcgen05 forward from 2048
    rows on) against the plugin's torch modules on the same parameters, same injected draws
    (sac_base.py:943-964): actions and per-dimension probabilities within 1e-5 of scale; the
    deterministic and offline-action branches too."""
    sac = _make(_plugin(tmp_path, PLUGIN_TEST, 'nn_plugin_act'), graph=True, batch_size=32)
    rng = np.random.RandomState(0)
    with torch.no_grad():
        for p in sac.model_policy.parameters():
            p.add_(torch.randn_like(p) * 0.05)  # in place: the kernels read the same storage
    for rows, tc in ((1, False), (37, False), (4096, False), (4096, True)):
        sac.actor_tensor_cores = tc
        p_tol = 2e-3 if tc else 1e-4  # 3xTF32 pre-activations (1.5e-6) amplified by |x - mu| / sigma^2
        obs = [rng.randn(rows, 6).astype(np.float32)]
        eps = rng.randn(rows, 2).astype(np.float32)
        pre_a, pre_h = np.zeros((rows, 2), np.float32), np.zeros((rows, 0), np.float32)
        for kw in (dict(eps=eps), dict(disable_sample=True), dict(offline_action=rng.rand(rows, 2).astype(np.float32))):
            a0, p0, h0 = sac.choose_action(obs, pre_a, pre_h, **kw)
            a1, p1, h1 = sac._choose_action_torch(obs, pre_a, pre_h, **kw)
            assert a0.shape == (rows, 2) and p0.shape == (rows, 2) and h0.shape == h1.shape == (rows, 0)
            assert np.max(np.abs(a0 - a1)) < 1e-5, (rows, kw.keys())
            # prob goes through atanh(tanh(x)): one ulp of the action is amplified by 1 / (1 - a^2) (x 50 at
            # |a| = 0.99) and again by |x - mu| / sigma^2 in the Gaussian, so the MAX relative difference of
            # two fp32 evaluations over thousands of samples is a heavy-tailed statistic: bound the bulk
            # tightly and the tail loosely (actions themselves agree to 1e-5 above)
            rel = (np.abs(p0 - p1) / np.maximum(1.0, np.abs(p1))).reshape(-1)
            assert np.median(rel) < (1e-5 if tc else 1e-6) and np.quantile(rel, 0.99) < p_tol and rel.max() < 100 * p_tol, \
                (rows, tc, kw.keys(), float(np.median(rel)), float(np.quantile(rel, 0.99)), float(rel.max()))
    # on-device Philox draws: different every call, finite, inside the squashed range
    a, p, _ = sac.choose_action([rng.randn(64, 6).astype(np.float32)], None, None)
    b, _, _ = sac.choose_action([rng.randn(64, 6).astype(np.float32)], None, None)
    assert np.all(np.abs(a) < 1) and np.all(np.isfinite(p)) and not np.array_equal(a, b)
    sac.close()


def test_recurrent_representation_learner(tmp_path):
    """seq_encoder=RNN with the envs/test/nn_rnn.py form of ModelRep: the plugin loads unchanged, the
    representation trains, the CUDA graph replays the eager sequence bit for bit, the window gather
    feeds the GRU kernels the stored observations / first hidden state, the next hidden states go
    back to pre_seq_hidden_state of the following rows (sac_base.py:2589-2596), choose_action runs
    one GRU step + the policy on the kernels and agrees with the torch modules."""
    from algorithm.utils.enums import SEQ_ENCODER
    nn = _plugin(tmp_path, PLUGIN_RNN, 'nn_plugin_rnn')
    kw = dict(seq_encoder=SEQ_ENCODER.RNN, burn_in_step=4, n_step=3, hidden_shape=(2, 8), batch_size=32,
              replay_config={'capacity': 1024})
    a = _make(nn, graph=False, **kw)
    b = _make(nn, graph=True, **kw)
    assert a.state_size == 8 and tuple(a.seq_hidden_state_shape) == (2, 8)
    assert a.get_initial_seq_hidden_state(3).shape == (3, 2, 8)
    rep0 = a._rep_flat.clone()
    hid0 = a.replay_buffer._columns['pre_seq_hidden_state'].clone()
    # one eager step, then check the data movement around the kernels against the ring
    assert a.train() == 1 and b.train() == 1
    torch.cuda.synchronize()
    rb, cap, L, bi = a.replay_buffer, a.replay_buffer.capacity, 8, 4
    ids = a._smp['ids'].cpu().numpy()
    pad = a._bt['padding_masks'].cpu().numpy().astype(bool)
    obs_ring = rb._columns['obs_o0'].cpu().numpy()
    got_obs = a._bt['obs'].cpu().numpy()
    hn_post = a._rw['hn_post'].cpu().numpy().reshape(len(ids), L, 2, 8)
    ring_h = rb._columns['pre_seq_hidden_state'].cpu().numpy().reshape(cap, 2, 8)
    store_ids = rb._store_ids.cpu().numpy()
    targets = {}
    for i, d in enumerate(ids):
        for t in range(L):  # observations are copied as stored, padded rows included (sac_base.py:2435-2453)
            assert np.array_equal(got_obs[i, t], obs_ring[(d - bi + t) % cap]), (i, t)
        h_first = a._bt['hidden'].cpu().numpy()[i, 0].reshape(2, 8)
        want_first = np.zeros((2, 8), np.float32) if pad[i, 0] else hid0.cpu().numpy().reshape(cap, 2, 8)[(d - bi) % cap]
        assert np.array_equal(h_first, want_first), i
        for t in range(L - 1):
            if not pad[i, t] and store_ids[(d + 1 - bi + t) % cap] == d + 1 - bi + t:
                targets.setdefault(int(d + 1 - bi + t), []).append((i, t))
    assert len(targets) > 50
    for tid, srcs in targets.items():
        if len({(hn_post[i, t]).tobytes() for i, t in srcs}) == 1:  # a single writer (or identical rows)
            i, t = srcs[0]
            assert np.array_equal(ring_h[tid % cap], hn_post[i, t]), (tid, srcs)
    for i in range(2, 7):
        assert a.train() == i and b.train() == i
    torch.cuda.synchronize()
    assert b._graph is not None and a._graph is None
    assert not torch.equal(rep0, a._rep_flat), 'the representation must train'
    assert not torch.equal(a._rep_flat, a._rept_flat)
    for x, y in ((a._rep_flat, b._rep_flat), (a._rept_flat, b._rept_flat), (a._q_flat, b._q_flat),
                 (a._pi_flat, b._pi_flat), (a._rep_m, b._rep_m), (a.replay_buffer._nodes, b.replay_buffer._nodes),
                 (a.replay_buffer._columns['pre_seq_hidden_state'], b.replay_buffer._columns['pre_seq_hidden_state']),
                 (a.replay_buffer._columns['mu_prob'], b.replay_buffer._columns['mu_prob'])):
        assert torch.equal(x, y)
    assert not torch.equal(hid0, a.replay_buffer._columns['pre_seq_hidden_state'])
    assert int(a._counters[4]) == 6
    assert all(np.isfinite(v) for v in b.last_step_stats().values())
    # parameters of the torch modules alias the flat buffers the kernels train
    w = a.model_rep.rnn._grus[1].weight_hh_l0
    assert w.data_ptr() >= a._rep_flat.data_ptr() and torch.equal(
        w, a._rep_flat[w.data_ptr() - a._rep_flat.data_ptr() >> 2:][:w.numel()].view_as(w))
    # actor side
    rng = np.random.RandomState(3)
    rows = 33
    obs = [rng.randn(rows, 6).astype(np.float32)]
    pre_a, pre_h = rng.rand(rows, 2).astype(np.float32), (rng.randn(rows, 2, 8) * 0.3).astype(np.float32)
    eps = rng.randn(rows, 2).astype(np.float32)
    a0, p0, h0 = a.choose_action(obs, pre_a, pre_h, eps=eps)
    a1, p1, h1 = a._choose_action_torch(obs, pre_a, pre_h, eps=eps)
    assert h0.shape == (rows, 2, 8) and np.max(np.abs(h0 - h1)) < 1e-5
    assert np.max(np.abs(a0 - a1)) < 1e-5
    assert np.median(np.abs(p0 - p1) / np.maximum(1.0, np.abs(p1))) < 1e-5
    # checkpoint keys of a run with a trained representation (sac_base.py:506-509)
    assert {'model_rep', 'model_target_rep', 'optimizer_rep'} <= set(a.ckpt_dict)
    sd = a.ckpt_dict['optimizer_rep'].state_dict()
    assert float(sd['state'][0]['step']) == 6 and len(sd['state']) == 8
    a.close(); b.close()


@pytest.mark.parametrize('name', ['vector', 'rnn'])
def test_resume_from_a_checkpoint_written_by_the_reference(tmp_path, name):
    """tests/golden/ckpt_<name>/ was written by the REFERENCE (`save_model(save_replay_buffer=True)` after 4
    train() steps, oracle/gen_golden.py:gen_ckpt_case): the B200 learner restores networks, optimizers,
    sum tree and transition storage from those files (sac_base.py:568-629; replay_buffer.py:96-111,
    220-227), reproduces the reference's deterministic action on recorded observations and trains on."""
    import shutil
    from tests.helpers import GOLDEN
    shutil.copytree(GOLDEN / f'ckpt_{name}', tmp_path / 'run')
    exp = np.load(tmp_path / 'run' / 'expect.npz')
    kw = {}
    if name == 'rnn':
        from algorithm.utils.enums import SEQ_ENCODER
        nn, kw = _plugin(tmp_path, PLUGIN_RNN, 'nn_plugin_ckpt_rnn'), dict(seq_encoder=SEQ_ENCODER.RNN, burn_in_step=3, n_step=2)
    else:
        nn = _plugin(tmp_path, PLUGIN_TEST, 'nn_plugin_ckpt')
    from algorithm.sac_base import SAC_Base
    sac = SAC_Base(obs_names=['vector'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2,
                   model_abs_dir=tmp_path / 'run', nn=nn, batch_size=8, summary_path=None, save_model_per_step=10 ** 9,
                   replay_config={'capacity': 64}, **kw)
    saved = torch.load(tmp_path / 'run' / 'model' / '4.pth', weights_only=True)
    assert sac.get_global_step() == int(exp['global_step']) == 4
    for key, mod in sac.ckpt_dict.items():
        assert key in saved, key
        if isinstance(mod, torch.nn.Module):
            for k, v in mod.state_dict().items():
                assert torch.equal(v.cpu(), saved[key][k]), (key, k)
    assert set(saved) == set(sac.ckpt_dict)
    assert float(sac.log_c_alpha) == float(saved['log_c_alpha'])
    cnt = sac._counters.cpu().tolist()
    assert cnt[0] == 4 and cnt[1] == 4 and cnt[2] == 4 and cnt[3] == 4 and (name != 'rnn' or cnt[4] == 4)
    st = saved['optimizer_policy']['state'][0]
    w0 = next(iter(sac.model_policy.parameters()))
    off = (w0.data_ptr() - sac._pi_flat.data_ptr()) // 4
    assert torch.equal(sac._pi_m[off:off + w0.numel()].view_as(w0).cpu(), st['exp_avg'])
    # replay: the reference's tree is nodes[1:], its storage columns are the rings
    rb = sac.replay_buffer
    tree = np.load(tmp_path / 'run' / 'model' / '4-rb_tree.npy')
    assert np.array_equal(rb._nodes.cpu().numpy()[1:], tree)
    store = np.load(tmp_path / 'run' / 'model' / '4-rb_storage.npz')
    assert rb.size == int(exp['rb_size']) == int(store['p_size'])
    for k in store.files:
        if k in ('p_size', 'p_id'):
            continue
        col = rb._store_ids if k == '_id' else rb._columns[k]
        assert np.array_equal(col.cpu().numpy().reshape(store[k].shape), store[k]), k
    a, p, h = sac.choose_action([exp['obs']], exp['pre_action'], exp['pre_hidden'], disable_sample=True)
    assert np.max(np.abs(a - exp['action'])) < 1e-5
    assert np.max(np.abs(p - exp['prob']) / np.maximum(1.0, np.abs(exp['prob']))) < 1e-4
    assert h.shape == exp['hidden'].shape and (h.size == 0 or np.max(np.abs(h - exp['hidden'])) < 1e-5)
    assert sac.train() == 5
    torch.cuda.synchronize()
    assert all(np.isfinite(v) for v in sac.last_step_stats().values())
    sac.close()


def test_on_policy_batch_buffer_path(tmp_path):
    """use_replay_buffer=False (sac_base.py:642-646, 2341-2349, 2509-2515): episodes go through the
    BatchBuffer, train() consumes one shuffled batch of windows per call and runs the same kernels with no
    IS weights / priorities — bit-identical to asac_sac_step driven through the C ABI on the same batch,
    parameters and Gaussian draws."""
    from algorithm.sac_base import SAC_Base
    from oracle.sac_oracle import SacBatch, SacHyper, SacNoise
    from tests.cuda_harness import SacCuda
    nn = _plugin(tmp_path, PLUGIN_TEST, 'nn_plugin_onpolicy')
    B, b, n, A = 16, 2, 3, 2
    sac = SAC_Base(obs_names=['o0'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=A, model_abs_dir=None, nn=nn,
                   batch_size=B, burn_in_step=b, n_step=n, use_replay_buffer=False, seed=3)
    assert not hasattr(sac, 'replay_buffer') and sac.train() == 0  # nothing buffered yet
    rng = np.random.RandomState(1)
    np.random.seed(5)
    for _ in range(2):
        sac.put_episode(**_episode(rng, [(6,)], A, 30))   # 2 x 29 windows -> 3 full batches, 10 left over
    seen = {}
    pop = sac.batch_buffer.get_batch

    def tap():
        seen['batch'] = pop()
        return seen['batch']
    sac.batch_buffer.get_batch = tap
    hp = SacHyper(state_size=6, action_size=A, burn_in_step=b, n_step=n, use_priority=False)
    cu = SacCuda(hp, B)
    cu.q.copy_(sac._q_flat); cu.qt.copy_(sac._qt_flat); cu.pi.copy_(sac._pi_flat)
    cu.log_alpha.copy_(sac._log_alpha_buf)
    assert sac.train() == 1
    torch.cuda.synchronize()
    (idx, last, pad, obs_list, actions, rewards, dones, mu, hidden) = seen['batch']
    assert idx.shape == (B, b + n) and obs_list[0].shape == (B, b + n + 1, 6) and hidden.shape == (B, b + n + 1, 0)
    assert bool(pad[:, :b].any()) and not bool(pad[:, b:].any())  # only burn-in rows are ever padding here
    noise = sac._noise.cpu()
    sizes = [B * (n + 1) * A, B * A, B * A, B * (n + 1) * A]
    offs = np.cumsum([0] + sizes)
    eps = [noise[offs[i]:offs[i + 1]] for i in range(4)]
    batch = SacBatch(states=obs_list[0].cpu(), actions=actions.cpu(), rewards=rewards.cpu(), dones=dones.cpu(),
                     mu_probs=mu.cpu(), last_masks=last.cpu(), padding_masks=pad.cpu(), priority_is=None)
    cu.step(cu.make_batch(batch, SacNoise(eps_y=eps[0].view(B, n + 1, A), eps_pi=eps[1].view(B, A),
                                          eps_alpha=eps[2].view(B, A), eps_td=eps[3].view(B, n + 1, A))))
    torch.cuda.synchronize()
    assert torch.equal(cu.q, sac._q_flat) and torch.equal(cu.qt, sac._qt_flat) and torch.equal(cu.pi, sac._pi_flat)
    assert torch.equal(cu.log_alpha, sac._log_alpha_buf)
    assert sac._counters.cpu().tolist()[:4] == [1, 1, 1, 1]
    assert sac.train() == 2 and sac.train() == 3
    assert sac.train() == 3  # the incomplete fourth batch waits for more windows
    sac.put_episode(**_episode(rng, [(6,)], A, 12))  # 10 left over + 11 windows -> one more batch
    assert sac.train() == 4
    torch.cuda.synchronize()
    assert all(np.isfinite(v) for v in sac.last_step_stats().values())
    sac.close()
