"""Live differential check of the CPU oracle against the MOUNTED reference (build container only;
skipped where /root/reference does not exist, e.g. on the GPU box): fixtures are minted on the spot
by oracle/gen_golden.py with seeds and shapes that differ from the committed ones and go through the
same comparisons as tests/test_oracle_golden.py.  Guards against an oracle that merely fits its own
committed fixtures.  CPU only."""
import numpy as np
import pytest

from oracle.ref_shims import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason='the reference checkout is not mounted')


@pytest.fixture()
def mint(tmp_path, monkeypatch):
    """gen_golden with its output directory redirected; the reference's `algorithm` package shadows the
    product's alias package of the same name only for the duration of the test."""
    import sys
    import oracle.gen_golden as gen
    from oracle.ref_shims import REFERENCE_ROOT
    monkeypatch.setattr(gen, 'GOLDEN', tmp_path)
    saved_path = list(sys.path)
    saved_modules = {n: m for n, m in sys.modules.items() if n == 'algorithm' or n.startswith('algorithm.')}
    for n in saved_modules:
        del sys.modules[n]
    sys.path.insert(0, str(REFERENCE_ROOT))

    def load(name):
        with np.load(tmp_path / name) as z:
            return {k: z[k] for k in z.files}
    yield gen, load
    for n in [n for n in sys.modules if n == 'algorithm' or n.startswith('algorithm.')]:
        del sys.modules[n]
    sys.modules.update(saved_modules)
    sys.path[:] = saved_path


def test_replay_trace_fresh_seed(mint):
    from tests.oracle_checks import check_per_trace
    gen, load = mint
    gen.gen_per_case('live', capacity=32, batch_size=4, prev_n=1, post_n=2, alpha=0.8,
                     episode_lens=[7, 11, 20, 6], n_rounds=3, seed=101)
    check_per_trace(load('per_live.npz'))


def test_sac_step_fresh_seed_and_shape(mint):
    from tests.oracle_checks import check_sac_steps
    gen, load = mint
    gen.gen_sac_case('live', S=7, A=3, E=2, hidden=32, depth=2, B=9, b=1, n=2, steps=2, seed=102,
                     v_lambda=0.9, v_rho=0.8, tau=0.01)
    check_sac_steps(load('sac_live.npz'))


def test_recurrent_sac_step_fresh_seed(mint):
    from tests.oracle_checks import check_recurrent_sac_steps
    gen, load = mint
    gen.gen_sac_rnn_case('live', So=6, A=2, E=2, B=6, b=4, n=2, steps=2, seed=103)
    check_recurrent_sac_steps(load('sac_live.npz'))


def test_discrete_and_hybrid_step_fresh_seed(mint):
    from tests.oracle_checks import check_hybrid_sac_steps
    gen, load = mint
    gen.gen_sac_discrete_case('live_d', S=5, d_action_sizes=[2, 5, 3], A=0, E=3, B=7, b=1, n=2, steps=2, seed=104,
                              v_rho=0.9, v_c=0.8)
    check_hybrid_sac_steps(load('sac_live_d.npz'))
    gen.gen_sac_discrete_case('live_h', S=5, d_action_sizes=[4], A=3, E=2, B=7, b=0, n=1, steps=1, seed=105,
                              use_n_step_is=False)
    check_hybrid_sac_steps(load('sac_live_h.npz'))
    gen.gen_sac_discrete_case('live_q', S=5, d_action_sizes=[3, 3], A=0, E=3, B=9, b=1, n=4, steps=2, seed=107,
                              discrete_dqn_like=True)
    check_hybrid_sac_steps(load('sac_live_q.npz'))


def test_batch_buffer_equals_reference(mint):
    """asac_b200.batch_buffer (on-policy path, use_replay_buffer=False) against the reference's
    episode_to_batch / BatchBuffer (utils/operators.py:105-207, batch_buffer.py:10-95) on the same episodes
    and the same NumPy seed: identical windows, padding quirk included, identical batches in identical order."""
    import torch
    from oracle.ref_shims import import_reference
    import_reference()
    from algorithm.batch_buffer import BatchBuffer as RefBuffer
    from algorithm.utils import episode_to_batch as ref_e2b
    from asac_b200.batch_buffer import BatchBuffer, episode_to_batch

    rng = np.random.RandomState(7)

    def episode(T, hidden=(2, 4)):
        idx = np.arange(T, dtype=np.int32)[None]
        last = np.zeros((1, T), dtype=bool); last[:, -1] = True
        return dict(ep_indexes=idx, ep_last_masks=last,
                    ep_obses_list=[rng.randn(1, T, 5).astype(np.float32), rng.randn(1, T, 2, 3).astype(np.float32)],
                    ep_actions=rng.rand(1, T, 3).astype(np.float32), ep_rewards=rng.randn(1, T).astype(np.float32),
                    ep_dones=rng.rand(1, T) < 0.2, ep_probs=rng.rand(1, T, 3).astype(np.float32),
                    ep_pre_seq_hidden_states=rng.randn(1, T, *hidden).astype(np.float32))

    pad_action = np.array([0.5, -0.5, 0.25], dtype=np.float32)
    for b, n, T in ((0, 1, 6), (3, 2, 9), (2, 4, 5)):
        ep = episode(T)
        args = dict(l_indexes=ep['ep_indexes'], l_last_masks=ep['ep_last_masks'], l_actions=ep['ep_actions'],
                    l_rewards=ep['ep_rewards'], l_dones=ep['ep_dones'], l_probs=ep['ep_probs'],
                    l_pre_seq_hidden_states=ep['ep_pre_seq_hidden_states'])
        want = ref_e2b(burn_in_step=b, n_step=n, padding_action=pad_action,
                       l_obses_list=[o.copy() for o in ep['ep_obses_list']], **args)
        got = episode_to_batch(b, n, pad_action, l_obses_list=ep['ep_obses_list'], **args)
        for i, (w, g) in enumerate(zip(want, got)):
            if isinstance(w, list):
                assert all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(w, g)), (b, n, i)
            else:
                assert np.array_equal(w, g) and w.dtype == g.dtype and w.shape == g.shape, (b, n, i)
    ours = BatchBuffer(2, 3, pad_action, batch_size=8, max_size=3)
    ref = RefBuffer(2, 3, pad_action, batch_size=8, max_size=3)
    eps = [episode(T) for T in (7, 12, 4, 30, 9)]
    np.random.seed(11)
    for ep in eps:
        ours.put_episode(**{k: ([o.copy() for o in v] if isinstance(v, list) else v) for k, v in ep.items()})
    np.random.seed(11)
    for ep in eps:
        ref.put_episode(**{k: ([o.copy() for o in v] if isinstance(v, list) else v) for k, v in ep.items()})
    n_batches = 0
    while True:
        a, r = ours.get_batch(), ref.get_batch()
        assert (a is None) == (r is None)
        if a is None:
            break
        n_batches += 1
        for x, y in zip(a, r):
            if isinstance(x, list):
                assert all(torch.equal(p, q) for p, q in zip(x, y))
            else:
                assert torch.equal(x, y) and x.shape[0] == 8
    assert n_batches == 3  # max_size keeps the newest three of the seven full batches


def test_padding_fresh_seed(mint):
    from tests.oracle_checks import check_padding
    gen, load = mint
    gen.gen_padding_case('live', burn_in_step=3, n_step=2, batch_size=12, capacity=64, seed=106)
    check_padding(load('pad_live.npz'))
