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
