"""CPU-only: the C-ABI library loads and exports every symbol include/asac_b200.h declares, the
ctypes mirror of its structs has the C layout, the host-side lowering is consistent with the
oracle's parameter naming, and the product path refuses to run without CUDA (no fallback)."""
import ctypes as C
import importlib.util
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / 'include' / 'asac_b200.h'


def _declared_symbols():
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(asac_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from asac_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f'{name} declared in {HEADER.name} but not exported'
        assert name in _lib.PROTOTYPES, f'{name} has no ctypes prototype'
    assert set(_lib.PROTOTYPES) == set(names)
    assert lib.asac_version() >= 100


def test_struct_layout_matches_c(tmp_path):
    """sizeof/offsetof from a C translation unit vs the ctypes mirror."""
    from asac_b200 import _lib
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "asac_b200.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(AsacSacConfig), sizeof(AsacSacParams), '
                   'sizeof(AsacSacBatch), sizeof(AsacSacWork), sizeof(AsacColumnTable), sizeof(AsacColumn), '
                   'sizeof(AsacWriteTable), sizeof(AsacPeerTable), sizeof(AsacGruShape), sizeof(AsacGruNet), '
                   'sizeof(AsacGruRep));\n'
                   'printf("%zu %zu %zu %zu %zu\\n", offsetof(AsacSacConfig, tau), offsetof(AsacSacConfig, gamma_ratio), '
                   'offsetof(AsacSacConfig, lambda_ratio), offsetof(AsacSacWork, y), offsetof(AsacColumnTable, col));\n'
                   'return 0;}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', str(ROOT / 'include'), str(src), '-o', str(exe)])
    sizes, offs = [list(map(int, line.split())) for line in subprocess.check_output([str(exe)]).decode().splitlines()]
    assert sizes == [C.sizeof(_lib.AsacSacConfig), C.sizeof(_lib.AsacSacParams), C.sizeof(_lib.AsacSacBatch),
                     C.sizeof(_lib.AsacSacWork), C.sizeof(_lib.AsacColumnTable), C.sizeof(_lib.AsacColumn),
                     C.sizeof(_lib.AsacWriteTable), C.sizeof(_lib.AsacPeerTable), C.sizeof(_lib.AsacGruShape),
                     C.sizeof(_lib.AsacGruNet), C.sizeof(_lib.AsacGruRep)]
    assert offs == [_lib.AsacSacConfig.tau.offset, _lib.AsacSacConfig.gamma_ratio.offset,
                    _lib.AsacSacConfig.lambda_ratio.offset, _lib.AsacSacWork.y.offset,
                    _lib.AsacColumnTable.col.offset]


def test_host_only_entry_points():
    from asac_b200 import _lib, lowering
    lib = _lib.load()
    for (i, h, d, o) in [(8, 64, 3, 1), (6, 64, 3, 4), (4, 64, 2, 1), (5, 32, 1, 6)]:
        shape = lowering.NetShape(i, h, d, o)
        assert lib.asac_mlp_param_count(i, h, d, o) == shape.count
        assert lib.asac_mlp_param_stride(i, h, d, o) == shape.stride
    assert lowering.NetShape(8, 64, 3, 1).count == 8961
    assert lowering.NetShape(6, 64, 3, 4).count == 9028
    cfg = _lib.AsacSacConfig()
    cfg.batch, cfg.seq_len, cfg.burn_in, cfg.n_step = 256, 2, 0, 1
    cfg.state_size, cfg.action_size, cfg.ensemble = 6, 2, 2
    cfg.q_hidden = cfg.pi_hidden = 64
    cfg.q_depth = cfg.pi_depth = 3
    cfg.bn_stride, cfg.update_target_per_step = 2, 1
    assert lib.asac_sac_tile_batch(C.byref(cfg)) == 4      # 64 batch tiles x E cluster ranks at B = 256
    cfg.batch = 4096
    assert lib.asac_sac_tile_batch(C.byref(cfg)) == 16     # full 16-row tiles once the batch allows
    cfg.batch = 256
    cfg.q_hidden = 48  # not a supported width -> error code + message, no crash
    assert lib.asac_sac_tile_batch(C.byref(cfg)) < 0
    assert b'hidden width' in lib.asac_last_error()
    # long R2D2-style windows shrink the tile instead of overflowing shared memory
    cfg.q_hidden, cfg.burn_in, cfg.n_step, cfg.seq_len, cfg.bn_stride = 64, 40, 5, 46, 46
    cfg.use_n_step_is = 1
    assert 1 <= lib.asac_sac_tile_batch(C.byref(cfg)) < 16
    # BASELINE configs[3] exactly: GRU state of width 8, trained representation (the post pass carries the
    # value rows a second time) -> 4 sequences per tile still fit; the peer buffers grow by the GRU's gradient
    cfg.state_size, cfg.rep_kind = 8, 1
    assert lib.asac_sac_tile_batch(C.byref(cfg)) == 4
    plain = lib.asac_peer_recv_words(C.byref(cfg), 8)
    cfg.rep_param_stride = 864
    assert lib.asac_peer_recv_words(C.byref(cfg), 8) == plain + 2 * 8 * 864


def _load_plugin(tmp_path, text):
    path = Path(text) if str(text).endswith('.py') else tmp_path / 'nn_plugin_cpu.py'
    if not str(text).endswith('.py'):
        path.write_text(text)
    spec = importlib.util.spec_from_file_location('nn_plugin_cpu', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_plugin_file_loads_unchanged_and_matches_oracle_forward(tmp_path):
    """envs/gym/pendulum/nn.py's text, imported through the `algorithm` alias package."""
    from asac_b200 import lowering
    from oracle.sac_oracle import policy_forward, q_forward
    nn = _load_plugin(tmp_path, Path(__file__).resolve().parent / 'golden' / 'plugins' / 'envs_gym_pendulum_nn.py')
    torch.manual_seed(0)
    q = nn.ModelQ(3, [], 1, False, None)
    pi = nn.ModelPolicy(3, [], 1, None)
    q_shape, q_params = lowering.analyze_q(q)
    pi_shape, pi_params = lowering.analyze_policy(pi)
    assert (q_shape.in_dim, q_shape.hidden, q_shape.depth, q_shape.out_dim) == (4, 64, 2, 1)
    assert (pi_shape.in_dim, pi_shape.hidden, pi_shape.depth, pi_shape.out_dim) == (3, 64, 2, 2)
    assert list(q.state_dict()) == ['c_dense.dense.0.linear.weight', 'c_dense.dense.0.linear.bias',
                                    'c_dense.dense.2.linear.weight', 'c_dense.dense.2.linear.bias',
                                    'c_dense.dense.4.weight', 'c_dense.dense.4.bias']
    with torch.no_grad():
        for p in list(q.parameters()) + list(pi.parameters()):
            p.add_(torch.randn_like(p) * 0.1)
    s, a = torch.randn(9, 3), torch.rand(9, 1)
    assert torch.allclose(q(s, a, [s])[1], q_forward(dict(q.state_dict()), 2, s, a), atol=1e-6)
    dist = pi(s, [s])[1]
    loc, scale = policy_forward(dict(pi.state_dict()), 2, s)
    assert torch.allclose(dist.loc, loc, atol=1e-6) and torch.allclose(dist.scale, scale, atol=1e-6)
    # binding: parameters become views of the flat buffer, in the kernel's layout
    flat_q = torch.zeros(q_shape.stride)
    expect = lowering.flat_from_state_dict(q_shape, q.state_dict(), policy=False)
    lowering.bind_parameters(q_params, flat_q)
    assert torch.equal(flat_q, expect)
    flat_pi = torch.zeros(pi_shape.stride)
    expect = lowering.flat_from_state_dict(pi_shape, pi.state_dict(), policy=True)
    lowering.bind_parameters(pi_params, flat_pi)
    assert torch.equal(flat_pi, expect)
    flat_q[0] = 42.
    assert float(q.c_dense.dense[0].linear.weight[0, 0]) == 42.
    back = lowering.state_dict_from_flat(pi_shape, flat_pi, policy=True)
    for k, v in pi.state_dict().items():
        assert torch.equal(back[k], v), k
    assert torch.allclose(q(s, a, [s])[1], q_forward(dict(q.state_dict()), 2, s, a), atol=1e-6)


def test_non_stock_networks_are_rejected(tmp_path):
    from asac_b200 import lowering
    import asac_b200.nn_models as m

    class QWithState(m.ModelQ):
        def _build_model(self):
            super()._build_model(c_state_depth=1)

    class QOverride(m.ModelQ):
        def forward(self, state, c_action, obs_list):
            return super().forward(state[..., :-1], c_action, obs_list)

    with pytest.raises(lowering.NotStockNetwork):
        lowering.analyze_q(QWithState(6, [], 2, False))
    with pytest.raises(lowering.NotStockNetwork):
        lowering.analyze_q(QOverride(6, [], 2, False))
    # discrete heads are lowered next to the continuous net (analyze_d_heads); unequal heads are not stock
    shape, params = lowering.analyze_q(m.ModelQ(6, [3], 2, False))
    assert shape == lowering.NetShape(8, 64, 3, 1) and len(params) == 8
    d_shapes, d_params = lowering.analyze_d_heads(m.ModelQ(6, [3, 4], 2, False), 'ModelQ')
    assert d_shapes == [lowering.NetShape(6, 64, 3, 3), lowering.NetShape(6, 64, 3, 4)] and len(d_params[1]) == 8
    assert lowering.analyze_q(m.ModelQ(6, [3], 0, False)) == (None, [])

    class QDeepHeads(m.ModelQ):
        def _build_model(self):
            super()._build_model(d_dense_depth=0)

    with pytest.raises(lowering.NotStockNetwork):
        lowering.analyze_d_heads(QDeepHeads(6, [3], 2, False), 'ModelQ')


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_path_fails_loudly_without_cuda(tmp_path):
    from asac_b200 import PrioritizedReplayBuffer, SAC_Base, _lib
    with pytest.raises(_lib.AsacError, match='no CPU fallback'):
        PrioritizedReplayBuffer(batch_size=4, capacity=16)
    nn = _load_plugin(tmp_path, 'import algorithm.nn_models as m\nModelRep = m.ModelSimpleRep\n'
                                'ModelQ = m.ModelQ\nModelPolicy = m.ModelPolicy\n')
    with pytest.raises(_lib.AsacError, match='no CPU fallback'):
        SAC_Base(obs_names=['v'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2, model_abs_dir=None, nn=nn)


def test_product_never_imports_the_oracle():
    pkg = ROOT / 'advanced-soft-actor-critic_b200'
    for path in pkg.rglob('*.py'):
        text = path.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), path


def test_reference_checkpoint_fixtures_have_the_expected_layout():
    """CPU-side check of tests/golden/ckpt_*: files and key names as the reference writes them
    (sac_base.py:493-566, 654-668; replay_buffer.py:96-111, 220-227) — what the GPU resume test loads."""
    import numpy as np
    import torch
    from tests.helpers import GOLDEN
    for name, rep in (('vector', False), ('rnn', True)):
        d = GOLDEN / f'ckpt_{name}' / 'model'
        saved = torch.load(d / '4.pth', weights_only=True)
        want = {'global_step', 'model_policy', 'optimizer_policy', 'log_d_alpha', 'log_c_alpha', 'optimizer_alpha'}
        want |= {f'{k}_{i}' for i in range(2) for k in ('model_q', 'model_target_q', 'optimizer_q')}
        if rep:
            want |= {'model_rep', 'model_target_rep', 'optimizer_rep'}
            assert set(saved['model_rep']) == {f'rnn._grus.{l}.{t}_l0' for l in range(2)
                                               for t in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')}
        assert set(saved) == want
        assert int(saved['global_step']) == 4
        tree = np.load(d / '4-rb_tree.npy')
        assert tree.shape == (2 * 64 - 1,) and tree.dtype == np.float32
        store = np.load(d / '4-rb_storage.npz')
        assert {'_id', 'index', 'last_mask', 'obs_vector', 'action', 'reward', 'done', 'mu_prob',
                'pre_seq_hidden_state', 'p_size', 'p_id'} == set(store.files)
        assert store['pre_seq_hidden_state'].shape[1:] == ((2, 8) if rep else (0,))


def test_recurrent_representation_lowering_on_the_host():
    """lowering.analyze_rep / GRU flat layout without a GPU: the envs/test/nn_rnn.py form lowers to a
    GruShape whose flat order is torch.nn.GRU's own, the state_dict round-trips through the flat
    vector, the `m.GRU` wrapper equals the oracle's restatement of the cell (and, when the reference
    is mounted, nothing else), and representations that are not ONE stock GRU are rejected."""
    import numpy as np
    from asac_b200 import _lib, lowering
    import asac_b200.nn_models as m
    from oracle.rep_oracle import gru_forward

    class Rep(m.ModelBaseRep):
        def _build_model(self):
            self.rnn = m.GRU(self.obs_shapes[0][0] + self.c_action_size, 8, 2)

        def forward(self, obs_list, pre_action, pre_seq_hidden_state, padding_mask=None):
            h0 = None if pre_seq_hidden_state is None else pre_seq_hidden_state[:, 0]
            return self.rnn(torch.cat([obs_list[0], pre_action], dim=-1), h0)

    rep = Rep(['v'], [(6,)], [], 2, False)
    shape, params = lowering.analyze_rep(rep, [(6,)], 2)
    assert shape == lowering.GruShape(6, 2, 8, 2) and shape.count == 864 == sum(p.numel() for p in params)
    cs = _lib.AsacGruShape(6, 2, 8, 2)
    assert _lib.load().asac_gru_param_count(C.byref(cs)) == shape.count
    assert _lib.load().asac_gru_backward_tile(C.byref(cs), 40) == 4  # config 4: 4 sequences per CTA fit
    sd = {k: v.detach().clone() for k, v in rep.state_dict().items()}
    flat = lowering.gru_flat_from_state_dict(shape, sd)
    back = lowering.gru_state_dict_from_flat(shape, flat)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    # binding: the nn.GRU parameters become views of the flat buffer, in flat order
    buf = torch.zeros(shape.stride)
    lowering.bind_parameters(params, buf)
    assert torch.equal(buf[:shape.count], flat[:shape.count])
    assert rep.rnn._grus[0].weight_ih_l0.data_ptr() == buf.data_ptr()
    # the wrapper (what the probe and the actor-side cross-check run) == the oracle's cell
    rng = np.random.RandomState(0)
    obs = torch.from_numpy(rng.randn(3, 5, 6).astype(np.float32))
    pre = torch.from_numpy(rng.rand(3, 5, 2).astype(np.float32))
    hid = torch.from_numpy(rng.randn(3, 5, 2, 8).astype(np.float32))
    with torch.no_grad():
        state, hn = rep([obs], pre, hid)
        s2, hn2 = gru_forward(sd, 2, torch.cat([obs, pre], -1), hid[:, 0])
    assert state.shape == (3, 5, 8) and hn.shape == (3, 5, 2, 8)
    assert torch.allclose(state, s2, atol=1e-6) and torch.allclose(hn, hn2, atol=1e-6)
    assert lowering.analyze_rep(m.ModelSimpleRep(['v'], [(6,)], [], 2, False), [(6,)], 2) is None

    class TwoGrus(Rep):
        def _build_model(self):
            super()._build_model()
            self.rnn2 = m.GRU(8, 8, 1)

    class GruAndDense(Rep):
        def _build_model(self):
            super()._build_model()
            self.dense = m.LinearLayers(8, 8, 1)

    class WrongInput(m.ModelBaseRep):
        def _build_model(self):
            self.rnn = m.GRU(6, 8, 1)  # no pre_action column

    for bad in (TwoGrus, GruAndDense, WrongInput):
        with pytest.raises(lowering.NotStockNetwork):
            lowering.analyze_rep(bad(['v'], [(6,)], [], 2, False), [(6,)], 2)
    # the torch wrapper itself supports the padding-mask (packed) form of seq_layers.py:60-103
    out, hn = m.GRU(8, 8, 1)(torch.zeros(1, 2, 8), None, torch.tensor([[True, False]]))
    assert out.shape == (1, 2, 8) and float(out[0, 0].abs().sum()) == 0.0


def test_batch_buffer_windows_without_the_reference():
    """asac_b200.batch_buffer on its own (the comparison with the reference's implementation is
    tests/test_oracle_vs_reference.py): window count, contents, the front-only padding flag, leftovers."""
    import numpy as np
    from asac_b200.batch_buffer import BatchBuffer, episode_to_batch
    T, b, n = 6, 2, 3
    idx = np.arange(T, dtype=np.int32)[None]
    last = np.zeros((1, T), dtype=bool); last[:, -1] = True
    obs = np.arange(T, dtype=np.float32).reshape(1, T, 1) + 100
    act = np.arange(T, dtype=np.float32).reshape(1, T, 1)
    out = episode_to_batch(b, n, np.array([9.], np.float32), idx, last, [obs], act, np.ones((1, T), np.float32),
                           np.zeros((1, T), bool), np.full((1, T, 1), 0.5, np.float32), np.zeros((1, T, 0), np.float32))
    bn_idx, bn_last, bn_pad, (bnx_obs,), bn_act, bn_rew, bn_done, bn_prob, bnx_h = out
    assert bn_idx.shape == (T - 1, b + n) and bnx_obs.shape == (T - 1, b + n + 1, 1) and bnx_h.shape == (T - 1, b + n + 1, 0)
    assert bn_idx[0].tolist() == [-1, -1, 0, 1, 2] and bn_idx[-1].tolist() == [2, 3, 4, 5, -1]
    assert bn_pad[0].tolist() == [True, True, False, False, False] and not bn_pad[2:].any()  # trailing pad: not flagged
    assert bn_last[-1].tolist() == [False, False, False, True, True] and bn_done[-1, -1] and bn_rew[-1, -1] == 0
    assert bn_act[0].ravel().tolist() == [9., 9., 0., 1., 2.] and bn_prob[0].ravel().tolist() == [1., 1., .5, .5, .5]
    assert bnx_obs[0].ravel().tolist() == [0., 0., 100., 101., 102., 103.]
    buf = BatchBuffer(b, n, np.array([9.], np.float32), batch_size=4)
    np.random.seed(0)
    buf.put_episode(idx, last, [obs], act, np.ones((1, T), np.float32), np.zeros((1, T), bool),
                    np.full((1, T, 1), 0.5, np.float32), np.zeros((1, T, 0), np.float32))
    first = buf.get_batch()
    assert first[0].shape == (4, b + n) and buf.get_batch() is None and buf._rest_batch[0].shape[0] == 1
    seen = sorted(int(r[2]) for r in first[0].tolist()) + [int(buf._rest_batch[0][0, 2])]
    assert sorted(seen) == [0, 1, 2, 3, 4]  # every window exactly once


def test_flat_adam_state_dict_loads_into_torch_adam():
    """The optimizer entries of a checkpoint written here must load into the reference's
    ``torch.optim.Adam`` objects (sac_base.py:296-300, 609-624): ``_FlatAdam.state_dict()`` (views of the flat
    moment buffers the CUDA Adam kernel updates) goes through ``Adam.load_state_dict`` and the next torch
    step continues from those moments; the reverse direction restores moments and step count."""
    from asac_b200.sac_base import _FlatAdam
    torch.manual_seed(0)
    shapes = [(4, 3), (4,), (2, 4), (2,)]
    n = sum(int(np.prod(s)) for s in shapes)
    flat, m_flat, v_flat = torch.randn(n), torch.rand(n) * 0.1, torch.rand(n) * 0.01
    counters = torch.tensor([9, 7, 0, 0, 0, 0, 0, 0], dtype=torch.int64)
    params, off = [], 0
    for s in shapes:
        k = int(np.prod(s))
        params.append(torch.nn.Parameter(flat[off:off + k].view(s)))
        off += k
    ours = _FlatAdam(params, flat, m_flat, v_flat, counters, 1, 3e-4)
    sd = ours.state_dict()
    assert set(sd) == {'state', 'param_groups'} and float(sd['state'][0]['step']) == 7
    clones = [torch.nn.Parameter(p.detach().clone()) for p in params]
    ref = torch.optim.Adam(clones, lr=3e-4)
    ref.load_state_dict(sd)
    assert torch.equal(ref.state[clones[2]]['exp_avg'], m_flat[16:24].view(2, 4))
    for c in clones:
        c.grad = torch.ones_like(c)
    ref.step()
    assert float(ref.state[clones[0]]['step']) == 8
    # and back: a state dict written by torch.optim.Adam restores the flat moments and the step counter
    m_flat.zero_(); v_flat.zero_(); counters[1] = 0
    ours.load_state_dict(ref.state_dict())
    assert int(counters[1]) == 8
    assert torch.equal(m_flat[16:24].view(2, 4), ref.state[clones[2]]['exp_avg'])
    assert torch.equal(v_flat[:12].view(4, 3), ref.state[clones[0]]['exp_avg_sq'])


def test_every_environment_switch_is_documented():
    """DESIGN.md §11 is the one list of A/B and debugging switches: every ASAC_* variable the product reads
    (library sources, host package, bench.py) must appear there."""
    import re
    root = Path(__file__).resolve().parent.parent
    pkg = root / 'advanced-soft-actor-critic_b200' / 'asac_b200'
    files = list((pkg / 'csrc').glob('*.cu')) + list((pkg / 'csrc').glob('*.cuh')) + list(pkg.glob('*.py')) + [root / 'bench.py']
    used = set()
    for f in files:
        text = f.read_text()
        used |= set(re.findall(r'getenv\("(ASAC_[A-Z0-9_]+)"\)', text))
        used |= set(re.findall(r"environ(?:\.get)?[\(\[]'(ASAC_[A-Z0-9_]+)'", text))
    assert len(used) >= 15, used
    design = (root / 'DESIGN.md').read_text()
    section = design[design.index('## 11. Switches'):]
    missing = sorted(v for v in used if v not in section)
    assert not missing, f'switches not documented in DESIGN.md §11: {missing}'
