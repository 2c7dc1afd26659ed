"""The drop-in boundary of the plugin ``nn`` surface, CPU only.

* every ``envs/**/nn*.py`` and ``tests/nn_*.py`` of the reference imports UNCHANGED through the alias
  package ``algorithm`` (needs the reference checkout: build container only);
* the verbatim plugin files committed under ``tests/golden/plugins`` are the reference's text and build;
* every restated layer / model class has the reference's ``state_dict`` keys and — after loading the
  reference module's weights — the reference's outputs (needs the reference checkout).
"""
import importlib.util
import sys
import types
from pathlib import Path

import pytest
import torch

from oracle.ref_shims import REFERENCE_ROOT, reference_available

PLUGINS = Path(__file__).resolve().parent / 'golden' / 'plugins'
needs_reference = pytest.mark.skipif(not reference_available(), reason='the reference checkout is not mounted')


def load_plugin(path: Path, name: str, package_root: Path | None = None):
    """Executes a plugin file the way sac_main.py:353-364 does; ``package_root`` gives relative imports
    (``from .nn_parking import *``) a package to resolve against."""
    if package_root is not None:
        parts = path.relative_to(package_root).with_suffix('').parts
        for i in range(len(parts)):
            pkg = '.'.join((name,) + parts[:i])
            if pkg not in sys.modules:
                mod = types.ModuleType(pkg)
                mod.__path__ = [str(package_root.joinpath(*parts[:i]))]
                sys.modules[pkg] = mod
        name = '.'.join((name,) + parts)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@needs_reference
def test_every_reference_plugin_file_imports_unchanged():
    import algorithm  # noqa: F401  (the alias package)
    files = sorted(list(REFERENCE_ROOT.glob('envs/**/nn*.py')) + list(REFERENCE_ROOT.glob('tests/nn_*.py')))
    assert len(files) >= 56
    failed = []
    for f in files:
        try:
            mod = load_plugin(f, 'refplug', REFERENCE_ROOT)
            assert hasattr(mod, 'ModelRep') or hasattr(mod, 'ModelOptionRep') or f.name.startswith('nn_'), f
        except Exception as e:  # noqa: BLE001
            failed.append(f'{f.relative_to(REFERENCE_ROOT)}: {type(e).__name__}: {e}')
    assert not failed, '\n'.join(failed)


@needs_reference
@pytest.mark.parametrize('fixture,ref', [('envs_test_nn_rnn.py', 'envs/test/nn_rnn.py'),
                                         ('envs_test_nn.py', 'envs/test/nn.py'),
                                         ('envs_gym_pendulum_nn.py', 'envs/gym/pendulum/nn.py'),
                                         ('tests_nn_conv_attn.py', 'tests/nn_conv_attn.py')])
def test_committed_plugin_files_are_verbatim(fixture, ref):
    assert (PLUGINS / fixture).read_bytes() == (REFERENCE_ROOT / ref).read_bytes()


def test_committed_plugin_files_build_on_the_alias():
    """No reference needed: the committed verbatim files construct their models (CPU torch modules)."""
    import asac_b200.nn_models as m
    rnn = load_plugin(PLUGINS / 'envs_test_nn_rnn.py', 'plug_nn_rnn')
    rep = rnn.ModelRep(['vector'], [(6,)], [], 2, False)
    assert isinstance(rep.rnn, m.GRU)
    state, hn = rep([torch.randn(3, 5, 6)], torch.rand(3, 5, 2), torch.zeros(3, 5, 2, 8))
    assert state.shape == (3, 5, 8) and hn.shape == (3, 5, 2, 8)
    assert rnn.ModelQ is m.ModelQ and rnn.ModelTermination is m.ModelTermination

    attn = load_plugin(PLUGINS / 'tests_nn_conv_attn.py', 'plug_nn_conv_attn')
    rep = attn.ModelRep(['vector', 'image'], [(10,), (3, 30, 30)], [], 2, False)
    idx = torch.arange(6).expand(2, 6)
    state, hn, weights = rep(4, idx, [torch.randn(2, 6, 10), torch.rand(2, 6, 3, 30, 30)], torch.rand(2, 6, 2), None)
    assert state.shape == (2, 4, 8) and hn.shape == (2, 4, 8) and len(weights) == 2

    pend = load_plugin(PLUGINS / 'envs_gym_pendulum_nn.py', 'plug_pendulum')
    q = pend.ModelQ(3, [], 1, False)
    assert sum(isinstance(b, m.ResBlock) for b in q.c_dense.dense) == 2


# ------------------------------------------------------------------ differential checks against the reference
@pytest.fixture(scope='module')
def ref_m():
    """The reference's own ``algorithm.nn_models`` (its `algorithm` shadows the alias only while importing)."""
    if not reference_available():
        pytest.skip('the reference checkout is not mounted')
    from oracle.ref_shims import install_shims
    install_shims()
    saved_path = list(sys.path)
    saved = {n: mod for n, mod in sys.modules.items() if n == 'algorithm' or n.startswith('algorithm.')}
    for n in saved:
        del sys.modules[n]
    sys.path[:] = [str(REFERENCE_ROOT)] + [p for p in saved_path
                                            if not (Path(p or '.') / 'algorithm' / '__init__.py').exists()]
    try:
        import algorithm.nn_models as ref
        import algorithm.nn_models.layers.seq_layers as ref_seq
        import algorithm.utils.operators as ref_ops
        import algorithm.utils.transform as ref_tf
        ref.seq, ref.ops, ref.tf = ref_seq, ref_ops, ref_tf
    finally:
        for n in [n for n in sys.modules if n == 'algorithm' or n.startswith('algorithm.')]:
            del sys.modules[n]
        sys.modules.update(saved)
        sys.path[:] = saved_path
    assert str(REFERENCE_ROOT) in ref.__file__
    return ref


def _pair(ref_cls, my_cls, *args, **kwargs):
    """Both modules built under the same torch seed; weights of the reference loaded into ours (strict)."""
    torch.manual_seed(11)
    a = ref_cls(*args, **kwargs)
    torch.manual_seed(11)
    b = my_cls(*args, **kwargs)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape, k
        assert torch.allclose(sa[k], sb[k], atol=1e-6), f'{k}: initialisation differs under the same seed'
    b.load_state_dict(sa, strict=True)
    return a.eval(), b.eval()


def _same(x, y, tol=1e-6):
    if isinstance(x, (tuple, list)):
        assert len(x) == len(y)
        for u, v in zip(x, y):
            _same(u, v, tol)
        return
    assert x.shape == y.shape
    assert torch.allclose(x, y, atol=tol, rtol=0), float((x - y).abs().max())


def _enum(mine, ref_member):
    return None if ref_member is None else mine[ref_member.name]


def test_public_names_cover_the_reference(ref_m):
    import asac_b200.nn_models as m
    missing = [n for n in dir(ref_m) if not n.startswith('_') and isinstance(getattr(ref_m, n), type)
               and getattr(ref_m, n).__module__.startswith('algorithm.') and not hasattr(m, n)]
    assert not missing, missing


@pytest.mark.parametrize('args,kw', [((6, 64, 3, 1), {}), ((5, [16, 16, 8], 0, None), {}), ((7, 32, 0, None), {}),
                                     ((8, 8, 2, 4), dict(residual=False))])
def test_linear_layers(ref_m, args, kw):
    import asac_b200.nn_models as m
    a, b = _pair(ref_m.LinearLayers, m.LinearLayers, *args, **kw)
    x = torch.randn(4, 3, args[0])
    _same(a(x), b(x))
    assert a.output_size == b.output_size


@pytest.mark.parametrize('masked', [False, True])
def test_gru_wrapper_with_and_without_padding(ref_m, masked):
    import asac_b200.nn_models as m
    a, b = _pair(ref_m.GRU, m.GRU, 5, 8, 2)
    x, h0 = torch.randn(5, 9, 5), torch.randn(5, 2, 8)
    mask = None
    if masked:  # left padding, right padding, both, none, everything
        mask = torch.zeros(5, 9, dtype=torch.bool)
        mask[0, :3] = True
        mask[1, -2:] = True
        mask[2, :4] = True; mask[2, -3:] = True
        mask[4, :] = True
    for h in (h0, None):
        _same(a(x, h, mask), b(x, h, mask))


@pytest.mark.parametrize('pe', [None, 'ABSOLUTE', 'ABSOLUTE_CAT', 'ROPE', 'ROPE2'])
@pytest.mark.parametrize('heads', [1, 2])
def test_multihead_attention(ref_m, pe, heads):
    import asac_b200.nn_models as m
    kw = dict(num_heads=heads, pe=None if pe is None else ref_m.seq.POSITIONAL_ENCODING[pe], qkv_dense_depth=1,
              out_dense_depth=1)
    torch.manual_seed(3)
    a = ref_m.MultiheadAttention(8, **kw)
    kw['pe'] = None if pe is None else m.POSITIONAL_ENCODING[pe]
    b = m.MultiheadAttention(8, **kw)
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())
    a.eval(); b.eval()
    q, k = torch.randn(3, 4, 8), torch.randn(3, 6, 8)
    qi, ki = torch.randint(0, 20, (3, 4)), torch.randint(0, 20, (3, 6))
    pad = torch.zeros(3, 6, dtype=torch.bool)
    pad[0, :2] = True
    pad[2, :] = True                       # a batch row with nothing to attend to
    causal = torch.ones(4, 6, dtype=torch.bool).triu(diagonal=3)
    per_row = torch.rand(3, 4, 6) > 0.6
    for mask in (None, causal, per_row):
        for p in (None, pad):
            _same(a(q, k, k, qi, ki, p, None if mask is None else mask.clone()),
                  b(q, k, k, qi, ki, p, None if mask is None else mask.clone()), 2e-6)
    _same(a(q, k, k), b(q, k, k), 2e-6)


@pytest.mark.parametrize('gate', [None, 'RESIDUAL', 'OUTPUT', 'RECURRENT', 'CAT'])
@pytest.mark.parametrize('layers', [1, 2, 3])
def test_episode_attention_stack_all_modes(ref_m, gate, layers):
    import asac_b200.nn_models as m
    norm = gate in ('RESIDUAL', 'CAT')
    pe = 'ROPE' if gate == 'OUTPUT' else None
    torch.manual_seed(5)
    a = ref_m.EpisodeMultiheadAttention(8, num_layers=layers, num_heads=2,
                                        pe=None if pe is None else ref_m.seq.POSITIONAL_ENCODING[pe],
                                        gate=None if gate is None else ref_m.seq.GATE[gate], use_layer_norm=norm)
    b = m.EpisodeMultiheadAttention(8, num_layers=layers, num_heads=2,
                                    pe=None if pe is None else m.POSITIONAL_ENCODING[pe],
                                    gate=None if gate is None else m.GATE[gate], use_layer_norm=norm)
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())
    a.eval(); b.eval()
    assert a.output_dim == b.output_dim and a.output_hidden_state_dim == b.output_hidden_state_dim
    B, L, q = 3, 7, 4
    key = torch.randn(B, L, 8)
    index = torch.arange(L).expand(B, L).clone()
    pad = torch.zeros(B, L, dtype=torch.bool)
    pad[1, :2] = True
    index[1, :2] = -1
    for cut in (True, False):
        for only_rest in (False, True):
            kw = dict(seq_q_len=q, cut_query=cut, query_only_attend_to_rest_key=only_rest, key_index=index,
                      key_padding_mask=pad)
            out_a = a(key, **kw)
            _same(out_a, b(key, **kw), 5e-6)
            hid = torch.randn(B, 2, a.output_hidden_state_dim)
            for prev in (False, True):
                _same(a(key, hidden_state=hid, is_prev_hidden_state=prev, **kw),
                      b(key, hidden_state=hid, is_prev_hidden_state=prev, **kw), 5e-6)
    _same(a(key), b(key), 5e-6)


@pytest.mark.parametrize('conv', ['small', 'simple', 'nature'])
def test_conv_layers(ref_m, conv):
    import asac_b200.nn_models as m
    a, b = _pair(ref_m.ConvLayers, m.ConvLayers, 84, 84, 3, conv, out_dense_depth=2, output_size=8)
    x = torch.rand(2, 3, 3, 84, 84)
    _same(a(x), b(x), 1e-5)
    assert a.conv_output_size == b.conv_output_size


def test_conv1d_transpose_transform_vit(ref_m):
    import asac_b200.nn_models as m
    a, b = _pair(ref_m.Conv1dLayers, m.Conv1dLayers, 100, 2, 'default', 64, 1, 16)
    x = torch.rand(2, 3, 100, 2)
    _same(a(x), b(x), 1e-5)
    up = lambda: torch.nn.ConvTranspose2d(4, 3, 4, 2)
    torch.manual_seed(2); a = ref_m.ConvTransposeLayers(8, 32, 1, 5, 5, 4, up())
    torch.manual_seed(2); b = m.ConvTransposeLayers(8, 32, 1, 5, 5, 4, up())
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 3, 8)
    _same(a(x), b(x), 1e-5)
    flip = lambda t: t.flip(-1)
    x = torch.rand(2, 3, 3, 8, 8)
    _same(ref_m.Transform(flip)(x), m.Transform(flip)(x))
    _same(ref_m.Transform()(x), m.Transform()(x))
    a, b = _pair(ref_m.VisionTransformer, m.VisionTransformer, 16, 3, 4, 1, 2, 8, 16)
    x = torch.rand(2, 2, 3, 16, 16)
    _same(a(x), b(x), 1e-5)
    for fn in ('conv1d_output_size', 'conv2d_output_shape', 'pool_out_shape', 'convtranspose_output_shape'):
        arg = 37 if fn == 'conv1d_output_size' else (37, 52)
        assert getattr(ref_m, fn)(arg, 5, 2) == getattr(m, fn)(arg, 5, 2)


def test_policy_q_and_auxiliary_heads(ref_m):
    import asac_b200.nn_models as m
    # hybrid policy / critic: discrete branches + continuous head
    a, b = _pair(ref_m.ModelPolicy, m.ModelPolicy, 6, [3, 2], 2)
    s = torch.randn(5, 6)
    (da, ca), (db, cb) = a(s, []), b(s, [])
    _same([da.logits, da.probs, ca.mean, ca.stddev], [db.logits, db.probs, cb.mean, cb.stddev])
    onehot = da.sample_deter()
    _same(onehot, db.sample_deter())
    _same([da.log_prob(onehot), da.entropy()], [db.log_prob(onehot), db.entropy()])
    a, b = _pair(ref_m.ModelQ, m.ModelQ, 6, [3, 2], 2, False)
    _same(a(s, torch.full((5, 2), 0.3), []), b(s, torch.full((5, 2), 0.3), []))
    a, b = _pair(ref_m.ModelTermination, m.ModelTermination, 6)
    _same(a(s, []), b(s, []))
    a, b = _pair(ref_m.ModelRND, m.ModelRND, 6, 5, 2)
    _same([a.cal_s_rnd(s), a.cal_d_rnd(s), a.cal_c_rnd(s, s[:, :2])], [b.cal_s_rnd(s), b.cal_d_rnd(s), b.cal_c_rnd(s, s[:, :2])])
    a, b = _pair(ref_m.ModelOptionSelectorRND, m.ModelOptionSelectorRND, 6, 3)
    _same(a.cal_rnd(s), b.cal_rnd(s))
    a, b = _pair(ref_m.ModelForwardDynamic, m.ModelForwardDynamic, 6, 2)
    _same(a(s, s[:, :2]), b(s, s[:, :2]))
    a, b = _pair(ref_m.ModelInverseDynamic, m.ModelInverseDynamic, 6, 2)
    _same(a(s, s), b(s, s))
    a, b = _pair(ref_m.ModelTransition, m.ModelTransition, 6, 0, 2, False)
    da, db = a([], s, s[:, :2]), b([], s, s[:, :2])
    _same([da.mean, da.stddev], [db.mean, db.stddev])
    a, b = _pair(ref_m.ModelReward, m.ModelReward, 6)
    _same(a(s), b(s))
    a, b = _pair(ref_m.ModelVOverOptions, m.ModelVOverOptions, 6, 4, False)
    _same(a(s), b(s))
    a, b = _pair(ref_m.ModelRepProjection, m.ModelRepProjection, 6)
    _same(a(s), b(s))
    a, b = _pair(ref_m.ModelRepPrediction, m.ModelRepPrediction, 6)
    _same(a(s), b(s))
    # NormalWithPadding
    pad = torch.tensor([[False, True]] * 4)
    loc, scale = torch.randn(4, 2), torch.rand(4, 2) + 0.1
    na, nb = ref_m.policy.NormalWithPadding(loc, scale, pad), m.NormalWithPadding(loc, scale, pad)
    v = torch.randn(4, 2)
    _same([na.log_prob(v), na.entropy()], [nb.log_prob(v), nb.entropy()])
    torch.manual_seed(1); ra = na.rsample()
    torch.manual_seed(1); rb = nb.rsample()
    _same(ra, rb)
    torch.manual_seed(1); sa = na.sample()
    torch.manual_seed(1); sb = nb.sample()
    _same(sa, sb)


def test_operators_and_enums(ref_m):
    from asac_b200.utils import operators as ops
    from asac_b200.utils.enums import convert_config_to_enum, convert_config_to_string
    dist = torch.distributions.Normal(torch.randn(4, 3), torch.rand(4, 3) + 0.2)
    x = torch.randn(4, 3) * 2
    _same(ref_m.ops.squash_correction_log_prob(dist, x), ops.squash_correction_log_prob(dist, x))
    _same(ref_m.ops.squash_correction_prob(dist, x), ops.squash_correction_prob(dist, x))
    lp = torch.randn(4, 3); lp[1, 2] = torch.inf
    _same(ref_m.ops.sum_log_prob(lp.clone(), keepdim=True), ops.sum_log_prob(lp.clone(), keepdim=True))
    _same(ref_m.ops.sum_entropy(lp.clone()), ops.sum_entropy(lp.clone()))
    pr = torch.rand(4, 3); pr[0, 0] = torch.inf; pr[2] = torch.tensor([torch.nan, 1., 1.])
    _same(ref_m.ops.prod_prob(pr.clone()), ops.prod_prob(pr.clone()))
    acts = torch.rand(3, 5, 2)
    for keep in (False, True):
        _same(ref_m.ops.gen_n_pre_actions(acts, keep), ops.gen_n_pre_actions(acts, keep))
        assert (ref_m.ops.gen_n_pre_actions(acts.numpy(), keep) == ops.gen_n_pre_actions(acts.numpy(), keep)).all()
    _same(ref_m.ops.gen_n_pre_actions(acts[:, :0], True), ops.gen_n_pre_actions(acts[:, :0], True))
    cfg = {'seq_encoder': 'ATTN', 'siamese': 'BYOL', 'curiosity': None, 'option_seq_encoder': 'RNN'}
    convert_config_to_enum(cfg)
    assert cfg['seq_encoder'].name == 'ATTN' and cfg['siamese'].name == 'BYOL' and cfg['curiosity'] is None
    convert_config_to_string(cfg)
    assert cfg == {'seq_encoder': 'ATTN', 'siamese': 'BYOL', 'curiosity': None, 'option_seq_encoder': 'RNN'}


def test_transforms(ref_m):
    from asac_b200.utils import transform as tf
    img = torch.rand(2, 3, 8, 8)
    for name, args in (('GaussianNoise', (0.1, 0.2)), ('SaltAndPepperNoise', (0.3, 0.7)), ('DepthNoise', (0.2,)),
                       ('DepthSaltAndPepperNoise', (1.0, 0.3))):
        torch.manual_seed(4); want = getattr(ref_m.tf, name)(*args)(img.clone())
        torch.manual_seed(4); got = getattr(tf, name)(*args)(img.clone())
        _same(want, got)
