"""GPU parity: the HBM segment tree / storage kernels and the PrioritizedReplayBuffer host class
against the golden traces minted from the reference and against the NumPy oracle.
Integer / index results and fp32 tree nodes are compared bit-exactly."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu

EP_KEYS = ['index', 'last_mask', 'obs_vector', 'obs_image', 'action', 'reward', 'done', 'mu_prob',
           'pre_seq_hidden_state']


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize('name', ['per_small.npz', 'per_zeros.npz'])
def test_replay_trace_bit_exact(name):
    """Same trace as tests/test_oracle_golden.py::test_per_trace_bit_exact, on the GPU buffer."""
    from asac_b200 import PrioritizedReplayBuffer
    g = load_golden(name)
    capacity, batch_size, prev_n, post_n, n_rounds, n_eps = [int(x) for x in g['meta']]
    rb = PrioritizedReplayBuffer(batch_size=batch_size, sample_prev_n=prev_n, sample_post_n=post_n,
                                 device='cuda:0', capacity=capacity, alpha=float(g['alpha']))
    for e in range(n_eps):
        rb.add({k: g[f'add{e}.{k}'] for k in EP_KEYS}, ignore_size=1)
        assert np.array_equal(_np(rb.tree_nodes()), g[f'add{e}.tree']), f'tree after add {e}'
        assert np.array_equal(_np(rb._store_ids), g[f'add{e}.ids'])
    lib = rb._lib
    for r in range(n_rounds):
        if r == 0 and 'zeroed.idx' in g:
            idx = torch.from_numpy(g['zeroed.idx']).cuda()
            zeros = torch.zeros(len(idx), device='cuda')
            assert lib.asac_tree_update(rb._nodes.data_ptr(), capacity, idx.data_ptr(), zeros.data_ptr(), len(idx),
                                        torch.cuda.current_stream().cuda_stream) == 0
            assert np.array_equal(_np(rb.tree_nodes()), g['zeroed.tree'])
        assert rb.beta == pytest.approx(float(g[f'r{r}.beta_before']))
        data_ids, batch, weights = rb.sample(unit_uniform=g[f'r{r}.u'])
        assert np.array_equal(_np(data_ids), g[f'r{r}.data_ids'])
        w, w_ref = _np(weights)[:, 0], g[f'r{r}.is_weights']
        assert np.max(np.abs(w.view(np.int32) - w_ref.view(np.int32))) <= 1, 'IS weights differ by more than 1 ulp'
        for k in EP_KEYS:
            assert np.array_equal(_np(batch[k]), g[f'r{r}.batch.{k}']), k
        # priority update.  NumPy evaluates clip(td)^alpha with a SIMD fp32 powf whose results are not
        # correctly rounded; the kernel evaluates in float64 and rounds once.  Priorities therefore
        # agree to an fp32 ulp or two (north star: priorities within 1e-5), not bit-for-bit:
        # (1) with the reference's own priorities injected the whole tree is bit-exact,
        # (2) with td errors the leaves agree within 4 ulp and every parent is exactly fp32(l + r).
        ref = g[f'r{r}.tree_after_update']
        saved = rb._nodes.clone()
        upd = g[f'r{r}.upd_ids'].astype(np.int64)
        p_ref = torch.from_numpy(ref[capacity - 1:][upd % capacity].copy()).cuda()
        t_ids = torch.from_numpy(upd).cuda()
        assert lib.asac_per_update(rb._nodes.data_ptr(), capacity, rb._store_ids.data_ptr(), t_ids.data_ptr(),
                                   p_ref.data_ptr(), len(upd), 0.01, 1.0, float(g['alpha']), 1,
                                   rb._per_state.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
        tree = _np(rb.tree_nodes())
        assert np.array_equal(tree, ref), f'round {r}: {np.sum(tree != ref)} nodes differ (injected priorities)'
        rb._nodes.copy_(saved)
        rb.update(g[f'r{r}.upd_ids'], g[f'r{r}.td'])
        tree = _np(rb.tree_nodes())
        leaves, leaves_ref = tree[capacity - 1:], ref[capacity - 1:]
        ulps = np.abs(leaves.view(np.int32).astype(np.int64) - leaves_ref.view(np.int32).astype(np.int64))
        assert ulps.max() <= 4, f'round {r}: priorities differ by {ulps.max()} ulp'
        assert np.max(np.abs(leaves - leaves_ref)) <= 1e-5
        heap = np.concatenate([[0.], tree]).astype(np.float32)  # 1-based
        parents = np.arange(1, capacity)
        assert np.array_equal(heap[parents], heap[2 * parents] + heap[2 * parents + 1]), 'parent != fp32(l + r)'
        rb._nodes[1:].copy_(torch.from_numpy(ref).cuda())  # continue from the reference's exact state
        rb.update_transitions(g[f'r{r}.upd_ids'], 'mu_prob', g[f'r{r}.new_mu'])
        assert np.array_equal(_np(rb._columns['mu_prob']), g[f'r{r}.mu_after'])
        rb.add({k: g[f'r{r}.ep.{k}'] for k in EP_KEYS}, ignore_size=1)
        assert np.array_equal(_np(rb.tree_nodes()), g[f'r{r}.tree_after_add'])
    assert np.array_equal(_np(rb._store_ids), g['final.ids'])
    assert rb.size == int(g['final.size'])
    assert rb._next_id == int(g['final.next_id'])
    rb.check_nan()
    rb.close()


def test_tree_full_capacity_against_oracle():
    """BASELINE capacity (524288): random updates with duplicates, rebuild, max, 4096 samples."""
    from asac_b200 import _lib
    from oracle.replay_oracle import SumTreeOracle
    lib = _lib.load()
    C = 524288
    rng = np.random.RandomState(0)
    s = torch.cuda.current_stream().cuda_stream
    nodes = torch.zeros(2 * C, device='cuda')
    oracle = SumTreeOracle(C)
    # bulk leaves + rebuild
    leaves = np.power(np.clip(np.abs(rng.randn(C)).astype(np.float32), 0.01, 1.0), np.float32(0.9))
    leaves[rng.rand(C) < 0.2] = 0
    nodes[C:] = torch.from_numpy(leaves).cuda()
    assert lib.asac_tree_rebuild(nodes.data_ptr(), C, s) == 0
    oracle.nodes[C - 1:] = leaves
    oracle.rebuild()
    assert np.array_equal(_np(nodes[1:]), oracle.nodes)
    # incremental updates (k = 256 with duplicates, then k = 3000 > one chunk)
    for k in (256, 3000):
        idx = rng.randint(0, C, size=k).astype(np.int64)
        idx[5] = idx[3]
        p = rng.rand(k).astype(np.float32)
        ti, tp = torch.from_numpy(idx).cuda(), torch.from_numpy(p).cuda()
        assert lib.asac_tree_update(nodes.data_ptr(), C, ti.data_ptr(), tp.data_ptr(), k, s) == 0
        oracle.update(idx, p)
        got = _np(nodes[1:])
        assert np.array_equal(got, oracle.nodes), f'k={k}: {np.sum(got != oracle.nodes)} nodes differ'
    out = torch.zeros(1, device='cuda')
    assert lib.asac_tree_leaf_max(nodes.data_ptr(), C, out.data_ptr(), s) == 0
    assert float(out.item()) == float(oracle.leaf_max)
    # sampling: injected uniforms, bit-exact leaves
    B = 4096
    u = rng.random_sample(B)
    tu = torch.from_numpy(u).cuda()
    slot = torch.zeros(B, dtype=torch.int32, device='cuda')
    pr = torch.zeros(B, device='cuda')
    assert lib.asac_tree_sample(nodes.data_ptr(), C, B, tu.data_ptr(), 0, None, slot.data_ptr(), pr.data_ptr(),
                                s) == 0
    leaf, p_ref = oracle.descend(oracle.draw(B, u))
    assert np.array_equal(_np(slot), leaf - (C - 1))
    assert np.array_equal(_np(pr), p_ref)
    # device-generated uniforms: every sample lands in its stratum and on a non-zero leaf
    counter = torch.zeros(1, dtype=torch.int64, device='cuda')
    assert lib.asac_tree_sample(nodes.data_ptr(), C, B, None, 77, counter.data_ptr(), slot.data_ptr(),
                                pr.data_ptr(), s) == 0
    sl = _np(slot).astype(np.int64)
    leaves = oracle.leaves().copy()
    assert (_np(pr) > 0).all() and np.array_equal(_np(pr), leaves[sl])
    cum = np.cumsum(leaves.astype(np.float64))
    seg = float(oracle.total) / B
    lo = np.concatenate([[0.], cum[:-1]])[sl]
    assert (cum[sl] >= np.arange(B) * seg * (1 - 1e-5)).all() and (lo <= (np.arange(B) + 1) * seg * (1 + 1e-5)).all()


def test_nan_td_error_raises_and_leaves_tree():
    from asac_b200 import PrioritizedReplayBuffer
    rb = PrioritizedReplayBuffer(batch_size=4, device='cuda:0', capacity=64)
    rb.add({'index': np.arange(20, dtype=np.int32), 'x': np.random.randn(20, 3).astype(np.float32)})
    before = rb.tree_nodes().clone()
    ids = np.arange(4, dtype=np.int64)
    rb.update(ids, np.array([0.5, np.nan, 0.1, 0.2], dtype=np.float32))
    with pytest.raises(Exception, match='td_error has nan'):
        rb.check_nan()
    assert torch.equal(before, rb.tree_nodes())
    assert rb.sample() is not None  # still usable


@pytest.mark.parametrize('name', ['pad_b2n3.npz', 'pad_b0n1.npz'])
def test_fused_gather_padding_matches_reference(name):
    """The gather kernel with role substitution == reference sample + _sample_from_replay_buffer
    padding (sac_base.py:2435-2453)."""
    from asac_b200 import PrioritizedReplayBuffer, _lib
    g = load_golden(name)
    b, n, B, capacity = [int(x) for x in g['meta']]
    L = b + n + 1
    raw = {k[4:]: v for k, v in g.items() if k.startswith('raw.')}
    # rebuild a ring whose rows at the sampled windows equal the raw gathered rows
    rb = PrioritizedReplayBuffer(batch_size=B, sample_prev_n=b, sample_post_n=n, device='cuda:0', capacity=capacity)
    ids = g['data_ids'].astype(np.int64)
    window = (ids[:, None] + np.arange(-b, n + 1)[None, :]).reshape(-1)
    slots = window % capacity
    cols = {}
    for k, v in raw.items():
        flat = v.reshape(B * L, *v.shape[2:])
        ring = np.zeros((capacity, *flat.shape[1:]), dtype=flat.dtype)
        ring[slots] = flat
        cols[k] = torch.from_numpy(ring).cuda()
    rb._columns = cols
    dev = 'cuda:0'
    A, S = raw['action'].shape[-1], raw['obs_vector'].shape[-1]
    out = {'index': torch.zeros(B, L, dtype=torch.int32, device=dev),
           'last_mask': torch.zeros(B, L, dtype=torch.uint8, device=dev),
           'action': torch.zeros(B, L, A, device=dev), 'reward': torch.zeros(B, L, device=dev),
           'done': torch.zeros(B, L, dtype=torch.uint8, device=dev), 'mu_prob': torch.zeros(B, L, A, device=dev),
           'states': torch.zeros(B, L, S, device=dev)}
    specs = [('index', out['index'], 4, 0, _lib.ROLE_INDEX), ('last_mask', out['last_mask'], 1, 0, _lib.ROLE_COPY),
             ('action', out['action'], 4 * A, 0, _lib.ROLE_ACTION), ('reward', out['reward'], 4, 0, _lib.ROLE_REWARD),
             ('done', out['done'], 1, 0, _lib.ROLE_DONE), ('mu_prob', out['mu_prob'], 4 * A, 0, _lib.ROLE_MU_PROB),
             ('obs_vector', out['states'], 4 * S, 0, _lib.ROLE_COPY)]
    mask = torch.zeros(B, L, dtype=torch.uint8, device=dev)
    rb._gather(torch.from_numpy(ids).cuda(), specs, torch.zeros(A, device=dev), mask)
    assert np.array_equal(_np(out['index'])[:, :-1], g['padded.bn_indexes'])
    assert np.array_equal(_np(mask)[:, :-1].astype(bool), g['padded.bn_padding_masks'])
    assert np.array_equal(_np(out['last_mask'])[:, :-1].astype(bool), g['padded.bn_last_masks'])
    assert np.array_equal(_np(out['action'])[:, :-1], g['padded.bn_actions'])
    assert np.array_equal(_np(out['reward'])[:, :-1], g['padded.bn_rewards'])
    assert np.array_equal(_np(out['done'])[:, :-1].astype(bool), g['padded.bn_dones'])
    assert np.array_equal(_np(out['mu_prob'])[:, :-1], g['padded.bn_mu_probs'])
    assert np.array_equal(_np(out['states']), g['padded.bnx_obs'])


def test_write_back_skips_padding_and_stale_ids():
    from asac_b200 import PrioritizedReplayBuffer
    rng = np.random.RandomState(1)
    C, B, b, n, A = 64, 8, 1, 2, 3
    rb = PrioritizedReplayBuffer(batch_size=B, sample_prev_n=b, sample_post_n=n, device='cuda:0', capacity=C)
    T = 90  # wraps the ring: ids 0..25 are overwritten
    rb.add({'index': np.arange(T, dtype=np.int32), 'mu_prob': rng.rand(T, A).astype(np.float32)})
    ref = _np(rb._columns['mu_prob']).copy()
    store = _np(rb._store_ids)
    ids = np.array([30, 40, 89, 10, 88, 27, 63, 64], dtype=np.int64)
    # overlapping windows write the same value to the same row (as pi_probs of one stored action do)
    wids = ids[:, None] - b + np.arange(b + n)[None, :]
    rows = (np.sin(wids[..., None] * 0.37 + np.arange(A)[None, None, :]) * 0.5 + 0.5).astype(np.float32)
    pad = rng.rand(B, b + n + 1) < 0.3
    pad[ids == 63] = False; pad[ids == 64] = False  # make the overlap observable
    rb.write_back(torch.from_numpy(ids).cuda(), 'mu_prob', torch.from_numpy(rows).cuda(), -b,
                  torch.from_numpy(pad.astype(np.uint8)).cuda())
    for i in range(B):
        for t in range(b + n):
            wid = ids[i] - b + t
            if not pad[i, t] and store[wid % C] == wid:
                ref[wid % C] = rows[i, t]
    assert np.array_equal(_np(rb._columns['mu_prob']), ref)


def _rows(rng, T, first_index=0):
    return {'index': (np.arange(T) + first_index).astype(np.int32), 'last_mask': np.arange(T) == T - 1,
            'obs_vector': rng.randn(T, 3).astype(np.float32), 'action': rng.rand(T, 2).astype(np.float32),
            'reward': rng.randn(T).astype(np.float32), 'done': rng.randint(0, 2, size=T).astype(bool),
            'mu_prob': rng.rand(T, 2).astype(np.float32),
            'pre_seq_hidden_state': np.zeros((T, 0), dtype=np.float32)}


def _same_state(rb, ref, what):
    assert np.array_equal(_np(rb.tree_nodes()), ref.tree.nodes), f'{what}: tree'
    assert np.array_equal(_np(rb._store_ids), ref.store.columns['_id']), f'{what}: ids'
    assert rb.size == ref.store.size and rb._next_id == ref.store.next_id, f'{what}: size / next id'
    for k, col in ref.store.columns.items():
        if k != '_id':
            assert np.array_equal(_np(rb._columns[k]), col), f'{what}: column {k}'


def test_ring_wrap_long_episodes_and_unsorted_updates_against_oracle():
    """Edge cases of replay_buffer.py the golden traces do not reach, against the NumPy oracle (itself
    pinned to the reference): ids wrapping at 10 * capacity (:28, 43-54), an episode longer than the
    ring (only the last `capacity` rows survive, :48-50), ring-tail / episode-tail zero priorities
    (:303-306), priority updates given in arbitrary order with duplicates and stale ids (sort path of
    the tree kernel), add_with_td_error (:317-337), and sample() before the buffer exceeds a batch."""
    from asac_b200 import PrioritizedReplayBuffer
    from oracle.replay_oracle import PerOracle
    C, B = 16, 4
    rb = PrioritizedReplayBuffer(batch_size=B, sample_prev_n=1, sample_post_n=2, device='cuda:0', capacity=C,
                                 alpha=0.7)
    ref = PerOracle(batch_size=B, sample_prev_n=1, sample_post_n=2, capacity=C, alpha=0.7)
    rng = np.random.RandomState(4)
    first = _rows(rng, 3)
    rb.add(first, ignore_size=1); ref.add(first, ignore_size=1)
    assert rb.sample() is None and ref.sample(rng.random_sample(B)) is None  # size 3 <= batch 4
    for step in range(40):  # 40 episodes of 5..9 rows: ids pass 10 * C = 160 more than once
        T = 5 + step % 5
        ep = _rows(rng, T)
        if step % 7 == 3:
            td = np.abs(rng.randn(T)).astype(np.float32)
            rb.add_with_td_error(td, ep, ignore_size=1); ref.add_with_td_error(td, ep, ignore_size=1)
        else:
            rb.add(ep, ignore_size=1); ref.add(ep, ignore_size=1)
        leaves, leaves_ref = _np(rb.tree_nodes())[C - 1:], ref.tree.nodes[C - 1:]
        if step % 7 == 3:  # np.power vs float64 pow: an ulp or two on the new leaves (see test above)
            ulp = np.abs(leaves.view(np.int32).astype(np.int64) - leaves_ref.view(np.int32).astype(np.int64))
            assert ulp.max() <= 4
            rb._nodes[1:].copy_(torch.from_numpy(ref.tree.nodes).cuda())
        _same_state(rb, ref, f'episode {step}')
        u = rng.random_sample(B)
        got, want = rb.sample(unit_uniform=u), ref.sample(u)
        assert np.array_equal(_np(got[0]), want[0]), f'step {step}: data ids'
        for k in want[1]:
            assert np.array_equal(_np(got[1][k]), want[1][k]), (step, k)
        # priority update in arbitrary order: shuffled, one duplicate, one id that is no longer resident
        ids = want[0][rng.permutation(B)].copy()
        ids[1] = ids[0]
        ids[2] -= 3 * C
        p = np.power(np.clip(np.abs(rng.randn(B)).astype(np.float32), 0.01, 1.0), np.float32(0.7))
        alive = ref.store.ids_at(ids) == ids
        ref.tree.update(ids[alive] % C, p[alive])
        t_ids, t_p = torch.from_numpy(ids).cuda(), torch.from_numpy(p).cuda()
        assert rb._lib.asac_per_update(rb._nodes.data_ptr(), C, rb._store_ids.data_ptr(), t_ids.data_ptr(),
                                       t_p.data_ptr(), B, 0.01, 1.0, 0.7, 1, rb._per_state.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream) == 0
        _same_state(rb, ref, f'update {step}')
    assert ref.store.next_id < 10 * C and rb._next_id == ref.store.next_id
    # an episode longer than the ring
    long_ep = _rows(rng, 3 * C + 5)
    rb.add(long_ep, ignore_size=1); ref.add(long_ep, ignore_size=1)
    _same_state(rb, ref, 'long episode')
    rb.close()


def test_capacity_rounding_clear_copy_and_storage_access():
    """The remaining public surface of PrioritizedReplayBuffer (replay_buffer.py:264, 401-410, 448-470):
    capacity rounded DOWN to a power of two, get_storage_data / get_storage_data_ids, copy() into a
    second buffer (same samples afterwards), clear() (empty again, sample() is None, is_full False)."""
    from asac_b200 import PrioritizedReplayBuffer
    B = 4
    rb = PrioritizedReplayBuffer(batch_size=B, sample_prev_n=0, sample_post_n=1, device='cuda:0', capacity=24, alpha=0.9)
    assert rb.capacity == 16  # 2 ** floor(log2(24))
    rng = np.random.RandomState(8)
    eps = [_rows(rng, 7), _rows(rng, 6)]
    for ep in eps:
        rb.add(ep, ignore_size=1)
    assert rb.size == 13 and not rb.is_full and rb.is_lg_batch_size and rb.get_curr_id() == 13
    ids = np.array([0, 5, 7, 12], dtype=np.int64)
    got = rb.get_storage_data(ids)
    stacked = {k: np.concatenate([e[k] for e in eps]) for k in eps[0]}
    for k, v in stacked.items():
        assert np.array_equal(_np(got[k]), v[ids]), k
    assert np.array_equal(_np(rb.get_storage_data_ids(ids)), ids)
    other = PrioritizedReplayBuffer(batch_size=B, sample_prev_n=0, sample_post_n=1, device='cuda:0', capacity=16, alpha=0.9)
    other.copy(rb)
    assert other.size == rb.size and torch.equal(other.tree_nodes(), rb.tree_nodes())
    u = rng.random_sample(B)
    a, b = rb.sample(unit_uniform=u), other.sample(unit_uniform=u)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    for k in a[1]:
        assert torch.equal(a[1][k], b[1][k]), k
    rb.add(_rows(rng, 9), ignore_size=1)  # 22 rows into 16 slots
    assert rb.is_full and rb.size == 16
    rb.clear()
    assert rb.size == 0 and not rb.is_full and rb.sample() is None and float(rb.tree_nodes()[1]) == 0.0
    rb.add(eps[0], ignore_size=1)  # usable again after clear()
    assert rb.size == 7
    rb.close(); other.close()


def test_sharded_importance_weights_are_normalised_over_all_shards():
    """Sharded replay (one tree per GPU, B draws each; SURVEY §7 asks for a statistical check): with the weights
    of asac_per_shard_weights — sampling probability p / total_of_the_shard, normalised by the smallest probability
    over BOTH shards' batches — the importance-weighted batch estimate of a per-transition quantity converges to its
    UNIFORM mean over the union of the shards (beta = 1), as replay_buffer.py:352-354 does for one buffer; the
    shard-local normalisation (each shard's own minimum) does not.  Two shards emulated on one GPU."""
    from asac_b200 import _lib
    from asac_b200._lib import check, ptr
    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    C, B, R = 1024, 64, 400
    rng = np.random.RandomState(3)
    pri = [rng.uniform(0.5, 1.0, C).astype(np.float32),                                  # flat priorities
           np.where(rng.rand(C) < 0.5, 0.05, 1.0).astype(np.float32)]                    # half of them 20x smaller
    shards = []
    for p in pri:
        nodes = torch.zeros(2 * C, device='cuda')
        idx = torch.arange(C, dtype=torch.int64, device='cuda')
        check(lib.asac_tree_update(ptr(nodes), C, ptr(idx), ptr(torch.from_numpy(p).cuda()), C, s), 'tree_update')
        shards.append(dict(nodes=nodes, ids=idx.clone(), state=torch.tensor([1.0, 0.0, 0.0, 0.0], dtype=torch.float64, device='cuda'),
                           counter=torch.zeros(1, dtype=torch.int64, device='cuda'),
                           slot=torch.zeros(B, dtype=torch.int32, device='cuda'), did=torch.zeros(B, dtype=torch.int64, device='cuda'),
                           p=torch.zeros(B, device='cuda'), w=torch.zeros(B, device='cuda'), wg=torch.zeros(B, device='cuda')))
    # the same transitions in ONE buffer of 2C slots drawing 2B per step: the reference's own scheme
    one = dict(nodes=torch.zeros(4 * C, device='cuda'), ids=torch.arange(2 * C, dtype=torch.int64, device='cuda'),
               state=torch.tensor([1.0, 0.0, 0.0, 0.0], dtype=torch.float64, device='cuda'),
               counter=torch.zeros(1, dtype=torch.int64, device='cuda'), slot=torch.zeros(2 * B, dtype=torch.int32, device='cuda'),
               did=torch.zeros(2 * B, dtype=torch.int64, device='cuda'), p=torch.zeros(2 * B, device='cuda'),
               w=torch.zeros(2 * B, device='cuda'))
    check(lib.asac_tree_update(ptr(one['nodes']), 2 * C, ptr(one['ids']), ptr(torch.from_numpy(np.concatenate(pri)).cuda()),
                               2 * C, s), 'tree_update')
    num_g = den_g = num_l = den_l = num_1 = den_1 = 0.0
    for r in range(R):
        for k, sh in enumerate(shards):
            check(lib.asac_per_sample(ptr(sh['nodes']), C, ptr(sh['ids']), B, None, 1000 + k, ptr(sh['counter']),
                                      ptr(sh['state']), ptr(sh['slot']), ptr(sh['did']), ptr(sh['p']), ptr(sh['w']), s),
                  'per_sample')
        gmin = torch.minimum(shards[0]['state'][2:3], shards[1]['state'][2:3]).clone()
        for sh in shards:
            check(lib.asac_per_shard_weights(ptr(sh['nodes']), B, ptr(sh['p']), ptr(sh['state']), ptr(gmin), ptr(sh['wg']), s),
                  'per_shard_weights')
        check(lib.asac_per_sample(ptr(one['nodes']), 2 * C, ptr(one['ids']), 2 * B, None, 77, ptr(one['counter']),
                                  ptr(one['state']), ptr(one['slot']), ptr(one['did']), ptr(one['p']), ptr(one['w']), s),
              'per_sample')
        # f = 1 on shard 1's transitions, 0 on shard 0's: uniform mean over the union = 0.5
        num_g += float(shards[1]['wg'].sum()); den_g += float(shards[0]['wg'].sum() + shards[1]['wg'].sum())
        num_l += float(shards[1]['w'].sum()); den_l += float(shards[0]['w'].sum() + shards[1]['w'].sum())
        num_1 += float(one['w'][one['slot'] >= C].sum()); den_1 += float(one['w'].sum())
        if r < 3:
            # deterministic part: with beta = 1, weight x sampling probability is ONE constant (the global minimum)
            # for every draw of every shard; with the shard-local rule each shard has its own constant
            for sh in shards:
                prob = sh['p'] / sh['nodes'][1]
                const = (sh['wg'].double() * prob.double())
                assert torch.allclose(const, gmin.expand_as(const), rtol=1e-5), (const.min(), const.max(), gmin)
            top = max(float(shards[0]['wg'].max()), float(shards[1]['wg'].max()))
            assert top == 1.0
            # one shard alone: the global rule IS the reference's rule, bit for bit
            own = torch.zeros(B, device='cuda')
            check(lib.asac_per_shard_weights(ptr(shards[0]['nodes']), B, ptr(shards[0]['p']), ptr(shards[0]['state']),
                                             ptr(shards[0]['state'][2:3].clone()), ptr(own), s), 'per_shard_weights')
            assert torch.equal(own, shards[0]['w'])
    est_g, est_l, est_1 = num_g / den_g, num_l / den_l, num_1 / den_1
    print(f'importance-weighted estimate of the shard-1 indicator (uniform mean 0.5): one buffer {est_1:.4f}, two shards '
          f'with global normalisation {est_g:.4f}, with shard-local normalisation {est_l:.4f}')
    # (normalising every batch by ITS minimum down-weights the batches that drew a rare transition: a few per cent of
    #  bias that the single buffer of the reference has as well)
    assert abs(est_1 - 0.5) < 0.06, est_1
    assert abs(est_g - 0.5) < 0.06, est_g
    assert abs(est_l - 0.5) > 0.2, est_l
