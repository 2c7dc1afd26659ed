"""Pins the CPU oracle (oracle/*.py) to the fixtures minted from the real reference
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle.replay_oracle import SumTreeOracle
from tests.helpers import load_golden
from tests.oracle_checks import (check_hybrid_sac_steps, check_padding, check_per_trace,
                                 check_recurrent_sac_steps, check_sac_steps)

@pytest.mark.parametrize('name', ['per_small.npz', 'per_zeros.npz'])
def test_per_trace_bit_exact(name):
    check_per_trace(load_golden(name))


def test_rebuild_equals_incremental():
    g = load_golden('per_zeros.npz')
    capacity = int(g['meta'][0])
    t = SumTreeOracle(capacity)
    ref = g['r1.tree_after_add']
    t.nodes[capacity - 1:] = ref[capacity - 1:]
    t.rebuild()
    assert np.array_equal(t.nodes, ref)


@pytest.mark.parametrize('name', ['pad_b2n3.npz', 'pad_b0n1.npz'])
def test_padding_matches_reference(name):
    check_padding(load_golden(name))


@pytest.mark.parametrize('name', ['sac_c2.npz', 'sac_c3.npz', 'sac_odd.npz', 'sac_nois.npz', 'sac_sub.npz',
                                  'sac_c2_b256.npz', 'sac_c3_b1024.npz'])  # the last two: BASELINE's full shapes
def test_sac_oracle_matches_reference(name):
    check_sac_steps(load_golden(name))


@pytest.mark.parametrize('name', ['sac_rnn.npz', 'sac_rnn_b0.npz', 'sac_rnn_c4.npz'])  # c4: B=256, b=40, n=5 (L=46)
def test_recurrent_sac_oracle_matches_reference(name):
    """The GRU-representation flow (envs/test/nn_rnn.py, seq_encoder=RNN) against the real reference:
    BPTT gradients of the representation, re-encoded states, next hidden states, td error on the
    target states."""
    check_recurrent_sac_steps(load_golden(name))


@pytest.mark.parametrize('name', ['sac_disc.npz', 'sac_hybrid.npz', 'sac_dqn.npz'])
def test_discrete_and_hybrid_oracle_matches_reference(name):
    """Groundwork for SURVEY §8f rank 4 (no CUDA path yet): discrete-only and hybrid action branches."""
    check_hybrid_sac_steps(load_golden(name))


def test_vectorized_descent_equals_scalar():
    rng = np.random.RandomState(5)
    t = SumTreeOracle(1024)
    leaves = rng.rand(1024).astype(np.float32)
    leaves[rng.rand(1024) < 0.3] = 0
    t.nodes[1023:] = leaves
    t.rebuild()
    v = t.draw(512, rng.random_sample(512))
    a, pa = t.descend(v)
    b, pb = t.descend_vectorized(v)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)
