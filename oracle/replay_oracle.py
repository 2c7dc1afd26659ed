"""NumPy restatement of the reference's prioritized replay (TEST INFRASTRUCTURE).

Follows ``/root/reference/algorithm/replay_buffer.py``:

* ``SumTreeOracle``      <- ``SumTree``                  (replay_buffer.py:145-242)
* ``RingStorageOracle``  <- ``DataStorage``              (replay_buffer.py:21-142)
* ``PerOracle``          <- ``PrioritizedReplayBuffer``  (replay_buffer.py:245-477),
  without the prefetch thread / lock / pinned-memory plumbing: ``sample`` here is
  the body of ``_prefetch_loop`` (:339-375) executed synchronously.
* ``pad_sampled_batch``  <- ``SAC_Base._sample_from_replay_buffer`` padding block
  (sac_base.py:2435-2453).

The node array is kept in the reference's own order (root at 0, children of i
at 2i+1 / 2i+2, leaves at [C-1, 2C-2]) so it can be compared bit-for-bit with a
reference ``*-rb_tree.npy`` file and with the CUDA tree after export.

All randomness is injected: ``sample`` takes the float64 uniforms the reference
would have drawn from ``np.random.uniform`` (replay_buffer.py:194).
"""
from __future__ import annotations

import math

import numpy as np


class SumTreeOracle:
    """replay_buffer.py:145-242."""

    def __init__(self, capacity: int):
        capacity = int(capacity)
        if capacity <= 0 or capacity & (capacity - 1):
            raise AssertionError('capacity must be a power of two')  # :148
        self.capacity = capacity
        self.levels = int(math.log2(capacity))  # child levels below the root (= depth - 1, :151)
        self.nodes = np.zeros(2 * capacity - 1, dtype=np.float32)  # :153

    # -- index helpers (:207-211)
    def leaf_of(self, data_idx):
        return np.asarray(data_idx) + (self.capacity - 1)

    def data_of(self, leaf_idx):
        return np.asarray(leaf_idx) - (self.capacity - 1)

    @property
    def total(self) -> np.float32:  # :233-235
        return self.nodes[0]

    @property
    def leaf_max(self) -> np.float32:  # :237-239
        return self.nodes[self.capacity - 1:].max()

    def leaves(self) -> np.ndarray:
        return self.nodes[self.capacity - 1:]

    def update(self, data_idx, p) -> None:
        """:172-183.  Leaves are overwritten (last duplicate wins, as NumPy fancy
        assignment does) and every ancestor is recomputed as fp32 ``left + right``."""
        data_idx = np.asarray(data_idx, dtype=np.int64).reshape(-1)
        p = np.asarray(p, dtype=np.float32).reshape(-1)
        touched = data_idx + (self.capacity - 1)
        for t, value in zip(touched, p):  # sequential == "last wins"
            self.nodes[t] = value
        level = np.unique(touched)
        for _ in range(self.levels):
            level = np.unique((level - 1) >> 1)
            self.nodes[level] = self.nodes[2 * level + 1] + self.nodes[2 * level + 2]

    def rebuild(self) -> None:
        """Full bottom-up recomputation; bit-identical to any sequence of
        ``update`` calls that produced the same leaves."""
        c = self.capacity
        lo = c - 1
        while lo > 0:
            parents = np.arange((lo - 1) // 2, lo)
            self.nodes[parents] = self.nodes[2 * parents + 1] + self.nodes[2 * parents + 2]
            lo = (lo - 1) // 2

    def strata_bounds(self, batch_size: int):
        """:190-193.  ``seg`` is an fp32 scalar, the bounds are float64."""
        seg = self.total / batch_size  # np.float32 / int -> np.float32
        lo = np.arange(batch_size)
        return lo * seg, (lo + 1) * seg  # int64 array * f32 scalar -> float64

    def draw(self, batch_size: int, unit_uniform: np.ndarray) -> np.ndarray:
        """``np.random.uniform(low, high)`` is ``low + (high-low)*u`` with
        ``u = random_sample()``; restated so tests can inject ``u``."""
        lo, hi = self.strata_bounds(batch_size)
        return lo + (hi - lo) * np.asarray(unit_uniform, dtype=np.float64)

    def descend(self, v: np.ndarray):
        """:195-205, one sample at a time (scalar restatement)."""
        v = np.array(v, dtype=np.float64).reshape(-1)
        out_idx = np.zeros(v.shape[0], dtype=np.int32)
        for s in range(v.shape[0]):
            node, x = 0, v[s]
            for _ in range(self.levels):
                left, right = 2 * node + 1, 2 * node + 2
                # float64 x against fp32 nodes (NumPy promotes the node to f64)
                if x <= np.float64(self.nodes[left]) or self.nodes[right] == 0:
                    node = left
                else:
                    x = x - np.float64(self.nodes[left])
                    node = right
            out_idx[s] = node
        return out_idx, self.nodes[out_idx]


    def descend_vectorized(self, v: np.ndarray):
        """:195-205 as the reference runs it: one vectorised pass per level (used by the timed
        CPU baseline; ``descend`` above is the scalar restatement the tests pin it to)."""
        v = np.array(v, dtype=np.float64).reshape(-1)
        node = np.zeros(v.shape[0], dtype=np.int32)
        for _ in range(self.levels):
            left, right = 2 * node + 1, 2 * node + 2
            go_left = np.logical_or(v <= self.nodes[left], self.nodes[right] == 0)
            node = np.where(go_left, left, right)
            v = np.where(go_left, v, v - self.nodes[left])
        return node, self.nodes[node]


class RingStorageOracle:
    """replay_buffer.py:21-142 (``DataStorage``)."""

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.max_id = 10 * self.capacity  # :28
        self.size = 0
        self.next_id = 0
        self.columns: dict[str, np.ndarray] | None = None

    def add(self, rows: dict[str, np.ndarray]) -> np.ndarray:  # :30-56
        n = next(iter(rows.values())).shape[0]
        if self.columns is None:
            self.columns = {'_id': np.zeros(self.capacity, dtype=np.int64)}
            for k, v in rows.items():
                self.columns[k] = np.zeros((self.capacity, *v.shape[1:]), dtype=v.dtype)
        ids = (np.arange(n) + self.next_id) % self.max_id
        slots = ids % self.capacity
        self.columns['_id'][slots] = ids
        for k, v in rows.items():
            self.columns[k][slots] = v
        self.size = min(self.size + n, self.capacity)
        self.next_id = int(ids[-1]) + 1
        if self.next_id == self.max_id:
            self.next_id = 0
        return slots

    def ids_at(self, ids) -> np.ndarray:  # :90-94
        return self.columns['_id'][np.asarray(ids) % self.capacity]

    def rows_at(self, ids) -> dict[str, np.ndarray]:  # :64-75
        slots = np.asarray(ids) % self.capacity
        return {k: v[slots] for k, v in self.columns.items() if k != '_id'}

    def write(self, ids, key: str, data) -> None:  # :61-62
        self.columns[key][np.asarray(ids) % self.capacity] = data


class PerOracle:
    """replay_buffer.py:245-477 without threads."""

    def __init__(self, batch_size=256, sample_prev_n=0, sample_post_n=0, capacity=524288,
                 alpha=0.9, beta=0.4, beta_increment_per_sampling=0.001,
                 td_error_min=0.01, td_error_max=1.):
        self.batch_size = batch_size
        self.prev_n, self.post_n = sample_prev_n, sample_post_n
        self.capacity = int(2 ** math.floor(math.log2(capacity)))  # :264
        self.alpha, self.beta = alpha, beta
        self.beta_inc = beta_increment_per_sampling
        self.td_min, self.td_max = td_error_min, td_error_max
        self.tree = SumTreeOracle(self.capacity)
        self.store = RingStorageOracle(self.capacity)
        self.vectorized = False  # True: level-vectorised descent (same results, reference speed)

    def _mask_tail(self, slots, probs, ignore_size):  # :303-306
        if ignore_size > 0:
            probs[(slots >= self.capacity - ignore_size) & (slots < self.capacity)] = 0
            probs[-ignore_size:] = 0

    def add(self, rows, ignore_size=0):  # :293-307
        max_p = self.td_max if self.store.size == 0 else self.tree.leaf_max
        slots = self.store.add(rows)
        probs = np.full(len(slots), max_p, dtype=np.float32)
        self._mask_tail(slots, probs, ignore_size)
        self.tree.update(slots, probs)
        return slots

    def priority_of(self, td_error) -> np.ndarray:  # :416-422
        clipped = np.clip(np.asarray(td_error).flatten(), self.td_min, self.td_max)
        if np.isnan(np.min(clipped)):
            raise Exception('td_error has nan')
        return np.power(clipped, self.alpha)

    def add_with_td_error(self, td_error, rows, ignore_size=0):  # :317-337
        slots = self.store.add(rows)
        probs = self.priority_of(td_error)
        self._mask_tail(slots, probs, ignore_size)
        self.tree.update(slots, probs)
        return slots

    @property
    def is_lg_batch_size(self):  # :467-470
        return self.store.size > self.batch_size

    def is_weights(self, p: np.ndarray) -> np.ndarray:  # :352-354
        w = p / self.tree.total
        self.beta = np.min([1., self.beta + self.beta_inc])
        return np.power(w / np.min(w), -self.beta).astype(np.float32)

    def sample(self, unit_uniform: np.ndarray):
        """Body of ``_prefetch_loop`` (:347-364) + ``sample`` (:396)."""
        if not self.is_lg_batch_size:
            return None
        v = self.tree.draw(self.batch_size, unit_uniform)
        leaf, p = self.tree.descend_vectorized(v) if self.vectorized else self.tree.descend(v)
        data_ids = self.store.ids_at(self.tree.data_of(leaf))
        weights = self.is_weights(p)
        offsets = np.arange(-self.prev_n, self.post_n + 1, dtype=np.int64)
        window = (data_ids[:, None] + offsets[None, :]).reshape(-1)
        span = self.prev_n + 1 + self.post_n
        batch = {k: a.reshape(self.batch_size, span, *a.shape[1:])
                 for k, a in self.store.rows_at(window).items()}
        return data_ids, batch, weights[:, None], leaf

    def update(self, data_ids, td_error):  # :412-427
        data_ids = np.asarray(data_ids)
        probs = self.priority_of(td_error)
        alive = self.store.ids_at(data_ids) == data_ids
        self.tree.update(data_ids[alive] % self.capacity, probs[alive])

    def update_transitions(self, data_ids, key, data):  # :429-434
        data_ids = np.asarray(data_ids)
        alive = self.store.ids_at(data_ids) == data_ids
        self.store.write(data_ids[alive], key, np.asarray(data)[alive])


def pad_sampled_batch(batch: dict[str, np.ndarray], burn_in_step: int,
                      padding_action: np.ndarray) -> dict[str, np.ndarray]:
    """sac_base.py:2435-2453: rows whose stored ``index`` does not continue the
    anchor row's episode are turned into padding."""
    out = {k: v.copy() for k, v in batch.items()}
    index = out['index']
    span = index.shape[1]
    rel = np.arange(span) - burn_in_step
    bad = (index - index[:, burn_in_step:burn_in_step + 1]) != rel[None, :]
    bad[:, burn_in_step] = False
    out['padding_mask'] = bad.copy()
    out['index'][bad] = -1
    out['action'][bad] = padding_action
    out['reward'][bad] = 0.
    out['done'][bad] = True
    out['mu_prob'][bad] = 1.
    if 'pre_seq_hidden_state' in out:
        out['pre_seq_hidden_state'][bad] = 0.
    return out
