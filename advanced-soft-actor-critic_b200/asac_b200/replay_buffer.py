"""HBM-resident prioritized replay with the reference's ``PrioritizedReplayBuffer`` API.

Mirrors ``algorithm/replay_buffer.py:245-477`` of the reference (same constructor arguments,
method names, return structure and error behaviour) but every array lives on the GPU and every
operation is a kernel of ``libasac_b200.so``:

* ``SumTree``      -> float32[2C] 1-based heap in HBM (``asac_tree_*`` / ``asac_per_*``)
* ``DataStorage``  -> one ``[C, *shape]`` CUDA tensor per key + int64 ``_id`` column
* prefetch thread + pinned H2D copy -> not needed: sampling is a two-kernel device operation
  ordered on the caller's stream, so there is no lock and no background thread either.

Differences a caller can observe (documented in DESIGN.md §4): ``sample()`` returns the data ids
as a CUDA int64 tensor instead of a NumPy array (``update`` / ``update_transitions`` accept
both), and the NaN check of ``update`` is reported asynchronously (``check_nan()``; it is polled
every ``nan_check_every`` updates and on ``save`` / ``close``).
"""
from __future__ import annotations

import ctypes as C
import functools
import logging
import math
import os
import threading
from pathlib import Path

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

_NP_TO_TORCH = {np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64,
                np.dtype('int32'): torch.int32, np.dtype('int64'): torch.int64,
                np.dtype('uint8'): torch.uint8, np.dtype('bool'): torch.bool,
                np.dtype('int8'): torch.int8, np.dtype('int16'): torch.int16,
                np.dtype('float16'): torch.float16}


_TORCH_TO_NP = {v: k for k, v in _NP_TO_TORCH.items()}


def _align16(n: int) -> int:
    return (n + 15) & ~15


def _locked(fn):
    """Host-side mutual exclusion for the mutating public methods: the reference guards its rings with a
    read/write lock (replay_buffer.py:273-274, utils/lock.py) because actor threads call ``add`` while the
    learner trains; here one re-entrant mutex serialises the bookkeeping (``_size``, ``_next_id``, the
    staging slot ring) — the device work itself is ordered by the stream."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self._mutex:
            return fn(self, *args, **kwargs)
    return wrapper


def _to_device(v, device) -> torch.Tensor:
    t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
    return t.to(device, non_blocking=True).contiguous()


class PrioritizedReplayBuffer:
    def __init__(self,
                 batch_size=256,
                 sample_prev_n=0,
                 sample_post_n=0,
                 device: torch.device | str | None = None,

                 capacity=524288,
                 alpha=0.9,
                 beta=0.4,
                 beta_increment_per_sampling=0.001,
                 td_error_min=0.01,
                 td_error_max=1.,
                 logger_parent_name='',
                 seed: int | None = None,
                 nan_check_every: int = 256):
        self._lib = _lib.load()  # raises when the CUDA library is missing: there is no CPU path
        if not torch.cuda.is_available():
            raise _lib.AsacError('PrioritizedReplayBuffer needs a CUDA device (no CPU fallback)')
        device = torch.device('cuda' if device is None else device)
        if device.type != 'cuda':
            raise _lib.AsacError(f'PrioritizedReplayBuffer lives in HBM; got device {device}')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = device
        self.batch_size = batch_size
        self.prev_n = sample_prev_n
        self.post_n = sample_post_n

        self.capacity = int(2 ** math.floor(math.log2(capacity)))  # replay_buffer.py:264
        self.alpha = alpha
        self._beta0 = float(beta)
        self.beta_increment_per_sampling = beta_increment_per_sampling
        self.td_error_min = td_error_min
        self.td_error_max = td_error_max
        self.max_id = 10 * self.capacity

        name = f'{logger_parent_name}.replay_buffer' if logger_parent_name != '' else 'replay_buffer'
        self._logger = logging.getLogger(name)

        with torch.cuda.device(self.device):
            self._nodes = torch.zeros(2 * self.capacity, dtype=torch.float32, device=device)
            self._store_ids = torch.zeros(self.capacity, dtype=torch.int64, device=device)
            self._per_state = torch.tensor([self._beta0, float(beta_increment_per_sampling), 0., 0.],
                                           dtype=torch.float64, device=device)
            self._draw_counter = torch.zeros(1, dtype=torch.int64, device=device)
            self._max_p = torch.zeros(1, dtype=torch.float32, device=device)
            self._td_max = torch.full((1,), float(td_error_max), dtype=torch.float32, device=device)
        self._columns: dict[str, torch.Tensor] | None = None
        self._size = 0
        self._next_id = 0
        self._seed = int(np.random.SeedSequence().entropy & 0x7FFFFFFFFFFFFFFF) if seed is None else int(seed)
        self._updates = 0
        self._ingest = None   # native ingest handle (asac_ingest_*), bound to the current rings
        self._columns_version = 0  # bumped whenever the rings change (the learner re-captures its graph)
        self._stage = [None, None, None, None]  # pinned host + device staging buffers, round robin
        self._stage_next = 0
        self._stage_limit = 64 << 20
        self._nan_check_every = nan_check_every
        self._closed = False
        self._mutex = threading.RLock()
        # a learner that defers its priority update registers a hook here; everything that READS the tree
        # from outside the learner's own schedule (sample / save / copy) applies the pending update first
        self._flush_hook = None
        # add() gives new rows the maximum leaf priority (replay_buffer.py:296-299).  With a deferred priority update
        # pending, that maximum is the one BEFORE the last step's update (the update runs on the next step's parallel
        # branch); ASAC_STRICT_ADD_ORDER=1 applies the pending update first — the reference's order when train() and
        # put_episode() alternate in one thread — at the price of putting the 12 us tree update back on the stream
        # in front of every add.  (The maximum saturates at td_error_max ** alpha in a running learner, so the two
        # orders give the same priorities except while it is still being reached.)
        self._strict_add_order = os.environ.get('ASAC_STRICT_ADD_ORDER', '0') == '1'

    # ------------------------------------------------------------------ helpers
    @property
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ids_tensor(self, ids) -> torch.Tensor:
        if isinstance(ids, torch.Tensor):
            return ids.to(self.device, dtype=torch.int64).contiguous().view(-1)
        return torch.from_numpy(np.ascontiguousarray(np.asarray(ids, dtype=np.int64))).to(self.device).view(-1)

    def _row_bytes(self, key: str) -> int:
        col = self._columns[key]
        return col[0].numel() * col.element_size()

    def _allocate(self, transitions: dict) -> None:
        self._drop_ingest()
        self._columns = {}
        for k, v in transitions.items():
            dtype = v.dtype if isinstance(v, torch.Tensor) else _NP_TO_TORCH[np.asarray(v).dtype]
            self._columns[k] = torch.zeros((self.capacity, *v.shape[1:]), dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ add
    def _store(self, transitions: dict) -> tuple[int, int]:
        """DataStorage.add (replay_buffer.py:30-56) without the priority part -> (first_id, T).

        Host arrays take the staged path: every column of the episode is packed into one pinned
        host buffer, crosses PCIe in ONE async copy and is written into the rings by ONE kernel
        (asac_storage_write_table).  Device tensors are written column by column."""
        T = int(next(iter(transitions.values())).shape[0])
        if self._columns is None:
            self._allocate(transitions)
        first_id = self._next_id
        # an episode longer than the ring overwrites its own head: only the last `capacity` rows
        # survive (NumPy fancy assignment keeps the last write, replay_buffer.py:48-50)
        skip = max(0, T - self.capacity)
        for k, v in transitions.items():
            col = self._columns[k]
            dtype = v.dtype if isinstance(v, torch.Tensor) else _NP_TO_TORCH[np.asarray(v).dtype]
            if dtype != col.dtype or tuple(v.shape[1:]) != tuple(col.shape[1:]):
                raise ValueError(f'column {k}: got {dtype}{tuple(v.shape[1:])}, '
                                 f'stored {col.dtype}{tuple(col.shape[1:])}')
        n = T - skip
        total = sum(_align16(n * self._row_bytes(k)) for k in transitions)
        if all(not isinstance(v, torch.Tensor) for v in transitions.values()) and \
                len(transitions) <= _lib.MAX_COLUMNS and total <= self._stage_limit:
            self._store_staged(transitions, first_id, skip, n, total)
            return first_id, T
        for k, v in transitions.items():
            src = _to_device(v[skip:] if skip else v, self.device)
            col = self._columns[k]
            check(self._lib.asac_storage_write_rows(ptr(col), self.capacity, (first_id + skip) % self.max_id,
                                                    ptr(src), n, self._row_bytes(k), self._stream),
                  'storage_write_rows')
        return first_id, T

    def _store_staged(self, transitions: dict, first_id: int, skip: int, n: int, total: int) -> None:
        slot = self._stage_next
        self._stage_next = (slot + 1) % len(self._stage)
        st = self._stage[slot]
        if st is None or st['host'].numel() < total:
            size = max(total, 1 << 16)
            host = torch.empty(size, dtype=torch.uint8).pin_memory()
            st = self._stage[slot] = {'host': host, 'np': host.numpy(),
                                      'dev': torch.empty(size, dtype=torch.uint8, device=self.device),
                                      'event': torch.cuda.Event()}
        else:
            st['event'].synchronize()  # the copy that last read this pinned buffer has finished
        table = _lib.AsacWriteTable()
        table.n_columns = len(transitions)
        off, base = 0, st['dev'].data_ptr()
        for i, (k, v) in enumerate(transitions.items()):
            rb = self._row_bytes(k)
            a = np.ascontiguousarray(v[skip:] if skip else v)
            st['np'][off:off + n * rb] = a.reshape(-1).view(np.uint8)
            table.col[i].ring = ptr(self._columns[k])
            table.col[i].rows = base + off
            table.col[i].row_bytes = rb
            off += _align16(n * rb)
        st['dev'][:off].copy_(st['host'][:off], non_blocking=True)
        st['event'].record(torch.cuda.current_stream(self.device))
        check(self._lib.asac_storage_write_table(C.byref(table), self.capacity, (first_id + skip) % self.max_id, n,
                                                 self._stream), 'storage_write_table')

    def _advance(self, first_id: int, T: int) -> None:
        self._size = min(self._size + T, self.capacity)
        last = (first_id + T - 1) % self.max_id
        self._next_id = last + 1
        if self._next_id == self.max_id:
            self._next_id = 0

    def _ingest_handle(self):
        """Native ingest context over the current rings (created lazily, dropped with them)."""
        if self._ingest is None:
            keys = list(self._columns)
            n = len(keys)
            rings = (C.c_void_p * n)(*[self._columns[k].data_ptr() for k in keys])
            rbs = (C.c_int64 * n)(*[self._row_bytes(k) for k in keys])
            h = C.c_void_p()
            check(self._lib.asac_ingest_create(C.byref(h), self.capacity, n, rings, rbs, ptr(self._nodes),
                                               ptr(self._store_ids), ptr(self._max_p), ptr(self._td_max)),
                  'ingest_create')
            self._ingest = h
            self._ingest_keys = keys
            self._ingest_spec = [(k, _TORCH_TO_NP[self._columns[k].dtype], tuple(self._columns[k].shape[1:]))
                                 for k in keys]
            self._ingest_ptrs = (C.c_void_p * n)()
        return self._ingest

    def _drop_ingest(self) -> None:
        """Called whenever the rings are (re)allocated or released."""
        self._columns_version += 1
        if getattr(self, '_ingest', None) is not None:
            self._lib.asac_ingest_destroy(self._ingest)
        self._ingest = None

    def _add_native(self, transitions: dict, ignore_size: int) -> bool:
        """add() for host arrays through asac_ingest_add: one call packs, copies and inserts.
        Returns False when the episode does not qualify (device tensors, key order, T > capacity)."""
        if self._columns is None:
            self._allocate(transitions)
        h = self._ingest_handle()
        if len(transitions) != len(self._ingest_spec):
            return False
        T, keep = -1, []
        for i, (k, dt, shape) in enumerate(self._ingest_spec):
            v = transitions.get(k)
            if not isinstance(v, np.ndarray) or v.dtype != dt or v.shape[1:] != shape:
                if isinstance(v, np.ndarray) and (v.dtype != dt or v.shape[1:] != shape):
                    raise ValueError(f'column {k}: got {v.dtype}{v.shape[1:]}, stored {dt}{shape}')
                return False
            if not v.flags.c_contiguous:
                v = np.ascontiguousarray(v)
            if T < 0:
                T = v.shape[0]
            elif v.shape[0] != T:
                raise ValueError('columns disagree in length')
            keep.append(v)
            self._ingest_ptrs[i] = v.__array_interface__['data'][0]
        if T <= 0 or T > self.capacity:
            return False
        first_id = self._next_id
        check(self._lib.asac_ingest_add(h, self._ingest_ptrs, T, first_id, int(ignore_size),
                                        1 if self._size == 0 else 0, self._stream), 'ingest_add')
        self._advance(first_id, T)
        return True

    @_locked
    def add(self, transitions: dict[str, np.ndarray], ignore_size=0) -> None:
        with torch.cuda.device(self.device):
            if self._strict_add_order:
                self._flush()
            if self._add_native(transitions, ignore_size):
                return
            if self._size == 0:
                max_p = self._td_max
            else:
                check(self._lib.asac_tree_leaf_max(ptr(self._nodes), self.capacity, ptr(self._max_p), self._stream),
                      'tree_leaf_max')
                max_p = self._max_p
            first_id, T = self._store(transitions)
            check(self._lib.asac_per_add(ptr(self._nodes), self.capacity, ptr(self._store_ids), first_id, T,
                                         ptr(max_p), int(ignore_size), self._stream), 'per_add')
            self._advance(first_id, T)

    @_locked
    def add_with_td_error(self, td_error: np.ndarray, transitions: dict[str, np.ndarray],
                          ignore_size: int = 0) -> None:
        with torch.cuda.device(self.device):
            td_host = np.asarray(td_error, dtype=np.float32).flatten()
            T = int(next(iter(transitions.values())).shape[0])
            if td_host.shape[0] != T:  # before anything is written
                raise ValueError('td_error and transitions disagree in length')
            if np.isnan(td_host).any():  # replay_buffer.py:418-420 raises before touching the tree
                self._logger.error('td_error has nan')
                raise Exception('td_error has nan')
            td = _to_device(td_host, self.device)
            first_id, T = self._store(transitions)
            # ids first (per_add with zero priorities), then the td-derived priorities with the
            # reference's tail masking (replay_buffer.py:330-336)
            zero = torch.zeros(1, dtype=torch.float32, device=self.device)
            check(self._lib.asac_per_add(ptr(self._nodes), self.capacity, ptr(self._store_ids), first_id, T,
                                         ptr(zero), 0, self._stream), 'per_add')
            ids = (torch.arange(T, device=self.device, dtype=torch.int64) + first_id) % self.max_id
            self._advance(first_id, T)
            self._update_priorities(ids, td, ignore_tail=int(ignore_size))

    # ------------------------------------------------------------------ sample
    def _gather(self, data_ids: torch.Tensor, specs: list[tuple[str, torch.Tensor, int, int, int]],
                padding_action: torch.Tensor | None, padding_mask: torch.Tensor) -> None:
        """specs: (key, out tensor, out_stride_bytes, out_offset_bytes, role)."""
        table = _lib.AsacColumnTable()
        if len(specs) > _lib.MAX_COLUMNS:
            raise NotImplementedError(f'more than {_lib.MAX_COLUMNS} stored keys')
        table.n_columns = len(specs)
        table.index_column = -1
        for i, (key, out, stride, offset, role) in enumerate(specs):
            col = self._columns[key]
            table.col[i].ring = ptr(col)
            table.col[i].out = out.data_ptr()
            table.col[i].row_bytes = self._row_bytes(key)
            table.col[i].out_stride = stride
            table.col[i].out_offset = offset
            table.col[i].role = role
            if key == 'index':
                table.index_column = i
        if table.index_column < 0:
            raise KeyError("the stored transitions have no int32 'index' column")
        check(self._lib.asac_storage_gather(C.byref(table), self.capacity, ptr(data_ids), int(data_ids.numel()),
                                            self.prev_n, self.post_n, ptr(padding_action), ptr(padding_mask),
                                            self._stream), 'storage_gather')

    def _draw(self, unit_uniform=None):
        B = self.batch_size
        dev = self.device
        slots = torch.empty(B, dtype=torch.int32, device=dev)
        data_ids = torch.empty(B, dtype=torch.int64, device=dev)
        p = torch.empty(B, dtype=torch.float32, device=dev)
        w = torch.empty(B, dtype=torch.float32, device=dev)
        u = None
        if unit_uniform is not None:
            u = _to_device(np.asarray(unit_uniform, dtype=np.float64), dev)
        check(self._lib.asac_per_sample(ptr(self._nodes), self.capacity, ptr(self._store_ids), B, ptr(u),
                                        self._seed, ptr(self._draw_counter), ptr(self._per_state), ptr(slots),
                                        ptr(data_ids), ptr(p), ptr(w), self._stream), 'per_sample')
        return slots, data_ids, p, w

    def sample(self, unit_uniform=None):
        """
        Returns (replay_buffer.py:377-396):
            data ids (torch.int64, CUDA): [batch, ]
            transitions: dict [batch, prev_n + 1 + post_n, *]
            priority weights: [batch, 1]
        ``unit_uniform`` (float64[batch] in [0,1)) replaces the on-device Philox draws (tests).
        """
        if not self.is_lg_batch_size:
            return None
        with self._mutex, torch.cuda.device(self.device):
            self._flush()
            _, data_ids, _, w = self._draw(unit_uniform)
            L = self.prev_n + 1 + self.post_n
            out, specs = {}, []
            for k, col in self._columns.items():
                t = torch.empty((self.batch_size, L, *col.shape[1:]), dtype=col.dtype, device=self.device)
                out[k] = t
                specs.append((k, t, self._row_bytes(k), 0, _lib.ROLE_COPY))
            mask = torch.empty((self.batch_size, L), dtype=torch.uint8, device=self.device)
            self._gather(data_ids, specs, None, mask)
        return data_ids, out, w.unsqueeze(-1)

    # ------------------------------------------------------------------ updates
    def _update_priorities(self, data_ids: torch.Tensor, td: torch.Tensor, ignore_tail: int = 0) -> None:
        k = int(data_ids.numel())
        for off in range(0, k, 1024):
            n = min(1024, k - off)
            ids_c, td_c, pre = data_ids[off:off + n], td[off:off + n], 0
            if ignore_tail > 0:
                # add_with_td_error: zero priority for the ring tail and the episode tail
                pr = torch.clamp(td_c, self.td_error_min, self.td_error_max).double() \
                    .pow(float(np.float32(self.alpha))).float()
                pr = torch.where(torch.isnan(td_c), td_c, pr)
                slot = ids_c % self.capacity
                pos = torch.arange(off, off + n, device=self.device)
                pr = torch.where((slot >= self.capacity - ignore_tail) | (pos >= k - ignore_tail),
                                 torch.zeros_like(pr), pr)
                td_c, pre = pr.contiguous(), 1
            check(self._lib.asac_per_update(ptr(self._nodes), self.capacity, ptr(self._store_ids),
                                            ptr(ids_c.contiguous()), ptr(td_c.contiguous()), n,
                                            float(self.td_error_min), float(self.td_error_max), float(self.alpha),
                                            pre, ptr(self._per_state), self._stream), 'per_update')
        self._updates += 1
        if self._nan_check_every and self._updates % self._nan_check_every == 0:
            self.check_nan()

    @_locked
    def update(self, data_ids, td_error) -> None:
        with torch.cuda.device(self.device):
            ids = self._ids_tensor(data_ids)
            td = td_error if isinstance(td_error, torch.Tensor) else torch.from_numpy(
                np.ascontiguousarray(np.asarray(td_error, dtype=np.float32)))
            td = td.to(self.device, dtype=torch.float32).contiguous().view(-1)
            if td.numel() != ids.numel():
                raise ValueError('data_ids and td_error disagree in length')
            self._update_priorities(ids, td)

    @_locked
    def update_transitions(self, data_ids, key: str, data) -> None:
        with torch.cuda.device(self.device):
            ids = self._ids_tensor(data_ids)
            col = self._columns[key]
            rows = _to_device(data, self.device).to(col.dtype).contiguous()
            if rows.shape[0] != ids.numel():
                raise ValueError('data_ids and data disagree in length')
            check(self._lib.asac_storage_scatter(ptr(col), self.capacity, ptr(self._store_ids), ptr(ids),
                                                 int(ids.numel()), 0, 1, ptr(rows), self._row_bytes(key), 1,
                                                 None, 0, self._stream), 'storage_scatter')

    def write_back(self, data_ids: torch.Tensor, key: str, rows: torch.Tensor, first_offset: int,
                   padding_mask: torch.Tensor, n_rows: int | None = None) -> None:
        """The learner's per-window write-backs (sac_base.py:2586-2605) in one launch:
        ring[data_id + first_offset + t] = rows[b, t] for t < n_rows (default: all of rows.shape[1])
        unless padded / overwritten."""
        col = self._columns[key]
        b_stride = rows.shape[1]
        n_rows = b_stride if n_rows is None else n_rows
        check(self._lib.asac_storage_scatter(ptr(col), self.capacity, ptr(self._store_ids), ptr(data_ids),
                                             int(data_ids.numel()), int(first_offset), int(n_rows), ptr(rows),
                                             self._row_bytes(key), int(b_stride), ptr(padding_mask),
                                             int(padding_mask.stride(0)), self._stream), 'storage_scatter')

    def check_nan(self) -> None:
        """Raises the reference's 'td_error has nan' (replay_buffer.py:418-420); synchronises."""
        if bool(self._per_state[3].item() != 0):
            self._per_state[3] = 0
            self._logger.error('td_error has nan')
            raise Exception('td_error has nan')

    # ------------------------------------------------------------------ storage access
    def get_curr_id(self) -> int:
        return self._next_id % self.capacity

    def get_storage_data(self, data_ids) -> dict[str, torch.Tensor]:
        """Rows of every key at ``data_ids`` (no residency check, replay_buffer.py:401-406)."""
        slots = self._ids_tensor(data_ids) % self.capacity
        return {k: col[slots] for k, col in self._columns.items()}

    def get_storage_data_ids(self, data_ids) -> torch.Tensor:
        return self._store_ids[self._ids_tensor(data_ids) % self.capacity]

    # ------------------------------------------------------------------ checkpoint interchange
    def _flush(self) -> None:
        if self._flush_hook is not None:
            self._flush_hook()

    @_locked
    def save(self, save_dir: Path, ckpt: int) -> None:
        """Writes the reference's files: ``<ckpt>-rb_tree.npy`` (float32[2C-1], root first) and
        ``<ckpt>-rb_storage.npz`` (replay_buffer.py:436-441,96-97,220-221)."""
        self._flush()
        self.check_nan()
        save_dir = Path(save_dir)
        np.save(save_dir.joinpath(f'{ckpt}-rb_tree.npy'), self._nodes[1:].cpu().numpy())
        cols = {'_id': self._store_ids.cpu().numpy()}
        for k, col in (self._columns or {}).items():
            cols[k] = col.cpu().numpy()
        np.savez(save_dir.joinpath(f'{ckpt}-rb_storage.npz'), **cols, p_size=self._size, p_id=self._next_id)

    @_locked
    def load(self, save_dir: Path, ckpt: int) -> None:
        """Reads the reference's files.  Everything is validated against THIS buffer before any state
        changes: node count, ``_id`` length, every column's leading dimension; float64 columns (a NumPy
        default that slips into hand-written episodes) are cast to float32, the dtype the kernels gather."""
        save_dir = Path(save_dir)
        tree_path = save_dir.joinpath(f'{ckpt}-rb_tree.npy')
        storage_path = save_dir.joinpath(f'{ckpt}-rb_storage.npz')
        tree = cols = None
        if tree_path.exists():
            tree = np.load(tree_path)
            if tree.ndim != 1 or tree.shape[0] != 2 * self.capacity - 1:
                raise ValueError(f'tree file holds {tree.shape} nodes, capacity {self.capacity} needs '
                                 f'{2 * self.capacity - 1}')
        if storage_path.exists():
            with np.load(storage_path) as saved:
                cols = {k: saved[k] for k in saved.files}
            for need in ('p_size', 'p_id', '_id'):
                if need not in cols:
                    raise ValueError(f'{storage_path.name} has no {need!r} entry')
            for k, v in cols.items():
                if k in ('p_size', 'p_id'):
                    continue
                if v.shape[:1] != (self.capacity,):
                    raise ValueError(f'{storage_path.name}: column {k!r} holds {v.shape[0] if v.ndim else 0} rows, '
                                     f'this buffer has capacity {self.capacity}')
                if v.dtype == np.float64:
                    cols[k] = v.astype(np.float32)
                elif v.dtype not in _NP_TO_TORCH:
                    raise ValueError(f'{storage_path.name}: column {k!r} has unsupported dtype {v.dtype}')
            if cols['_id'].dtype != np.int64:
                raise ValueError(f"{storage_path.name}: '_id' is {cols['_id'].dtype}, expected int64")
            if not 0 <= int(cols['p_size']) <= self.capacity or not 0 <= int(cols['p_id']) < self.max_id:
                raise ValueError(f'{storage_path.name}: p_size / p_id out of range')
        with torch.cuda.device(self.device):
            if tree is not None:
                self._nodes[1:].copy_(torch.from_numpy(tree.astype(np.float32)))
                self._nodes[0] = 0
            if cols is not None:
                self._drop_ingest()
                self._store_ids.copy_(torch.from_numpy(cols['_id']))
                self._columns = {k: torch.from_numpy(v).to(self.device).contiguous() for k, v in cols.items()
                                 if k not in ('p_size', 'p_id', '_id')}
                self._size = int(cols['p_size'])
                self._next_id = int(cols['p_id'])

    @_locked
    def clear(self) -> None:
        self._drop_ingest()
        self._size = 0
        self._next_id = 0
        self._columns = None
        self._nodes.zero_()
        self._store_ids.zero_()

    @_locked
    def copy(self, src: 'PrioritizedReplayBuffer') -> None:
        if src.capacity != self.capacity:
            raise ValueError('capacity mismatch')
        src._flush()
        self._drop_ingest()
        self._nodes.copy_(src._nodes)
        self._store_ids.copy_(src._store_ids)
        self._columns = None if src._columns is None else {k: v.to(self.device).clone()
                                                           for k, v in src._columns.items()}
        self._size, self._next_id = src._size, src._next_id

    @property
    def beta(self) -> float:
        return float(self._per_state[0].item())

    @property
    def is_full(self) -> bool:
        return self._size == self.capacity

    @property
    def size(self) -> int:
        return self._size

    @property
    def is_lg_batch_size(self) -> bool:
        return self._size > self.batch_size

    def tree_nodes(self) -> torch.Tensor:
        """float32[2C-1] in the reference's node order (root first)."""
        return self._nodes[1:]

    def close(self):
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        self._drop_ingest()
        self._columns = None
