"""On-policy batch buffer with the reference's API (``algorithm/batch_buffer.py:10-95``) for
``SAC_Base(use_replay_buffer=False)``: every window of an episode is used exactly once, in shuffled
order, ``batch_size`` windows at a time.

Host-side NumPy, like the reference's: the work is slicing one episode into ``ep_len - 1`` windows
of ``burn_in_step + n_step`` rows — a few KB — after which the learner uploads one batch per
``train()`` into the same device buffers the replay path gathers into, and the same kernels run.
"""
from __future__ import annotations

import threading

import numpy as np
import torch


def episode_to_batch(burn_in_step: int, n_step: int, padding_action: np.ndarray, l_indexes: np.ndarray,
                     l_last_masks: np.ndarray, l_obses_list: list[np.ndarray], l_actions: np.ndarray,
                     l_rewards: np.ndarray, l_dones: np.ndarray, l_probs: np.ndarray,
                     l_pre_seq_hidden_states: np.ndarray):
    """utils/operators.py:105-207.  Inputs ``[1, ep_len, ...]``; the episode is padded by ``burn_in_step``
    rows in front and ``n_step - 1`` behind (index -1, last_mask / done True, reward 0, prob 1, the padding
    action, zero observations / hidden states) and window ``i`` (``i < ep_len - 1``) is rows ``i .. i + bn - 1``
    (``bn + 1`` rows for observations and hidden states).  Returns ``(bn_indexes, bn_last_masks,
    bn_padding_masks, bnx_obses_list, bn_actions, bn_rewards, bn_dones, bn_probs, bnx_pre_seq_hidden_states)``.

    Reference quirk kept on purpose: only the FRONT padding is flagged in ``bn_padding_masks`` (the mask
    there is built as ones(b) ++ zeros(len of the already padded last-mask row) ++ ones(n - 1), so no window
    reaches its trailing ones); the rows padded behind the episode are excluded by their ``last_mask``."""
    b, n = int(burn_in_step), int(n_step)
    bn = b + n
    ep_len = l_indexes.shape[1]
    n_windows = ep_len - 1

    def pad(x: np.ndarray, fill) -> np.ndarray:
        front = np.empty((b, *x.shape[2:]), dtype=x.dtype)
        back = np.empty((n - 1, *x.shape[2:]), dtype=x.dtype)
        front[...] = fill
        back[...] = fill
        return np.concatenate([front, x[0], back], axis=0)

    rows = np.arange(n_windows)[:, None] + np.arange(bn)[None, :]          # [windows, bn]
    rows_x = np.arange(n_windows)[:, None] + np.arange(bn + 1)[None, :]    # [windows, bn + 1]
    padding_action = np.asarray(padding_action).reshape(-1).astype(l_actions.dtype)
    padded_len = b + ep_len + n - 1
    padding_mask_row = np.arange(padded_len) < b
    return (pad(l_indexes, -1)[rows],
            pad(l_last_masks, True)[rows],
            padding_mask_row[rows],
            [pad(o, 0)[rows_x] for o in l_obses_list],
            pad(l_actions, padding_action)[rows],
            pad(l_rewards, 0)[rows],
            pad(l_dones, True)[rows],
            pad(l_probs, 1)[rows],
            pad(l_pre_seq_hidden_states, 0)[rows_x])


def _map(batch: list, fn) -> list:
    return [[fn(x) for x in item] if isinstance(item, list) else fn(item) for item in batch]


class BatchBuffer:
    _rest_batch = None

    def __init__(self, burn_in_step: int, n_step: int, padding_action: np.ndarray, batch_size: int,
                 device: torch.device | None = None, max_size: int = 10):
        self.burn_in_step = burn_in_step
        self.n_step = n_step
        self.padding_action = padding_action
        self.batch_size = batch_size
        self.device = device
        self.max_size = max_size
        self._lock = threading.Lock()
        self._batch_list: list[list] = []

    def put_episode(self, ep_indexes, ep_last_masks, ep_obses_list, ep_actions, ep_rewards, ep_dones, ep_probs,
                    ep_pre_seq_hidden_states) -> None:
        """batch_buffer.py:31-85: windows of the episode (plus what the last call left over), shuffled with
        ``np.random.permutation``, cut into full batches; an incomplete tail waits for the next episode; at most
        ``max_size`` batches are kept (oldest dropped)."""
        with self._lock:
            batch = list(episode_to_batch(self.burn_in_step, self.n_step, self.padding_action, ep_indexes,
                                          ep_last_masks, list(ep_obses_list), ep_actions, ep_rewards, ep_dones,
                                          ep_probs, ep_pre_seq_hidden_states))
            if self._rest_batch is not None:
                rest, self._rest_batch = self._rest_batch, None
                batch = [[np.concatenate([r, x]) for r, x in zip(ri, bi)] if isinstance(bi, list)
                         else np.concatenate([ri, bi]) for ri, bi in zip(rest, batch)]
            total = batch[0].shape[0]
            order = np.random.permutation(total)
            batch = _map(batch, lambda x: x[order])
            for start in range(0, total, self.batch_size):
                piece = _map(batch, lambda x: x[start:start + self.batch_size])
                if start + self.batch_size > total:
                    self._rest_batch = piece
                else:
                    self._batch_list.append(piece)
                    if len(self._batch_list) > self.max_size:
                        self._batch_list.pop(0)

    def get_batch(self):
        """The oldest complete batch as torch tensors (on ``device`` when it is a CUDA device), or None."""
        with self._lock:
            if not self._batch_list:
                return None
            batch = _map(self._batch_list.pop(0), torch.from_numpy)
            if self.device is not None and torch.device(self.device).type == 'cuda':
                batch = _map(batch, lambda t: t.to(self.device))
            return batch
