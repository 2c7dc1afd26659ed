"""Step records and the abstract multi-agent environment interface
(reference: algorithm/env_wrapper/env_wrapper.py:6-102).  ``ma_*`` = dict keyed by agent-group name."""
from __future__ import annotations

from pathlib import Path

import numpy as np

__all__ = ['DecisionStep', 'TerminalStep', 'EnvWrapper']


class DecisionStep:
    """Agents that need an action: ids ``(NAgents,)``, observations ``[(NAgents, *shape), ...]``, the reward of
    the previous step, and the dataset's action for offline environments."""

    def __init__(self, ma_agent_ids: dict[str, np.ndarray], ma_obs_list: dict[str, list[np.ndarray]],
                 ma_last_reward: dict[str, np.ndarray], ma_offline_action: dict[str, np.ndarray] | None = None):
        vars(self).update(ma_agent_ids=ma_agent_ids, ma_obs_list=ma_obs_list, ma_last_reward=ma_last_reward,
                          ma_offline_action=ma_offline_action)


class TerminalStep(DecisionStep):
    """Agents whose episode ended; ``ma_max_reached`` marks time-limit terminations."""

    def __init__(self, ma_agent_ids, ma_obs_list, ma_last_reward, ma_max_reached: dict[str, np.ndarray],
                 ma_offline_action=None):
        super().__init__(ma_agent_ids, ma_obs_list, ma_last_reward, ma_offline_action)
        self.ma_max_reached = ma_max_reached


class EnvWrapper:
    def __init__(self, train_mode: bool = True, env_name: str = None, env_args: dict | None = None,
                 n_envs: int = 1, model_abs_dir: Path | None = None):
        self.train_mode = train_mode
        self.env_name = env_name
        self.env_args = {} if env_args is None else env_args
        self.n_envs = n_envs
        self.model_abs_dir = model_abs_dir

    def init(self):
        """-> (ma_obs_names, ma_obs_shapes, ma_obs_dtypes, ma_d_action_sizes, ma_c_action_size)."""
        raise NotImplementedError()

    def reset(self, reset_config: dict | None = None):
        """-> (ma_agent_ids, ma_obs_list[, ma_offline_action])."""
        raise NotImplementedError()

    def step(self, ma_d_action: dict[str, np.ndarray], ma_c_action: dict[str, np.ndarray]):
        """one-hot discrete / continuous actions per group -> (DecisionStep, TerminalStep, all_envs_done)."""
        raise NotImplementedError()

    def close(self):
        raise NotImplementedError()

    def send_option(self, option: dict[str, int]):
        pass
