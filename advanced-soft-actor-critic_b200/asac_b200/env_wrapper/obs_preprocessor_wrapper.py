"""Pass-through environment decorator that observation preprocessors subclass
(reference: algorithm/env_wrapper/obs_preprocessor_wrapper.py:9-60)."""
from __future__ import annotations

from .env_wrapper import DecisionStep, EnvWrapper, TerminalStep  # noqa: F401

__all__ = ['ObsPreprocessorWrapper']


class ObsPreprocessorWrapper(EnvWrapper):
    def __init__(self, env: EnvWrapper):
        self._env = env

    def init(self):
        return self._env.init()

    def reset(self, reset_config: dict | None = None):
        return self._env.reset(reset_config)

    def step(self, ma_d_action, ma_c_action):
        return self._env.step(ma_d_action, ma_c_action)

    def close(self):
        self._env.close()

    def send_option(self, option):
        self._env.send_option(option)
