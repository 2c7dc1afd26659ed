from asac_b200.replay_buffer import PrioritizedReplayBuffer  # noqa: F401
