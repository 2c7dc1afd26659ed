from asac_b200.sac_base import SAC_Base  # noqa: F401
