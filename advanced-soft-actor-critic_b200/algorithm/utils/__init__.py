from asac_b200.config_enums import SEQ_ENCODER  # noqa: F401
