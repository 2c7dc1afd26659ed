"""Import alias so that plugin files written for the reference (``import algorithm.nn_models as m``,
``from algorithm.sac_base import SAC_Base``) resolve to the B200 implementation unchanged."""
