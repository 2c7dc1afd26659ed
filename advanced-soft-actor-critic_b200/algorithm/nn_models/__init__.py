from asac_b200.nn_models import *  # noqa: F401,F403
from asac_b200.nn_models import __all__  # noqa: F401
