"""Host-side helpers mirroring the reference's ``algorithm/utils`` names that plugin files and callers
import: enums (+ yaml converters), distribution operators, image transforms, visualisation aids."""
from .enums import *  # noqa: F401,F403
from .operators import *  # noqa: F401,F403
