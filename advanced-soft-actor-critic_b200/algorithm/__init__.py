"""Import alias: code written for the reference (``import algorithm.nn_models as m``,
``from algorithm.nn_models.layers.seq_layers import GATE``, ``from algorithm.sac_base import SAC_Base``,
``from algorithm.utils.visualization.image import ImageVisual`` ...) resolves to the B200 implementation.

``algorithm.<x>`` IS ``asac_b200.<x>`` — the same module object under a second name, installed by the
finder below — so ``isinstance`` checks in the learner's lowering see the plugin's classes as its own,
and there is one place per class, not a tree of re-export stubs."""
import importlib
import importlib.abc
import importlib.util
import sys

_TARGET = 'asac_b200'
_ASAC_ALIAS = True  # the finder only serves THIS package (a test may swap the reference's own `algorithm` in)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target: str):
        self._target = target

    def create_module(self, spec):
        return importlib.import_module(self._target)

    def exec_module(self, module):  # already executed under its own name
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(__name__ + '.') or not getattr(sys.modules.get(__name__), '_ASAC_ALIAS', False):
            return None
        real = _TARGET + fullname[len(__name__):]
        try:
            real_spec = importlib.util.find_spec(real)
        except ModuleNotFoundError:
            return None
        if real_spec is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real),
                                               is_package=real_spec.submodule_search_locations is not None)
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
