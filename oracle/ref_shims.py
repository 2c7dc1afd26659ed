"""Import the real reference (read-only mount) in the build container.

TEST INFRASTRUCTURE, build-container only: ``/root/reference`` does not exist on
the GPU box, so nothing reachable from ``-m gpu`` tests, ``smoke()`` or
``bench.py`` calls this.  It is used by ``oracle/gen_golden.py`` (fixture
generation) and by the optional ``tests/test_oracle_vs_reference.py`` (skipped
when the mount is absent).

Three shims are needed: matplotlib is not installed;
``PrioritizedReplayBuffer.__init__`` calls ``torch.cuda.Stream()`` and
``_prefetch_loop`` calls ``.pin_memory()`` unconditionally
(replay_buffer.py:276,363-366), which both need an NVIDIA driver.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path('/root/reference')


def reference_available() -> bool:
    return (REFERENCE_ROOT / 'algorithm' / 'sac_base.py').exists()


def install_shims() -> None:
    import torch

    sys.dont_write_bytecode = True
    for name in ('matplotlib', 'matplotlib.backends', 'matplotlib.backends.backend_agg', 'matplotlib.figure'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['matplotlib.backends.backend_agg'].FigureCanvasAgg = object
    sys.modules['matplotlib.figure'].Figure = object
    if not torch.cuda.is_available():
        torch.cuda.Stream = lambda *a, **k: None
        torch.Tensor.pin_memory = lambda self, *a, **k: self
    try:  # first import of torchvision walks sys.modules with inspect and trips over a namespace package
        import torchvision  # noqa: F401  (the reference's `algorithm`): get it over with beforehand
    except Exception:  # noqa: BLE001
        pass
    try:  # no tqdm monitor thread: the fixture generators patch Thread.start while the reference starts up,
        import tqdm  # and a monitor created meanwhile would fail to join at interpreter exit
        tqdm.tqdm.monitor_interval = 0
    except Exception:  # noqa: BLE001
        pass
    root = str(REFERENCE_ROOT)
    if root not in sys.path:
        sys.path.insert(0, root)


def load_reference_nn(rel_path: str):
    """Load a plugin ``nn`` module (e.g. ``envs/test/nn.py``) the way
    ``sac_main.py:353-364`` does."""
    install_shims()
    path = REFERENCE_ROOT / rel_path
    spec = importlib.util.spec_from_file_location('ref_nn_' + path.stem + '_' + path.parent.name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    """Returns (SAC_Base, replay_buffer module, None).  ``tests/get_synthesis_data.py`` is loaded
    by file path by its users (the name ``tests`` collides with this repo's own test package)."""
    install_shims()
    # a product alias package named ``algorithm`` may already be importable; make sure the
    # reference's own package wins inside this process.
    for name in [n for n in sys.modules if n == 'algorithm' or n.startswith('algorithm.')]:
        mod = sys.modules[name]
        if not str(getattr(mod, '__file__', '')).startswith(str(REFERENCE_ROOT)):
            del sys.modules[name]
    # The reference's ``algorithm`` directory has no __init__.py (a namespace package), and a regular
    # package of the same name ANYWHERE on sys.path beats a namespace package: hide the directories that
    # hold the product's alias package while importing (e.g. under pytest).  Once imported, submodules
    # resolve through ``algorithm.__path__``, not sys.path.
    saved = list(sys.path)
    sys.path[:] = [str(REFERENCE_ROOT)] + [p for p in saved
                                            if not (Path(p or '.') / 'algorithm' / '__init__.py').exists()]
    try:
        from algorithm import replay_buffer as ref_rb
        from algorithm.sac_base import SAC_Base
    finally:
        sys.path[:] = saved
    return SAC_Base, ref_rb, None
