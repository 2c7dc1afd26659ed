"""pytest configuration: registers the ``gpu`` marker and puts the product package
directory (``advanced-soft-actor-critic_b200/``) and the repo root on ``sys.path``."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / 'advanced-soft-actor-critic_b200'
for p in (str(ROOT), str(PKG_DIR)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
