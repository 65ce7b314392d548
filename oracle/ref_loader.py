"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (oscar-franzen/adobo).

Only usable in the build container, where the read-only checkout lives at
/root/reference.  It is used by tests/golden/make_golden.py to produce the
committed golden vectors and by the `not gpu` tests (skipped when the checkout is
absent) to pin the oracle restatement (oracle/adobo_oracle.py) to the live
reference.  Nothing in adobo_b200/ imports this file.

The reference cannot be imported as a package in this image: eleven third-party
packages are absent and adobo/preproc.py:28 imports a symbol that SciPy removed.
None of those are touched by the hot path, so the recipe is:
  1. stub the absent third-party modules with MagicMock,
  2. register a bare `adobo` package object so adobo/__init__.py is skipped,
  3. restore `np.float` (adobo/irlbpy/irlb.py:85 uses it),
  4. import the five hot-path modules from the read-only tree, unmodified.
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("ADOBO_REFERENCE_ROOT", "/root/reference")

_ABSENT = [
    'leidenalg', 'igraph', 'umap', 'statsmodels', 'statsmodels.api', 'statsmodels.nonparametric',
    'statsmodels.nonparametric.kernel_regression', 'matplotlib', 'matplotlib.pyplot',
    'matplotlib.widgets', 'mpl_toolkits', 'mpl_toolkits.axes_grid1', 'seaborn', 'datatable', 'fa2',
    'patsy', 'mplcursors', 'community',
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'adobo'))


def load():
    """Returns a namespace with the reference modules data, normalize, hvg, dr, clustering, irlbpy."""
    if not available():
        raise RuntimeError('reference checkout not present at %s' % REFERENCE_ROOT)
    for m in _ABSENT:
        if m not in sys.modules:
            try:
                importlib.import_module(m)
            except Exception:
                sys.modules[m] = MagicMock()
    if 'adobo' not in sys.modules or not hasattr(sys.modules['adobo'], '__path__'):
        pkg = types.ModuleType('adobo')
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'adobo')]
        pkg.__version__ = '0.2.70'
        sys.modules['adobo'] = pkg
    if not hasattr(np, 'float'):
        np.float = float
    ns = types.SimpleNamespace()
    for name in ('data', 'normalize', 'hvg', 'dr', 'clustering', 'irlbpy'):
        setattr(ns, name, importlib.import_module('adobo.' + name))
    return ns
