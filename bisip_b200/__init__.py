"""bisip_b200 — B200-native implementation of BISIP's MCMC likelihood hot path.

Exposes the reference package's public names (reference ``__init__.py:10-29``): the model classes,
the ``plotlib`` mixin, ``DataFiles`` and ``test_run`` — plus ``ColeCole`` (alias of
``PeltonColeCole``), the survey-scale ``BatchInversion`` and the ``products`` helpers.
"""
from . import cython_funcs, products
from .batch import BatchInversion, fit_sharded
from .data import DataFiles
from .models import (ColeCole, Dias2000, Inversion, PeltonColeCole, PolynomialDecomposition,
                     Shin2015)
from .plotlib import plotlib
from .selftest import test_run

__version__ = "0.1.0"

_REFERENCE_NAMES = ['Inversion', 'PolynomialDecomposition', 'PeltonColeCole', 'Dias2000', 'Shin2015',
                    'plotlib', 'test_run', 'DataFiles']
__all__ = tuple(_REFERENCE_NAMES + ['ColeCole', 'BatchInversion', 'fit_sharded', 'products', 'cython_funcs'])
