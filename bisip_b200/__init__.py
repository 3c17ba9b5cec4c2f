"""bisip_b200 — B200-native implementation of BISIP's MCMC likelihood hot path.

Drop-in for the reference package's public names (reference ``__init__.py:10-29``), plus
``ColeCole`` (alias of ``PeltonColeCole``) and the survey-scale ``BatchInversion``.
"""
from .models import Inversion
from .models import PolynomialDecomposition
from .models import PeltonColeCole
from .models import ColeCole
from .models import Dias2000
from .models import Shin2015
from .plotlib import plotlib
from .data import DataFiles
from .batch import BatchInversion
from .selftest import test_run

__version__ = "0.1.0"

__all__ = (
    'Inversion',
    'PolynomialDecomposition',
    'PeltonColeCole',
    'ColeCole',
    'Dias2000',
    'Shin2015',
    'BatchInversion',
    'plotlib',
    'test_run',
    'DataFiles',
)
