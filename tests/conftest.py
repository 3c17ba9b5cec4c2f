import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gold_fl():
    with np.load(os.path.join(GOLD, "forward_logprob.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def gold_ld():
    with np.load(os.path.join(GOLD, "load_data.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def gold_post():
    with np.load(os.path.join(GOLD, "posterior.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def data_files():
    from bisip_b200 import DataFiles
    return DataFiles()


@pytest.fixture(scope="session")
def golden_dir():
    import pathlib
    return pathlib.Path(GOLD)
