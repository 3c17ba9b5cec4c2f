"""Bundled example spectra — mirror of reference ``data.py:13-19`` (``DataFiles``).

The six example spectra ship as one packed table (``data/examples.npz``, written by
``tests/golden/make_golden.py`` from the reference's ``data/*.dat``); ``DataFiles``
materialises them on first use as CSV files with the reference's format
(``docs/user/data_format.rst``: header line, then ``freq, amp, pha, amp_err, pha_err``
with ``%.18e`` so every value round-trips bit-exactly) and maps name -> path like the
reference does.
"""
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PACK = os.path.join(_HERE, "data", "examples.npz")
HEADER = "freq, amp, pha, amp_err, pha_err"


def example_tables():
    """dict name -> (n, 5) float64 table."""
    with np.load(_PACK) as z:
        return {k: z[k] for k in sorted(z.files)}


def _materialise():
    out = os.path.join(_HERE, "data", "_generated")
    try:
        os.makedirs(out, exist_ok=True)
        probe = os.path.join(out, ".w")
        open(probe, "w").close()
        os.remove(probe)
    except OSError:
        out = os.path.join(tempfile.gettempdir(), "bisip_b200_data")
        os.makedirs(out, exist_ok=True)
    paths = {}
    for name, tab in example_tables().items():
        fp = os.path.join(out, name + ".dat")
        if not os.path.exists(fp):
            tmp = fp + f".{os.getpid()}.tmp"
            np.savetxt(tmp, tab, fmt="%.18e", delimiter=",", header=HEADER, comments="")
            os.replace(tmp, fp)
        paths[name] = fp
    return paths


class DataFiles(dict):
    """``DataFiles()['SIP-K389175']`` -> path of a bundled example data file."""

    def __init__(self, *args, **kwargs):
        super(DataFiles, self).__init__(*args, **kwargs)
        self.update(_materialise())
