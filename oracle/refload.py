"""TEST INFRASTRUCTURE — loader for the UNMODIFIED reference package built by
``oracle/build_ref.sh`` into ``oracle/_ref/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module.  Nothing under ``bisip_b200/`` does.

The reference imports ``emcee``, ``matplotlib`` and ``corner`` at module top
(reference ``models.py:10``, ``plotlib.py:11-12``); none is installed in this image and
there is no network, so stubs are registered first: empty modules for matplotlib / corner,
and ``oracle/emcee_restatement.py`` for emcee (so the reference's own ``fit()`` runs).  The forward /
log-likelihood / log-prior / log-probability code that then runs is the reference's own
(``models.py:59-76``, ``cython_funcs.pyx:33-108``).  If a real ``emcee`` is ever importable it is used instead of the restatement.
"""
import importlib
import os
import sys
import types
import warnings

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    d = os.path.join(_REF_DIR, "bisip")
    return os.path.isdir(d) and any(f.startswith("cython_funcs") and f.endswith(".so")
                                    for f in os.listdir(d))


def load():
    """Return the reference ``bisip`` module (imported from oracle/_ref)."""
    if "bisip" in sys.modules and getattr(sys.modules["bisip"], "__graft_ref__", False):
        return sys.modules["bisip"]
    if not available():
        raise ImportError("oracle/_ref is not built: run oracle/build_ref.sh where "
                          "/root/reference exists")
    if "emcee" not in sys.modules:
        try:
            importlib.import_module("emcee")
        except Exception:            # not installed in this image: use the restatement
            from . import emcee_restatement
            sys.modules["emcee"] = emcee_restatement
    for name in ("matplotlib", "matplotlib.pyplot", "corner"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["corner"], "corner"):
        sys.modules["corner"].corner = lambda *a, **k: None
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, _REF_DIR)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("bisip")
    finally:
        sys.path.remove(_REF_DIR)
    mod.__graft_ref__ = True
    return mod


def data_file(name="SIP-K389175"):
    return os.path.join(_REF_DIR, "bisip", "data", name + ".dat")
