"""TEST INFRASTRUCTURE — ctypes front-end of the C restatement ``oracle/bisip_oracle.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module; the product package ``bisip_b200`` never does.

``Problem`` mirrors the argument tuple the reference hands to emcee
(``model_args = (forward, bounds, w, zn, zn_err)``, reference ``models.py:108-109``) plus the
per-model constants (``taus``, ``log_taus``, ``c_exp``: ``models.py:197-209``; ``n_modes``:
``models.py:245``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "bisip_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libbisip_oracle.so")

MODEL_IDS = {"colecole": 0, "dias": 1, "shin": 2, "decomp": 3}


def build(force=False):
    """gcc the restatement into oracle/_build/ (git-ignored *.so; travels to the GPU box)."""
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC)):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
                           _SRC, "-o", _LIB, "-lm"])
    return _LIB


class _CProblem(C.Structure):
    _fields_ = [("model", C.c_int), ("ndim", C.c_int), ("N", C.c_int), ("n_modes", C.c_int),
                ("S", C.c_int), ("D", C.c_int), ("c_exp", C.c_double),
                ("w", C.c_void_p), ("taus", C.c_void_p), ("log_taus", C.c_void_p),
                ("y", C.c_void_p), ("yerr", C.c_void_p), ("bounds", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_log_probability.restype = C.c_double
        _lib.oracle_ensemble_run.restype = C.c_int
    return _lib


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Problem:
    """One spectrum + one model = everything the log-probability needs."""

    def __init__(self, model, w, y, yerr, bounds, n_modes=1, taus=None, log_taus=None, c_exp=1.0):
        self.model = model
        self.w = _f64(w)
        self.N = self.w.shape[0]
        self.y = _f64(y).reshape(2, self.N)
        self.yerr = _f64(yerr).reshape(2, self.N)
        self.bounds = _f64(bounds)
        self.ndim = self.bounds.shape[1]
        self.taus = _f64(taus if taus is not None else np.zeros(1))
        self.log_taus = _f64(log_taus if log_taus is not None else np.zeros((1, 1)))
        self.c = _CProblem(MODEL_IDS[model], self.ndim, self.N, int(n_modes),
                           self.taus.shape[0], self.log_taus.shape[0], float(c_exp),
                           _ptr(self.w), _ptr(self.taus), _ptr(self.log_taus),
                           _ptr(self.y), _ptr(self.yerr), _ptr(self.bounds))

    # -- forward / log-prob ------------------------------------------------------------
    def forward(self, theta):
        theta = _f64(theta)
        single = theta.ndim == 1
        th = theta.reshape(-1, self.ndim)
        Z = np.empty((th.shape[0], 2, self.N))
        lib().oracle_forward_many(C.byref(self.c), _ptr(th), C.c_int(th.shape[0]), _ptr(Z))
        return Z[0] if single else Z

    def log_probability(self, theta):
        theta = _f64(theta)
        single = theta.ndim == 1
        th = theta.reshape(-1, self.ndim)
        out = np.empty(th.shape[0])
        lib().oracle_log_probability_many(C.byref(self.c), _ptr(th), C.c_int(th.shape[0]), _ptr(out))
        return out[0] if single else out

    # -- sampler (Philox stream identical to the CUDA kernel) ----------------------------
    def run(self, p0, nsteps, seed=0, spectrum=0, a=2.0, discard=0, thin=1, step0=0):
        coords = _f64(p0).copy()
        W = coords.shape[0]
        first = discard + thin - 1
        nkeep = 0 if nsteps <= first else (nsteps - first + thin - 1) // thin
        chain = np.empty((nkeep, W, self.ndim))
        logp = np.empty((nkeep, W))
        acc = np.zeros(W, dtype=np.int32)
        lpf = np.empty(W)
        nan = lib().oracle_ensemble_run(C.byref(self.c), _ptr(coords), C.c_int(W), C.c_int(nsteps),
                                        C.c_int(step0), C.c_uint64(seed), C.c_uint32(spectrum),
                                        C.c_double(a), C.c_int(discard), C.c_int(thin),
                                        _ptr(chain), _ptr(logp), _ptr(acc), _ptr(lpf))
        return dict(chain=chain, log_prob=logp, accepted=acc, coords=coords, lp=lpf, nan=bool(nan))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)
