"""The reference's native surface — drop-in for ``bisip.cython_funcs`` (reference ``cython_funcs.pyx:49-108``).

The reference's only native code is four Cython ``def`` functions that ``models.py`` imports by name
(``models.py:13-16``) and calls with keyword arguments (``models.py:228-229, 267-271, 305, 345-349``):

    ColeCole_cyth(w, R0, m, lt, c)                      cython_funcs.pyx:49
    Dias2000_cyth(w, R0, m, log_tau, eta, delta)        cython_funcs.pyx:64
    Decomp_cyth(w, taus, log_taus, c_exp, R0, a)        cython_funcs.pyx:75
    Shin2015_cyth(w, R, log_Q, n)                       cython_funcs.pyx:96

This module exports the same four names with the same positional-or-keyword signatures, the same argument
checks (array arguments are typed ``ndarray[float64, ndim=...]`` buffers there: another dtype or rank raises
``ValueError``, a non-array raises ``TypeError``; scalars are C doubles: ``TypeError`` for non-numbers) and the
same result: a FRESH float64 ``(2, N)`` array, rows ``[real; imag]``.  The arithmetic runs in the CUDA forward
kernels through the C ABI (``bisip_forward``); there is no CPU fallback.

``sys.modules['bisip.cython_funcs'] = bisip_b200.cython_funcs`` makes the unmodified reference ``models.py`` run
on the GPU kernels (tests/test_gpu_cython_shim.py does exactly that).
"""
import numpy as np

from . import _lib, engine

__all__ = ['ColeCole_cyth', 'Dias2000_cyth', 'Decomp_cyth', 'Shin2015_cyth']


def _buf(name, a, ndim):
    """Cython's typed-buffer argument check: exact type family, dtype and rank."""
    if not isinstance(a, np.ndarray):
        raise TypeError(f"Argument '{name}' has incorrect type (expected numpy.ndarray, got {type(a).__name__})")
    if a.dtype != np.float64:
        raise ValueError(f"Buffer dtype mismatch, expected 'DTYPE_t' but got '{a.dtype.name}'")
    if a.ndim != ndim:
        raise ValueError(f"Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})")
    return a


def _dbl(name, v):
    try:
        return float(v)
    except (TypeError, ValueError):
        raise TypeError(f"Argument '{name}': must be real number, not {type(v).__name__}") from None


def _forward(spec, theta, w):
    dev = _lib.require_cuda()
    th = _lib.dev_f64(np.asarray(theta, dtype=np.float64).reshape(1, 1, -1), dev)
    return engine.forward(spec, th, _lib.dev_const(w, dev))[0, 0].cpu().numpy()


def ColeCole_cyth(w, R0, m, lt, c):
    """Pelton Cole-Cole impedance: ``R0*(1 - sum_i m_i*(1 - 1/(1 + (1j*w*exp(lt_i))**c_i)))``.
    The number of modes is ``m.shape[0]``, as in the reference (``lt`` and ``c`` are read up to that length)."""
    w, m, lt, c = _buf('w', w, 1), _buf('m', m, 1), _buf('lt', lt, 1), _buf('c', c, 1)
    R0 = _dbl('R0', R0)
    K = m.shape[0]
    if K == 0:                                   # empty sum: Z = R0 (1 - 0)
        return np.array([np.full(w.shape[0], R0), np.zeros(w.shape[0])])
    if lt.shape[0] < K or c.shape[0] < K:
        raise ValueError('lt and c must hold one entry per mode (m.shape[0])')
    theta = np.concatenate([[R0], m, lt[:K], c[:K]])
    return _forward(engine.ModelSpec(model=_lib.MODEL_COLECOLE, ndim=1 + 3 * K, n_modes=K), theta, w)


def Dias2000_cyth(w, R0, m, log_tau, eta, delta):
    """Dias (2000) impedance (``cython_funcs.pyx:36-40``)."""
    w = _buf('w', w, 1)
    theta = [_dbl('R0', R0), _dbl('m', m), _dbl('log_tau', log_tau), _dbl('eta', eta), _dbl('delta', delta)]
    return _forward(engine.ModelSpec(model=_lib.MODEL_DIAS, ndim=5), theta, w)


def Decomp_cyth(w, taus, log_taus, c_exp, R0, a):
    """Debye / Warburg polynomial decomposition: ``M_k = sum_i a_i log_taus[i,k]``,
    ``Z = R0*(1 - sum_k M_k*(1 - 1/(1 + (1j*w*taus_k)**c_exp)))`` (``cython_funcs.pyx:75-94``)."""
    w, taus, log_taus, a = _buf('w', w, 1), _buf('taus', taus, 1), _buf('log_taus', log_taus, 2), _buf('a', a, 1)
    c_exp, R0 = _dbl('c_exp', c_exp), _dbl('R0', R0)
    D = a.shape[0]
    if log_taus.shape[0] < D or log_taus.shape[1] < taus.shape[0]:
        raise ValueError('log_taus must be (a.shape[0], taus.shape[0])')
    dev = _lib.require_cuda()
    spec = engine.ModelSpec(model=_lib.MODEL_DECOMP, ndim=1 + D, taus=_lib.dev_const(taus, dev),
                            log_taus=_lib.dev_const(log_taus[:D, :taus.shape[0]], dev), c_exp=c_exp,
                            precision=_lib.PREC_FP64)
    return _forward(spec, np.concatenate([[R0], a]), w)


def Shin2015_cyth(w, R, log_Q, n):
    """Shin (2015): ``sum_i (1/z_cpe_i + 1/R_i)**-1``, ``z_cpe_i = 1/(exp(log_Q_i)*(1j*w)**n_i)`` over
    ``R.shape[0]`` elements (the model class always passes two; ``cython_funcs.pyx:96-108``).  The CUDA kernel
    evaluates pairs: other counts are summed pair by pair, a last single element as half of a doubled pair
    (halving is exact)."""
    w, R, log_Q, n = _buf('w', w, 1), _buf('R', R, 1), _buf('log_Q', log_Q, 1), _buf('n', n, 1)
    D = R.shape[0]
    if log_Q.shape[0] < D or n.shape[0] < D:
        raise ValueError('log_Q and n must hold one entry per element (R.shape[0])')
    spec = engine.ModelSpec(model=_lib.MODEL_SHIN, ndim=6)
    Z = np.zeros((2, w.shape[0]))
    for i in range(0, D, 2):
        j = i + 1 if i + 1 < D else i
        z = _forward(spec, [R[i], R[j], log_Q[i], log_Q[j], n[i], n[j]], w)
        Z = Z + (z if j != i else 0.5 * z)
    return Z
