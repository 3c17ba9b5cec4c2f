"""Synthetic SIP spectra of the benchmark shape.

Frequencies ``f = logspace(4, -2, N)`` Hz (descending, like the bundled files), ``w = 2 pi f``.
Spectrum ``b`` draws from ``rng = default_rng(1_000_003*b + 17)``: theta_true (model-specific
box below), then the (2, N) standard-normal noise.  ``Z_true = forward(theta_true)``,
``sigma = 0.01*|Z_true|`` on both parts, ``y = Z_true + sigma*noise``, and finally the
reference normalisation (``utils.py:138-142``): divide data and errors by ``max|y|``.
"""
import numpy as np

from .batch import default_bounds, tau_grid

# central 50 % of the default boxes; PD coefficients scaled so chargeabilities stay O(1e-2)
_PD_SCALE = np.array([1.0, 1.0, 0.3, 0.1, 0.03, 0.01, 0.003, 0.001])


def frequencies(N=64):
    f = np.logspace(4, -2, N)
    return f, 2 * np.pi * f


def true_box(model, poly_deg=4, n_modes=1):
    _, b = default_bounds(model, poly_deg, n_modes)
    lo, hi = b
    mid, half = 0.5 * (lo + hi), 0.25 * (hi - lo)
    lo_c, hi_c = mid - half, mid + half
    if model == 'decomp':
        lo_c[0], hi_c[0] = 0.95, 1.05
        lo_c[1:] = -0.02 * _PD_SCALE[:poly_deg + 1]
        hi_c[1:] = 0.02 * _PD_SCALE[:poly_deg + 1]
    return lo_c, hi_c


def draws(model, b0, b1, N=64, poly_deg=4, n_modes=1):
    """theta_true (n, ndim) and noise (n, 2, N) for global spectrum indices [b0, b1)."""
    lo_c, hi_c = true_box(model, poly_deg, n_modes)
    n = b1 - b0
    theta = np.empty((n, lo_c.shape[0]))
    noise = np.empty((n, 2, N))
    for i, b in enumerate(range(b0, b1)):
        rng = np.random.default_rng(1_000_003 * b + 17)
        theta[i] = rng.uniform(lo_c, hi_c)
        noise[i] = rng.standard_normal((2, N))
    return theta, noise


def assemble(Z_true, noise, rel_err=0.01):
    """Z_true (n, 2, N), noise (n, 2, N) -> zn, zn_err (n, 2, N)."""
    amp = np.sqrt(Z_true[:, 0] ** 2 + Z_true[:, 1] ** 2)          # (n, N)
    sigma = np.repeat((rel_err * amp)[:, None, :], 2, axis=1)
    y = Z_true + sigma * noise
    nf = np.max(np.sqrt(y[:, 0] ** 2 + y[:, 1] ** 2), axis=1)[:, None, None]
    return y / nf, sigma / nf


def make(model, b0, b1, forward, N=64, poly_deg=4, n_modes=1, n_tau=None):
    """Build spectra [b0, b1).  ``forward(theta (n, ndim), w) -> (n, 2, N)`` is supplied by the
    caller: the CUDA batched forward in the product / bench, the oracle in CPU-only tests."""
    f, w = frequencies(N)
    theta, noise = draws(model, b0, b1, N, poly_deg, n_modes)
    zn, zn_err = assemble(np.asarray(forward(theta, w)), noise)
    return dict(freq=f, w=w, theta_true=theta, zn=zn, zn_err=zn_err)
