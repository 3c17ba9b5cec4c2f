"""Synthetic SIP spectra of the benchmark shape.

Frequencies ``f = logspace(4, -2, N)`` Hz (descending, like the bundled files), ``w = 2 pi f``.
Spectrum ``b`` draws from ``rng = default_rng(1_000_003*b + 17)``: theta_true (model-specific,
below), then the (2, N) standard-normal noise.  ``Z_true = forward(theta_true)``,
``sigma = 0.01*|Z_true|`` on both parts, ``y = Z_true + sigma*noise``, and finally the
reference normalisation (``utils.py:138-142``): divide data and errors by ``max|y|``.

theta_true:
* Cole-Cole / Dias / Shin: uniform in the central 50 % of the default box.
* Polynomial decomposition: SURVEY §8d asks for coefficients such that "chargeabilities stay O(1e-2)
  like App. C"; its literal recipe (a_i uniform in +-0.02*{1,1,0.3,0.1,0.03}) does not do that — with
  |log_tau| up to 6 it gives |m(tau)| ~ 1 per relaxation time, total chargeability of tens, |Z| up to
  2 R0, and after the reference normalisation r0 falls outside its prior box [0.9, 1.1] for most
  spectra.  Here the coefficients are the least-squares polynomial (degree poly_deg, over the tau
  grid) of a smooth positive relaxation-time distribution: a Gaussian bump in log_tau (centre in
  [-4.5, -1.5], width in [0.8, 1.6]: little chargeability slower than the lowest frequency, so that the
  normalised r0 stays inside its prior box) on a flat floor (5-30 % of the total), scaled to a total
  chargeability in [0.2, 0.8] — the range of the reference's own fits (0.55-1.34, App. C.2).  The
  constant is in the polynomial basis, so the fitted polynomial keeps that total exactly; r0 is
  uniform in [0.95, 1.05].  theta_true is the POLYNOMIAL (what forward() sees), so the truth lies
  inside the model and inside the prior box.
"""
import numpy as np

from .batch import default_bounds, tau_grid


def frequencies(N=64):
    f = np.logspace(4, -2, N)
    return f, 2 * np.pi * f


def true_box(model, poly_deg=4, n_modes=1):
    """Box theta_true is uniform in (vector models).  For 'decomp' only r0 is drawn from its entry;
    the coefficients come from ``decomp_truth``."""
    _, b = default_bounds(model, poly_deg, n_modes)
    lo, hi = b
    mid, half = 0.5 * (lo + hi), 0.25 * (hi - lo)
    lo_c, hi_c = mid - half, mid + half
    if model == 'decomp':
        lo_c[0], hi_c[0] = 0.95, 1.05
    return lo_c, hi_c


def decomp_truth(u, log_tau, poly_deg):
    """u: 5 uniforms in [0,1) -> theta_true = [r0, a0..aP] (see the module docstring)."""
    r0 = 0.95 + 0.10 * u[0]
    mu = -4.5 + 3.0 * u[1]
    width = 0.8 + 0.8 * u[2]
    m_tot = 0.2 + 0.6 * u[3]
    floor = 0.05 + 0.25 * u[4]
    S = log_tau.shape[0]
    g = np.exp(-0.5 * ((log_tau - mu) / width) ** 2)
    m = m_tot * ((1.0 - floor) * g / g.sum() + floor / S)
    V = np.vander(log_tau, poly_deg + 1, increasing=True)
    scale = np.abs(V).max(0)                                   # column scaling: powers reach 6^poly_deg
    a = np.linalg.lstsq(V / scale, m, rcond=None)[0] / scale
    return np.concatenate([[r0], a])


def draws(model, b0, b1, N=64, poly_deg=4, n_modes=1, n_tau=None):
    """theta_true (n, ndim) and noise (n, 2, N) for global spectrum indices [b0, b1)."""
    lo_c, hi_c = true_box(model, poly_deg, n_modes)
    n = b1 - b0
    theta = np.empty((n, lo_c.shape[0]))
    noise = np.empty((n, 2, N))
    log_tau = tau_grid(frequencies(N)[1], n_tau, poly_deg)[0] if model == 'decomp' else None
    for i, b in enumerate(range(b0, b1)):
        rng = np.random.default_rng(1_000_003 * b + 17)
        if model == 'decomp':
            theta[i] = decomp_truth(rng.uniform(size=5), log_tau, poly_deg)
        else:
            theta[i] = rng.uniform(lo_c, hi_c)
        noise[i] = rng.standard_normal((2, N))
    return theta, noise


def assemble(Z_true, noise, rel_err=0.01):
    """Z_true (n, 2, N), noise (n, 2, N) -> zn, zn_err (n, 2, N), norm_factor (n,)."""
    amp = np.sqrt(Z_true[:, 0] ** 2 + Z_true[:, 1] ** 2)          # (n, N)
    sigma = np.repeat((rel_err * amp)[:, None, :], 2, axis=1)
    y = Z_true + sigma * noise
    nf = np.max(np.sqrt(y[:, 0] ** 2 + y[:, 1] ** 2), axis=1)[:, None, None]
    return y / nf, sigma / nf, nf[:, 0, 0]


def make(model, b0, b1, forward, N=64, poly_deg=4, n_modes=1, n_tau=None):
    """Build spectra [b0, b1).  ``forward(theta (n, ndim), w) -> (n, 2, N)`` is supplied by the
    caller: the CUDA batched forward in the product / bench, the oracle in CPU-only tests.
    ``n_tau`` (decomposition): size of the tau grid the caller's forward uses (None = 2N)."""
    f, w = frequencies(N)
    theta, noise = draws(model, b0, b1, N, poly_deg, n_modes, n_tau)
    zn, zn_err, nf = assemble(np.asarray(forward(theta, w)), noise)
    return dict(freq=f, w=w, theta_true=theta, zn=zn, zn_err=zn_err, norm_factor=nf)
