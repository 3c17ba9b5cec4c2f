"""emcee-compatible facade over the on-device ensemble sampler.

The reference leaks the whole ``emcee.EnsembleSampler`` through ``Inversion.sampler``
(reference ``models.py:154-159``) and calls ``run_mcmc`` / ``get_chain`` on it
(``models.py:118, 137``).  This class keeps that surface — ``run_mcmc``, ``get_chain``,
``get_log_prob``, ``acceptance_fraction``, ``iteration``, ``chain``, ``get_autocorr_time`` —
with emcee's storage and slicing semantics (bit-exact indexing), while
the moves themselves run in ``bisip_ensemble_run`` (CUDA, Philox RNG).  emcee is a
third-party dependency whose source is not part of the reference tree; its error behaviour
is restated from its published 3.x API.
"""
import numpy as np
import torch

from . import _lib, engine


def walkers_independent(coords):
    """emcee.ensemble.walkers_independent: condition number of the centred, scaled walkers."""
    if not np.all(np.isfinite(coords)):
        return False
    C = coords - np.mean(coords, axis=0)[None, :]
    C_colmax = np.amax(np.abs(C), axis=0)
    if np.any(C_colmax == 0):
        return False
    C /= C_colmax
    C_colsum = np.sqrt(np.sum(C ** 2, axis=0))
    C /= C_colsum
    return np.linalg.cond(C.astype(float)) <= 1e8


def slice_chain(arr, iteration, flat=False, thin=1, discard=0):
    """emcee Backend.get_value: arr[discard+thin-1 : iteration : thin], optional flatten."""
    v = arr[discard + thin - 1:iteration:thin]
    if flat:
        s = list(v.shape[1:])
        s[0] = int(np.prod(v.shape[:2]))
        return v.reshape(s)
    return v


class EnsembleSampler:
    """One spectrum, ``nwalkers`` walkers; the hot loop lives on the GPU."""

    def __init__(self, nwalkers, ndim, spec, w, y, yerr, bounds, seed=None, a=2.0, device=None,
                 spectrum_index=0):
        self.nwalkers = int(nwalkers)
        self.ndim = int(ndim)
        self.a = float(a)
        self.device = _lib.require_cuda(device)
        self.spec = spec
        self._w = _lib.dev_f64(w, self.device)
        self._y = _lib.dev_f64(y, self.device).reshape(1, 2, -1)
        self._yerr = _lib.dev_f64(yerr, self.device).reshape(1, 2, -1)
        self._bounds = _lib.dev_f64(bounds, self.device)
        if seed is None:
            # like emcee, tie the stream to NumPy's global generator so that
            # np.random.seed(...) before fit() makes the run repeatable
            seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
        self.seed = int(seed)
        self.spectrum_index = int(spectrum_index)
        self.reset()

    def reset(self):
        self.iteration = 0
        self._chain = np.empty((0, self.nwalkers, self.ndim))
        self._log_prob = np.empty((0, self.nwalkers))
        self._chain_dev = None
        self.accepted = np.zeros(self.nwalkers)
        self._last = None

    # ------------------------------------------------------------------ run
    def run_mcmc(self, initial_state, nsteps, progress=False, **kwargs):
        """emcee's ``run_mcmc(initial_state, nsteps, progress=...)``; returns the last ``(coords, log_prob)`` pair
        (emcee returns a ``State``, which unpacks to the same two leading items).  ``progress`` is accepted and
        ignored (the whole chain is one kernel launch).  emcee's other ``sample`` keywords change what is stored
        (``thin_by``, ``store``, ``tune``, ``skip_initial_state_check`` ...) and have no counterpart here: passing
        one raises instead of silently returning a chain with different semantics."""
        if kwargs:
            raise NotImplementedError('run_mcmc keyword(s) not supported by the on-device sampler: '
                                      + ', '.join(sorted(kwargs)) + ' (use get_chain(discard=, thin=) instead)')
        if initial_state is None:
            if self._last is None:
                raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
            coords = self._last[0]
        else:
            coords = np.array(initial_state, dtype=np.float64, copy=True)
        if coords.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions {0}".format(coords.shape))
        if self.nwalkers < 2 * self.ndim:
            raise RuntimeError("It is unadvisable to use a red-blue move with fewer walkers than "
                               "twice the number of dimensions.")
        if np.any(np.isinf(coords)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(coords)):
            raise ValueError("At least one parameter value was NaN")
        if not walkers_independent(coords.copy()):
            raise ValueError("Initial state has a large condition number. Make sure that your walkers "
                             "are linearly independent for the best performance")
        cdev = _lib.dev_f64(coords, self.device).reshape(1, self.nwalkers, self.ndim).clone()
        res = engine.ensemble_run(self.spec, cdev, self._w, self._y, self._yerr, self._bounds,
                                  nsteps=int(nsteps), seed=self.seed, spectrum0=self.spectrum_index,
                                  a=self.a, discard=0, thin=1, step0=self.iteration)
        flags = int(res["flags"][0].item())
        if flags & 2:
            raise ValueError("The initial log_prob was NaN")
        if flags & 1:
            raise ValueError("Probability function returned NaN")
        chain = res["chain"][0]
        self._chain_dev = chain if self._chain_dev is None else torch.cat((self._chain_dev, chain), 0)
        self._chain = np.concatenate((self._chain, chain.cpu().numpy()))
        self._log_prob = np.concatenate((self._log_prob, res["log_prob"][0].cpu().numpy()))
        self.accepted = self.accepted + res["accepted"][0].cpu().numpy()
        self.iteration += int(nsteps)
        self._last = (res["coords"][0].cpu().numpy(), res["lp"][0].cpu().numpy())
        return self._last

    # ------------------------------------------------------------------ access
    def get_chain(self, **kwargs):
        return slice_chain(self._chain, self.iteration, **kwargs)

    def get_log_prob(self, **kwargs):
        return slice_chain(self._log_prob, self.iteration, **kwargs)

    def get_chain_device(self, discard=0, thin=1):
        """(n_keep*nwalkers, ndim) flat CUDA tensor of the kept steps."""
        v = self._chain_dev[discard + thin - 1:self.iteration:thin]
        return v.reshape(-1, self.ndim).contiguous()

    def get_last_sample(self):
        return self._last

    @property
    def chain(self):
        """emcee's deprecated (nwalkers, nsteps, ndim) view."""
        return np.swapaxes(self._chain[:self.iteration], 0, 1)

    @property
    def flatchain(self):
        return self.get_chain(flat=True)

    @property
    def lnprobability(self):
        return np.swapaxes(self._log_prob[:self.iteration], 0, 1)

    @property
    def acceptance_fraction(self):
        return self.accepted / float(self.iteration)

    def get_autocorr_time(self, discard=0, thin=1, c=5, tol=50, quiet=False):
        """Integrated autocorrelation time per parameter (Sokal windowing, as emcee.autocorr)."""
        x = self.get_chain(discard=discard, thin=thin)
        tau = thin * integrated_time(x, c=c, tol=tol, quiet=quiet)
        return tau


def _next_pow_two(n):
    i = 1
    while i < n:
        i = i << 1
    return i


def function_1d(x):
    x = np.atleast_1d(x)
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    acf /= acf[0]
    return acf


def _auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return int(np.argmin(m))
    return len(taus) - 1


def integrated_time(x, c=5, tol=50, quiet=False):
    x = np.atleast_1d(x)
    if len(x.shape) == 1:
        x = x[:, np.newaxis, np.newaxis]
    if len(x.shape) == 2:
        x = x[:, :, np.newaxis]
    n_t, n_w, n_d = x.shape
    tau_est = np.empty(n_d)
    windows = np.empty(n_d, dtype=int)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += function_1d(x[:, k, d])
        f /= n_w
        taus = 2.0 * np.cumsum(f) - 1.0
        windows[d] = _auto_window(taus, c)
        tau_est[d] = taus[windows[d]]
    flag = tol * tau_est > n_t
    if np.any(flag) and not quiet:
        raise RuntimeError("The chain is shorter than {0} times the integrated autocorrelation time "
                           "for {1} parameter(s). Use this estimate with caution and run a longer "
                           "chain!\nN/{0} = {2:.0f};\ntau: {3}".format(tol, int(np.sum(flag)), n_t / tol, tau_est))
    return tau_est


def integrated_time_batch(chain, c=5, thin=1):
    """``integrated_time`` for a batch of chains at once, on whatever device ``chain`` lives on:
    chain (B, n_t, n_w, n_d) float64 tensor -> (B, n_d) tensor of integrated autocorrelation times (times ``thin``).
    Same estimator as emcee.autocorr (walker-averaged normalised ACF through an FFT, Sokal's automatic window with
    constant ``c``); used by ``BatchInversion.get_autocorr_time`` so that survey-scale batches get a convergence
    diagnostic without a Python loop over spectra.  No ``tol`` check: callers compare ``n_t`` with the result."""
    B, n_t, n_w, n_d = chain.shape
    n = _next_pow_two(n_t)
    x = chain - chain.mean(1, keepdim=True)
    f = torch.fft.rfft(x, n=2 * n, dim=1)
    acf = torch.fft.irfft(f * f.conj(), n=2 * n, dim=1)[:, :n_t]
    acf = acf / acf[:, :1]
    rho = acf.mean(2)                                              # (B, n_t, n_d)
    taus = 2.0 * torch.cumsum(rho, 1) - 1.0
    m = torch.arange(n_t, device=chain.device, dtype=chain.dtype)[None, :, None] < c * taus
    first_false = torch.argmin(m.to(torch.int8), dim=1)            # emcee: argmin(m) if any(m) else n_t - 1
    window = torch.where(m.any(1), first_false, torch.full_like(first_false, n_t - 1))
    return thin * taus.gather(1, window[:, None, :])[:, 0]
