"""Survey-scale inversion: many independent spectra, one CTA each, sharded one block per GPU.

The reference handles batches as a serial Python loop over files
(``docs/tutorials/decomposition.ipynb`` cells 9-10: ``for fp in files: model = ...; model.fit()``).
``BatchInversion`` keeps the per-spectrum semantics of ``Inversion.fit`` +
``get_param_percentile / mean / std`` (reference ``models.py:84-119``, ``utils.py:37-85``) but
runs all spectra of a shard in one ``bisip_ensemble_run`` launch and reduces the kept chain to
summaries on the device (``bisip_column_stats``); chains never visit the host unless asked for.

Multi-GPU: spectra are independent, so rank ``r`` of ``G`` owns the contiguous block
``shard_range(B, r, G)`` and no collective runs during sampling.  The Philox counter carries
the *global* spectrum index, so results do not depend on ``G`` or on the sub-batch size.
``gather`` all-gathers the per-spectrum summaries (NCCL on GPU tensors; gloo on CPU tensors
in the unit tests).
"""
import numpy as np
import torch

from . import _lib, engine

_MODEL_IDS = {'decomp': _lib.MODEL_DECOMP, 'colecole': _lib.MODEL_COLECOLE, 'dias': _lib.MODEL_DIAS,
              'shin': _lib.MODEL_SHIN}


def default_bounds(model, poly_deg=5, n_modes=1):
    """Default parameter boxes of the reference models (``models.py:212-213, 249-252, 287-291,
    325-331``) as ((names), (2, ndim) array)."""
    if model == 'decomp':
        names = ['r0'] + [f'a{i}' for i in range(poly_deg + 1)]
        box = [[0.9, 1.1]] + [[-1, 1]] * (poly_deg + 1)
    elif model == 'colecole':
        names = (['r0'] + [f'm{i+1}' for i in range(n_modes)] + [f'log_tau{i+1}' for i in range(n_modes)]
                 + [f'c{i+1}' for i in range(n_modes)])
        box = [[0.9, 1.1]] + [[0.0, 1.0]] * n_modes + [[-15, 5]] * n_modes + [[0.0, 1.0]] * n_modes
    elif model == 'dias':
        names = ['r0', 'm', 'log_tau', 'eta', 'delta']
        box = [[0.9, 1.1], [0, 1], [-20, 0], [0, 150], [0, 1]]
    elif model == 'shin':
        names = ['R1', 'R2', 'log_Q1', 'log_Q2', 'n1', 'n2']
        box = [[0.0, 1.0], [0.0, 1.0], [-15, -13], [-7, -5], [0, 1], [0, 1]]
    else:
        raise ValueError(f'unknown model {model!r}')
    return names, np.array(box, dtype=np.float64).T


def tau_grid(w, n_tau=None, poly_deg=5):
    """Relaxation-time grid and power table of PolynomialDecomposition (reference
    ``models.py:201-209``) for one frequency vector: (log_tau, taus, log_taus)."""
    w = np.asarray(w, dtype=np.float64)
    lo = np.floor(min(np.log10(1. / w)) - 1)
    hi = np.floor(max(np.log10(1. / w)) + 1)
    log_tau = np.linspace(lo, hi, 2 * len(w) if n_tau is None else int(n_tau))
    log_taus = np.array([log_tau ** i for i in range(poly_deg + 1)])
    return log_tau, 10 ** log_tau, log_taus


def shard_range(n, rank, world):
    """Contiguous block of spectrum indices owned by ``rank``: [lo, hi)."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather(local, n_total, rank, world, group=None):
    """All-gather per-spectrum result tensors (dict name -> tensor with leading dim = local
    shard size) into full-length tensors ordered by global spectrum index."""
    import torch.distributed as dist
    per = -(-n_total // world)
    out = {}
    for key in sorted(local):
        t = local[key]
        pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[key] = torch.cat(parts, 0)[:n_total]
    return out


class BatchInversion:
    """Invert ``B`` spectra that share a model (and its bounds) on one GPU.

    Args:
        model (str): 'decomp' (PolynomialDecomposition), 'colecole', 'dias' or 'shin'.
        w: angular frequencies, (N,) shared by all spectra or (B, N).
        zn, zn_err: normalised data and errors, (B, 2, N), rows [real; imag]
            (the ``data['zn']`` / ``data['zn_err']`` arrays of the reference, ``utils.py:141-142``).
        nwalkers, nsteps: as ``Inversion``.
        bounds: optional (2, ndim) override of the default parameter boxes.
        poly_deg, c_exp, n_tau, precision: PolynomialDecomposition options.
        n_modes: Cole-Cole modes.
        seed: Philox key.  spectrum_offset: global index of the first spectrum (for shards).
    """

    def __init__(self, model, w, zn, zn_err, nwalkers=32, nsteps=5000, bounds=None, poly_deg=5, c_exp=1.0,
                 n_tau=None, precision='fp64', n_modes=1, seed=0, spectrum_offset=0, a=2.0, device=None):
        if model not in _MODEL_IDS:
            raise ValueError(f'unknown model {model!r}')
        self.model = model
        self.nwalkers, self.nsteps = int(nwalkers), int(nsteps)
        self.seed, self.spectrum_offset, self.a = int(seed), int(spectrum_offset), float(a)
        self.device = _lib.require_cuda(device)
        self.param_names, dflt = default_bounds(model, poly_deg, n_modes)
        self.param_bounds = dflt if bounds is None else np.asarray(bounds, dtype=np.float64)
        self.ndim = self.param_bounds.shape[1]
        self.w = np.asarray(w, dtype=np.float64)
        self._zn, self._zn_err = zn, zn_err
        self.n_spectra = int(zn.shape[0])
        self.poly_deg, self.c_exp, self.n_modes, self.precision = poly_deg, c_exp, n_modes, precision
        taus = log_taus = None
        if model == 'decomp':
            if self.w.ndim == 1:
                self.log_tau, taus, log_taus = tau_grid(self.w, n_tau, poly_deg)
            else:
                grids = [tau_grid(wb, n_tau, poly_deg) for wb in self.w]
                self.log_tau = np.stack([g[0] for g in grids])
                taus = np.stack([g[1] for g in grids])
                log_taus = np.stack([g[2] for g in grids])
        self.taus, self.log_taus = taus, log_taus
        self.results = None

    # ------------------------------------------------------------------ helpers
    def _spec(self, lo=None, hi=None):
        dev = self.device
        taus = log_taus = None
        if self.taus is not None:
            per_spectrum = self.taus.ndim == 2
            taus = _lib.dev_f64(self.taus[lo:hi] if per_spectrum else self.taus, dev)
            log_taus = _lib.dev_f64(self.log_taus[lo:hi] if per_spectrum else self.log_taus, dev)
        return engine.ModelSpec(model=_MODEL_IDS[self.model], ndim=self.ndim, n_modes=self.n_modes, taus=taus,
                                log_taus=log_taus, c_exp=float(self.c_exp), precision=_lib.PRECISIONS[self.precision])

    def _to_dev(self, a, lo, hi):
        if isinstance(a, torch.Tensor):
            return a[lo:hi].to(self.device, dtype=torch.float64, non_blocking=True).contiguous()
        return _lib.dev_f64(a[lo:hi], self.device)

    def draw_p0(self, lo, hi):
        """Uniform starting positions inside the bounds, like ``Inversion.fit`` does with
        ``np.random.uniform`` (reference ``models.py:104-106``), drawn per spectrum from
        ``default_rng(seed + global_index)`` so they do not depend on sharding."""
        out = np.empty((hi - lo, self.nwalkers, self.ndim))
        lob, hib = self.param_bounds
        for i in range(lo, hi):
            rng = np.random.default_rng([self.seed, self.spectrum_offset + i])
            out[i - lo] = rng.uniform(lob, hib, (self.nwalkers, self.ndim))
        return out

    def max_batch(self, n_keep, keep_chain):
        free, _ = torch.cuda.mem_get_info(self.device)
        # kept chain (+ log-prob when the chain is kept); the statistics read it in place: no workspace
        per = n_keep * self.nwalkers * (self.ndim + (1 if keep_chain else 0)) * 8 + self.nwalkers * self.ndim * 8 * 4 + 4096
        return max(1, min(65535, int(0.7 * free // per)))

    # ------------------------------------------------------------------ fit
    def fit(self, p0=None, discard=0, thin=1, percentiles=(2.5, 50, 97.5), keep_chain=False, batch_size=None):
        """Sample every spectrum and summarise the kept chain on the device.

        Returns (and stores in ``self.results``) a dict of host arrays:
        ``percentiles`` (B, len(p), ndim), ``mean`` (B, ndim), ``std`` (B, ndim),
        ``acceptance_fraction`` (B,), ``flags`` (B,), and with ``keep_chain`` the kept
        ``chain`` (B, n_keep, W, ndim) and ``log_prob`` (B, n_keep, W).
        """
        dev_res = self.fit_device(p0, discard, thin, percentiles, keep_chain, batch_size)
        self.percentiles = tuple(float(q) for q in percentiles)
        self.results = {k: v.cpu().numpy() for k, v in dev_res.items()}
        return self.results

    def fit_device(self, p0=None, discard=0, thin=1, percentiles=(2.5, 50, 97.5), keep_chain=False,
                   batch_size=None):
        """Same as ``fit`` but leaves the summaries on the GPU (for the NCCL gather)."""
        B, W, ndim = self.n_spectra, self.nwalkers, self.ndim
        nk = engine.n_keep(self.nsteps, discard, thin)
        if nk == 0:
            raise ValueError('discard/thin leave no samples')
        bs = int(batch_size) if batch_size else self.max_batch(nk, keep_chain)
        bounds = _lib.dev_f64(self.param_bounds, self.device)
        shared_w = self.w.ndim == 1
        w_all = _lib.dev_f64(self.w, self.device) if shared_w else None
        acc = {k: [] for k in ('percentiles', 'mean', 'std', 'acceptance_fraction', 'flags')}
        if keep_chain:
            acc['chain'], acc['log_prob'] = [], []
        for lo in range(0, B, bs):
            hi = min(B, lo + bs)
            y = self._to_dev(self._zn, lo, hi)
            ye = self._to_dev(self._zn_err, lo, hi)
            w = w_all if shared_w else _lib.dev_f64(self.w[lo:hi], self.device)
            coords = self._to_dev(p0, lo, hi).clone() if p0 is not None else _lib.dev_f64(self.draw_p0(lo, hi), self.device)
            res = engine.ensemble_run(self._spec(lo, hi), coords, w, y, ye, bounds, nsteps=self.nsteps, seed=self.seed,
                                      spectrum0=self.spectrum_offset + lo, a=self.a, discard=discard, thin=thin,
                                      store_chain=True, store_logp=keep_chain)
            st = engine.column_stats(res['chain'].reshape(hi - lo, nk * W, ndim), p=list(percentiles),
                                     want_mean=True, want_std=True)
            acc['percentiles'].append(st['pct'])
            acc['mean'].append(st['mean'])
            acc['std'].append(st['std'])
            acc['acceptance_fraction'].append(res['accepted'].to(torch.float64).mean(1) / float(self.nsteps))
            acc['flags'].append(res['flags'])
            if keep_chain:
                acc['chain'].append(res['chain'])
                acc['log_prob'].append(res['log_prob'])
        return {k: torch.cat(v, 0) for k, v in acc.items()}

    # ------------------------------------------------------------------ results
    def rtd(self, stat='mean'):
        """Relaxation-time distribution of every spectrum from the fitted coefficients:
        ``stat`` = 'mean' or an index into the stored percentiles.  Returns (m (B, n_tau), total_m (B,))."""
        from .products import relaxation_time_distribution
        if self.model != 'decomp':
            raise ValueError('rtd() is defined for the polynomial decomposition only')
        if self.results is None:
            raise AssertionError('Model is not fitted! Fit the model to a dataset before attempting to plot results.')
        a = self.results['mean'] if stat == 'mean' else self.results['percentiles'][:, int(stat)]
        lt = self.log_taus
        m = (relaxation_time_distribution(a[:, 1:], lt) if lt.ndim == 2 else
             np.stack([relaxation_time_distribution(a[b, 1:], lt[b]) for b in range(a.shape[0])]))
        return m, m.sum(-1)

    def get_autocorr_time(self, c=5, thin=1):
        """Integrated autocorrelation time of every parameter of every spectrum, (B, ndim), from the kept chain of
        the last ``fit(..., keep_chain=True)`` (emcee's estimator; ``thin`` = the thinning that fit used, so the
        result is in steps).  Evaluated on the GPU in sub-batches."""
        from .sampler import integrated_time_batch
        if self.results is None or 'chain' not in self.results:
            raise AssertionError('get_autocorr_time needs fit(..., keep_chain=True)')
        ch = self.results['chain']
        out = np.empty((ch.shape[0], ch.shape[-1]))
        step = max(1, int(2 ** 27 // max(1, ch[0].size)))            # ~1 GB of float64 per sub-batch
        for lo in range(0, ch.shape[0], step):
            t = torch.from_numpy(ch[lo:lo + step]).to(self.device)
            out[lo:lo + step] = integrated_time_batch(t, c=c, thin=thin).cpu().numpy()
        return out

    def to_csv(self, path, ids=None):
        """One row per spectrum: id, acceptance, flags, mean/std/percentiles of every parameter."""
        from .products import batch_table
        if self.results is None:
            raise AssertionError('Model is not fitted! Fit the model to a dataset before attempting to plot results.')
        cols, table = batch_table(self.param_names, self.results, ids, self.percentiles)
        np.savetxt(path, table, header=','.join(cols), delimiter=',', comments='')
        return cols

    # ------------------------------------------------------------------ construction from files
    @classmethod
    def from_files(cls, model, filepaths, headers=1, ph_units='mrad', **kwargs):
        """Vectorised ingest of many data files that share a frequency grid or not
        (reference ``utils.py:108-146`` applied per file)."""
        from .utils import prepare_data
        data = [prepare_data(np.loadtxt(fp, skiprows=headers, delimiter=','), ph_units) for fp in filepaths]
        ws = np.stack([d['w'] for d in data])
        w = ws[0] if np.all(ws == ws[0]) else ws
        inv = cls(model, w, np.stack([d['zn'] for d in data]), np.stack([d['zn_err'] for d in data]), **kwargs)
        inv.data = data
        return inv
