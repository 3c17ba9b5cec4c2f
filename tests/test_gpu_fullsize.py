"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run
5e9 log-prob evaluations; these can be checked at any size):

* calibration — theta_true of every synthetic spectrum is a draw from the model, so it must fall inside the
  central 95 % posterior interval for 95 % of the spectra and (mean - truth)/std must be ~N(0,1).  This
  exercises the forward model, the likelihood, the prior, the stretch move, the chain layout, discard / thin
  and the percentile kernel together; any bias in one of them moves the coverage far from 0.95.
* shard independence — spectra taken out of the middle of the batch and run alone with their global
  index give bit-identical summaries (what multi-GPU sharding relies on).
* same-stream oracle parity on one spectrum of the big batch (chain prefix).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(model, B, **kw):
    import torch
    from bisip_b200 import _lib, engine, synthetic
    from bisip_b200.batch import BatchInversion
    dev = torch.device("cuda:0")
    _, w = synthetic.frequencies(64)
    probe = BatchInversion(model, w, np.zeros((1, 2, 64)), np.ones((1, 2, 64)), device=dev, **kw)

    def fwd(th, ww):
        return engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
    syn = synthetic.make(model, 0, B, fwd, N=64, poly_deg=kw.get('poly_deg', 4), n_tau=kw.get('n_tau'))
    truth = syn['theta_true'].copy()
    nres = 2 if model == 'shin' else 1                 # the resistance-like parameters scale with 1/norm_factor
    truth[:, :nres] /= syn['norm_factor'][:, None]
    return w, syn, truth


def _calibration(res, truth):
    lo, hi = res['percentiles'][:, 0], res['percentiles'][:, 2]
    cov = ((lo <= truth) & (truth <= hi)).mean(0)
    z = (res['mean'] - truth) / res['std']
    return cov, z.mean(0), z.std(0)


def _subset_identical(model, w, syn, res, lo, hi, fit_kw, **ctor_kw):
    from bisip_b200.batch import BatchInversion
    part = BatchInversion(model, w, syn['zn'][lo:hi], syn['zn_err'][lo:hi], spectrum_offset=lo, **ctor_kw)
    r = part.fit(**fit_kw)
    for key in ('percentiles', 'mean', 'std', 'acceptance_fraction', 'flags'):
        np.testing.assert_array_equal(r[key], res[key][lo:hi], err_msg=key)


def test_c5_shard_full_size():
    """BASELINE config 5, one GPU's shard at full size: 12,500 spectra x 256 walkers x 2000 steps,
    Debye decomposition poly_deg 4, 64 taus (5.1e9 log-prob evaluations)."""
    from bisip_b200.batch import BatchInversion
    B = 12500
    ctor = dict(nwalkers=256, nsteps=2000, poly_deg=4, n_tau=64, seed=0xB151B)
    fit_kw = dict(discard=1000, thin=10)
    w, syn, truth = _make('decomp', B, poly_deg=4, n_tau=64)
    assert np.all((truth[:, 0] > 0.9) & (truth[:, 0] < 1.1))          # the truth lies inside the prior box
    inv = BatchInversion('decomp', w, syn['zn'], syn['zn_err'], **ctor)
    res = inv.fit(**fit_kw)
    assert res['percentiles'].shape == (B, 3, 6) and np.all(res['flags'] == 0)
    assert np.all(np.isfinite(res['mean'])) and np.all(res['std'] > 0)
    acc = res['acceptance_fraction']
    assert 0.40 < acc.min() and acc.max() < 0.60
    cov, zm, zs = _calibration(res, truth)
    # binomial sd of the coverage over 12,500 spectra is 0.002; measured 0.949-0.951
    assert np.all(np.abs(cov - 0.95) < 0.01), cov
    assert np.all(np.abs(zm) < 0.15), zm
    assert np.all(np.abs(zs - 1.0) < 0.05), zs
    _subset_identical('decomp', w, syn, res, 6000, 6004, fit_kw, **ctor)
    # derived product at scale: total chargeability of the posterior-mean RTD recovers the truth's
    from bisip_b200 import products
    _, tot = inv.rtd()
    tot_true = products.total_chargeability(truth[:, 1:], inv.log_taus)
    assert np.median(np.abs(tot - tot_true) / tot_true) < 0.05          # measured 0.026 (1 % data noise)


@pytest.mark.parametrize("model,min_cov", [('dias', 0.90), ('shin', 0.80)])
def test_c3_full_size(model, min_cov):
    """BASELINE config 3: Dias2000 / Shin2015 on 1,024 synthetic 64-frequency spectra, 128 walkers x 2000 steps.
    These posteriors are bounded, skewed and partly prior-dominated (Shin's log_Q boxes are 2 units wide), so
    the 95 % interval over-covers for some parameters; the bar is a lower bound per parameter plus an
    unbiased standardised error."""
    from bisip_b200.batch import BatchInversion
    from oracle import oracle
    B = 1024
    ctor = dict(nwalkers=128, nsteps=2000, seed=77)
    fit_kw = dict(discard=1000, thin=5)
    w, syn, truth = _make(model, B)
    inv = BatchInversion(model, w, syn['zn'], syn['zn_err'], **ctor)
    res = inv.fit(**fit_kw)
    assert np.all(res['flags'] == 0) and np.all(np.isfinite(res['mean']))
    cov, zm, zs = _calibration(res, truth)
    assert np.all(cov >= min_cov), cov
    assert np.all(np.abs(zm) < 0.6), zm
    _subset_identical(model, w, syn, res, 1000, 1003, fit_kw, **ctor)
    # chain prefix of spectrum 1000 against the oracle on the same Philox stream
    b, T = 1000, 40
    one = BatchInversion(model, w, syn['zn'][b:b + 1], syn['zn_err'][b:b + 1], nwalkers=128, nsteps=T, seed=77,
                         spectrum_offset=b)
    p0 = one.draw_p0(0, 1)
    r1 = one.fit(p0=p0, keep_chain=True)
    prob = oracle.Problem(model, w, syn['zn'][b], syn['zn_err'][b], one.param_bounds)
    ref = prob.run(p0[0], T, seed=77, spectrum=b)
    np.testing.assert_array_equal(r1['chain'][0], ref['chain'])


def test_c5_shard_full_size_collapsed():
    """The same full-size shard with precision='fp64-collapsed' (z = (L K) a, csrc/decomp_collapsed.cuh): calibrated,
    independent of how the batch is cut, and — being FP64 to rounding — the SAME chains as the two-stage DMMA path:
    summaries of a 512-spectra slice are identical except where a ~1e-13 log-prob rounding difference flipped one
    of that slice's 5e8 accept tests (expected ~0.01 spectra; allowed: 1 %, and those within Monte-Carlo error)."""
    from bisip_b200 import engine
    from bisip_b200.batch import BatchInversion
    B = 12500
    ctor = dict(nwalkers=256, nsteps=2000, poly_deg=4, n_tau=64, seed=0xB151B, precision='fp64-collapsed')
    fit_kw = dict(discard=1000, thin=10)
    w, syn, truth = _make('decomp', B, poly_deg=4, n_tau=64)
    inv = BatchInversion('decomp', w, syn['zn'], syn['zn_err'], **ctor)
    assert engine.decomp_kernel_kind(inv._spec(), 64, 256) == 'fp64-collapsed'
    res = inv.fit(**fit_kw)
    assert np.all(res['flags'] == 0) and np.all(np.isfinite(res['mean'])) and np.all(res['std'] > 0)
    acc = res['acceptance_fraction']
    assert 0.40 < acc.min() and acc.max() < 0.60
    cov, zm, zs = _calibration(res, truth)
    assert np.all(np.abs(cov - 0.95) < 0.01), cov
    assert np.all(np.abs(zm) < 0.15), zm
    assert np.all(np.abs(zs - 1.0) < 0.05), zs
    _subset_identical('decomp', w, syn, res, 6000, 6004, fit_kw, **ctor)
    ref = BatchInversion('decomp', w, syn['zn'][:512], syn['zn_err'][:512], **dict(ctor, precision='fp64')).fit(**fit_kw)
    same = np.all(res['percentiles'][:512] == ref['percentiles'], axis=(1, 2))
    assert same.mean() >= 0.99, same.mean()
    shift = np.abs(res['percentiles'][:512, 1] - ref['percentiles'][:, 1]) / ref['std']
    assert shift.max() < 0.35, shift.max()


@pytest.mark.parametrize("prec", ['3xtf32', 'tf32'])
def test_c5_shard_full_size_tcgen05(prec):
    """The same full-size shard on the tcgen05 kernel (TF32 / 3xTF32 operands, FP32 accumulators in tensor memory):
    the posterior must be as well calibrated as the FP64 one, summaries must not depend on how the batch is cut, and
    the medians must agree with an FP64 run of a slice within Monte-Carlo error."""
    from bisip_b200 import engine
    from bisip_b200.batch import BatchInversion
    B = 12500
    ctor = dict(nwalkers=256, nsteps=2000, poly_deg=4, n_tau=64, seed=0xB151B, precision=prec)
    fit_kw = dict(discard=1000, thin=10)
    w, syn, truth = _make('decomp', B, poly_deg=4, n_tau=64)
    inv = BatchInversion('decomp', w, syn['zn'], syn['zn_err'], **ctor)
    assert engine.decomp_kernel_kind(inv._spec(), 64, 256) == 'tcgen05'
    res = inv.fit(**fit_kw)
    assert np.all(res['flags'] == 0) and np.all(np.isfinite(res['mean'])) and np.all(res['std'] > 0)
    acc = res['acceptance_fraction']
    assert 0.40 < acc.min() and acc.max() < 0.60
    cov, zm, zs = _calibration(res, truth)
    assert np.all(np.abs(cov - 0.95) < 0.01), cov
    assert np.all(np.abs(zm) < 0.15), zm
    assert np.all(np.abs(zs - 1.0) < 0.05), zs
    _subset_identical('decomp', w, syn, res, 6000, 6004, fit_kw, **ctor)
    ref = BatchInversion('decomp', w, syn['zn'][:512], syn['zn_err'][:512], **dict(ctor, precision='fp64')).fit(**fit_kw)
    shift = np.abs(res['percentiles'][:512, 1] - ref['percentiles'][:, 1]) / ref['std']
    assert shift.max() < 0.35 and np.median(shift) < 0.05, (shift.max(), np.median(shift))
    assert np.abs(res['std'][:512] / ref['std'] - 1).max() < 0.25


def test_c4_256_taus_many_spectra():
    """BASELINE config 4 shape: 256 taus, 256 walkers, thousands of spectra (2-CTA cluster path), FP64 DMMA
    calibration plus the TF32 / 3xTF32 log-probability tolerance over every spectrum at its truth."""
    from bisip_b200 import _lib, engine
    from bisip_b200.batch import BatchInversion
    B = 2000
    ctor = dict(nwalkers=256, nsteps=1500, poly_deg=4, n_tau=256, seed=4)
    fit_kw = dict(discard=750, thin=5)
    w, syn, truth = _make('decomp', B, poly_deg=4, n_tau=256)
    inv = BatchInversion('decomp', w, syn['zn'], syn['zn_err'], **ctor)
    res = inv.fit(**fit_kw)
    assert np.all(res['flags'] == 0)
    cov, zm, zs = _calibration(res, truth)
    assert np.all(np.abs(cov - 0.95) < 0.025), cov          # binomial sd over 2,000 spectra: 0.005
    assert np.all(np.abs(zm) < 0.2), zm
    _subset_identical('decomp', w, syn, res, 1500, 1502, fit_kw, **ctor)
    dev = inv.device
    args = (_lib.dev_f64(truth[:, None, :], dev), _lib.dev_f64(w, dev), _lib.dev_f64(syn['zn'], dev),
            _lib.dev_f64(syn['zn_err'], dev), _lib.dev_f64(inv.param_bounds, dev))
    lp64 = engine.log_probability(inv._spec(), *args).cpu().numpy()[:, 0]
    assert np.all(np.isfinite(lp64))
    for prec, tol in (('tf32', 3e-3), ('3xtf32', 5e-5)):
        alt = BatchInversion('decomp', w, syn['zn'][:1], syn['zn_err'][:1], precision=prec, **ctor)
        lp = engine.log_probability(alt._spec(), *args).cpu().numpy()[:, 0]
        err = np.abs(lp - lp64) / np.maximum(1, np.abs(lp64))
        assert 0 < err.max() <= tol, (prec, err.max())
