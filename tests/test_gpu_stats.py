"""Device chain statistics vs NumPy: percentiles must be BIT-EXACT (exact order statistics +
NumPy's lerp), mean/std within 1e-13 relative (different summation order)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,ncol,B", [(1, 3, 2), (2, 1, 1), (257, 6, 3), (4096, 7, 2), (25600, 6, 4), (70001, 2, 1)])
def test_column_stats_vs_numpy(n, ncol, B):
    from bisip_b200 import _lib, engine
    dev = _lib.require_cuda()
    rng = np.random.default_rng(n)
    x = rng.standard_normal((B, n, ncol)) * np.array([10.0 ** (i - 3) for i in range(ncol)])
    x[:, : n // 3, 0] = np.round(x[:, : n // 3, 0], 1)        # ties
    if ncol > 1:
        x[..., 1] = -np.abs(x[..., 1])                          # all negative
    p = [0, 2.5, 16, 50, 84, 97.5, 99.9, 100]
    st = engine.column_stats(_lib.dev_f64(x, dev), p=p, want_mean=True, want_std=True)
    np.testing.assert_array_equal(st['pct'].cpu().numpy(), np.percentile(x, p, axis=1).transpose(1, 0, 2))
    # mean of a zero-centred sample cancels: tolerance is relative to the sample scale
    np.testing.assert_allclose(st['mean'].cpu().numpy(), x.mean(1), rtol=1e-13, atol=1e-14 * np.abs(x).max())
    np.testing.assert_allclose(st['std'].cpu().numpy(), x.std(1), rtol=1e-12, atol=1e-300)


def test_more_than_16_percentiles():
    from bisip_b200 import _lib, engine
    dev = _lib.require_cuda()
    x = np.random.default_rng(0).standard_normal((1, 999, 2))
    p = np.linspace(0, 100, 41)
    st = engine.column_stats(_lib.dev_f64(x, dev), p=p)
    np.testing.assert_array_equal(st['pct'][0].cpu().numpy(), np.percentile(x[0], p, axis=0))


def test_param_and_model_percentiles_api(gold_fl, data_files):
    """get_param_percentile/mean/std and get_model_percentile against NumPy on the same chain
    (reference utils.py:17-85), incl. the parse_chain warning / error contract."""
    import bisip_b200 as bb
    m = bb.PolynomialDecomposition(data_files['SIP-K389175'], nwalkers=32, poly_deg=4, nsteps=300, seed=1)
    np.random.seed(1)
    m.fit()
    ch = m.get_chain(discard=100, thin=2, flat=True)
    np.testing.assert_array_equal(m.get_param_percentile(discard=100, thin=2), np.percentile(ch, [2.5, 50, 97.5], axis=0))
    np.testing.assert_array_equal(m.get_param_percentile(50, chain=ch), np.percentile(ch, 50, axis=0))
    np.testing.assert_allclose(m.get_param_mean(ch), ch.mean(0), rtol=1e-13)
    np.testing.assert_allclose(m.get_param_std(discard=100, thin=2), ch.std(0), rtol=1e-12)
    with pytest.warns(UserWarning):
        m.get_param_mean()
    with pytest.raises(ValueError):
        m.get_param_mean(m.get_chain())                        # 3-D chain
    with pytest.raises(ValueError):
        m.get_param_mean(ch, discard=10)
    # model percentiles: forward over the whole flat chain, then percentile over samples
    # (fused kernel: the decomposition column is evaluated in its collapsed FP64 form, 1e-12 from the two-stage forward)
    mp = m.get_model_percentile([2.5, 50, 97.5], ch)
    Z = m.forward(ch, m.data['w'])
    want = np.percentile(Z, [2.5, 50, 97.5], axis=0)
    assert np.max(np.abs(mp - want)) <= 1e-12 * np.max(np.abs(want))
    assert mp.shape == (3, 2, m.data['N'])
    assert m.get_model_percentile(50, ch).shape == (2, m.data['N'])


@pytest.mark.parametrize("kind", ["constant", "two_values", "mostly_equal", "outlier", "bimodal_far", "denormal_span",
                                  "huge_span", "with_inf"])
def test_column_stats_hard_distributions(kind):
    """Exactness must not depend on the equal-width binning of the shared-memory select: constant columns, heavy
    ties (the selected bin overflows the candidate pool -> radix fallback), a far outlier that squeezes the bulk into
    one bin, two far modes, spans that overflow (max - min) and infinities."""
    from bisip_b200 import _lib, engine
    dev = _lib.require_cuda()
    rng = np.random.default_rng(5)
    n = 20000
    x = rng.standard_normal((2, n, 3))
    if kind == "constant":
        x[..., 0] = 1.25
    elif kind == "two_values":
        x[..., 0] = np.where(rng.random((2, n)) < 0.3, -1.0, 2.0)
    elif kind == "mostly_equal":
        x[..., 0] = np.where(rng.random((2, n)) < 0.9, 0.5, x[..., 0])
    elif kind == "outlier":
        x[..., 0] = 1.0 + 1e-6 * x[..., 0]
        x[:, 17, 0] = -1e9
    elif kind == "bimodal_far":
        x[..., 0] = np.where(rng.random((2, n)) < 0.5, 1e-3 * x[..., 0], 1e6 + x[..., 0])
    elif kind == "denormal_span":
        x[..., 0] = x[..., 0] * 1e-310
    elif kind == "huge_span":
        x[..., 0] = x[..., 0] * 1e308
    elif kind == "with_inf":
        x[:, 5, 0] = np.inf
        x[:, 9, 0] = -np.inf
    p = [0, 2.5, 50, 97.5, 100]
    st = engine.column_stats(_lib.dev_f64(x, dev), p=p)
    with np.errstate(invalid='ignore', over='ignore'):
        want = np.percentile(x, p, axis=1).transpose(1, 0, 2)
    np.testing.assert_array_equal(st['pct'].cpu().numpy(), want)


def test_column_stats_nan_propagates():
    from bisip_b200 import _lib, engine
    dev = _lib.require_cuda()
    x = np.random.default_rng(1).standard_normal((1, 5000, 2))
    x[0, 77, 1] = np.nan
    st = engine.column_stats(_lib.dev_f64(x, dev), p=[50.0])
    got = st['pct'].cpu().numpy()
    assert np.isnan(got[0, 0, 1]) and got[0, 0, 0] == np.percentile(x[0, :, 0], 50)


@pytest.mark.parametrize("case", ['colecole_k2', 'dias', 'shin', 'decomp_p4_debye', 'decomp_p4_warburg'])
def test_fused_model_percentile_matches_forward_then_percentile(case, gold_fl, gold_ld):
    """bisip_model_percentile (forward + select fused per model column) against bisip_forward + np.percentile:
    to rounding (1e-14) for the vector models, 1e-12 for the decomposition (collapsed column form)."""
    from bisip_b200 import _lib, engine
    from helpers import CASES
    dev = _lib.require_cuda()
    model, kw = CASES[case]
    lo, hi = gold_fl[f'{case}/bounds']
    rng = np.random.default_rng(3)
    n = 6001
    centre = gold_fl[f'{case}/theta'][0]
    th = centre + 0.02 * (hi - lo) * rng.standard_normal((2, n, lo.shape[0]))
    th = np.clip(th, lo + 1e-9, hi - 1e-9)
    w = gold_ld['SIP-K389175/w']
    spec = engine.ModelSpec(model={'decomp': _lib.MODEL_DECOMP, 'colecole': _lib.MODEL_COLECOLE, 'dias': _lib.MODEL_DIAS,
                                   'shin': _lib.MODEL_SHIN}[model], ndim=lo.shape[0], n_modes=kw.get('n_modes', 1),
                            taus=_lib.dev_f64(gold_fl[f'{case}/taus'], dev) if model == 'decomp' else None,
                            log_taus=_lib.dev_f64(gold_fl[f'{case}/log_taus'], dev) if model == 'decomp' else None,
                            c_exp=kw.get('c_exp', 1.0))
    thd, wd = _lib.dev_f64(th, dev), _lib.dev_f64(w, dev)
    p = [2.5, 16, 50, 84, 97.5]
    got = engine.model_percentile(spec, thd, wd, p).cpu().numpy()
    Z = engine.forward(spec, thd, wd).cpu().numpy()                      # (2, n, 2, N)
    want = np.percentile(Z, p, axis=1).transpose(1, 0, 2, 3)
    assert got.shape == want.shape == (2, 5, 2, w.shape[0])
    # vector models: the same formulas as the batched forward (the compiler may contract a*b+c differently in the two
    # kernels: last-bit differences); decomposition: collapsed column form vs the two-stage contraction
    tol = 1e-12 if model == 'decomp' else 1e-14
    assert np.max(np.abs(got - want)) <= tol * np.max(np.abs(want))


def test_model_percentile_long_chain_falls_back(data_files):
    """Chains beyond one CTA's shared memory (~27,000 samples) go through bisip_forward + bisip_column_stats."""
    import bisip_b200 as bb
    from bisip_b200 import _lib, engine
    m = bb.Dias2000(data_files['SIP-K389172'], nwalkers=32, nsteps=10)
    rng = np.random.default_rng(0)
    lo, hi = m.param_bounds
    ch = rng.uniform(lo + 0.3 * (hi - lo), hi - 0.3 * (hi - lo), (40000, 5))
    dev = _lib.require_cuda()
    assert engine.model_percentile(m._spec(dev), _lib.dev_f64(ch[None], dev), _lib.dev_f64(m.data['w'], dev), [50]) is None
    mp = m.get_model_percentile([2.5, 50, 97.5], ch)
    np.testing.assert_array_equal(mp, np.percentile(m.forward(ch, m.data['w']), [2.5, 50, 97.5], axis=0))
