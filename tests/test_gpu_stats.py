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
    mp = m.get_model_percentile([2.5, 50, 97.5], ch)
    Z = m.forward(ch, m.data['w'])
    np.testing.assert_array_equal(mp, np.percentile(Z, [2.5, 50, 97.5], axis=0))
    assert mp.shape == (3, 2, m.data['N'])
    assert m.get_model_percentile(50, ch).shape == (2, m.data['N'])
