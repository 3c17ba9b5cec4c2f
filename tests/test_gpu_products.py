"""GPU tests of the derived products: RTD / total chargeability from a fit,
quickstart CSV export, batch result table."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rtd_total_chargeability_and_csv(tmp_path, data_files):
    import bisip_b200 as bb
    from bisip_b200 import products
    np.random.seed(42)
    m = bb.PolynomialDecomposition(data_files['SIP-K389175'], nwalkers=32, poly_deg=4, nsteps=1000)
    m.fit()
    mean = m.get_param_mean(discard=500)
    rtd = m.get_rtd(discard=500)
    assert rtd.shape == (40,)
    ref = sum(mean[1 + p] * m.log_tau ** p for p in range(5))          # tutorial's get_m
    np.testing.assert_allclose(rtd, ref, rtol=1e-13, atol=1e-16)
    tot = m.get_total_chargeability(discard=500)
    assert tot == pytest.approx(ref.sum(), rel=1e-12)
    # reference notebook (seed 42, unseeded MT19937 stream differs): 0.655028, MC scatter ~2 %
    assert tot == pytest.approx(0.655028, rel=0.1)
    # RTD percentiles over the chain: exact order statistics of a @ log_taus
    flat = m.get_chain(discard=500, thin=5, flat=True)
    pr = m.get_rtd(p=[2.5, 50, 97.5], chain=flat)
    want = np.percentile(products.relaxation_time_distribution(flat[:, 1:], m.log_taus), [2.5, 50, 97.5], axis=0)
    np.testing.assert_array_equal(pr, want)
    assert m.get_rtd(p=50, chain=flat).shape == (40,)
    # quickstart-style CSV
    out = tmp_path / 'res.csv'
    table = m.save_results(out, discard=500, thin=2)
    names, back = products.load_percentiles_csv(out)
    assert names == m.param_names
    np.testing.assert_array_equal(back, table)
    np.testing.assert_array_equal(table, m.get_param_percentile(discard=500, thin=2))


def test_batch_rtd_and_table(tmp_path, gold_fl):
    from bisip_b200 import products, synthetic
    from bisip_b200.batch import BatchInversion
    tag = 'syn_decomp_s64'
    _, w = synthetic.frequencies(64)
    inv = BatchInversion('decomp', w, gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err'], nwalkers=32, nsteps=200,
                         poly_deg=4, n_tau=64, seed=3)
    with pytest.raises(AssertionError):
        inv.rtd()
    res = inv.fit(discard=100, thin=2, percentiles=(16, 50, 84))
    m, tot = inv.rtd()
    B = res['mean'].shape[0]
    assert m.shape == (B, 64) and tot.shape == (B,)
    np.testing.assert_array_equal(m[1], products.relaxation_time_distribution(res['mean'][1, 1:], inv.log_taus))
    m50, _ = inv.rtd(stat=1)
    np.testing.assert_array_equal(m50[0], products.relaxation_time_distribution(res['percentiles'][0, 1, 1:], inv.log_taus))
    cols = inv.to_csv(tmp_path / 'batch.csv', ids=np.arange(100, 100 + B))
    assert 'a4_p84' in cols and 'r0_mean' in cols
    table = np.loadtxt(tmp_path / 'batch.csv', delimiter=',', skiprows=1)
    assert table.shape == (B, len(cols))
    np.testing.assert_array_equal(table[:, cols.index('r0_p50')], res['percentiles'][:, 1, 0])
    cc = BatchInversion('colecole', w, gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err'], nwalkers=32, nsteps=10)
    with pytest.raises(ValueError):
        cc.rtd()


def test_batch_autocorr_time_on_the_kept_chain():
    """Convergence diagnostic for batches: integrated autocorrelation times of every spectrum on
    the GPU equal the per-chain emcee-style estimator, and the stretch move's tau is O(10) steps for 6 parameters."""
    from bisip_b200 import sampler, synthetic, engine, _lib
    from bisip_b200.batch import BatchInversion
    _, w = synthetic.frequencies(64)
    probe = BatchInversion('decomp', w, np.zeros((1, 2, 64)), np.ones((1, 2, 64)), poly_deg=4, n_tau=64)
    fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], probe.device),
                                        _lib.dev_f64(ww, probe.device))[:, 0].cpu().numpy()
    syn = synthetic.make('decomp', 0, 6, fwd, N=64, poly_deg=4, n_tau=64)
    inv = BatchInversion('decomp', w, syn['zn'], syn['zn_err'], nwalkers=64, nsteps=3000, poly_deg=4, n_tau=64, seed=9)
    with pytest.raises(AssertionError):
        inv.get_autocorr_time()
    res = inv.fit(discard=1000, thin=2, keep_chain=True)
    tau = inv.get_autocorr_time(thin=2)
    assert tau.shape == (6, 6) and np.all(np.isfinite(tau))
    for b in range(6):
        np.testing.assert_allclose(tau[b], 2 * sampler.integrated_time(res['chain'][b], quiet=True), rtol=1e-9)
    assert 5 < np.median(tau) < 200
