"""Derived products and result tables: RTD / total chargeability of the
polynomial decomposition and the quickstart CSV layout.  CPU only."""
import numpy as np
import pytest

from bisip_b200 import products

# posterior means and total_m printed by the reference tutorial (docs/tutorials/decomposition.ipynb
# cells 22 and 27; 6 printed decimals => totals reproduce to ~0.5 %)
NOTEBOOK = {
    'SIP-K389170': ([0.014033, -0.002350, -0.004634, -0.000338, 0.000219], 1.342871),
    'SIP-K389172': ([0.017680, -0.011578, -0.003509, 0.002061, 0.000526], 1.210898),
    'SIP-K389173': ([0.003390, -0.000995, -0.000499, 0.000396, 0.000172], 0.786556),
    'SIP-K389174': ([0.007547, -0.003206, -0.001871, 0.000548, 0.000242], 0.930805),
    'SIP-K389175': ([0.006870, -0.003937, -0.001338, 0.000741, 0.000219], 0.655028),
    'SIP-K389176': ([0.002426, -0.001084, -0.000048, 0.000564, 0.000159], 0.545386),
}


def _log_taus(poly_deg=4):
    log_tau = np.linspace(-6, 2, 40)          # decomposition.ipynb cell 25 printout (N=20 -> 40 taus)
    return log_tau, np.array([log_tau ** i for i in range(poly_deg + 1)])


def test_rtd_matches_tutorial_loop():
    log_tau, log_taus = _log_taus()
    rng = np.random.default_rng(0)
    a = rng.normal(size=(7, 5))
    m = products.relaxation_time_distribution(a, log_taus)
    ref = np.zeros((7, 40))
    for p in range(5):                        # the tutorial's get_m, literally
        ref += a[:, p, None] * log_tau ** p
    np.testing.assert_array_equal(m, ref)
    np.testing.assert_allclose(products.total_chargeability(a, log_taus), ref.sum(1), rtol=1e-15)
    assert products.relaxation_time_distribution(a[0], log_taus).shape == (40,)
    with pytest.raises(ValueError):
        products.relaxation_time_distribution(a[:, :4], log_taus)


@pytest.mark.parametrize("name", sorted(NOTEBOOK))
def test_total_chargeability_notebook_values(name):
    _, log_taus = _log_taus()
    a, total = NOTEBOOK[name]
    assert products.total_chargeability(np.array(a), log_taus) == pytest.approx(total, rel=1e-2)


def test_quickstart_csv_layout(tmp_path, golden_dir):
    """save_percentiles_csv reproduces the reference's quickstart_results.csv byte for byte."""
    gold = (golden_dir / 'quickstart_results.csv').read_text()
    names, table = products.load_percentiles_csv(golden_dir / 'quickstart_results.csv')
    assert names == ['r0', 'm1', 'log_tau1', 'c1'] and table.shape == (3, 4)
    out = tmp_path / 'out.csv'
    products.save_percentiles_csv(out, names, table)
    assert out.read_text() == gold
    with pytest.raises(ValueError):
        products.save_percentiles_csv(out, names[:3], table)


def test_batch_table_columns():
    B, ndim = 3, 2
    res = {'mean': np.arange(6.).reshape(B, ndim), 'std': np.ones((B, ndim)),
           'percentiles': np.arange(18.).reshape(B, 3, ndim), 'acceptance_fraction': np.full(B, 0.4),
           'flags': np.zeros(B, dtype=np.int32)}
    cols, table = products.batch_table(['r0', 'm'], res, ids=[10, 11, 12])
    assert cols == ['spectrum', 'acceptance_fraction', 'flags', 'r0_mean', 'm_mean', 'r0_std', 'm_std',
                    'r0_p2.5', 'm_p2.5', 'r0_p50', 'm_p50', 'r0_p97.5', 'm_p97.5']
    assert table.shape == (3, len(cols))
    np.testing.assert_array_equal(table[:, 0], [10, 11, 12])
    np.testing.assert_array_equal(table[1, 7:9], res['percentiles'][1, 0])
