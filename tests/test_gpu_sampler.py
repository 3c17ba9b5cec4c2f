"""GPU sampler tests: the on-device stretch-move ensemble vs the C oracle running the SAME
Philox stream, emcee storage/slicing semantics, and Monte-Carlo-error agreement with the
reference fit() (posterior.npz: reference log-probability driven by the emcee restatement).
"""
import numpy as np
import pytest

from helpers import CASES, make_model, oracle_problem

pytestmark = pytest.mark.gpu


def _gpu_run(case, gold_fl, data_files, W, T, seed, **kw):
    m = make_model(case, data_files['SIP-K389175'], nwalkers=W, nsteps=T, seed=seed)
    rng = np.random.default_rng(seed)
    lo, hi = gold_fl[f'{case}/bounds']
    p0 = rng.uniform(lo, hi, (W, lo.shape[0]))
    m.fit(p0=p0)
    return m, p0


@pytest.mark.parametrize("case,W,T", [('decomp_p4_debye', 32, 200), ('decomp_p4_warburg', 64, 100),
                                      ('decomp_p5_debye', 256, 60), ('colecole_k2', 64, 200),
                                      ('colecole_k1', 33, 150), ('dias', 32, 200), ('shin', 128, 100)])
def test_chain_matches_oracle_same_stream(case, W, T, gold_fl, gold_ld, data_files):
    """Same p0, same Philox stream => same accept/reject decisions => same chain.
    Positions are produced by identical unfused arithmetic, so they agree to rounding of
    nothing (bit-exact) unless a log-prob rounding difference flips an acceptance, which at
    ~1e-13 relative lp error has probability ~1e-9 per run."""
    seed = 1234
    m, p0 = _gpu_run(case, gold_fl, data_files, W, T, seed)
    prob = oracle_problem(case, gold_fl, gold_ld)
    ref = prob.run(p0, T, seed=seed, spectrum=0)
    chain = m.get_chain()
    assert chain.shape == (T, W, p0.shape[1])
    np.testing.assert_array_equal(chain, ref['chain'])
    np.testing.assert_array_equal(m.sampler.accepted, ref['accepted'])
    lp = m.sampler.get_log_prob()
    fin = np.isfinite(ref['log_prob'])
    assert np.array_equal(fin, np.isfinite(lp))
    assert np.max(np.abs(lp[fin] - ref['log_prob'][fin]) / np.maximum(1, np.abs(ref['log_prob'][fin]))) <= 1e-12


@pytest.mark.parametrize("tag,W,T", [('syn_decomp_s128', 64, 40), ('syn_decomp_s256', 256, 12), ('syn_decomp_s128', 33, 30)])
def test_large_tau_grid_chain_matches_oracle(tag, W, T, gold_fl):
    """n_tau > 64: stage-1 recompute + column split over a CTA cluster (DSMEM exchange of partial
    chi^2).  Same Philox stream as the oracle => identical chain."""
    from bisip_b200 import synthetic
    from bisip_b200.batch import BatchInversion
    from oracle import oracle
    _, w = synthetic.frequencies(64)
    S, c_exp = int(tag.split('_s')[1]), float(gold_fl[f'{tag}/c_exp'])
    zn, ze, bounds = gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err'], gold_fl[f'{tag}/bounds']
    inv = BatchInversion('decomp', w, zn, ze, nwalkers=W, nsteps=T, poly_deg=4, n_tau=S, c_exp=c_exp, seed=77)
    p0 = inv.draw_p0(0, 4)
    res = inv.fit(p0=p0, keep_chain=True)
    for b in (0, 3):
        prob = oracle.Problem('decomp', w, zn[b], ze[b], bounds, taus=gold_fl[f'{tag}/taus'],
                              log_taus=gold_fl[f'{tag}/log_taus'], c_exp=c_exp)
        ref = prob.run(p0[b], T, seed=77, spectrum=b)
        np.testing.assert_array_equal(res['chain'][b], ref['chain'])
        fin = np.isfinite(ref['log_prob'])
        assert np.max(np.abs(res['log_prob'][b][fin] - ref['log_prob'][fin]) / np.maximum(1, np.abs(ref['log_prob'][fin]))) <= 1e-12


def test_reference_default_tau_grid_n64(gold_fl):
    """The reference hard-codes n_tau = 2N (models.py:203): 128 taus for a 64-frequency spectrum."""
    from bisip_b200 import synthetic
    from bisip_b200.batch import BatchInversion
    _, w = synthetic.frequencies(64)
    inv = BatchInversion('decomp', w, gold_fl['syn_decomp_s128/zn'], gold_fl['syn_decomp_s128/zn_err'], nwalkers=32,
                         nsteps=20, poly_deg=4)
    assert inv.taus.shape == (128,)
    np.testing.assert_array_equal(inv.taus, gold_fl['syn_decomp_s128/taus'])
    r = inv.fit()
    assert np.all(r['flags'] == 0) and np.all(np.isfinite(r['mean']))


def test_logp_consistent_with_chain(gold_fl, data_files):
    """Stored log-probabilities equal the log-probability recomputed at the stored positions."""
    m, _ = _gpu_run('decomp_p4_debye', gold_fl, data_files, 64, 300, 5)
    ch = m.get_chain(flat=True)
    lp = m._log_probability(ch, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
    st = m.sampler.get_log_prob(flat=True)
    fin = np.isfinite(st)
    assert np.array_equal(fin, np.isfinite(lp))
    assert np.max(np.abs(lp[fin] - st[fin]) / np.maximum(1, np.abs(st[fin]))) <= 1e-12


def test_chain_layout_and_slicing(gold_fl, data_files):
    """emcee layout (nsteps, nwalkers, ndim); get_chain(discard, thin, flat) = chain[discard+thin-1::thin]
    (SURVEY App. B.4); stored notebook shapes (1500,32,4) and (24000,4) for T=2000."""
    import bisip_b200 as bb
    m = bb.PeltonColeCole(data_files['SIP-K389172'], nwalkers=32, nsteps=2000, headers=9, seed=3)
    assert m.data['N'] == 12
    np.random.seed(0)
    m.fit()
    full = m.get_chain()
    assert full.shape == (2000, 32, 4) and full.dtype == np.float64
    assert m.get_chain(discard=500).shape == (1500, 32, 4)
    flat = m.get_chain(discard=500, thin=2, flat=True)
    assert flat.shape == (24000, 4)
    np.testing.assert_array_equal(flat, full[501::2].reshape(-1, 4))
    np.testing.assert_array_equal(m.get_chain(discard=7, thin=5), full[11::5])
    assert m.sampler.chain.shape == (32, 2000, 4)
    # device-side storage of kept steps only == host slicing of the full chain
    from bisip_b200 import _lib, engine
    dev = _lib.require_cuda()
    res = engine.ensemble_run(m._spec(dev), _lib.dev_f64(m.p0[None], dev).clone(), _lib.dev_f64(m.data['w'], dev),
                              _lib.dev_f64(m.data['zn'][None], dev), _lib.dev_f64(m.data['zn_err'][None], dev),
                              _lib.dev_f64(m.param_bounds, dev), nsteps=2000, seed=m.sampler.seed, discard=500, thin=7)
    np.testing.assert_array_equal(res['chain'][0].cpu().numpy(), full[506::7])
    np.testing.assert_array_equal(res['log_prob'][0].cpu().numpy(), m.sampler.get_log_prob()[506::7])


def test_results_independent_of_batching(gold_fl):
    """Spectrum b of a batch == the same spectrum run alone with spectrum0=b (Philox counter
    carries the global index) — the property multi-GPU sharding relies on."""
    from bisip_b200 import synthetic
    from bisip_b200.batch import BatchInversion
    tag = 'syn_decomp_s64'
    _, w = synthetic.frequencies(64)
    zn, ze = gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err']
    full = BatchInversion('decomp', w, zn, ze, nwalkers=32, nsteps=100, poly_deg=4, n_tau=64, seed=11)
    r_full = full.fit(discard=20, thin=2, keep_chain=True)
    part = BatchInversion('decomp', w, zn[2:4], ze[2:4], nwalkers=32, nsteps=100, poly_deg=4, n_tau=64, seed=11,
                          spectrum_offset=2)
    r_part = part.fit(discard=20, thin=2, keep_chain=True, batch_size=1)
    np.testing.assert_array_equal(r_full['chain'][2:4], r_part['chain'])
    np.testing.assert_array_equal(r_full['percentiles'][2:4], r_part['percentiles'])
    assert r_full['chain'].shape == (4, 40, 32, 6)


@pytest.mark.parametrize("tag,ctor", [
    ('c1_decomp', lambda bb, df: bb.PolynomialDecomposition(df['SIP-K389175'], nwalkers=32, poly_deg=4, nsteps=1000)),
    ('c2_colecole', lambda bb, df: bb.PeltonColeCole(df['SIP-K389174'], nwalkers=64, n_modes=2, nsteps=2000)),
    ('dias', lambda bb, df: bb.Dias2000(df['SIP-K389172'], nwalkers=32, nsteps=2000)),
])
def test_posterior_matches_reference_within_mc_error(tag, ctor, gold_post, data_files):
    """Acceptance fraction and posterior mean / percentiles agree with the reference fit()
    (several seeds each) within Monte-Carlo error: |mean_gpu - mean_ref| <= 5 combined standard
    errors of the seed-to-seed scatter (plus a small floor)."""
    import bisip_b200 as bb
    edits = {'c2_colecole': {'log_tau1': [-5, 5], 'log_tau2': [-15, -10]}, 'dias': {'eta': [0, 25], 'log_tau': [-15, -5]}}
    discard = 500 if tag == 'c1_decomp' else 1000
    means, pcts, accs = [], [], []
    for s in range(6):
        np.random.seed(100 + s)
        m = ctor(bb, data_files)
        m.params.update(edits.get(tag, {}))
        m.fit()
        ch = m.get_chain(discard=discard, flat=True)
        means.append(ch.mean(0)); pcts.append(np.percentile(ch, [2.5, 50, 97.5], axis=0))
        accs.append(m.sampler.acceptance_fraction.mean())
    means, pcts, accs = np.array(means), np.array(pcts), np.array(accs)
    rm, rp, ra = gold_post[f'{tag}/mean'], gold_post[f'{tag}/pct'], gold_post[f'{tag}/acc']
    post_sd = gold_post[f'{tag}/std'].mean(0)

    def close(a, b, floor):
        se = np.sqrt(a.var(0, ddof=1) / len(a) + b.var(0, ddof=1) / len(b))
        return np.all(np.abs(a.mean(0) - b.mean(0)) <= 5 * se + floor)
    assert close(means, rm, 0.05 * post_sd)
    assert close(pcts, rp, 0.15 * post_sd)
    assert abs(accs.mean() - ra.mean()) <= 0.03


def test_sampler_errors(data_files):
    import bisip_b200 as bb
    fp = data_files['SIP-K389175']
    m = bb.Dias2000(fp, nwalkers=8, nsteps=10)
    with pytest.raises(AssertionError):
        m.get_chain()
    with pytest.raises(RuntimeError):             # nwalkers < 2*ndim (emcee red-blue move)
        m.fit()
    m = bb.Dias2000(fp, nwalkers=32, nsteps=10)
    with pytest.raises(ValueError):               # wrong p0 shape
        m.fit(p0=np.zeros((31, 5)))
    with pytest.raises(ValueError):               # linearly dependent walkers
        m.fit(p0=np.ones((32, 5)))
    with pytest.raises(NotImplementedError):
        m.fit(pool=object())
    assert not m.fitted
