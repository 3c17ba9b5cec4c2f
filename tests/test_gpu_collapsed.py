"""GPU parity tests of precision='fp64-collapsed' (csrc/decomp_collapsed.cuh): the decomposition re-associated
to z = (L K) a with G = L K built once per spectrum.  It is an FP64 path and is held to the SAME bar as the
two-stage DMMA path (north_star: forward and log-prob within 1e-12 relative), against the golden vectors
generated from the reference (Decomp_cyth, cython_funcs.pyx:75-94; _log_probability, models.py:59-76) and
against the C oracle; the sampler must reproduce the oracle's chain on the same Philox stream.
"""
import numpy as np
import pytest

from helpers import CASES, lp_err, make_model, normwise, oracle_problem

pytestmark = pytest.mark.gpu

TOL = 1e-12
PREC = 'fp64-collapsed'
DECOMP_CASES = [c for c in CASES if c.startswith('decomp')]


@pytest.mark.parametrize("case", DECOMP_CASES)
def test_collapsed_forward_and_logprob_match_reference_golden(case, gold_fl, data_files):
    m = make_model(case, data_files['SIP-K389175'], precision=PREC)
    th = gold_fl[f'{case}/theta']
    Z = m.forward(th, m.data['w'])
    assert Z.shape == gold_fl[f'{case}/Z'].shape
    assert normwise(Z, gold_fl[f'{case}/Z']).max() <= TOL
    ref = gold_fl[f'{case}/lp']
    lp = m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
    assert np.array_equal(np.isneginf(lp), np.isneginf(ref))
    assert lp_err(lp, ref).max() <= TOL
    ll = m._log_likelihood(th, m.forward, m.data['w'], m.data['zn'], m.data['zn_err'])
    fin = np.isfinite(ref)
    assert lp_err(ll[fin], ref[fin]).max() <= TOL


@pytest.mark.parametrize("case", DECOMP_CASES)
def test_collapsed_matches_oracle_on_random_theta(case, gold_fl, gold_ld, data_files):
    prob = oracle_problem(case, gold_fl, gold_ld)
    m = make_model(case, data_files['SIP-K389175'], precision=PREC)
    rng = np.random.default_rng(7)
    lo, hi = gold_fl[f'{case}/bounds']
    th = rng.uniform(lo, hi, (301, lo.shape[0]))          # 301: a ragged last 128-row chunk
    assert normwise(m.forward(th, m.data['w']), prob.forward(th)).max() <= TOL
    lp = m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
    assert lp_err(lp, prob.log_probability(th)).max() <= TOL


@pytest.mark.parametrize("tag", ['syn_decomp_s64', 'syn_decomp_s128', 'syn_decomp_s256'])
def test_collapsed_synthetic_batch(tag, gold_fl):
    """Bench shape (N=64; 64, 128 and 256 taus — no cluster and no tau limit in this form) through the C ABI, and
    really the collapsed kernel: the library reports it, and the values differ from the two-stage path by rounding."""
    from bisip_b200 import _lib, engine, synthetic
    from bisip_b200.batch import BatchInversion
    _, w = synthetic.frequencies(64)
    kw = dict(poly_deg=4, n_tau=int(tag.split('_s')[1]), c_exp=float(gold_fl[f'{tag}/c_exp']))
    inv = {p: BatchInversion('decomp', w, gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err'], precision=p, **kw)
           for p in ('fp64', PREC)}
    dev = inv[PREC].device
    assert engine.decomp_kernel_kind(inv[PREC]._spec(), 64, 256) == 'fp64-collapsed'
    assert engine.decomp_kernel_kind(inv['fp64']._spec(), 64, 256) in ('dmma', 'dmma-cluster')
    th = gold_fl[f'{tag}/theta']
    thb = _lib.dev_f64(np.broadcast_to(th, (4,) + th.shape).copy(), dev)
    wd = _lib.dev_f64(w, dev)
    Z = {p: engine.forward(inv[p]._spec(), thb, wd).cpu().numpy() for p in inv}
    for b in range(4):
        assert normwise(Z[PREC][b], gold_fl[f'{tag}/Z']).max() <= TOL
    args = (thb, wd, _lib.dev_f64(gold_fl[f'{tag}/zn'], dev), _lib.dev_f64(gold_fl[f'{tag}/zn_err'], dev),
            _lib.dev_f64(gold_fl[f'{tag}/bounds'], dev))
    lp = {p: engine.log_probability(inv[p]._spec(), *args).cpu().numpy() for p in inv}
    ref = gold_fl[f'{tag}/lp']
    assert np.array_equal(np.isneginf(lp[PREC]), np.isneginf(ref))
    assert lp_err(lp[PREC], ref).max() <= TOL
    assert not np.array_equal(Z[PREC], Z['fp64'])          # a different summation order, not the same kernel


@pytest.mark.parametrize("case,W,T", [('decomp_p4_debye', 32, 200), ('decomp_p4_warburg', 64, 100),
                                      ('decomp_p5_debye', 256, 60), ('decomp_p3_c07', 33, 120),
                                      ('decomp_p4_debye', 300, 30), ('decomp_p4_debye', 130, 40)])
def test_collapsed_chain_matches_oracle_same_stream(case, W, T, gold_fl, gold_ld, data_files):
    """Same p0 and Philox stream as the oracle => same accept decisions => identical chain (positions come from
    the same unfused proposal arithmetic; a log-prob rounding difference of ~1e-14 flips a decision with
    probability ~1e-10 per run).  Walker counts cover the 128-thread CTA (<= 128 walkers), the 256-thread CTA,
    odd counts and more walkers than threads."""
    seed = 4321
    m = make_model(case, data_files['SIP-K389175'], nwalkers=W, nsteps=T, seed=seed, precision=PREC)
    rng = np.random.default_rng(seed)
    lo, hi = gold_fl[f'{case}/bounds']
    p0 = rng.uniform(lo, hi, (W, lo.shape[0]))
    m.fit(p0=p0)
    ref = oracle_problem(case, gold_fl, gold_ld).run(p0, T, seed=seed, spectrum=0)
    chain = m.get_chain()
    assert chain.shape == (T, W, p0.shape[1])
    np.testing.assert_array_equal(chain, ref['chain'])
    np.testing.assert_array_equal(m.sampler.accepted, ref['accepted'])
    lp = m.sampler.get_log_prob()
    fin = np.isfinite(ref['log_prob'])
    assert np.array_equal(fin, np.isfinite(lp))
    assert np.max(np.abs(lp[fin] - ref['log_prob'][fin]) / np.maximum(1, np.abs(ref['log_prob'][fin]))) <= TOL


def test_collapsed_batch_equals_two_stage_batch():
    """A batch of synthetic spectra at the bench shape: the collapsed and the two-stage DMMA sampler take the same
    decisions from the same p0 / seed, so kept chains, acceptance and summaries are identical."""
    from bisip_b200 import _lib, engine, synthetic
    from bisip_b200.batch import BatchInversion
    dev = _lib.require_cuda()
    N, S, P, W, B = 64, 64, 4, 256, 24
    _, w = synthetic.frequencies(N)
    kw = dict(poly_deg=P, n_tau=S, c_exp=1.0)
    probe = BatchInversion('decomp', w, np.zeros((1, 2, N)), np.ones((1, 2, N)), device=dev, **kw)
    fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
    syn = synthetic.make('decomp', 0, B, fwd, N=N, poly_deg=P, n_tau=S)
    inv = {p: BatchInversion('decomp', w, syn['zn'], syn['zn_err'], nwalkers=W, nsteps=120, seed=5, device=dev,
                             precision=p, **kw) for p in ('fp64', PREC)}
    p0 = inv['fp64'].draw_p0(0, B)
    res = {p: inv[p].fit(p0=p0.copy(), discard=60, thin=2, keep_chain=True) for p in inv}
    assert np.all(res[PREC]['flags'] == 0)
    np.testing.assert_array_equal(res[PREC]['chain'], res['fp64']['chain'])
    np.testing.assert_array_equal(res[PREC]['acceptance_fraction'], res['fp64']['acceptance_fraction'])
    np.testing.assert_array_equal(res[PREC]['percentiles'], res['fp64']['percentiles'])
