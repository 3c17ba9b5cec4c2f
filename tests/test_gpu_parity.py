"""GPU parity tests: CUDA path (through the C ABI) vs the golden vectors generated from the
reference and vs the C oracle, on identical theta and data.

Tolerances (north_star: forward and log-prob within 1e-12 relative in FP64 mode):
  forward  : max|dZ| / max|Z| per theta           <= 1e-12
  log-prob : |dlp| / max(1,|lp|)                   <= 1e-12 ; -inf must match exactly
"""
import numpy as np
import pytest

from helpers import CASES, lp_err, make_model, normwise, oracle_problem

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.mark.parametrize("case", list(CASES))
def test_forward_matches_reference_golden(case, gold_fl, data_files):
    m = make_model(case, data_files['SIP-K389175'])
    th = gold_fl[f'{case}/theta']
    Z = m.forward(th, m.data['w'])
    assert Z.shape == gold_fl[f'{case}/Z'].shape
    assert normwise(Z, gold_fl[f'{case}/Z']).max() <= TOL
    z1 = m.forward(th[0], m.data['w'])           # single theta -> (2, N), like the reference
    assert z1.shape == (2, m.data['N'])
    np.testing.assert_array_equal(z1, Z[0])


@pytest.mark.parametrize("case", list(CASES))
def test_log_probability_matches_reference_golden(case, gold_fl, data_files):
    m = make_model(case, data_files['SIP-K389175'])
    th = gold_fl[f'{case}/theta']
    ref = gold_fl[f'{case}/lp']
    lp = m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
    assert np.array_equal(np.isneginf(lp), np.isneginf(ref))
    assert np.isneginf(ref).sum() >= 8          # outside + on-face thetas are in the fixture
    assert lp_err(lp, ref).max() <= TOL
    # scalar call returns a float, log-likelihood ignores the prior
    assert isinstance(m._log_probability(th[0], m.forward, m.param_bounds, m.data['w'], m.data['zn'],
                                         m.data['zn_err']), float)
    ll = m._log_likelihood(th, m.forward, m.data['w'], m.data['zn'], m.data['zn_err'])
    fin = np.isfinite(ref)
    assert lp_err(ll[fin], ref[fin]).max() <= TOL
    assert np.all(np.isfinite(ll))


@pytest.mark.parametrize("case", list(CASES))
def test_forward_and_logprob_match_oracle(case, gold_fl, gold_ld, data_files):
    prob = oracle_problem(case, gold_fl, gold_ld)
    m = make_model(case, data_files['SIP-K389175'])
    rng = np.random.default_rng(7)
    lo, hi = gold_fl[f'{case}/bounds']
    th = rng.uniform(lo, hi, (300, lo.shape[0]))
    assert normwise(m.forward(th, m.data['w']), prob.forward(th)).max() <= TOL
    lp = m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
    assert lp_err(lp, prob.log_probability(th)).max() <= TOL


def test_drop_in_classes_match_live_reference_shape_sweep(data_files):
    """The drop-in classes against the LIVE unmodified reference (oracle/_ref travels to the GPU box) on the same theta
    and data over the axes the goldens do not span: all six bundled files, polynomial degrees 0-9 with Debye / Warburg /
    fractional exponents, 1-5 Cole-Cole modes, thetas on the faces of the box and outside it.  Same constructor
    arguments, same attribute names, same call signatures on both sides."""
    from oracle import refload
    if not refload.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import bisip_b200 as bb
    ref = refload.load()
    rng = np.random.default_rng(11)
    files = ['SIP-K389170', 'SIP-K389172', 'SIP-K389173', 'SIP-K389174', 'SIP-K389175', 'SIP-K389176']
    ctors = [('PolynomialDecomposition', dict(poly_deg=d, c_exp=c))
             for d, c in ((0, 1.0), (1, 0.5), (2, 1.0), (3, 0.7), (5, 0.5), (7, 1.0), (9, 1.0))]
    ctors += [('PeltonColeCole', dict(n_modes=k)) for k in (1, 2, 3, 4, 5)] + [('Dias2000', {}), ('Shin2015', {})]
    for i, (cls, kw) in enumerate(ctors):
        name = files[i % len(files)]
        r = getattr(ref, cls)(refload.data_file(name), **kw)
        m = getattr(bb, cls)(data_files[name], **kw)
        assert m.param_names == r.param_names
        np.testing.assert_array_equal(m.param_bounds, r.param_bounds)
        for key in ('w', 'zn', 'zn_err'):
            np.testing.assert_array_equal(m.data[key], r.data[key])
        B = r.param_bounds
        th = rng.uniform(B[0], B[1], (40, B.shape[1]))
        if cls == 'PolynomialDecomposition':
            th[20:, 1:] *= 0.02           # coefficients of realistic size as well as the full box
        th[0, 0] = B[0, 0]                # on the lower face
        th[1, -1] = B[1, -1]              # on the upper face
        th[2, 0] = B[1, 0] + 0.5          # outside
        Zr = np.stack([r.forward(t, r.data['w']) for t in th])
        lr = np.array([r._log_probability(t, r.forward, B, r.data['w'], r.data['zn'], r.data['zn_err']) for t in th])
        assert normwise(m.forward(th, m.data['w']), Zr).max() <= TOL, (cls, kw)
        lp = m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
        assert np.array_equal(np.isneginf(lp), np.isneginf(lr)) and np.isneginf(lr[:3]).all()
        assert lp_err(lp, lr).max() <= TOL, (cls, kw)


def test_forward_on_box_faces_matches_live_reference():
    """One parameter at a time (and pairs: corners) on a face of the default box, all four models: the reference's C complex
    arithmetic is finite there (1/R = inf, 1/delta = inf, 1/(1-m) = inf ...) and the CUDA forward must return the same
    values, not NaN — forward() and the prior-free _log_likelihood() (342 points; the sweep is tools/face_sweep.py)."""
    import importlib.util
    import os
    from oracle import refload
    if not refload.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('face_sweep', os.path.join(root, 'tools', 'face_sweep.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    bad, worst, npts = mod.sweep(TOL)
    assert npts == 342 and not bad, bad[:3]
    assert worst <= TOL


def test_survey_known_answers(gold_fl, data_files):
    """SURVEY.md App. C.1 log-probabilities."""
    known = {'decomp_p4_debye': 384.579610803116, 'decomp_p4_warburg': -26.40360134097351,
             'colecole_k2': 119.69117931325292, 'dias': 113.72993198177019, 'shin': -325.39343206612887}
    for case, val in known.items():
        m = make_model(case, data_files['SIP-K389175'])
        th = gold_fl[f'{case}/theta'][0]
        lp = m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
        assert abs(lp - val) <= 1e-12 * max(1, abs(val))
    m = make_model('decomp_p4_debye', data_files['SIP-K389175'])
    th = gold_fl['decomp_p4_debye/theta'][0].copy()
    th[0] = 1.1                                   # exactly on the upper bound -> strict prior
    assert m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'],
                              m.data['zn_err']) == -np.inf


def test_raw_kernels_other_frequencies(gold_fl, data_files):
    """forward(theta, w) must honour the w argument (SURVEY App. C.1 raw kernel calls)."""
    import bisip_b200 as bb
    fp = data_files['SIP-K389175']
    w = gold_fl['raw/w']
    cc = bb.PeltonColeCole(fp, n_modes=1).forward(np.array([1.0, 0.3, -2.0, 0.5]), w)
    di = bb.Dias2000(fp).forward(np.array([1.0, 0.25, -10.0, 5.0, 0.5]), w)
    sh = bb.Shin2015(fp).forward(np.array([0.5, 0.5, -14.0, -6.0, 0.5, 0.5]), w)
    for got, key in ((cc, 'colecole'), (di, 'dias'), (sh, 'shin')):
        assert normwise(got[None], gold_fl[f'raw/{key}'][None]).max() <= TOL


SYN = ['syn_decomp_s64', 'syn_decomp_s128', 'syn_decomp_s256', 'syn_colecole', 'syn_dias', 'syn_shin']


@pytest.mark.parametrize("tag", SYN)
def test_synthetic_batch_logprob(tag, gold_fl):
    """Batched C-ABI call over 4 spectra x n theta at the bench shape (N=64)."""
    import torch
    from bisip_b200 import _lib, engine, synthetic
    from bisip_b200.batch import BatchInversion
    model = tag.split('_')[1]
    _, w = synthetic.frequencies(64)
    kw = dict(poly_deg=4, n_tau=int(tag.split('_s')[1]), c_exp=float(gold_fl[f'{tag}/c_exp'])) if model == 'decomp' else {}
    inv = BatchInversion(model, w, gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err'], **kw)
    dev = inv.device
    th = gold_fl[f'{tag}/theta']
    thb = _lib.dev_f64(np.broadcast_to(th, (4,) + th.shape).copy(), dev)
    spec = inv._spec()
    Z = engine.forward(spec, thb, _lib.dev_f64(w, dev)).cpu().numpy()
    for b in range(4):
        assert normwise(Z[b], gold_fl[f'{tag}/Z']).max() <= TOL
    lp = engine.log_probability(spec, thb, _lib.dev_f64(w, dev), _lib.dev_f64(gold_fl[f'{tag}/zn'], dev),
                                _lib.dev_f64(gold_fl[f'{tag}/zn_err'], dev),
                                _lib.dev_f64(gold_fl[f'{tag}/bounds'], dev)).cpu().numpy()
    ref = gold_fl[f'{tag}/lp']
    assert np.array_equal(np.isneginf(lp), np.isneginf(ref))
    assert lp_err(lp, ref).max() <= TOL
    torch.cuda.synchronize()


def test_decomp_kernel_matrix(gold_fl, gold_ld):
    """K = 1 - 1/(1+(i w tau)^c) against NumPy complex arithmetic."""
    from bisip_b200 import _lib, engine
    dev = _lib.require_cuda()
    w = gold_ld['SIP-K389175/w']
    for case, c in (('decomp_p4_debye', 1.0), ('decomp_p4_warburg', 0.5), ('decomp_p3_c07', 0.7)):
        taus = gold_fl[f'{case}/taus']
        K = engine.decomp_kernel_matrix(_lib.dev_f64(w, dev), _lib.dev_f64(taus, dev), c).cpu().numpy()
        ref = 1 - 1 / (1 + (1j * w[None, :] * taus[:, None]) ** c)
        N = len(w)
        assert np.max(np.abs(K[:, :N] - ref.real)) <= 1e-14
        assert np.max(np.abs(K[:, N:] - ref.imag)) <= 1e-14


@pytest.mark.parametrize("tag", ['syn_decomp_s64', 'syn_decomp_s128', 'syn_decomp_s256'])
@pytest.mark.parametrize("prec,ztol,lptol", [('tf32', 1e-3, 3e-3), ('3xtf32', 2e-5, 5e-5),
                                             ('tf32-mma', 1e-3, 3e-3), ('3xtf32-mma', 2e-5, 5e-5)])
def test_reduced_precision_within_stated_tolerance(tag, prec, ztol, lptol, gold_fl):
    """TF32 / 3xTF32 stage 2 (FP64 stage 1, FP32 accumulate) against the reference golden values, on the tcgen05
    kernel ('tf32', '3xtf32' where the problem fits one tile) and on the mma.sync tiles ('*-mma', and the fallback).
    Stated tolerances (measured in profiles/r01_tf32_study.md, with head-room):
      forward  max|dZ|/max|Z| <= 1e-3 (tf32), 2e-5 (3xtf32);  log-prob relative <= 3e-3 / 5e-5."""
    from bisip_b200 import _lib, engine, synthetic
    from bisip_b200.batch import BatchInversion
    _, w = synthetic.frequencies(64)
    inv = BatchInversion('decomp', w, gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err'], poly_deg=4,
                         n_tau=int(tag.split('_s')[1]), c_exp=float(gold_fl[f'{tag}/c_exp']), precision=prec)
    dev = inv.device
    th = gold_fl[f'{tag}/theta']
    thb = _lib.dev_f64(np.broadcast_to(th, (4,) + th.shape).copy(), dev)
    Z = engine.forward(inv._spec(), thb, _lib.dev_f64(w, dev)).cpu().numpy()
    err = normwise(Z[0], gold_fl[f'{tag}/Z']).max()
    assert 1e-9 < err <= ztol          # > 1e-9: really the reduced-precision path, not FP64
    lp = engine.log_probability(inv._spec(), thb, _lib.dev_f64(w, dev), _lib.dev_f64(gold_fl[f'{tag}/zn'], dev),
                                _lib.dev_f64(gold_fl[f'{tag}/zn_err'], dev), _lib.dev_f64(gold_fl[f'{tag}/bounds'], dev)).cpu().numpy()
    ref = gold_fl[f'{tag}/lp']
    assert np.array_equal(np.isneginf(lp), np.isneginf(ref))
    assert lp_err(lp, ref).max() <= lptol


def test_3xtf32_sampler_runs_and_agrees_statistically(gold_fl):
    from bisip_b200 import synthetic
    from bisip_b200.batch import BatchInversion
    _, w = synthetic.frequencies(64)
    tag = 'syn_decomp_s64'
    out = {}
    for prec in ('fp64', '3xtf32'):
        inv = BatchInversion('decomp', w, gold_fl[f'{tag}/zn'], gold_fl[f'{tag}/zn_err'], nwalkers=64, nsteps=1500,
                             poly_deg=4, n_tau=64, precision=prec, seed=3)
        out[prec] = inv.fit(discard=750, thin=3)
        assert np.all(out[prec]['flags'] == 0)
    shift = np.abs(out['3xtf32']['percentiles'][:, 1] - out['fp64']['percentiles'][:, 1]) / out['fp64']['std']
    assert shift.max() < 0.5             # medians agree within Monte-Carlo error
    assert np.abs(out['3xtf32']['acceptance_fraction'] - out['fp64']['acceptance_fraction']).max() < 0.03


TCGEN05_SHAPES = [  # (n_freq, n_tau, poly_deg, walkers, c_exp, precision, expected kernel)
    (64, 64, 4, 256, 1.0, '3xtf32', 'tcgen05'),      # C5 shape: one M=128 tile per half-step, two CTAs per SM
    (64, 64, 4, 256, 1.0, 'tf32', 'tcgen05'),
    (20, 40, 4, 32, 1.0, '3xtf32', 'tcgen05'),       # bundled-file shape (C1): 16-row half-steps, padded columns / taus
    (33, 50, 3, 66, 0.5, '3xtf32', 'tcgen05'),       # odd everything, Warburg
    (64, 8, 0, 14, 1.0, '3xtf32', 'tcgen05'),        # a single K step, poly_deg 0
    (5, 3, 1, 8, 1.0, '3xtf32', 'tcgen05'),          # fewer taus than a K step, fewer frequencies than a column group
    (9, 4, 5, 16, 1.0, 'tf32', 'tcgen05'),           # more coefficients than taus: rank-deficient tau table
    (64, 64, 7, 254, 1.0, '3xtf32', 'tcgen05'),      # poly_deg 7 (8 coefficients), 127-row half-steps
    (64, 128, 4, 256, 1.0, '3xtf32', 'tcgen05'),     # two 64-tau chunks: double-buffered A, one CTA per SM
    (64, 256, 4, 128, 1.0, 'tf32', 'tcgen05'),       # four chunks (C4 tau grid)
    (64, 256, 4, 128, 1.0, '3xtf32', 'tcgen05-cluster'),   # K planes exceed one CTA: real | imaginary columns over a 2-CTA cluster
    (64, 256, 4, 256, 0.5, '3xtf32', 'tcgen05-cluster'),   # C4 shape, Warburg
    (40, 256, 5, 100, 1.0, '3xtf32', 'tcgen05-cluster'),   # 48 columns per CTA: the thread halves do not split them
    (64, 512, 4, 64, 1.0, '3xtf32', 'mma-tf32'),     # too large even for the pair -> mma.sync tiles
    (64, 64, 4, 258, 1.0, '3xtf32', 'mma-tf32'),     # 129-row half-steps do not fit the 128-lane tile
    (96, 64, 4, 64, 1.0, '3xtf32', 'mma-tf32'),      # 192 columns > 128
    (64, 64, 4, 256, 1.0, '3xtf32-mma', 'mma-tf32'),
    (64, 64, 4, 256, 1.0, 'fp64', 'dmma'),
    (64, 128, 4, 256, 1.0, 'fp64', 'dmma-cluster'),
]


@pytest.mark.parametrize("N,S,P,W,c_exp,prec,kind", TCGEN05_SHAPES)
def test_tcgen05_dispatch_and_agreement(N, S, P, W, c_exp, prec, kind):
    """Every shape class of the tcgen05 decomposition kernel (and each documented fallback): the library reports
    the kernel it will launch, its log-probability agrees with the FP64 DMMA path within the stated TF32 / 3xTF32
    tolerance, and a short sampler run from the same p0 / seed produces no NaN flag and the same acceptance."""
    from bisip_b200 import _lib, engine, synthetic
    from bisip_b200.batch import BatchInversion
    dev = _lib.require_cuda()
    _, w = synthetic.frequencies(N)
    kw = dict(poly_deg=P, n_tau=S, c_exp=c_exp)
    probe = BatchInversion('decomp', w, np.zeros((1, 2, N)), np.ones((1, 2, N)), device=dev, **kw)
    fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
    B = 5
    syn = synthetic.make('decomp', 0, B, fwd, N=N, poly_deg=P, n_tau=S)
    inv = {p: BatchInversion('decomp', w, syn['zn'], syn['zn_err'], nwalkers=W, nsteps=60, seed=11, device=dev, precision=p, **kw)
           for p in ('fp64', prec)}
    assert engine.decomp_kernel_kind(inv[prec]._spec(), N, W) == kind
    truth = syn['theta_true'].copy()
    truth[:, 0] /= syn['norm_factor']
    rng = np.random.default_rng(5)
    th = truth[:, None, :] * (1 + 2e-3 * rng.standard_normal((B, 333, P + 2)))
    th[:, -1, 0] = 2.0                                                      # one row outside the prior box
    args = (_lib.dev_f64(th, dev), _lib.dev_f64(w, dev), _lib.dev_f64(syn['zn'], dev), _lib.dev_f64(syn['zn_err'], dev),
            _lib.dev_f64(inv['fp64'].param_bounds, dev))
    lp = {p: engine.log_probability(inv[p]._spec(), *args).cpu().numpy() for p in inv}
    assert np.all(np.isneginf(lp[prec][:, -1]))
    fin = np.isfinite(lp['fp64'])                    # tiny tau grids can put a synthetic truth outside the prior box
    assert np.array_equal(np.isfinite(lp[prec]), fin) and not np.any(np.isnan(lp[prec])) and fin.sum() >= 300
    tol = 0.0 if prec == 'fp64' else (3e-3 if prec.startswith('tf32') else 5e-5)
    assert lp_err(lp[prec][fin], lp['fp64'][fin]).max() <= tol
    Z = {p: engine.forward(inv[p]._spec(), args[0], args[1]).cpu().numpy() for p in inv}
    ztol = 0.0 if prec == 'fp64' else (1e-3 if prec.startswith('tf32') else 2e-5)
    assert max(normwise(Z[prec][b], Z['fp64'][b]).max() for b in range(B)) <= ztol
    p0 = inv['fp64'].draw_p0(0, B)
    res = {p: inv[p].fit(p0=p0.copy(), discard=30, thin=1) for p in inv}
    assert np.all(res[prec]['flags'] == 0)
    assert np.abs(res[prec]['acceptance_fraction'] - res['fp64']['acceptance_fraction']).max() < 0.1


@pytest.mark.gpu
def test_user_forward_callable(data_files):
    """The reference's _log_likelihood / _log_probability take ANY forward callable (models.py:59-62, :71-76): the
    callable runs on the host, the Gaussian reduction in bisip_gauss_loglike; tolerance 1e-12 relative against the
    reference expression evaluated in NumPy on the same model rows."""
    import bisip_b200 as bb
    m = bb.PeltonColeCole(data_files['SIP-K389175'], nwalkers=32, nsteps=10, n_modes=1)
    w, y, yerr = m.data['w'], m.data['zn'], m.data['zn_err']
    calls = []

    def debye(theta, x):               # a model the library does not have
        calls.append(1)
        z = theta[0] * (1 - theta[1] * (1 - 1 / (1 + 1j * x * 10.0 ** theta[2])))
        return np.array([z.real, z.imag])

    def ref_ll(theta):
        s2 = yerr ** 2
        return -0.5 * np.sum((y - debye(theta, w)) ** 2 / s2 + 2 * np.log(s2))

    rng = np.random.default_rng(5)
    lo, hi = m.param_bounds
    th = rng.uniform(lo, hi, (7, 4))
    want = np.array([ref_ll(t) for t in th])
    got = m._log_likelihood(th, debye, w, y, yerr)
    assert got.shape == (7,)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    one = m._log_likelihood(th[0], debye, w, y, yerr)
    assert isinstance(one, float) and abs(one - want[0]) <= 1e-12 * abs(want[0])
    # prior first: outside (or on a face of) the box the callable is never run
    th[2, 1] = hi[1]
    th[5, 0] = lo[0] - 1.0
    calls.clear()
    lp = m._log_probability(th, debye, m.param_bounds, w, y, yerr)
    assert len(calls) == 5 and np.isneginf(lp[[2, 5]]).all()
    keep = [0, 1, 3, 4, 6]
    np.testing.assert_allclose(lp[keep], want[keep], rtol=1e-12)
    assert m._log_probability(th[2], debye, m.param_bounds, w, y, yerr) == -np.inf
    with pytest.raises(ValueError):
        m._log_likelihood(th[0], lambda t, x: np.zeros(3), w, y, yerr)
