"""CPU tests that pin the oracle (test infrastructure) against the reference:
 * C restatement vs the golden vectors generated from the UNMODIFIED reference build
 * Philox4x32-10 known-answer vectors (Random123)
 * where oracle/_ref exists (this container): restatement vs the live reference
 * sampler restatement: Philox-stream C sampler vs MT19937 emcee restatement, MC-error level
"""
import numpy as np
import pytest

from helpers import CASES, lp_err, normwise, oracle_problem
from oracle import oracle, refload


@pytest.mark.parametrize("case", list(CASES))
def test_c_oracle_matches_reference_golden(case, gold_fl, gold_ld):
    prob = oracle_problem(case, gold_fl, gold_ld)
    th = gold_fl[f'{case}/theta']
    Z = prob.forward(th)
    # same libm calls in the same order as the Cython reference: forward agrees to the last bits
    assert normwise(Z, gold_fl[f'{case}/Z']).max() <= 4e-16
    lp = prob.log_probability(th)
    ref = gold_fl[f'{case}/lp']
    assert np.array_equal(np.isneginf(lp), np.isneginf(ref))
    assert lp_err(lp, ref).max() <= 1e-14          # NumPy sums pairwise, the C loop in order


def test_survey_known_answers_oracle(gold_fl, gold_ld):
    known = {'decomp_p4_debye': 384.579610803116, 'decomp_p4_warburg': -26.40360134097351,
             'colecole_k2': 119.69117931325292, 'dias': 113.72993198177019, 'shin': -325.39343206612887}
    for case, val in known.items():
        prob = oracle_problem(case, gold_fl, gold_ld)
        assert abs(prob.log_probability(gold_fl[f'{case}/theta'][0]) - val) <= 2e-13 * abs(val)
    assert gold_ld['SIP-K389175/norm_factor'] == 41229.19000000001
    np.testing.assert_allclose(gold_ld['SIP-K389175/w'][:3], [37699.11184307752, 18849.55592153876, 9424.77796076938], rtol=0, atol=0)


def test_synthetic_golden_oracle(gold_fl):
    from bisip_b200 import synthetic
    _, w = synthetic.frequencies(64)
    for tag, model in (('syn_decomp_s64', 'decomp'), ('syn_decomp_s128', 'decomp'), ('syn_decomp_s256', 'decomp'),
                       ('syn_colecole', 'colecole'), ('syn_dias', 'dias'), ('syn_shin', 'shin')):
        for b in range(4):
            prob = oracle.Problem(model, w, gold_fl[f'{tag}/zn'][b], gold_fl[f'{tag}/zn_err'][b], gold_fl[f'{tag}/bounds'],
                                  taus=gold_fl.get(f'{tag}/taus'), log_taus=gold_fl.get(f'{tag}/log_taus'),
                                  c_exp=float(gold_fl.get(f'{tag}/c_exp', 1.0)))
            assert lp_err(prob.log_probability(gold_fl[f'{tag}/theta']), gold_fl[f'{tag}/lp'][b]).max() <= 1e-13
        assert normwise(prob.forward(gold_fl[f'{tag}/theta']), gold_fl[f'{tag}/Z']).max() <= 1e-15


def test_philox_known_answers():
    assert oracle.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert oracle.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert oracle.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


@pytest.mark.skipif(not refload.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_c_oracle_matches_live_reference(gold_fl, gold_ld):
    bisip = refload.load()
    fp = refload.data_file()
    rng = np.random.default_rng(3)
    for case, ctor in (('decomp_p4_debye', lambda: bisip.PolynomialDecomposition(fp, poly_deg=4)),
                       ('colecole_k2', lambda: bisip.PeltonColeCole(fp, n_modes=2)),
                       ('dias', lambda: bisip.Dias2000(fp)), ('shin', lambda: bisip.Shin2015(fp))):
        m = ctor()
        prob = oracle_problem(case, gold_fl, gold_ld)
        B = m.param_bounds
        for t in rng.uniform(B[0], B[1], (50, B.shape[1])):
            assert normwise(prob.forward(t)[None], m.forward(t, m.data['w'])[None]).max() <= 4e-16
            ref = m._log_probability(t, m.forward, B, m.data['w'], m.data['zn'], m.data['zn_err'])
            assert lp_err(prob.log_probability(t), ref) <= 1e-14


@pytest.mark.skipif(not refload.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_c_oracle_matches_live_reference_shape_sweep():
    """The restatement against the live reference over the model-size axes the goldens do not span: all six bundled data
    files, polynomial degrees 0-9 with Debye / Warburg / fractional exponents, 1-5 Cole-Cole modes, and thetas on the
    faces of the box and outside it (strict prior: -inf without a forward call, reference models.py:64-76)."""
    bisip = refload.load()
    rng = np.random.default_rng(11)
    files = ['SIP-K389170', 'SIP-K389172', 'SIP-K389173', 'SIP-K389174', 'SIP-K389175', 'SIP-K389176']
    ctors = [('decomp', dict(poly_deg=d, c_exp=c)) for d, c in ((0, 1.0), (1, 0.5), (2, 1.0), (3, 0.7), (5, 0.5), (7, 1.0), (9, 1.0))]
    ctors += [('colecole', dict(n_modes=k)) for k in (1, 2, 3, 4, 5)] + [('dias', {}), ('shin', {})]
    cls = {'decomp': bisip.PolynomialDecomposition, 'colecole': bisip.PeltonColeCole, 'dias': bisip.Dias2000,
           'shin': bisip.Shin2015}
    for i, (model, kw) in enumerate(ctors):
        m = cls[model](refload.data_file(files[i % len(files)]), **kw)
        B = m.param_bounds
        prob = oracle.Problem(model, m.data['w'], m.data['zn'], m.data['zn_err'], B, n_modes=kw.get('n_modes', 1),
                              taus=getattr(m, 'taus', None), log_taus=getattr(m, 'log_taus', None), c_exp=kw.get('c_exp', 1.0))
        th = rng.uniform(B[0], B[1], (12, B.shape[1]))
        if model == 'decomp':            # coefficients of realistic size too (the full box gives |Z| ~ 1e3 and more)
            th[6:, 1:] *= 0.02
        th[0, 0] = B[0, 0]               # on the lower face
        th[1, -1] = B[1, -1]             # on the upper face
        th[2, 0] = B[1, 0] + 0.5         # outside
        for t in th:
            Zr = m.forward(t, m.data['w'])
            assert normwise(prob.forward(t)[None], Zr[None]).max() <= 1e-15, (model, kw)
            ref = m._log_probability(t, m.forward, B, m.data['w'], m.data['zn'], m.data['zn_err'])
            assert lp_err(prob.log_probability(t), ref) <= 1e-13, (model, kw)
        assert np.isneginf([prob.log_probability(t) for t in th[:3]]).all()


def test_oracle_sampler_semantics(gold_fl, gold_ld):
    """Storage/slicing of the C sampler and invariants of the stretch move."""
    prob = oracle_problem('colecole_k1', gold_fl, gold_ld)
    rng = np.random.default_rng(5)
    lo, hi = gold_fl['colecole_k1/bounds']
    p0 = rng.uniform(lo, hi, (16, 4))
    full = prob.run(p0, 120, seed=9)
    assert full['chain'].shape == (120, 16, 4) and full['log_prob'].shape == (120, 16)
    part = prob.run(p0, 120, seed=9, discard=20, thin=7)
    np.testing.assert_array_equal(part['chain'], full['chain'][26::7])
    # positions only ever change to in-bounds proposals; stored lp matches a recomputation
    ch = full['chain'].reshape(-1, 4)
    assert np.all((ch > lo) & (ch < hi) | np.isin(ch, p0))
    lp = prob.log_probability(ch)
    np.testing.assert_allclose(lp, full['log_prob'].reshape(-1), rtol=0, atol=0)
    assert 0 < full['accepted'].sum() < 120 * 16
    # different spectrum index / seed => different stream
    assert not np.array_equal(prob.run(p0, 10, seed=9, spectrum=1)['chain'], full['chain'][:10])


def test_oracle_sampler_vs_emcee_restatement_mc(gold_fl, gold_ld, gold_post):
    """The Philox-stream sampler and the MT19937 emcee restatement (which produced posterior.npz
    by driving the reference fit()) sample the same posterior: means agree within MC error."""
    prob = oracle_problem('decomp_p4_debye', gold_fl, gold_ld)
    lo, hi = gold_fl['decomp_p4_debye/bounds']
    means, accs = [], []
    for s in range(4):
        p0 = np.random.default_rng(s).uniform(lo, hi, (32, 6))
        r = prob.run(p0, 1000, seed=100 + s)
        means.append(r['chain'][500:].reshape(-1, 6).mean(0))
        accs.append(r['accepted'].mean() / 1000)
    means = np.array(means)
    rm = gold_post['c1_decomp/mean']
    sd = gold_post['c1_decomp/std'].mean(0)
    se = np.sqrt(means.var(0, ddof=1) / 4 + rm.var(0, ddof=1) / len(rm))
    assert np.all(np.abs(means.mean(0) - rm.mean(0)) <= 5 * se + 0.05 * sd)
    assert abs(np.mean(accs) - gold_post['c1_decomp/acc'].mean()) <= 0.03


# ---- the sampler restatement against REAL emcee output (the reference's stored notebook cells) -------------------
@pytest.mark.parametrize("name,nseeds", [('quickstart_cc1', 24), ('dias_K389172', 24), ('pelton_cc2_K389174', 12),
                                         ('decomp_debye_p4_K389175', 6)])
def test_sampler_restatement_matches_real_emcee_notebook_output(name, nseeds):
    """oracle/bisip_oracle.c's stretch-move sampler (the stream the CUDA kernel reproduces bit for bit) run on the
    notebook's exact configuration: every number the notebook printed (posterior mean / std / percentiles / total
    chargeability, produced by the real emcee package) lies within 3 combined standard errors of the sampler's own
    seed distribution (tests/anchors.py).  This is what pins the restatement to emcee rather than to itself."""
    import anchors as A
    anchor = A.load()[name]
    pr = A.problem(anchor)
    kw = dict(n_modes=anchor.get('n_modes', 1))
    if anchor['model'] == 'decomp':
        kw.update(taus=pr['taus'], log_taus=pr['log_taus'], c_exp=anchor['c_exp'])
    prob = oracle.Problem(anchor['model'], pr['data']['w'], pr['data']['zn'], pr['data']['zn_err'], pr['bounds'], **kw)
    chains = [prob.run(A.p0_for(anchor, pr['bounds'], s), anchor['nsteps'], seed=0xE3CEE, spectrum=s)['chain']
              for s in range(nseeds)]
    runs = [A.run_stats(anchor, c, pr['log_taus']) for c in chains]
    z = A.zscores(anchor, runs, tau=A.autocorr_time(chains[:4], anchor))
    for key, v in z.items():
        assert np.all(np.abs(v) <= 3.0), (name, key, np.round(v, 2))


def test_emcee_anchor_fixture_matches_the_notebooks():
    """The committed fixture is what make_emcee_anchors.py extracts from the reference tree (where it exists)."""
    import json
    import os
    import subprocess
    import sys
    if not os.path.isdir('/root/reference/docs/tutorials'):
        pytest.skip('reference tree not present')
    import anchors as A
    before = json.load(open(A.ANCHOR_FILE))
    gen = os.path.join(os.path.dirname(A.ANCHOR_FILE), 'make_emcee_anchors.py')
    subprocess.check_call([sys.executable, gen], stdout=subprocess.DEVNULL)
    assert json.load(open(A.ANCHOR_FILE)) == before
    names = [a['name'] for a in before['anchors']]
    assert len(names) == 9 and 'quickstart_cc1' in names
