#!/usr/bin/env python
"""Generate the golden fixtures from the UNMODIFIED reference (run where /root/reference
exists, after ``oracle/build_ref.sh``):

    python tests/golden/make_golden.py

Writes
  bisip_b200/data/examples.npz      raw tables of the six bundled example spectra
  tests/golden/load_data.npz        reference ``load_data`` outputs (utils.py:108-146)
  tests/golden/forward_logprob.npz  reference forward / _log_probability at seeded theta
                                    (models.py:59-76 + cython_funcs.pyx) on SIP-K389175 and on
                                    synthetic 64-frequency spectra
  tests/golden/posterior.npz        posterior summaries of the reference ``fit()`` driven by
                                    oracle/emcee_restatement.py (MC-error anchors)

Everything is seeded; re-running reproduces the files bit for bit on the same NumPy/glibc.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refload          # noqa: E402
from bisip_b200 import synthetic    # noqa: E402  (host-only helpers; no GPU touched)

GOLD = os.path.join(ROOT, "tests", "golden")
FILES = ["SIP-K389170", "SIP-K389172", "SIP-K389173", "SIP-K389174", "SIP-K389175", "SIP-K389176"]


def ref_lp(m, th):
    return m._log_probability(th, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])


def theta_set(bounds, rng, n_in=48):
    lo, hi = bounds
    ndim = lo.shape[0]
    inside = rng.uniform(lo, hi, (n_in, ndim))
    near = lo + (hi - lo) * rng.uniform(0.45, 0.55, (8, ndim))       # central, well conditioned
    outside = rng.uniform(lo, hi, (6, ndim))
    for i in range(6):
        d = i % ndim
        outside[i, d] = hi[d] + 0.1 * (hi[d] - lo[d]) if i % 2 else lo[d] - 0.1 * (hi[d] - lo[d])
    on_face = rng.uniform(lo, hi, (2, ndim))
    on_face[0, 0] = hi[0]                                            # strict bound -> -inf
    on_face[1, ndim - 1] = lo[ndim - 1]
    return np.concatenate([inside, near, outside, on_face])


def main():
    bisip = refload.load()
    rng = np.random.default_rng(20261017)

    # ---- raw tables + load_data ---------------------------------------------------------
    tables, ld = {}, {}
    for name in FILES:
        fp = refload.data_file(name)
        tables[name] = np.loadtxt(fp, skiprows=1, delimiter=',')
        m = bisip.Dias2000(fp)
        for k in ('zn', 'zn_err', 'w', 'Z', 'Z_err'):
            ld[f'{name}/{k}'] = np.asarray(m.data[k])
        ld[f'{name}/norm_factor'] = np.float64(m.data['norm_factor'])
    for units in ('rad', 'deg'):
        m = bisip.Dias2000(refload.data_file("SIP-K389175"), ph_units=units)
        ld[f'units_{units}/zn'] = m.data['zn']
        ld[f'units_{units}/zn_err'] = m.data['zn_err']
    m = bisip.Dias2000(refload.data_file("SIP-K389172"), headers=9)
    ld['headers9/zn'] = m.data['zn']
    ld['headers9/w'] = m.data['w']
    os.makedirs(os.path.join(ROOT, "bisip_b200", "data"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "bisip_b200", "data", "examples.npz"), **tables)
    np.savez_compressed(os.path.join(GOLD, "load_data.npz"), **ld)

    # ---- forward / log-prob on the bundled spectrum ---------------------------------------
    fp = refload.data_file("SIP-K389175")
    fl = {}
    cases = {
        'decomp_p4_debye': lambda: bisip.PolynomialDecomposition(fp, poly_deg=4, c_exp=1.0),
        'decomp_p4_warburg': lambda: bisip.PolynomialDecomposition(fp, poly_deg=4, c_exp=0.5),
        'decomp_p5_debye': lambda: bisip.PolynomialDecomposition(fp),
        'decomp_p3_c07': lambda: bisip.PolynomialDecomposition(fp, poly_deg=3, c_exp=0.7),
        'colecole_k1': lambda: bisip.PeltonColeCole(fp, n_modes=1),
        'colecole_k2': lambda: bisip.PeltonColeCole(fp, n_modes=2),
        'colecole_k3': lambda: bisip.PeltonColeCole(fp, n_modes=3),
        'dias': lambda: bisip.Dias2000(fp),
        'shin': lambda: bisip.Shin2015(fp),
    }
    survey_theta = {   # SURVEY.md App. C.1 known answers
        'decomp_p4_debye': [0.997613, 0.006870, -0.003937, -0.001338, 0.000741, 0.000219],
        'decomp_p4_warburg': [0.997613, 0.006870, -0.003937, -0.001338, 0.000741, 0.000219],
        'colecole_k2': [1.0, 0.14, 0.9, -1.5, -12.8, 0.45, 0.6],
        'dias': [1.0, 0.25, -10.0, 5.0, 0.5],
        'shin': [0.5, 0.5, -14.0, -6.0, 0.5, 0.5],
    }
    for name, ctor in cases.items():
        m = ctor()
        th = theta_set(m.param_bounds.astype(float), rng)
        if name in survey_theta:
            th = np.concatenate([np.array([survey_theta[name]]), th])
        fl[f'{name}/theta'] = th
        fl[f'{name}/bounds'] = m.param_bounds.astype(float)
        fl[f'{name}/Z'] = np.stack([m.forward(t, m.data['w']) for t in th])
        fl[f'{name}/lp'] = np.array([ref_lp(m, t) for t in th])
        if name.startswith('decomp'):
            fl[f'{name}/taus'] = m.taus
            fl[f'{name}/log_taus'] = m.log_taus
    # raw kernel calls of SURVEY App. C.1
    cy = sys.modules['bisip.cython_funcs']
    wk = 2 * np.pi * np.array([1e-2, 1.0, 1e2, 6e3])
    fl['raw/w'] = wk
    fl['raw/colecole'] = cy.ColeCole_cyth(wk, 1.0, np.array([0.3]), np.array([-2.0]), np.array([0.5]))
    fl['raw/dias'] = cy.Dias2000_cyth(wk, 1.0, 0.25, -10.0, 5.0, 0.5)
    fl['raw/shin'] = cy.Shin2015_cyth(wk, np.array([0.5, 0.5]), np.array([-14.0, -6.0]), np.array([0.5, 0.5]))

    # ---- synthetic 64-frequency spectra (bench shape), reference forward as truth ---------------
    N = 64
    _, w64 = synthetic.frequencies(N)
    m0 = bisip.Dias2000(fp)                     # any instance: we call the Cython kernels directly

    def ref_forward(model, poly_deg=4, n_tau=None, c_exp=1.0, n_modes=1):
        from bisip_b200.batch import tau_grid
        _, taus, log_taus = tau_grid(w64, n_tau, poly_deg)

        def f(theta, w):
            out = np.empty((len(theta), 2, len(w)))
            for i, t in enumerate(theta):
                if model == 'decomp':
                    out[i] = cy.Decomp_cyth(w, taus, log_taus, c_exp, R0=t[0], a=np.ascontiguousarray(t[1:]))
                elif model == 'colecole':
                    K = n_modes
                    out[i] = cy.ColeCole_cyth(w, R0=t[0], m=np.ascontiguousarray(t[1:1 + K]),
                                              lt=np.ascontiguousarray(t[1 + K:1 + 2 * K]),
                                              c=np.ascontiguousarray(t[1 + 2 * K:]))
                elif model == 'dias':
                    out[i] = cy.Dias2000_cyth(w, *t)
                else:
                    out[i] = cy.Shin2015_cyth(w, R=np.ascontiguousarray(t[:2]), log_Q=np.ascontiguousarray(t[2:4]),
                                              n=np.ascontiguousarray(t[4:]))
            return out
        return f, taus, log_taus

    from bisip_b200.batch import default_bounds
    syn_cases = [('decomp', dict(poly_deg=4, n_tau=64)), ('decomp', dict(poly_deg=4, n_tau=128)),
                 ('decomp', dict(poly_deg=4, n_tau=256, c_exp=0.5)), ('colecole', dict(n_modes=1)),
                 ('dias', {}), ('shin', {})]
    for model, kw in syn_cases:
        tag = f"syn_{model}" + (f"_s{kw['n_tau']}" if 'n_tau' in kw else '')
        f, taus, log_taus = ref_forward(model, **kw)
        syn = synthetic.make(model, 0, 4, f, N=N, poly_deg=kw.get('poly_deg', 4), n_modes=kw.get('n_modes', 1),
                             n_tau=kw.get('n_tau'))
        _, bounds = default_bounds(model, kw.get('poly_deg', 4), kw.get('n_modes', 1))
        fl[f'{tag}/zn'], fl[f'{tag}/zn_err'] = syn['zn'], syn['zn_err']
        fl[f'{tag}/theta_true'] = syn['theta_true']
        fl[f'{tag}/bounds'] = bounds
        if model == 'decomp':
            fl[f'{tag}/taus'], fl[f'{tag}/log_taus'] = taus, log_taus
            fl[f'{tag}/c_exp'] = np.float64(kw.get('c_exp', 1.0))
        th = np.concatenate([theta_set(bounds, rng, n_in=12)[:20],
                             syn['theta_true'] * (1 + 1e-3 * rng.standard_normal(syn['theta_true'].shape))])
        fl[f'{tag}/theta'] = th
        Zs = f(th, w64)
        fl[f'{tag}/Z'] = Zs
        lps = np.empty((4, len(th)))
        for b in range(4):
            for i, t in enumerate(th):
                lps[b, i] = m0._log_probability(t, lambda tt, ww: f(tt[None], ww)[0], bounds, w64,
                                                syn['zn'][b], syn['zn_err'][b])
        fl[f'{tag}/lp'] = lps
    np.savez_compressed(os.path.join(GOLD, "forward_logprob.npz"), **fl)

    # ---- posterior summaries from the reference fit() (emcee restatement drives it) --------
    post = {}

    def run(tag, ctor, setup, nseeds, discard, thin=1):
        means, stds, pcts, accs = [], [], [], []
        for s in range(nseeds):
            np.random.seed(42 + s)
            m = ctor()
            setup(m)
            m.fit()
            ch = m.get_chain(discard=discard, thin=thin, flat=True)
            means.append(ch.mean(0))
            stds.append(ch.std(0))
            pcts.append(np.percentile(ch, [2.5, 50, 97.5], axis=0))
            accs.append(m.sampler.acceptance_fraction.mean())
        post[f'{tag}/mean'] = np.array(means)
        post[f'{tag}/std'] = np.array(stds)
        post[f'{tag}/pct'] = np.array(pcts)
        post[f'{tag}/acc'] = np.array(accs)
        post[f'{tag}/bounds'] = m.param_bounds.astype(float)
        print(tag, np.mean(means, 0), np.mean(accs))

    f175, f174, f172 = (refload.data_file(n) for n in ("SIP-K389175", "SIP-K389174", "SIP-K389172"))
    run('c1_decomp', lambda: bisip.PolynomialDecomposition(f175, nwalkers=32, poly_deg=4, nsteps=1000),
        lambda m: None, nseeds=6, discard=500)
    run('c2_colecole', lambda: bisip.PeltonColeCole(f174, nwalkers=64, n_modes=2, nsteps=2000),
        lambda m: m.params.update({'log_tau1': [-5, 5], 'log_tau2': [-15, -10]}), nseeds=4, discard=1000)
    run('dias', lambda: bisip.Dias2000(f172, nwalkers=32, nsteps=2000),
        lambda m: m.params.update({'eta': [0, 25], 'log_tau': [-15, -5]}), nseeds=4, discard=1000)
    np.savez_compressed(os.path.join(GOLD, "posterior.npz"), **post)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
