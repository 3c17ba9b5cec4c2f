"""The CUDA sampler against REAL emcee output: the posterior numbers stored in the reference's tutorial notebooks
(docs/tutorials/{quickstart,decomposition,pelton,dias}.ipynb, quickstart_results.csv; extracted into
tests/golden/emcee_anchors.json by tests/golden/make_emcee_anchors.py).

For each notebook configuration the GPU sampler runs the same model / data file / bounds / walkers / steps with many
independent seeds (one launch: the seeds are batch entries), computes the statistics the notebook cell computed
(same discard / thin / flat), and every stored emcee number must lie within 3 combined standard errors of the seed
distribution (tests/anchors.py: measured seed scatter, not smaller than sigma / sqrt(ESS), plus the rounding of the
printed value).  The same is asserted through the drop-in class API (Inversion.fit + get_param_*).
"""
import numpy as np
import pytest

import anchors as A

pytestmark = pytest.mark.gpu
ANCHORS = A.load()
NSEEDS = 256        # centre and scatter of the seed distribution to ~6 %: the z-scores are then properties of the notebook run


def _batch_chains(anchor, pr, seeds):
    from bisip_b200.batch import BatchInversion
    n = len(seeds)
    d = pr['data']
    inv = BatchInversion(anchor['model'], d['w'], np.repeat(d['zn'][None], n, 0), np.repeat(d['zn_err'][None], n, 0),
                         nwalkers=anchor['nwalkers'], nsteps=anchor['nsteps'], bounds=pr['bounds'],
                         poly_deg=anchor.get('poly_deg', 5), c_exp=anchor.get('c_exp', 1.0),
                         n_modes=anchor.get('n_modes', 1), seed=0xE3CEE, spectrum_offset=seeds[0])
    p0 = np.stack([A.p0_for(anchor, pr['bounds'], s) for s in seeds])
    res = inv.fit(p0=p0, keep_chain=True)
    assert np.all(res['flags'] == 0)
    return res


_BATCH_RUNS = {}


def _batch_runs(name):
    """Statistics of NSEEDS independent GPU runs of an anchor's configuration (cached for the session)."""
    if name not in _BATCH_RUNS:
        anchor = ANCHORS[name]
        pr = A.problem(anchor)
        res = _batch_chains(anchor, pr, list(range(NSEEDS)))
        chains = res['chain']
        assert chains.shape == (NSEEDS, anchor['nsteps'], anchor['nwalkers'], len(anchor['param_names']))
        runs = [A.run_stats(anchor, c, pr['log_taus']) for c in chains]
        _BATCH_RUNS[name] = (runs, A.autocorr_time(chains[:8], anchor), float(res['acceptance_fraction'].mean()))
    return _BATCH_RUNS[name]


@pytest.mark.parametrize("name", sorted(ANCHORS))
def test_gpu_sampler_matches_real_emcee_notebook_output(name):
    anchor = ANCHORS[name]
    runs, tau, acc = _batch_runs(name)
    z = A.zscores(anchor, runs, tau=tau)
    for key, v in z.items():
        assert np.all(np.abs(v) <= 3.0), (name, key, np.round(v, 2))
    # acceptance of the stretch move at these dimensions (emcee: 0.2 - 0.6 is healthy)
    assert 0.25 < acc < 0.7


@pytest.mark.parametrize("name,nseeds", [('quickstart_cc1', 10), ('dias_K389172', 10), ('pelton_cc2_K389174', 8),
                                         ('decomp_debye_p4_K389170', 8)])
def test_class_api_matches_real_emcee_notebook_output(name, nseeds, data_files):
    """The notebook's own lines, through the drop-in classes: Model(filepath, ...); params.update(...); fit();
    get_param_mean / std / percentile(discard=, thin=) — p0 and the Philox key drawn from NumPy's global generator
    like the reference does.  A handful of such runs must be draws of the distribution that the 256-seed batch (the
    one z-tested against the notebook above) samples: same data ingest, bounds edits, discard / thin and statistics."""
    import bisip_b200 as bb
    anchor = ANCHORS[name]
    pr = A.problem(anchor)
    ctor = {'colecole': lambda: bb.PeltonColeCole(data_files[anchor['file']], nwalkers=anchor['nwalkers'],
                                                  nsteps=anchor['nsteps'], headers=anchor['headers'],
                                                  n_modes=anchor.get('n_modes', 1)),
            'dias': lambda: bb.Dias2000(data_files[anchor['file']], nwalkers=anchor['nwalkers'], nsteps=anchor['nsteps']),
            'decomp': lambda: bb.PolynomialDecomposition(data_files[anchor['file']], nwalkers=anchor['nwalkers'],
                                                         nsteps=anchor['nsteps'], poly_deg=anchor.get('poly_deg'),
                                                         c_exp=anchor.get('c_exp'))}[anchor['model']]
    runs, chains = [], []
    for s in range(nseeds):
        np.random.seed(4200 + s)
        m = ctor()
        m.params.update(anchor['bounds_edits'])
        m.fit()
        kw = dict(discard=anchor['discard'], thin=anchor['thin'])
        st = {'mean': m.get_param_mean(**kw), 'std': m.get_param_std(**kw)}
        if 'pct' in anchor:
            st['pct'] = m.get_param_percentile(p=anchor['p'], **kw)
        if 'total_m' in anchor:
            st['total_m'] = np.array([m.get_total_chargeability(**kw)])
        runs.append(st)
        chains.append(m.get_chain())
    batch, tau, _ = _batch_runs(name)
    for key in runs[0]:
        a = np.array([r[key] for r in runs])
        b = np.array([r[key] for r in batch])
        centre = np.median(b, axis=0)
        scatter = np.maximum(1.4826 * np.median(np.abs(b - centre), axis=0),
                             A.ess_standard_errors(anchor, np.median([r['std'] for r in batch], axis=0), tau).get(key, 0.0))
        se = scatter * np.sqrt((np.pi / 2) * (1.0 / len(a) + 1.0 / len(b)))
        zz = (np.median(a, axis=0) - centre) / se
        assert np.all(np.abs(zz) <= 3.5), (name, key, np.round(zz, 2))


def test_package_test_run_executes(capsys):
    """bisip_b200.test_run — the reference's only test (tests/test_module.py:20-100), same three fits."""
    import bisip_b200 as bb
    np.random.seed(7)
    bb.test_run()
    out = capsys.readouterr().out
    assert 'Testing ColeCole model' in out and 'Testing Dias model' in out and 'Testing Debye Decomposition' in out
    assert 'All tests passed.' in out
