"""``test_run`` — the package smoke run, mirroring reference ``tests/test_module.py:20-100``:
fit Cole-Cole, Dias and Debye models on a bundled spectrum, print mean +/- std, exercise the
plot helpers when matplotlib is present."""
import warnings

import numpy as np


def test_run(dias=True, colecole=True, debye=True, nsteps_scale=1.0):
    from . import DataFiles, Dias2000, PeltonColeCole, PolynomialDecomposition
    fp = DataFiles()['SIP-K389175']
    model = None

    def report(m, values, errs):
        for n, v, u in zip(m.param_names, values, errs):
            print(f'{n}: {v:.5f} +/- {u:.5f}')

    if colecole:
        print('Testing ColeCole model')
        model = PeltonColeCole(fp, nwalkers=32, n_modes=2, nsteps=int(1000 * nsteps_scale))
        model.fit()
        d = int(800 * nsteps_scale)
        report(model, model.get_param_mean(discard=d), model.get_param_std(discard=d))
    if dias:
        print('Testing Dias model')
        model = Dias2000(fp, nwalkers=32, nsteps=int(2000 * nsteps_scale))
        start = np.tile([1.0, 0.25, -10, 5, 0.5], (32, 1))
        start += 1e-1 * start * (np.random.rand(*start.shape) - 1)
        model.fit(p0=start)
        chain = model.get_chain(discard=int(1000 * nsteps_scale), thin=1, flat=True)
        report(model, model.get_param_mean(chain), model.get_param_std(chain))
    if debye:
        print('Testing Debye Decomposition')
        model = PolynomialDecomposition(fp, nwalkers=32, poly_deg=4, nsteps=int(2000 * nsteps_scale))
        model.params.update(a0=[-2, 2])      # bounds edited in place before fit()
        model.fit()
        chain = model.get_chain(discard=int(1000 * nsteps_scale), thin=1, flat=True)
        report(model, model.get_param_mean(chain), model.get_param_std(chain))
    if model is None:
        return
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        warnings.warn('matplotlib was not found: plot helpers not exercised')
        print('All tests passed.')
        return
    print('Testing plotlib with last results')
    for fig in (model.plot_data(feature='phase'), model.plot_traces(), model.plot_histograms(chain),
                model.plot_fit(chain)):
        plt.close(fig)
    try:
        plt.close(model.plot_corner(chain))
    except ImportError:
        warnings.warn('The `corner` package was not found. Install it with `conda install corner`')
    print('All tests passed.')


test_run.__test__ = False   # not a pytest test: it needs a GPU and is run by tests/test_gpu_emcee_anchors.py::test_package_test_run_executes
